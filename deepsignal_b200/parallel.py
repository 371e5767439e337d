"""Read sharding across the GPUs of one box: reads are independent, weights are
replicated, there is NO collective on the data path -- only a host-side gather of the result text that
keeps each read's lines contiguous and in input order, like the reference's single writer process
(call_modifications.py:283-297; its CPU fan-out is call_modifications.py:461-476).

Two ways to use several GPUs, both built on the same ownership rules:

* threads  -- ``run_sharded``: one worker thread + one ``dsb_ctx`` per device inside one host process
  (``deepsignal call_mods --devices 0,1,...``).  Independent contexts are thread-safe and ctypes releases
  the GIL during library calls, so the devices run concurrently; a re-order buffer in front of the writer
  restores input order.
* processes -- under ``torchrun`` (RANK / WORLD_SIZE set) every rank processes the batches it owns on
  ``cuda:LOCAL_RANK`` and rank 0 writes after ``gather_in_order`` (gloo / the CPU side of nccl).
"""
from __future__ import annotations

import os
import queue
import threading


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous [lo, hi) of ``n`` reads for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def batch_owner(batch_index: int, world_size: int) -> int:
    """Round-robin owner of a features batch when the input is a stream (feature file)."""
    return batch_index % world_size


def range_owner(batch_index: int, n_batches: int, world_size: int) -> int:
    """Owner of batch ``batch_index`` when the ``n_batches`` batches are split into contiguous ranges (packed reads)."""
    for r in range(world_size):
        lo, hi = shard_range(n_batches, r, world_size)
        if lo <= batch_index < hi:
            return r
    raise IndexError(batch_index)


def gather_in_order(local_items, rank: int, world_size: int, group=None):
    """``local_items``: list of (global_index, payload).  Rank 0 returns all payloads sorted by
    global index, other ranks return None."""
    if world_size == 1:
        return [p for _, p in sorted(local_items, key=lambda t: t[0])]
    import torch.distributed as dist
    bucket = [None] * world_size if rank == 0 else None
    dist.gather_object(local_items, bucket, dst=0, group=group)
    if rank != 0:
        return None
    merged = [it for part in bucket for it in part]
    return [p for _, p in sorted(merged, key=lambda t: t[0])]


def run_sharded(batches, n_workers, owner, make_worker, process, emit, depth=4):
    """Thread-per-device pipeline.

    ``batches``      iterable of work items in input order (item i has sequence number i)
    ``owner(i)``     -> worker index of item i
    ``make_worker(w)`` -> per-worker state (an Engine on device w); called inside the worker thread
    ``process(state, item)`` -> result
    ``emit(result)`` called on the calling thread in input order
    Returns the number of items.  The first exception of any worker is re-raised here."""
    inq = [queue.Queue(maxsize=depth) for _ in range(n_workers)]
    outq = queue.Queue()
    errors = []

    def work(w):
        state = None
        try:
            state = make_worker(w)
            while True:
                it = inq[w].get()
                if it is None:
                    break
                seq, item = it
                outq.put((seq, process(state, item)))
        except BaseException as e:      # noqa: BLE001 - forwarded to the caller
            errors.append(e)
            outq.put((-1, None))
            while True:                 # keep draining so the feeder never blocks on a dead worker
                if inq[w].get() is None:
                    break
        finally:
            if state is not None and hasattr(state, "close"):
                state.close()

    threads = [threading.Thread(target=work, args=(w,), daemon=True) for w in range(n_workers)]
    for t in threads:
        t.start()
    pending, next_seq, total = {}, 0, 0

    def drain(block):
        nonlocal next_seq
        while True:
            try:
                seq, res = outq.get(block, timeout=0.05 if block else None) if block else outq.get_nowait()
            except queue.Empty:
                return
            if seq < 0:
                raise errors[0]
            pending[seq] = res
            while next_seq in pending:
                emit(pending.pop(next_seq))
                next_seq += 1
            if block:
                return

    try:
        for i, item in enumerate(batches):
            total += 1
            w = owner(i)
            while True:
                try:
                    inq[w].put((i, item), timeout=0.05)
                    break
                except queue.Full:
                    drain(False)
                    if errors:
                        raise errors[0]
            drain(False)
        for q in inq:
            q.put(None)
        while next_seq < total:
            drain(True)
            if errors:
                raise errors[0]
    finally:
        for q in inq:
            try:
                q.put_nowait(None)
            except queue.Full:
                pass
    for t in threads:
        t.join()
    return total

"""call modifications from extracted features (or packed reads) with the B200 engine.

Host-side mirror of the reference's ``deepsignal/call_modifications.py``: same function
names, argument meaning, feature-file parsing (:35-91), per-sample output formatting
(:183-190) and 10-column result TSV (chrom, pos, strand, pos_in_strand, read_name,
read_strand, prob_0, prob_1, called_label, k_mer).  The TensorFlow session of the
reference (``sess.run([activation_logits, prediction], feed_dict)``, :168-178) is replaced
by one call into libdsb200 (``Engine.forward`` -> ``dsb_forward_host``); nothing on the
hot path is computed in Python.

Differences that are deliberate:
* one reader thread + one GPU worker (thread + ``dsb_ctx``) per device instead of N TensorFlow
  processes (the reference's process fan-out exists to use CPU cores; the GPU engine is fed by
  large batches instead); ``--devices 0,1,...`` or ``torchrun`` shard the reads over several GPUs
  (deepsignal_b200/parallel.py).  ``--nproc`` is accepted and only bounds host parsing threads.
* a features batch is forwarded in one call rather than in ``batch_size`` slices: the model is
  in inference mode, so results are batch-composition invariant; ``batch_num`` is still
  reported as ceil(n / batch_size) like the reference.
"""
from __future__ import absolute_import

import os
import sys
import threading
import time
import queue

import numpy as np

from .utils.process_utils import base2code_dna, code2base_dna, str2bool, display_args  # noqa: F401
from .utils.process_utils import get_motif_seqs, get_contig2len

queen_size_border = 100


def _parse_feature_line(words):
    return ("\t".join(words[0:6]),
            [base2code_dna[x] for x in words[6]],          # KeyError on non-ACGTN like the reference (:52)
            [float(x) for x in words[7].split(",")],
            [float(x) for x in words[8].split(",")],
            [int(x) for x in words[9].split(",")],
            [float(x) for x in words[10].split(",")],
            int(words[11]))


def _read_features_file(features_file, features_batch_q, f5_batch_num=20):
    """Group feature-file lines into batches of ``f5_batch_num`` *reads* (reference :35-91).

    Puts 7-tuples (sampleinfo, kmers, base_means, base_stds, base_signal_lens, cent_signals,
    labels) on the queue and finally the string "kill"."""
    r_num = 0
    cols = ([], [], [], [], [], [], [])
    with open(features_file, "r") as rf:
        readid_pre = None
        for line in rf:
            words = line.strip().split("\t")
            if readid_pre is None:
                readid_pre = words[4]
            elif words[4] != readid_pre:
                r_num += 1
                readid_pre = words[4]
                if r_num % f5_batch_num == 0:
                    features_batch_q.put(cols)
                    cols = ([], [], [], [], [], [], [])
            for dst, v in zip(cols, _parse_feature_line(words)):
                dst.append(v)
    if len(cols[0]) > 0:
        features_batch_q.put(cols)
    features_batch_q.put("kill")


def _call_mods(features_batch, engine, model, init_learning_rate, batch_size):
    """Mirror of reference ``_call_mods`` (:149-194): returns (pred_str, accuracy, batch_num).

    ``engine`` plays the role of the TF session, ``model`` carries the precision ("fp32", "fp16x2",
    "fp16", "bf16").  ``init_learning_rate`` is accepted for signature compatibility and unused (the
    reference only feeds it to a placeholder)."""
    sampleinfo, kmers, base_means, base_stds, base_signal_lens, cent_signals, labels = features_batch
    labels = np.reshape(labels, (len(labels)))
    n = len(sampleinfo)
    if n == 0:
        return [], float("nan"), 0
    kmers_a = np.asarray(kmers, dtype=np.int32)
    probs, prediction = engine.forward(kmers_a, np.asarray(base_means, np.float32), np.asarray(base_stds, np.float32),
                                       np.asarray(base_signal_lens, np.float32), np.asarray(cent_signals, np.float32),
                                       precision=getattr(model, "precision", "fp16x2"))
    prediction = prediction.astype(np.int64)
    batch_num = 0
    accuracys = []
    for i in range(0, n, batch_size):
        accuracys.append(float(np.mean(labels[i:i + batch_size] == prediction[i:i + batch_size])))
        batch_num += 1
    # prob_0/(prob_0+prob_1) in float32, printed with numpy's shortest repr — exactly reference :185-190
    p0, p1 = probs[:, 0], probs[:, 1]
    s = p0 + p1
    p0n, p1n = p0 / s, p1 / s
    pred_str = []
    for idx in range(n):
        pred_str.append("\t".join([sampleinfo[idx], str(p0n[idx]), str(p1n[idx]), str(prediction[idx]),
                                   "".join([code2base_dna[x] for x in kmers[idx]])]))
    return pred_str, float(np.mean(accuracys)), batch_num


def _write_predstr_to_file(write_fp, predstr_q):
    with open(write_fp, "w") as wf:
        while True:
            pred_str = predstr_q.get()
            if isinstance(pred_str, str) and pred_str == "kill":
                break
            for one_pred_str in pred_str:
                wf.write(one_pred_str + "\n")
            wf.flush()


class _ModelHandle(object):
    """What the reference's ``Model(base_num, signal_num, class_num, ...)`` is to ``_call_mods``."""

    def __init__(self, base_num, signal_num, class_num, is_cnn=True, is_base=True, is_rnn=True, precision="fp16x2"):
        if not (is_cnn or is_rnn):
            raise ValueError("at least one of is_cnn/is_rnn should be True")
        if class_num != 2:
            raise ValueError("the call_mods engine supports class_num == 2 (the reference default)")
        self.base_num, self.signal_num, self.class_num = base_num, signal_num, class_num
        self.is_cnn, self.is_base, self.is_rnn, self.precision = is_cnn, is_base, is_rnn, precision


def _load_engine(model_path, kmer_len, cent_signals_len, is_cnn, is_rnn, is_base, device=0):
    from .engine import Engine
    from .weights import load_checkpoint
    eng = Engine(device)
    eng.load_weights(load_checkpoint(model_path), is_cnn=is_cnn, is_rnn=is_rnn, is_base=is_base,
                     kmer_len=kmer_len, signal_len=cent_signals_len)
    return eng


def _read_batches(reads_path, f5_args):
    """fast5-directory mode of the reference (:94-122,437-457) on packed reads / fast5 files: yields ReadBatch slices of
    ``f5_batch_num * 64`` reads (the reference hands ``f5_batch_num`` files to each CPU worker, :104-108)."""
    from .extract_features import load_reads
    (is_recursive, corrected_group, basecall_subgroup, reference_path, is_dna, normalize_method, motifs, mod_loc,
     methy_label, f5_batch_num, position_file) = f5_args
    chrom2len = get_contig2len(reference_path) if reference_path is not None else None
    reads, errornum = load_reads(reads_path, chrom2len, is_recursive, corrected_group, basecall_subgroup,
                                 return_errors=True)
    step = max(1, int(f5_batch_num)) * 64
    spans = [(lo, min(reads.n_reads, lo + step)) for lo in range(0, reads.n_reads, step)]
    return reads, spans, errornum


class _Work(object):
    """One unit of work for a device worker: kind in {"reads", "tuple", "block"} + payload."""
    __slots__ = ("kind", "payload")

    def __init__(self, kind, payload):
        self.kind, self.payload = kind, payload


def _process_work(engine, model, work, ctx):
    """Runs on the worker that owns ``work``: (extraction +) forward + formatting.  Returns (lines, batch_num, errors)."""
    if work.kind == "block":        # C++-parsed feature-file block -> forward -> C++-formatted result lines
        from .fast_tsv import format_calls
        block = work.payload
        probs, prediction = engine.forward(block.kmer, block.means, block.stds, block.lens, block.signals,
                                           precision=model.precision)
        return [format_calls(block, probs, prediction).decode().rstrip("\n")], -(-block.n // ctx["batch_size"]), 0
    if work.kind == "reads":        # packed reads -> extraction kernels -> forward, all on this worker's device
        from .extract_features import extract_batch_arrays, failed_reads
        arr, info = extract_batch_arrays(engine, work.payload, ctx["motif_seqs"], ctx["mod_loc"], ctx["kmer_len"],
                                         ctx["cent_signals_len"], ctx["normalize_method"], ctx["positions"], want64=False)
        errors = int(np.sum(failed_reads(arr)))
        if len(info) == 0:
            return [], 0, errors
        if os.environ.get("DSB_PY_TSV", "0") != "1":
            # forward + the C++ result formatter (byte-identical to the Python formatting of _call_mods, ~5 us per site faster)
            from .fast_tsv import block_from_info, format_calls
            probs, prediction = engine.forward(arr["kmer"], arr["means"], arr["stds"], arr["lens"], arr["signals"],
                                               precision=model.precision)
            text = format_calls(block_from_info(info, arr["kmer"]), probs, prediction).decode().rstrip("\n")
            return [text], -(-len(info) // ctx["batch_size"]), errors
        batch = (info, arr["kmer"], arr["means"], arr["stds"], arr["lens"], arr["signals"], [ctx["methy_label"]] * len(info))
        pred_str, _, nb = _call_mods(batch, engine, model, ctx["learning_rate"], ctx["batch_size"])
        return pred_str, nb, errors
    pred_str, _, nb = _call_mods(work.payload, engine, model, ctx["learning_rate"], ctx["batch_size"])
    return pred_str, nb, 0


def _iter_work(input_path, kmer_len, cent_signals_len, nproc, f5_args, state):
    """Work items of one call_mods run, in input order; ``state`` receives n_reads / n_failed for the final report."""
    f5_batch_num = f5_args[9]
    if os.path.isdir(input_path) or input_path.endswith(".npz"):
        reads, spans, errornum = _read_batches(input_path, f5_args)
        state["n_reads"], state["n_failed"], state["n_items"] = reads.n_reads + errornum, errornum, len(spans)
        for lo, hi in spans:
            yield _Work("reads", reads.slice(lo, hi))
    elif input_path.endswith(".bin"):
        # the reference's 1628-byte binary sample records (scripts/generate_binary_feature_file.py): no read / position
        # columns exist in that format, so the six info columns are printed as '.'
        from .binary_format import read_binary
        rec = read_binary(input_path, kmer_len, cent_signals_len)
        step = 1 << 18
        for lo in range(0, len(rec["labels"]), step):
            sl = slice(lo, lo + step)
            n_here = len(rec["labels"][sl])
            yield _Work("tuple", (["\t".join(["."] * 6)] * n_here, rec["kmer"][sl], rec["means"][sl], rec["stds"][sl],
                                  rec["lens"][sl], rec["signals"][sl], rec["labels"][sl]))
    elif os.environ.get("DSB_PY_TSV", "0") != "1":
        # fast path: C++ text codec (csrc/tsv.cpp); identical parsing rules and byte-identical result lines
        from .fast_tsv import iter_feature_blocks
        for block in iter_feature_blocks(input_path, kmer_len, cent_signals_len, nthreads=max(1, nproc)):
            yield _Work("block", block)
    else:
        features_batch_q = queue.Queue(maxsize=queen_size_border)

        def reader():       # a malformed line (e.g. a base outside ACGTN, reference :52) must fail the run, not hang it
            try:
                _read_features_file(input_path, features_batch_q, f5_batch_num * 64)
            except BaseException as e:      # noqa: BLE001 - re-raised by the consumer below
                features_batch_q.put(e)
        p_rf = threading.Thread(target=reader, daemon=True)
        p_rf.start()
        while True:
            features_batch = features_batch_q.get()
            if isinstance(features_batch, BaseException):
                raise features_batch
            if isinstance(features_batch, str) and features_batch == "kill":
                break
            yield _Work("tuple", features_batch)
        p_rf.join()


def call_mods(input_path, model_path, result_file, kmer_len, cent_signals_len,
              batch_size, learning_rate, class_num, nproc, is_gpu, is_rnn, is_base, is_cnn,
              f5_args, precision="fp16x2", device=0, devices=None):
    """Same positional signature as the reference ``call_mods`` (:417-495) plus ``precision`` / ``device`` /
    ``devices``.  ``is_gpu`` is accepted for compatibility; this implementation always runs on the GPU and fails
    loudly if none is present.

    Reads are independent, so the work shards with no collective (reference fan-out: :461-476):
    ``devices=[0, 1, ...]`` runs one worker thread + engine per GPU in this process; under ``torchrun``
    (WORLD_SIZE > 1) every rank handles the batches it owns on ``cuda:LOCAL_RANK`` and rank 0 writes the file.
    Either way the result file is byte-identical to a single-GPU run: lines keep input order, each read's lines
    stay contiguous, and per-sample results do not depend on batch composition (inference-mode graph)."""
    from . import parallel
    start = time.time()
    model_path = os.path.abspath(model_path)
    input_path = os.path.abspath(input_path)
    model = _ModelHandle(kmer_len, cent_signals_len, class_num, is_cnn, is_base, is_rnn, precision)
    (_, _, _, _, is_dna, normalize_method, motifs, mod_loc, methy_label, f5_batch_num, position_file) = f5_args
    positions = None
    if position_file is not None:
        from .extract_features import _read_position_file
        positions = _read_position_file(position_file)
    ctx = dict(batch_size=batch_size, learning_rate=learning_rate, kmer_len=kmer_len, cent_signals_len=cent_signals_len,
               motif_seqs=get_motif_seqs(motifs, is_dna), mod_loc=mod_loc, normalize_method=normalize_method,
               positions=positions, methy_label=methy_label)
    state = {"n_reads": 0, "n_failed": 0, "n_items": None}
    rank, local_rank, world_size = parallel.world()
    devices = [int(d) for d in devices] if devices else [local_rank if world_size > 1 else int(device)]
    make_engine = lambda w: _load_engine(model_path, kmer_len, cent_signals_len, is_cnn, is_rnn, is_base, devices[w])
    counters = {"batches": 0, "errors": 0}

    if world_size > 1:
        # ---- one process per GPU (torchrun): own batches round-robin, rank 0 gathers the text and writes it
        import torch.distributed as dist
        own_group = not dist.is_initialized()
        if own_group:
            dist.init_process_group(backend="gloo")
        engine = make_engine(0)
        local, stats = [], [0, 0]
        work_iter = _iter_work(input_path, kmer_len, cent_signals_len, nproc, f5_args, state)
        for i, work in enumerate(work_iter):
            owner = (parallel.range_owner(i, state["n_items"], world_size) if state["n_items"] is not None
                     else parallel.batch_owner(i, world_size))
            if owner != rank:
                continue
            lines, nb, err = _process_work(engine, model, work, ctx)
            local.append((i, lines))
            stats[0] += nb
            stats[1] += err
        engine.close()
        merged = parallel.gather_in_order(local, rank, world_size)
        all_stats = [None] * world_size
        dist.all_gather_object(all_stats, stats)
        counters["batches"] = sum(s[0] for s in all_stats)
        counters["errors"] = sum(s[1] for s in all_stats)
        if rank == 0:
            with open(result_file, "w") as wf:
                for lines in merged:
                    for one in lines:
                        wf.write(one + "\n")
        if own_group:
            dist.barrier()
            dist.destroy_process_group()
    else:
        # ---- one worker thread + engine per device in this process; the writer keeps input order
        pred_str_q = queue.Queue(maxsize=queen_size_border)
        p_w = threading.Thread(target=_write_predstr_to_file, args=(result_file, pred_str_q), daemon=True)
        p_w.start()

        def emit(res):
            lines, nb, err = res
            pred_str_q.put(lines)
            counters["batches"] += nb
            counters["errors"] += err

        nw = len(devices)
        items = _iter_work(input_path, kmer_len, cent_signals_len, nproc, f5_args, state)

        def owner(i):
            return parallel.range_owner(i, state["n_items"], nw) if state["n_items"] is not None else parallel.batch_owner(i, nw)
        try:
            parallel.run_sharded(items, nw, owner, make_engine, lambda eng, work: _process_work(eng, model, work, ctx), emit)
        finally:
            pred_str_q.put("kill")
            p_w.join()
    if rank == 0:
        if state["n_reads"]:
            print("%d of %d fast5 files failed.." % (state["n_failed"] + counters["errors"], state["n_reads"]))
        print("call_mods process {} ending, proceed {} batches".format(os.getpid(), counters["batches"]))
        print("call_mods costs %.2f seconds.." % (time.time() - start))


def main():
    from .deepsignal import main as cli_main
    sys.argv.insert(1, "call_mods")
    return cli_main()


if __name__ == "__main__":
    sys.exit(main())

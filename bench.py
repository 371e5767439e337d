#!/usr/bin/env python
"""bench.py — CpG samples/sec of the call_mods hot path on N B200s (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[2] — the call_mods forward on synthetic
bn17/sn360 samples with the seed-17 random-init ("stress") checkpoint.  A step is one pass of
`--samples-per-step` samples per GPU through `dsb_forward_dev` (inputs resident in HBM, 1.8 GB per
2^20 samples — far larger than the 126 MB L2, so no L2 flush is needed between steps).  `e2e` is the
same metric through `dsb_forward_host` with pinned HOST buffers: host->device copies of the five
input arrays and the device->host read of probs/pred are inside the timed region.

Reads shard across GPUs with no collective (one process per GPU; torch.distributed is used only
for the barrier and the max-over-ranks of the device time) -> "scaling": "weak".

`--impl reference` times the CPU restatement of the reference's un-folded TF graph (oracle/model_oracle.py,
torch CPU fp32, all host threads, batch 512 like call_modifications.py:157) — TensorFlow itself
cannot be installed in this image, so kind = "port".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_INCEPT_A = 90 * (512 + 480) + 2 * 90 * 960 + 5 * 45 * 960 + 3 * 23 * 960   # 544,320 B per sample, see roofline below
FLOP_PER_SAMPLE = 198.61e6          # executed MACs*2 after the exact folds
BYTES_IN_PER_SAMPLE = 17 * 4 * 4 + 360 * 4   # 1712
BYTES_OUT_PER_SAMPLE = 12


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(bf16_tflops=d.get("bf16_tflops", 1590.0), bf16_tflops_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm_gbs=d.get("hbm_gbs", 6650.0), source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def cpu_baseline(n_samples, ckpt, threads=None):
    """TF-equivalent CPU restatement (un-folded graph, fp32, batch 512) timed on the host cores."""
    import torch
    from oracle import model_oracle as mo
    from deepsignal_b200.synth import synth_features
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    f = synth_features(n_samples, seed=20260103)
    lens = f["lens"].astype(np.float32)
    mo.forward(ckpt, f["kmer"][:64], f["means"][:64], f["stds"][:64], lens[:64], f["signals"][:64], dtype=torch.float32)
    t0 = time.perf_counter()
    for i in range(0, n_samples, 512):
        j = min(n_samples, i + 512)
        mo.forward(ckpt, f["kmer"][i:j], f["means"][i:j], f["stds"][i:j], lens[i:j], f["signals"][i:j], dtype=torch.float32)
    dt = time.perf_counter() - t0
    return n_samples / dt, cores, dt


def device_inputs(torch, n, dev, seed):
    """cfg-3 distributions generated on the device."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    kmer = torch.randint(0, 4, (n, 17), device=dev, dtype=torch.int32, generator=g)
    kmer[:, 8] = 1
    kmer[:, 9] = 2
    u = torch.rand((n, 17), device=dev, generator=g).clamp_(1e-7, 1.0)
    lens = 3 + torch.ceil(torch.log(u) / np.log(1 - 1 / 5.9)).clamp_(min=1)          # geometric(p=1/5.9)
    stall = torch.rand((n, 17), device=dev, generator=g) < 0.01
    lens = torch.where(stall, lens * 20, lens).to(torch.float32)
    means = torch.randn((n, 17), device=dev, generator=g)
    stds = (torch.randn((n, 17), device=dev, generator=g) * 0.05 + 0.15).abs_() + 0.01
    signals = torch.randn((n, 360), device=dev, generator=g)
    nreal = lens.sum(dim=1).clamp_(max=360)
    signals.masked_fill_(torch.arange(360, device=dev)[None, :] >= nreal[:, None], 0.0)
    return kmer.contiguous(), means.contiguous(), stds.contiguous(), lens.contiguous(), signals.contiguous()


def bench_reads(a, eng, torch, dev, rank, world, barrier, max_over_ranks, ckpt):
    """configs[1] (extraction only) and configs[3] (extraction + forward) on synthetic resquiggled reads."""
    from deepsignal_b200.synth import synth_reads, SEED_CFG2, SEED_CFG4
    from oracle import features_oracle as fo
    rb = synth_reads(a.reads, seed=(SEED_CFG2 if a.workload == "extract" else SEED_CFG4) + rank)
    dr = eng.reads_to_device(rb, dev)
    n_raw, n_ev = int(rb.raw_off[-1]), int(rb.ev_off[-1])
    cap = max(1024, int(n_ev * 0.03))
    bufs = eng.alloc_site_bufs(cap, dev)

    def step():
        if a.workload == "extract":
            return eng.extract_torch(dr, rb.n_reads, bufs)
        return eng.extract_call_torch(dr, rb.n_reads, precision=a.precision, site_bufs=bufs)[0]

    for _ in range(a.warmup):
        n_sites = step()
    sampler = ClockSampler(torch.cuda.current_device())
    barrier()
    sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        n_sites = step()
    e1.record()
    barrier()
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    sites_s = world * n_sites * a.steps / (ms * 1e-3)
    alg_bytes = 2.0 * n_raw + 9.0 * n_ev + 1677.0 * n_sites          # SURVEY §8d per-read figure
    pk = peaks()
    out = {
        "metric": "CpG samples/sec ({})".format("extract_features" if a.workload == "extract" else "extract + call_mods"),
        "value": sites_s, "unit": "samples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if a.workload == "extract" else a.precision, "data": "synthetic",
        "config": {"workload": "BASELINE configs[{}]: {} synthetic tombo-style R9.4 reads per GPU, {} raw samples, {} bases, {} CpG sites".format(
            1 if a.workload == "extract" else 3, rb.n_reads, n_raw, n_ev, n_sites),
            "reads_per_s": world * rb.n_reads * a.steps / (ms * 1e-3),
            "l2": "raw signal {} MB + outputs {} MB per step exceed the 126 MB L2".format(2 * n_raw >> 20, int(1677 * n_sites) >> 20)},
        "clocks": clocks, "gpu_launches": int(launches)}
    if a.workload == "extract":
        ach = alg_bytes * a.steps / (ms * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                           "traffic": None, "kernel": "extraction kernel chain (5 launches per batch)",
                           "peak_source": pk["source"]}
        if rank == 0 and not a.no_cpu_baseline:
            k = min(rb.n_reads, 1500)          # ~10 s of single-core NumPy work
            t0 = time.perf_counter()
            ns = 0
            for i in range(k):
                ns += len(fo.extract_read(rb.read_dict(i), ["CG"]))
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": ns / dt, "unit": "samples/s", "cores": 1, "kind": "port",
                                   "sample": "{} reads ({} sites) through the NumPy restatement of _extract_features, {:.1f} s".format(k, ns, dt)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--samples-per-step", type=int, default=1 << 20)
    ap.add_argument("--e2e-samples", type=int, default=1 << 19)
    ap.add_argument("--max-chunk", type=int, default=37888)
    ap.add_argument("--cpu-samples", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="forward", choices=["forward", "extract", "e2e"],
                    help="forward = BASELINE configs[2] (default, the metric's workload); extract = configs[1] "
                         "(4k reads, extraction only); e2e = configs[3] (extract + call_mods, read-sharded)")
    ap.add_argument("--reads", type=int, default=4000)
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from deepsignal_b200 import weights as W
    ckpt = W.random_checkpoint(17, "stress")
    workload = "call_mods forward, bn17/sn360 synthetic samples (BASELINE configs[2]), seed-17 random-init checkpoint"

    if a.impl == "reference":
        if rank != 0:
            return 0
        per_step = max(512, a.cpu_samples // 4)
        for _ in range(max(1, min(a.warmup, 1))):
            cpu_baseline(512, ckpt)
        t_tot, n_tot = 0.0, 0
        cores = os.cpu_count() or 1
        for _ in range(a.steps):
            v, cores, dt = cpu_baseline(per_step, ckpt)
            t_tot += dt
            n_tot += per_step
        val = n_tot / t_tot
        sample = "{} samples per step, torch-CPU fp32 restatement of the un-folded graph, batch 512".format(per_step)
        print(json.dumps({
            "impl": "reference", "metric": "CpG samples/sec (call_mods forward)", "value": val, "unit": "samples/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_tot / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "samples_per_step": per_step},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist
    from deepsignal_b200.engine import Engine

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = Engine(local_rank)
    eng.load_weights(ckpt)
    eng.set_max_chunk(a.max_chunk)
    if a.workload != "forward":
        out = bench_reads(a, eng, torch, dev, rank, world, barrier, max_over_ranks, ckpt)
        if rank == 0:
            print(json.dumps(out))
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return 0
    n = a.samples_per_step
    x = device_inputs(torch, n, dev, 20260103 + rank)
    probs = torch.empty((n, 2), dtype=torch.float32, device=dev)
    pred = torch.empty((n,), dtype=torch.int32, device=dev)

    def step():
        eng.forward_torch(*x, precision=a.precision, probs=probs, pred=pred)

    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * n * a.steps / (ms * 1e-3)
    frac_called = float(pred.float().mean().item())

    # ---- per-kernel roofline: one extra profiled pass (CUDA events around every launch, single stream)
    pk = peaks()
    eng.set_profiling(True)
    step()
    torch.cuda.synchronize(dev)
    prof = eng.profile()
    eng.set_profiling(False)
    tc = [prof[k] for k in ("gemm", "lstm_gemm", "incept_a")]
    tc_ms = sum(v["ms"] for v in tc)
    tc_flops = sum(v["flops"] for v in tc)
    tc_launches = sum(v["launches"] for v in tc)
    traffic_a, traffic_tc = None, None
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")      # ncu dram__bytes_{read,write}.sum of the same command, per launch
    if os.path.exists(tp):
        with open(tp) as fh:
            tj = json.load(fh)
        ka = [v for k, v in tj.get("kernels", {}).items() if k.startswith("incept_a_kernel")]
        if ka:
            traffic_a = ka[0]["dram_MB"] * 1e6 / ka[0]["launches"] * (n / tj["chunk_samples"] if n < tj["chunk_samples"] else 1.0)
        traffic_tc = tj.get("tensor_core_dram_bytes_per_launch")
    # dominant kernel = incept_a_kernel (11 launches per chunk, ~35 % of the step): HBM-bound.  Algorithmic bytes per sample
    # (DESIGN.md section 4): sum over the 11 blocks of L_b real rows x (2*cin_b in + 480 out) = 544,320 B.
    a_prof = prof["incept_a"]
    a_bytes = ALG_BYTES_INCEPT_A * n
    a_ach = a_bytes / (a_prof["ms"] * 1e-3) / 1e9 if a_prof["ms"] > 0 else 0.0
    chunks = -(-n // a.max_chunk)
    roofline = {"bound": "hbm", "achieved": a_ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": a_ach / pk["hbm_gbs"],
                "traffic": traffic_a, "traffic_source": "profiles/r1_traffic.json (ncu dram bytes per incept_a_kernel launch, one 37888-sample chunk; algorithmic: %.0f)" % (a_bytes / max(a_prof["launches"], 1)),
                "kernel": "incept_a_kernel (inception block, first half: 1x1 convs + in-SM max pool; 11 launches per chunk)",
                "launches_profiled": a_prof["launches"], "avg_launch_ms": a_prof["ms"] / max(a_prof["launches"], 1),
                "share_of_step": a_prof["ms"] / max(sum(v["ms"] for v in prof.values()), 1e-9),
                "peak_source": pk["source"] + ", STREAM-style copy",
                "breakdown_ms": {k: v["ms"] for k, v in prof.items()},
                # all tensor-core kernels together (front end, inception A/B, LSTM) against the tensor roofline
                "tensor": {"achieved": tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                           "frac": (tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0) / pk["bf16_tflops_sustained"],
                           "launches_profiled": tc_launches, "traffic": traffic_tc,
                           "peak_source": "sustained bf16 figure (kernels timed inside a long step)"},
                "whole_step_frac_of_tensor_peak": value / world * FLOP_PER_SAMPLE / 1e12 / pk["bf16_tflops_sustained"],
                "chunks_per_step": chunks}

    # ---- end to end through the host entry point (pinned host buffers, copies inside the timed region)
    ne = a.e2e_samples
    hx = [t[:ne].cpu().pin_memory() for t in x]
    hprobs = torch.empty((ne, 2), dtype=torch.float32).pin_memory()
    hpred = torch.empty((ne,), dtype=torch.int32).pin_memory()

    def e2e_step():
        eng.forward_ptrs(ne, hx[0].data_ptr(), hx[1].data_ptr(), hx[2].data_ptr(), hx[3].data_ptr(), hx[4].data_ptr(),
                         hprobs.data_ptr(), hpred.data_ptr(), precision=a.precision, host=True)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_val = world * ne * a.steps / t_e2e
    same = bool(torch.equal(hpred.to(dev), pred[:ne]))

    out = {
        "metric": "CpG samples/sec (call_mods forward)", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
        "config": {"workload": workload, "samples_per_step_per_gpu": n, "max_chunk": a.max_chunk,
                   "l2": "inputs (1.8 GB per 2^20 samples) and per-chunk activations exceed the 126 MB L2; no flush needed",
                   "parallelism": "reads sharded, {} replica(s), no collective".format(world),
                   "frac_called_modified": frac_called, "e2e_matches_device_path": same},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": ne * BYTES_IN_PER_SAMPLE,
                "d2h_bytes_per_step": ne * BYTES_OUT_PER_SAMPLE, "samples_per_step_per_gpu": ne},
        "roofline": roofline,
    }
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, cores, dt = cpu_baseline(a.cpu_samples, ckpt)
        out["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                               "sample": "{} samples, torch-CPU fp32 restatement of the un-folded reference graph, batch 512, {:.1f} s".format(a.cpu_samples, dt)}
    if rank == 0:
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

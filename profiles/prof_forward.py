#!/usr/bin/env python
"""Per-class kernel time of the forward for one chunk size (CUDA events inside libdsb200, serialised streams).

    python profiles/prof_forward.py [--chunk 37888] [--chunks 4] [--precision fp16x2] [--opt fused_blocks=0 ...]

Prints one JSON line: wall ms per chunk (multi-stream, un-profiled) and the profiled per-class sums per chunk.
Used for the ablation tables in profiles/*.md (--opt sets fields of dsb_plan_options, e.g. ablate_a=2)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsignal_b200.engine import Engine          # noqa: E402
from deepsignal_b200.weights import random_checkpoint  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=37888)
    ap.add_argument("--chunks", type=int, default=4)
    ap.add_argument("--precision", default="fp16x2")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--opt", nargs="*", default=[], help="dsb_plan_options fields, name=value")
    a = ap.parse_args()
    opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in a.opt}
    n = a.chunk * a.chunks
    eng = Engine(0)
    eng.load_weights(random_checkpoint(17))
    eng.set_max_chunk(a.chunk)
    if opts:
        eng.set_plan_options(**opts)
    g = torch.Generator(device="cuda").manual_seed(1)
    kmer = torch.randint(0, 4, (n, 17), device="cuda", dtype=torch.int32, generator=g)
    means = torch.randn(n, 17, device="cuda", generator=g)
    stds = torch.rand(n, 17, device="cuda", generator=g)
    lens = torch.randint(1, 60, (n, 17), device="cuda", generator=g).float()
    signals = torch.randn(n, 360, device="cuda", generator=g)
    run = lambda: eng.forward_torch(kmer, means, stds, lens, signals, precision=a.precision)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(a.reps):
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    eng.set_profiling(True)
    run()
    torch.cuda.synchronize()
    prof = eng.profile()
    eng.set_profiling(False)
    out = {"precision": a.precision, "chunk": a.chunk, "ms_per_chunk": best / a.chunks, "samples_per_s": n / best * 1e3, "opts": opts, "lib": os.path.basename(os.environ.get("DSB_LIB", "libdsb200.so"))}
    for k, v in prof.items():
        if v["launches"]:
            out[k + "_ms_per_chunk"] = round(v["ms"] / a.chunks, 4)
            out[k + "_launches"] = v["launches"] // a.chunks
    print(json.dumps(out))


if __name__ == "__main__":
    main()

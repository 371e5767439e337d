#!/usr/bin/env python
"""SURVEY section 8(d) cfg 5: throughput of the forward over batch sizes 512 ... 262,144 (device-resident inputs,
one dsb_forward_dev call per batch, best of a few repetitions).  Writes one JSON line per batch."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsignal_b200.engine import Engine              # noqa: E402
from deepsignal_b200.weights import random_checkpoint  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    eng = Engine(0)
    eng.load_weights(random_checkpoint(17))
    nmax = 262144
    g = torch.Generator(device="cuda").manual_seed(1)
    kmer = torch.randint(0, 4, (nmax, 17), device="cuda", dtype=torch.int32, generator=g)
    means = torch.randn(nmax, 17, device="cuda", generator=g)
    stds = torch.rand(nmax, 17, device="cuda", generator=g)
    lens = torch.randint(1, 60, (nmax, 17), device="cuda", generator=g).float()
    signals = torch.randn(nmax, 360, device="cuda", generator=g)
    n = 512
    while n <= nmax:
        run = lambda: eng.forward_torch(kmer[:n], means[:n], stds[:n], lens[:n], signals[:n], precision=prec)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        best = 1e30
        reps = max(3, min(50, (1 << 19) // n))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
        print(json.dumps({"batch": n, "precision": prec, "ms": best, "samples_per_s": n / best * 1e3,
                          "tensor_frac_of_sustained_peak": n / best * 1e3 * 198.61e6 / 1403e12}))
        n *= 2


if __name__ == "__main__":
    main()

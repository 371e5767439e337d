#!/usr/bin/env python
"""Small end-to-end pass of every device path for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python profiles/sanitize.py

One 300-sample forward in each precision mode (fused and un-fused plans), one extraction + fused extract->forward
on 6 short reads, one site-frequency reduce.  Sizes are tiny because the tools slow kernels down by 10-100x."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsignal_b200.engine import Engine            # noqa: E402
from deepsignal_b200 import weights as W             # noqa: E402
from deepsignal_b200.synth import synth_features, synth_reads  # noqa: E402


def main():
    modes = sys.argv[1:] or ["fp16x2", "fp16", "bf16", "fp32"]
    eng = Engine(0)
    eng.load_weights(W.random_checkpoint(17, "stress"))
    f = synth_features(300, seed=3)
    x = (f["kmer"], f["means"], f["stds"], f["lens"].astype(np.float32), f["signals"])
    for prec in modes:
        p, c = eng.forward(*x, precision=prec)
        print(prec, "fused", float(p.sum()))
        if prec != "fp32":
            eng.set_plan_options(fused_blocks=0, fused_front=0, lstm_seq=0)
            p2, _ = eng.forward(*[a[:130] for a in x], precision=prec)
            print(prec, "unfused", float(np.abs(p2 - p[:130]).max()))
            eng.set_plan_options(fused_blocks=1, fused_front=1, lstm_seq=1)
    rb = synth_reads(6, seed=5, nb_lo=200, nb_hi=400, cg_frac=0.04)
    got = eng.extract(rb, ["CG"])
    print("extract sites", got["n_sites"])
    import torch
    dev = torch.device("cuda", 0)
    dr = eng.reads_to_device(rb, dev)
    n, bufs = eng.extract_call_torch(dr, rb.n_reads, precision=modes[0], capacity=4096)
    torch.cuda.synchronize()
    print("extract+call sites", n, float(bufs["probs"][:n].sum()))
    rng = np.random.default_rng(0)
    m = 5000
    r = eng.site_frequency(rng.integers(0, 3, m), rng.integers(0, 200, m), rng.random(m), rng.random(m), rng.integers(0, 2, m), np.ones(m))
    print("site_frequency sites", len(r["first"]))
    eng.close()


if __name__ == "__main__":
    main()

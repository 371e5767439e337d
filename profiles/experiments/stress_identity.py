"""Repeated-run determinism check of the default path (races in the warp-specialised kernels would show up as run-to-run
differences): the same inputs through the forward many times, all outputs must be bit-identical to the first."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepsignal_b200.engine import Engine
from deepsignal_b200 import weights as W
from deepsignal_b200.synth import synth_features

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
eng = Engine(0)
eng.load_weights(W.random_checkpoint(17, "stress"))
bad = 0
for n, prec in ((300, "bf16"), (5000, "bf16"), (5000, "fp16"), (40000, "bf16")):
    f = synth_features(n, seed=11)
    x = (f["kmer"], f["means"], f["stds"], f["lens"].astype(np.float32), f["signals"])
    ref, pref = eng.forward(*x, precision=prec)
    for r in range(reps):
        p, pp = eng.forward(*x, precision=prec)
        if not (np.array_equal(p, ref) and np.array_equal(pp, pref)):
            bad += 1
            print("MISMATCH n=%d %s rep %d max|d|=%g rows=%d" % (n, prec, r, np.abs(p - ref).max(), int((p != ref).any(axis=1).sum())))
    print("n=%d %s: %d repetitions identical" % (n, prec, reps))
print("mismatching runs:", bad)
sys.exit(1 if bad else 0)

"""A/B check of the CTA-pair (cta_group::2) LSTM GEMM against the single-CTA kernel (DSB_LSTM_2CTA toggles at plan build)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepsignal_b200.engine import Engine
from deepsignal_b200 import weights as W
from deepsignal_b200.synth import synth_features

ck = W.random_checkpoint(17, "stress")
f = synth_features(3000, seed=9)
x = (f["kmer"], f["means"], f["stds"], f["lens"].astype(np.float32), f["signals"])
eng = Engine(0)
os.environ["DSB_LSTM_2CTA"] = "0"
eng.load_weights(ck)
a, pa = eng.forward(*x, precision="bf16")
os.environ["DSB_LSTM_2CTA"] = "1"
eng.load_weights(ck)
b, pb = eng.forward(*x, precision="bf16")
print("pair vs single: max|dp| = %.3e identical=%s flips=%d" % (np.abs(a - b).max(), np.array_equal(a, b), int((pa != pb).sum())))

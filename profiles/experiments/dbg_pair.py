"""Debug aid: CTA-pair LSTM sequence kernel vs the single-CTA one (run each mode in its own process)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepsignal_b200.engine import Engine
from deepsignal_b200 import weights as W
from deepsignal_b200.synth import synth_features

n = int(sys.argv[1]) if len(sys.argv) > 1 else 700
f = synth_features(n, seed=3)
x = (f["kmer"], f["means"], f["stds"], f["lens"].astype(np.float32), f["signals"])
ck = W.random_checkpoint(17, "stress")
eng = Engine(0)
eng.load_weights(ck)
p, _ = eng.forward(*x, precision=sys.argv[2] if len(sys.argv) > 2 else "bf16")
np.save(sys.argv[3], p)
print("done", n, float(p.sum()))

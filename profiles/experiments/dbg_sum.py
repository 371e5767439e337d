import os, sys, numpy as np
sys.path.insert(0, ".")
from deepsignal_b200.engine import Engine
from deepsignal_b200 import weights as W
from deepsignal_b200.synth import synth_features
f = synth_features(99, seed=3)
x = (f["kmer"], f["means"], f["stds"], f["lens"].astype(np.float32), f["signals"])
ck = W.random_checkpoint(17, "stress", is_rnn=False)
eng = Engine(0)
eng.load_weights(ck, is_rnn=False)
eng.forward(*x, precision="bf16")

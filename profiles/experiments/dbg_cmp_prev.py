"""Debug aid: run the same forward through the in-tree libdsb200.so and a previous build (profiles/experiments/_old/
libdsb200_prev.so or $DSB_PREV_LIB, not tracked) in two processes and compare the probabilities bit for bit."""
import os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CODE = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import deepsignal_b200._lib as L
if sys.argv[1] != "cur": L.LIB_PATH = sys.argv[1]
from deepsignal_b200.engine import Engine
from deepsignal_b200 import weights as W
from deepsignal_b200.synth import synth_features
f = synth_features(int(sys.argv[3]), seed=3)
x = (f["kmer"], f["means"], f["stds"], f["lens"].astype(np.float32), f["signals"])
eng = Engine(0); eng.load_weights(W.random_checkpoint(17, "stress"))
p, _ = eng.forward(*x, precision=sys.argv[4]); np.save(sys.argv[2], p)
''' % ROOT
for prec in ("bf16", "fp16"):
    for n in (99, 1500):
        subprocess.check_call([sys.executable, "-c", CODE, "cur", "/tmp/cur.npy", str(n), prec])
        subprocess.check_call([sys.executable, "-c", CODE, os.environ.get("DSB_PREV_LIB", os.path.join(HERE, "_old", "libdsb200_prev.so")), "/tmp/prev.npy", str(n), prec])
        a, b = np.load("/tmp/cur.npy"), np.load("/tmp/prev.npy")
        print(prec, n, "identical to previous build:", np.array_equal(a, b), float(np.abs(a - b).max()))

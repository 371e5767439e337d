"""Debug aid: fused front kernel vs separate launches (run on the GPU box)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepsignal_b200.engine import Engine
from deepsignal_b200 import weights as W
from deepsignal_b200.synth import synth_features

n = int(sys.argv[1]) if len(sys.argv) > 1 else 99
kind = sys.argv[2] if len(sys.argv) > 2 else "stress"
f = synth_features(n, seed=3)
x = (f["kmer"], f["means"], f["stds"], f["lens"].astype(np.float32), f["signals"])
print("signals finite:", np.isfinite(x[4]).all(), "max|signal|", np.abs(x[4]).max())
variant = sys.argv[3] if len(sys.argv) > 3 else "both"
kw = dict(is_cnn=variant != "rnn", is_rnn=variant != "cnn")
ck = W.random_checkpoint(17, kind, **kw)
eng = Engine(0)
variant = sys.argv[3] if len(sys.argv) > 3 else "both"
kw = dict(is_cnn=variant != "rnn", is_rnn=variant != "cnn")
outs = {}
for mode in ("0", "1", "0", "1"):
    os.environ["DSB_FRONT"] = mode
    eng.load_weights(ck, **kw)
    p, _ = eng.forward(*x, precision="bf16")
    outs.setdefault(mode, []).append(p.copy())
from oracle import model_oracle as mo
ref, _ = mo.forward(ck, *x, **kw)
print("vs oracle: mode0", np.abs(outs["0"][0] - ref).max(), "mode1", np.abs(outs["1"][0] - ref).max())
print("0 vs 0:", np.abs(outs["0"][0] - outs["0"][1]).max(), " 1 vs 1:", np.abs(outs["1"][0] - outs["1"][1]).max(),
      " 0 vs 1:", np.abs(outs["0"][0] - outs["1"][0]).max())
bad = np.nonzero(np.abs(outs["0"][0] - outs["1"][0]).max(axis=1) > 0)[0]
print("differing samples:", len(bad), bad[:20])

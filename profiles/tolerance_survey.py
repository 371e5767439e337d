#!/usr/bin/env python
"""max |dp| of the fp16x2 (and fp16) path against the float64 oracle over several random-init 'stress' checkpoints and the
three model variants: how much of the 2e-3 tolerance the default mode uses on checkpoints other than the timed one.

    python profiles/tolerance_survey.py            # on the GPU box; prints one JSON line per case
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsignal_b200 import weights as W          # noqa: E402
from deepsignal_b200.engine import Engine         # noqa: E402
from deepsignal_b200.synth import synth_features  # noqa: E402
from oracle import model_oracle as mo             # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    eng = Engine(0)
    f = synth_features(n, seed=104)
    x = (f["kmer"], f["means"], f["stds"], f["lens"], f["signals"])
    for seed in (17, 23, 31, 47, 59):
        for kw in (dict(), dict(is_base=False), dict(is_rnn=False), dict(is_cnn=False)):
            ck = W.random_checkpoint(seed, "stress", **kw)
            eng.load_weights(ck, **kw)
            ref, _ = mo.forward(ck, *x, **kw)
            out = {"seed": seed, "variant": kw, "n": n}
            for prec in ("fp16x3", "fp16x2", "fp16"):
                got, pred = eng.forward(*x, precision=prec)
                out[prec] = float(np.abs(got - ref).max())
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

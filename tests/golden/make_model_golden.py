#!/usr/bin/env python
"""Generate tests/golden/model_ref.json by executing the REFERENCE's own graph code (deepsignal/model.py + layers.py,
unmodified) over the NumPy stand-in for TensorFlow 1.x in oracle/tf_shim.py.

Run in the build container only (needs the reference sources: /root/reference, or the git-ignored copy oracle/_ref made
by oracle/make_ref.py):

    python tests/golden/make_model_golden.py

What this pins: the wiring of the forward (reference model.py:25-100, layers.py:20-264) -- scopes and variable names,
layer order, strides and paddings, channel counts, the concat order of the five inception branches, which BiLSTM
timesteps feed the joint model, the two bias-free dense layers, the sigmoid.  What it cannot pin: TensorFlow's own
kernels (TF 1.x is not installable here); the TF op semantics restated in oracle/tf_shim.py are listed in its header.
tests/test_model_golden.py checks oracle/model_oracle.py against the vectors written here.
"""
import hashlib
import importlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = [  # (name, checkpoint seed, variant, feature seed, n, Model(...) switches)
    ("stress17", 17, "stress", 101, 6, {}),
    ("tfdefault23", 23, "tf_default", 102, 4, {}),
    ("stress31_no_base", 31, "stress", 103, 3, {"is_base": False}),      # model.py:57-76: features only, no embedding
    ("stress31_no_rnn", 31, "stress", 103, 3, {"is_rnn": False}),        # model.py:88-94: signal model alone
    ("stress31_no_cnn", 31, "stress", 103, 3, {"is_cnn": False}),        # event model alone (the signal model is still built)
]


def reference_dir():
    for cand in (os.environ.get("DEEPSIGNAL_REFERENCE"), "/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "deepsignal", "model.py")):
            return os.path.join(cand, "deepsignal")
    return None


def run_reference_model(ckpt, kmer, means, stds, lens, signals, ref_pkg_dir, **model_kw):
    """-> (activation_logits [n,2] f64, prediction [n], names of the variables the graph code asked for, in order)."""
    from oracle import tf_shim
    saved = tf_shim.install()
    pkg_name = "_dsref_graph"
    try:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [ref_pkg_dir]
        sys.modules[pkg_name] = pkg
        for sub in ("layers", "model"):
            sys.modules.pop(pkg_name + "." + sub, None)
        ref_model = importlib.import_module(pkg_name + ".model")
        n = len(kmer)
        # placeholders in creation order (model.py:31-42): base_int, means, stds, sanums, signals, labels, lr, keep_prob, training
        tf_shim.begin(ckpt, [kmer, means, stds, lens, signals, np.zeros(n, np.int32), 0.001, 1.0, False])
        m = ref_model.Model(base_num=kmer.shape[1], signal_num=signals.shape[1], class_num=2, **model_kw)
        return np.asarray(m.activation_logits, np.float64), np.asarray(m.prediction), tf_shim.variables_used()
    finally:
        for k in [k for k in sys.modules if k == pkg_name or k.startswith(pkg_name + ".")]:
            del sys.modules[k]
        tf_shim.uninstall(saved)


def case_inputs(seed, n):
    from deepsignal_b200.synth import synth_features
    f = synth_features(n, seed=seed)
    return f["kmer"], f["means"], f["stds"], f["lens"], f["signals"]


def main():
    from deepsignal_b200 import weights as W
    ref = reference_dir()
    if ref is None:
        raise SystemExit("reference sources not found (set DEEPSIGNAL_REFERENCE)")
    out = {"generator": "tests/golden/make_model_golden.py", "cases": {}}
    for name, ck_seed, variant, f_seed, n, kw in CASES:
        ck = W.random_checkpoint(ck_seed, variant, **kw)
        kmer, means, stds, lens, signals = case_inputs(f_seed, n)
        act, pred, used = run_reference_model(ck, kmer, means, stds, lens, signals, ref, **kw)
        out["cases"][name] = {
            "checkpoint": [ck_seed, variant], "model_kw": kw, "features_seed": f_seed, "n": n,
            "activation_logits": [[repr(float(v)) for v in row] for row in act],
            "prediction": [int(v) for v in pred],
            "n_variables": len(used),
            "variables_sha1": hashlib.sha1("\n".join(used).encode()).hexdigest(),
            "first_variables": used[:4], "last_variables": used[-2:],
        }
        print(name, "n_variables", len(used), "act[0]", act[0])
    with open(os.path.join(HERE, "model_ref.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()

"""Checkpoint weight names / shapes of the call_mods model and synthetic checkpoints.

The names are the TF-1.x variable names the reference graph creates with
``model_prefix="model"`` (reference: deepsignal/model.py:43,47,61 and
deepsignal/layers.py:66-71,75-84,87-139,181-235,258-262).  A "checkpoint" here
is a plain ``dict[str, np.ndarray(float32)]`` keyed by those names; it can be
stored as ``.npz`` (``save_npz`` / ``load_checkpoint``).  The library folds the
tensors itself (``dsb_load_weights``), so what crosses the C ABI is exactly what a
TF ``Saver`` bundle holds.
"""
from __future__ import annotations

import numpy as np

VOCAB_SIZE = 1024       # reference model.py:19
EMBEDDING_SIZE = 128    # reference model.py:20
HIDDEN = 256            # reference model.py:51
LAYER_NUM = 3           # reference model.py:50
BN_EPS = 1e-3           # tf.contrib.layers.batch_norm default epsilon (layers.py:81)
INCEPTION_BLOCKS = 11
TIMES = 16              # reference layers.py:87 (times=16)

# (sub-scope, conv name, bn name, kernel width, Cin ("in" = block input), Cout)
INCEPTION_CONVS = [
    ("branch1_maxpooling", "conv1a_1x1", "bn", 1, "in", 48),
    ("branch2_1x1", "conv0b_1x1", "bn", 1, "in", 48),
    ("branch3_1x3", "conv0c_1x1", "bn1", 1, "in", 32),
    ("branch3_1x3", "conv1c_1x3", "bn2", 3, 32, 48),
    ("branch4_1x5", "conv0d_1x1", "bn1", 1, "in", 32),
    ("branch4_1x5", "conv1d_1x5", "bn2", 5, 32, 48),
    ("branch5_residual_1x3", "convstem_1x1", "bn0", 1, "in", 48),
    ("branch5_residual_1x3", "conv0e_1x1", "bn1", 1, "in", 32),
    ("branch5_residual_1x3", "conv1e_1x3", "bn2", 3, 32, 64),
    ("branch5_residual_1x3", "conv2e_1x1", "bn3", 1, 64, 48),
]
STEM_CONVS = [  # (scope, k, Cin, Cout)
    ("conv_layer1", 7, 1, 64),
    ("conv_layer2", 1, 64, 128),
    ("conv_layer3", 3, 128, 256),
]
BN_FIELDS = ("gamma", "beta", "moving_mean", "moving_variance")


def conv_out_len(n: int, stride: int) -> int:
    return -(-n // stride)


def cnn_lengths(signal_num: int = 360):
    """Positions after each stage: conv1 s2, pool s2, (3 blocks), pool s2, (5), pool s2, (3)."""
    l1 = conv_out_len(signal_num, 2)
    l2 = conv_out_len(l1, 2)
    l3 = conv_out_len(l2, 2)
    l4 = conv_out_len(l3, 2)
    return l1, l2, l3, l4


def inception_scope(i: int, prefix: str = "model") -> str:
    s = prefix + "signalm"
    return "{}incp_layer{}/{}{}".format(s, i, s, i)


def variable_specs(base_num: int = 17, signal_num: int = 360, class_num: int = 2,
                   is_cnn: bool = True, is_rnn: bool = True, is_base: bool = True,
                   prefix: str = "model"):
    """Ordered ``[(name, shape)]`` of every inference-time variable.

    The signal model's variables exist even with ``is_cnn=False`` (the reference
    builds it unconditionally, model.py:83-85); they are simply unused then.
    """
    specs = []
    if is_rnn:
        if is_base:
            specs.append((prefix + "embedding", (VOCAB_SIZE, EMBEDDING_SIZE)))
        in0 = (EMBEDDING_SIZE if is_base else 0) + 3
        for d in ("fw", "bw"):
            for l in range(LAYER_NUM):
                cin = in0 if l == 0 else HIDDEN
                base = "{}em/{}/multi_rnn_cell/cell_{}/lstm_cell/".format(prefix, d, l)
                specs.append((base + "kernel", (cin + HIDDEN, 4 * HIDDEN)))
                specs.append((base + "bias", (4 * HIDDEN,)))
    s = prefix + "signalm"
    for scope, k, cin, cout in STEM_CONVS:
        specs.append(("{}{}/conv/kernel".format(s, scope), (1, k, cin, cout)))
        for f in BN_FIELDS:
            specs.append(("{}{}/bn/{}".format(s, scope, f), (cout,)))
    for i in range(1, INCEPTION_BLOCKS + 1):
        cin_block = 256 if i == 1 else 240
        sc = inception_scope(i, prefix)
        for sub, conv, bn, k, cin, cout in INCEPTION_CONVS:
            cin = cin_block if cin == "in" else cin
            specs.append(("{}{}/{}/kernel".format(sc, sub, conv), (1, k, cin, cout)))
            for f in BN_FIELDS:
                specs.append(("{}{}/{}/{}".format(sc, sub, bn, f), (cout,)))
    joint = 0
    if is_rnn:
        joint += 2 * HIDDEN
    if is_cnn:
        joint += cnn_lengths(signal_num)[3] * 240
    specs.append(("dense/kernel", (joint, joint)))
    specs.append(("dense_1/kernel", (joint, class_num)))
    return specs


def _glorot(rng, shape, fan_in, fan_out, gain=1.0):
    lim = gain * np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def random_checkpoint(seed: int = 17, variant: str = "stress", **model_kw):
    """Synthetic "random-init checkpoint" (SURVEY §8d cfg-1).

    variant "tf_default": TF-1.x default initialisers — glorot-uniform conv / dense /
    LSTM kernels, zero LSTM bias, truncated-normal(σ=√(2/1024)) embedding
    (model.py:62), BN γ=1 β=0 μ=0 σ²=1.
    variant "stress": He-scaled conv kernels, non-trivial BN statistics, small random
    LSTM bias and a rescaled final FC so logits spread over a few units instead of
    sitting at sigmoid≈0.5 (needed for a meaningful call-flip test).
    """
    if variant not in ("tf_default", "stress"):
        raise ValueError("variant must be tf_default or stress")
    rng = np.random.default_rng(seed)
    ck = {}
    for name, shape in variable_specs(**model_kw):
        leaf = name.rsplit("/", 1)[-1]
        if name.endswith("embedding"):
            std = np.sqrt(2.0 / VOCAB_SIZE)
            w = rng.normal(0.0, std, size=shape)
            bad = np.abs(w) > 2 * std  # truncated normal: redraw beyond 2 sigma
            while bad.any():
                w[bad] = rng.normal(0.0, std, size=int(bad.sum()))
                bad = np.abs(w) > 2 * std
            if variant == "stress":
                w *= 8.0
            ck[name] = w.astype(np.float32)
        elif leaf == "kernel" and len(shape) == 4:
            _, k, cin, cout = shape
            gain = np.sqrt(2.0) * 1.05 if variant == "stress" else 1.0
            if variant == "stress":
                ck[name] = (rng.normal(0, 1, size=shape) * gain / np.sqrt(k * cin)).astype(np.float32)
            else:
                ck[name] = _glorot(rng, shape, k * cin, k * cout)
        elif leaf == "kernel":
            gain = 1.0
            if variant == "stress" and name == "dense_1/kernel":
                gain = 0.05
            ck[name] = _glorot(rng, shape, shape[0], shape[1], gain)
        elif leaf == "bias":
            ck[name] = (rng.normal(0, 0.1, size=shape).astype(np.float32) if variant == "stress"
                        else np.zeros(shape, np.float32))
        elif leaf == "gamma":
            ck[name] = (rng.uniform(0.5, 1.5, size=shape).astype(np.float32) if variant == "stress"
                        else np.ones(shape, np.float32))
        elif leaf == "beta":
            ck[name] = (rng.normal(0, 0.1, size=shape).astype(np.float32) if variant == "stress"
                        else np.zeros(shape, np.float32))
        elif leaf == "moving_mean":
            ck[name] = (rng.normal(0, 0.1, size=shape).astype(np.float32) if variant == "stress"
                        else np.zeros(shape, np.float32))
        elif leaf == "moving_variance":
            ck[name] = (rng.uniform(0.5, 1.5, size=shape).astype(np.float32) if variant == "stress"
                        else np.ones(shape, np.float32))
        else:  # pragma: no cover
            raise AssertionError(name)
    return ck


def save_npz(path: str, ckpt: dict) -> None:
    np.savez(path, **{k.replace("/", "|"): v for k, v in ckpt.items()})


def load_npz(path: str) -> dict:
    with np.load(path) as z:
        return {k.replace("|", "/"): np.asarray(z[k], dtype=np.float32) for k in z.files}


def load_checkpoint(model_path: str) -> dict:
    """Load a checkpoint by the path given to ``--model_path``.

    Accepts an ``.npz`` written by :func:`save_npz`, or a TF-V2 tensor bundle prefix
    (``….ckpt`` with ``.index`` / ``.data-00000-of-00001``; restore site in the
    reference: call_modifications.py:210-211) read by :mod:`deepsignal_b200.tf_bundle`.
    """
    import os
    if model_path.endswith(".npz") and os.path.exists(model_path):
        return load_npz(model_path)
    if os.path.exists(model_path + ".npz"):
        return load_npz(model_path + ".npz")
    if os.path.exists(model_path + ".index"):
        from .tf_bundle import read_bundle
        return {k: np.asarray(v, np.float32) for k, v in read_bundle(model_path).items()
                if v.dtype == np.float32}
    raise FileNotFoundError("no checkpoint at {} (.npz or TF bundle)".format(model_path))

"""ORACLE (test infrastructure, never shipped or measured as the product).

CPU restatement of the reference's call_mods forward graph — the *un-folded*
TensorFlow-1.x graph of deepsignal/model.py:25-108 + deepsignal/layers.py:20-264 —
in PyTorch on the CPU (float64 for parity checks, float32 for the CPU baseline leg
of bench.py).  TensorFlow itself is a third-party dependency that is absent from
/root/reference and cannot be installed here (tensorflow 1.8–1.13.1,
README.md:28), so the TF op semantics below are restated from its published
behaviour and every rule is unit-tested against a hand-computed case in
tests/test_model_oracle.py:

* ``padding="SAME"``: out = ceil(in/s); pad_total = max((out-1)*s + k - in, 0);
  pad_left = pad_total // 2, the remainder goes right (tf.layers.conv2d /
  max_pooling2d / average_pooling2d; call sites layers.py:91-135,184-235).
* conv kernels are HWIO ``[1,k,Cin,Cout]``, cross-correlation, no bias.
* max-pool pads with -inf; avg-pool divides by the number of valid cells.
* ``tf.contrib.layers.batch_norm`` inference: γ(x-μ)/sqrt(σ²+0.001)+β (layers.py:80-84).
* ``LSTMCell``: z=[x,h]·K+b ; i,j,f,o = split(z,4) ; c'=σ(f+1)·c+σ(i)·tanh(j) ;
  h'=σ(o)·tanh(c') (forget_bias=1, no peephole / projection; layers.py:49-50).
* ``bidirectional_dynamic_rnn`` over two independent 3-layer stacks, zero initial
  state, full-length sequences; output = concat(fw_top[:, -1], bw_top[:, 0])
  (layers.py:66-71,171-172).
* flatten of [B,1,23,240] is position-major (layers.py:237-238); joint =
  concat(event, signal) -> dense (no bias) -> dense_1 (no bias) (layers.py:250-264);
  probs = sigmoid(logits); prediction = argmax (ties -> 0) (model.py:100,108).

PARITY PIN STATUS: the reference ships no golden vectors / known-answer tests for
this path and TF cannot run here.  The WIRING of this restatement is
pinned to the reference's own source: tests/golden/make_model_golden.py executes
deepsignal/model.py + layers.py UNMODIFIED over a NumPy stand-in for the ~35 TF 1.x
calls they make (oracle/tf_shim.py, written independently of this file), and
tests/test_model_golden.py requires this oracle to reproduce those outputs to 1e-10
on five cases (two full-model checkpoints and the is_base / is_rnn / is_cnn = False
variants), with exactly the variable names the reference's scopes generate.  The TF op semantics themselves (the bullets above) remain restated and
hand-checked only: **parity unpinned against real TF kernels**.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from deepsignal_b200 import weights as W


def same_pads(n: int, k: int, s: int):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv_same(x, kernel_hwio, stride=1):
    """x [B,C,L]; kernel [1,k,Cin,Cout] (TF HWIO) -> [B,Cout,ceil(L/s)]."""
    k = kernel_hwio.shape[1]
    w = kernel_hwio[0].permute(2, 1, 0).contiguous()  # [Cout,Cin,k]
    pl, pr = same_pads(x.shape[-1], k, stride)
    return F.conv1d(F.pad(x, (pl, pr)), w, stride=stride)


def maxpool_same(x, k, s):
    pl, pr = same_pads(x.shape[-1], k, s)
    return F.max_pool1d(F.pad(x, (pl, pr), value=float("-inf")), k, s)


def avgpool_same(x, k, s=1):
    pl, pr = same_pads(x.shape[-1], k, s)
    num = F.avg_pool1d(F.pad(x, (pl, pr)), k, s) * k
    ones = torch.ones(1, 1, x.shape[-1], dtype=x.dtype)
    cnt = F.avg_pool1d(F.pad(ones, (pl, pr)), k, s) * k
    return num / cnt


def batch_norm(x, ck, scope):
    g, b, m, v = (ck[scope + "/" + f] for f in W.BN_FIELDS)
    inv = g / torch.sqrt(v + W.BN_EPS)
    return (x - m[None, :, None]) * inv[None, :, None] + b[None, :, None]


def lstm_cell(x, h, c, kernel, bias):
    z = torch.cat([x, h], dim=1) @ kernel + bias
    i, j, f, o = torch.split(z, W.HIDDEN, dim=1)
    c2 = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def lstm_stack(x, ck, direction, prefix="model"):
    """x [B,T,in] in the order the stack consumes it -> top-layer outputs [B,T,256]."""
    B, T, _ = x.shape
    seq = x
    for l in range(W.LAYER_NUM):
        base = "{}em/{}/multi_rnn_cell/cell_{}/lstm_cell/".format(prefix, direction, l)
        K, b = ck[base + "kernel"], ck[base + "bias"]
        h = torch.zeros(B, W.HIDDEN, dtype=x.dtype)
        c = torch.zeros(B, W.HIDDEN, dtype=x.dtype)
        outs = []
        for t in range(T):
            h, c = lstm_cell(seq[:, t], h, c, K, b)
            outs.append(h)
        seq = torch.stack(outs, dim=1)
    return seq


def event_model(base_int, means, stds, sanums, ck, is_base=True, prefix="model"):
    feats = [means[..., None], stds[..., None], sanums[..., None]]
    if is_base:
        feats = [ck[prefix + "embedding"][base_int.long()]] + feats
    x = torch.cat(feats, dim=2)                      # [B,17,131]
    fw = lstm_stack(x, ck, "fw", prefix)
    bw = torch.flip(lstm_stack(torch.flip(x, dims=[1]), ck, "bw", prefix), dims=[1])
    return torch.cat([fw[:, -1, :], bw[:, 0, :]], dim=1)


def inception_block(x, ck, i, prefix="model"):
    sc = W.inception_scope(i, prefix)

    def cbr(inp, sub, conv, bn, relu=True):
        y = batch_norm(conv_same(inp, ck["{}{}/{}/kernel".format(sc, sub, conv)]), ck,
                       "{}{}/{}".format(sc, sub, bn))
        return F.relu(y) if relu else y

    b1 = cbr(maxpool_same(x, 3, 1), "branch1_maxpooling", "conv1a_1x1", "bn")
    b2 = cbr(x, "branch2_1x1", "conv0b_1x1", "bn")
    b3 = cbr(cbr(x, "branch3_1x3", "conv0c_1x1", "bn1"), "branch3_1x3", "conv1c_1x3", "bn2")
    b4 = cbr(cbr(x, "branch4_1x5", "conv0d_1x1", "bn1"), "branch4_1x5", "conv1d_1x5", "bn2")
    r = "branch5_residual_1x3"
    stem = cbr(x, r, "convstem_1x1", "bn0", relu=False)
    e = cbr(cbr(cbr(x, r, "conv0e_1x1", "bn1"), r, "conv1e_1x3", "bn2"), r, "conv2e_1x1", "bn3", relu=False)
    b5 = F.relu(stem + e)
    return torch.cat([b1, b2, b3, b4, b5], dim=1)


def signal_model(signals, ck, prefix="model", return_stages=False):
    s = prefix + "signalm"
    x = signals[:, None, :]                          # [B,1,360]
    stages = {}
    x = F.relu(batch_norm(conv_same(x, ck[s + "conv_layer1/conv/kernel"], 2), ck, s + "conv_layer1/bn"))
    stages["conv1"] = x
    x = maxpool_same(x, 3, 2)
    stages["pool1"] = x
    x = F.relu(batch_norm(conv_same(x, ck[s + "conv_layer2/conv/kernel"]), ck, s + "conv_layer2/bn"))
    stages["conv2"] = x
    x = F.relu(batch_norm(conv_same(x, ck[s + "conv_layer3/conv/kernel"]), ck, s + "conv_layer3/bn"))
    stages["conv3"] = x
    for i in range(1, 12):
        x = inception_block(x, ck, i, prefix)
        stages["incp%d" % i] = x
        if i in (3, 8):
            x = maxpool_same(x, 3, 2)
            stages["pool_after%d" % i] = x
    x = avgpool_same(x, 7, 1)
    stages["avgpool"] = x
    flat = x.permute(0, 2, 1).reshape(x.shape[0], -1)   # position-major, channel-minor
    return (flat, stages) if return_stages else flat


def forward(ckpt: dict, kmer, means, stds, sanums, signals, dtype=torch.float64,
            is_cnn=True, is_rnn=True, is_base=True, prefix="model", return_logits=False):
    """Returns (activation_logits [B,2] = sigmoid(logits), prediction [B] int64)."""
    if not (is_cnn or is_rnn):
        raise ValueError("at least one of is_cnn/is_rnn should be True")
    ck = {k: torch.as_tensor(np.asarray(v), dtype=dtype) for k, v in ckpt.items()}
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).to(dtype)  # placeholders are float32
    parts = []
    with torch.no_grad():
        if is_rnn:
            parts.append(event_model(torch.as_tensor(np.asarray(kmer)), t(means), t(stds), t(sanums),
                                     ck, is_base, prefix))
        if is_cnn:
            parts.append(signal_model(t(signals), ck, prefix))
        joint = torch.cat(parts, dim=1)
        logits = (joint @ ck["dense/kernel"]) @ ck["dense_1/kernel"]
        probs = torch.sigmoid(logits)
        pred = torch.argmax(probs, dim=1)
    if return_logits:
        return probs.numpy(), pred.numpy(), logits.numpy()
    return probs.numpy(), pred.numpy()


def host_normalise(probs):
    """Reference call_modifications.py:185-187: p/(p0+p1) in the sess.run dtype (float32)."""
    p = np.asarray(probs, dtype=np.float32)
    s = p[:, 0] + p[:, 1]
    return p[:, 0] / s, p[:, 1] / s

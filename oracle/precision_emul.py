"""ORACLE-SIDE TOOL (test infrastructure, never shipped or measured as the product).

Error-budget emulator for the 16-bit tensor-core path: the float64 restatement of the
call_mods forward (reference deepsignal/model.py:57-108, deepsignal/layers.py:87-264) with the
BatchNorm folded into the conv weights exactly as csrc/fold.cu does, plus a switchable
rounding to fp16 / bf16 at every point where the CUDA path stores a 16-bit value:

  w_front   conv2 / conv3 weights          a_front   P1, C2, X0 activations
  w_blk     inception weights              a_planes  P3 / P4 / P5 / stem planes
  w_lstm    LSTM kernels + embedding LUT   a_u       b5b output (U)
  a_feat    mean / std / len feature rows  a_y       block outputs (optionally only some blocks)
  a_h       h_t of every LSTM step         a_b1in    the pooled block input that feeds branch 1 (not in ALL_SITES: a no-op
                                                     when a_y already rounds the block input; study site of a rejected
                                                     variant of the split mode, profiles/r2_ablation.md)

`split_*` names in the site set select hi+lo operands (error ~2^-22) for that class instead of
a single rounding.  Used by tests/test_precision_budget.py and to decide where the kernels keep
extra precision.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from deepsignal_b200 import weights as W
from oracle import model_oracle as MO

ALL_SITES = ("w_front", "w_blk", "w_lstm", "a_front", "a_planes", "a_u", "a_y", "a_feat", "a_h")


def make_q(kind: str):
    """fp16 / bf16: one rounding.  fp16+lo8: hi = fp16(x) plus lo = e5m2(fp16(x - hi)) — the 24-bit activation format
    (fp16 high part + the high byte of the fp16 low part) of the split-operand mode's HBM layout."""
    if kind == "fp16+lo8":
        def q(x):
            hi = x.to(torch.float16).to(torch.float64)
            lo = (x - hi).to(torch.float16).to(torch.float8_e5m2).to(torch.float64)
            return hi + lo
        return q
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[kind]

    def q(x):
        return x.to(dt).to(torch.float64)
    return q


class Emul:
    def __init__(self, ckpt: dict, kind: str = "fp16", sites=ALL_SITES, y_blocks=None, prefix="model", act_kind=None):
        self.ck = {k: torch.as_tensor(np.asarray(v), dtype=torch.float64) for k, v in ckpt.items()}
        self.q = make_q(kind)
        self.q_act = make_q(act_kind) if act_kind else self.q      # rounding of the a_* (activation) sites
        self.sites = set(sites)
        self.y_blocks = y_blocks          # None = all blocks round (w_blk, a_planes, a_u, a_y); else only these block numbers
        self.cur_block = None
        self.prefix = prefix

    def r(self, site, x):
        if site not in self.sites:
            return x
        if self.cur_block is not None and self.y_blocks is not None and self.cur_block not in self.y_blocks:
            return x
        return self.q_act(x) if site.startswith("a_") else self.q(x)

    def folded_conv(self, conv, bn):
        ck = self.ck
        g, b, m, v = (ck[bn + "/" + f] for f in W.BN_FIELDS)
        scale = g / torch.sqrt(v + W.BN_EPS)
        # fold.cu keeps the folded weights / bias as float32 before the 16-bit rounding
        k = (ck[conv + "/kernel"] * scale[None, None, None, :]).to(torch.float32).to(torch.float64)
        bias = (b - m * scale).to(torch.float32).to(torch.float64)
        return k, bias

    def cb(self, x, conv, bn, wsite, relu=True, stride=1):
        k, bias = self.folded_conv(conv, bn)
        y = MO.conv_same(x, self.r(wsite, k), stride) + bias[None, :, None]
        return F.relu(y) if relu else y

    def block(self, x, i):
        sc = W.inception_scope(i, self.prefix)
        self.cur_block = i
        c = lambda inp, sub, conv, bn, relu=True: self.cb(inp, sc + sub + "/" + conv, sc + sub + "/" + bn, "w_blk", relu)
        r5 = "branch5_residual_1x3"
        b1 = c(self.r("a_b1in", MO.maxpool_same(x, 3, 1)), "branch1_maxpooling", "conv1a_1x1", "bn")
        b2 = c(x, "branch2_1x1", "conv0b_1x1", "bn")
        p3 = self.r("a_planes", c(x, "branch3_1x3", "conv0c_1x1", "bn1"))
        p4 = self.r("a_planes", c(x, "branch4_1x5", "conv0d_1x1", "bn1"))
        p5 = self.r("a_planes", c(x, r5, "conv0e_1x1", "bn1"))
        stem = self.r("a_planes", c(x, r5, "convstem_1x1", "bn0", relu=False))
        b3 = c(p3, "branch3_1x3", "conv1c_1x3", "bn2")
        b4 = c(p4, "branch4_1x5", "conv1d_1x5", "bn2")
        u = self.r("a_u", c(p5, r5, "conv1e_1x3", "bn2"))
        e = c(u, r5, "conv2e_1x1", "bn3", relu=False)
        b5 = F.relu(stem + e)
        y = self.r("a_y", torch.cat([b1, b2, b3, b4, b5], dim=1))
        self.cur_block = None
        return y

    def signal_model(self, signals):
        s = self.prefix + "signalm"
        x = signals[:, None, :]
        x = self.cb(x, s + "conv_layer1/conv", s + "conv_layer1/bn", "none", stride=2)   # conv1 runs on split parts: exact
        x = self.r("a_front", MO.maxpool_same(x, 3, 2))
        x = self.r("a_front", self.cb(x, s + "conv_layer2/conv", s + "conv_layer2/bn", "w_front"))
        x = self.r("a_front", self.cb(x, s + "conv_layer3/conv", s + "conv_layer3/bn", "w_front"))
        for i in range(1, 12):
            x = self.block(x, i)
            if i in (3, 8):
                x = MO.maxpool_same(x, 3, 2)
        x = MO.avgpool_same(x, 7, 1)
        return x.permute(0, 2, 1).reshape(x.shape[0], -1)

    def lstm_stack(self, x_feat, code, direction):
        """x_feat [B,T,3] (mean,std,len), code [B,T] in consumption order."""
        ck = self.ck
        B, T, _ = x_feat.shape
        seq = None
        for l in range(W.LAYER_NUM):
            base = "{}em/{}/multi_rnn_cell/cell_{}/lstm_cell/".format(self.prefix, direction, l)
            K, b = ck[base + "kernel"], ck[base + "bias"]
            if l == 0:
                emb = ck[self.prefix + "embedding"]
                lut = (emb[:5] @ K[:128]).to(torch.float32).to(torch.float64)     # fold.cu: 5 LUT rows
                lut = self.r("w_lstm", lut)
                Kx = self.r("w_lstm", K[128:131])
                Kh = self.r("w_lstm", K[131:])
                xin = lut[code.long()] + self.r("a_feat", x_feat) @ Kx            # [B,T,1024]
            else:
                Kx = self.r("w_lstm", K[:256])
                Kh = self.r("w_lstm", K[256:])
                xin = seq @ Kx
            h = torch.zeros(B, W.HIDDEN, dtype=torch.float64)
            c = torch.zeros(B, W.HIDDEN, dtype=torch.float64)
            outs = []
            for t in range(T):
                z = xin[:, t] + h @ Kh + b
                i, j, f, o = torch.split(z, W.HIDDEN, dim=1)
                c = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
                h = self.r("a_h", torch.sigmoid(o) * torch.tanh(c))
                outs.append(h)
            seq = torch.stack(outs, dim=1)
        return seq

    def event_model(self, kmer, means, stds, lens):
        x = torch.stack([means, stds, lens], dim=2)
        fw = self.lstm_stack(x, kmer, "fw")
        bw = self.lstm_stack(torch.flip(x, dims=[1]), torch.flip(kmer, dims=[1]), "bw")
        return torch.cat([fw[:, -1, :], bw[:, -1, :]], dim=1)

    def forward(self, kmer, means, stds, lens, signals, parts=("rnn", "cnn")):
        t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).to(torch.float64)
        with torch.no_grad():
            ev = self.event_model(torch.as_tensor(np.asarray(kmer)), t(means), t(stds), t(lens))
            sg = self.signal_model(t(signals))
            joint = torch.cat([ev, sg], dim=1)
            logits = (joint @ self.ck["dense/kernel"]) @ self.ck["dense_1/kernel"]
            return torch.sigmoid(logits).numpy(), logits.numpy()


def report(p_ref, p, label=""):
    d = np.abs(p - p_ref)
    flips = int(((p[:, 1] > p[:, 0]) != (p_ref[:, 1] > p_ref[:, 0])).sum())
    return dict(label=label, max=float(d.max()), mean=float(d.mean()), rms=float(np.sqrt((d ** 2).mean())), flips=flips, n=len(p))

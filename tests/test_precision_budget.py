"""CPU checks of oracle/precision_emul.py, the error-budget emulator behind the choice of the split-operand (fp16x2)
tensor-core mode: with no rounding site switched on it must reproduce the float64 oracle, and the
ordering bf16 > fp16 > (fp16 weights + split activations) that the GPU measurements show must come out of it."""
import numpy as np

from deepsignal_b200 import weights as W
from deepsignal_b200.synth import synth_features
from oracle import model_oracle as mo
from oracle import precision_emul as pe


def _args(n, seed):
    f = synth_features(n, seed=seed)
    return f["kmer"], f["means"], f["stds"], f["lens"].astype(np.float32), f["signals"]


def test_emulator_without_rounding_equals_the_oracle():
    ck = W.random_checkpoint(17, "stress")
    x = _args(48, 7)
    ref, _ = mo.forward(ck, *x)
    got, _ = pe.Emul(ck, "fp16", sites=()).forward(*x)
    assert np.abs(got - ref).max() <= 1e-6          # the BN fold is kept in float32 like csrc/fold.cu: not bit-identical


def test_error_ordering_of_the_modes():
    ck = W.random_checkpoint(17, "stress")
    x = _args(96, 8)
    ref, _ = pe.Emul(ck, "fp16", sites=()).forward(*x)
    weights_only = ("w_front", "w_blk", "w_lstm", "a_h", "a_feat")
    err = {}
    err["bf16"] = pe.report(ref, pe.Emul(ck, "bf16").forward(*x)[0])["rms"]
    err["fp16"] = pe.report(ref, pe.Emul(ck, "fp16").forward(*x)[0])["rms"]
    err["fp16x2"] = pe.report(ref, pe.Emul(ck, "fp16", sites=weights_only).forward(*x)[0])["rms"]
    lo8 = pe.Emul(ck, "fp16", sites=("w_front", "w_blk", "w_lstm", "a_front", "a_planes", "a_u", "a_y"), act_kind="fp16+lo8")
    err["fp16+lo8"] = pe.report(ref, lo8.forward(*x)[0])["rms"]
    print(err)
    assert err["bf16"] > 4 * err["fp16"] > 4 * 1.2 * err["fp16x2"] > 0
    assert 1e-4 < err["fp16"] < 1e-3                # the measured fp16 figure is rms 3.9e-4
    assert abs(err["fp16+lo8"] - err["fp16x2"]) <= 0.1 * err["fp16x2"]   # 3 significant bits of lo are enough


def test_fp16_lo8_rounding_is_a_24_bit_format():
    q = pe.make_q("fp16+lo8")
    import torch
    v = torch.tensor([1.0 + 2.0 ** -13, 3.14159265, 1e-3, 0.0, 100.123], dtype=torch.float64)
    r = q(v)
    # 11 bits of hi + 3 significant bits of lo; below |v| ~ 0.03 the lo part is an e5m2 subnormal: absolute floor 2^-17
    assert bool(((r - v).abs() <= torch.maximum(v.abs() * 2.0 ** -13, torch.tensor(2.0 ** -17, dtype=torch.float64))).all())
    assert float(q(torch.tensor([0.5], dtype=torch.float64))[0]) == 0.5

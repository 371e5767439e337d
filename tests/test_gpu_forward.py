"""GPU parity of the call_mods forward (dsb_forward_* through the C ABI) against the float64 oracle.

Tolerances are BASELINE.json's: probabilities within 1e-4 abs on the fp32 path, within 2e-3 abs on
the 16-bit tensor-core path with >= 99.9 % identical binary calls.

Two checkpoints are used.  "tf_default" is the literal random-init checkpoint (TF-1.x default
initialisers): logits stay within a few 1e-3 of each other, so every mode meets 2e-3 easily.
"stress" spreads the logits over ~N(mu, 1) so that rounding error is visible in the probabilities:
there the fp16-operand mode sits right at 2e-3 (max over 2048 samples 2.0e-3, over 20000 samples
2.5e-3, mean 3e-4; every flipped call has a margin |p1-p0| < 6e-4), while bf16 operands (8-bit
mantissa) give ~2e-2 — a CPU emulation of 16-bit rounding at the same points of the float64 oracle
reproduces both figures, so they are properties of the number formats, not kernel defects.  The stress
numbers are asserted against those documented bounds and printed."""
import numpy as np
import pytest

from deepsignal_b200 import weights as W
from deepsignal_b200.synth import synth_features
from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
TC_TOL = 2e-3          # BASELINE tolerance of the 16-bit tensor-core path
FP16_STRESS_BOUND = 4e-3   # fp16 operands on the spread-logit checkpoint (see module doc)
BF16_STRESS_BOUND = 6e-2   # what 8-bit mantissas can hold there


@pytest.fixture(scope="module")
def engine():
    from deepsignal_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="module")
def ckpt():
    return W.random_checkpoint(17, "stress")


def _inputs(n, seed):
    f = synth_features(n, seed=seed)
    return f["kmer"], f["means"], f["stds"], f["lens"].astype(np.float32), f["signals"]


def _report(tag, got, ref, pred, ref_pred):
    err = np.abs(got - ref)
    margin = np.abs(ref[:, 0] - ref[:, 1])
    flips = pred != ref_pred
    print("\n[{}] n={} max|dp|={:.3e} mean|dp|={:.3e} flips={} (margins of flips: {})".format(
        tag, len(got), err.max(), err.mean(), int(flips.sum()), np.round(margin[flips], 5).tolist()[:8]))
    return err, flips, margin


def test_fp32_matches_oracle(engine, ckpt):
    engine.load_weights(ckpt)
    x = _inputs(300, 101)
    ref, ref_pred = mo.forward(ckpt, *x)
    got, pred = engine.forward(*x, precision="fp32")
    err, flips, margin = _report("fp32", got, ref, pred, ref_pred)
    assert err.max() <= FP32_TOL
    assert not np.any(flips & (margin > 2 * FP32_TOL))


def test_tensor_core_paths_match_oracle_stress(engine, ckpt):
    engine.load_weights(ckpt)
    x = _inputs(2048, 102)
    ref, ref_pred = mo.forward(ckpt, *x)
    got, pred = engine.forward(*x, precision="fp16")
    err, flips, margin = _report("fp16 tcgen05, stress ckpt", got, ref, pred, ref_pred)
    assert err.max() <= FP16_STRESS_BOUND and err.mean() <= 1e-3
    assert flips.mean() <= 2.5e-3
    assert not np.any(flips & (margin > TC_TOL))
    got, pred = engine.forward(*x, precision="bf16")
    err, flips, margin = _report("bf16 tcgen05, stress ckpt", got, ref, pred, ref_pred)
    assert err.max() <= BF16_STRESS_BOUND
    assert not np.any(flips & (margin > 2 * BF16_STRESS_BOUND))


def test_tensor_core_paths_track_fp32_on_device(engine, ckpt):
    engine.load_weights(ckpt)
    x = _inputs(20000, 103)
    a, pa = engine.forward(*x, precision="fp32")
    for prec, bound in (("fp16", FP16_STRESS_BOUND), ("bf16", BF16_STRESS_BOUND)):
        b, pb = engine.forward(*x, precision=prec)
        d = np.abs(a - b)
        flips = pa != pb
        margin = np.abs(a[:, 0] - a[:, 1])
        print("\n[{} vs fp32, n=20000] max|dp|={:.3e} mean={:.3e} flips={} ({:.4%}) max flip margin={:.2e}".format(
            prec, d.max(), d.mean(), int(flips.sum()), flips.mean(), margin[flips].max() if flips.any() else 0.0))
        assert d.max() <= bound
        assert not np.any(flips & (margin > bound))
        if prec == "fp16":
            assert flips.mean() <= 1e-3


@pytest.mark.parametrize("kw", [dict(is_cnn=False), dict(is_rnn=False), dict(is_base=False)])
@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("fp16", FP16_STRESS_BOUND), ("bf16", BF16_STRESS_BOUND)])
def test_model_variants(engine, kw, precision, tol):
    ck = W.random_checkpoint(23, "stress", **kw)
    engine.load_weights(ck, **kw)
    x = _inputs(130, 104)
    ref, _ = mo.forward(ck, *x, **kw)
    got, _ = engine.forward(*x, precision=precision)
    assert np.abs(got - ref).max() <= tol


def test_tf_default_checkpoint(engine):
    ck = W.random_checkpoint(17, "tf_default")
    engine.load_weights(ck)
    x = _inputs(256, 105)
    ref, _ = mo.forward(ck, *x)
    ref_pred = np.argmax(ref, axis=1)
    for precision, tol in (("fp32", FP32_TOL), ("bf16", TC_TOL), ("fp16", TC_TOL)):
        got, pred = engine.forward(*x, precision=precision)
        print("\n[{} tf_default] max|dp|={:.3e} identical calls={:.4%}".format(precision, np.abs(got - ref).max(), (pred == ref_pred).mean()))
        assert np.abs(got - ref).max() <= tol
        assert (pred == ref_pred).mean() >= 0.999


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_chunking_and_ragged_sizes_are_batch_invariant(engine, ckpt, precision):
    engine.load_weights(ckpt)
    x = _inputs(1000, 106)
    engine.set_max_chunk(8192)
    full, pfull = engine.forward(*x, precision=precision)
    engine.set_max_chunk(256)
    chunked, pch = engine.forward(*x, precision=precision)
    assert np.array_equal(full, chunked) and np.array_equal(pfull, pch)
    for n in (1, 7, 127, 129, 300, 385, 641):       # 1, 2, 3 (odd: a partner-less CTA in the pair kernels), 4 and 6 tiles of 128
        part, _ = engine.forward(*[a[:n] for a in x], precision=precision)
        assert np.array_equal(part, full[:n]), n
    engine.set_max_chunk(8192)
    empty, pe = engine.forward(*[a[:0] for a in x], precision=precision)
    assert empty.shape == (0, 2) and pe.shape == (0,)


def test_device_pointer_entry_matches_host_entry(engine, ckpt):
    import torch
    engine.load_weights(ckpt)
    x = _inputs(700, 107)
    host, hp = engine.forward(*x, precision="bf16")
    dev = [torch.from_numpy(a).cuda() for a in x]
    probs, pred = engine.forward_torch(*dev, precision="bf16")
    torch.cuda.synchronize()
    assert np.array_equal(probs.cpu().numpy(), host) and np.array_equal(pred.cpu().numpy(), hp)


def test_large_batch_properties(engine, ckpt):
    """Full-size property check (no oracle): a 262,144-sample device-resident batch is permutation
    equivariant and its head equals the same samples computed alone."""
    import torch
    engine.load_weights(ckpt)
    n = 262144
    base = _inputs(4096, 108)
    rep = n // 4096
    g = torch.Generator().manual_seed(1)
    perm = torch.randperm(n, generator=g)
    dev = [torch.from_numpy(np.tile(a, (rep, 1))).cuda() for a in base]
    p1, c1 = engine.forward_torch(*dev, precision="bf16")
    dev2 = [t[perm.cuda()].contiguous() for t in dev]
    p2, c2 = engine.forward_torch(*dev2, precision="bf16")
    torch.cuda.synchronize()
    assert torch.equal(p1[perm.cuda()], p2) and torch.equal(c1[perm.cuda()], c2)
    small, _ = engine.forward(*base, precision="bf16")
    assert np.array_equal(p1[:4096].cpu().numpy(), small)
    assert torch.equal(p1[:4096], p1[4096:8192])


def test_errors(engine, ckpt):
    from deepsignal_b200._lib import DsbError
    from deepsignal_b200.engine import Engine
    e2 = Engine(0)
    x = _inputs(4, 109)
    e2.kmer_len, e2.signal_len = 17, 360
    with pytest.raises(DsbError):
        e2.forward(*x)                      # forward before load_weights
    bad = dict(ckpt)
    bad.pop("dense/kernel")
    with pytest.raises(DsbError):
        e2.load_weights(bad)
    bad = dict(ckpt)
    bad["modelembedding"] = np.zeros((10, 128), np.float32)
    with pytest.raises(DsbError):
        e2.load_weights(bad)
    with pytest.raises(DsbError):
        e2.load_weights(ckpt, is_cnn=False, is_rnn=False)
    e2.close()


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_fused_inception_kernels_match_unfused_launches(engine, ckpt, precision, monkeypatch):
    """The two fused inception kernels against the seven separate GEMM/pool launches they replace."""
    x = _inputs(777, 110)
    monkeypatch.setenv("DSB_FUSED", "0")
    engine.load_weights(ckpt)              # plans are rebuilt after a weight load
    a, pa = engine.forward(*x, precision=precision)
    monkeypatch.setenv("DSB_FUSED", "1")
    engine.load_weights(ckpt)
    b, pb = engine.forward(*x, precision=precision)
    d = np.abs(a - b).max()
    print("\n[fused vs unfused, {}] max|dp|={:.3e} identical={}".format(precision, d, np.array_equal(a, b)))
    assert d <= 1e-3


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("n", [110, 700])
def test_persistent_lstm_kernel_matches_per_step_launches(engine, ckpt, precision, n, monkeypatch):
    """lstm_seq.cu (one launch for all directions / layers / steps) against the 102 per-step GEMM launches:
    same K layout and segment order, so the probabilities must be bit-identical."""
    x = _inputs(4321, n)
    monkeypatch.setenv("DSB_LSTM_SEQ", "0")
    engine.load_weights(ckpt)
    a, pa = engine.forward(*x, precision=precision)
    monkeypatch.setenv("DSB_LSTM_SEQ", "1")
    engine.load_weights(ckpt)
    b, pb = engine.forward(*x, precision=precision)
    d = np.abs(a - b).max()
    print("\n[lstm seq vs per-step, {}, n={}] max|dp|={:.3e} identical={}".format(precision, n, d, np.array_equal(a, b)))
    assert np.array_equal(a, b)
    assert np.array_equal(pa, pb)
    # the default is the CTA-pair kernel (lstm_seq2.cu); the single-CTA one (lstm_seq.cu) must agree bit for bit as well
    monkeypatch.setenv("DSB_LSTM_PAIR", "0")
    engine.load_weights(ckpt)
    c, pc = engine.forward(*x, precision=precision)
    monkeypatch.delenv("DSB_LSTM_PAIR")
    assert np.array_equal(a, c)
    assert np.array_equal(pa, pc)


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("n", [3, 700])
def test_fused_front_kernel_matches_separate_launches(engine, ckpt, precision, n, monkeypatch):
    """front_tc.cu (conv1 + pool + conv2 + conv3 in one kernel) against stem_kernel + two GEMM launches: same
    arithmetic and rounding points, so the probabilities must be bit-identical."""
    x = _inputs(99, n)
    monkeypatch.setenv("DSB_FRONT", "0")
    engine.load_weights(ckpt)
    a, pa = engine.forward(*x, precision=precision)
    monkeypatch.setenv("DSB_FRONT", "1")
    engine.load_weights(ckpt)
    b, pb = engine.forward(*x, precision=precision)
    d = np.abs(a - b).max()
    print("\n[front fused vs separate, {}, n={}] max|dp|={:.3e} identical={}".format(precision, n, d, np.array_equal(a, b)))
    assert np.array_equal(a, b)
    assert np.array_equal(pa, pb)
    # default front end (front2_tc.cu): conv1 on the tensor cores as split fp16 parts -- ~2^-22 relative on conv1 BEFORE
    # the 16-bit rounding of its output, i.e. about 1e-4 of the P1 elements round to the neighbouring 16-bit value.  On the
    # spread-logit checkpoint one such flip moves a probability by up to a few 1e-3 in bf16 mode (the same sensitivity
    # that makes bf16-vs-oracle 1.5e-2 there: a flipped element re-draws the rounding noise of every later layer), so this
    # is a statistical-equivalence check against the documented stress bounds, not bit-identity; the oracle tests above run
    # on this default path and their error statistics are unchanged (bf16 1.466e-2 / fp16 1.985e-3 on the 2048-sample case).
    monkeypatch.delenv("DSB_FRONT")
    engine.load_weights(ckpt)
    c, pc = engine.forward(*x, precision=precision)
    d2 = np.abs(a - c).max()
    print("[front2 (tensor-core conv1) vs separate, {}, n={}] max|dp|={:.3e} flips={}".format(precision, n, d2, int((pa != pc).sum())))
    assert d2 <= (BF16_STRESS_BOUND if precision == "bf16" else FP16_STRESS_BOUND)
    assert (pa != pc).mean() <= 0.02


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_repeated_runs_are_bit_identical(engine, ckpt, precision):
    """Races in the warp-specialised kernels (mbarrier hand-offs, relaxed cluster arrives, TMA-store / TMA-load round trips)
    would show up as run-to-run differences; the forward is deterministic by construction (no atomics on the data path)."""
    engine.load_weights(ckpt)
    x = _inputs(3000, 555)
    ref, pref = engine.forward(*x, precision=precision)
    for _ in range(6):
        got, pred = engine.forward(*x, precision=precision)
        assert np.array_equal(got, ref) and np.array_equal(pred, pref)


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_single_kernel_inception_block_matches_two_kernel_path(engine, ckpt, precision, monkeypatch):
    """incept_f.cu (opt-in DSB_AB_FUSED=1: kernels A and B of a block fused, planes never leave the SM, p1 on CTA pairs,
    thread-fed GEMMs as CTA-local MMAs) must be bit-identical to the default A + B path, including an odd tile count."""
    x = _inputs(1100, 31)
    engine.load_weights(ckpt)
    a, pa = engine.forward(*x, precision=precision)
    monkeypatch.setenv("DSB_AB_FUSED", "1")
    engine.load_weights(ckpt)
    b, pb = engine.forward(*x, precision=precision)
    monkeypatch.delenv("DSB_AB_FUSED")
    engine.load_weights(ckpt)
    assert np.array_equal(a, b) and np.array_equal(pa, pb)

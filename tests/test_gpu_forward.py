"""GPU parity of the call_mods forward (dsb_forward_* through the C ABI) against the float64 oracle.

Tolerances are BASELINE.json's: probabilities within 1e-4 abs on the fp32 path, within 2e-3 abs on the
16-bit tensor-core path with >= 99.9 % identical binary calls.  The tensor-core mode that carries that
claim — the default of the CLI, the engine and bench.py — is "fp16x2" (split hi + lo fp16 activations,
fp16 weights, fp32 accumulation); it is held to 2e-3 on BOTH checkpoints at every size below.

Two checkpoints are used.  "tf_default" is the literal random-init checkpoint (TF-1.x default
initialisers): logits stay within a few 1e-3 of zero, so every mode meets 2e-3 easily.  "stress" (the
checkpoint bench.py times) spreads the logits over ~N(1, 1) so that rounding error is visible in the
probabilities.  There the single-rounding modes do NOT meet 2e-3 (fp16 ~2.5e-3, bf16 ~1.5e-2: 11- and
8-bit significands over 25 rounding points, reproduced by oracle/precision_emul.py); they stay available
as faster, looser modes and their error is only characterised here (FAST_MODE_BOUNDS), never claimed as parity."""
import numpy as np
import pytest

from deepsignal_b200 import weights as W
from deepsignal_b200.synth import synth_features
from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
TC_TOL = 2e-3          # BASELINE tolerance of the 16-bit tensor-core path: asserted for the default mode fp16x2
X3_TOL = 1e-4          # three-term split mode (weights hi + lo too): the fp32 path's tolerance, whatever the checkpoint
MIN_IDENTICAL = 0.999
FAST_MODE_BOUNDS = {"fp16": 4e-3, "bf16": 6e-2}     # characterisation of the non-default single-rounding modes


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
    print("\n[{}] n={} max|dp|={:.3e} mean|dp|={:.3e} flips={} ({:.4%}; margins of flips: {})".format(
        tag, len(got), err.max(), err.mean(), int(flips.sum()), flips.mean(), np.round(margin[flips], 5).tolist()[:8]))
    return err, flips, margin


def test_default_precision_is_the_mode_held_to_the_tolerance():
    from deepsignal_b200 import _lib
    assert _lib.DEFAULT_PRECISION == "fp16x2"


def test_fp32_matches_oracle(engine, ckpt):
    engine.load_weights(ckpt)
    x = _inputs(300, 101)
    ref, ref_pred = mo.forward(ckpt, *x)
    got, pred = engine.forward(*x, precision="fp32")
    err, flips, margin = _report("fp32", got, ref, pred, ref_pred)
    assert err.max() <= FP32_TOL
    assert not np.any(flips & (margin > 2 * FP32_TOL))


def test_default_tensor_core_mode_meets_tolerance_vs_oracle_stress(engine, ckpt):
    """fp16x2 against the float64 oracle on the checkpoint the bench times: <= 2e-3, >= 99.9 % identical calls."""
    engine.load_weights(ckpt)
    x = _inputs(4096, 102)
    ref = np.concatenate([mo.forward(ckpt, *[a[i:i + 1024] for a in x])[0] for i in range(0, 4096, 1024)])
    ref_pred = np.argmax(ref, axis=1)
    got, pred = engine.forward(*x, precision="fp16x2")
    err, flips, margin = _report("fp16x2 tcgen05, stress ckpt", got, ref, pred, ref_pred)
    assert err.max() <= TC_TOL
    assert (~flips).mean() >= MIN_IDENTICAL
    assert not np.any(flips & (margin > TC_TOL))


def test_fast_modes_characterised_vs_oracle_stress(engine, ckpt):
    """Single-rounding fp16 / bf16: faster, not parity modes.  Their error on the stress checkpoint is printed and
    checked against the documented ranges only so that a regression (a kernel bug) shows up."""
    engine.load_weights(ckpt)
    x = _inputs(2048, 102)
    ref, ref_pred = mo.forward(ckpt, *x)
    for prec in ("fp16", "bf16"):
        got, pred = engine.forward(*x, precision=prec)
        err, flips, margin = _report(prec + " tcgen05 (fast mode), stress ckpt", got, ref, pred, ref_pred)
        assert err.max() <= FAST_MODE_BOUNDS[prec]
        assert not np.any(flips & (margin > 2 * FAST_MODE_BOUNDS[prec]))


def test_default_mode_tolerance_at_100k_samples(engine, ckpt):
    """BASELINE configs[2]'s tolerance report size: 100,000 cfg-3 samples.  The float64 oracle needs minutes for that on
    the host, so the reference here is the library's fp32 path — itself pinned to the float64 oracle on a 2,048-sample
    subset inside this test (<= 1e-4; measured ~1e-6)."""
    engine.load_weights(ckpt)
    n = 100000
    x = _inputs(n, 20260103)
    a, pa = engine.forward(*x, precision="fp32")
    sub = slice(0, 2048)
    ref, _ = mo.forward(ckpt, *[t[sub] for t in x])
    pin = np.abs(a[sub] - ref).max()
    assert pin <= FP32_TOL
    b, pb = engine.forward(*x, precision="fp16x2")
    d = np.abs(a - b)
    flips = pa != pb
    margin = np.abs(a[:, 0] - a[:, 1])
    buckets = [(0, 1e-3), (1e-3, 2e-3), (2e-3, 5e-3), (5e-3, 1.0)]
    hist = {"{:g}-{:g}".format(lo, hi): int((flips & (margin >= lo) & (margin < hi)).sum()) for lo, hi in buckets}
    print("\n[fp16x2 vs fp32 path, n=100000, stress] fp32-vs-f64-oracle pin={:.2e}  max|dp|={:.3e} mean={:.3e} identical={:.5%} "
          "flips by |p1-p0|: {}".format(pin, d.max(), d.mean(), (~flips).mean(), hist))
    assert d.max() + pin <= TC_TOL
    assert (~flips).mean() >= MIN_IDENTICAL
    assert not np.any(flips & (margin > TC_TOL))
    for prec in ("fp16", "bf16"):
        c, pc = engine.forward(*x, precision=prec)
        dc = np.abs(a - c)
        print("[{} (fast mode) vs fp32 path, n=100000] max|dp|={:.3e} mean={:.3e} identical={:.5%}".format(
            prec, dc.max(), dc.mean(), (pa == pc).mean()))
        assert dc.max() <= FAST_MODE_BOUNDS[prec]


@pytest.mark.parametrize("kw", [dict(is_cnn=False), dict(is_rnn=False), dict(is_base=False)])
@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("fp16x3", X3_TOL), ("fp16x2", TC_TOL), ("fp16", FAST_MODE_BOUNDS["fp16"]),
                                           ("bf16", FAST_MODE_BOUNDS["bf16"])])
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
    for precision, tol in (("fp32", FP32_TOL), ("fp16x2", TC_TOL), ("bf16", TC_TOL), ("fp16", TC_TOL)):
        got, pred = engine.forward(*x, precision=precision)
        print("\n[{} tf_default] max|dp|={:.3e} identical calls={:.4%}".format(precision, np.abs(got - ref).max(), (pred == ref_pred).mean()))
        assert np.abs(got - ref).max() <= tol
        assert (pred == ref_pred).mean() >= MIN_IDENTICAL


@pytest.mark.parametrize("precision", ["fp32", "fp16x3", "fp16x2", "bf16", "fp16"])
def test_chunking_and_ragged_sizes_are_batch_invariant(engine, ckpt, precision):
    engine.load_weights(ckpt)
    x = _inputs(1000, 106)
    engine.set_max_chunk(8192)
    full, pfull = engine.forward(*x, precision=precision)
    engine.set_max_chunk(256)
    chunked, pch = engine.forward(*x, precision=precision)
    assert np.array_equal(full, chunked) and np.array_equal(pfull, pch)
    for n in (1, 7, 127, 129, 300, 385, 641):       # 1, 2, 3 (odd: a partner-less CTA in the pair kernel), 4 and 6 tiles of 128
        part, _ = engine.forward(*[a[:n] for a in x], precision=precision)
        assert np.array_equal(part, full[:n]), n
    engine.set_max_chunk(8192)
    empty, pe = engine.forward(*[a[:0] for a in x], precision=precision)
    assert empty.shape == (0, 2) and pe.shape == (0,)


def test_device_pointer_entry_matches_host_entry(engine, ckpt):
    import torch
    engine.load_weights(ckpt)
    x = _inputs(700, 107)
    for precision in ("fp16x2", "bf16"):
        host, hp = engine.forward(*x, precision=precision)
        dev = [torch.from_numpy(a).cuda() for a in x]
        probs, pred = engine.forward_torch(*dev, precision=precision)
        torch.cuda.synchronize()
        assert np.array_equal(probs.cpu().numpy(), host) and np.array_equal(pred.cpu().numpy(), hp)
    assert engine.nonfinite_count() == 0


@pytest.mark.parametrize("precision", ["fp16x2", "bf16"])
def test_large_batch_properties(engine, ckpt, precision):
    """Full-size property check (no oracle): a 262,144-sample device-resident batch is permutation
    equivariant and its head equals the same samples computed alone."""
    import torch
    engine.load_weights(ckpt)
    engine.set_max_chunk(37888)
    n = 262144
    base = _inputs(4096, 108)
    rep = n // 4096
    g = torch.Generator().manual_seed(1)
    perm = torch.randperm(n, generator=g)
    dev = [torch.from_numpy(np.tile(a, (rep, 1))).cuda() for a in base]
    p1, c1 = engine.forward_torch(*dev, precision=precision)
    dev2 = [t[perm.cuda()].contiguous() for t in dev]
    p2, c2 = engine.forward_torch(*dev2, precision=precision)
    torch.cuda.synchronize()
    assert torch.equal(p1[perm.cuda()], p2) and torch.equal(c1[perm.cuda()], c2)
    small, _ = engine.forward(*base, precision=precision)
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
    e2.load_weights(ckpt)
    with pytest.raises(DsbError):
        e2.forward(*x, precision=7)         # unknown precision code reaches the library and is rejected there
    e2.close()


@pytest.mark.parametrize("precision", ["fp16x3", "fp16x2", "fp16", "bf16"])
def test_outlier_inputs_have_a_defined_outcome(engine, ckpt, precision):
    """16-bit overflow handling: inputs far outside the training range (signal +-50 sigma, a 5000-sample stall) must
    either give finite probabilities within tolerance of the fp32 path or be reported as DSB_ERR_NONFINITE — never
    silent garbage.  Inputs large enough to overflow fp16 (1e6) must raise."""
    from deepsignal_b200._lib import DsbError, DSB_ERR_NONFINITE
    engine.load_weights(ckpt)
    kmer, means, stds, lens, signals = [a.copy() for a in _inputs(256, 111)]
    signals[::7, ::5] = 50.0
    signals[::11, 1::9] = -50.0
    lens[::3, 4] = 5000.0
    means[::5, 2] = 25.0
    x = (kmer, means, stds, lens, signals)
    ref, _ = engine.forward(*x, precision="fp32")
    assert np.isfinite(ref).all()
    try:
        got, _ = engine.forward(*x, precision=precision)
    except DsbError as e:
        assert e.code == DSB_ERR_NONFINITE
    else:
        assert np.isfinite(got).all()
        print("\n[{} outliers] max|dp| vs fp32 = {:.3e}".format(precision, np.abs(got - ref).max()))
        if precision in ("fp16x2", "fp16x3"):
            assert np.abs(got - ref).max() <= (TC_TOL if precision == "fp16x2" else X3_TOL)
    if precision != "bf16":                  # bf16 has the fp32 exponent range
        signals[0, :] = 1e6
        with pytest.raises(DsbError) as ei:
            engine.forward(kmer, means, stds, lens, signals, precision=precision)
        assert ei.value.code == DSB_ERR_NONFINITE
        good, _ = engine.forward(*_inputs(64, 112), precision=precision)      # the counter was reset: the next call is clean
        assert np.isfinite(good).all()


@pytest.mark.parametrize("precision", ["fp16x2", "bf16", "fp16"])
def test_fused_inception_kernels_match_unfused_launches(engine, ckpt, precision):
    """The two fused inception kernels against the seven separate GEMM/pool launches they replace (same rounding
    points; the accumulation order of the p1 GEMM differs, so agreement is to rounding noise, not bit-identity)."""
    x = _inputs(777, 110)
    engine.load_weights(ckpt)
    engine.set_plan_options(fused_blocks=0)
    a, pa = engine.forward(*x, precision=precision)
    engine.set_plan_options(fused_blocks=1)
    b, pb = engine.forward(*x, precision=precision)
    d = np.abs(a - b).max()
    print("\n[fused vs unfused, {}] max|dp|={:.3e} identical={}".format(precision, d, np.array_equal(a, b)))
    assert d <= (1e-5 if precision == "fp16x2" else 1e-3)


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("n", [110, 700])
def test_persistent_lstm_kernel_matches_per_step_launches(engine, ckpt, precision, n):
    """lstm_seq2.cu (one launch for all directions / layers / steps) against the 102 per-step GEMM launches:
    same K layout and segment order, so the probabilities must be bit-identical."""
    x = _inputs(4321, n)
    engine.load_weights(ckpt)
    engine.set_plan_options(lstm_seq=0)
    a, pa = engine.forward(*x, precision=precision)
    engine.set_plan_options(lstm_seq=1)
    b, pb = engine.forward(*x, precision=precision)
    d = np.abs(a - b).max()
    print("\n[lstm seq vs per-step, {}, n={}] max|dp|={:.3e} identical={}".format(precision, n, d, np.array_equal(a, b)))
    assert np.array_equal(a, b)
    assert np.array_equal(pa, pb)


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("n", [3, 700])
def test_fused_front_kernel_matches_separate_launches(engine, ckpt, precision, n):
    """front2_tc.cu (conv1 + pool + conv2 + conv3 in one kernel, conv1 on the tensor cores as split fp16 parts: ~2^-22
    relative before the 16-bit rounding of its output) against stem_kernel + two GEMM launches.  About 1e-4 of the P1
    elements round to the neighbouring 16-bit value, which re-draws the rounding noise of the later layers in these
    single-rounding modes: a statistical-equivalence check, not bit-identity."""
    x = _inputs(99, n)
    engine.load_weights(ckpt)
    engine.set_plan_options(fused_front=0)
    a, pa = engine.forward(*x, precision=precision)
    engine.set_plan_options(fused_front=1)
    c, pc = engine.forward(*x, precision=precision)
    d2 = np.abs(a - c).max()
    print("\n[front2 (tensor-core conv1) vs separate, {}, n={}] max|dp|={:.3e} flips={}".format(precision, n, d2, int((pa != pc).sum())))
    assert d2 <= FAST_MODE_BOUNDS[precision]
    assert (pa != pc).mean() <= 0.02


@pytest.mark.parametrize("precision", ["fp16x3", "fp16x2", "bf16", "fp16"])
def test_repeated_runs_are_bit_identical(engine, ckpt, precision):
    """Races in the warp-specialised kernels (mbarrier hand-offs, relaxed cluster arrives, TMA-store / TMA-load round trips)
    would show up as run-to-run differences; the forward is deterministic by construction (no atomics on the data path)."""
    engine.load_weights(ckpt)
    x = _inputs(3000, 555)
    ref, pref = engine.forward(*x, precision=precision)
    for _ in range(6):
        got, pred = engine.forward(*x, precision=precision)
        assert np.array_equal(got, ref) and np.array_equal(pred, pref)


def test_three_term_split_mode_is_checkpoint_independent(engine):
    """fp16x3 (activations AND weights as hi + lo fp16, three product terms on tcgen05): within the fp32 path's 1e-4 on
    checkpoints where the default fp16x2 mode -- whose residual is the fp16 rounding of the CNN weights -- leaves 2e-3
    (profiles/r2_tolerance_survey.jsonl), and identical calls."""
    x = _inputs(1024, 104)
    for seed in (17, 47, 59):
        ck = W.random_checkpoint(seed, "stress")
        engine.load_weights(ck)
        ref, ref_pred = mo.forward(ck, *x)
        got, pred = engine.forward(*x, precision="fp16x3")
        err = np.abs(got - ref).max()
        print("seed", seed, "fp16x3 max|dp|", err)
        assert err <= X3_TOL
        margin = np.abs(ref[:, 0] - ref[:, 1])
        assert not np.any((pred != ref_pred) & (margin > X3_TOL))

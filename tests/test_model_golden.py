"""oracle/model_oracle.py against the golden vectors that the REFERENCE's own graph code produced (deepsignal/model.py +
layers.py executed unmodified over oracle/tf_shim.py by tests/golden/make_model_golden.py).

This pins the oracle's wiring of the forward to the reference's source.  TensorFlow itself cannot run here, so the
semantics of the individual TF ops remain restated (once, in oracle/tf_shim.py, independently of model_oracle.py which is
written on torch ops): the primitive-level known-answer tests are in tests/test_model_oracle.py and below."""
import hashlib
import importlib.util
import json
import os

import numpy as np
import pytest

from deepsignal_b200 import weights as W
from oracle import model_oracle as mo
from oracle import tf_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "model_ref.json")


def _gen():
    spec = importlib.util.spec_from_file_location("make_model_golden", os.path.join(ROOT, "tests", "golden", "make_model_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cases():
    return json.load(open(GOLDEN))["cases"]


@pytest.mark.parametrize("name", ["stress17", "tfdefault23", "stress31_no_base", "stress31_no_rnn", "stress31_no_cnn"])
def test_oracle_matches_reference_graph_code(name):
    c = _cases()[name]
    kw = c.get("model_kw", {})
    ck = W.random_checkpoint(*c["checkpoint"], **kw)
    kmer, means, stds, lens, signals = _gen().case_inputs(c["features_seed"], c["n"])
    want = np.asarray([[float(v) for v in row] for row in c["activation_logits"]], np.float64)
    got, pred = mo.forward(ck, kmer, means, stds, lens, signals, **kw)
    # both sides are float64; they differ only in summation order (einsum / matmul vs conv1d)
    assert np.max(np.abs(got - want)) <= 1e-10
    assert pred.tolist() == c["prediction"]
    # the variable names the reference's scopes generate are exactly the checkpoint's (weights.variable_specs)
    names = [n for n, _ in W.variable_specs(**kw)]
    assert c["n_variables"] == len(names) == len(ck)      # incl. the signal model's variables when is_cnn=False (model.py:83-85)
    assert set(names) == set(ck)


def test_golden_is_reproducible_from_reference_sources():
    """Re-runs the reference's graph code when its sources are present (build container, or oracle/_ref on the GPU box)."""
    gen = _gen()
    ref = gen.reference_dir()
    if ref is None:
        pytest.skip("reference sources not available")
    c = _cases()["stress17"]
    ck = W.random_checkpoint(*c["checkpoint"])
    kmer, means, stds, lens, signals = [a[:3] for a in gen.case_inputs(c["features_seed"], c["n"])]
    act, pred, used = gen.run_reference_model(ck, kmer, means, stds, lens, signals, ref)
    want = np.asarray([[float(v) for v in row] for row in c["activation_logits"][:3]], np.float64)
    assert np.max(np.abs(act - want)) <= 1e-12            # per-sample independence: the first 3 of the 6 golden rows
    assert hashlib.sha1("\n".join(used).encode()).hexdigest() == c["variables_sha1"]


def test_shim_primitives_known_answers():
    """Hand-computed answers for the TF op semantics restated in oracle/tf_shim.py."""
    # SAME padding (TF): n=5, k=3, s=2 -> out 3, pad (1, 1); n=6, k=3, s=2 -> out 3, pad (0, 1); n=90, k=7, s=1 -> (3, 3)
    assert tf_shim._same_pad(5, 3, 2) == (3, 1, 1)
    assert tf_shim._same_pad(6, 3, 2) == (3, 0, 1)
    assert tf_shim._same_pad(90, 7, 1) == (90, 3, 3)
    assert tf_shim._same_pad(360, 7, 2) == (180, 2, 3)
    x = np.arange(1.0, 7.0).reshape(1, 1, 6, 1)
    # max pool k=3 s=2 SAME on [1..6]: windows (1,2,3) (3,4,5) (5,6,pad) -> 3, 5, 6
    assert tf_shim.max_pooling2d(x, [1, 3], strides=2).ravel().tolist() == [3.0, 5.0, 6.0]
    # average pool k=3 s=1 SAME divides by the number of valid elements: (1+2)/2, 2, 3, 4, 5, (5+6)/2
    assert tf_shim.average_pooling2d(x, [1, 3], strides=1).ravel().tolist() == [1.5, 2.0, 3.0, 4.0, 5.0, 5.5]
    # conv k=3 s=2 SAME, kernel [1, 10, 100]: pad (0, 1): 1+20+300, 3+40+500, 5+60+0
    tf_shim.begin({"c/kernel": np.asarray([1.0, 10.0, 100.0]).reshape(1, 3, 1, 1)}, [])
    assert tf_shim.conv2d(x, 1, [1, 3], strides=2, use_bias=False, name="c").ravel().tolist() == [321.0, 543.0, 65.0]
    # LSTMCell: gate order i, j, f, o and forget_bias = 1
    cell = tf_shim.LSTMCell(1)
    kernel = np.zeros((2, 4))
    bias = np.asarray([0.0, 100.0, -1.0, 100.0])          # i=0 -> 0.5, j -> tanh = 1, f + 1 = 0 -> 0.5, o -> 1
    c2, h2 = cell.step(np.zeros((1, 1)), np.full((1, 1), 2.0), np.zeros((1, 1)), kernel, bias)
    assert abs(c2[0, 0] - (0.5 * 2.0 + 0.5 * 1.0)) < 1e-12 and abs(h2[0, 0] - np.tanh(1.5)) < 1e-12
    # batch norm, inference: (x - mean) / sqrt(var + 1e-3) * gamma + beta
    tf_shim.begin({"bn/beta": np.asarray([1.0]), "bn/gamma": np.asarray([2.0]), "bn/moving_mean": np.asarray([3.0]),
                   "bn/moving_variance": np.asarray([0.999])}, [])
    y = tf_shim.batch_norm(inputs=np.asarray([[5.0]]), is_training=False, scope="bn", center=True, scale=True)
    assert abs(float(y[0, 0]) - 5.0) < 1e-12


def test_same_padding_rule_against_an_independent_tf_port():
    """TensorFlow's asymmetric "SAME" padding (the extra sample goes to the right) restated in oracle/tf_shim.py and
    oracle/model_oracle.py, against HuggingFace transformers' `apply_tf_padding` -- written by a third party to load
    TF-slim checkpoints into torch and validated against them -- for every (length, kernel, stride) the model uses.
    The same port's default batch-norm epsilon (TF-slim = tf.contrib.layers.batch_norm) is the 1e-3 used here."""
    torch = pytest.importorskip("torch")
    mnet = pytest.importorskip("transformers.models.mobilenet_v1.modeling_mobilenet_v1")
    from transformers import MobileNetV1Config
    shapes = [(360, 7, 2), (180, 3, 2), (90, 1, 1), (90, 3, 1), (90, 5, 1), (90, 3, 2), (45, 3, 1), (45, 5, 1), (45, 3, 2),
              (23, 3, 1), (23, 5, 1), (23, 7, 1), (5, 3, 2), (6, 3, 2)]
    for n, k, s in shapes:
        conv = torch.nn.Conv2d(1, 1, (1, k), stride=(1, s))
        x = torch.zeros(1, 1, 1, n)
        x[..., 0] = 1.0
        y = mnet.apply_tf_padding(x, conv)
        left = int(torch.nonzero(y[0, 0, 0])[0])
        right = y.shape[-1] - n - left
        out, pl, pr = tf_shim._same_pad(n, k, s)
        assert (pl, pr) == (left, right), (n, k, s)
        assert mo.same_pads(n, k, s) == (left, right), (n, k, s)
        assert out == -(-n // s)
    assert MobileNetV1Config().layer_norm_eps == W.BN_EPS == 0.001

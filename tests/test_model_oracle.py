"""Hand-computed known-answer cases for every TF-1.x rule the model oracle restates
(SURVEY §8a notes).  The reference holds no golden vectors for the forward and TF
cannot run here, so these are the pins the oracle has ("parity unpinned" vs real TF)."""
import numpy as np
import torch

from oracle import model_oracle as mo
from deepsignal_b200 import weights as W

T64 = torch.float64


def test_same_padding_rule():
    assert mo.same_pads(360, 7, 2) == (2, 3)
    assert mo.same_pads(180, 3, 2) == (0, 1)
    assert mo.same_pads(90, 3, 2) == (0, 1)
    assert mo.same_pads(45, 3, 2) == (1, 1)
    assert mo.same_pads(90, 3, 1) == (1, 1)
    assert mo.same_pads(45, 5, 1) == (2, 2)
    assert mo.same_pads(23, 7, 1) == (3, 3)
    assert W.cnn_lengths(360) == (180, 90, 45, 23)


def test_conv_is_cross_correlation_hwio():
    x = torch.tensor([[[1., 2., 3., 4., 5.]]], dtype=T64)
    k = torch.tensor([1., 10., 100.], dtype=T64).reshape(1, 3, 1, 1)
    assert mo.conv_same(x, k).flatten().tolist() == [210., 321., 432., 543., 54.]
    assert mo.conv_same(x, k, 2).flatten().tolist() == [210., 432., 54.]
    x4 = x[:, :, :4]
    assert mo.conv_same(x4, k, 2).flatten().tolist() == [321., 43.]   # pad (0,1)
    # HWIO: kernel[0, t, cin, cout]
    x2 = torch.tensor([[[1., 2.], [10., 20.]]], dtype=T64)            # [B=1, C=2, L=2]
    k2 = torch.zeros(1, 1, 2, 3, dtype=T64)
    k2[0, 0, 0, 1] = 1.0
    k2[0, 0, 1, 2] = 2.0
    y = mo.conv_same(x2, k2)
    assert y[0, 0].tolist() == [0., 0.] and y[0, 1].tolist() == [1., 2.] and y[0, 2].tolist() == [20., 40.]


def test_pooling_rules():
    x = torch.tensor([[[-5., -3., -4.]]], dtype=T64)
    assert mo.maxpool_same(x, 3, 1).flatten().tolist() == [-3., -3., -3.]      # -inf padding
    x = torch.arange(1., 7.).reshape(1, 1, 6).to(T64)
    assert mo.maxpool_same(x, 3, 2).flatten().tolist() == [3., 5., 6.]         # pad (0,1)
    x = torch.arange(1., 6.).reshape(1, 1, 5).to(T64)
    assert mo.maxpool_same(x, 3, 2).flatten().tolist() == [2., 4., 5.]         # pad (1,1)
    x = torch.tensor([[[1., 2., 3., 4.]]], dtype=T64)
    y = mo.avgpool_same(x, 7, 1).flatten().tolist()                            # valid-count divisor
    assert y == [2.5, 2.5, 2.5, 2.5]
    x = torch.arange(1., 10.).reshape(1, 1, 9).to(T64)
    y = mo.avgpool_same(x, 7, 1).flatten()
    assert abs(y[0].item() - (1 + 2 + 3 + 4) / 4) < 1e-12 and abs(y[4].item() - 5.0) < 1e-12


def _zero_lstm_ckpt(in0=131):
    ck = {}
    for d in ("fw", "bw"):
        for l in range(3):
            cin = in0 if l == 0 else 256
            base = "modelem/%s/multi_rnn_cell/cell_%d/lstm_cell/" % (d, l)
            ck[base + "kernel"] = torch.zeros(cin + 256, 1024, dtype=T64)
            ck[base + "bias"] = torch.zeros(1024, dtype=T64)
    return ck


def test_lstm_gate_order_and_forget_bias():
    ck = _zero_lstm_ckpt()
    for d in ("fw", "bw"):
        for l in range(3):
            b = ck["modelem/%s/multi_rnn_cell/cell_%d/lstm_cell/bias" % (d, l)]
            b[0:256] = 50.0        # i  -> sigmoid 1
            b[256:512] = 50.0      # j  -> tanh 1
            b[512:768] = -1.0      # f  -> sigmoid(f + 1.0) = 0.5
            b[768:1024] = 50.0     # o  -> sigmoid 1
    x = torch.zeros(2, 17, 131, dtype=T64)
    out = mo.lstm_stack(x, ck, "fw")
    c = 0.0
    for t in range(17):
        c = 0.5 * c + 1.0
        assert abs(out[0, t, 0].item() - np.tanh(c)) < 1e-12
    cell_h, cell_c = mo.lstm_cell(torch.zeros(1, 131, dtype=T64), torch.zeros(1, 256, dtype=T64),
                                  torch.full((1, 256), 2.0, dtype=T64),
                                  torch.zeros(387, 1024, dtype=T64), torch.zeros(1024, dtype=T64))
    sig1 = 1 / (1 + np.exp(-1.0))
    assert abs(cell_c[0, 0].item() - 2.0 * sig1) < 1e-12              # i=j=0 adds 0.5*tanh(0)=0
    assert abs(cell_h[0, 0].item() - 0.5 * np.tanh(2.0 * sig1)) < 1e-12


def test_bidirectional_output_selection():
    """fw emits the state after t=16, bw the state after consuming t=16..0 (i.e. at t=0)."""
    ck = _zero_lstm_ckpt()
    ck["modelembedding"] = torch.zeros(1024, 128, dtype=T64)
    for d in ("fw", "bw"):
        for l in range(3):
            base = "modelem/%s/multi_rnn_cell/cell_%d/lstm_cell/" % (d, l)
            K, b = ck[base + "kernel"], ck[base + "bias"]
            b[0:256] = 50.0
            b[512:768] = -50.0     # forget everything: c_t = tanh(j_t)
            b[768:1024] = 50.0
            if l == 0:
                K[128, 256:512] = 100.0        # 'means' input row -> j gate
            else:
                K[0:256, 256:512] = torch.eye(256, dtype=T64) * 100.0
    means = torch.zeros(1, 17, dtype=T64)
    means[0, 0] = 1.0
    means[0, 16] = -1.0
    means[0, 8] = 0.3
    z = torch.zeros(1, 17, dtype=T64)
    out = mo.event_model(torch.zeros(1, 17, dtype=torch.int64), means, z, z, ck)
    assert out.shape == (1, 512)
    assert out[0, 0].item() < -0.5 and out[0, 256].item() > 0.5


def test_embedding_and_input_concat_order():
    ck = _zero_lstm_ckpt()
    emb = torch.zeros(1024, 128, dtype=T64)
    emb[3, 5] = 1.0
    ck["modelembedding"] = emb
    K = ck["modelem/fw/multi_rnn_cell/cell_0/lstm_cell/kernel"]
    b = ck["modelem/fw/multi_rnn_cell/cell_0/lstm_cell/bias"]
    b[0:256] = 50.0
    b[768:1024] = 50.0
    K[5, 256] = 1.0       # embedding dim 5  -> j[0]
    K[128, 257] = 1.0     # mean             -> j[1]
    K[129, 258] = 1.0     # std              -> j[2]
    K[130, 259] = 1.0     # signal length    -> j[3]
    kmer = torch.full((1, 17), 3, dtype=torch.int64)
    one = torch.ones(1, 17, dtype=T64)
    x = torch.cat([emb[kmer], (0.1 * one)[..., None], (0.2 * one)[..., None], (0.3 * one)[..., None]], dim=2)
    h, c = mo.lstm_cell(x[:, 0], torch.zeros(1, 256, dtype=T64), torch.zeros(1, 256, dtype=T64), K, b)
    for j, v in enumerate((1.0, 0.1, 0.2, 0.3)):
        assert abs(c[0, j].item() - np.tanh(v)) < 1e-12


def test_shapes_flatten_and_concat_order():
    ckpt = W.random_checkpoint(seed=1, variant="stress")
    ck = {k: torch.as_tensor(v, dtype=T64) for k, v in ckpt.items()}
    x = torch.randn(2, 360, dtype=T64, generator=torch.Generator().manual_seed(0))
    flat, st = mo.signal_model(x, ck, return_stages=True)
    assert st["conv1"].shape == (2, 64, 180) and st["pool1"].shape == (2, 64, 90)
    assert st["conv2"].shape == (2, 128, 90) and st["conv3"].shape == (2, 256, 90)
    assert st["incp3"].shape == (2, 240, 90) and st["pool_after3"].shape == (2, 240, 45)
    assert st["incp8"].shape == (2, 240, 45) and st["pool_after8"].shape == (2, 240, 23)
    assert st["avgpool"].shape == (2, 240, 23) and flat.shape == (2, 5520)
    assert flat[1, 7 * 240 + 11].item() == st["avgpool"][1, 11, 7].item()      # position-major
    # concat order [b1,b2,b3,b4,b5]: zero branch-2 conv -> channels 48..95 become relu(beta-term)
    sc = W.inception_scope(1)
    ck2 = dict(ck)
    ck2[sc + "branch2_1x1/conv0b_1x1/kernel"] = torch.zeros_like(ck[sc + "branch2_1x1/conv0b_1x1/kernel"])
    y = mo.inception_block(st["conv3"], ck2, 1)
    g, b, m, v = (ck[sc + "branch2_1x1/bn/" + f] for f in W.BN_FIELDS)
    expect = torch.relu((0 - m) * g / torch.sqrt(v + 1e-3) + b)
    assert torch.allclose(y[0, 48:96, 0], expect, atol=1e-12)
    y_ref = mo.inception_block(st["conv3"], ck, 1)
    assert torch.equal(y[:, :48], y_ref[:, :48]) and torch.equal(y[:, 96:], y_ref[:, 96:])
    assert (y_ref >= 0).all()


def test_forward_variants_and_tie_break():
    rng = np.random.default_rng(0)
    n = 3
    kmer = rng.integers(0, 4, (n, 17)).astype(np.int32)
    means, stds = rng.normal(size=(n, 17)), np.abs(rng.normal(size=(n, 17)))
    lens = rng.integers(3, 30, (n, 17)).astype(np.float64)
    sig = rng.normal(size=(n, 360))
    for kw in (dict(), dict(is_cnn=False), dict(is_rnn=False), dict(is_base=False)):
        ck = W.random_checkpoint(seed=2, variant="stress", **kw)
        probs, pred = mo.forward(ck, kmer, means, stds, lens, sig, **kw)
        assert probs.shape == (n, 2) and pred.shape == (n,)
        assert np.array_equal(pred, np.argmax(probs, axis=1))
    ck = W.random_checkpoint(seed=2, variant="tf_default")
    ck["dense_1/kernel"] = np.zeros_like(ck["dense_1/kernel"])
    probs, pred = mo.forward(ck, kmer, means, stds, lens, sig)
    assert np.all(probs == 0.5) and np.all(pred == 0)                  # argmax ties -> class 0


def test_variable_names_and_param_count():
    specs = dict(W.variable_specs())
    assert specs["modelembedding"] == (1024, 128)
    assert specs["modelem/fw/multi_rnn_cell/cell_0/lstm_cell/kernel"] == (387, 1024)
    assert specs["modelem/bw/multi_rnn_cell/cell_2/lstm_cell/kernel"] == (512, 1024)
    assert specs["modelsignalmconv_layer1/conv/kernel"] == (1, 7, 1, 64)
    assert specs["modelsignalmincp_layer1/modelsignalm1branch1_maxpooling/conv1a_1x1/kernel"] == (1, 1, 256, 48)
    assert specs["modelsignalmincp_layer11/modelsignalm11branch5_residual_1x3/conv1e_1x3/kernel"] == (1, 3, 32, 64)
    assert specs["modelsignalmincp_layer4/modelsignalm4branch4_1x5/bn2/moving_variance"] == (48,)
    assert specs["dense/kernel"] == (6032, 6032) and specs["dense_1/kernel"] == (6032, 2)
    conv_params = sum(int(np.prod(s)) for n, s in specs.items() if n.endswith("kernel") and len(s) == 4)
    macs = 0
    L = {"conv_layer1": 180, "conv_layer2": 90, "conv_layer3": 90}
    for n, s in specs.items():
        if n.endswith("kernel") and len(s) == 4:
            if "incp_layer" in n:
                i = int(n.split("incp_layer")[1].split("/")[0])
                l = 90 if i <= 3 else (45 if i <= 8 else 23)
            else:
                l = L[n.split("signalm")[1].split("/")[0]]
            macs += l * int(np.prod(s))
    assert macs == 54_625_536                                           # SURVEY §8a CNN MACs/sample
    assert conv_params > 900_000

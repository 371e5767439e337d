"""GPU parity of the extraction kernels (dsb_extract_host through the C ABI) against the golden
fixtures produced by the reference's own code and against the NumPy oracle.  Integer features
(k-mer codes, slice lengths, site locations / positions, window indices) must be bit-exact; float
features within 1e-4 abs (BASELINE.json) — in practice they are bit-identical in float64."""
import numpy as np
import pytest

from deepsignal_b200.synth import CONTIGS, synth_reads
from deepsignal_b200.utils.process_utils import get_motif_seqs
from oracle import features_oracle as fo

pytestmark = pytest.mark.gpu
CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4}


@pytest.fixture(scope="module")
def engine():
    from deepsignal_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


def _oracle(rb, motifs, mod_loc=0, chrom2len=CONTIGS, method="mad", kmer_len=17, slen=360):
    feats = []
    for i in range(rb.n_reads):
        feats += [(i,) + f for f in fo.extract_read(rb.read_dict(i), motifs, mod_loc, chrom2len, kmer_len, slen, 1, None, method)]
    return feats


def _compare(got, want, K=17, exact_floats=True):
    assert got["n_sites"] == len(want)
    if not want:
        return
    assert np.array_equal(got["site_read"], np.asarray([w[0] for w in want]))
    assert np.array_equal(got["pos"], np.asarray([w[2] for w in want]))
    assert np.array_equal(got["pos_in_strand"], np.asarray([w[4] for w in want]))
    assert np.array_equal(got["kmer"], np.asarray([[CODE[c] for c in w[7]] for w in want]))
    assert np.array_equal(got["lens_i32"], np.asarray([w[10] for w in want]))
    assert np.array_equal(got["lens"], np.asarray([w[10] for w in want], dtype=np.float32))
    means = np.asarray([w[8] for w in want])
    stds = np.asarray([w[9] for w in want])
    sig = np.asarray([w[11] for w in want])
    dm, ds, dx = (np.nanmax(np.abs(got["means64"] - means)), np.nanmax(np.abs(got["stds64"] - stds)),
                  np.abs(got["signals64"] - sig).max())
    print("\nmax abs diff means {:.3e} stds {:.3e} signals {:.3e}; bit-identical: {} {} {}".format(
        dm, ds, dx, np.array_equal(got["means64"], means, equal_nan=True),
        np.array_equal(got["stds64"], stds, equal_nan=True), np.array_equal(got["signals64"], sig)))
    assert dm <= 1e-4 and ds <= 1e-4 and dx <= 1e-4
    assert np.array_equal(got["means"], means.astype(np.float32), equal_nan=True) or dm > 0
    assert np.abs(got["signals"] - sig.astype(np.float32)).max() <= 1e-4
    if exact_floats:
        assert np.array_equal(got["signals64"], sig)
    # window indices
    for i, w in enumerate(want):
        mode, start, nreal = fo.central_window(w[10], got["signals"].shape[1])
        assert tuple(got["window"][i]) == (mode, start, nreal)


def test_golden_reference_fixture(engine, golden, golden_reads):
    a, m = golden
    got = engine.extract(golden_reads, ["CG"])
    assert got["n_sites"] == m["extract_n"]
    info = m["extract_info"]
    assert got["pos"].tolist() == [x[1] for x in info]
    assert got["pos_in_strand"].tolist() == [x[3] for x in info]
    assert ["".join("ACGTN"[c] for c in row) for row in got["kmer"]] == [x[6] for x in info]
    assert np.array_equal(got["lens_i32"], a["extract_lens"])
    assert np.abs(got["means64"] - a["extract_means"]).max() <= 1e-4
    assert np.abs(got["stds64"] - a["extract_stds"]).max() <= 1e-4
    assert np.abs(got["signals64"] - a["extract_cent"]).max() <= 1e-4
    print("\nbit-identical to reference: means {} stds {} signals {}".format(
        np.array_equal(got["means64"], a["extract_means"]), np.array_equal(got["stds64"], a["extract_stds"]),
        np.array_equal(got["signals64"], a["extract_cent"])))
    # normalisation constants equal numpy's median / MAD on the rescaled signal
    for i in range(golden_reads.n_reads):
        x = a["rescaled_%d" % i]
        assert got["read_shift"][i] == np.median(x)
        assert got["read_scale"][i] == fo.mad(x)


def test_oracle_parity_cfg2_reads(engine):
    rb = synth_reads(200, seed=20260102)
    want = _oracle(rb, ["CG"])
    got = engine.extract(rb, ["CG"])
    _compare(got, want)
    modes = set(got["window"][:, 0].tolist())
    assert modes == {0, 1, 2}, modes        # all three window branches exercised


def test_zscore_no_reference_and_motif_variants(engine):
    rb = synth_reads(40, seed=77, nb_lo=400, nb_hi=900, cg_frac=0.03, with_reference=False)
    got = engine.extract(rb, ["CG"], normalize_method="zscore")
    want = _oracle(rb, ["CG"], chrom2len=None, method="zscore")
    _compare(got, want, exact_floats=False)
    assert np.all(got["pos_in_strand"] == -1)
    for motif, loc in (("CHG", 0), ("CCWGG", 1), ("CHH,CHG", 0), ("NCG", 1)):
        ms = get_motif_seqs(motif)
        got = engine.extract(rb, ms, mod_loc=loc)
        want = _oracle(rb, ms, mod_loc=loc, chrom2len=None)
        _compare(got, want)


def test_other_kmer_and_signal_lengths(engine):
    rb = synth_reads(30, seed=78, nb_lo=300, nb_hi=700, cg_frac=0.03)
    for k, s in ((9, 128), (21, 400)):
        got = engine.extract(rb, ["CG"], kmer_len=k, signal_len=s)
        want = _oracle(rb, ["CG"], kmer_len=k, slen=s)
        _compare(got, want, K=k)


def test_edge_cases(engine):
    from deepsignal_b200._lib import DsbError
    rb = synth_reads(12, seed=79, nb_lo=10, nb_hi=40, cg_frac=0.1)     # reads shorter than / near the k-mer
    got = engine.extract(rb, ["CG"])
    want = _oracle(rb, ["CG"])
    _compare(got, want)
    # events that run past the end of the raw signal are clipped like a python slice
    rb = synth_reads(6, seed=80, nb_lo=200, nb_hi=300, cg_frac=0.05)
    rb.raw_off = rb.raw_off.copy()
    cut = rb.raw_off[1:] - rb.raw_off[:-1]
    new_raw, offs = [], [0]
    for i in range(rb.n_reads):
        keep = int(cut[i]) - 400
        new_raw.append(rb.raw[rb.raw_off[i]:rb.raw_off[i] + keep])
        offs.append(offs[-1] + keep)
    rb.raw = np.concatenate(new_raw)
    rb.raw_off = np.asarray(offs, np.int64)
    got = engine.extract(rb, ["CG"])
    want = _oracle(rb, ["CG"])
    assert got["n_sites"] == len(want)
    assert np.array_equal(got["lens_i32"], np.asarray([w[10] for w in want]))
    # capacity too small -> error code, not silent truncation
    rb = synth_reads(20, seed=81)
    with pytest.raises(DsbError) as ei:
        engine.extract(rb, ["CG"], capacity=5)
    assert ei.value.code == -5
    # no motif hit at all
    got = engine.extract(rb, ["GGGGGGGGGGGG"])
    assert got["n_sites"] == 0


def test_wide_adc_range_uses_global_histogram(engine):
    rb = synth_reads(4, seed=82, nb_lo=300, nb_hi=500, cg_frac=0.03)
    raw = rb.raw.astype(np.int32)
    raw[::97] = 30000
    raw[5::101] = -30000
    rb.raw = raw.astype(np.int16)
    got = engine.extract(rb, ["CG"])
    want = _oracle(rb, ["CG"])
    _compare(got, want)

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


def test_very_long_slices_are_bit_identical(engine):
    """Stalls of hundreds to tens of thousands of samples: summed by the whole warp (leaf blocks of numpy's pairwise tree in
    parallel, combined in tree order) and, beyond the leaf scratch, serially -- np.mean / np.std must still match bit for bit."""
    force = {30: 129, 31: 128, 60: 1000, 61: 257, 90: 5000, 120: 20000, 150: 40000}
    rb = synth_reads(2, seed=83, nb_lo=180, nb_hi=220, cg_frac=0.08, force_lens=force)
    got = engine.extract(rb, ["CG"])
    want = _oracle(rb, ["CG"])
    _compare(got, want)
    lens = np.asarray([w[10] for w in want])
    assert lens.max() == 40000 and (lens == 20000).any() and (lens == 5000).any() and (lens == 129).any()
    assert np.array_equal(got["means64"], np.asarray([w[8] for w in want]), equal_nan=True)
    assert np.array_equal(got["stds64"], np.asarray([w[9] for w in want]), equal_nan=True)


def test_wide_adc_range_uses_global_histogram(engine):
    rb = synth_reads(4, seed=82, nb_lo=300, nb_hi=500, cg_frac=0.03)
    raw = rb.raw.astype(np.int32)
    raw[::97] = 30000
    raw[5::101] = -30000
    rb.raw = raw.astype(np.int16)
    got = engine.extract(rb, ["CG"])
    want = _oracle(rb, ["CG"])
    _compare(got, want)


def test_cfg2_full_size_4000_reads(engine):
    """BASELINE configs[1] at its stated size: 4,000 synthetic reads (~117 k CpG sites).  The NumPy oracle needs ~50 ms per
    read, so it checks a random 300-read subset site for site; the whole batch is checked through size-independent
    properties: the result of the 4,000-read call equals the concatenation of four 1,000-read calls (batch invariance),
    integer features are consistent with the event table, and every window obeys the reference's three-branch rule."""
    rb = synth_reads(4000, seed=20260102)
    got = engine.extract(rb, ["CG"])
    n = got["n_sites"]
    assert n > 100000
    # (a) batch invariance
    off = 0
    for lo in range(0, 4000, 1000):
        part = engine.extract(rb.slice(lo, lo + 1000), ["CG"])
        m = part["n_sites"]
        sl = slice(off, off + m)
        assert np.array_equal(part["site_read"] + lo, got["site_read"][sl])
        for k in ("site_loc", "pos", "pos_in_strand", "kmer", "lens_i32", "window"):
            assert np.array_equal(part[k], got[k][sl]), k
        for k in ("means64", "stds64", "signals64"):
            assert np.array_equal(part[k], got[k][sl], equal_nan=True), k
        off += m
    assert off == n
    # (b) integer features against the event table
    r, loc = got["site_read"].astype(np.int64), got["site_loc"].astype(np.int64)
    base_idx = rb.ev_off[r][:, None] + loc[:, None] + np.arange(-8, 9)[None, :]
    code = np.full(256, 255, np.int32)
    for ch, c in CODE.items():
        code[ord(ch)] = c
    assert np.array_equal(got["kmer"], code[rb.ev_base[base_idx]])
    assert np.all(got["kmer"][:, 8] == CODE["C"]) and np.all(got["kmer"][:, 9] == CODE["G"])
    assert np.array_equal(got["lens_i32"], rb.ev_len[base_idx])          # synthetic events lie inside the raw signal
    assert np.all(np.diff(got["site_read"]) >= 0)                       # ordered by read, then by location
    same = np.diff(got["site_read"]) == 0
    assert np.all(np.diff(got["site_loc"])[same] > 0)
    # (c) the window rule (extract_features.py:154-190) on every site
    tot = got["lens_i32"].sum(axis=1)
    mid = got["lens_i32"][:, 8]
    mode = np.where(tot < 360, 0, np.where(mid >= 360, 1, 2))
    assert np.array_equal(got["window"][:, 0], mode)
    assert np.all(got["window"][mode == 0, 2] == tot[mode == 0]) and np.all(got["window"][mode != 0, 2] == 360)
    assert set(np.unique(mode).tolist()) == {0, 1, 2}
    short = np.nonzero(mode == 0)[0][:2000]
    for i in short[:200]:
        assert np.all(got["signals64"][i, tot[i]:] == 0.0)
    # (d) oracle on a random subset of reads
    rng = np.random.default_rng(4)
    pick = np.sort(rng.choice(4000, size=300, replace=False))
    for i in pick[:300]:
        want = fo.extract_read(rb.read_dict(int(i)), ["CG"], 0, CONTIGS)
        rows = np.nonzero(got["site_read"] == i)[0]
        assert len(rows) == len(want)
        if not want:
            continue
        assert np.array_equal(got["pos"][rows], np.asarray([w[1] for w in want]))
        assert np.array_equal(got["pos_in_strand"][rows], np.asarray([w[3] for w in want]))
        assert np.array_equal(got["lens_i32"][rows], np.asarray([w[9] for w in want]))
        assert np.array_equal(got["means64"][rows], np.asarray([w[7] for w in want]))
        assert np.array_equal(got["stds64"][rows], np.asarray([w[8] for w in want]))
        assert np.array_equal(got["signals64"][rows], np.asarray([w[10] for w in want]))


def test_positions_filter_matches_reference_rule(engine):
    """--positions (extract_features.py:262-263): only sites whose chrom||pos||strand key is listed survive."""
    from deepsignal_b200.extract_features import extract_batch_arrays, key_sep
    rb = synth_reads(30, seed=83, nb_lo=400, nb_hi=900, cg_frac=0.03)
    everything = _oracle(rb, ["CG"])
    keys = [key_sep.join([w[1], str(w[2]), w[3]]) for w in everything]
    positions = set(keys[::3]) | {key_sep.join(["chrNope", "12", "+"])}
    arr, info = extract_batch_arrays(engine, rb, ["CG"], 0, 17, 360, "mad", positions)
    want = []
    for i in range(rb.n_reads):
        want += fo.extract_read(rb.read_dict(i), ["CG"], 0, CONTIGS, 17, 360, 1, positions)
    assert 0 < len(want) < len(everything)
    assert arr["n_sites"] == len(want) == len(info)
    assert [x.split("\t")[:3] for x in info] == [[w[0], str(w[1]), w[2]] for w in want]
    assert np.array_equal(arr["lens_i32"], np.asarray([w[9] for w in want]))
    assert np.array_equal(arr["signals64"], np.asarray([w[10] for w in want]))


def test_unusable_reads_are_dropped_and_counted(engine):
    """A read whose normalisation is undefined — no raw samples, non-positive scaling, or MAD = 0 (constant signal, numpy
    would print inf / nan features) — yields no features and counts as a failed read; its neighbours are unaffected."""
    from deepsignal_b200.extract_features import _extract_features
    rb = synth_reads(8, seed=84, nb_lo=300, nb_hi=500, cg_frac=0.04)
    clean = _oracle(rb, ["CG"])
    bad = synth_reads(8, seed=84, nb_lo=300, nb_hi=500, cg_frac=0.04)
    bad.scaling = bad.scaling.copy()
    bad.scaling[2] = 0.0                                   # non-physical scaling
    raw = bad.raw.copy()
    raw[bad.raw_off[5]:bad.raw_off[6]] = 480               # constant signal: MAD = 0
    bad.raw = raw
    feats, err = _extract_features(bad, None, None, "mad", ["CG"], 0, CONTIGS, 17, 360, 1, None, engine=engine)
    assert err == 2
    want = [w for w in clean if w[0] not in (2, 5)]
    assert len(feats) == len(want) > 0
    assert [f[1] for f in feats] == [w[2] for w in want]
    assert all(np.array_equal(f[10], w[11]) for f, w in zip(feats, want))
    got = engine.extract(bad, ["CG"])
    assert not np.any(got["site_read"] == 2)                # the kernels emit nothing for a read without a LUT
    assert not np.isfinite(got["read_scale"][2]) and got["read_scale"][5] == 0.0

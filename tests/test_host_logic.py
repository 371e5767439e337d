"""CPU tests of the host-side mirror of the reference interface (feature-file parsing, result
formatting, CLI flags, read sharding) — no GPU and no oracle compute involved except as data."""
import io
import os
import sys

import numpy as np
import pytest

from deepsignal_b200 import call_modifications as cm
from deepsignal_b200 import extract_features as ef
from deepsignal_b200 import parallel
from oracle import features_oracle as fo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class ListQ(list):
    def put(self, x):
        self.append(x)


def test_read_features_file_matches_reference(golden):
    a, m = golden
    q = ListQ()
    cm._read_features_file(os.path.join(GOLDEN, "features_small.tsv"), q, f5_batch_num=2)
    assert q[-1] == "kill"
    assert [len(b[0]) for b in q[:-1]] == m["tsv_batches"]
    b0 = q[0]
    assert b0[0] == m["tsv_b0_sampleinfo"]
    assert np.array_equal(np.asarray(b0[1], np.int32), a["tsv_b0_kmers"])
    assert np.array_equal(np.asarray(b0[2]), a["tsv_b0_means"])
    assert np.array_equal(np.asarray(b0[4]), a["tsv_b0_lens"])
    assert np.array_equal(np.asarray(b0[5]), a["tsv_b0_signals"])


def test_call_mods_formatting_matches_reference(golden):
    """The reference's _call_mods with a fake session gave golden pred_str; ours with a fake engine must too."""
    a, m = golden
    q = ListQ()
    cm._read_features_file(os.path.join(GOLDEN, "features_small.tsv"), q, f5_batch_num=2)
    fake = a["call_fake_probs"]

    class FakeEngine(object):
        def forward(self, kmer, means, stds, lens, signals, precision="fp16"):
            n = len(kmer)
            assert means.dtype == np.float32 and lens.dtype == np.float32 and signals.shape == (n, 360)
            return fake[:n], np.argmax(fake[:n], axis=1).astype(np.int32)

    model = cm._ModelHandle(17, 360, 2)
    pred_str, acc, nb = cm._call_mods(q[0], FakeEngine(), model, 0.001, 512)
    assert pred_str == m["call_pred_str"]
    assert nb == 1 and 0.0 <= acc <= 1.0
    cols = pred_str[0].split("\t")
    assert len(cols) == 10 and cols[6] == "0.5" and cols[8] == "0"      # tie -> class 0


def test_features_to_str_matches_reference(golden, golden_reads):
    _, m = golden
    from deepsignal_b200.synth import CONTIGS
    feats = []
    for i in range(golden_reads.n_reads):
        feats += fo.extract_read(golden_reads.read_dict(i), ["CG"], 0, dict(CONTIGS))
    assert [ef._features_to_str(f) for f in feats[:8]] == m["extract_tsv"]
    assert len(m["extract_tsv"][0].split("\t")) == 12


def test_reads_container_roundtrip(tmp_path, golden_reads):
    p = str(tmp_path / "reads.npz")
    ef.save_reads(p, golden_reads)
    rb = ef.load_reads(p, chrom2len={"chrSynA": 5_000_000})
    assert rb.n_reads == golden_reads.n_reads and rb.read_id == golden_reads.read_id
    assert np.array_equal(rb.raw, golden_reads.raw) and np.array_equal(rb.ev_off, golden_reads.ev_off)
    assert set(rb.chrom_len.tolist()) <= {5_000_000, -1}
    rb2 = ef.load_reads(str(tmp_path))
    assert rb2.n_reads == rb.n_reads and np.all(rb2.chrom_len == -1)
    part = ef._concat_reads([golden_reads.slice(0, 2), golden_reads.slice(2, golden_reads.n_reads)])
    assert np.array_equal(part.raw_off, golden_reads.raw_off) and np.array_equal(part.ev_start, golden_reads.ev_start)


def test_cli_flags_and_defaults(monkeypatch):
    from deepsignal_b200 import deepsignal as cli
    seen = {}
    monkeypatch.setattr(cli, "main_extraction", lambda args: seen.update(extract=args))
    monkeypatch.setattr(cli, "main_call_mods", lambda args: seen.update(call=args))
    # set_defaults binds the functions at parser-build time, so patch before main() builds it
    monkeypatch.setattr(sys, "argv", ["deepsignal", "extract", "-i", "d", "-o", "o.tsv"])
    cli.main()
    e = seen["extract"]
    assert (e.kmer_len, e.cent_signals_len, e.motifs, e.mod_loc, e.normalize_method, e.methy_label) == (17, 360, "CG", 0, "mad", 1)
    assert (e.corrected_group, e.basecall_subgroup, e.recursively, e.is_dna, e.w_is_dir, e.w_batch_num, e.f5_batch_num, e.nproc) == \
        ("RawGenomeCorrected_000", "BaseCalled_template", "yes", "yes", "no", 200, 50, 1)
    monkeypatch.setattr(sys, "argv", ["deepsignal", "call_mods", "-i", "f.tsv", "-m", "m.ckpt", "-o", "r.tsv"])
    cli.main()
    c = seen["call"]
    assert (c.kmer_len, c.cent_signals_len, c.batch_size, c.learning_rate, c.class_num, c.is_gpu) == (17, 360, 512, 0.001, 2, "no")
    assert (c.is_cnn, c.is_rnn, c.is_base, c.f5_batch_num, c.precision) == ("yes", "yes", "yes", 50, "fp16x2")


def test_shard_ranges_cover_without_overlap():
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world_size, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    n_reads = 11
    lo, hi = parallel.shard_range(n_reads, rank, world_size)
    # every rank "calls" its reads: two result lines per read, reads stay contiguous
    local = [(r, ["read%d\tline%d" % (r, k) for k in range(2)]) for r in range(lo, hi)]
    merged = parallel.gather_in_order(local, rank, world_size)
    if rank == 0:
        with open(os.path.join(out_dir, "merged.txt"), "w") as fh:
            for lines in merged:
                fh.write("\n".join(lines) + "\n")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_host_gather(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = open(os.path.join(str(tmp_path), "merged.txt")).read().splitlines()
    assert got == ["read%d\tline%d" % (r, k) for r in range(11) for k in range(2)]


def test_binary_record_layout_matches_reference_struct(tmp_path):
    """1628-byte record of the reference: struct format '<17B17f17f17H360f1B'
    (scripts/generate_binary_feature_file.py:52-54)."""
    import struct
    from deepsignal_b200 import binary_format as bf
    from deepsignal_b200.synth import synth_features
    assert bf.record_dtype().itemsize == 1628 == struct.calcsize("<17B17f17f17H360f1B")
    f = synth_features(5, seed=3)
    p = str(tmp_path / "x.bin")
    bf.write_binary(p, f["kmer"], f["means"], f["stds"], f["lens"], f["signals"], [1, 0, 1, 1, 0])
    raw = open(p, "rb").read()
    assert len(raw) == 5 * 1628
    fmt = "<17B17f17f17H360f1B"
    for i in range(5):
        vals = struct.unpack(fmt, raw[i * 1628:(i + 1) * 1628])
        assert list(vals[:17]) == f["kmer"][i].tolist()
        assert np.array_equal(np.asarray(vals[17:34], np.float32), f["means"][i])
        assert list(vals[51:68]) == f["lens"][i].tolist()
        assert np.array_equal(np.asarray(vals[68:428], np.float32), f["signals"][i])
    back = bf.read_binary(p)
    assert np.array_equal(back["kmer"], f["kmer"]) and np.array_equal(back["signals"], f["signals"])
    assert back["lens"].dtype == np.float32 and back["labels"].tolist() == [1, 0, 1, 1, 0]


def test_fast_tsv_parser_matches_reference_reader(golden):
    from deepsignal_b200.fast_tsv import iter_feature_blocks
    a, m = golden
    p = os.path.join(GOLDEN, "features_small.tsv")
    blocks = list(iter_feature_blocks(p, max_samples=5, nthreads=2, block_bytes=7000))   # forces carry-over across blocks
    assert sum(b.n for b in blocks) == 12 and len(blocks) >= 3
    kmer = np.concatenate([b.kmer for b in blocks])
    means = np.concatenate([b.means for b in blocks])
    sig = np.concatenate([b.signals for b in blocks])
    lens = np.concatenate([b.lens for b in blocks])
    q = ListQ()
    cm._read_features_file(p, q, f5_batch_num=10 ** 9)
    ref = q[0]
    assert np.array_equal(kmer, np.asarray(ref[1], np.int32))
    assert np.array_equal(means, np.asarray(ref[2], np.float32))
    assert np.array_equal(lens, np.asarray(ref[4], np.float32))
    assert np.array_equal(sig, np.asarray(ref[5], np.float32))
    infos = []
    for b in blocks:
        infos += b.info_strings()
    assert infos == ref[0]
    with pytest.raises(ValueError):
        bad = os.path.join(os.path.dirname(p), "..", "..", "README.md")
        list(iter_feature_blocks(bad))


def test_fast_tsv_formatter_matches_numpy_float32_text(golden):
    from deepsignal_b200.fast_tsv import iter_feature_blocks, format_calls
    a, m = golden
    p = os.path.join(GOLDEN, "features_small.tsv")
    block = next(iter_feature_blocks(p, max_samples=10 ** 4))
    fake = a["call_fake_probs"]
    n0 = len(m["call_pred_str"])
    got = format_calls(block, np.vstack([fake, fake])[:block.n], np.argmax(np.vstack([fake, fake])[:block.n], axis=1)).decode().splitlines()
    assert got[:n0] == m["call_pred_str"]                       # byte-identical to the reference's _call_mods output
    # float32 text == str(np.float32) over many magnitudes
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.random(20000, dtype=np.float32), np.float32(10.0) ** rng.uniform(-12, 0, 20000).astype(np.float32),
                           np.asarray([0.5, 1.0, 1e-4, 9.9999e-5, 1e-8, 0.99999994, 3.4e-38, 1e-45, 0.0], np.float32)])
    p1 = vals.astype(np.float32)
    p0 = (np.float32(1.0) - p1).astype(np.float32)
    probs = np.stack([p0, p1], axis=1)
    blk = type(block)()
    n = len(vals)
    import ctypes
    blk.n, blk.kmer, blk.text = n, np.zeros((n, 17), np.int32), b"x"
    blk.text_ptr = ctypes.cast(ctypes.c_char_p(blk.text), ctypes.c_void_p).value
    blk.info_off = np.tile(np.asarray([[0, 1]], np.int64), (n, 1))
    lines = format_calls(blk, probs, np.zeros(n, np.int32)).decode().splitlines()
    s = p0 + p1
    for i in range(n):
        c = lines[i].split("\t")
        assert c[1] == str(p0[i] / s[i]) and c[2] == str(p1[i] / s[i]), (i, c, str(p0[i] / s[i]), str(p1[i] / s[i]))


def test_tf_bundle_round_trip_and_checkpoint_loader(tmp_path):
    """SSTable index + data shard written and read back; `load_checkpoint` picks a `.ckpt` prefix up."""
    from deepsignal_b200 import tf_bundle, weights as W
    ck = {k: v for k, v in list(W.random_checkpoint(3, "tf_default").items())[:40]}
    ck["modelglobal_step"] = np.asarray(7, np.int32)            # non-float entries are ignored by the loader
    prefix = str(tmp_path / "bn_17.sn_360.epoch_9.ckpt")
    tf_bundle.write_bundle(prefix, ck)
    idx = tf_bundle.read_index(prefix)
    assert set(idx) == set(ck)
    name = "modelsignalmconv_layer1/conv/kernel"
    assert idx[name]["shape"] == [1, 7, 1, 64] and idx[name]["dtype"] == tf_bundle.DT_FLOAT
    back = tf_bundle.read_bundle(prefix)
    for k, v in ck.items():
        assert back[k].dtype == np.asarray(v).dtype and np.array_equal(back[k], v), k
    loaded = W.load_checkpoint(prefix)
    assert "modelglobal_step" not in loaded and np.array_equal(loaded[name], ck[name])
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283            # CRC-32C check value
    with open(prefix + ".index", "r+b") as fh:
        fh.seek(-1, 2)
        fh.write(b"\x00")
    with pytest.raises(ValueError):
        tf_bundle.read_index(prefix)


def test_site_frequency_matches_reference_script(tmp_path):
    """call_modification_frequency: byte-identical to the reference script's output (golden from make_golden.py)."""
    import json
    from deepsignal_b200 import call_modification_frequency as cmf
    ref = json.load(open(os.path.join(GOLDEN, "freq_ref.json")))
    src = os.path.join(GOLDEN, "calls_small.tsv")
    for tag, (cf, srt, bed) in dict(plain=(0.0, False, False), sort=(0.0, True, False), bed=(0.0, True, True), cf=(0.3, False, False)).items():
        _, stats = cmf.calculate_mods_frequency([src], cf)
        out = str(tmp_path / (tag + ".txt"))
        cmf.write_sitekey2stats(stats, out, srt, bed)
        assert open(out).read() == ref[tag], tag


# ---------------------------------------------------------------------------------------------------------
# read sharding in the product entry point (call_modifications.call_mods): the multi-worker result file must be
# byte-identical to the single-worker one.  The engine is replaced by a deterministic stand-in (no GPU here).
# ---------------------------------------------------------------------------------------------------------
class _StubEngine(object):
    def __init__(self, device):
        self.device = device

    def forward(self, kmer, means, stds, lens, signals, precision="fp16x2"):
        z = (np.asarray(means, np.float32).sum(axis=1) + np.asarray(signals, np.float32)[:, :8].sum(axis=1)
             + np.asarray(lens, np.float32).sum(axis=1) * 0.01)
        p1 = (1.0 / (1.0 + np.exp(-z))).astype(np.float32)
        probs = np.stack([np.float32(1.0) - p1, p1], axis=1).astype(np.float32)
        return probs, np.argmax(probs, axis=1).astype(np.int32)

    def close(self):
        pass


def _f5_args(f5_batch_num=1):
    return (True, "RawGenomeCorrected_000", "BaseCalled_template", None, True, "mad", "CG", 0, 1, f5_batch_num, None)


def _run_call_mods(tsv, out, devices=None, py_tsv=False):
    os.environ["DSB_PY_TSV"] = "1" if py_tsv else "0"
    try:
        cm.call_mods(tsv, "unused.ckpt", out, 17, 360, 512, 0.001, 2, 2, False, True, True, True, _f5_args(), devices=devices)
    finally:
        os.environ.pop("DSB_PY_TSV", None)
    return open(out).read()


@pytest.mark.parametrize("py_tsv", [False, True])
def test_call_mods_devices_sharding_is_byte_identical(tmp_path, monkeypatch, py_tsv):
    from deepsignal_b200.synth import synth_features, write_feature_tsv
    monkeypatch.setattr(cm, "_load_engine", lambda model_path, k, s, c, r, b, device=0: _StubEngine(device))
    monkeypatch.setattr("deepsignal_b200.fast_tsv.BLOCK_SAMPLES", 128)
    tsv = str(tmp_path / "f.tsv")
    write_feature_tsv(tsv, synth_features(1500, seed=5, sites_per_read=7))
    one = _run_call_mods(tsv, str(tmp_path / "one.tsv"), None, py_tsv)
    assert len(one.splitlines()) == 1500
    for devs in ([0, 1], [0, 1, 2, 3, 4, 5, 6, 7]):
        many = _run_call_mods(tsv, str(tmp_path / "many.tsv"), devs, py_tsv)
        assert many == one, devs
    # each read's lines stay contiguous and in input order
    names = [ln.split("\t")[4] for ln in one.splitlines()]
    assert names == sorted(names)


def test_call_mods_worker_error_propagates(tmp_path, monkeypatch):
    from deepsignal_b200.synth import synth_features, write_feature_tsv

    class Boom(_StubEngine):
        def forward(self, *a, **k):
            raise RuntimeError("device fell over")
    monkeypatch.setattr(cm, "_load_engine", lambda model_path, k, s, c, r, b, device=0: Boom(device))
    tsv = str(tmp_path / "f.tsv")
    write_feature_tsv(tsv, synth_features(60, seed=5, sites_per_read=7))
    with pytest.raises(RuntimeError, match="device fell over"):
        _run_call_mods(tsv, str(tmp_path / "o.tsv"), [0, 1])


def test_feature_file_with_non_acgtn_kmer_is_rejected(tmp_path, monkeypatch):
    """reference call_modifications.py:52: base2code_dna[x] raises KeyError on a base outside ACGTN."""
    from deepsignal_b200.synth import synth_features, write_feature_tsv
    monkeypatch.setattr(cm, "_load_engine", lambda model_path, k, s, c, r, b, device=0: _StubEngine(device))
    tsv = str(tmp_path / "f.tsv")
    write_feature_tsv(tsv, synth_features(8, seed=5))
    txt = open(tsv).read().splitlines()
    w = txt[3].split("\t")
    w[6] = w[6][:5] + "X" + w[6][6:]
    txt[3] = "\t".join(w)
    open(tsv, "w").write("\n".join(txt) + "\n")
    with pytest.raises((KeyError, ValueError)):
        _run_call_mods(tsv, str(tmp_path / "o.tsv"), None, py_tsv=True)
    with pytest.raises((KeyError, ValueError)):
        _run_call_mods(tsv, str(tmp_path / "o2.tsv"), None, py_tsv=False)


def _gloo_call_mods_worker(rank, world_size, port, tsv, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world_size))
    from deepsignal_b200 import call_modifications as cm2
    cm2._load_engine = lambda model_path, k, s, c, r, b, device=0: _StubEngine(device)
    cm2.call_mods(tsv, "unused.ckpt", out, 17, 360, 512, 0.001, 2, 2, False, True, True, True, _f5_args())


def test_call_mods_two_rank_process_sharding_is_byte_identical(tmp_path, monkeypatch):
    """torchrun-style: two ranks over gloo, rank 0 writes; same bytes as the single-process run."""
    import socket
    import torch.multiprocessing as mp
    from deepsignal_b200.synth import synth_features, write_feature_tsv
    monkeypatch.setattr(cm, "_load_engine", lambda model_path, k, s, c, r, b, device=0: _StubEngine(device))
    tsv = str(tmp_path / "f.tsv")
    write_feature_tsv(tsv, synth_features(900, seed=6, sites_per_read=9))
    one = _run_call_mods(tsv, str(tmp_path / "one.tsv"))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "two.tsv")
    mp.spawn(_gloo_call_mods_worker, args=(2, port, tsv, out), nprocs=2, join=True)
    assert open(out).read() == one


def _crc32c_bitwise(data):
    """Second, table-free implementation of CRC-32C (Castagnoli, reflected polynomial 0x82F63B78)."""
    crc = 0xFFFFFFFF
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xFFFFFFFF


def test_tf_bundle_snappy_blocks_crc_and_prefix_compression(tmp_path):
    """Index blocks as TF's table builder writes them: prefix-shared keys (restart interval 16), several data blocks,
    Snappy-compressed (type 1) with a masked crc32c trailer.  The Snappy decoder is also fed hand-assembled streams."""
    from deepsignal_b200 import tf_bundle, weights as W
    # --- hand-assembled Snappy streams (format description: literal 00, copy-1 01, copy-2 10)
    assert tf_bundle.snappy_decompress(bytes([5, 4 << 2]) + b"hello") == b"hello"
    # "abcd" literal, then copy-1: len 4+((tag>>2)&7)=8, offset 4 -> overlapping repeat
    assert tf_bundle.snappy_decompress(bytes([12, 3 << 2]) + b"abcd" + bytes([(4 << 2) | 1, 4])) == b"abcdabcdabcd"
    # copy-2: tag = ((len-1)<<2)|2, offset LE16
    assert tf_bundle.snappy_decompress(bytes([10, 4 << 2]) + b"01234" + bytes([(4 << 2) | 2, 5, 0])) == b"0123401234"
    # long literal with a 1-byte length field (tag 60)
    blob = bytes(range(200))
    assert tf_bundle.snappy_decompress(tf_bundle._put_varint(200) + bytes([60 << 2, 199]) + blob) == blob
    with pytest.raises(ValueError):
        tf_bundle.snappy_decompress(bytes([4, (4 << 2) | 2, 9, 0]))         # copy before any output
    rng = np.random.default_rng(0)
    for payload in (b"", b"a", b"model/signalm/" * 300, bytes(rng.integers(0, 256, 5000, dtype=np.uint8)),
                    (b"incp_layer" + bytes(rng.integers(0, 4, 3000, dtype=np.uint8))) * 3):
        assert tf_bundle.snappy_decompress(tf_bundle.snappy_compress(payload)) == payload
    # --- crc32c: table implementation vs the bitwise one
    for payload in (b"", b"123456789", bytes(rng.integers(0, 256, 4097, dtype=np.uint8))):
        assert tf_bundle.crc32c(payload) == _crc32c_bitwise(payload)
    # --- compressed, multi-block bundle with the full variable list of the model (names share long prefixes)
    ck = W.random_checkpoint(3, "tf_default")
    small = {k: (v if v.size < 5000 else v.reshape(-1)[:64].copy()) for k, v in ck.items()}
    prefix = str(tmp_path / "c.ckpt")
    tf_bundle.write_bundle(prefix, small, compress=True, block_entries=37)
    raw_prefix = str(tmp_path / "r.ckpt")
    tf_bundle.write_bundle(raw_prefix, small, compress=False, block_entries=37)
    assert os.path.getsize(prefix + ".index") < 0.6 * os.path.getsize(raw_prefix + ".index")   # the names compress well
    for pfx in (prefix, raw_prefix):
        back = tf_bundle.read_bundle(pfx)
        assert set(back) == set(small)
        for k, v in small.items():
            assert np.array_equal(back[k], v), k
    # a flipped payload bit is caught by the block crc
    data = bytearray(open(prefix + ".index", "rb").read())
    data[10] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="crc32c"):
        tf_bundle.read_index(prefix)


def test_block_from_info_formats_like_call_mods():
    """The reads path of call_mods formats its result lines with the C++ codec from in-memory samples: same bytes as the
    Python formatting of _call_mods (reference call_modifications.py:183-190)."""
    from deepsignal_b200.fast_tsv import block_from_info, format_calls
    rng = np.random.default_rng(3)
    n = 500
    info = ["chr%d\t%d\t%s\t%d\tread_%d\tt" % (i % 3, 1000 + i, "+-"[i % 2], -1 if i % 5 else 7 * i, i // 9) for i in range(n)]
    kmer = rng.integers(0, 5, (n, 17)).astype(np.int32)
    eng = _StubEngine(0)
    z = [rng.random((n, 17), dtype=np.float32) for _ in range(3)]
    sig = rng.random((n, 360), dtype=np.float32)
    probs, pred = eng.forward(kmer, z[0], z[1], z[2], sig)
    got = format_calls(block_from_info(info, kmer), probs, pred).decode().splitlines()
    want, _, _ = cm._call_mods((info, kmer, z[0], z[1], z[2], sig, [1] * n), eng, cm._ModelHandle(17, 360, 2), 0.001, 512)
    assert got == want
    bad = kmer.copy()
    bad[3, 4] = 255
    with pytest.raises(KeyError):
        block_from_info(info, bad)


def test_synth_force_lens_does_not_disturb_the_random_stream():
    """`force_lens` only overwrites dwell times: every read generated without it must stay byte-identical (the golden
    fixtures and the bench workloads are seeded)."""
    from deepsignal_b200.synth import synth_reads
    a = synth_reads(3, seed=5, nb_lo=100, nb_hi=150)
    b = synth_reads(3, seed=5, nb_lo=100, nb_hi=150, force_lens={})
    c = synth_reads(3, seed=5, nb_lo=100, nb_hi=150, force_lens={7: 300})
    assert np.array_equal(a.raw, b.raw) and np.array_equal(a.ev_len, b.ev_len)
    assert np.array_equal(a.ev_base, c.ev_base) and c.ev_len[7] == 300
    keep = np.ones(len(a.ev_len), bool)
    keep[7] = False
    assert np.array_equal(a.ev_len[keep], c.ev_len[keep])


def test_snappy_codec_against_a_real_snappy_implementation():
    """deepsignal_b200/tf_bundle.py carries a pure-Python Snappy (TF's table builder compresses index blocks with it);
    pyarrow links the real library: each side must read what the other writes."""
    pa = pytest.importorskip("pyarrow")
    if not pa.Codec.is_available("snappy"):
        pytest.skip("pyarrow built without snappy")
    from deepsignal_b200 import tf_bundle
    codec = pa.Codec("snappy")
    rng = np.random.default_rng(3)
    names = b"".join(b"modelsignalmincp_layer%d/modelsignalm%dbranch3_1x3/bn1/moving_variance\x00" % (i, i) for i in range(1, 12))
    cases = [b"a", b"abc" * 1000, bytes(rng.integers(0, 256, 70000, dtype=np.uint8)), names * 7,
             bytes(rng.integers(0, 4, 200000, dtype=np.uint8))]
    for d in cases:
        assert tf_bundle.snappy_decompress(codec.compress(d, asbytes=True)) == d
        assert codec.decompress(tf_bundle.snappy_compress(d), decompressed_size=len(d), asbytes=True) == d

"""End-to-end drop-in checks on the GPU: the `deepsignal extract` / `deepsignal call_mods` commands
produce the reference's file formats; extraction text is identical to the oracle's, called
probabilities agree with the float64 oracle within the path's tolerance."""
import os
import subprocess
import sys

import numpy as np
import pytest

from deepsignal_b200 import weights as W
from deepsignal_b200.synth import CONTIGS, synth_features, synth_reads, write_feature_tsv
from deepsignal_b200.extract_features import save_reads
from oracle import features_oracle as fo
from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    r = subprocess.run([sys.executable, "-m", "deepsignal_b200.deepsignal"] + list(args), cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_extract_cli_text_identical_to_oracle(tmp_path):
    rb = synth_reads(24, seed=31, nb_lo=300, nb_hi=900, cg_frac=0.03)
    reads_p = str(tmp_path / "reads.npz")
    save_reads(reads_p, rb)
    ref_fa = str(tmp_path / "ref.fa")
    with open(ref_fa, "w") as fh:
        for name, n in CONTIGS.items():
            fh.write(">%s synthetic\n" % name)
            fh.write("N" * 10 + "\n")          # lengths come from the file: make them differ from CONTIGS
    out = str(tmp_path / "features.tsv")
    _run("extract", "-i", reads_p, "-o", out, "--reference_path", ref_fa, "--motifs", "CG", "--mod_loc", "0")
    want = []
    for i in range(rb.n_reads):
        want += [fo.features_to_str(f) for f in fo.extract_read(rb.read_dict(i), ["CG"], 0, {k: 10 for k in CONTIGS})]
    got = open(out).read().splitlines()
    assert len(got) == len(want) and len(got) > 50
    assert got == want                     # every float prints identically: the float64 features are bit-exact


def test_call_mods_cli_from_feature_file_and_from_reads(tmp_path):
    ck = W.random_checkpoint(17, "stress")
    model_p = str(tmp_path / "bn_17.sn_360.epoch_0.ckpt.npz")
    W.save_npz(model_p, ck)
    f = synth_features(300, seed=41, sites_per_read=7)
    feats_p = str(tmp_path / "features.tsv")
    write_feature_tsv(feats_p, f)
    res_p = str(tmp_path / "calls.tsv")
    _run("call_mods", "-i", feats_p, "-m", model_p, "-o", res_p, "--precision", "fp32", "--f5_batch_num", "3")
    lines = open(res_p).read().splitlines()
    assert len(lines) == 300
    ref, ref_pred = mo.forward(ck, f["kmer"], f["means"], f["stds"], f["lens"], f["signals"])
    p0n, p1n = mo.host_normalise(ref)
    for i, ln in enumerate(lines):
        c = ln.split("\t")
        assert len(c) == 10
        assert c[4] == "read_%06d" % f["read_index"][i] and c[9] == "".join("ACGTN"[k] for k in f["kmer"][i])
        assert abs(float(c[6]) - p0n[i]) <= 2e-4 and abs(float(c[7]) - p1n[i]) <= 2e-4
        if abs(ref[i, 0] - ref[i, 1]) > 1e-3:
            assert int(c[8]) == ref_pred[i]
    # fast5-directory mode of the reference == packed reads here: extraction and forward fused on the GPU
    rb = synth_reads(10, seed=42, nb_lo=300, nb_hi=700, cg_frac=0.03)
    reads_p = str(tmp_path / "reads.npz")
    save_reads(reads_p, rb)
    res2 = str(tmp_path / "calls2.tsv")
    _run("call_mods", "-i", reads_p, "-m", model_p, "-o", res2, "--precision", "fp32")
    feats = []
    for i in range(rb.n_reads):
        feats += fo.extract_read(rb.read_dict(i), ["CG"])
    lines2 = open(res2).read().splitlines()
    assert len(lines2) == len(feats) and len(feats) > 20
    kmer = np.asarray([[fo_code for fo_code in map("ACGTN".index, ft[6])] for ft in feats], np.int32)
    ref2, _ = mo.forward(ck, kmer, np.asarray([ft[7] for ft in feats]), np.asarray([ft[8] for ft in feats]),
                         np.asarray([ft[9] for ft in feats], np.float64), np.asarray([ft[10] for ft in feats]))
    q0, _ = mo.host_normalise(ref2)
    for i, ln in enumerate(lines2):
        c = ln.split("\t")
        assert c[1] == str(feats[i][1]) and c[4] == feats[i][4] and c[3] == "-1"
        assert abs(float(c[6]) - q0[i]) <= 2e-4


def test_call_mods_binary_records_path(tmp_path):
    """the reference's 1628-byte binary sample records as a zero-parse call_mods input."""
    from deepsignal_b200 import binary_format as bf
    ck = W.random_checkpoint(17, "stress")
    model_p = str(tmp_path / "m.ckpt.npz")
    W.save_npz(model_p, ck)
    f = synth_features(700, seed=43)
    bin_p = str(tmp_path / "samples.bin")
    bf.write_binary(bin_p, f["kmer"], f["means"], f["stds"], f["lens"], f["signals"], np.ones(700, np.int64))
    res_p = str(tmp_path / "calls.tsv")
    _run("call_mods", "-i", bin_p, "-m", model_p, "-o", res_p, "--precision", "fp16x2")
    lines = open(res_p).read().splitlines()
    assert len(lines) == 700
    ref, ref_pred = mo.forward(ck, f["kmer"], f["means"], f["stds"], f["lens"], f["signals"])
    p0n, p1n = mo.host_normalise(ref)
    for i, ln in enumerate(lines):
        c = ln.split("\t")
        assert len(c) == 10 and c[:6] == ["."] * 6 and c[9] == "".join("ACGTN"[k] for k in f["kmer"][i])
        assert abs(float(c[6]) - p0n[i]) <= 2e-3 and abs(float(c[7]) - p1n[i]) <= 2e-3


def test_call_mods_devices_flag_matches_single_gpu(tmp_path):
    """--devices shards the reads over several GPUs; the result file must be byte-identical to the 1-GPU run.
    Needs two GPUs for the real thing; on one GPU the same device is used twice (two engines, two worker threads)."""
    import torch
    ck = W.random_checkpoint(17, "stress")
    model_p = str(tmp_path / "m.ckpt.npz")
    W.save_npz(model_p, ck)
    rb = synth_reads(300, seed=44, nb_lo=300, nb_hi=700, cg_frac=0.03)
    reads_p = str(tmp_path / "reads.npz")
    save_reads(reads_p, rb)
    one = str(tmp_path / "one.tsv")
    _run("call_mods", "-i", reads_p, "-m", model_p, "-o", one, "--f5_batch_num", "1")
    devs = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    two = str(tmp_path / "two.tsv")
    _run("call_mods", "-i", reads_p, "-m", model_p, "-o", two, "--f5_batch_num", "1", "--devices", devs)
    a, b = open(one).read(), open(two).read()
    assert len(a.splitlines()) > 500 and a == b


def test_site_frequency_gpu_matches_reference_script(tmp_path):
    """the (chrom, pos)-keyed aggregation as a GPU segmented reduce must reproduce the reference script's text
    (golden produced by tests/golden/make_golden.py from scripts/call_modification_frequency.py) byte for byte, and agree
    with the host NumPy path on a larger synthetic result file."""
    import json
    from deepsignal_b200 import call_modification_frequency as cmf
    golden = os.path.join(ROOT, "tests", "golden")
    ref = json.load(open(os.path.join(golden, "freq_ref.json")))
    src = os.path.join(golden, "calls_small.tsv")
    for tag, (cf, srt, bed) in dict(plain=(0.0, False, False), sort=(0.0, True, False), bed=(0.0, True, True), cf=(0.3, False, False)).items():
        _, stats = cmf.calculate_mods_frequency([src], cf, device=0)
        out = str(tmp_path / (tag + ".txt"))
        cmf.write_sitekey2stats(stats, out, srt, bed)
        assert open(out).read() == ref[tag], tag
    # larger: 200k records over 3 contigs, ~20x coverage, both strands
    rng = np.random.default_rng(9)
    n = 200000
    chrom = rng.choice(["chrA", "chrB", "chr10"], size=n)
    pos = rng.integers(0, 3000, size=n)
    p1 = rng.random(n).round(6)
    big = str(tmp_path / "big.tsv")
    with open(big, "w") as fh:
        for i in range(n):
            fh.write("\t".join([chrom[i], str(pos[i]), "+", str(pos[i]), "r%d" % (i % 977), "t", str(1.0 - p1[i]), str(p1[i]),
                                str(int(p1[i] > 0.5)), "ACGTACGTCGACGTACG"]) + "\n")
    for cf in (0.0, 0.4):
        _, host = cmf.calculate_mods_frequency([big], cf)
        _, dev = cmf.calculate_mods_frequency([big], cf, device=0)
        for k in host:
            assert np.array_equal(host[k], dev[k]), (cf, k)

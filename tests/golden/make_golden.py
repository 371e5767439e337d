#!/usr/bin/env python
"""Generate tests/golden/*.npz|json by executing the REFERENCE's own functions.

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_golden.py

The reference imports h5py / statsmodels / tensorflow at module scope and uses the
``np.int`` / ``np.float`` aliases numpy>=1.24 removed, so this script installs
``sys.modules`` stubs (robust.mad restated from statsmodels, see
oracle/features_oracle.py) and the two aliases, then calls the reference code
unmodified.  The fast5 accessors (``_get_label_raw``, ``_get_scaling_of_a_read``,
``_get_alignment_info_from_fast5``) are replaced by readers of our in-memory
synthetic reads so that the reference's ``_extract_features`` loop itself
(extract_features.py:215-286) produces the golden feature tuples.
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("DEEPSIGNAL_REFERENCE", "/root/reference")


def install_stubs():
    np.int = int
    np.float = float
    h5py = types.ModuleType("h5py")
    sm = types.ModuleType("statsmodels")
    robust = types.ModuleType("statsmodels.robust")
    robust.mad = lambda a: np.median(np.fabs(np.asarray(a) - np.median(a)) / 0.6744897501960817)
    sm.robust = robust
    tf = types.ModuleType("tensorflow")
    contrib = types.ModuleType("tensorflow.contrib")
    rnn = types.ModuleType("tensorflow.contrib.rnn")
    rnn.LSTMCell = rnn.GRUCell = object
    layers = types.ModuleType("tensorflow.contrib.layers")
    layers.batch_norm = object
    contrib.rnn, contrib.layers, contrib.framework = rnn, layers, types.ModuleType("tensorflow.contrib.framework")
    tf.contrib = contrib
    tf.float32 = "float32"
    for name, mod in [("h5py", h5py), ("statsmodels", sm), ("statsmodels.robust", robust),
                      ("tensorflow", tf), ("tensorflow.contrib", contrib), ("tensorflow.contrib.rnn", rnn),
                      ("tensorflow.contrib.layers", layers),
                      ("tensorflow.contrib.framework", contrib.framework)]:
        sys.modules[name] = mod
    sys.path.insert(0, REF)


def main():
    install_stubs()
    from deepsignal import extract_features as ref_ef
    from deepsignal import call_modifications as ref_cm
    from deepsignal.utils import process_utils as ref_pu
    from deepsignal_b200.synth import synth_reads, synth_features, write_feature_tsv

    out = {}
    meta = {}

    # ---- a-1/a-2: rescale + normalise ------------------------------------------------
    rb = synth_reads(6, seed=4242, nb_lo=260, nb_hi=520, cg_frac=0.04, long_stall_frac=0.0)
    for i in range(rb.n_reads):
        rd = rb.read_dict(i)
        scaling = rd["range"] / rd["digitisation"]
        resc = ref_ef._rescale_signals(rd["raw"], scaling, rd["offset"])
        out["rescaled_%d" % i] = resc
        out["norm_mad_%d" % i] = ref_ef._normalize_signals(resc, "mad")
        out["norm_zscore_%d" % i] = ref_ef._normalize_signals(resc, "zscore")

    # ---- a-8: window rule on curated length lists (no branch-2: unseeded RNG) ---------
    rng = np.random.default_rng(99)
    cases = []
    for total_hint in (5, 9, 21, 22, 30, 60):
        for trial in range(6):
            lens = (1 + rng.geometric(1.0 / total_hint, size=17)).tolist()
            if trial == 4:
                lens[:8] = [1] * 8           # short left side -> right extension
            if trial == 5:
                lens[9:] = [1] * 8           # short right side -> left extension
            lens[8] = min(lens[8], 359)
            cases.append(lens)
    cases.append([21] * 17 + [])             # T = 357 < 360
    cases.append([21] * 16 + [24])           # T = 360 exactly
    cases.append([0] * 8 + [300] + [10] * 8)  # empty left side
    win_out = []
    for lens in cases:
        vals = [rng.normal(size=n).round(6) for n in lens]
        cent = ref_ef._get_central_signals(vals, 360)
        win_out.append(np.asarray(cent, dtype=np.float64))
        out["win_vals_%d" % (len(win_out) - 1)] = np.concatenate(vals) if sum(lens) else np.zeros(0)
    out["win_lens"] = np.asarray(cases, dtype=np.int64)
    out["win_cent"] = np.stack(win_out)

    # ---- a-4: motifs ------------------------------------------------------------------
    motif_cases = ["CG", "CHG", "CHH,CG", "GATC", "CCWGG", "A", "NCG"]
    meta["motif_seqs"] = {m: sorted(ref_pu.get_motif_seqs(m, True)) for m in motif_cases}
    seq = rb.read_dict(0)["ev_base"].decode()
    meta["motif_scan_seq"] = seq
    meta["motif_scan"] = {m + "|" + str(loc): ref_pu.get_refloc_of_methysite_in_motif(
        seq, set(ref_pu.get_motif_seqs(m, True)), loc) for m, loc in
        [("CG", 0), ("CG", 1), ("CHG", 0), ("CCWGG", 1), ("A", 0)]}

    # ---- full reference _extract_features loop on in-memory reads --------------------
    store = {"fn%d" % i: rb.read_dict(i) for i in range(rb.n_reads)}

    def fake_label_raw(fn, cg, sg):
        rd = store[fn]
        starts = [int(s) + rd["read_start_rel_to_raw"] for s in rd["ev_start"]]
        lengths = np.asarray(rd["ev_len"]).astype(np.int)
        base = [chr(c) for c in rd["ev_base"]]
        return rd["raw"], list(zip(starts, lengths, base))

    ref_ef._get_label_raw = fake_label_raw
    ref_ef._get_scaling_of_a_read = lambda fn: (store[fn]["range"] / store[fn]["digitisation"], store[fn]["offset"])
    ref_ef._get_alignment_info_from_fast5 = lambda fn, cg, sg: (
        store[fn]["read_id"], store[fn]["strand"], store[fn]["alignstrand"], store[fn]["chrom"],
        store[fn]["chrom_start"])
    from deepsignal_b200.synth import CONTIGS
    feats, err = ref_ef._extract_features(list(store), "RawGenomeCorrected_000", "BaseCalled_template", "mad",
                                          ref_pu.get_motif_seqs("CG", True), 0, dict(CONTIGS), 17, 360, 1, None)
    assert err == 0, err
    assert len(feats) > 20, len(feats)
    meta["extract_n"] = len(feats)
    meta["extract_info"] = [[f[0], int(f[1]), f[2], int(f[3]), f[4], f[5], f[6]] for f in feats]
    out["extract_means"] = np.asarray([f[7] for f in feats], dtype=np.float64)
    out["extract_stds"] = np.asarray([f[8] for f in feats], dtype=np.float64)
    out["extract_lens"] = np.asarray([f[9] for f in feats], dtype=np.int64)
    out["extract_cent"] = np.asarray([np.asarray(f[10], dtype=np.float64) for f in feats])
    meta["extract_tsv"] = [ref_ef._features_to_str(f) for f in feats[:8]]
    for k in ("raw", "raw_off", "scaling", "offset", "ev_start", "ev_len", "ev_base", "ev_off",
              "chrom_start", "strand_minus", "chrom_len"):
        out["reads_" + k] = getattr(rb, k)
    meta["reads_chrom"], meta["reads_id"], meta["reads_strand"] = rb.chrom, rb.read_id, rb.read_strand
    # zscore + no reference + positions filter variants
    feats_z, _ = ref_ef._extract_features(list(store)[:2], "x", "y", "zscore", ["CG"], 0, None, 17, 360, 0, None)
    out["extract_z_means"] = np.asarray([f[7] for f in feats_z], dtype=np.float64)
    out["extract_z_cent"] = np.asarray([np.asarray(f[10], dtype=np.float64) for f in feats_z])
    meta["extract_z_info"] = [[f[0], int(f[1]), f[2], int(f[3]), f[4], f[5], f[6]] for f in feats_z]

    # ---- call_mods host code: TSV parse + pred_str formatting with a fake session ----
    sf = synth_features(12, seed=7, sites_per_read=5)
    tsv = os.path.join(HERE, "features_small.tsv")
    write_feature_tsv(tsv, sf)

    class Q(list):
        def put(self, x): self.append(x)
        def qsize(self): return 0

    q = Q()
    ref_cm._read_features_file(tsv, q, f5_batch_num=2)
    assert q[-1] == "kill"
    meta["tsv_batches"] = [len(b[0]) for b in q[:-1]]
    b0 = q[0]
    meta["tsv_b0_sampleinfo"] = b0[0]
    out["tsv_b0_kmers"] = np.asarray(b0[1], dtype=np.int32)
    out["tsv_b0_means"] = np.asarray(b0[2], dtype=np.float64)
    out["tsv_b0_lens"] = np.asarray(b0[4], dtype=np.int64)
    out["tsv_b0_signals"] = np.asarray(b0[5], dtype=np.float64)

    prng = np.random.default_rng(5)
    fake_probs = prng.random((len(b0[0]), 2)).astype(np.float32)
    fake_probs[0] = [0.5, 0.5]
    fake_probs[1] = [1e-8, 0.99999994]

    class Sess:
        def run(self, fetches, feed_dict=None):
            n = len(feed_dict["base_int"])
            return fake_probs[:n], np.argmax(fake_probs[:n], axis=1)

    model = types.SimpleNamespace(base_int="base_int", means="means", stds="stds", sanums="sanums",
                                  signals="signals", labels="labels", lr="lr", training="training",
                                  keep_prob="keep_prob", activation_logits="al", prediction="pred")
    pred_str, acc, nb = ref_cm._call_mods(b0, Sess(), model, 0.001, 512)
    meta["call_pred_str"] = pred_str
    out["call_fake_probs"] = fake_probs

    # ---- post-processing: the reference's own frequency script on a synthetic call_mods file ----
    import subprocess
    rng2 = np.random.default_rng(4)
    lines = []
    for i in range(400):
        chrom_i = ["chrA", "chrB", "chr10", "chr2"][rng2.integers(0, 4)]
        pos_i = int(rng2.integers(100, 140))
        p1 = float(np.float32(rng2.random()))
        p0 = float(np.float32(1 - p1))
        lines.append("\t".join([chrom_i, str(pos_i), "+-"[pos_i % 2], str(pos_i + 7), "read%d" % i, "t", str(np.float32(p0)),
                                 str(np.float32(p1)), str(int(p1 > p0)), "ACGTACGTCGACGTACG"]))
    calls = os.path.join(HERE, "calls_small.tsv")
    with open(calls, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    freq = {}
    for tag, extra in (("plain", []), ("sort", ["--sort"]), ("bed", ["--sort", "--bed"]), ("cf", ["--prob_cf", "0.3"])):
        tmp_out = os.path.join(HERE, "_freq_tmp.txt")
        subprocess.check_call([sys.executable, os.path.join(REF, "scripts", "call_modification_frequency.py"), "-i", calls,
                               "-o", tmp_out] + extra, cwd=os.path.join(REF, "scripts"), stdout=subprocess.DEVNULL)
        freq[tag] = open(tmp_out).read()
        os.remove(tmp_out)
    with open(os.path.join(HERE, "freq_ref.json"), "w") as fh:
        json.dump(freq, fh)

    np.savez_compressed(os.path.join(HERE, "features_ref.npz"), **out)
    with open(os.path.join(HERE, "features_ref.json"), "w") as fh:
        json.dump(meta, fh, indent=0)
    print("wrote golden fixtures:", len(out), "arrays;", meta["extract_n"], "reference sites")


if __name__ == "__main__":
    main()

"""the feature extraction module, B200 engine behind the reference's interface.

output format (reference extract_features.py:1-5, README.md:141-153):
chrom, pos, alignstrand, pos_in_strand, readname, read_strand, k_mer, signal_means,
signal_stds, signal_lens, cent_signals, methy_label

Host-side mirror of ``deepsignal/extract_features.py``: ``_extract_features`` returns the same
12-tuples (:278-280) and ``_features_to_str`` renders the same text (:289-303), but the per-read
math (:143-190, :211-276) runs in libdsb200's CUDA kernels (``Engine.extract`` ->
``dsb_extract_host``).  Reads come as packed arrays (:class:`deepsignal_b200.synth.ReadBatch`,
stored as ``.npz``); a directory of tombo-resquiggled fast5 files is read with h5py when h5py is
importable (it is not in the build image, so that path is untested there).
"""
from __future__ import absolute_import

import os
import sys
import time

import numpy as np

from .synth import ReadBatch
from .utils.process_utils import get_motif_seqs, get_contig2len, str2bool  # noqa: F401

key_sep = "||"
_READ_FIELDS = ("raw", "raw_off", "scaling", "offset", "ev_start", "ev_len", "ev_base", "ev_off",
                "chrom_start", "strand_minus", "chrom_len")


def save_reads(path, rb):
    np.savez_compressed(path, chrom=np.asarray(rb.chrom), read_id=np.asarray(rb.read_id),
                        read_strand=np.asarray(rb.read_strand), **{k: getattr(rb, k) for k in _READ_FIELDS})


def _reads_from_npz(path):
    with np.load(path, allow_pickle=False) as z:
        return ReadBatch(*[z[k] for k in _READ_FIELDS], [str(x) for x in z["chrom"]],
                         [str(x) for x in z["read_id"]], [str(x) for x in z["read_strand"]])


def _concat_reads(parts):
    if len(parts) == 1:
        return parts[0]
    raw_off, ev_off = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)]
    r0 = e0 = 0
    for p in parts:
        raw_off.append(p.raw_off[1:] + r0)
        ev_off.append(p.ev_off[1:] + e0)
        r0 += int(p.raw_off[-1])
        e0 += int(p.ev_off[-1])
    cat = lambda k: np.concatenate([getattr(p, k) for p in parts])
    return ReadBatch(cat("raw"), np.concatenate(raw_off), cat("scaling"), cat("offset"), cat("ev_start"),
                     cat("ev_len"), cat("ev_base"), np.concatenate(ev_off), cat("chrom_start"),
                     cat("strand_minus"), cat("chrom_len"), sum((p.chrom for p in parts), []),
                     sum((p.read_id for p in parts), []), sum((p.read_strand for p in parts), []))


def _reads_from_fast5s(fast5_files, corrected_group, basecall_subgroup):
    """Pull the arrays the reference reads from each fast5 (:35-72,121-140,193-208).  Needs h5py."""
    import h5py  # noqa: F401  (absent in the build image)
    raws, ev_s, ev_l, ev_b, rows, errors = [], [], [], [], [], 0
    for fn in fast5_files:
        try:
            with h5py.File(fn, "r") as h5:
                rd = list(h5["Raw/Reads"].values())[0]
                raw = np.asarray(rd["Signal"][()], dtype=np.int16)
                read_id = rd.attrs["read_id"]
                read_id = read_id.decode() if isinstance(read_id, bytes) else str(read_id)
                grp = h5["/Analyses/" + corrected_group + "/" + basecall_subgroup]
                ev = grp["Events"]
                rel = int(ev.attrs["read_start_rel_to_raw"])
                al = grp["Alignment"].attrs
                dec = lambda v: v.decode() if isinstance(v, bytes) else str(v)
                ch = h5["UniqueGlobalKey/channel_id"].attrs
                row = (dec(al["mapped_chrom"]), int(al["mapped_start"]), dec(al["mapped_strand"]), read_id,
                       "t" if basecall_subgroup.endswith("template") else "c",
                       float(ch["range"]) / float(ch["digitisation"]), float(ch["offset"]))
                starts = np.asarray(ev["start"], dtype=np.int64) + rel
                lengths = np.asarray(ev["length"], dtype=np.int64)
                bases = np.frombuffer(b"".join(bytes(x) for x in ev["base"]), dtype=np.uint8)
                if not (len(starts) == len(lengths) == len(bases)):
                    raise ValueError("Events columns differ in length")
        except Exception:
            errors += 1
        else:       # every field of the read parsed: only now do the per-read lists grow (they must stay aligned)
            rows.append(row)
            raws.append(raw)
            ev_s.append(starts)
            ev_l.append(lengths)
            ev_b.append(bases)
    n = len(raws)
    off = lambda parts: np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64)
    cat = lambda parts, dt: (np.concatenate(parts).astype(dt) if n else np.zeros(0, dt))
    rb = ReadBatch(cat(raws, np.int16), off(raws), np.asarray([r[5] for r in rows], np.float64),
                   np.asarray([r[6] for r in rows], np.float64), cat(ev_s, np.int32), cat(ev_l, np.int32),
                   cat(ev_b, np.uint8), off(ev_s), np.asarray([r[1] for r in rows], np.int64),
                   np.asarray([1 if r[2] == "-" else 0 for r in rows], np.uint8), np.full(n, -1, np.int64),
                   [r[0] for r in rows], [r[3] for r in rows], [r[4] for r in rows])
    return rb, errors


def load_reads(path, chrom2len=None, is_recursive=True, corrected_group="RawGenomeCorrected_000",
               basecall_subgroup="BaseCalled_template", return_errors=False):
    """``path``: a packed-reads ``.npz``, a directory of them, or a directory of fast5 files."""
    errors = 0
    if os.path.isfile(path):
        rb = _reads_from_npz(path)
    else:
        files = []
        for root, _, fns in os.walk(path):
            files += [os.path.join(root, f) for f in sorted(fns)]
            if not is_recursive:
                break
        npz = [f for f in files if f.endswith(".npz")]
        f5 = [f for f in files if f.endswith(".fast5")]
        if npz:
            rb = _concat_reads([_reads_from_npz(f) for f in npz])
        elif f5:
            rb, errors = _reads_from_fast5s(f5, corrected_group, basecall_subgroup)
        else:
            raise FileNotFoundError("no .npz packed reads or .fast5 files under {}".format(path))
    # chromosome lengths come from --reference_path (reference :240-248): None -> pos_in_strand = -1
    if chrom2len is None:
        rb.chrom_len = np.full(rb.n_reads, -1, np.int64)
    else:
        missing = sorted(set(c for c in rb.chrom if c not in chrom2len))
        if missing:
            print("warning - chrom_name in fast5 not in provided reference genome!")
        rb.chrom_len = np.asarray([chrom2len.get(c, -1) for c in rb.chrom], np.int64)
    return (rb, errors) if return_errors else rb


def _read_position_file(position_file):
    postions = set()
    with open(position_file, "r") as rf:
        for line in rf:
            words = line.strip().split("\t")
            postions.add(key_sep.join(words[:3]))
    return postions


def failed_reads(arr):
    """Reads the extraction could not normalise: no raw samples / non-positive scaling (the kernels emit no sites for
    them) or a zero / non-finite scale (MAD = 0 on a constant signal: numpy would print inf / nan features).  They
    count as errors and contribute no features, like a read the reference's try/except drops (:281-283)."""
    sc = arr["read_scale"]
    return ~np.isfinite(sc) | (sc == 0)


def extract_batch_arrays(engine, reads, motif_seqs, methyloc, kmer_len, raw_signals_len, normalize_method,
                         positions=None, want64=True):
    """GPU extraction of one read batch -> (arrays dict, sampleinfo strings); applies --positions.
    ``want64=False`` skips the float64 copies that only the feature-file text needs (call_mods feeds float32)."""
    if kmer_len % 2 == 0:
        raise ValueError("kmer_len must be odd")
    arr = engine.extract(reads, motif_seqs, mod_loc=methyloc, kmer_len=kmer_len, signal_len=raw_signals_len,
                         normalize_method=normalize_method, want64=want64)
    n = arr["n_sites"]
    strand = np.where(reads.strand_minus[arr["site_read"]] != 0, "-", "+") if n else np.zeros(0, "<U1")
    bad_read = failed_reads(arr)
    info, keep = [], []
    for i in range(n):
        r = int(arr["site_read"][i])
        if bad_read[r]:
            continue
        chrom, pos = reads.chrom[r], int(arr["pos"][i])
        if positions is not None and key_sep.join([chrom, str(pos), str(strand[i])]) not in positions:
            continue
        keep.append(i)
        info.append("\t".join([chrom, str(pos), str(strand[i]), str(int(arr["pos_in_strand"][i])),
                               reads.read_id[r], reads.read_strand[r]]))
    if len(keep) != n:
        keep = np.asarray(keep, dtype=np.int64)
        arr = {k: (v[keep] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v) for k, v in arr.items()}
        arr["n_sites"] = len(keep)
    return arr, info


def _extract_features(reads, corrected_group, basecall_subgroup, normalize_method,
                      motif_seqs, methyloc, chrom2len, kmer_len, raw_signals_len,
                      methy_label, positions, engine=None):
    """Mirror of the reference ``_extract_features`` (:215-286): returns (features_list, error).

    ``reads`` is a :class:`ReadBatch` (the reference takes fast5 paths); reads whose normalisation
    is undefined (empty signal, non-positive scaling) count as errors like a raised exception would.
    """
    from .engine import Engine
    own = engine is None
    if own:
        engine = Engine(0)
    try:
        if chrom2len is not None:
            reads.chrom_len = np.asarray([chrom2len.get(c, -1) for c in reads.chrom], np.int64)
        arr, info = extract_batch_arrays(engine, reads, motif_seqs, methyloc, kmer_len, raw_signals_len,
                                         normalize_method, positions)
    finally:
        if own:
            engine.close()
    error = int(np.sum(failed_reads(arr)))
    code2base = "ACGTN"
    features_list = []
    for i in range(arr["n_sites"]):
        chrom, pos, alignstrand, pos_in_strand, readname, strand = info[i].split("\t")
        k_mer = "".join(code2base[c] if c < 5 else "?" for c in arr["kmer"][i])
        features_list.append((chrom, int(pos), alignstrand, int(pos_in_strand), readname, strand, k_mer,
                              list(arr["means64"][i]), list(arr["stds64"][i]), [int(x) for x in arr["lens_i32"][i]],
                              arr["signals64"][i], methy_label))
    return features_list, error


def _features_to_str(features):
    """Same text as reference ``_features_to_str`` (:289-303)."""
    chrom, pos, alignstrand, pos_in_strand, readname, strand, k_mer, signal_means, signal_stds, \
        signal_lens, cent_signals, methy_label = features
    means_text = ",".join([str(x) for x in np.around(signal_means, decimals=6)])
    stds_text = ",".join([str(x) for x in np.around(signal_stds, decimals=6)])
    signal_len_text = ",".join([str(x) for x in signal_lens])
    cent_signals_text = ",".join([str(x) for x in cent_signals])
    return "\t".join([chrom, str(pos), alignstrand, str(pos_in_strand), readname, strand, k_mer, means_text,
                      stds_text, signal_len_text, cent_signals_text, str(methy_label)])


def extract_features(fast5_dir, is_recursive, reference_path, is_dna,
                     batch_size, write_fp, nproc,
                     corrected_group, basecall_subgroup, normalize_method,
                     motifs, methyloc, kmer_len, raw_signals_len, methy_label,
                     position_file, w_is_dir, w_batch_num, device=0):
    """Same signature as the reference ``extract_features`` (:424-478); ``batch_size`` is the number
    of reads per GPU batch (reference: per worker process), ``nproc`` is accepted and unused."""
    from .engine import Engine
    start = time.time()
    print("parse the motifs string..")
    motif_seqs = get_motif_seqs(motifs, is_dna)
    print("read genome reference file..")
    chrom2len = get_contig2len(reference_path) if reference_path is not None else None
    positions = _read_position_file(position_file) if position_file is not None else None
    reads, errornum = load_reads(fast5_dir, chrom2len, is_recursive, corrected_group, basecall_subgroup,
                                 return_errors=True)
    print("{} fast5 files in total..".format(reads.n_reads + errornum))
    engine = Engine(device)
    step = max(1, int(batch_size)) * 64
    file_count, batch_count = 0, 0
    if w_is_dir:
        os.makedirs(write_fp, exist_ok=True)
        wf = open(os.path.join(write_fp, "0.tsv"), "w")
    else:
        wf = open(write_fp, "w")
    try:
        for lo in range(0, reads.n_reads, step):
            rb = reads.slice(lo, min(reads.n_reads, lo + step))
            feats, err = _extract_features(rb, corrected_group, basecall_subgroup, normalize_method, motif_seqs,
                                           methyloc, None, kmer_len, raw_signals_len, methy_label, positions,
                                           engine=engine)
            errornum += err
            if w_is_dir and batch_count >= w_batch_num:
                wf.close()
                file_count += 1
                wf = open(os.path.join(write_fp, str(file_count) + ".tsv"), "w")
                batch_count = 0
            for f in feats:
                wf.write(_features_to_str(f) + "\n")
            batch_count += 1
    finally:
        wf.close()
        engine.close()
    print("%d of %d fast5 files failed..\nextract_features costs %.1f seconds.." %
          (errornum, reads.n_reads + (errornum if errornum else 0), time.time() - start))


def main():
    from .deepsignal import main as cli_main
    sys.argv.insert(1, "extract")
    return cli_main()


if __name__ == "__main__":
    sys.exit(main())

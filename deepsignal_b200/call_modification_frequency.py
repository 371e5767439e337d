#! /usr/bin/env python
"""calculate modification frequency at genome level from call_mods result files.

Same inputs, options and output text as the reference's ``scripts/call_modification_frequency.py``
(:17-79) with ``scripts/txt_formater.py`` (ModRecord / SiteStats / split_key): per (chromosome, position) key,
sum of prob_0 / prob_1, counts of modified / unmodified calls, coverage, Rmet and the k-mer of the first record;
``--bed`` writes bedMethyl.  The reduction runs either on the host (NumPy: ``np.unique`` + ``np.add.at``, the default, no GPU
needed) or, with ``--device N``, on the GPU as a segmented reduce keyed on (chromosome, position) through
``dsb_site_frequency_host`` (csrc/sitefreq.cu: stable radix sort + one thread per site adding its records in line order);
both give the reference's text byte for byte."""
from __future__ import absolute_import

import argparse
import gzip
import os
import sys

import numpy as np

key_sep = "||"


def _read_records(mods_files):
    rows = []
    for fn in mods_files:
        opener = gzip.open(fn, "rt") if fn.endswith(".gz") else open(fn, "r")
        with opener as fh:
            for line in fh:
                w = line.strip().split("\t")
                if len(w) >= 10:
                    rows.append(w[:10])
    return rows


def calculate_mods_frequency(mods_files, prob_cf, device=None):
    """Returns (keys, stats) with keys in first-appearance order like the reference's dict.
    ``device``: CUDA device index -> aggregate on the GPU (libdsb200); None -> NumPy on the host."""
    rows = _read_records(mods_files)
    count = len(rows)
    if count == 0:
        print("0 calls..")
        return [], {}
    chrom = np.asarray([r[0] for r in rows])
    pos = np.asarray([int(r[1]) for r in rows], np.int64)
    p0 = np.asarray([float(r[6]) for r in rows], np.float64)
    p1 = np.asarray([float(r[7]) for r in rows], np.float64)
    lab = np.asarray([int(r[8]) for r in rows], np.int64)
    use = np.abs(p0 - p1) >= prob_cf                     # ModRecord.is_record_callable
    if device is not None:
        from .engine import Engine
        names, chrom_id = np.unique(chrom, return_inverse=True)
        eng = Engine(int(device))
        try:
            r = eng.site_frequency(chrom_id.astype(np.int32), pos, p0, p1, (lab == 1).astype(np.uint8), use.astype(np.uint8))
        finally:
            eng.close()
        firsts = r["first"]
        stats = dict(chrom=chrom[firsts], pos=pos[firsts], strand=np.asarray([rows[i][2] for i in firsts]),
                     pos_in_strand=np.asarray([int(rows[i][3]) for i in firsts], np.int64), prob_0=r["prob_0"], prob_1=r["prob_1"],
                     met=r["met"], unmet=r["coverage"] - r["met"], coverage=r["coverage"], kmer=np.asarray([rows[i][9] for i in firsts]))
        used = int(use.sum())
        print("{:.2f}% ({} of {}) calls used..".format(used / float(count) * 100, used, count))
        return [key_sep.join([str(chrom[i]), str(int(pos[i]))]) for i in firsts], stats
    idx = np.flatnonzero(use)
    keys_all = np.char.add(np.char.add(chrom, key_sep), pos.astype(str))
    uk, first, inv = np.unique(keys_all[idx], return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")             # first-appearance order
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    g = rank[inv]
    n = len(uk)
    s0, s1 = np.zeros(n), np.zeros(n)
    # sequential accumulation order per key == the reference's "+=" over lines
    np.add.at(s0, g, p0[idx])
    np.add.at(s1, g, p1[idx])
    met = np.bincount(g, weights=(lab[idx] == 1), minlength=n).astype(np.int64)
    cov = np.bincount(g, minlength=n).astype(np.int64)
    firsts = idx[first[order]]
    stats = dict(chrom=chrom[firsts], pos=pos[firsts], strand=np.asarray([rows[i][2] for i in firsts]),
                 pos_in_strand=np.asarray([int(rows[i][3]) for i in firsts], np.int64), prob_0=s0, prob_1=s1,
                 met=met, unmet=cov - met, coverage=cov, kmer=np.asarray([rows[i][9] for i in firsts]))
    print("{:.2f}% ({} of {}) calls used..".format(len(idx) / float(count) * 100, len(idx), count))
    return list(uk[order]), stats


def write_sitekey2stats(stats, result_file, is_sort, is_bed):
    n = len(stats.get("pos", []))
    order = np.arange(n)
    if is_sort and n:
        order = np.lexsort((stats["pos"], stats["chrom"]))        # split_key -> (chrom, int(pos))
    with open(result_file, "w") as wf:
        for i in order:
            cov = int(stats["coverage"][i])
            if cov <= 0:
                continue
            chrom, pos = str(stats["chrom"][i]), int(stats["pos"][i])
            rmet = float(stats["met"][i]) / cov
            if is_bed:
                wf.write("\t".join([chrom, str(pos), str(pos + 1), ".", str(cov), str(stats["strand"][i]), str(pos),
                                    str(pos + 1), "0,0,0", str(cov), str(int(round(rmet * 100, 0)))]) + "\n")
            else:
                wf.write("%s\t%d\t%s\t%d\t%.3f\t%.3f\t%d\t%d\t%d\t%.4f\t%s\n" % (
                    chrom, pos, stats["strand"][i], stats["pos_in_strand"][i], stats["prob_0"][i], stats["prob_1"][i],
                    stats["met"][i], stats["unmet"][i], cov, rmet, stats["kmer"][i]))


def main():
    parser = argparse.ArgumentParser(description="calculate frequency of interested sites at genome level")
    parser.add_argument("--input_path", "-i", action="append", type=str, required=True)
    parser.add_argument("--result_file", "-o", action="store", type=str, required=True)
    parser.add_argument("--bed", action="store_true", default=False)
    parser.add_argument("--sort", action="store_true", default=False)
    parser.add_argument("--prob_cf", type=float, action="store", required=False, default=0.0)
    parser.add_argument("--file_uid", type=str, action="store", required=False, default=None)
    parser.add_argument("--device", type=int, default=None, help="CUDA device index: aggregate on the GPU (default: host NumPy)")
    args = parser.parse_args()
    mods_files = []
    for ipath in args.input_path:
        input_path = os.path.abspath(ipath)
        if os.path.isdir(input_path):
            for ifile in os.listdir(input_path):
                if args.file_uid is None or ifile.find(args.file_uid) != -1:
                    mods_files.append("/".join([input_path, ifile]))
        elif os.path.isfile(input_path):
            mods_files.append(input_path)
        else:
            raise ValueError()
    print("get {} input file(s)..".format(len(mods_files)))
    _, stats = calculate_mods_frequency(mods_files, args.prob_cf, device=args.device)
    write_sitekey2stats(stats, args.result_file, args.sort, args.bed)


if __name__ == "__main__":
    sys.exit(main())

"""ORACLE (test infrastructure, never shipped or measured as the product).

NumPy float64 restatement of the reference's per-read feature math
(deepsignal/extract_features.py:143-190, 211-303 and
deepsignal/utils/process_utils.py:95-110) over in-memory read tables — the fast5
(HDF5) container itself cannot be opened here (no h5py), so a "read" is the set of
arrays the reference pulls out of the file (extract_features.py:35-72,121-140,
193-208):

    raw            int16[N]      Raw/Reads/*/Signal
    digitisation, range, offset  UniqueGlobalKey/channel_id attrs (float64)
    read_start_rel_to_raw        Events attr
    ev_start, ev_len, ev_base    Events table (start is relative to read_start_rel_to_raw)
    read_id, strand('t'/'c'), alignstrand('+'/'-'), chrom, chrom_start   Alignment attrs

Third-party arithmetic restated (absent from /root/reference):
* statsmodels>=0.9 ``robust.mad(a)`` = ``median(|a - median(a)| / c)`` with
  c = Phi^-1(3/4) = 0.6744897501960817 (requirements.txt:3; call site
  extract_features.py:147).  Taken from the published statsmodels source on recall —
  statsmodels is not installed here; dividing before or after the outer median
  differs by at most 1 ulp.
* numpy ``median / mean / std / around`` are called directly (numpy is present).

PARITY PIN STATUS: pinned against the reference's own functions executed in this
container (tests/golden/make_golden.py imports /root/reference with h5py /
statsmodels stubs and np.int/np.float shims) -> tests/golden/features_ref.npz.

One deliberate, documented deviation: when the centre base alone has >= 360 samples
the reference draws ``sorted(random.sample(range(n), 360))`` from an unseeded RNG
(extract_features.py:165-168), which is irreproducible.  The oracle and the CUDA path
both take the deterministic sorted subset ``idx_j = (j*n)//360``; tests check the
invariant the reference guarantees (a strictly increasing 360-subset of the centre
base's samples).
"""
from __future__ import annotations

import numpy as np

MAD_C = 0.6744897501960817   # scipy.stats.norm.ppf(0.75)
KEY_SEP = "||"               # reference extract_features.py:32


def rescale_signals(raw, scaling, offset):
    """reference extract_features.py:211-212 (offset is the float64 channel attr)."""
    return np.array(np.float64(scaling) * (np.asarray(raw) + np.float64(offset)), dtype=np.float64)


def mad(a):
    a = np.asarray(a, dtype=np.float64)
    return np.median(np.fabs(a - np.median(a)) / MAD_C)


def normalize_signals(signals, normalize_method="mad"):
    """reference extract_features.py:143-151."""
    if normalize_method == "zscore":
        sshift, sscale = np.mean(signals), float(np.std(signals))
    elif normalize_method == "mad":
        sshift, sscale = np.median(signals), float(mad(signals))
    else:
        raise ValueError("")
    return np.around((signals - sshift) / sscale, decimals=6)


def subsample_indices(n: int, m: int):
    """Deterministic stand-in for sorted(random.sample(range(n), m)), n >= m."""
    return [(j * n) // m for j in range(m)]


def central_window(lens, rawsignal_num=360):
    """Window of the concatenated k-mer signal, reference extract_features.py:154-190.

    Returns (mode, start, n_real): mode 0 = all T<360 samples then zero fill,
    mode 1 = centre-base subsample, mode 2 = contiguous [start, start+360) of the
    concatenation.  ``start`` indexes the concatenation of the k-mer's slices.
    """
    lens = [int(x) for x in lens]
    total = sum(lens)
    if total < rawsignal_num:
        return 0, 0, total
    mid = int((len(lens) - 1) / 2)
    if lens[mid] >= rawsignal_num:
        return 1, sum(lens[:mid]), rawsignal_num
    left_len = (rawsignal_num - lens[mid]) // 2
    right_len = rawsignal_num - left_len
    L, R = sum(lens[:mid]), sum(lens[mid:])
    if left_len > L:
        right_len = right_len + left_len - L
        left_len = L
    elif right_len > R:
        left_len = left_len + right_len - R
        right_len = R
    assert left_len + right_len == rawsignal_num
    return 2, L - left_len, rawsignal_num


def get_central_signals(signals_list, rawsignal_num=360):
    """reference extract_features.py:154-190 with the deterministic branch-2 rule."""
    lens = [len(x) for x in signals_list]
    mode, start, n_real = central_window(lens, rawsignal_num)
    cat = np.concatenate([np.asarray(x, dtype=np.float64) for x in signals_list]) if sum(lens) else np.zeros(0)
    if mode == 0:
        return np.append(cat, np.zeros(rawsignal_num - n_real))
    if mode == 1:
        mid = int((len(lens) - 1) / 2)
        centre = np.asarray(signals_list[mid], dtype=np.float64)
        return centre[subsample_indices(len(centre), rawsignal_num)]
    return cat[start:start + rawsignal_num]


def motif_sites(seq: str, motifset, methyloc: int):
    """reference process_utils.py:95-110."""
    motifset = set(motifset)
    m = len(next(iter(motifset)))
    return [i + methyloc for i in range(len(seq) - m + 1) if seq[i:i + m] in motifset]


def extract_read(read: dict, motif_seqs, methyloc=0, chrom2len=None, kmer_len=17,
                 raw_signals_len=360, methy_label=1, positions=None, normalize_method="mad"):
    """One read of reference ``_extract_features`` (extract_features.py:228-280).

    Returns the list of 12-tuples the reference appends (same field order, :278-280).
    """
    if kmer_len % 2 == 0:
        raise ValueError("kmer_len must be odd")
    nbh = (kmer_len - 1) // 2
    scaling = np.float64(read["range"]) / np.float64(read["digitisation"])
    sig = rescale_signals(read["raw"], scaling, read["offset"])
    norm = normalize_signals(sig, normalize_method)
    starts = np.asarray(read["ev_start"], dtype=np.int64) + int(read["read_start_rel_to_raw"])
    lens = np.asarray(read["ev_len"], dtype=np.int64)
    bases = read["ev_base"]
    seq = bases.decode("ascii") if isinstance(bases, (bytes, bytearray)) else "".join(
        chr(c) for c in np.asarray(bases, dtype=np.uint8))
    signal_list = [norm[s:s + l] for s, l in zip(starts.tolist(), lens.tolist())]
    chrom, alignstrand = read["chrom"], read["alignstrand"]
    chromlen = None
    if chrom2len is not None:
        chromlen = chrom2len.get(chrom)   # KeyError -> warning + None in the reference (:244-248)
    out = []
    nb = len(seq)
    for loc in motif_sites(seq, motif_seqs, methyloc):
        if not (nbh <= loc < nb - nbh):
            continue
        if alignstrand == "-":
            pos = int(read["chrom_start"]) + nb - 1 - loc
            pos_in_strand = chromlen - 1 - pos if chromlen is not None else -1
        else:
            pos = int(read["chrom_start"]) + loc
            pos_in_strand = pos if chromlen is not None else -1
        if positions is not None and KEY_SEP.join([chrom, str(pos), alignstrand]) not in positions:
            continue
        k_mer = seq[loc - nbh:loc + nbh + 1]
        k_signals = signal_list[loc - nbh:loc + nbh + 1]
        signal_lens = [len(x) for x in k_signals]
        signal_means = [np.mean(x) for x in k_signals]
        signal_stds = [np.std(x) for x in k_signals]
        cent = get_central_signals(k_signals, raw_signals_len)
        out.append((chrom, pos, alignstrand, pos_in_strand, read["read_id"], read["strand"],
                    k_mer, signal_means, signal_stds, signal_lens, cent, methy_label))
    return out


def features_to_str(features) -> str:
    """reference extract_features.py:289-303 (12 tab-separated columns)."""
    chrom, pos, alignstrand, pos_in_strand, readname, strand, k_mer, means, stds, lens, cent, label = features
    means_text = ",".join(str(x) for x in np.around(means, decimals=6))
    stds_text = ",".join(str(x) for x in np.around(stds, decimals=6))
    lens_text = ",".join(str(x) for x in lens)
    cent_text = ",".join(str(x) for x in cent)
    return "\t".join([chrom, str(pos), alignstrand, str(pos_in_strand), readname, strand, k_mer,
                      means_text, stds_text, lens_text, cent_text, str(label)])

"""Seeded synthetic inputs for the configurations of BASELINE.json.

* :func:`synth_features` — bn17/sn360 samples as the five model-input arrays
  (cfg 1, 3, 5); :func:`write_feature_tsv` renders them in the reference's
  12-column feature-file layout (extract_features.py:289-303, README.md:141-153).
* :func:`synth_reads` — tombo-style resquiggled R9.4 reads as packed arrays
  (cfg 2, 4): what the reference reads out of a fast5 (extract_features.py:35-72,
  193-208), held in memory because HDF5 cannot be read here.
All distributions are our synthetic assumptions (stated in SURVEY §8d), not data.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SEED_CFG1 = 20260101
SEED_CFG2 = 20260102
SEED_CFG3 = 20260103
SEED_CFG4 = 20260104

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _dwell(rng, n):
    """3 + Geometric(p=1/5.9) samples per base (mean ~8.9), 1 % stalls x20."""
    d = 3 + rng.geometric(1.0 / 5.9, size=n)
    stall = rng.random(n) < 0.01
    d[stall] *= 20
    return d.astype(np.int64)


def synth_features(n: int, seed: int = SEED_CFG1, kmer_len: int = 17, signal_num: int = 360,
                   sites_per_read: int = 50):
    """Model-input arrays for ``n`` CpG samples.

    Returns dict(kmer int32[n,k], means f32, stds f32, lens int32[n,k], signals f32[n,s],
    read_index int32[n]).  Values are rounded to 6 dp like a feature file's text.
    """
    rng = np.random.default_rng(seed)
    half = (kmer_len - 1) // 2
    kmer = rng.integers(0, 4, size=(n, kmer_len), dtype=np.int32)
    kmer[:, half] = 1       # C
    kmer[:, half + 1] = 2   # G
    lens = _dwell(rng, n * kmer_len).reshape(n, kmer_len).astype(np.int32)
    means = np.around(rng.normal(0.0, 1.0, size=(n, kmer_len)), 6).astype(np.float32)
    stds = np.around(np.abs(rng.normal(0.15, 0.05, size=(n, kmer_len))) + 0.01, 6).astype(np.float32)
    signals = np.around(rng.normal(0.0, 1.0, size=(n, signal_num)), 6).astype(np.float32)
    total = lens.sum(axis=1)
    nreal = np.minimum(total, signal_num)
    signals[np.arange(signal_num)[None, :] >= nreal[:, None]] = 0.0
    read_index = (np.arange(n) // sites_per_read).astype(np.int32)
    return dict(kmer=kmer, means=means, stds=stds, lens=lens, signals=signals, read_index=read_index)


def write_feature_tsv(path: str, feats: dict, label: int = 1, chrom: str = "chrSyn") -> None:
    """Render :func:`synth_features` output as a reference-layout feature file."""
    code2base = "ACGTN"
    with open(path, "w") as fh:
        n = feats["kmer"].shape[0]
        for i in range(n):
            kmer = "".join(code2base[c] for c in feats["kmer"][i])
            cols = [chrom, str(1000 + i), "+", str(1000 + i), "read_%06d" % feats["read_index"][i], "t", kmer,
                    ",".join(str(np.float64(x)) for x in np.around(feats["means"][i].astype(np.float64), 6)),
                    ",".join(str(np.float64(x)) for x in np.around(feats["stds"][i].astype(np.float64), 6)),
                    ",".join(str(int(x)) for x in feats["lens"][i]),
                    ",".join(str(np.float64(x)) for x in np.around(feats["signals"][i].astype(np.float64), 6)),
                    str(label)]
            fh.write("\t".join(cols) + "\n")


@dataclass
class ReadBatch:
    """Packed resquiggled reads (host arrays).  Event starts are absolute raw indices."""
    raw: np.ndarray            # int16[total_raw]
    raw_off: np.ndarray        # int64[R+1]
    scaling: np.ndarray        # float64[R]  range / digitisation
    offset: np.ndarray         # float64[R]
    ev_start: np.ndarray       # int32[total_ev]  start + read_start_rel_to_raw
    ev_len: np.ndarray         # int32[total_ev]
    ev_base: np.ndarray        # uint8[total_ev]  ASCII
    ev_off: np.ndarray         # int64[R+1]
    chrom_start: np.ndarray    # int64[R]
    strand_minus: np.ndarray   # uint8[R]  1 = mapped to '-'
    chrom_len: np.ndarray      # int64[R]  -1 when no reference was given
    chrom: list = field(default_factory=list)     # str per read
    read_id: list = field(default_factory=list)   # str per read
    read_strand: list = field(default_factory=list)  # 't' / 'c'

    @property
    def n_reads(self) -> int:
        return len(self.raw_off) - 1

    def read_dict(self, i: int) -> dict:
        """Read ``i`` in the oracle's dict form (event starts back to relative, rel=0)."""
        r0, r1 = int(self.raw_off[i]), int(self.raw_off[i + 1])
        e0, e1 = int(self.ev_off[i]), int(self.ev_off[i + 1])
        return dict(raw=self.raw[r0:r1], digitisation=1.0, range=float(self.scaling[i]),
                    offset=float(self.offset[i]), read_start_rel_to_raw=0,
                    ev_start=self.ev_start[e0:e1], ev_len=self.ev_len[e0:e1],
                    ev_base=self.ev_base[e0:e1].tobytes(), read_id=self.read_id[i],
                    strand=self.read_strand[i], alignstrand="-" if self.strand_minus[i] else "+",
                    chrom=self.chrom[i], chrom_start=int(self.chrom_start[i]))

    def slice(self, lo: int, hi: int) -> "ReadBatch":
        r0, r1 = int(self.raw_off[lo]), int(self.raw_off[hi])
        e0, e1 = int(self.ev_off[lo]), int(self.ev_off[hi])
        return ReadBatch(self.raw[r0:r1], self.raw_off[lo:hi + 1] - r0, self.scaling[lo:hi], self.offset[lo:hi],
                         self.ev_start[e0:e1], self.ev_len[e0:e1], self.ev_base[e0:e1],
                         self.ev_off[lo:hi + 1] - e0, self.chrom_start[lo:hi], self.strand_minus[lo:hi],
                         self.chrom_len[lo:hi], self.chrom[lo:hi], self.read_id[lo:hi], self.read_strand[lo:hi])


CONTIGS = {"chrSynA": 5_000_000, "chrSynB": 3_000_000, "chrSynC": 2_000_000, "chrSynD": 1_000_000}


def synth_reads(n_reads: int, seed: int = SEED_CFG2, nb_lo: int = 1500, nb_hi: int = 4500,
                cg_frac: float = 0.01, long_stall_frac: float = 0.002, with_reference: bool = True,
                force_lens: dict | None = None) -> ReadBatch:
    """cfg-2 / cfg-4 generator: CpG-depleted sequence (~``cg_frac`` CG dinucleotides,
    ~30 sites per 3 kb read), 4-level current table + N(0, 12 ADC) noise around 480,
    contiguous events after a random ``read_start_rel_to_raw`` lead-in.  ``long_stall_frac``
    of the bases dwell 360..700 samples so every branch of the 360-sample window rule
    (extract_features.py:154-190) is exercised.  ``force_lens`` = {global event index: samples}."""
    rng = np.random.default_rng(seed)
    nb = rng.integers(nb_lo, nb_hi + 1, size=n_reads)
    ev_off = np.zeros(n_reads + 1, np.int64)
    ev_off[1:] = np.cumsum(nb)
    total_ev = int(ev_off[-1])
    # sequence: iid ACGT, break accidental CG, then plant CG at ~cg_frac of positions
    codes = rng.integers(0, 4, size=total_ev)
    is_cg = (codes[:-1] == 1) & (codes[1:] == 2)
    codes[1:][is_cg] = 3                      # C G -> C T
    plant = np.flatnonzero(rng.random(total_ev - 1) < cg_frac)
    codes[plant] = 1
    codes[plant + 1] = 2
    ev_base = _BASES[codes]
    ev_len = _dwell(rng, total_ev)
    long_stall = rng.random(total_ev) < long_stall_frac
    ev_len[long_stall] = rng.integers(360, 701, size=int(long_stall.sum()))
    for ev_index, n_samples in (force_lens or {}).items():      # tests: exact dwell of chosen events (draws nothing from rng)
        ev_len[ev_index] = n_samples
    lead = rng.integers(0, 2001, size=n_reads)
    tail = rng.integers(0, 200, size=n_reads)
    csum = np.zeros(total_ev + 1, np.int64)
    csum[1:] = np.cumsum(ev_len)
    per_read_sig = csum[ev_off[1:]] - csum[ev_off[:-1]]
    n_raw = lead + per_read_sig + tail
    raw_off = np.zeros(n_reads + 1, np.int64)
    raw_off[1:] = np.cumsum(n_raw)
    read_of_ev = np.repeat(np.arange(n_reads), nb)
    ev_start = (csum[:-1] - csum[ev_off[:-1]][read_of_ev] + lead[read_of_ev]).astype(np.int32)
    # raw signal: level by base inside events, neutral level in lead/tail
    total_raw = int(raw_off[-1])
    level_tab = np.array([430.0, 470.0, 500.0, 540.0])
    level = np.full(total_raw, 480.0)
    ev_abs = raw_off[:-1][read_of_ev] + ev_start
    seg_level = level_tab[codes]
    idx = np.repeat(ev_abs, ev_len) + (np.arange(int(csum[-1])) - np.repeat(csum[:-1], ev_len))
    level[idx] = np.repeat(seg_level, ev_len)
    raw = np.rint(level + rng.normal(0.0, 12.0, size=total_raw)).astype(np.int16)
    names = list(CONTIGS)
    chrom_idx = rng.integers(0, len(names), size=n_reads)
    chrom = [names[c] for c in chrom_idx]
    clen = np.array([CONTIGS[c] for c in chrom], np.int64)
    chrom_start = (rng.random(n_reads) * (clen - nb - 1)).astype(np.int64)
    return ReadBatch(
        raw=raw, raw_off=raw_off,
        scaling=np.full(n_reads, 1467.61 / 8192.0, np.float64),
        offset=rng.integers(-20, 31, size=n_reads).astype(np.float64),
        ev_start=ev_start, ev_len=ev_len.astype(np.int32), ev_base=ev_base.copy(), ev_off=ev_off,
        chrom_start=chrom_start, strand_minus=(rng.random(n_reads) < 0.5).astype(np.uint8),
        chrom_len=clen if with_reference else np.full(n_reads, -1, np.int64),
        chrom=chrom, read_id=["read-%08d" % i for i in range(n_reads)],
        read_strand=["t"] * n_reads)

"""Fast feature-file reader / result writer on top of the library's text codec
(``dsb_tsv_parse_features`` / ``dsb_tsv_format_calls``; csrc/tsv.cpp).  Same parsing rules as the
reference reader (call_modifications.py:47-85) and byte-identical result lines (:183-190)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib


class FeatureBlock(object):
    """``n`` parsed samples plus the raw text they came from (for the six info columns)."""
    __slots__ = ("n", "kmer", "means", "stds", "lens", "signals", "labels", "info_off", "text", "text_ptr")

    def info_strings(self):
        t = self.text
        o = self.text_ptr - C.cast(C.c_char_p(t), C.c_void_p).value
        return [t[o + s:o + e].decode() for s, e in self.info_off]


BLOCK_SAMPLES = 1 << 18     # samples per parsed block = per forward call of call_mods


def iter_feature_blocks(path, kmer_len=17, signal_len=360, max_samples=None, nthreads=None, block_bytes=None):
    lib = _lib.load()
    nthreads = nthreads or min(32, os.cpu_count() or 1)
    max_samples = max_samples or BLOCK_SAMPLES
    block_bytes = block_bytes or max(1 << 20, int(max_samples * (signal_len * 11 + kmer_len * 30 + 256)))
    K, S = int(kmer_len), int(signal_len)
    carry = b""
    with open(path, "rb") as fh:
        while True:
            data = fh.read(block_bytes)
            eof = not data
            buf = carry + data
            if eof and buf and not buf.endswith(b"\n"):
                buf += b"\n"
            if not buf:
                return
            pos = 0
            base_addr = C.cast(C.c_char_p(buf), C.c_void_p).value      # bytes object: stable address while `buf` lives
            while pos < len(buf):
                cap = int(min(max_samples, (len(buf) - pos) // (2 * S + 2 * K + 20) + 16))   # >= lines that can fit
                kmer = np.empty((cap, K), np.int32)
                means = np.empty((cap, K), np.float32)
                stds = np.empty((cap, K), np.float32)
                lens = np.empty((cap, K), np.float32)
                signals = np.empty((cap, S), np.float32)
                labels = np.empty(cap, np.int32)
                info = np.empty((cap, 2), np.int64)
                consumed = C.c_int64(0)
                n = lib.dsb_tsv_parse_features(C.c_void_p(base_addr + pos), len(buf) - pos, K, S, cap, kmer.ctypes.data,
                                               means.ctypes.data, stds.ctypes.data, lens.ctypes.data, signals.ctypes.data,
                                               labels.ctypes.data, info.ctypes.data, C.byref(consumed), nthreads)
                if n < 0:
                    if n <= -1000:
                        raise ValueError("malformed feature line near sample {} of a block of {}".format(-n - 1000, path))
                    raise _lib.DsbError(n, "dsb_tsv_parse_features failed")
                if n > 0:
                    b = FeatureBlock()
                    b.n, b.kmer, b.means, b.stds, b.lens, b.signals, b.labels = n, kmer[:n], means[:n], stds[:n], lens[:n], signals[:n], labels[:n]
                    b.info_off, b.text, b.text_ptr = info[:n], buf, base_addr + pos
                    yield b
                if consumed.value == 0:
                    break
                pos += consumed.value
            carry = buf[pos:]
            if eof:
                return


def block_from_info(info_strings, kmer) -> FeatureBlock:
    """A block for :func:`format_calls` from in-memory samples: ``info_strings`` = the six tab-joined info columns per
    sample (what the reference calls ``sampleinfo``), ``kmer`` int32 [n, k] base codes (all must be 0..4)."""
    kmer = np.ascontiguousarray(kmer, np.int32)
    if kmer.size and (kmer.min() < 0 or kmer.max() > 4):
        raise KeyError("base outside ACGTN in a k-mer")          # the reference's code2base_dna lookup raises KeyError here
    enc = [s.encode() for s in info_strings]
    lens = np.fromiter((len(e) for e in enc), dtype=np.int64, count=len(enc))
    ends = np.cumsum(lens)
    b = FeatureBlock()
    b.n, b.kmer = len(enc), kmer
    b.text = b"".join(enc) + b"\0"
    b.text_ptr = C.cast(C.c_char_p(b.text), C.c_void_p).value
    b.info_off = np.stack([ends - lens, ends], axis=1).astype(np.int64)
    b.means = b.stds = b.lens = b.signals = b.labels = None
    return b


def format_calls(block, probs, pred) -> bytes:
    lib = _lib.load()
    n = block.n
    probs = np.ascontiguousarray(probs, np.float32)
    pred = np.ascontiguousarray(pred, np.int32)
    kmer = np.ascontiguousarray(block.kmer, np.int32)
    info = np.ascontiguousarray(block.info_off, np.int64)
    cap = int((info[:, 1] - info[:, 0]).sum()) + n * (kmer.shape[1] + 96) + 64
    out = C.create_string_buffer(cap)
    w = lib.dsb_tsv_format_calls(C.c_void_p(block.text_ptr), info.ctypes.data, kmer.ctypes.data, kmer.shape[1], probs.ctypes.data,
                                 pred.ctypes.data, n, out, cap)
    if w < 0:
        raise _lib.DsbError(w, "dsb_tsv_format_calls failed")
    return out.raw[:w]

"""The reference's binary sample record (§8f-2): ``'<' + kB + kf + kf + kH + sf + 1B`` = 1628 bytes for
bn17/sn360 — bases (uint8 codes), means (f32), stds (f32), signal lengths (uint16), signals (f32),
label (uint8).  Writer: reference scripts/generate_binary_feature_file.py:52-54; TF-side parser:
deepsignal/utils/tf_utils.py:7-28; same layout in deepsignal/denoise.py:27-28.

Here it is a zero-parse ``call_mods`` input: the file is memory-mapped as a NumPy structured array and the
five model-input arrays are strided views that go straight to ``Engine.forward``.
"""
from __future__ import annotations

import numpy as np


def record_dtype(base_num: int = 17, signal_num: int = 360) -> np.dtype:
    return np.dtype([("bases", "u1", (base_num,)), ("means", "<f4", (base_num,)), ("stds", "<f4", (base_num,)),
                     ("lens", "<u2", (base_num,)), ("signals", "<f4", (signal_num,)), ("label", "u1")])


def write_binary(path: str, kmer, means, stds, lens, signals, labels, base_num=17, signal_num=360) -> None:
    dt = record_dtype(base_num, signal_num)
    n = len(kmer)
    rec = np.zeros(n, dtype=dt)
    rec["bases"] = np.asarray(kmer, dtype=np.uint8)
    rec["means"] = np.asarray(means, dtype=np.float32)
    rec["stds"] = np.asarray(stds, dtype=np.float32)
    rec["lens"] = np.asarray(lens, dtype=np.uint16)      # struct 'H': lengths above 65535 cannot be stored
    rec["signals"] = np.asarray(signals, dtype=np.float32)
    rec["label"] = np.asarray(labels, dtype=np.uint8)
    rec.tofile(path)


def read_binary(path: str, base_num=17, signal_num=360, mmap=True):
    """Returns dict(kmer i32[n,k], means f32, stds f32, lens f32, signals f32, labels i32)."""
    dt = record_dtype(base_num, signal_num)
    rec = np.memmap(path, dtype=dt, mode="r") if mmap else np.fromfile(path, dtype=dt)
    return dict(kmer=rec["bases"].astype(np.int32), means=np.ascontiguousarray(rec["means"]),
                stds=np.ascontiguousarray(rec["stds"]), lens=rec["lens"].astype(np.float32),
                signals=np.ascontiguousarray(rec["signals"]), labels=rec["label"].astype(np.int32))

"""Engine — one libdsb200 context bound to one GPU.

Thin host object: owns the ``dsb_ctx``, hands numpy (host) or torch (device) buffers
to the C ABI and returns numpy / torch results.  All arithmetic happens in the
library's CUDA kernels; nothing here computes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DsbError


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _np(a, dtype, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected shape {}, got {}".format(shape, a.shape))
    return a


class Engine(object):
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if self.lib.dsb_abi_version() != _lib.ABI_VERSION:
            raise ImportError("libdsb200 ABI version mismatch")
        h = C.c_void_p()
        rc = self.lib.dsb_create(int(device), C.byref(h))
        if rc != 0:
            raise DsbError(rc, (self.lib.dsb_last_error(None) or b"").decode())
        self.h = h
        self.device = int(device)
        self.kmer_len = None
        self.signal_len = None
        self.flags = None

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            raise DsbError(rc, (self.lib.dsb_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.dsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ model
    def load_weights(self, ckpt: dict, is_cnn=True, is_rnn=True, is_base=True, kmer_len=17, signal_len=360):
        """``ckpt``: TF-variable-name -> float32 array (see deepsignal_b200.weights)."""
        names = sorted(ckpt)
        arrs = [np.ascontiguousarray(ckpt[k], dtype=np.float32) for k in names]
        tens = (_lib.Tensor * len(names))()
        keep = []
        for i, (k, a) in enumerate(zip(names, arrs)):
            kb = k.encode()
            keep.append(kb)
            tens[i].name = kb
            tens[i].data = a.ctypes.data
            tens[i].ndim = a.ndim
            for d in range(a.ndim):
                tens[i].shape[d] = a.shape[d]
        flags = (_lib.FLAG_CNN if is_cnn else 0) | (_lib.FLAG_RNN if is_rnn else 0) | (_lib.FLAG_BASE if is_base else 0)
        self._check(self.lib.dsb_load_weights(self.h, tens, len(names), flags, int(kmer_len), int(signal_len)))
        self.kmer_len, self.signal_len, self.flags = int(kmer_len), int(signal_len), flags

    def set_max_chunk(self, n: int):
        self._check(self.lib.dsb_set_max_chunk(self.h, int(n)))

    def set_plan_options(self, **kw):
        """Change how the forward plan is built (``dsb_plan_options``): fused_blocks / fused_front / lstm_seq select
        the fused kernels (default) or the one-launch-per-op references; ablate_* are timing experiments."""
        o = _lib.PlanOptions()
        self._check(self.lib.dsb_get_plan_options(self.h, C.byref(o)))
        for k, v in kw.items():
            if not hasattr(o, k):
                raise KeyError(k)
            setattr(o, k, int(v))
        self._check(self.lib.dsb_set_plan_options(self.h, C.byref(o)))

    def nonfinite_count(self) -> int:
        """Samples with non-finite logits since the last call (after a device-pointer forward; sync the stream first)."""
        n = C.c_int64(0)
        self._check(self.lib.dsb_nonfinite_count(self.h, C.byref(n)))
        return int(n.value)

    def workspace_bytes(self, n: int, precision=_lib.DEFAULT_PRECISION) -> int:
        return int(self.lib.dsb_workspace_bytes(self.h, int(n), _lib.PRECISIONS.get(precision, precision)))

    def forward(self, kmer, means, stds, lens, signals, precision=_lib.DEFAULT_PRECISION):
        """Host arrays in, host arrays out (``dsb_forward_host``): probs f32[n,2], pred int32[n]."""
        signals = _np(signals, np.float32)
        n = signals.shape[0]
        K, S = self.kmer_len, self.signal_len
        signals = _np(signals, np.float32, (n, S))
        kmer = _np(kmer, np.int32, (n, K)) if kmer is not None else None
        means = _np(means, np.float32, (n, K)) if means is not None else None
        stds = _np(stds, np.float32, (n, K)) if stds is not None else None
        lens = _np(lens, np.float32, (n, K)) if lens is not None else None
        probs = np.empty((n, 2), np.float32)
        pred = np.empty((n,), np.int32)
        self._check(self.lib.dsb_forward_host(self.h, _lib.PRECISIONS.get(precision, precision), n, _ptr(kmer), _ptr(means),
                                              _ptr(stds), _ptr(lens), _ptr(signals), _ptr(probs), _ptr(pred)))
        return probs, pred

    def forward_ptrs(self, n, kmer, means, stds, lens, signals, probs, pred, precision=_lib.DEFAULT_PRECISION, host=False, stream=0):
        """Raw-pointer call (ints): device pointers + stream -> dsb_forward_dev, host pointers -> dsb_forward_host."""
        vp = lambda x: C.c_void_p(int(x)) if x else None
        if host:
            self._check(self.lib.dsb_forward_host(self.h, _lib.PRECISIONS.get(precision, precision), int(n), vp(kmer), vp(means),
                                                  vp(stds), vp(lens), vp(signals), vp(probs), vp(pred)))
        else:
            self._check(self.lib.dsb_forward_dev(self.h, _lib.PRECISIONS.get(precision, precision), int(n), vp(kmer), vp(means),
                                                 vp(stds), vp(lens), vp(signals), vp(probs), vp(pred), vp(stream)))

    def forward_torch(self, kmer, means, stds, lens, signals, precision=_lib.DEFAULT_PRECISION, probs=None, pred=None):
        """Device-resident torch tensors in / out (``dsb_forward_dev`` on torch's current stream)."""
        import torch
        n = signals.shape[0]
        dev = signals.device
        if probs is None:
            probs = torch.empty((n, 2), dtype=torch.float32, device=dev)
        if pred is None:
            pred = torch.empty((n,), dtype=torch.int32, device=dev)
        for t, dt in ((kmer, torch.int32), (means, torch.float32), (stds, torch.float32), (lens, torch.float32),
                      (signals, torch.float32)):
            if t is not None and (t.dtype != dt or not t.is_contiguous() or t.device != dev):
                raise ValueError("inputs must be contiguous {} tensors on {}".format(dt, dev))
        stream = torch.cuda.current_stream(dev).cuda_stream
        dp = lambda t: 0 if t is None else t.data_ptr()
        self.forward_ptrs(n, dp(kmer), dp(means), dp(stds), dp(lens), dp(signals), dp(probs), dp(pred),
                          precision=precision, host=False, stream=stream)
        return probs, pred

    # ------------------------------------------------------------------ profiling
    def set_profiling(self, on: bool):
        self._check(self.lib.dsb_set_profiling(self.h, 1 if on else 0))

    def profile(self):
        out = {}
        for cls, name in enumerate(_lib.PROFILE_CLASSES):
            ms, n, fl = C.c_double(), C.c_int64(), C.c_double()
            self._check(self.lib.dsb_get_profile(self.h, cls, C.byref(ms), C.byref(n), C.byref(fl)))
            out[name] = dict(ms=ms.value, launches=n.value, flops=fl.value)
        return out

    def launch_count(self) -> int:
        return int(self.lib.dsb_launch_count(self.h))

    # ------------------------------------------------------------------ site frequencies
    def site_frequency(self, chrom_id, pos, p0, p1, label, use):
        """GPU segmented reduce over call_mods records keyed by (chrom_id, pos) (``dsb_site_frequency_host``).
        Returns dict(first, prob_0, prob_1, met, coverage) in first-appearance order of the sites."""
        n = len(pos)
        chrom_id, pos = _np(chrom_id, np.int32, (n,)), _np(pos, np.int64, (n,))
        p0, p1 = _np(p0, np.float64, (n,)), _np(p1, np.float64, (n,))
        label, use = _np(label, np.uint8, (n,)), _np(use, np.uint8, (n,))
        cap = max(1, n)
        first, met, cov = np.empty(cap, np.int64), np.empty(cap, np.int64), np.empty(cap, np.int64)
        s0, s1 = np.empty(cap, np.float64), np.empty(cap, np.float64)
        ns = C.c_int64(0)
        self._check(self.lib.dsb_site_frequency_host(self.h, n, _ptr(chrom_id), _ptr(pos), _ptr(p0), _ptr(p1), _ptr(label), _ptr(use),
                                                     cap, C.byref(ns), _ptr(first), _ptr(s0), _ptr(s1), _ptr(met), _ptr(cov)))
        k = int(ns.value)
        order = np.argsort(first[:k], kind="stable")
        return dict(first=first[:k][order], prob_0=s0[:k][order], prob_1=s1[:k][order], met=met[:k][order], coverage=cov[:k][order])

    # ------------------------------------------------------------------ extraction
    @staticmethod
    def _motif_bytes(motif_seqs):
        motif_seqs = list(motif_seqs)
        mlen = len(motif_seqs[0])
        if any(len(m) != mlen for m in motif_seqs):
            raise ValueError("all motifs must have the same length")
        return "".join(motif_seqs).encode("ascii"), len(motif_seqs), mlen

    def extract(self, reads, motif_seqs=("CG",), mod_loc=0, kmer_len=17, signal_len=360,
                normalize_method="mad", want64=True, capacity=None):
        """Run the extraction kernels on a :class:`deepsignal_b200.synth.ReadBatch` (host arrays).

        Returns a dict of numpy arrays, one row per targeted site (ordered by read, then location).
        """
        R = reads.n_reads
        rd = _lib.Reads()
        host = dict(
            raw=_np(reads.raw, np.int16), raw_off=_np(reads.raw_off, np.int64), scaling=_np(reads.scaling, np.float64),
            offset=_np(reads.offset, np.float64), ev_start=_np(reads.ev_start, np.int32),
            ev_len=_np(reads.ev_len, np.int32), ev_base=_np(reads.ev_base, np.uint8), ev_off=_np(reads.ev_off, np.int64),
            chrom_start=_np(reads.chrom_start, np.int64), strand_minus=_np(reads.strand_minus, np.uint8),
            chrom_len=_np(reads.chrom_len, np.int64))
        rd.n_reads = R
        for k, a in host.items():
            setattr(rd, k, a.ctypes.data)
        mb, nm, ml = self._motif_bytes(motif_seqs)
        prm = _lib.ExtractParams(int(kmer_len), int(signal_len), {"mad": 0, "zscore": 1}[normalize_method],
                                 int(mod_loc), nm, ml, mb)
        K, S = int(kmer_len), int(signal_len)
        if capacity is None:
            # Targeted sites are a few per cent of the bases (CpG: 1-2 %), while the hard upper bound — every event — would
            # cost ~5 KB of output buffer per BASE on host and device.  Start from 4 % and, if the library reports
            # DSB_ERR_CAPACITY (it returns the exact count), run again with exactly that many rows.
            n_ev = max(1, int(self.lib.dsb_extract_max_sites(C.byref(rd))))
            try:
                return self.extract(reads, motif_seqs, mod_loc, kmer_len, signal_len, normalize_method, want64,
                                    capacity=min(n_ev, max(1024, n_ev // 25)))
            except DsbError as e:
                if e.code != _lib.DSB_ERR_CAPACITY:
                    raise
                return self.extract(reads, motif_seqs, mod_loc, kmer_len, signal_len, normalize_method, want64,
                                    capacity=e.n_sites)
        out = dict(
            site_read=np.empty(capacity, np.int32), site_loc=np.empty(capacity, np.int32),
            pos=np.empty(capacity, np.int64), pos_in_strand=np.empty(capacity, np.int64),
            kmer=np.empty((capacity, K), np.int32), means=np.empty((capacity, K), np.float32),
            stds=np.empty((capacity, K), np.float32), lens=np.empty((capacity, K), np.float32),
            signals=np.empty((capacity, S), np.float32), lens_i32=np.empty((capacity, K), np.int32),
            window=np.empty((capacity, 3), np.int32))
        if want64:
            out.update(means64=np.empty((capacity, K), np.float64), stds64=np.empty((capacity, K), np.float64),
                       signals64=np.empty((capacity, S), np.float64))
        st = _lib.Sites()
        st.capacity = capacity
        for f in _lib.SITE_FIELDS:
            setattr(st, f, out[f].ctypes.data if f in out else None)
        n_sites = C.c_int64(0)
        shift = np.empty(R, np.float64)
        scale = np.empty(R, np.float64)
        rc = self.lib.dsb_extract_host(self.h, C.byref(rd), C.byref(prm), C.byref(st), C.byref(n_sites),
                                       _ptr(shift), _ptr(scale))
        if rc == _lib.DSB_ERR_CAPACITY:
            err = DsbError(rc, (self.lib.dsb_last_error(self.h) or b"").decode())
            err.n_sites = int(n_sites.value)
            raise err
        self._check(rc)
        n = int(n_sites.value)
        res = {k: v[:n] for k, v in out.items()}
        res["read_shift"], res["read_scale"], res["n_sites"] = shift, scale, n
        return res

    def extract_call_torch(self, dev_reads: dict, n_reads: int, motif_seqs=("CG",), mod_loc=0,
                           normalize_method="mad", precision=_lib.DEFAULT_PRECISION, capacity=None, site_bufs=None):
        """Fused extraction -> forward on device-resident reads (torch tensors keyed like ``dsb_reads``)."""
        import torch
        rd = _lib.Reads()
        rd.n_reads = int(n_reads)
        for k in ("raw", "raw_off", "scaling", "offset", "ev_start", "ev_len", "ev_base", "ev_off", "chrom_start",
                  "strand_minus", "chrom_len"):
            t = dev_reads.get(k)
            setattr(rd, k, t.data_ptr() if t is not None else None)
        mb, nm, ml = self._motif_bytes(motif_seqs)
        K, S = self.kmer_len, self.signal_len
        prm = _lib.ExtractParams(K, S, {"mad": 0, "zscore": 1}[normalize_method], int(mod_loc), nm, ml, mb)
        dev = dev_reads["raw"].device
        if site_bufs is None:
            site_bufs = self.alloc_site_bufs(capacity, dev)
        capacity = site_bufs["capacity"]
        st = _lib.Sites()
        st.capacity = capacity
        for f in _lib.SITE_FIELDS:
            setattr(st, f, site_bufs[f].data_ptr() if f in site_bufs else None)
        n_sites = C.c_int64(0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._check(self.lib.dsb_extract_call_dev(self.h, _lib.PRECISIONS.get(precision, precision), C.byref(rd), C.byref(prm),
                                                  C.byref(st), C.byref(n_sites), C.c_void_p(site_bufs["probs"].data_ptr()),
                                                  C.c_void_p(site_bufs["pred"].data_ptr()), C.c_void_p(stream)))
        return int(n_sites.value), site_bufs

    def reads_to_device(self, reads, dev):
        """Upload a :class:`ReadBatch` as the torch tensors ``dsb_extract_dev`` / ``dsb_extract_call_dev`` take."""
        import torch
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        return dict(raw=t(reads.raw, np.int16), raw_off=t(reads.raw_off, np.int64), scaling=t(reads.scaling, np.float64),
                    offset=t(reads.offset, np.float64), ev_start=t(reads.ev_start, np.int32), ev_len=t(reads.ev_len, np.int32),
                    ev_base=t(reads.ev_base, np.uint8), ev_off=t(reads.ev_off, np.int64),
                    chrom_start=t(reads.chrom_start, np.int64), strand_minus=t(reads.strand_minus, np.uint8),
                    chrom_len=t(reads.chrom_len, np.int64))

    def extract_torch(self, dev_reads: dict, n_reads: int, site_bufs: dict, motif_seqs=("CG",), mod_loc=0,
                      kmer_len=17, signal_len=360, normalize_method="mad"):
        """Device-resident extraction (``dsb_extract_dev`` on torch's current stream) into ``site_bufs``
        (from :meth:`alloc_site_bufs`); returns the number of sites."""
        import torch
        rd = _lib.Reads()
        rd.n_reads = int(n_reads)
        for k in ("raw", "raw_off", "scaling", "offset", "ev_start", "ev_len", "ev_base", "ev_off", "chrom_start",
                  "strand_minus", "chrom_len"):
            t = dev_reads.get(k)
            setattr(rd, k, t.data_ptr() if t is not None else None)
        mb, nm, ml = self._motif_bytes(motif_seqs)
        prm = _lib.ExtractParams(int(kmer_len), int(signal_len), {"mad": 0, "zscore": 1}[normalize_method],
                                 int(mod_loc), nm, ml, mb)
        st = _lib.Sites()
        st.capacity = site_bufs["capacity"]
        for f in _lib.SITE_FIELDS:
            setattr(st, f, site_bufs[f].data_ptr() if f in site_bufs else None)
        n_sites = C.c_int64(0)
        stream = torch.cuda.current_stream(dev_reads["raw"].device).cuda_stream
        self._check(self.lib.dsb_extract_dev(self.h, C.byref(rd), C.byref(prm), C.byref(st), C.byref(n_sites),
                                             None, None, C.c_void_p(stream)))
        return int(n_sites.value)

    def alloc_site_bufs(self, capacity: int, dev):
        import torch
        K, S = self.kmer_len or 17, self.signal_len or 360
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        return dict(capacity=int(capacity), site_read=mk((capacity,), torch.int32), site_loc=mk((capacity,), torch.int32),
                    pos=mk((capacity,), torch.int64), pos_in_strand=mk((capacity,), torch.int64),
                    kmer=mk((capacity, K), torch.int32), means=mk((capacity, K), torch.float32),
                    stds=mk((capacity, K), torch.float32), lens=mk((capacity, K), torch.float32),
                    signals=mk((capacity, S), torch.float32), probs=mk((capacity, 2), torch.float32),
                    pred=mk((capacity,), torch.int32))

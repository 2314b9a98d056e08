"""Long files conditioned in parts (BASELINE config 5): host side of the C ABI's afx_part_* calls.

A file is cut into sample-range parts (afx_part_plan); every part is downmixed / resampled / reduced by one
context (one GPU); the per-file reductions of TSampleAnalyser::LoadSample (SampleAnalyser.cpp:612-669) and
CalcEffectiveLength (:1715-1756) are combined between three phases -- in one process by a plain loop
(`analyze_in_parts`), across processes by an all-gather of a 112-byte record per phase (`analyze_sharded`; no
data-path collective beyond those scalars and the <= 3.5 MB analysis window).  The analysis of the <= 20 s
behind the trim point then runs on one context.

The part worker is pluggable (`job_factory`) so the protocol can be exercised on CPU with a numpy worker
(tests/test_longfile_gloo.py); the product worker is `PartJob`, which has no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api

INT64_MAX = (1 << 63) - 1


def plan_parts(nframes: int, src_rate: int, n_parts: int, sample_rate: int = 44100):
    """-> list of (src_begin, src_end, out_begin, out_end); host arithmetic only (no device)."""
    L = api.load_library()
    arr = (api.AfxPart * n_parts)()
    rc = L.afx_part_plan(sample_rate, nframes, src_rate, n_parts, arr)
    if rc != 0:
        raise api.AfxError("afx_part_plan failed: %d" % rc)
    return [(p.src_begin, p.src_end, p.out_begin, p.out_end) for p in arr]


def new_sums() -> api.AfxPartSums:
    s = api.AfxPartSums()
    api.load_library().afx_part_sums_init(C.byref(s))
    return s


def merge_sums(parts) -> api.AfxPartSums:
    L = api.load_library()
    acc = new_sums()
    for p in parts:
        L.afx_part_sums_merge(C.byref(acc), C.byref(p))
    return acc


def sums_to_array(s: api.AfxPartSums) -> np.ndarray:
    """112-byte record as 14 float64 / int64 lanes (the all-gather payload)."""
    return np.frombuffer(bytes(s), dtype=np.uint8).copy()


def sums_from_array(a: np.ndarray) -> api.AfxPartSums:
    return api.AfxPartSums.from_buffer_copy(np.ascontiguousarray(a, dtype=np.uint8).tobytes())


class PartJob:
    """One part on one context (GPU): afx_part_open .. afx_part_close."""

    def __init__(self, an: api.SampleAnalyser, whole: api.AfxFile, part, pcm_slice: np.ndarray):
        self._an, self._L = an, an._L
        self._slice = np.ascontiguousarray(pcm_slice)
        self.part = api.AfxPart(*part)
        self._whole = whole
        h = C.c_void_p()
        an._check(self._L.afx_part_open(an._ctx, C.byref(whole), C.byref(self.part),
                                        self._slice.ctypes.data if self._slice.size else None, C.byref(h)))
        self._h = h

    def peak(self) -> api.AfxPartSums:
        out = api.AfxPartSums()
        self._an._check(self._L.afx_part_peak(self._h, C.byref(out)))
        return out

    def trim(self, g: api.AfxPartSums) -> api.AfxPartSums:
        out = api.AfxPartSums()
        self._an._check(self._L.afx_part_trim(self._h, C.byref(g), C.byref(out)))
        return out

    def effective(self, g: api.AfxPartSums) -> api.AfxPartSums:
        out = api.AfxPartSums()
        self._an._check(self._L.afx_part_effective(self._h, C.byref(g), C.byref(out)))
        return out

    def read(self, begin: int, count: int, dst: np.ndarray) -> int:
        n = self._L.afx_part_read(self._h, begin, count, dst.ctypes.data)
        if n < 0:
            self._an._check(int(n))
        return int(n)

    def close(self):
        if self._h:
            self._L.afx_part_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def window_of(an: api.SampleAnalyser, whole: api.AfxFile, g: api.AfxPartSums):
    b, c = C.c_int64(), C.c_int64()
    an._check(an._L.afx_part_window(an._ctx, C.byref(whole), C.byref(g), C.byref(b), C.byref(c)))
    return b.value, c.value


def analyze_conditioned(an: api.SampleAnalyser, whole: api.AfxFile, g: api.AfxPartSums, mono: np.ndarray, begin: int) -> api.Batch:
    mono = np.ascontiguousarray(mono, dtype=np.float32)
    h = C.c_void_p()
    an._check(an._L.afx_analyze_conditioned(an._ctx, C.byref(whole), C.byref(g), mono.ctypes.data if mono.size else None,
                                            begin, mono.size, C.byref(h)))
    b = api.Batch.__new__(api.Batch)
    b._an, b._L, b._keep, b.n_files, b._files, b._h = an, an._L, mono, 1, None, h
    return b


def describe_whole(nframes: int, channels: int, rate: int, dtype, file_size: int | None = None) -> api.AfxFile:
    fmt = api.AFX_PCM_I16 if np.dtype(dtype) == np.int16 else api.AFX_PCM_F32
    item = np.dtype(dtype).itemsize
    fs = file_size if file_size is not None else 44 + nframes * channels * item
    return api.AfxFile(None, nframes, channels, rate, fmt, 16, fs)


def analyze_in_parts(analysers, pcm: np.ndarray, rate: int, n_parts: int | None = None, file_size: int | None = None,
                     job_factory=PartJob):
    """One process driving len(analysers) contexts (GPUs): part p runs on analysers[p % len]; returns the Batch of
    the analysis (file 0), made on analysers[0]."""
    a = pcm if pcm.ndim == 2 else pcm[:, None]
    n_parts = n_parts or len(analysers)
    whole = describe_whole(a.shape[0], a.shape[1], rate, a.dtype, file_size)
    parts = plan_parts(a.shape[0], rate, n_parts, analysers[0].sample_rate)
    jobs = [job_factory(analysers[p % len(analysers)], whole, parts[p], a[parts[p][0]:parts[p][1]]) for p in range(n_parts)]
    try:
        g = merge_sums([j.peak() for j in jobs])
        g = merge_sums([j.trim(g) for j in jobs])
        g = merge_sums([j.effective(g) for j in jobs])
        begin, count = window_of(analysers[0], whole, g)
        win = np.zeros(count, dtype=np.float32)
        got = sum(j.read(begin, count, win) for j in jobs)
        if got != count:
            raise api.AfxError("long file: the parts delivered %d of %d window samples" % (got, count))
    finally:
        for j in jobs:
            j.close()
    return analyze_conditioned(analysers[0], whole, g, win, begin)


def analyze_sharded(an, whole: api.AfxFile, part, pcm_slice, dist, analysis_rank: int = 0, job_factory=PartJob,
                    group=None, device=None, finish=analyze_conditioned, window=window_of, job=None):
    """One process per part (torch.distributed, rank = part index).  Returns the analysis Batch on `analysis_rank`,
    None elsewhere.  The only traffic: three all-gathers of a 112-byte record and a gather of the window pieces.
    `job`: a part job opened earlier (its host -> device copy is asynchronous, so the next file's part can be in
    flight while this file goes through its phases)."""
    import torch
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if job is None:
        job = job_factory(an, whole, part, pcm_slice)

    def combine(mine: api.AfxPartSums) -> api.AfxPartSums:
        t = torch.from_numpy(sums_to_array(mine))
        if device is not None:
            t = t.to(device)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t, group=group)
        return merge_sums([sums_from_array(o.cpu().numpy()) for o in outs])

    try:
        g = combine(job.peak())
        g = combine(job.trim(g))
        g = combine(job.effective(g))
        begin, count = window(an, whole, g)
        # every part contributes the window samples it owns; pieces are disjoint, so a sum-reduce to the analysis rank
        # of zero-filled buffers assembles the window exactly (x + 0 == x in float32)
        win = np.zeros(max(count, 1), dtype=np.float32)
        job.read(begin, count, win)
        t = torch.from_numpy(win)
        if device is not None:
            t = t.to(device)
        dist.reduce(t, dst=analysis_rank, op=dist.ReduceOp.SUM, group=group)
        win = t.cpu().numpy()[:count]
    finally:
        job.close()
    if rank != analysis_rank:
        return None
    return finish(an, whole, g, win, begin)

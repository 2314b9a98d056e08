"""ctypes binding of libafec_b200.so (include/afec_b200.h) and a thin host-side mirror of the
reference's extractor interface (`TSampleAnalyser`, Export/SampleAnalyser.h:28-64) for tests and
bench.  There is no fallback: importing works without a GPU, but creating an analyser raises when the
CUDA library is missing or no device is present."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import layout

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libafec_b200.so")

AFX_PCM_I16, AFX_PCM_F32, AFX_PCM_U8, AFX_PCM_I24, AFX_PCM_I32, AFX_PCM_F32U = 0, 1, 2, 3, 4, 5
AFX_PCM_I8, AFX_PCM_I16BE, AFX_PCM_I24BE, AFX_PCM_I32BE, AFX_PCM_F32UBE = 6, 7, 8, 9, 10
FEAT_SPECTRAL, FEAT_AMPLITUDE, FEAT_PEAKS, FEAT_BANDS = 1, 2, 4, 8
FEAT_PITCH, FEAT_AUTOCORR, FEAT_RHYTHM, FEAT_STATS = 16, 32, 64, 128
FEAT_ALL = 0xFF
FEAT_HIGHLEVEL = 0x100   # on top of FEAT_ALL: model-free high-level descriptors + the classification feature vector
FEAT_PACK = 0x200        # on top of FEAT_ALL: the row's msgpack BLOB images packed on the GPU
FEAT_EXT_MELCHROMA = 0x400   # extension: MFCC-13 over 40 mel filters + chroma-12 (needs FEAT_SPECTRAL)
N_BLOBS = 122
HAVE_RESAMPLE = True   # k_resample + host block plan (libresample HQ restatement)

EXPORTS = [
    "afx_abi_version", "afx_pcm_bytes", "afx_create", "afx_destroy", "afx_last_error", "afx_trim", "afx_host_alloc", "afx_host_free",
    "afx_batch_create", "afx_batch_upload", "afx_batch_compute", "afx_batch_download", "afx_batch_download_rows", "afx_batch_sync",
    "afx_analyze", "afx_batch_result", "afx_batch_free", "afx_batch_timings", "afx_batch_counters",
    "afx_batch_kernel_times", "afx_batch_conditioned", "afx_measure_fp64_peak", "afx_debug_fft",
    "afx_part_plan", "afx_part_sums_init", "afx_part_sums_merge", "afx_part_open", "afx_part_peak", "afx_part_trim",
    "afx_part_effective", "afx_part_window", "afx_part_read", "afx_part_close", "afx_analyze_conditioned", "afx_debug_rs_plan_check",
]


class AfxConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("sample_rate", C.c_int32), ("fft_size", C.c_int32),
                ("hop_size", C.c_int32), ("features", C.c_uint32), ("reserved", C.c_uint32)]


class AfxFile(C.Structure):
    _fields_ = [("pcm", C.c_void_p), ("nframes", C.c_int64), ("channels", C.c_int32), ("src_rate", C.c_int32),
                ("format", C.c_int32), ("bit_depth", C.c_int32), ("file_size", C.c_int64)]


class AfxFileResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_frames", C.c_int32), ("n_rhythm_frames", C.c_int32),
                ("hl_status", C.c_int32), ("header", C.POINTER(C.c_double)),
                ("fs", C.POINTER(C.c_double) * layout.N_FS), ("fv", C.POINTER(C.c_double) * layout.N_FV),
                ("stats", C.POINTER(C.c_double)), ("highlevel", C.POINTER(C.c_double)), ("hl_pitch", C.POINTER(C.c_double)),
                ("hl_signature", C.POINTER(C.c_double)), ("hl_features", C.POINTER(C.c_double)),
                ("ext_mfcc", C.POINTER(C.c_double)), ("ext_chroma", C.POINTER(C.c_double)), ("ext_chroma_index", C.POINTER(C.c_double)),
                ("packed", C.POINTER(C.c_ubyte)), ("packed_off", C.POINTER(C.c_uint32))]


class AfxPart(C.Structure):
    _fields_ = [("src_begin", C.c_int64), ("src_end", C.c_int64), ("out_begin", C.c_int64), ("out_end", C.c_int64)]


class AfxPartSums(C.Structure):
    _fields_ = [("maxabs", C.c_float), ("reserved", C.c_int32), ("sumsq", C.c_double), ("first", C.c_int64),
                ("last", C.c_int64), ("eff_first", C.c_int64 * 3), ("eff_last", C.c_int64 * 3)]


class AfxError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AfxError("libafec_b200.so is not built (run `python -m afec_b200.build`); "
                       "there is no CPU fallback for the descriptor path")
    L = C.CDLL(LIB_PATH)
    L.afx_abi_version.restype = C.c_int
    L.afx_pcm_bytes.argtypes = [C.c_int32]
    L.afx_create.argtypes = [C.POINTER(AfxConfig), C.POINTER(C.c_void_p)]
    L.afx_destroy.argtypes = [C.c_void_p]
    L.afx_destroy.restype = None
    L.afx_last_error.argtypes = [C.c_void_p]
    L.afx_trim.argtypes = [C.c_void_p]
    L.afx_last_error.restype = C.c_char_p
    L.afx_host_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    L.afx_host_free.argtypes = [C.c_void_p, C.c_void_p]
    L.afx_batch_create.argtypes = [C.c_void_p, C.POINTER(AfxFile), C.c_int32, C.POINTER(C.c_void_p)]
    for name in ("afx_batch_upload", "afx_batch_compute", "afx_batch_download", "afx_batch_download_rows", "afx_batch_sync"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.afx_analyze.argtypes = [C.c_void_p, C.POINTER(AfxFile), C.c_int32, C.POINTER(C.c_void_p)]
    L.afx_batch_result.argtypes = [C.c_void_p, C.c_int32, C.POINTER(AfxFileResult)]
    L.afx_batch_free.argtypes = [C.c_void_p]
    L.afx_batch_free.restype = None
    L.afx_batch_timings.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.afx_batch_counters.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 5
    L.afx_batch_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int32]
    L.afx_batch_conditioned.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
    L.afx_batch_conditioned.restype = C.c_int64
    L.afx_measure_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.afx_debug_fft.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.afx_part_plan.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.POINTER(AfxPart)]
    L.afx_part_sums_init.argtypes = [C.POINTER(AfxPartSums)]
    L.afx_part_sums_init.restype = None
    L.afx_part_sums_merge.argtypes = [C.POINTER(AfxPartSums), C.POINTER(AfxPartSums)]
    L.afx_part_sums_merge.restype = None
    L.afx_part_open.argtypes = [C.c_void_p, C.POINTER(AfxFile), C.POINTER(AfxPart), C.c_void_p, C.POINTER(C.c_void_p)]
    L.afx_part_peak.argtypes = [C.c_void_p, C.POINTER(AfxPartSums)]
    L.afx_part_trim.argtypes = [C.c_void_p, C.POINTER(AfxPartSums), C.POINTER(AfxPartSums)]
    L.afx_part_effective.argtypes = [C.c_void_p, C.POINTER(AfxPartSums), C.POINTER(AfxPartSums)]
    L.afx_part_window.argtypes = [C.c_void_p, C.POINTER(AfxFile), C.POINTER(AfxPartSums), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.afx_part_read.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    L.afx_part_read.restype = C.c_int64
    L.afx_part_close.argtypes = [C.c_void_p]
    L.afx_part_close.restype = None
    L.afx_analyze_conditioned.argtypes = [C.c_void_p, C.POINTER(AfxFile), C.POINTER(AfxPartSums), C.c_void_p, C.c_int64, C.c_int64,
                                          C.POINTER(C.c_void_p)]
    L.afx_debug_rs_plan_check.argtypes = [C.c_int32, C.c_int64, C.c_int32]
    _lib = L
    return L


class PinnedArena:
    """Pinned host memory from afx_host_alloc, exposed as a numpy uint8 array (decode ring slot)."""

    def __init__(self, analyser: "SampleAnalyser", nbytes: int):
        self._an = analyser
        p = C.c_void_p()
        analyser._check(analyser._L.afx_host_alloc(analyser._ctx, max(1, nbytes), C.byref(p)))
        self.ptr = p.value
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(1, nbytes)).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self._an._L.afx_host_free(self._an._ctx, C.c_void_p(self.ptr))
            self.ptr = None


class Batch:
    def __init__(self, analyser: "SampleAnalyser", files, keepalive):
        self._an = analyser
        self._L = analyser._L
        self._keep = keepalive
        self.n_files = len(files)
        arr = (AfxFile * max(1, len(files)))(*files)
        self._files = arr
        h = C.c_void_p()
        analyser._check(self._L.afx_batch_create(analyser._ctx, arr, len(files), C.byref(h)))
        self._h = h

    def upload(self):
        self._an._check(self._L.afx_batch_upload(self._h))

    def compute(self):
        self._an._check(self._L.afx_batch_compute(self._h))

    def download(self):
        self._an._check(self._L.afx_batch_download(self._h))

    def download_rows(self):
        self._an._check(self._L.afx_batch_download_rows(self._h))

    def sync(self):
        self._an._check(self._L.afx_batch_sync(self._h))

    def packed_blobs(self, i: int) -> list:
        """File i's AFX_N_BLOBS msgpack BLOB images (contexts created with FEAT_PACK), in afec-ll.db column order."""
        r = self.raw_result(i)
        if r.status != 0:
            return []
        if not r.packed:
            raise AfxError("the context was not created with FEAT_PACK")
        off = np.ctypeslib.as_array(r.packed_off, (N_BLOBS + 1,))
        raw = np.ctypeslib.as_array(r.packed, (int(off[-1]),))
        return [raw[int(off[k]):int(off[k + 1])].tobytes() for k in range(N_BLOBS)]

    def run(self):
        self.upload(); self.compute(); self.download(); self.sync()
        return self

    def timings(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._L.afx_batch_timings(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def counters(self):
        v = [C.c_int64() for _ in range(5)]
        self._L.afx_batch_counters(self._h, *[C.byref(x) for x in v])
        return dict(zip(["kernel_launches", "h2d_bytes", "d2h_bytes", "main_frames", "rhythm_frames"],
                        [x.value for x in v]))

    def kernel_times(self):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = self._L.afx_batch_kernel_times(self._h, names, ms, 64)
        return [(names[i].decode(), ms[i]) for i in range(max(0, n))]

    def conditioned(self, i: int) -> np.ndarray:
        n = self._L.afx_batch_conditioned(self._h, i, None, 0)
        if n < 0:
            raise AfxError("afx_batch_conditioned failed: %d" % n)
        out = np.zeros(int(n), dtype=np.float64)
        if n:
            self._L.afx_batch_conditioned(self._h, i, out.ctypes.data, int(n))
        return out

    def raw_result(self, i: int) -> AfxFileResult:
        r = AfxFileResult()
        self._an._check(self._L.afx_batch_result(self._h, i, C.byref(r)))
        return r

    def result(self, i: int) -> layout.FileResult:
        """Copy file i's values into a layout.FileResult (missing feature groups stay zero)."""
        r = self.raw_result(i)
        out = layout.FileResult(status=r.status, F=r.n_frames, Fr=r.n_rhythm_frames)
        out.header = np.ctypeslib.as_array(r.header, (layout.N_HEADER,)).copy()
        F, Fr = r.n_frames, r.n_rhythm_frames
        for s in range(layout.N_FS):
            ln = F if s < layout.N_FS_MAIN else Fr
            if r.fs[s] and ln > 0:
                out.fs.append(np.ctypeslib.as_array(r.fs[s], (ln,)).copy())
            else:
                out.fs.append(np.zeros(ln))
        for v, (_, nb) in enumerate(layout.FRAMED_VECTORS):
            if r.fv[v] and F > 0:
                out.fv.append(np.ctypeslib.as_array(r.fv[v], (F * nb,)).copy().reshape(F, nb))
            else:
                out.fv.append(np.zeros((F, nb)))
        if r.stats:
            out.stats = np.ctypeslib.as_array(r.stats, (layout.N_SERIES * layout.N_STATS,)).copy().reshape(
                layout.N_SERIES, layout.N_STATS)
        return out

    def extension(self, i: int):
        """File i's extension outputs (contexts created with FEAT_EXT_MELCHROMA): (mfcc [F, 13], chroma [F, 12], chroma_index [F])."""
        r = self.raw_result(i)
        if r.status != 0:
            return None
        if not r.ext_mfcc:
            raise AfxError("the context was not created with FEAT_EXT_MELCHROMA")
        F = r.n_frames
        if F <= 0:
            return np.zeros((0, 13)), np.zeros((0, 12)), np.zeros(0)
        return (np.ctypeslib.as_array(r.ext_mfcc, (F * 13,)).copy().reshape(F, 13),
                np.ctypeslib.as_array(r.ext_chroma, (F * 12,)).copy().reshape(F, 12),
                np.ctypeslib.as_array(r.ext_chroma_index, (F,)).copy())

    def highlevel(self, i: int) -> layout.HighLevelResult:
        """File i's high-level derivations + classification features (contexts created with FEAT_HIGHLEVEL)."""
        r = self.raw_result(i)
        out = layout.HighLevelResult(status=r.status if r.status else (100 if r.hl_status else 0), F=r.n_frames)
        if r.status != 0:
            return out
        if not r.highlevel:
            raise AfxError("the context was not created with FEAT_HIGHLEVEL")
        F = r.n_frames
        out.scalars = np.ctypeslib.as_array(r.highlevel, (layout.N_HL,)).copy()
        out.pitch = np.ctypeslib.as_array(r.hl_pitch, (F,)).copy() if F > 0 else np.zeros(0)
        out.peak = np.ctypeslib.as_array(r.fs[1], (F,)).copy() if F > 0 else np.zeros(0)
        out.signature = np.ctypeslib.as_array(r.hl_signature, (layout.HL_SIGNATURE_FRAMES * layout.HL_SIGNATURE_BANDS,)).copy().reshape(
            layout.HL_SIGNATURE_FRAMES, layout.HL_SIGNATURE_BANDS)
        out.features = np.ctypeslib.as_array(r.hl_features, (layout.HL_N_FEATURES,)).copy()
        return out

    def free(self):
        if self._h:
            self._L.afx_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SampleAnalyser:
    """Host mirror of `TSampleAnalyser(SampleRate, FftFrameSize, HopFrameSize)`
    (Export/SampleAnalyser.h:33-36); `analyze_pcm` is `Analyze` minus the file decoding."""

    def __init__(self, sample_rate: int = 44100, fft_size: int = 2048, hop_size: int = 1024,
                 device: int = 0, features: int = FEAT_ALL):
        self._L = load_library()
        self._ctx = C.c_void_p()
        cfg = AfxConfig(device, sample_rate, fft_size, hop_size, features, 0)
        rc = self._L.afx_create(C.byref(cfg), C.byref(self._ctx))
        if rc != 0:
            msg = self._L.afx_last_error(None)
            raise AfxError("afx_create failed (%d): %s" % (rc, msg.decode() if msg else ""))
        self.sample_rate, self.fft_size, self.hop_size, self.features = sample_rate, fft_size, hop_size, features

    def _check(self, rc: int):
        if rc != 0:
            msg = self._L.afx_last_error(self._ctx)
            raise AfxError("afec_b200 call failed (%d): %s" % (rc, msg.decode() if msg else ""))

    @staticmethod
    def describe(pcm: np.ndarray, rate: int, bit_depth: int = 16, file_size: int | None = None):
        """numpy int16 / float32 array [n] or [n, channels] -> (AfxFile, keepalive array)."""
        a = np.asarray(pcm)
        if a.ndim == 1:
            a = a[:, None]
        if a.dtype == np.int16:
            fmt = AFX_PCM_I16
        elif a.dtype == np.float32:
            fmt = AFX_PCM_F32
        else:
            raise TypeError("PCM must be int16 or float32 (16-bit range)")
        a = np.ascontiguousarray(a)
        n, ch = a.shape
        fs = file_size if file_size is not None else 44 + a.size * a.itemsize
        return AfxFile(a.ctypes.data if a.size else None, n, ch, rate, fmt, bit_depth, fs), a

    def batch(self, pcms, rates, file_sizes=None) -> Batch:
        files, keep = [], []
        for i, (p, r) in enumerate(zip(pcms, rates)):
            f, a = self.describe(p, r, file_size=None if file_sizes is None else file_sizes[i])
            files.append(f)
            keep.append(a)
        return Batch(self, files, keep)

    def batch_from_descriptors(self, files, keepalive=None) -> Batch:
        return Batch(self, files, keepalive)

    def analyze_pcm(self, pcms, rates, file_sizes=None) -> list:
        b = self.batch(pcms, rates, file_sizes).run()
        out = [b.result(i) for i in range(b.n_files)]
        b.free()
        return out

    def fp64_peak_tflops(self) -> float:
        v = C.c_double()
        self._check(self._L.afx_measure_fp64_peak(self._ctx, C.byref(v)))
        return v.value

    def debug_fft(self, x: np.ndarray) -> np.ndarray:
        """Forward FFT of complex128 rows [batch, n] (n = 256 / 1024 / 2048) through the kernels' FFT core."""
        a = np.ascontiguousarray(x, dtype=np.complex128)
        if a.ndim == 1:
            a = a[None, :]
        out = np.empty_like(a)
        self._check(self._L.afx_debug_fft(self._ctx, a.shape[1], a.shape[0], a.ctypes.data, out.ctypes.data))
        return out

    def pinned(self, nbytes: int) -> PinnedArena:
        return PinnedArena(self, nbytes)

    def trim(self):
        """Release the context's device buffers (they grow again on demand)."""
        self._check(self._L.afx_trim(self._ctx))

    def close(self):
        if self._ctx:
            self._L.afx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

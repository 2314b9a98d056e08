"""TEST INFRASTRUCTURE: ctypes binding of oracle/libafec_oracle.so (plain-C restatement)
and a runner for oracle/_ref/afec_ref (the unmodified reference, when it was built)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
import wave

import numpy as np

from afec_b200 import layout

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libafec_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "afec_ref")

_lib = None


def build() -> None:
    """Compile the C restatement (gcc) if missing or stale."""
    src = os.path.join(HERE, "afec_oracle.c")
    if (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libafec_oracle.so"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.afxo_analyze.restype = C.c_long
        L.afxo_analyze.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p, C.c_long,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_long)]
        L.afxo_stats13.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.afxo_peaks.restype = C.c_int
        L.afxo_peaks.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        for name in ("afxo_variance", "afxo_centroid", "afxo_spread", "afxo_skewness",
                     "afxo_kurtosis", "afxo_flatness", "afxo_sum", "afxo_mean", "afxo_median",
                     "afxo_gmean", "afxo_min", "afxo_max"):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.c_void_p, C.c_int]
        L.afxo_flux.restype = C.c_double
        L.afxo_flux.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.afxo_fft.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.afxo_set_fft_variant.argtypes = [C.c_int]
        L.afxo_set_fft_variant.restype = None
        L.afxo_tables.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.afxo_condition.restype = C.c_int
        L.afxo_condition.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                     C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.afxo_highlevel.restype = C.c_int
        L.afxo_highlevel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def to_planar_f32(pcm: np.ndarray) -> np.ndarray:
    """pcm: [nframes] or [nframes, channels], int16 or float32 (16-bit range) -> [channels, nframes] f32."""
    a = np.asarray(pcm)
    if a.ndim == 1:
        a = a[:, None]
    return np.ascontiguousarray(a.T.astype(np.float32))


def analyze(pcm: np.ndarray, src_rate: int = 44100, hop: int = 1024, sample_rate: int = 44100,
            fft_size: int = 2048, file_size: int = 0, bit_depth: int = 16) -> layout.FileResult:
    planar = to_planar_f32(pcm)
    ch, n = planar.shape
    F, Fr, need = C.c_int(0), C.c_int(0), C.c_long(0)
    L = lib()
    cap = layout.record_doubles(n // hop + 8, n // 128 + 8) + 2 * (8 * 2048)
    out = np.zeros(cap, dtype=np.float64)
    rc = L.afxo_analyze(planar.ctypes.data, ch, n, src_rate, sample_rate, fft_size, hop, file_size,
                        bit_depth, out.ctypes.data, cap, C.byref(F), C.byref(Fr), C.byref(need))
    if rc == -2:
        cap = need.value
        out = np.zeros(cap, dtype=np.float64)
        rc = L.afxo_analyze(planar.ctypes.data, ch, n, src_rate, sample_rate, fft_size, hop, file_size,
                            bit_depth, out.ctypes.data, cap, C.byref(F), C.byref(Fr), C.byref(need))
    if rc < 0:
        return layout.FileResult(status=int(-rc))
    hdr = b"AFXD" + np.array([0, F.value, Fr.value], dtype="<i4").tobytes()
    r, _ = layout.parse_record(memoryview(hdr + out[:rc].tobytes()), 0)
    return r


def condition(pcm: np.ndarray, src_rate: int = 44100, sample_rate: int = 44100, fft_size: int = 2048):
    planar = to_planar_f32(pcm)
    ch, n = planar.shape
    L = lib()
    off, pk, rms = C.c_int(0), C.c_float(0), C.c_float(0)
    cap = int(n * max(1.0, sample_rate / src_rate)) + 4 * fft_size
    data = np.zeros(cap, dtype=np.float64)
    ln = L.afxo_condition(planar.ctypes.data, ch, n, src_rate, sample_rate, fft_size, data.ctypes.data,
                          cap, C.byref(off), C.byref(pk), C.byref(rms))
    return data[:ln].copy(), off.value, pk.value, rms.value


def set_fft_variant(v: int) -> None:
    """0: the restatement's decimation-in-time FFT; 1: an identical transform with another rounding order (tests that
    show which outputs are decided by FFT rounding noise)."""
    lib().afxo_set_fft_variant(int(v))


_silence_pad = None


def silence_pad() -> np.ndarray:
    """The values the reference pads short files with in the classification features: the last frame of a silent
    0.5-s sample analysed at hop 1024 (SampleClassificationDescriptors.cpp:330-368) -- 14 frequency bands, then
    spectral_rms, spectral_flatness, spectral_flux, spectral_contrast, spectral_complexity, f0_confidence, amplitude_rms."""
    global _silence_pad
    if _silence_pad is None:
        r = analyze(np.zeros(22050, dtype=np.int16), hop=1024)
        last = r.F - 1
        _silence_pad = np.array(list(r.series("frequency_bands")[last][:14]) +
                                [r.series(n)[last] for n in ("spectral_rms", "spectral_flatness", "spectral_flux", "spectral_contrast",
                                                             "spectral_complexity", "f0_confidence", "amplitude_rms")], dtype=np.float64)
    return _silence_pad


def highlevel(r: layout.FileResult, peak_value: float, rms_value: float, sample_rate: int = 44100) -> layout.HighLevelResult:
    """High-level derivations + classification features of one low-level result (afxo_highlevel)."""
    out = layout.HighLevelResult(F=r.F)
    if r.status != 0:
        out.status = r.status
        return out
    body = layout.record_body(r)
    pad = silence_pad()
    out.pitch = np.zeros(r.F)
    out.signature = np.zeros((layout.HL_SIGNATURE_FRAMES, layout.HL_SIGNATURE_BANDS))
    out.features = np.zeros(layout.HL_N_FEATURES)
    rc = lib().afxo_highlevel(body.ctypes.data, r.F, r.Fr, sample_rate, float(peak_value), float(rms_value), pad.ctypes.data,
                              out.scalars.ctypes.data, out.pitch.ctypes.data, out.signature.ctypes.data, out.features.ctypes.data)
    out.peak = np.array(r.series("amplitude_peak"), dtype=np.float64)
    out.status = 0 if rc == 0 else 100 - rc
    return out


def reference_analyze_highlevel(pcms, rates, hop: int = 1024, tmpdir: str | None = None) -> list:
    """Run the real reference with kHighLevelDescriptors (no classification models) -> [(FileResult, HighLevelResult)]."""
    assert have_reference()
    with tempfile.TemporaryDirectory(dir=tmpdir) as d:
        paths = []
        for i, (p, r) in enumerate(zip(pcms, rates)):
            path = os.path.join(d, "f%05d.wav" % i)
            write_wav(path, p, r)
            paths.append(path)
        out = os.path.join(d, "dump.bin")
        subprocess.run([REF_BIN, "dumphl", str(hop), out] + paths, check=True, env=dict(os.environ, HOME=d),
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return layout.load_dump_highlevel(out)


def scalar_stat(name: str, x) -> float:
    x = np.ascontiguousarray(x, dtype=np.float64)
    return float(getattr(lib(), "afxo_" + name)(x.ctypes.data if len(x) else None, len(x)))


def stats13(x) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(13)
    lib().afxo_stats13(x.ctypes.data, len(x), out.ctypes.data)
    return out


def peaks(x, thr: float):
    x = np.ascontiguousarray(x, dtype=np.float64)
    bins = np.zeros(len(x) + 8, dtype=np.int32)
    vals = np.zeros(len(x) + 8)
    n = lib().afxo_peaks(x.ctypes.data, len(x), float(thr), bins.ctypes.data, vals.ctypes.data)
    return list(zip(bins[:n].tolist(), vals[:n].tolist()))


# --------------------------------------------------------------------------------------------
# the unmodified reference binary (travels to the GPU box inside oracle/_ref/)

def have_reference() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def write_wav(path: str, pcm: np.ndarray, rate: int) -> None:
    a = np.asarray(pcm)
    if a.ndim == 1:
        a = a[:, None]
    assert a.dtype == np.int16
    with wave.open(path, "wb") as w:
        w.setnchannels(a.shape[1])
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(np.ascontiguousarray(a).tobytes())


def reference_analyze(pcms, rates, hop: int = 1024, tmpdir: str | None = None) -> list:
    """Run the real reference on int16 PCM arrays (written as WAV) -> list[FileResult]."""
    assert have_reference()
    with tempfile.TemporaryDirectory(dir=tmpdir) as d:
        paths = []
        for i, (p, r) in enumerate(zip(pcms, rates)):
            path = os.path.join(d, "f%05d.wav" % i)
            write_wav(path, p, r)
            paths.append(path)
        out = os.path.join(d, "dump.bin")
        env = dict(os.environ, HOME=d)
        subprocess.run([REF_BIN, "dump", str(hop), out] + paths, check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return layout.load_dump(out)

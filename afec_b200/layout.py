"""Descriptor naming, ordering and the flat "AFXD" result layout.

Mirrors the reference's low-level descriptor set and order
(Source/Crawler/FeatureExtraction/Source/SampleDescriptors.cpp:154-203,
Export/SampleDescriptors.h:396-466).  The same layout is produced by the CUDA
library (include/afec_b200.h), the CPU oracle (oracle/afec_oracle.c) and the
reference harness (oracle/ref_harness.cpp), so parity checks are array compares.

Per file:
  header   : 32 float64 scalars (HEADER_NAMES, rest reserved)
  fs[s]    : 24 framed scalar series; s < 22 have F values, s = 22, 23 have Fr
  fv[v]    : 7 framed vector series, frame-major [F][nbands]
  stats    : 136 series x 13 statistics (24 scalar series, then each vector
             descriptor band-major)
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

HEADER_NAMES = [
    "file_size", "file_length", "file_sample_rate", "file_channel_count", "file_bit_depth",
    "effectve_length_48dB", "effectve_length_24dB", "effectve_length_12dB", "analyzation_offset",
    "rhythm_complex_onset_count", "rhythm_complex_onset_contrast",
    "rhythm_complex_onset_frequency_mean", "rhythm_complex_onset_strength",
    "rhythm_complex_tempo", "rhythm_complex_tempo_confidence",
    "rhythm_percussive_onset_count", "rhythm_percussive_onset_contrast",
    "rhythm_percussive_onset_frequency_mean", "rhythm_percussive_onset_strength",
    "rhythm_percussive_tempo", "rhythm_percussive_tempo_confidence",
    "rhythm_final_tempo", "rhythm_final_tempo_confidence",
    # conditioning outputs (not DB columns; reference TSampleData members)
    "peak_value", "rms_value", "data_offset", "data_length",
]
N_HEADER = 32

FRAMED_SCALARS = [
    "amplitude_silence", "amplitude_peak", "amplitude_rms", "amplitude_envelope",
    "spectral_rms", "spectral_centroid", "spectral_rolloff", "spectral_spread",
    "spectral_skewness", "spectral_kurtosis", "spectral_flatness", "spectral_inharmonicity",
    "spectral_complexity", "spectral_contrast", "spectral_flux", "f0", "f0_confidence",
    "failsafe_f0", "tristimulus1", "tristimulus2", "tristimulus3", "auto_correlation",
    "rhythm_complex_onsets", "rhythm_percussive_onsets",
]
N_FS = 24
N_FS_MAIN = 22  # the last two run on the rhythm frame grid (Fr)

FRAMED_VECTORS = [
    ("spectral_rms_bands", 14), ("spectral_flatness_bands", 14), ("spectral_flux_bands", 14),
    ("spectral_complexity_bands", 14), ("spectral_contrast_bands", 14),
    ("frequency_bands", 28), ("cepstrum_bands", 14),
]
N_FV = 7
FV_TOTAL_BANDS = sum(n for _, n in FRAMED_VECTORS)  # 112

STAT_NAMES = ["min", "max", "median", "mean", "gmean", "variance", "centroid", "spread",
              "skewness", "kurtosis", "flatness", "dmean", "dvariance"]
N_STATS = 13
N_SERIES = N_FS + FV_TOTAL_BANDS  # 136

# integer-valued outputs that must match bit-exactly
INTEGER_SERIES = {"amplitude_silence", "spectral_rolloff", "spectral_complexity",
                  "spectral_complexity_bands"}


@dataclass
class FileResult:
    status: int = 0
    F: int = 0
    Fr: int = 0
    header: np.ndarray = field(default_factory=lambda: np.zeros(N_HEADER))
    fs: list = field(default_factory=list)       # 24 arrays
    fv: list = field(default_factory=list)       # 7 arrays [F][nb]
    stats: np.ndarray = field(default_factory=lambda: np.zeros((N_SERIES, N_STATS)))

    def scalar(self, name: str) -> float:
        return float(self.header[HEADER_NAMES.index(name)])

    def series(self, name: str) -> np.ndarray:
        if name in FRAMED_SCALARS:
            return self.fs[FRAMED_SCALARS.index(name)]
        for i, (n, _) in enumerate(FRAMED_VECTORS):
            if n == name:
                return self.fv[i]
        raise KeyError(name)

    def series_stats(self, name: str) -> np.ndarray:
        """13 stats (scalar series) or [nbands][13] (vector series)."""
        if name in FRAMED_SCALARS:
            return self.stats[FRAMED_SCALARS.index(name)]
        off = N_FS
        for n, nb in FRAMED_VECTORS:
            if n == name:
                return self.stats[off:off + nb]
            off += nb
        raise KeyError(name)


def record_doubles(F: int, Fr: int) -> int:
    return (N_HEADER + N_FS_MAIN * F + 2 * Fr + FV_TOTAL_BANDS * F + N_SERIES * N_STATS)


def parse_record(buf: memoryview, pos: int):
    """Parse one AFXD record at byte offset pos -> (FileResult, new_pos)."""
    if bytes(buf[pos:pos + 4]) != b"AFXD":
        raise ValueError("bad AFXD magic at %d" % pos)
    status, = struct.unpack_from("<i", buf, pos + 4)
    pos += 8
    r = FileResult(status=status)
    if status != 0:
        return r, pos
    F, Fr = struct.unpack_from("<ii", buf, pos)
    pos += 8
    n = record_doubles(F, Fr)
    d = np.frombuffer(buf, dtype="<f8", count=n, offset=pos).copy()
    pos += 8 * n
    r.F, r.Fr = F, Fr
    o = 0
    r.header = d[o:o + N_HEADER]; o += N_HEADER
    for s in range(N_FS):
        ln = F if s < N_FS_MAIN else Fr
        r.fs.append(d[o:o + ln]); o += ln
    for _, nb in FRAMED_VECTORS:
        r.fv.append(d[o:o + F * nb].reshape(F, nb)); o += F * nb
    r.stats = d[o:o + N_SERIES * N_STATS].reshape(N_SERIES, N_STATS); o += N_SERIES * N_STATS
    assert o == n
    return r, pos


def parse_dump(data: bytes) -> list:
    buf = memoryview(data)
    pos, out = 0, []
    while pos < len(buf):
        r, pos = parse_record(buf, pos)
        out.append(r)
    return out


def load_dump(path: str) -> list:
    with open(path, "rb") as f:
        return parse_dump(f.read())


# ---- high-level derivations that need no classification model (SampleAnalyser.cpp:1232-1606) and the
# classification feature vector (SampleClassificationDescriptors.cpp:330-560) --------------------------------
HL_SCALARS = ["base_note", "base_note_confidence", "peak_db", "rms_db", "bpm", "bpm_confidence", "brightness",
              "noisiness", "harmonicity", "spectral_flatness", "spectral_flux", "spectral_complexity",
              "spectral_contrast", "spectral_inharmonicity", "pitch_confidence"]
N_HL = 16                   # 15 scalars + 1 reserved
HL_SIGNATURE_FRAMES, HL_SIGNATURE_BANDS = 64, 14
HL_N_FEATURES = 1680        # 35 rows of 48 (SampleClassificationDescriptors.cpp:536-547 pads to the time-series width)


@dataclass
class HighLevelResult:
    status: int = 0
    F: int = 0
    scalars: np.ndarray = field(default_factory=lambda: np.zeros(N_HL))
    pitch: np.ndarray = field(default_factory=lambda: np.zeros(0))          # [F] MIDI notes
    peak: np.ndarray = field(default_factory=lambda: np.zeros(0))           # [F] == amplitude_peak
    signature: np.ndarray = field(default_factory=lambda: np.zeros((HL_SIGNATURE_FRAMES, HL_SIGNATURE_BANDS)))
    features: np.ndarray = field(default_factory=lambda: np.zeros(HL_N_FEATURES))

    def scalar(self, name: str) -> float:
        return float(self.scalars[HL_SCALARS.index(name)])


def record_body(r: "FileResult") -> np.ndarray:
    """The flat float64 body of an AFXD record (what oracle.afxo_highlevel takes)."""
    return np.ascontiguousarray(np.concatenate([r.header] + [np.ravel(a) for a in r.fs] + [np.ravel(a) for a in r.fv]
                                               + [np.ravel(r.stats)]), dtype=np.float64)


def parse_highlevel(buf: memoryview, pos: int):
    """Parse one AFXH record at byte offset pos -> (HighLevelResult, new_pos)."""
    if bytes(buf[pos:pos + 4]) != b"AFXH":
        raise ValueError("bad AFXH magic at %d" % pos)
    status, = struct.unpack_from("<i", buf, pos + 4)
    pos += 8
    r = HighLevelResult(status=status)
    if status != 0:
        return r, pos
    F, nf = struct.unpack_from("<ii", buf, pos)
    pos += 8
    n = N_HL + 2 * F + HL_SIGNATURE_FRAMES * HL_SIGNATURE_BANDS + nf
    d = np.frombuffer(buf, dtype="<f8", count=n, offset=pos).copy()
    pos += 8 * n
    r.F = F
    o = 0
    r.scalars = d[o:o + N_HL]; o += N_HL
    r.pitch = d[o:o + F]; o += F
    r.peak = d[o:o + F]; o += F
    r.signature = d[o:o + HL_SIGNATURE_FRAMES * HL_SIGNATURE_BANDS].reshape(HL_SIGNATURE_FRAMES, HL_SIGNATURE_BANDS)
    o += HL_SIGNATURE_FRAMES * HL_SIGNATURE_BANDS
    r.features = d[o:o + nf]
    return r, pos


def load_dump_highlevel(path: str) -> list:
    """A `afec_ref dumphl` file: per input file an AFXD record followed by an AFXH record -> [(FileResult, HighLevelResult)]."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    pos, out = 0, []
    while pos < len(buf):
        ll, pos = parse_record(buf, pos)
        hl, pos = parse_highlevel(buf, pos)
        out.append((ll, hl))
    return out

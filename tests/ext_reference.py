"""CPU restatement (numpy, float64) of the mel-40 / MFCC-13 / chroma-12 extension (afec_b200/csrc/afx_ext.cu; BASELINE
configs[2]).  The reference has no counterpart: this file IS the definition the CUDA kernels are held
to -- "validated against a builder-written CPU restatement" (north_star item 4).  Tests only.

Per main frame, from the magnitude spectrum Mag[0..1023] of the path (2 x Hann window, FFT / N; SampleAnalyser.cpp:814-847):
  E_m   = sum_k Mag[k] W_mel[m][k]            40 triangular equal-gain filters, 20 Hz .. 15.5 kHz, mel = 1127 ln(1 + f / 700),
                                              integer peak bins int(f_peak / 22050 * 1024) (LibXtract init.c:237-378 laid over all bins)
  mfcc_n = sum_m log(max(E_m, 2e-42)) cos(pi n (m + 1/2) / 40), n = 0..12
  C_c   = sum_k Mag[k]^2 W_chr[c][k]          bins 65.4 Hz .. 8372 Hz shared linearly between the two nearest semitones
  chroma = C / max(C) (0 when silent), chroma_index = first argmax
"""
import numpy as np

NBINS, NFFT, SR = 1024, 2048, 44100.0
NMEL, NCHR, NMFCC = 40, 12, 13


def weights():
    """-> (W_mel [40, 1024], W_chr [12, 1024]) as float32 values held in float64 (the kernels' weights are float32)."""
    wm = np.zeros((NMEL, NBINS), dtype=np.float32)
    nyq, fmin, fmax = SR / 2.0, 20.0, 15500.0
    mel_max, mel_min = 1127 * np.log(1 + fmax / 700), 1127 * np.log(1 + fmin / 700)
    bw = (mel_max - mel_min) / NMEL
    lin = [fmin if n == 0 else 700 * (np.exp((mel_min + bw * n) / 1127) - 1) for n in range(NMEL + 2)]
    peak = [int(v / nyq * NBINS) for v in lin]
    for n in range(NMEL):
        p0, p1, p2 = peak[n], peak[n + 1], peak[n + 2]
        for k in range(p0, min(p1, NBINS - 1) + 1):
            wm[n, k] = np.float32((k - p0) / (p1 - p0)) if p1 > p0 else np.float32(1.0)
        for k in range(p1 + 1, min(p2, NBINS - 1) + 1):
            wm[n, k] = np.float32((p2 - k) / (p2 - p1)) if p2 > p1 else np.float32(0.0)
    wc = np.zeros((NCHR, NBINS), dtype=np.float32)
    for k in range(1, NBINS):
        f = k * SR / NFFT
        if f < 65.40639132514966 or f > 8372.018089619156:
            continue
        pitch = 69.0 + 12.0 * np.log2(f / 440.0)
        lower = np.floor(pitch)
        frac = pitch - lower
        c0 = int(lower) % 12
        wc[c0, k] += np.float32(1.0 - frac)
        wc[(c0 + 1) % 12, k] += np.float32(frac)
    return wm.astype(np.float64), wc.astype(np.float64)


def magnitude_spectra(mdata: np.ndarray, hop: int) -> np.ndarray:
    """[F, 1024] magnitude spectra of the conditioned signal as the path forms them (SampleAnalyser.cpp:814-847)."""
    n = np.arange(NFFT)
    win = (0.5 * (1.0 - np.cos(2.0 * np.pi * n / (NFFT - 1)))) * 2.0
    L = min(len(mdata), 882000)
    F = (L - NFFT) // hop + 1
    out = np.zeros((max(F, 0), NBINS))
    for t in range(F):
        X = np.fft.rfft(mdata[t * hop:t * hop + NFFT] * win) / NFFT
        out[t] = np.abs(X[:NBINS])
    return out


def analyze(mdata: np.ndarray, hop: int):
    """-> (mfcc [F, 13], chroma [F, 12], chroma_index [F], mel energies [F, 40], chroma energies [F, 12])"""
    wm, wc = weights()
    mag = magnitude_spectra(np.asarray(mdata, dtype=np.float64), hop)
    E = mag @ wm.T
    C = (mag * mag) @ wc.T
    m = np.arange(NMEL)
    dct = np.cos(np.pi * np.arange(NMFCC)[:, None] * (m[None, :] + 0.5) / NMEL)
    mfcc = np.log(np.maximum(E, 2e-42)) @ dct.T
    mx = C.max(axis=1, keepdims=True) if len(C) else np.zeros((0, 1))
    with np.errstate(invalid="ignore", divide="ignore"):
        chroma = np.where(mx > 0, C / np.where(mx > 0, mx, 1.0), 0.0)
    idx = np.argmax(C, axis=1).astype(np.float64) if len(C) else np.zeros(0)
    return mfcc, chroma, idx, E, C

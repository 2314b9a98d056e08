"""Seeded synthetic PCM corpus: decaying sines, band-limited noise
bursts and clicks with optional leading silence -- exercises trim, F0, onsets, whitening."""
from __future__ import annotations

import numpy as np


def one_shot(seed: int, seconds: float = 3.0, rate: int = 44100, channels: int = 1) -> np.ndarray:
    """int16 PCM, shape [n] (mono) or [n, channels]."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * rate))
    t = np.arange(n) / rate
    x = np.zeros(n)
    f = float(np.exp(rng.uniform(np.log(40.0), np.log(4000.0))))
    tau = float(rng.uniform(0.05, 2.0))
    x += np.sin(2 * np.pi * f * t) * np.exp(-t / tau)
    # a couple of harmonics
    for h in (2, 3):
        x += rng.uniform(0.0, 0.4) * np.sin(2 * np.pi * f * h * t) * np.exp(-t / (tau / h))
    # band-limited noise burst (one-pole low-passed white noise)
    nb = rng.standard_normal(n)
    a = float(rng.uniform(0.05, 0.9))
    lp = np.empty(n)
    acc = 0.0
    # vectorised one-pole via cumulative trick is unstable; do a cheap FIR approximation instead
    k = np.exp(-np.arange(64) * (1 - a))
    lp = np.convolve(nb, k / k.sum(), mode="same")
    del acc
    x += rng.uniform(0.02, 0.5) * lp * np.exp(-t / rng.uniform(0.05, 1.0))
    # clicks
    for _ in range(int(rng.integers(0, 9))):
        p = int(rng.integers(0, max(1, n - 64)))
        x[p:p + 32] += rng.uniform(0.2, 1.0) * np.hanning(32) * rng.choice([-1.0, 1.0])
    # leading silence 0..200 ms on ~half of the files
    if rng.uniform() < 0.5:
        lead = int(rng.uniform(0, 0.2) * rate)
        x = np.concatenate([np.zeros(lead), x])[:n]
    peak = float(rng.uniform(0.1, 1.0))
    x = x / (np.max(np.abs(x)) + 1e-12) * peak
    pcm = np.round(x * 32767.0).astype(np.int16)
    if channels == 1:
        return pcm
    # decorrelated channels: per-channel gain + small delay
    out = np.zeros((n, channels), dtype=np.int16)
    for c in range(channels):
        g = float(rng.uniform(0.5, 1.0))
        d = int(rng.integers(0, 32))
        out[d:, c] = np.round(pcm[:n - d].astype(np.float64) * g).astype(np.int16)
    return out


def corpus(n_files: int, seconds: float = 3.0, rate: int = 44100, channels: int = 1,
           seed0: int = 0, min_seconds: float | None = None) -> list:
    """List of int16 arrays.  With min_seconds set, durations are U(min_seconds, seconds)."""
    out = []
    for i in range(n_files):
        if min_seconds is not None:
            rng = np.random.default_rng(10_000_019 + seed0 + i)
            sec = float(rng.uniform(min_seconds, seconds))
        else:
            sec = seconds
        out.append(one_shot(seed0 + i, sec, rate, channels))
    return out


def tiled_corpus(n_files: int, n_unique: int, seconds: float = 3.0, rate: int = 44100,
                 channels: int = 1, seed0: int = 0, min_seconds: float | None = None) -> list:
    """n_files entries cycling over n_unique generated files (bench-sized corpora, cheap to build)."""
    base = corpus(min(n_unique, n_files), seconds, rate, channels, seed0, min_seconds)
    return [base[i % len(base)] for i in range(n_files)]

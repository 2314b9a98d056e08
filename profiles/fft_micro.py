"""Micro-benchmark of the FFT core (afx_debug_fft): run under ncu to read k_debug_fft durations."""
import sys, numpy as np
sys.path.insert(0, '.')
from afec_b200 import api
an = api.SampleAnalyser(44100, 2048, 1024)
rng = np.random.default_rng(0)
for n, batch in ((2048, 20000), (1024, 40000), (256, 160000)):
    x = rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))
    an.debug_fft(x); an.debug_fft(x)
an.close()

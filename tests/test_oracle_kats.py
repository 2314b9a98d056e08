"""Known-answer tests lifted from the reference's only value-pinning test,
Source/Crawler/FeatureExtraction/Test/TestStatistics.cpp:16-113, applied to the CPU oracle."""
import numpy as np
import pytest


def test_min_max_peaks(oracle_lib):
    o = oracle_lib
    seq = [1, 2, 2, 2, 0, 5, 6]
    assert o.scalar_stat("min", seq) == 0
    assert o.scalar_stat("max", seq) == 6
    assert o.peaks(seq, 0) == [(2, 2.0), (6, 6.0)]      # TestStatistics.cpp:26-31
    assert o.peaks(seq, 2) == [(6, 6.0)]


@pytest.mark.parametrize("seq", [[1, 2, 3, 4, 5, 6], [6, 5, 4, 3, 2, 1], [3, 2, 4, 6, 5, 1], [4, 3, 6, 5, 2, 1]])
def test_sum_variance_median_mean(oracle_lib, seq):
    o = oracle_lib
    assert o.scalar_stat("sum", seq) == 21
    assert abs(o.scalar_stat("variance", seq) - 2.9) <= 0.1
    assert abs(np.sqrt(o.scalar_stat("variance", seq)) - 1.7) <= 0.1
    assert o.scalar_stat("median", seq) == 3            # lower median
    assert abs(o.scalar_stat("mean", seq) - 21.0 / 6) <= 1e-16
    assert abs(o.scalar_stat("gmean", seq) - 3) <= 1.0


def test_centroid(oracle_lib):
    o = oracle_lib
    assert abs(o.scalar_stat("centroid", [1, 2, 3, 4, 5, 6]) - 3.0) <= 1.0
    assert abs(o.scalar_stat("centroid", [1, 1, 1, 1, 6, 8]) - 4.0) <= 1.0
    assert abs(o.scalar_stat("centroid", [1, 20, 4, 6, 5, 1]) - 2.0) <= 1.0
    assert abs(o.scalar_stat("centroid", [1, 1, 1, 1, 1, 1]) - 2.5) <= 0.001


def test_single_item(oracle_lib):
    o = oracle_lib
    v = [123456789.0]
    assert o.scalar_stat("sum", v) == v[0]
    assert o.scalar_stat("median", v) == v[0]
    assert o.scalar_stat("variance", v) == 0
    assert o.scalar_stat("mean", v) == v[0]
    assert o.scalar_stat("gmean", v) == v[0]
    assert o.scalar_stat("centroid", v) == 0
    assert o.scalar_stat("spread", v) == 0


def test_empty(oracle_lib):
    o = oracle_lib
    for name in ("sum", "median", "variance", "mean", "gmean", "centroid", "spread"):
        assert o.scalar_stat(name, []) == 0


def test_calc13_conventions(oracle_lib):
    """TStatistics::Calc (Statistics.cpp:12-90): length-1 series only set min/max/mean."""
    s = oracle_lib.stats13([4.0])
    assert s[0] == 4 and s[1] == 4 and s[3] == 4
    assert (np.delete(s, [0, 1, 3]) == 0).all()
    assert (oracle_lib.stats13([]) == 0).all()
    s = oracle_lib.stats13([1.0, 3.0])
    assert s[11] == 0 and s[12] == 0 and s[2] == 1.0      # no derivative stats for n == 2


def test_fft_roundtrip(oracle_lib):
    """Source/Core/AudioTypes/Test/TestFourier.cpp:16-83: forward o inverse = identity (1e-4)."""
    import ctypes
    rng = np.random.default_rng(0)
    re = rng.uniform(-1, 1, 2048)
    im = np.zeros(2048)
    re0 = re.copy()
    L = oracle_lib.lib()
    L.afxo_fft(re.ctypes.data, im.ctypes.data, 2048, 1)
    ref = np.fft.ifft(re0) * 2048            # exp(+i) convention
    assert np.allclose(re, ref.real, atol=1e-9) and np.allclose(im, ref.imag, atol=1e-9)
    L.afxo_fft(re.ctypes.data, im.ctypes.data, 2048, -1)
    assert np.allclose(re / 2048, re0, atol=1e-4) and np.allclose(im / 2048, 0, atol=1e-4)

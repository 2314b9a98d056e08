"""k_pack (afx_pack.cu, AFX_FEAT_PACK): the BLOB images of an afec-ll.db row packed on the GPU must be the msgpack bytes
the reference stores -- checked against Python's msgpack of the very values the same batch returns as arrays (every
double 0xcb + big endian, array headers 0x9X / 0xdc as msgpack-c; tests/test_host_sink.py pins that encoding against the
reference-written database)."""
import msgpack
import numpy as np
import pytest

from afec_b200 import api, layout, synth

pytestmark = pytest.mark.gpu


def expected_blobs(r: layout.FileResult) -> list:
    out = []
    for s in range(layout.N_FS):
        out.append(msgpack.packb([float(v) for v in r.fs[s]], use_single_float=False))
    si = layout.N_FS
    for v, (_, nb) in enumerate(layout.FRAMED_VECTORS):
        out.append(msgpack.packb([[float(x) for x in row] for row in r.fv[v]], use_single_float=False))
        for k in range(layout.N_STATS):
            out.append(msgpack.packb([float(x) for x in r.stats[si:si + nb, k]], use_single_float=False))
        si += nb
    return out


def test_packed_rows_are_the_msgpack_of_the_arrays():
    pcms = [synth.one_shot(1500 + i, 0.05 + 0.6 * i) for i in range(5)]
    pcms += [synth.one_shot(1510, 21.0), synth.one_shot(1511, 0.02), np.zeros(30000, dtype=np.int16), np.zeros((0,), dtype=np.int16),
             synth.one_shot(1512, 0.4, rate=48000, channels=2)]
    rates = [44100] * 9 + [48000]
    an = api.SampleAnalyser(44100, 2048, 1024, features=api.FEAT_ALL | api.FEAT_PACK)
    b = an.batch(pcms, rates).run()
    for i in range(len(pcms)):
        r = b.result(i)
        got = b.packed_blobs(i)
        if r.status != 0:
            assert got == []
            continue
        want = expected_blobs(r)
        assert len(got) == api.N_BLOBS == len(want)
        for k, (g, w) in enumerate(zip(got, want)):
            assert g == w, "file %d blob %d: %d vs %d bytes" % (i, k, len(g), len(w))
    # rows-only download: the same packed rows, no arrays
    b.free()
    b = an.batch(pcms[:3], rates[:3])
    b.upload(); b.compute(); b.download_rows(); b.sync()
    full = an.batch  # noqa: F841
    for i in range(3):
        raw = b.raw_result(i)
        assert not raw.fs[0] and not raw.fv[0] and raw.stats and raw.packed
    rows = [b.packed_blobs(i) for i in range(3)]
    b.free()
    b2 = an.batch(pcms[:3], rates[:3]).run()
    assert rows == [b2.packed_blobs(i) for i in range(3)]
    b2.free()
    an.close()


def test_pack_needs_the_full_low_level_set():
    with pytest.raises(api.AfxError):
        api.SampleAnalyser(44100, 2048, 1024, features=api.FEAT_SPECTRAL | api.FEAT_PACK)
    an = api.SampleAnalyser(44100, 2048, 1024, features=api.FEAT_ALL)
    b = an.batch([synth.one_shot(1, 0.2)], [44100])
    b.upload(); b.compute()
    with pytest.raises(api.AfxError):
        b.download_rows()
    b.free()
    an.close()

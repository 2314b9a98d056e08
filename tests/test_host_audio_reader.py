"""The host readers (afec_b200/host/audio_reader.cpp): WAV and AIFF / AIFC headers, sample types, error messages
(WaveFile.cpp:372-407, AifFile.cpp:150-372) -- CPU only; the device-side sample conversion of the raw bytes is tested on
the GPU (tests/test_gpu_host.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import audio_files
from afec_b200 import build as afx_build, synth


@pytest.fixture(scope="module")
def host():
    L = C.CDLL(afx_build.HOST_LIB)
    L.afxh_probe_audio.argtypes = [C.c_char_p, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.c_char_p, C.c_int]
    L.afxh_read_audio.argtypes = [C.c_char_p, C.c_void_p, C.c_longlong]
    L.afxh_read_audio.restype = C.c_longlong
    return L


def probe(L, path):
    fr, nb = C.c_longlong(), C.c_longlong()
    ch, rate, bits, fmt = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    err = C.create_string_buffer(256)
    rc = L.afxh_probe_audio(path.encode(), C.byref(fr), C.byref(ch), C.byref(rate), C.byref(bits), C.byref(fmt), C.byref(nb), err, 256)
    return rc, dict(frames=fr.value, channels=ch.value, rate=rate.value, bits=bits.value, format=fmt.value, nbytes=nb.value), err.value.decode()


@pytest.mark.parametrize("name", sorted(audio_files.format_cases()))
def test_probe_and_read(host, tmp_path, name):
    writer, kind, kw, code = audio_files.format_cases()[name]
    pcm = synth.one_shot(33, 0.05, channels=2)
    values = audio_files.quantise(pcm, kind)
    path = str(tmp_path / name)
    writer(path, values, kind, 48000, **kw)
    rc, info, err = probe(host, path)
    assert rc == 0, err
    bits = {"u8": 8, "i8": 8, "i16": 16, "i24": 24, "i32": 32, "f32": 32, "f64": 64}[kind]
    bps = {0: 2, 1: 4, 2: 1, 3: 3, 4: 4, 5: 4, 6: 1, 7: 2, 8: 3, 9: 4, 10: 4}[code]
    assert info == dict(frames=len(pcm), channels=2, rate=48000, bits=bits, format=code, nbytes=len(pcm) * 2 * bps)
    buf = np.zeros(info["nbytes"], dtype=np.uint8)
    assert host.afxh_read_audio(path.encode(), buf.ctypes.data, len(buf)) == len(buf)
    if kind == "f64":            # the one host-side conversion: float32 in 16-bit range
        assert np.array_equal(buf.view(np.float32).reshape(-1, 2), audio_files.to_float16range(values, kind))
    else:                        # raw bytes as in the file
        big = name.endswith((".aif", ".aiff")) or (name.endswith(".aifc") and not name.startswith("sowt"))
        assert buf.tobytes() == audio_files._sample_bytes(values, kind, big)


def test_error_messages(host, tmp_path):
    p = str(tmp_path / "x.wav")
    open(p, "wb").write(b"this is not a wave file")
    assert probe(host, p)[2] == "Not a valid WAV file."
    p = str(tmp_path / "x.aiff")
    open(p, "wb").write(b"FORM\x00\x00\x00\x04AIFF")
    assert probe(host, p)[2] == "This is not a valid AIFF file!"
    p = str(tmp_path / "in24.aifc")
    audio_files.write_aiff(p, audio_files.quantise(synth.one_shot(1, 0.01), "i24"), "i24", 44100, compression="in24")
    assert probe(host, p)[2] == "Unsupported compressed AIFC file type."         # AifFile.cpp:192-211 does not list it (four-cc order)
    assert probe(host, str(tmp_path / "missing.wav"))[2] == "Failed to open the file for reading."
    p = str(tmp_path / "adpcm.wav")
    audio_files.write_wav(p, audio_files.quantise(synth.one_shot(1, 0.01), "i16"), "i16", 44100)
    raw = bytearray(open(p, "rb").read()); raw[20] = 2                            # format tag 2 = ADPCM
    open(p, "wb").write(bytes(raw))
    assert probe(host, p)[2] == "Unsupported file format."

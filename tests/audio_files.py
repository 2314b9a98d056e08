"""Writers for small test files in the sample formats the host readers accept (tests only), and the numpy restatement
of the reference's sample conversion (CoreFileFormats/Export/SampleConverter.h:392-518) for each of them."""
import struct

import numpy as np


def _frames(pcm):
    a = np.asarray(pcm)
    return a[:, None] if a.ndim == 1 else a


def quantise(pcm16: np.ndarray, kind: str) -> np.ndarray:
    """int16 PCM -> the integer / float sample values a file of `kind` holds."""
    a = _frames(pcm16).astype(np.int64)
    if kind == "u8":
        return ((a >> 8) + 128).astype(np.uint8)
    if kind == "i8":
        return (a >> 8).astype(np.int8)
    if kind == "i16":
        return a.astype(np.int16)
    if kind == "i24":
        return (a * 256 + 77).astype(np.int32)                    # uses the low byte too
    if kind == "i32":
        return (a * 65536 + 12345).astype(np.int32)
    if kind == "f32":
        return (a / 32768.0 * 1.02).astype(np.float32)            # a little over full scale: exercises the clamp
    if kind == "f64":
        return (a / 32768.0 * 1.02).astype(np.float64)
    raise ValueError(kind)


def to_float16range(values: np.ndarray, kind: str) -> np.ndarray:
    """What the reference's decoders hand LoadSample for those sample values: float32 in 16-bit range."""
    v = values
    if kind == "u8":
        return ((v.astype(np.int32) - 128) << 8).astype(np.float32)
    if kind == "i8":
        return (v.astype(np.int32) << 8).astype(np.float32)
    if kind == "i16":
        return v.astype(np.float32)
    if kind == "i24":
        return ((v.astype(np.int64) << 8).astype(np.float64) * 32768.0 / 2147483648.0).astype(np.float32)
    if kind == "i32":
        return np.clip((v.astype(np.float64) * 32768.0 / 2147483648.0).astype(np.float32), -32768.0, 32767.0).astype(np.float32)
    if kind in ("f32", "f64"):
        return np.clip(v.astype(np.float64) * 32768.0, -32768.0, 32767.0).astype(np.float32)
    raise ValueError(kind)


def _sample_bytes(values: np.ndarray, kind: str, big: bool) -> bytes:
    if kind in ("u8", "i8"):
        return values.tobytes()
    if kind == "i24":
        b = values.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
        return (b[:, ::-1] if big else b).tobytes()
    dt = {"i16": "i2", "i32": "i4", "f32": "f4", "f64": "f8"}[kind]
    return values.astype((">" if big else "<") + dt).tobytes()


def write_wav(path: str, values: np.ndarray, kind: str, rate: int, extensible: bool = False):
    ch = values.shape[1]
    bits = {"u8": 8, "i16": 16, "i24": 24, "i32": 32, "f32": 32, "f64": 64}[kind]
    tag = 3 if kind in ("f32", "f64") else 1
    data = _sample_bytes(values, kind, False)
    block = ch * bits // 8
    if extensible:
        sub = struct.pack("<H", tag) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
        fmt = struct.pack("<HHIIHHHHI", 0xFFFE, ch, rate, rate * block, block, bits, 22, bits, 0) + sub
    else:
        fmt = struct.pack("<HHIIHH", tag, ch, rate, rate * block, block, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 4) + b"abcd" + \
        b"data" + struct.pack("<I", len(data)) + data + (b"\x00" if len(data) & 1 else b"")
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def _extended(rate: float) -> bytes:
    import math
    m, e = math.frexp(rate)                      # rate = m * 2^e, 0.5 <= m < 1
    return struct.pack(">HQ", e - 1 + 16383, int(m * (1 << 64)))


def write_aiff(path: str, values: np.ndarray, kind: str, rate: int, compression: str | None = None):
    """AIFF (compression None) or AIFC with the given four-cc: 'NONE', 'twos', 'sowt' (little endian), 'fl32', 'fl64'."""
    ch = values.shape[1]
    bits = {"i8": 8, "i16": 16, "i24": 24, "i32": 32, "f32": 32, "f64": 64}[kind]
    big = compression != "sowt"
    data = _sample_bytes(values, kind, big)
    comm = struct.pack(">hIh", ch, values.shape[0], bits) + _extended(float(rate))
    if compression is not None:
        name = b"\x04test\x00"                    # pascal string, padded to even length
        comm += compression.encode() + name
    ssnd = struct.pack(">II", 0, 0) + data
    body = (b"AIFC" if compression is not None else b"AIFF")
    if compression is not None:
        body += b"FVER" + struct.pack(">II", 4, 0xA2805140)
    body += b"COMM" + struct.pack(">I", len(comm)) + comm + b"SSND" + struct.pack(">I", len(ssnd)) + ssnd + (b"\x00" if len(ssnd) & 1 else b"")
    with open(path, "wb") as f:
        f.write(b"FORM" + struct.pack(">I", len(body)) + body)


# file name -> (writer, kind, kwargs, expected AFX_PCM_* code)
def format_cases():
    return {
        "u8.wav": (write_wav, "u8", {}, 2), "i16.wav": (write_wav, "i16", {}, 0), "i24.wav": (write_wav, "i24", {}, 3),
        "i32.wav": (write_wav, "i32", {}, 4), "f32.wav": (write_wav, "f32", {}, 5), "f64.wav": (write_wav, "f64", {}, 1),
        "i24_ext.wav": (write_wav, "i24", {"extensible": True}, 3),
        "i8.aiff": (write_aiff, "i8", {}, 6), "i16.aif": (write_aiff, "i16", {}, 7), "i24.aiff": (write_aiff, "i24", {}, 8),
        "i32.aiff": (write_aiff, "i32", {}, 9), "sowt.aifc": (write_aiff, "i16", {"compression": "sowt"}, 0),
        "none.aifc": (write_aiff, "i16", {"compression": "NONE"}, 7), "fl32.aifc": (write_aiff, "f32", {"compression": "fl32"}, 10),
        "fl64.aifc": (write_aiff, "f64", {"compression": "fl64"}, 1),
    }

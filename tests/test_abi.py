"""The C-ABI library loads and exports every symbol include/afec_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from afec_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "afec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(afx_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    from afec_b200 import api
    assert sorted(api.EXPORTS) == syms


def test_abi_version(lib_path):
    lib = ctypes.CDLL(lib_path)
    assert lib.afx_abi_version() == 3


def test_create_fails_loudly_without_gpu(lib_path):
    """No CPU fallback: without a device afx_create must fail (on a GPU box it succeeds)."""
    import torch
    from afec_b200 import api
    if torch.cuda.is_available():
        a = api.SampleAnalyser()
        a.close()
    else:
        with pytest.raises(api.AfxError):
            api.SampleAnalyser()


def test_bad_config_rejected(lib_path):
    from afec_b200 import api
    L = api.load_library()
    ctx = ctypes.c_void_p()
    for cfg in (api.AfxConfig(0, 48000, 2048, 1024, 0xFF, 0), api.AfxConfig(0, 44100, 1024, 512, 0xFF, 0),
                api.AfxConfig(0, 44100, 2048, 1000, 0xFF, 0)):
        assert L.afx_create(ctypes.byref(cfg), ctypes.byref(ctx)) == -1
        assert b"afx_create" in L.afx_last_error(None)


def test_host_adapter_library_loads(lib_path):
    """The C++ host adapter (reference extractor interface + sqlite sink) links against the CUDA library
    and exports its C entry points; the crawler executable exists."""
    from afec_b200 import build
    host = ctypes.CDLL(build.HOST_LIB)
    for s in ("afxh_schema", "afxh_write_row", "afxh_extract_files", "afxh_extract_one"):
        assert hasattr(host, s), "missing export " + s
    assert os.access(build.CRAWLER, os.X_OK)

"""End to end through the C++ host adapter on the GPU: WAV files -> afec-b200-crawler -> afec-ll.db,
compared with the database the unmodified reference wrote for the same files (tests/golden/ref_ll.db)."""
import json
import os
import shutil
import sqlite3
import subprocess
import time

import pytest

import db_cases
import dbcompare
from afec_b200 import build as afx_build

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ll.db")


def run_crawler(args):
    r = subprocess.run([afx_build.CRAWLER] + args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


@pytest.mark.parametrize("load", ["sqlite", "direct"])
def test_crawler_database_is_a_drop_in(tmp_path, load):
    """`direct`: the fresh database is written in sqlite's file format by the repo's own writer (--direct-load,
    afec_b200/host/direct_db_writer.cpp); everything after that -- the comparison with the reference's database, the
    incremental crawls that delete and replace rows -- goes through sqlite on that file."""
    d = str(tmp_path)
    os.makedirs(os.path.join(d, "sub"))
    paths = db_cases.write_files(d)
    shutil.move(paths[1], os.path.join(d, "sub", os.path.basename(paths[1])))
    db = os.path.join(d, "afec-ll.db")
    out = run_crawler(["-l", "low", "-o", db] + (["--direct-load"] if load == "direct" else []) + [d])
    assert "4 files found, 4 to analyse" in out
    stats = json.loads(out.strip().splitlines()[-1])
    assert stats["files"] == 3 and stats["failed"] == 1
    got, sql, pragmas = dbcompare.rows(db)
    want, ref_sql, _ = dbcompare.rows(GOLDEN)
    assert sql == ref_sql
    assert pragmas == {"user_version": 2, "encoding": "UTF-8", "journal_mode": "wal"}
    assert set(got) == set(want)
    for name in want:
        errs = dbcompare.compare_row(got[name], want[name])
        assert not errs, name + ":\n" + "\n".join(errs[:20])
    c = sqlite3.connect(db)
    names = sorted(r[0] for r in c.execute("select filename from assets"))
    assert "sub/pad_stereo.wav" in names and "kick.wav" in names          # relative to the database directory
    assert c.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
    c.close()

    # incremental crawl (Crawler.cpp:934-998): nothing to do, then one modified and one vanished file
    assert "0 to analyse" in run_crawler(["-o", db, d])
    os.remove(os.path.join(d, "hat_48k.wav"))
    future = time.time() + 5
    os.utime(os.path.join(d, "kick.wav"), (future, future))
    out = run_crawler(["-o", db, d])
    assert "3 files found, 1 to analyse, 1 removed" in out
    c = sqlite3.connect(db)
    assert c.execute("select count(*) from assets").fetchone()[0] == 3
    c.close()


def test_single_file_extract_entry_point(tmp_path):
    """TSampleAnalyser::Extract(FileName, pPool, PoolLock) on one file, plus the load-failure row."""
    import ctypes as C
    L = C.CDLL(afx_build.HOST_LIB)
    L.afxh_extract_one.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    d = str(tmp_path)
    db_cases.write_files(d)
    db = os.path.join(d, "one.db")
    for name in ("kick.wav", "_Not A Wavefile.wav"):
        assert L.afxh_extract_one(db.encode(), os.path.join(d, name).encode(), 1024, 0) == 0
    got, _, _ = dbcompare.rows(db)
    want, _, _ = dbcompare.rows(GOLDEN)
    for name in ("kick.wav", "_Not A Wavefile.wav"):
        errs = dbcompare.compare_row(got[name], want[name])
        assert not errs, "\n".join(errs[:20])


def test_long_file_in_parts_writes_the_same_row(tmp_path):
    """TGpuSampleAnalyser::AnalyzeInParts (the long-file path of the C++ adapter, BASELINE config 5): the row it writes
    equals the row of the whole-file Extract() -- every BLOB byte for byte."""
    import ctypes as C
    import numpy as np
    from afec_b200 import synth
    from oracle import oracle
    L = C.CDLL(afx_build.HOST_LIB)
    L.afxh_extract_one.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    L.afxh_extract_one_in_parts.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]
    d = str(tmp_path)
    clip = synth.one_shot(41, 6.0, rate=96000, channels=2)
    wav = os.path.join(d, "long_96k_stereo.wav")
    oracle.write_wav(wav, np.ascontiguousarray(np.tile(clip, (4, 1))), 96000)
    a, b = os.path.join(d, "whole.db"), os.path.join(d, "parts.db")
    assert L.afxh_extract_one(a.encode(), wav.encode(), 1024, 0) == 0
    dev = (C.c_int * 1)(0)
    assert L.afxh_extract_one_in_parts(b.encode(), wav.encode(), 1024, dev, 1, 3) == 0
    got, _, _ = dbcompare.rows(b)
    want, _, _ = dbcompare.rows(a)
    g, w = got["long_96k_stereo.wav"], want["long_96k_stereo.wav"]
    assert w["status"] == "succeeded" and g["status"] == "succeeded"
    for k in w:
        if k == "modtime":
            continue
        assert g[k] == w[k], k


def test_sample_formats_through_the_crawler_vs_the_reference(tmp_path):
    """WAV (8 / 16 / 24 / 32-bit, float 32 / 64, extensible) and AIFF / AIFC files: afec-b200-crawler (raw bytes to the GPU,
    sample conversion on the device) against the database the unmodified reference writes for the same files."""
    from oracle import oracle
    if not oracle.have_reference():
        pytest.skip("oracle/_ref not built")
    import audio_files
    from afec_b200 import synth
    d = str(tmp_path / "lib")
    os.makedirs(d)
    pcm = synth.one_shot(64, 0.5, rate=48000, channels=2)
    paths = []
    for name, (writer, kind, kw, _) in sorted(audio_files.format_cases().items()):
        p = os.path.join(d, name)
        writer(p, audio_files.quantise(pcm, kind), kind, 48000, **kw)
        paths.append(p)
    bad = os.path.join(d, "in24.aifc")
    audio_files.write_aiff(bad, audio_files.quantise(pcm, "i24"), "i24", 48000, compression="in24")
    ref_db = str(tmp_path / "ref.db")
    subprocess.run([oracle.REF_BIN, "db", "1024", ref_db] + paths + [bad], check=True, env=dict(os.environ, HOME=str(tmp_path)),
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    db = os.path.join(d, "afec-ll.db")
    run_crawler(["-o", db, d])
    got, _, _ = dbcompare.rows(db)
    want, _, _ = dbcompare.rows(ref_db)
    assert set(got) == set(want)
    for name in want:
        errs = dbcompare.compare_row(got[name], want[name])
        assert not errs, name + ":\n" + "\n".join(errs[:20])
    assert want["in24.aifc"]["status"].startswith("error: Sample failed to load")


def test_extract_batch_on_two_devices(tmp_path):
    """TGpuSampleAnalyser over devices {0, 1}: the decode -> GPU -> sink pipeline with slots on both GPUs writes the same
    rows as one device (needs two visible GPUs)."""
    import ctypes as C
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from afec_b200 import synth
    from oracle import oracle
    L = C.CDLL(afx_build.HOST_LIB)
    L.afxh_extract_files.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int,
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    d = str(tmp_path)
    names = []
    for i in range(48):
        p = os.path.join(d, "f%03d.wav" % i)
        oracle.write_wav(p, synth.one_shot(2000 + i, 0.2 + 0.05 * i, channels=1 + i % 2), 44100)
        names.append(p.encode())
    arr = (C.c_char_p * len(names))(*names)
    out = {}
    for tag, devs in (("one", [0]), ("two", [0, 1])):
        db = os.path.join(d, tag + ".db")
        dv = (C.c_int * len(devs))(*devs)
        os.environ["AFXH_MAX_BATCH_FILES"] = "5"                      # many chunks: both devices get work
        failed = L.afxh_extract_files(db.encode(), b"", arr, len(names), 1024, dv, len(devs), 2, None, None, None)
        del os.environ["AFXH_MAX_BATCH_FILES"]
        assert failed == 0
        out[tag] = dbcompare.rows(db)[0]
    assert set(out["one"]) == set(out["two"]) and len(out["one"]) == 48
    for name, w in out["one"].items():
        g = out["two"][name]
        for k in w:
            if k != "modtime":
                assert g[k] == w[k], (name, k)

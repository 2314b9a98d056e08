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


def test_crawler_database_is_a_drop_in(tmp_path):
    d = str(tmp_path)
    os.makedirs(os.path.join(d, "sub"))
    paths = db_cases.write_files(d)
    shutil.move(paths[1], os.path.join(d, "sub", os.path.basename(paths[1])))
    db = os.path.join(d, "afec-ll.db")
    out = run_crawler(["-l", "low", "-o", db, d])
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

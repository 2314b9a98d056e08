"""The host-side sink (afec_b200/host: TSqliteSampleDescriptorPool, msgpack packing) against a database
written by the unmodified reference (tests/golden/ref_ll.db).  No GPU needed: rows are fed from the CPU
oracle's values through afxh_write_row."""
import ctypes as C
import os
import sqlite3

import msgpack
import numpy as np
import pytest

import db_cases
import dbcompare
from afec_b200 import build as afx_build

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ll.db")


@pytest.fixture(scope="module")
def host():
    afx_build.build()
    L = C.CDLL(afx_build.HOST_LIB)
    L.afxh_schema.argtypes = [C.c_char_p, C.c_int]
    L.afxh_write_row.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_char_p]
    return L


def test_schema_is_the_reference_schema(host):
    n = host.afxh_schema(None, 0)
    buf = C.create_string_buffer(n + 1)
    host.afxh_schema(buf, n + 1)
    ours = "CREATE TABLE assets(" + buf.value.decode() + ")"
    _, ref_sql, _ = dbcompare.rows(GOLDEN)
    assert ours == ref_sql
    assert ours.count(",") + 1 == 461


def write_oracle_row(host, oracle_lib, db, directory, name, pcm, rate):
    path = os.path.join(directory, name)
    r = oracle_lib.analyze(pcm, src_rate=rate, file_size=os.path.getsize(path))
    fs = np.concatenate([np.ravel(a) for a in r.fs]) if r.fs else np.zeros(0)
    fv = np.concatenate([np.ravel(a) for a in r.fv]) if r.fv else np.zeros(0)
    hdr = np.ascontiguousarray(r.header, dtype=np.float64)
    st = np.ascontiguousarray(r.stats, dtype=np.float64)
    rc = host.afxh_write_row(db.encode(), (directory + "/").encode(), path.encode(), b"wav", r.F, r.Fr,
                             hdr.ctypes.data, fs.ctypes.data, fv.ctypes.data, st.ctypes.data, None)
    assert rc == 0


def test_rows_match_reference_database(host, oracle_lib, tmp_path):
    d = str(tmp_path)
    db_cases.write_files(d)
    db = os.path.join(d, "afec-ll.db")
    for name, c in db_cases.cases().items():
        if isinstance(c, bytes):
            rc = host.afxh_write_row(db.encode(), (d + "/").encode(), os.path.join(d, name).encode(), None, 0, 0,
                                     None, None, None, None, b"Sample failed to load: Not a valid WAV file.")
            assert rc == 0
        else:
            write_oracle_row(host, oracle_lib, db, d, name, c[0], c[1])
    got, sql, pragmas = dbcompare.rows(db)
    want, ref_sql, _ = dbcompare.rows(GOLDEN)
    assert sql == ref_sql
    assert pragmas == {"user_version": 2, "encoding": "UTF-8", "journal_mode": "wal"}
    assert set(got) == set(want)
    for name in want:
        errs = dbcompare.compare_row(got[name], want[name])
        assert not errs, name + ":\n" + "\n".join(errs[:20])
    # file names are stored relative to the base path, modtime is the file's mtime
    c = sqlite3.connect(db)
    names = sorted(r[0] for r in c.execute("select filename from assets"))
    assert names == sorted(db_cases.cases())
    mt = c.execute("select modtime from assets where filename='kick.wav'").fetchone()[0]
    assert mt == int(os.stat(os.path.join(d, "kick.wav")).st_mtime)
    assert c.execute("select count(*) from assets where status!='succeeded'").fetchone()[0] == 1
    c.close()


def test_msgpack_blobs_are_bit_exact(host, oracle_lib, tmp_path):
    """VR / VVR BLOBs decode to exactly the doubles that went in; sizes follow msgpack-c's array16 / fixarray rules."""
    d = str(tmp_path)
    db_cases.write_files(d)
    db = os.path.join(d, "x.db")
    pcm, rate = db_cases.cases()["kick.wav"]
    write_oracle_row(host, oracle_lib, db, d, "kick.wav", pcm, rate)
    r = oracle_lib.analyze(pcm, src_rate=rate, file_size=os.path.getsize(os.path.join(d, "kick.wav")))
    c = sqlite3.connect(db)
    blob, vv, vstat = c.execute("select spectral_centroid_VR, frequency_bands_VVR, cepstrum_bands_mean_VR from assets").fetchone()
    assert len(blob) == 3 + 9 * r.F and blob[0] == 0xdc
    assert np.array_equal(np.array(msgpack.unpackb(blob)), r.series("spectral_centroid"))
    assert np.array_equal(np.array(msgpack.unpackb(vv)), r.series("frequency_bands"))
    assert len(vstat) == 1 + 9 * 14 and vstat[0] == 0x9e
    assert np.array_equal(np.array(msgpack.unpackb(vstat)), r.series_stats("cepstrum_bands")[:, 3])
    c.close()


def test_old_database_versions_are_rebuilt_and_newer_refused(host, tmp_path):
    """SqliteSampleDescriptorPool.cpp:1239-1300: user_version < 2 -> dropped and recreated, > 2 -> refused."""
    db = str(tmp_path / "old.db")
    c = sqlite3.connect(db)
    c.execute("create table assets(filename text primary key, junk integer)")
    c.execute("insert into assets values('x', 1)")
    c.execute("pragma user_version = 1")
    c.commit(); c.close()
    rc = host.afxh_write_row(db.encode(), b"", b"/nonexistent/f.wav", None, 0, 0, None, None, None, None, b"Sample failed to load: x")
    assert rc == 0
    c = sqlite3.connect(db)
    assert c.execute("pragma user_version").fetchone()[0] == 2
    assert [r[0] for r in c.execute("select filename from assets")] == ["/nonexistent/f.wav"]
    c.execute("pragma user_version = 3"); c.commit(); c.close()
    assert host.afxh_write_row(db.encode(), b"", b"/nonexistent/g.wav", None, 0, 0, None, None, None, None, b"x") == -2


def test_sink_throughput_hook_writes_readable_rows(host, tmp_path):
    """afxh_sink_bench (how fast can rows go into afec-ll.db) inserts synthetic rows through the same
    InsertSample path; the rows must read back with every BLOB a well-formed msgpack array of the declared length."""
    host.afxh_sink_bench.restype = C.c_double
    host.afxh_sink_bench.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    db = str(tmp_path / "sink.db")
    secs = host.afxh_sink_bench(db.encode(), 12, 128, 1030, 5)
    assert secs > 0.0
    con = sqlite3.connect(db)
    assert con.execute("SELECT count(*) FROM assets WHERE status = 'succeeded'").fetchone()[0] == 12
    row = con.execute("SELECT spectral_centroid_VR, rhythm_complex_onsets_VR, cepstrum_bands_VVR, frequency_bands_mean_VR "
                      "FROM assets LIMIT 1").fetchone()
    cen, ons, cep, fbm = (msgpack.unpackb(b) for b in row)
    assert len(cen) == 128 and len(ons) == 1030 and len(cep) == 128 and len(cep[0]) == 14 and len(fbm) == 28
    assert con.execute("PRAGMA user_version").fetchone()[0] == 2
    con.close()


def _dump_table(path):
    con = sqlite3.connect(path)
    rows = {r[0]: r for r in con.execute("SELECT * FROM assets")}
    pragmas = {k: con.execute("pragma " + k).fetchone()[0] for k in ("user_version", "encoding", "journal_mode")}
    con.close()
    return rows, pragmas


def test_packed_rows_bulk_load_and_shards_write_the_same_bytes(host, tmp_path):
    """The round-2 sink paths -- rows whose BLOBs arrive packed (as the GPU delivers them), journal-less bulk load of a
    fresh database, several shard writers + merge -- all end in a database whose every column of every row equals the
    row-by-row WAL path byte for byte, with the reference's pragmas in the file header."""
    host.afxh_sink_bench2.restype = C.c_double
    host.afxh_sink_bench2.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    host.afxh_merge_shards.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int]
    n, F, Fr = 23, 40, 330
    base = str(tmp_path / "plain.db")
    assert host.afxh_sink_bench2(base.encode(), n, F, Fr, 1, 0, 1, None) > 0          # one transaction per row, WAL, host packing
    want, pragmas = _dump_table(base)
    assert len(want) == n and pragmas == {"user_version": 2, "encoding": "UTF-8", "journal_mode": "wal"}
    for tag, mode, shards in (("packed", 2, 1), ("bulk", 1, 1), ("packed_bulk", 3, 1), ("sharded", 3, 3)):
        db = str(tmp_path / (tag + ".db"))
        assert host.afxh_sink_bench2(db.encode(), n, F, Fr, 7, mode, shards, None) > 0
        if shards > 1:
            names = [(db + ".%d" % k).encode() for k in range(1, shards)]
            parts = [len(_dump_table(x.decode())[0]) for x in names] + [len(_dump_table(db)[0])]
            assert sum(parts) == n and min(parts) > 0                                 # disjoint shards, each a valid afec-ll.db
            assert host.afxh_merge_shards(db.encode(), (C.c_char_p * len(names))(*names), len(names), 1) == n - parts[-1]
            assert not any(os.path.exists(x.decode()) for x in names)
        got, pr = _dump_table(db)
        assert pr == pragmas, tag
        assert got == want, tag


def test_bulk_load_only_touches_an_empty_database(host, tmp_path):
    host.afxh_sink_bench2.restype = C.c_double
    host.afxh_sink_bench2.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    db = str(tmp_path / "twice.db")
    assert host.afxh_sink_bench2(db.encode(), 5, 10, 80, 2, 3, 1, None) > 0
    # a second run into the now non-empty database: BeginBulkLoad declines, rows are replaced through the WAL path
    assert host.afxh_sink_bench2(db.encode(), 5, 10, 80, 2, 3, 1, None) > 0
    rows, pragmas = _dump_table(db)
    assert len(rows) == 5 and pragmas["journal_mode"] == "wal"


def test_page_size_option_changes_the_file_not_the_rows(host, tmp_path, monkeypatch):
    """AFX_SINK_PAGE_SIZE / `afec-b200-crawler --page-size` (a sink-throughput option, off by default: the reference writes
    sqlite's 4096-byte pages): a new database takes the page size, every column of every row stays the same bytes, the file
    passes sqlite's integrity check; a value that is not a power of two in 512 .. 65536 is ignored."""
    import sqlite3
    host.afxh_sink_bench2.restype = C.c_double
    host.afxh_sink_bench2.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    n, F, Fr = 9, 40, 330
    base = str(tmp_path / "plain.db")
    assert host.afxh_sink_bench2(base.encode(), n, F, Fr, 4, 3, 1, None) > 0
    want, pragmas = _dump_table(base)
    for value, expect in (("32768", 32768), ("65536", 65536), ("5000", 4096), ("junk", 4096)):
        monkeypatch.setenv("AFX_SINK_PAGE_SIZE", value)
        db = str(tmp_path / ("p%s.db" % value))
        assert host.afxh_sink_bench2(db.encode(), n, F, Fr, 4, 3, 1, None) > 0
        got, pr = _dump_table(db)
        assert got == want and pr == pragmas, value
        con = sqlite3.connect(db)
        assert con.execute("PRAGMA page_size").fetchone()[0] == expect, value
        assert con.execute("PRAGMA integrity_check").fetchone()[0] == "ok"
        con.close()


def _sink(host):
    host.afxh_sink_bench2.restype = C.c_double
    host.afxh_sink_bench2.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    return host.afxh_sink_bench2


@pytest.mark.parametrize("page_size", [None, "32768", "1024"])
@pytest.mark.parametrize("n,F", [(1, 3), (7, 40), (230, 12), (60, 260), (2500, 1)])
def test_direct_load_writes_the_database_sqlite_would(host, tmp_path, monkeypatch, n, F, page_size):
    """TSqliteSampleDescriptorPool::BeginDirectLoad (afec-b200-crawler --direct-load): the rows of a fresh database are written
    in sqlite's FILE FORMAT by this repo's writer (direct_db_writer.cpp: table leaves + overflow chains, the filename index,
    interior pages, header) instead of through sqlite.  The result must be the database sqlite would have written: every
    column of every row (succeeded and failed ones) equal to the WAL path's, the reference's pragmas, sqlite's own integrity
    check clean (it walks both b-trees, every overflow chain, the free list and the index / table correspondence), look-ups
    through the index, and sqlite must be able to go on editing the file.  Sizes: one leaf as root, several leaves, an index
    of more than one level (2500 rows), rows of ~400 KB; three page sizes."""
    bench = _sink(host)
    if page_size:
        monkeypatch.setenv("AFX_SINK_PAGE_SIZE", page_size)
    want_db, got_db = str(tmp_path / "sqlite.db"), str(tmp_path / "direct.db")
    assert bench(want_db.encode(), n, F, 8 * F, 16, 8 | 0, 1, None) > 0                 # row by row through sqlite, WAL, host packing
    assert bench(got_db.encode(), n, F, 8 * F, 16, 8 | 4 | 2, 1, None) > 0              # direct load, rows arrive packed
    want, pragmas = _dump_table(want_db)
    got, pr = _dump_table(got_db)
    assert len(got) == n + (n + 2) // 3 and got == want and pr == pragmas
    con = sqlite3.connect(got_db)
    assert con.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
    assert con.execute("PRAGMA page_size").fetchone()[0] == int(page_size or 4096)
    for i in (0, n // 2, n - 1):
        name = "/nonexistent/f%07d.wav" % i
        plan = " ".join(r[-1] for r in con.execute("EXPLAIN QUERY PLAN SELECT status FROM assets WHERE filename = ?", (name,)))
        assert "sqlite_autoindex_assets_1" in plan
        assert con.execute("SELECT status FROM assets WHERE filename = ?", (name,)).fetchall() == [("succeeded",)]
    assert con.execute("SELECT status FROM assets WHERE filename = '/nonexistent/bad0000000.wav'").fetchall() == [("error: Sample failed to load: test",)]
    # sqlite goes on with the file: delete, replace, insert
    con.execute("DELETE FROM assets WHERE filename = '/nonexistent/f0000000.wav'")
    con.execute("INSERT OR REPLACE INTO assets(filename, modtime, status) VALUES ('/nonexistent/bad0000000.wav', 7, 'error: again')")
    con.execute("INSERT INTO assets(filename, modtime, status) VALUES ('zzz', 7, 'error: new')")
    con.commit()
    assert con.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
    assert con.execute("SELECT count(*) FROM assets").fetchone()[0] == len(got)
    con.close()


def test_direct_load_needs_an_empty_database_and_takes_a_name_twice(host, tmp_path):
    bench = _sink(host)
    db = str(tmp_path / "twice.db")
    assert bench(db.encode(), 4, 10, 80, 2, 4 | 2, 1, None) > 0
    assert bench(db.encode(), 4, 10, 80, 2, 4 | 2, 1, None) == -3.0           # not empty any more: BeginDirectLoad declines
    rows, _ = _dump_table(db)
    assert len(rows) == 4
    # the same file name twice in one load: the direct writer has no b-tree to replace a row in, so the pool keeps the later row
    # aside and puts it through sqlite's INSERT OR REPLACE when the load ends -- the database is the one sqlite alone writes
    db2, db3 = str(tmp_path / "dup.db"), str(tmp_path / "dup_sqlite.db")
    assert bench(db2.encode(), 4, 10, 80, 2, 16 | 4 | 2, 1, None) > 0
    assert bench(db3.encode(), 4, 10, 80, 2, 16 | 2, 1, None) > 0
    rows, pr = _dump_table(db2)
    want, pragmas = _dump_table(db3)
    assert rows == want and pr == pragmas
    assert len(rows) == 4 and sum(r[2] == "error: again" for r in rows.values()) == 1
    con = sqlite3.connect(db2)
    assert con.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
    con.close()


def test_direct_load_shards_and_merge(host, tmp_path):
    """Direct loads of several shard files side by side, merged through sqlite afterwards (ATTACH + INSERT ... SELECT reads the
    directly written files): the rows are those of one WAL-path database."""
    bench = _sink(host)
    host.afxh_merge_shards.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int]
    n, F = 41, 30
    base = str(tmp_path / "plain.db")
    assert bench(base.encode(), n, F, 8 * F, 1, 0, 1, None) > 0
    db = str(tmp_path / "sharded.db")
    assert bench(db.encode(), n, F, 8 * F, 5, 4 | 2, 3, None) > 0
    names = [(db + ".%d" % k).encode() for k in (1, 2)]
    assert host.afxh_merge_shards(db.encode(), (C.c_char_p * 2)(*names), 2, 1) > 0
    got, pr = _dump_table(db)
    want, pragmas = _dump_table(base)
    assert got == want and pr == pragmas


def test_direct_load_after_a_load_that_was_cut_off(host, tmp_path):
    """The header (page count) is written last: a direct load that dies leaves an EMPTY database with a tail of pages that are
    not part of it.  The next load writes over that tail."""
    bench = _sink(host)
    db = str(tmp_path / "cut.db")
    assert bench(db.encode(), 0, 5, 40, 2, 2, 1, None) >= 0                   # schema only
    con = sqlite3.connect(db); con.execute("PRAGMA wal_checkpoint(TRUNCATE)"); con.close()
    with open(db, "ab") as f:
        f.write(os.urandom(4096 * 37 + 100))
    con = sqlite3.connect(db)
    assert con.execute("SELECT count(*) FROM assets").fetchone()[0] == 0 and con.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
    con.close()
    assert bench(db.encode(), 9, 20, 160, 2, 4 | 2, 1, None) > 0
    con = sqlite3.connect(db)
    assert con.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
    assert con.execute("SELECT count(*) FROM assets").fetchone()[0] == 9
    con.close()
    assert os.path.getsize(db) % 4096 == 0


@pytest.mark.parametrize("page_size,name_len,n", [(None, 60, 900), (None, 1500, 400), (None, 3900, 60), ("1024", 400, 700), ("512", 200, 1200)])
def test_direct_load_index_with_long_names(host, tmp_path, monkeypatch, page_size, name_len, n):
    """File names longer than an index page's local limit (1002 bytes at 4096-byte pages, 230 at 1024, 102 at 512) make the index
    records spill into overflow pages, in leaves and -- for the keys that move up -- in interior pages; names arrive unsorted.
    The prime 7919 permutes the numbering, so n must not be a multiple of it."""
    if page_size:
        monkeypatch.setenv("AFX_SINK_PAGE_SIZE", page_size)
    host.afxh_direct_failed_rows.argtypes = [C.c_char_p, C.c_int, C.c_int]
    db = str(tmp_path / "names.db")
    assert host.afxh_direct_failed_rows(db.encode(), n, name_len) == 0
    con = sqlite3.connect(db)
    assert con.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
    assert con.execute("SELECT count(*), count(DISTINCT filename), min(length(filename)), max(length(filename)) FROM assets").fetchone() == (n, n, max(24, name_len), max(24, name_len))
    names = [r[0] for r in con.execute("SELECT filename FROM assets ORDER BY filename")]          # walks the index
    assert names == sorted(names)
    for probe in (names[0], names[n // 2], names[-1]):
        assert con.execute("SELECT status FROM assets WHERE filename = ?", (probe,)).fetchall() == [("error: Sample failed to load: test",)]
    con.execute("DELETE FROM assets WHERE filename = ?", (names[1],))
    con.execute("INSERT INTO assets(filename, modtime, status) VALUES (?, 1, 'error: new')", (names[1] + "x",))
    con.commit()
    assert con.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
    con.close()


def test_direct_load_under_sanitizers(tmp_path):
    """The direct writer handles raw page buffers: the same load (packed-on-host rows, failed rows, a name twice, long names, three
    page sizes) under AddressSanitizer + UndefinedBehaviorSanitizer must report nothing and leave a database sqlite accepts."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    hostdir = os.path.join(os.path.dirname(here), "afec_b200", "host")
    exe = str(tmp_path / "harness")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           "-I" + hostdir, "-I" + os.path.join(os.path.dirname(here), "include"), os.path.join(here, "direct_load_harness.cpp"),
           os.path.join(hostdir, "sqlite_pool.cpp"), os.path.join(hostdir, "direct_db_writer.cpp"), os.path.join(hostdir, "descriptors.cpp"),
           afx_build.SQLITE, "-pthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("this g++ has no sanitizer runtime")
    assert r.returncode == 0, r.stderr[-2000:]
    for page_size, rows, frames, name_len in (("4096", 300, 40, 60), ("4096", 40, 300, 2000), ("512", 200, 10, 700), ("65536", 150, 30, 24)):
        db = str(tmp_path / ("s%s_%d.db" % (page_size, rows)))
        env = dict(os.environ, AFX_SINK_PAGE_SIZE=page_size, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0")
        r = subprocess.run([exe, db, str(rows), str(frames), str(name_len)], capture_output=True, text=True, env=env)
        assert r.returncode == 0 and "ERROR" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
        con = sqlite3.connect(db)
        assert con.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
        assert con.execute("SELECT count(*), sum(status = 'error: twice') FROM assets").fetchone() == (rows, 1)
        con.close()

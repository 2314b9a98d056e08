"""Generates tests/golden/ref_ll.db: an afec-ll.db written by the UNMODIFIED reference
(oracle/_ref/afec_ref `db` mode = TSampleAnalyser::Extract into TSqliteSampleDescriptorPool) for the
files of tests/db_cases.py.  Run in the build container only:

    python tests/golden/make_golden_db.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import db_cases  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    assert oracle.have_reference(), "build oracle/_ref first (oracle/build_ref.sh)"
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_ll.db")
    with tempfile.TemporaryDirectory() as d:
        paths = db_cases.write_files(d)
        db = os.path.join(d, "ll.db")
        subprocess.run([oracle.REF_BIN, "db", "1024", db] + paths, check=True, env=dict(os.environ, HOME=d),
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        import sqlite3
        c = sqlite3.connect(db)
        c.execute("PRAGMA wal_checkpoint(TRUNCATE)")
        # file names are machine specific: keep the base name only
        for (name,) in c.execute("select filename from assets").fetchall():
            c.execute("update assets set filename=?, modtime=0 where filename=?", (os.path.basename(name), name))
        c.commit()
        c.execute("PRAGMA journal_mode=DELETE")
        c.close()
        shutil.copy(db, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

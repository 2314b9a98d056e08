// afec-ll.db writer: the reference's TSqliteSampleDescriptorPool for the low-level descriptor set.
// Reference: SqliteSampleDescriptorPool.cpp:1116-1350 (open / schema / version handling), :1551-1578
// (modification dates), :1582-1651 (InsertSample), :1655-1685 (InsertFailedSample), :1696-1733 (remove);
// Database.cpp:337-351 (pragmas).  One behavioural extension: BeginBulk / EndBulk wrap many inserts in
// one transaction (the reference commits one per file); rows and bytes are identical.
#include "afx_host.h"
#include "sqlite3_min.h"
#include "direct_db_writer.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <sys/stat.h>
#include <unistd.h>

namespace afec {

struct TSqliteSampleDescriptorPool::Impl {
  sqlite3* db = nullptr;
  sqlite3_stmt* insert = nullptr;
  sqlite3_stmt* insert_failed = nullptr;
  std::string file;
  int bulk = 0;
  bool bulk_load = false;
  std::vector<unsigned char> blob;    // one row's msgpack blobs, back to back (bound SQLITE_STATIC until the step)
  std::unique_ptr<TDirectDbWriter> direct;   // BeginDirectLoad .. EndBulkLoad: rows go straight into the file (db is closed meanwhile)
  std::vector<TDbValue> values;
  // rows of a direct load whose file name came before in the same load: kept (deep copies) and put through sqlite's INSERT OR
  // REPLACE when the load ends -- the direct writer has no b-tree to replace a row in
  struct Deferred { std::vector<TDbValue> values; std::vector<std::string> store; };
  std::vector<Deferred> deferred;
  void add_direct(const std::string& key);
};

void TSqliteSampleDescriptorPool::Impl::add_direct(const std::string& key)
{
  if (!direct->HasKey(key)) { direct->AddRow(values); return; }
  deferred.emplace_back();
  Deferred& d = deferred.back();
  d.values = values;
  d.store.reserve(values.size());
  for (TDbValue& v : d.values)
    if (v.mKind == TDbValue::kText || v.mKind == TDbValue::kBlob) { d.store.emplace_back((const char*)v.mData, v.mSize); v.mData = d.store.back().data(); }
}

static void check(sqlite3* db, int rc, const char* what)
{
  if (rc != SQLITE_OK && rc != SQLITE_DONE && rc != SQLITE_ROW)
    throw TReadableException(std::string("Database error (") + what + "): " + (db ? sqlite3_errmsg(db) : "?"));
}
static void exec(sqlite3* db, const std::string& sql)
{
  char* err = nullptr;
  const int rc = sqlite3_exec(db, sql.c_str(), nullptr, nullptr, &err);
  if (rc != SQLITE_OK) { std::string m = err ? err : "?"; sqlite3_free(err); throw TReadableException("Database error: " + m + " in '" + sql.substr(0, 80) + "'"); }
}
static long long scalar_int(sqlite3* db, const std::string& sql)
{
  sqlite3_stmt* st = nullptr;
  check(db, sqlite3_prepare_v2(db, sql.c_str(), -1, &st, nullptr), "prepare");
  long long v = 0;
  if (sqlite3_step(st) == SQLITE_ROW) v = sqlite3_column_int64(st, 0);
  sqlite3_finalize(st);
  return v;
}

std::vector<std::string> TSqliteSampleDescriptorPool::ColumnNamesAndTypes()
{
  // Descriptors(kLowLevelDescriptors) x Values(): "<name>_<R|S|VR|VVR> <type>" (SqliteSampleDescriptorPool.cpp:1313-1350)
  std::vector<std::string> c = { "filename TEXT PRIMARY KEY", "modtime INTEGER", "status TEXT", "file_type_S TEXT" };
  static const char* const kHeaderTypes[9] = { "INTEGER", "REAL", "INTEGER", "INTEGER", "INTEGER", "REAL", "REAL", "REAL", "REAL" };
  for (int i = 0; i < 9; ++i) c.push_back(std::string(kHeaderNames[i]) + "_R " + kHeaderTypes[i]);
  auto framed = [&](int s) {
    c.push_back(std::string(kFramedScalarNames[s]) + "_VR BLOB");
    for (int k = 0; k < AFX_N_STATS; ++k) c.push_back(std::string(kFramedScalarNames[s]) + "_" + kStatNames[k] + "_R REAL");
  };
  for (int s = 0; s < AFX_N_FS_MAIN; ++s) framed(s);
  for (int t = 0; t < 2; ++t) {                       // onsets + the six scalars of each onset type
    framed(AFX_N_FS_MAIN + t);
    for (int k = 0; k < 6; ++k) c.push_back(std::string(kHeaderNames[9 + 6 * t + k]) + "_R REAL");
  }
  c.push_back("rhythm_final_tempo_R REAL"); c.push_back("rhythm_final_tempo_confidence_R REAL");
  for (int v = 0; v < AFX_N_FV; ++v) {
    c.push_back(std::string(kFramedVectorNames[v]) + "_VVR BLOB");
    for (int k = 0; k < AFX_N_STATS; ++k) c.push_back(std::string(kFramedVectorNames[v]) + "_" + kStatNames[k] + "_VR BLOB");
  }
  return c;
}

TSqliteSampleDescriptorPool::TSqliteSampleDescriptorPool() : mImpl(new Impl()) {}
TSqliteSampleDescriptorPool::~TSqliteSampleDescriptorPool() { Close(); }

void TSqliteSampleDescriptorPool::Close()
{
  Impl& I = *mImpl;
  if (I.direct) { try { EndBulkLoad(); } catch (...) { I.direct.reset(); I.bulk_load = false; } }
  if (!I.db) return;
  if (I.bulk) { try { exec(I.db, "COMMIT"); } catch (...) {} I.bulk = 0; }
  if (I.bulk_load) { try { EndBulkLoad(); } catch (...) {} }
  if (I.insert) sqlite3_finalize(I.insert);
  if (I.insert_failed) sqlite3_finalize(I.insert_failed);
  I.insert = I.insert_failed = nullptr;
  sqlite3_close(I.db);
  I.db = nullptr;
}

static bool open_db(sqlite3** db, const std::string& name, bool ro)
{
  const int flags = ro ? SQLITE_OPEN_READONLY : (SQLITE_OPEN_READWRITE | SQLITE_OPEN_CREATE);
  if (sqlite3_open_v2(name.c_str(), db, flags, nullptr) != SQLITE_OK) { if (*db) sqlite3_close(*db); *db = nullptr; return false; }
  sqlite3_busy_timeout(*db, 60000);
  if (!ro) {                                            // Database.cpp:337-351
    exec(*db, "PRAGMA encoding = utf8;");
    // AFX_SINK_PAGE_SIZE / `--page-size` (a power of two, 512 .. 65536; takes effect on a database that does not exist yet): a row
    // is 50-250 overflow pages of the default 4096 bytes, each its own pwrite at commit; 32 KB pages measured +35 % rows/s
    // (DESIGN.md section 6).  Not the default: the reference writes sqlite's default page size.
    if (const char* ps = getenv("AFX_SINK_PAGE_SIZE")) {
      const int n = atoi(ps);
      if (n >= 512 && n <= 65536 && (n & (n - 1)) == 0) exec(*db, "PRAGMA page_size = " + std::to_string(n) + ";");
    }
    exec(*db, "PRAGMA journal_mode = WAL;");
    exec(*db, "PRAGMA synchronous = NORMAL;");
    // connection-local tuning (nothing of it is stored in the file): a row is 0.2-1 MB of BLOBs, so the default 2 MB
    // page cache and the 1000-page auto-checkpoint make every few inserts spill and checkpoint
    exec(*db, "PRAGMA cache_size = -262144;");
    exec(*db, "PRAGMA wal_autocheckpoint = 65536;");
  }
  return true;
}

bool TSqliteSampleDescriptorPool::Open(const std::string& DatabaseName, bool ReadOnly)
{
  Close();
  Impl& I = *mImpl;
  I.file = DatabaseName;
  if (!open_db(&I.db, DatabaseName, ReadOnly)) return false;
  if (ReadOnly) return true;
  // InitializeDatabase, SqliteSampleDescriptorPool.cpp:1224-1366
  bool create = false;
  if (scalar_int(I.db, "SELECT count(name) FROM sqlite_master WHERE type='table' AND name='assets'") != 1) create = true;
  else {
    const int version = (int)scalar_int(I.db, "PRAGMA user_version");
    if (version > kCurrentVersion)
      throw TReadableException("Unknown database version: " + std::to_string(version) + ". The database maybe got created by a newer version of the crawler.");
    if (version < kCurrentVersion) {                    // older layout: drop everything and start over
      create = true;
      sqlite3_close(I.db); I.db = nullptr;
      const bool deleted = (unlink(DatabaseName.c_str()) == 0);
      unlink((DatabaseName + "-wal").c_str()); unlink((DatabaseName + "-shm").c_str());
      if (!open_db(&I.db, DatabaseName, false)) throw TReadableException("Failed to (re)open the database file for upgrading.");
      if (!deleted) { try { exec(I.db, "DROP table 'assets'"); exec(I.db, "VACUUM"); } catch (...) {} }
    }
  }
  if (create) {
    exec(I.db, "BEGIN");
    exec(I.db, "PRAGMA user_version = '" + std::to_string((int)kCurrentVersion) + "'");
    std::string sql = "CREATE TABLE assets(";
    const std::vector<std::string> cols = ColumnNamesAndTypes();
    for (size_t i = 0; i < cols.size(); ++i) { if (i) sql += ","; sql += cols[i]; }
    sql += ")";
    exec(I.db, sql);
    exec(I.db, "COMMIT");
  }
  return true;
}

void TSqliteSampleDescriptorPool::SetBasePath(const std::string& BasePath)
{
  mBasePath = BasePath;
  if (!mBasePath.empty() && mBasePath.back() != '/') mBasePath += '/';
}

std::string TSqliteSampleDescriptorPool::RelativeFilenamePath(const std::string& FileName) const
{
  if (!mBasePath.empty() && FileName.compare(0, mBasePath.size(), mBasePath) == 0) return FileName.substr(mBasePath.size());
  return FileName;
}

bool TSqliteSampleDescriptorPool::IsEmpty() const { return NumberOfSamples() == 0; }
int TSqliteSampleDescriptorPool::NumberOfSamples() const
{
  return mImpl->db ? (int)scalar_int(mImpl->db, "SELECT count(*) FROM assets") : 0;
}

std::vector<std::pair<std::string, int>> TSqliteSampleDescriptorPool::SampleModificationDates() const
{
  std::vector<std::pair<std::string, int>> ret;
  sqlite3_stmt* st = nullptr;
  check(mImpl->db, sqlite3_prepare_v2(mImpl->db, "SELECT filename, modtime FROM assets", -1, &st, nullptr), "prepare");
  while (sqlite3_step(st) == SQLITE_ROW) {
    std::string name = (const char*)sqlite3_column_text(st, 0);
    if (!mBasePath.empty() && name.compare(0, mBasePath.size(), mBasePath) != 0) name = mBasePath + name;
    ret.push_back({ name, sqlite3_column_int(st, 1) });
  }
  sqlite3_finalize(st);
  return ret;
}

void TSqliteSampleDescriptorPool::BeginBulk() { if (mImpl->direct) return; if (mImpl->bulk++ == 0) exec(mImpl->db, "BEGIN"); }
void TSqliteSampleDescriptorPool::EndBulk() { if (mImpl->direct) return; if (mImpl->bulk > 0 && --mImpl->bulk == 0) exec(mImpl->db, "COMMIT"); }

// append helpers: msgpack straight into the row buffer (same bytes as PackVR / PackVVR)
static inline void put_array_header(std::vector<unsigned char>& out, size_t n)
{
  if (n < 16) out.push_back((unsigned char)(0x90u | n));
  else if (n < 65536) { out.push_back(0xdc); out.push_back((unsigned char)(n >> 8)); out.push_back((unsigned char)n); }
  else { out.push_back(0xdd); out.push_back((unsigned char)(n >> 24)); out.push_back((unsigned char)(n >> 16)); out.push_back((unsigned char)(n >> 8)); out.push_back((unsigned char)n); }
}
static inline void put_doubles(std::vector<unsigned char>& out, const double* v, size_t n)
{
  const size_t o = out.size();
  out.resize(o + 9 * n);
  unsigned char* p = out.data() + o;
  for (size_t i = 0; i < n; ++i) {
    uint64_t u; memcpy(&u, v + i, 8);
    u = __builtin_bswap64(u);
    *p++ = 0xcb; memcpy(p, &u, 8); p += 8;
  }
}

// One row as values in column order (ColumnNamesAndTypes).  Texts and BLOBs are referenced, not copied: they stay valid until
// the row has been stepped / written -- a row packed on the GPU goes from the pinned download buffer into sqlite's pages (or,
// in a direct load, into the file's) with no copy in between.
static void row_values(std::vector<TDbValue>& v, const std::string& rel, int modtime, const std::string& file_type,
                       const double* header, const double (*stats)[AFX_N_STATS], const unsigned char* const* blob_ptr, const int* blob_len)
{
  static const std::string kSucceeded = "succeeded";
  v.clear();
  int blob = 0;
  v.push_back(TDbValue::Text(rel));
  v.push_back(TDbValue::Int(modtime));
  v.push_back(TDbValue::Text(kSucceeded));
  v.push_back(TDbValue::Text(file_type));
  v.push_back(TDbValue::Int((int)header[0]));        // file_size
  v.push_back(TDbValue::Real(header[1]));            // file_length
  v.push_back(TDbValue::Int((int)header[2]));
  v.push_back(TDbValue::Int((int)header[3]));
  v.push_back(TDbValue::Int((int)header[4]));
  for (int k = 5; k < 9; ++k) v.push_back(TDbValue::Real(header[k]));
  auto framed = [&](int s) {
    v.push_back(TDbValue::Blob(blob_ptr[blob], (size_t)blob_len[blob])); ++blob;
    for (int k = 0; k < AFX_N_STATS; ++k) v.push_back(TDbValue::Real(stats[s][k]));
  };
  for (int s = 0; s < AFX_N_FS_MAIN; ++s) framed(s);
  for (int t = 0; t < 2; ++t) {
    framed(AFX_N_FS_MAIN + t);
    for (int k = 0; k < 6; ++k) v.push_back(TDbValue::Real(header[9 + 6 * t + k]));
  }
  v.push_back(TDbValue::Real(header[21])); v.push_back(TDbValue::Real(header[22]));
  for (int f = 0; f < AFX_N_FV; ++f)
    for (int k = 0; k < 1 + AFX_N_STATS; ++k) { v.push_back(TDbValue::Blob(blob_ptr[blob], (size_t)blob_len[blob])); ++blob; }
}

static void step_values(sqlite3* db, sqlite3_stmt* st, const std::vector<TDbValue>& v)
{
  int p = 1;
  for (const TDbValue& x : v) {
    switch (x.mKind) {
      case TDbValue::kNull: sqlite3_bind_null(st, p); break;
      case TDbValue::kInt: sqlite3_bind_int64(st, p, x.mInt); break;
      case TDbValue::kReal: sqlite3_bind_double(st, p, x.mReal); break;
      case TDbValue::kText: sqlite3_bind_text(st, p, (const char*)x.mData, (int)x.mSize, SQLITE_STATIC); break;
      case TDbValue::kBlob: sqlite3_bind_blob(st, p, x.mData, (int)x.mSize, SQLITE_STATIC); break;
    }
    ++p;
  }
  const int rc = sqlite3_step(st);
  sqlite3_reset(st); sqlite3_clear_bindings(st);
  check(db, rc, "insert");
}

void TSqliteSampleDescriptorPool::PrepareInsert()
{
  Impl& I = *mImpl;
  if (!I.db) throw TReadableException("Database is not open");
  if (I.insert) return;
  const std::vector<std::string> cols = ColumnNamesAndTypes();
  std::string sql = "INSERT OR REPLACE into assets(", q;
  for (size_t i = 0; i < cols.size(); ++i) {
    if (i) { sql += ","; q += ","; }
    sql += cols[i].substr(0, cols[i].find(' ')); q += "?";
  }
  sql += ") values(" + q + ")";
  check(I.db, sqlite3_prepare_v2(I.db, sql.c_str(), -1, &I.insert, nullptr), "prepare insert");
}

void TSqliteSampleDescriptorPool::InsertRow(const std::string& FileName, const std::string& FileType, const double* Header,
                                            const double (*Stats)[AFX_N_STATS], const unsigned char* const* BlobPtr, const int* BlobLen)
{
  Impl& I = *mImpl;
  const std::string rel = RelativeFilenamePath(FileName);
  row_values(I.values, rel, ModificationStatTime(FileName), FileType, Header, Stats, BlobPtr, BlobLen);
  if (I.direct) { I.add_direct(rel); return; }
  PrepareInsert();
  const bool own_txn = (I.bulk == 0);
  if (own_txn) exec(I.db, "BEGIN");
  try {
    step_values(I.db, I.insert, I.values);
    if (own_txn) exec(I.db, "COMMIT");
  } catch (...) {
    sqlite3_reset(I.insert); sqlite3_clear_bindings(I.insert);
    if (own_txn) { try { exec(I.db, "ROLLBACK"); } catch (...) {} }
    throw;
  }
}

void TSqliteSampleDescriptorPool::InsertSample(const std::string& FileName, const TSampleDescriptors& R)
{
  Impl& I = *mImpl;
  // All BLOBs of the row are packed back to back into one buffer (the buffer may move while it grows, so the pointers
  // are taken at the end); column order as ColumnNamesAndTypes()
  I.blob.clear();
  size_t want = 64;
  for (int s = 0; s < AFX_N_FS; ++s) want += 5 + 9 * R.mFramedScalars[s].size();
  for (int v = 0; v < AFX_N_FV; ++v) want += 5 + (size_t)R.mFrames * (3 + 9 * (size_t)kFramedVectorBands[v]) + AFX_N_STATS * (5 + 9 * (size_t)kFramedVectorBands[v]);
  I.blob.reserve(want);
  size_t off[AFX_N_BLOBS + 1]; int nb_ = 0;
  for (int s = 0; s < AFX_N_FS; ++s) {
    off[nb_++] = I.blob.size();
    put_array_header(I.blob, R.mFramedScalars[s].size());
    put_doubles(I.blob, R.mFramedScalars[s].data(), R.mFramedScalars[s].size());
  }
  int series = AFX_N_FS;
  for (int v = 0; v < AFX_N_FV; ++v) {
    const int nb = kFramedVectorBands[v];
    off[nb_++] = I.blob.size();
    put_array_header(I.blob, (size_t)R.mFrames);
    for (size_t f = 0; f < (size_t)R.mFrames; ++f) { put_array_header(I.blob, (size_t)nb); put_doubles(I.blob, R.mFramedVectors[v].data() + f * nb, (size_t)nb); }
    double col[28];
    for (int k = 0; k < AFX_N_STATS; ++k) {
      for (int b = 0; b < nb; ++b) col[b] = R.mStats[series + b][k];
      off[nb_++] = I.blob.size();
      put_array_header(I.blob, (size_t)nb);
      put_doubles(I.blob, col, (size_t)nb);
    }
    series += nb;
  }
  off[nb_] = I.blob.size();
  const unsigned char* ptr[AFX_N_BLOBS]; int len[AFX_N_BLOBS];
  for (int k = 0; k < AFX_N_BLOBS; ++k) { ptr[k] = I.blob.data() + off[k]; len[k] = (int)(off[k + 1] - off[k]); }
  InsertRow(FileName, R.mFileType, R.mHeader, R.mStats, ptr, len);
}

void TSqliteSampleDescriptorPool::InsertPackedSample(const std::string& FileName, const std::string& FileType, const afx_file_result& R)
{
  if (!R.packed || !R.packed_off || !R.header || !R.stats) throw TReadableException("InsertPackedSample: the result holds no packed row (AFX_FEAT_PACK)");
  const unsigned char* ptr[AFX_N_BLOBS]; int len[AFX_N_BLOBS];
  for (int k = 0; k < AFX_N_BLOBS; ++k) { ptr[k] = R.packed + R.packed_off[k]; len[k] = (int)(R.packed_off[k + 1] - R.packed_off[k]); }
  InsertRow(FileName, FileType, R.header, reinterpret_cast<const double (*)[AFX_N_STATS]>(R.stats), ptr, len);
}

// A fresh database can be filled without a journal: nothing exists that a crash could damage (the crawl starts over), and
// sqlite then writes every page once instead of twice (WAL frame + checkpoint).  The file is switched to the reference's
// journal_mode = WAL / synchronous = NORMAL again when the load ends, so the finished afec-ll.db has the same header
// pragmas as one written row by row.
bool TSqliteSampleDescriptorPool::BeginBulkLoad()
{
  Impl& I = *mImpl;
  if (!I.db || I.bulk_load || I.bulk || NumberOfSamples() != 0) return false;
  exec(I.db, "PRAGMA journal_mode = OFF;");
  exec(I.db, "PRAGMA synchronous = OFF;");
  I.bulk_load = true;
  return true;
}
void TSqliteSampleDescriptorPool::EndBulkLoad()
{
  Impl& I = *mImpl;
  if (I.direct) {                                       // the file is complete: back to sqlite (open_db sets WAL / NORMAL)
    std::unique_ptr<TDirectDbWriter> w = std::move(I.direct);
    I.bulk_load = false;
    w->Finish();
    if (!open_db(&I.db, I.file, false)) throw TReadableException("Failed to reopen the database after the direct load");
    if (!I.deferred.empty()) {                          // names that came twice: the later row replaces the earlier one, as it would have
      std::vector<Impl::Deferred> rows;
      rows.swap(I.deferred);
      PrepareInsert();
      exec(I.db, "BEGIN");
      try { for (const Impl::Deferred& d : rows) step_values(I.db, I.insert, d.values); exec(I.db, "COMMIT"); }
      catch (...) { try { exec(I.db, "ROLLBACK"); } catch (...) {} throw; }
    }
    return;
  }
  if (!I.db || !I.bulk_load) return;
  if (I.bulk) { exec(I.db, "COMMIT"); I.bulk = 0; }
  exec(I.db, "PRAGMA journal_mode = WAL;");
  exec(I.db, "PRAGMA synchronous = NORMAL;");
  I.bulk_load = false;
}

// The same for an EMPTY database, without sqlite in the data path: the rows are written into the file in sqlite's format
// (TDirectDbWriter), sequentially, one copy per row.  Until EndBulkLoad / Close only the Insert* calls and BeginBulk /
// EndBulk (no-ops) may be used; a row whose file name came before in the same load is kept aside and goes through sqlite's
// INSERT OR REPLACE when the load ends.
bool TSqliteSampleDescriptorPool::BeginDirectLoad()
{
  Impl& I = *mImpl;
  if (!I.db || I.bulk_load || I.bulk || I.direct || NumberOfSamples() != 0) return false;
  if (scalar_int(I.db, "PRAGMA auto_vacuum") != 0) return false;
  if (scalar_int(I.db, "SELECT count(*) FROM sqlite_master WHERE tbl_name='assets'") != 2) return false;     // the table and its key index
  const long long troot = scalar_int(I.db, "SELECT rootpage FROM sqlite_master WHERE type='table' AND name='assets'");
  const long long iroot = scalar_int(I.db, "SELECT rootpage FROM sqlite_master WHERE type='index' AND tbl_name='assets'");
  if (troot < 2 || iroot < 2) return false;
  if (I.insert) sqlite3_finalize(I.insert);
  if (I.insert_failed) sqlite3_finalize(I.insert_failed);
  I.insert = I.insert_failed = nullptr;
  exec(I.db, "PRAGMA wal_checkpoint(TRUNCATE);");
  exec(I.db, "PRAGMA journal_mode = DELETE;");          // no log beside the file while it is written directly
  sqlite3_close(I.db); I.db = nullptr;
  try { I.direct.reset(new TDirectDbWriter(I.file, (uint32_t)troot, (uint32_t)iroot)); }
  catch (const std::exception&) {                       // not a file this writer takes: back to sqlite, the caller loads through it
    if (!open_db(&I.db, I.file, false)) { I.db = nullptr; throw; }
    return false;
  }
  I.bulk_load = true;
  return true;
}

// Rows of other afec-ll.db files (written side by side by several sink threads) are appended to this one
int TSqliteSampleDescriptorPool::MergeFrom(const std::vector<std::string>& ShardFiles, bool DeleteShards)
{
  Impl& I = *mImpl;
  if (!I.db) throw TReadableException("Database is not open");
  int merged = 0;
  for (size_t k = 0; k < ShardFiles.size(); ++k) {
    const std::string alias = "shard" + std::to_string(k);
    std::string quoted = ShardFiles[k]; for (size_t q = 0; (q = quoted.find('\'', q)) != std::string::npos; q += 2) quoted.insert(q, "'");
    exec(I.db, "ATTACH DATABASE '" + quoted + "' AS " + alias);
    exec(I.db, "BEGIN");
    exec(I.db, "INSERT OR REPLACE INTO main.assets SELECT * FROM " + alias + ".assets");
    merged += sqlite3_changes(I.db);
    exec(I.db, "COMMIT");
    exec(I.db, "DETACH DATABASE " + alias);
    if (DeleteShards) { unlink(ShardFiles[k].c_str()); unlink((ShardFiles[k] + "-wal").c_str()); unlink((ShardFiles[k] + "-shm").c_str()); }
  }
  return merged;
}

void TSqliteSampleDescriptorPool::InsertFailedSample(const std::string& FileName, const std::string& Reason)
{
  Impl& I = *mImpl;
  if (I.direct) {
    const std::string rel = RelativeFilenamePath(FileName), status = "error: " + Reason;
    static const size_t kColumns = ColumnNamesAndTypes().size();
    I.values.assign(kColumns, TDbValue::Null());
    I.values[0] = TDbValue::Text(rel); I.values[1] = TDbValue::Int(ModificationStatTime(FileName)); I.values[2] = TDbValue::Text(status);
    I.add_direct(rel);
    return;
  }
  if (!I.db) throw TReadableException("Database is not open");
  if (!I.insert_failed)
    check(I.db, sqlite3_prepare_v2(I.db, "INSERT OR REPLACE into assets(filename, modtime, status) values (?,?,?)", -1, &I.insert_failed, nullptr), "prepare");
  sqlite3_stmt* st = I.insert_failed;
  const std::string rel = RelativeFilenamePath(FileName), status = "error: " + Reason;
  const bool own_txn = (I.bulk == 0);
  if (own_txn) exec(I.db, "BEGIN");
  sqlite3_bind_text(st, 1, rel.c_str(), -1, SQLITE_TRANSIENT);
  sqlite3_bind_int(st, 2, ModificationStatTime(FileName));
  sqlite3_bind_text(st, 3, status.c_str(), -1, SQLITE_TRANSIENT);
  const int rc = sqlite3_step(st);
  sqlite3_reset(st); sqlite3_clear_bindings(st);
  if (own_txn) exec(I.db, (rc == SQLITE_DONE) ? "COMMIT" : "ROLLBACK");
  check(I.db, rc, "insert failed sample");
}

void TSqliteSampleDescriptorPool::RemoveSample(const std::string& FileName) { RemoveSamples(std::vector<std::string>(1, FileName)); }

void TSqliteSampleDescriptorPool::RemoveSamples(const std::vector<std::string>& FileNames)
{
  Impl& I = *mImpl;
  sqlite3_stmt* st = nullptr;
  // the reference matches with a case-folding MATCH operator; Linux file names are case sensitive: '='
  check(I.db, sqlite3_prepare_v2(I.db, "DELETE from assets WHERE filename = ?", -1, &st, nullptr), "prepare delete");
  size_t i = 0;
  while (i < FileNames.size()) {                          // batches of 100, :1713-1725
    exec(I.db, "BEGIN");
    for (int k = 0; k < 100 && i < FileNames.size(); ++k, ++i) {
      const std::string rel = RelativeFilenamePath(FileNames[i]);
      sqlite3_bind_text(st, 1, rel.c_str(), -1, SQLITE_TRANSIENT);
      sqlite3_step(st); sqlite3_reset(st);
    }
    exec(I.db, "COMMIT");
  }
  sqlite3_finalize(st);
}

}  // namespace afec

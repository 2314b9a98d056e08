// Small C entry points over the host classes, for language bindings and tests (ctypes).
#include "afx_host.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <algorithm>
#include <memory>
#include <thread>
#include <sys/stat.h>
#include <cstdlib>
#include <cstring>

using namespace afec;

extern "C" {

// number of columns and the CREATE TABLE column list ("name TYPE,name TYPE,...")
int afxh_schema(char* out, int cap)
{
  std::string s;
  const auto cols = TSqliteSampleDescriptorPool::ColumnNamesAndTypes();
  for (size_t i = 0; i < cols.size(); ++i) { if (i) s += ","; s += cols[i]; }
  if (out && cap > 0) { strncpy(out, s.c_str(), (size_t)cap - 1); out[cap - 1] = 0; }
  return (int)s.size();
}

// write one row from flat arrays laid out like afx_file_result (header[32], fs concatenated, fv concatenated,
// stats[136][13]) -- lets the sink be tested without a GPU.  status_reason != NULL writes a failed row.
int afxh_write_row(const char* db, const char* base_path, const char* filename, const char* file_type,
                   int n_frames, int n_rhythm_frames, const double* header, const double* fs, const double* fv,
                   const double* stats, const char* failed_reason)
{
  try {
    TSqliteSampleDescriptorPool pool;
    if (!pool.Open(db)) return -1;
    if (base_path && *base_path) pool.SetBasePath(base_path);
    if (failed_reason) { pool.InsertFailedSample(filename, failed_reason); return 0; }
    TSampleDescriptors d;
    d.mFileName = filename; d.mFileType = file_type ? file_type : "";
    d.mFrames = n_frames; d.mRhythmFrames = n_rhythm_frames;
    memcpy(d.mHeader, header, sizeof(d.mHeader));
    const double* p = fs;
    for (int s = 0; s < AFX_N_FS; ++s) { const int n = s < AFX_N_FS_MAIN ? n_frames : n_rhythm_frames; d.mFramedScalars[s].assign(p, p + n); p += n; }
    p = fv;
    for (int v = 0; v < AFX_N_FV; ++v) { const size_t n = (size_t)n_frames * kFramedVectorBands[v]; d.mFramedVectors[v].assign(p, p + n); p += n; }
    memcpy(d.mStats, stats, sizeof(d.mStats));
    pool.InsertSample(filename, d);
    return 0;
  } catch (const std::exception&) { return -2; }
}

// header probe of a WAV / AIFF file (no GPU): 0, or -1 with the loader's message in err
int afxh_probe_audio(const char* filename, long long* frames, int* channels, int* rate, int* bits, int* format, long long* data_bytes,
                     char* err, int errcap)
{
  try {
    TAudioInfo I;
    ProbeAudioFile(filename, I);
    if (frames) *frames = I.mFrames;
    if (channels) *channels = I.mChannels;
    if (rate) *rate = I.mSampleRate;
    if (bits) *bits = I.mBitDepth;
    if (format) *format = I.mFormat;
    if (data_bytes) *data_bytes = (long long)I.mDataBytes;
    return 0;
  } catch (const std::exception& e) {
    if (err && errcap > 0) { strncpy(err, e.what(), (size_t)errcap - 1); err[errcap - 1] = 0; }
    return -1;
  }
}

// the sample bytes of a WAV / AIFF file as the extractor uploads them (raw, format as probed); returns bytes written or < 0
long long afxh_read_audio(const char* filename, unsigned char* dst, long long cap)
{
  try {
    TAudioInfo I;
    ProbeAudioFile(filename, I);
    if ((long long)I.mDataBytes > cap) return -2;
    ReadAudioData(I, dst);
    return (long long)I.mDataBytes;
  } catch (const std::exception&) { return -1; }
}

// run the batched extractor over a list of files; returns failed count or < 0
int afxh_extract_files(const char* db, const char* base_path, const char* const* files, int n_files, int hop,
                       const int* devices, int n_devices, int slots_per_device, double* audio_seconds, double* seconds,
                       long long* main_frames)
{
  try {
    TSqliteSampleDescriptorPool pool;
    if (!pool.Open(db)) return -1;
    if (base_path && *base_path) pool.SetBasePath(base_path);
    std::vector<int> dev(devices, devices + n_devices);
    TGpuSampleAnalyser an(44100, 2048, hop, dev, slots_per_device);
    std::vector<std::string> names(files, files + n_files);
    if (getenv("AFXH_MAX_BATCH_FILES")) an.SetMaxBatchFiles(std::max(1, atoi(getenv("AFXH_MAX_BATCH_FILES"))));   // tests: many small chunks
    std::mutex lock; TGpuSampleAnalyser::TProgress pr;
    const int failed = an.ExtractBatch(names, &pool, lock, &pr);
    if (audio_seconds) *audio_seconds = pr.mAudioSeconds;
    if (seconds) *seconds = pr.mSeconds;
    if (main_frames) *main_frames = pr.mMainFrames;
    return failed;
  } catch (const std::exception&) { return -2; }
}

// TSampleAnalyser::Extract on one file (single-file entry point)
int afxh_extract_one(const char* db, const char* filename, int hop, int device)
{
  try {
    TSqliteSampleDescriptorPool pool;
    if (!pool.Open(db)) return -1;
    TGpuSampleAnalyser an(44100, 2048, hop, std::vector<int>(1, device), 1);
    std::mutex lock;
    an.Extract(filename, &pool, lock);
    return 0;
  } catch (const std::exception&) { return -2; }
}

// sink throughput: insert n_rows synthetic rows of a file with `frames` main and `rframes` rhythm
// frames, `bulk` rows per transaction; returns seconds (< 0 on error).  No GPU involved.
double afxh_sink_bench(const char* db, int n_rows, int frames, int rframes, int bulk)
{
  try {
    TSqliteSampleDescriptorPool pool;
    if (!pool.Open(db)) return -1.0;
    TSampleDescriptors d;
    d.mFileType = "wav"; d.mFrames = frames; d.mRhythmFrames = rframes;
    for (int s = 0; s < AFX_N_FS; ++s) { const int n = s < AFX_N_FS_MAIN ? frames : rframes; d.mFramedScalars[s].resize(n); for (int i = 0; i < n; ++i) d.mFramedScalars[s][i] = 0.001 * i + s; }
    for (int v = 0; v < AFX_N_FV; ++v) { const size_t n = (size_t)frames * kFramedVectorBands[v]; d.mFramedVectors[v].resize(n); for (size_t i = 0; i < n; ++i) d.mFramedVectors[v][i] = 1e-3 * (double)i; }
    for (int s = 0; s < AFX_N_SERIES; ++s) for (int k = 0; k < AFX_N_STATS; ++k) d.mStats[s][k] = s + 0.01 * k;
    const auto t0 = std::chrono::steady_clock::now();
    char name[64];
    for (int i = 0; i < n_rows; ++i) {
      if (bulk > 1 && i % bulk == 0) pool.BeginBulk();
      snprintf(name, sizeof(name), "/nonexistent/f%07d.wav", i);
      d.mFileName = name;
      pool.InsertSample(name, d);
      if (bulk > 1 && (i % bulk == bulk - 1 || i == n_rows - 1)) pool.EndBulk();
    }
    pool.Close();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  } catch (const std::exception&) { return -2.0; }
}

// sink throughput with the round-2 paths: `mode` bit 0 = journal-less bulk load of the fresh database, bit 1 = rows arrive
// packed (as afx_file_result.packed delivers them from the GPU; packed once here, outside the timed region), bit 2 = direct load
// (TSqliteSampleDescriptorPool::BeginDirectLoad: the file is written in sqlite's format without sqlite), bit 3 = a failed row after
// every third row, bit 4 = the first name once more at the end (-4 when it is refused); `shards` pools written by as many threads.  Returns seconds (< 0 on error); *bytes = database bytes written.
double afxh_sink_bench2(const char* db, int n_rows, int frames, int rframes, int bulk, int mode, int shards, long long* bytes)
{
  try {
    TSampleDescriptors d;
    d.mFileType = "wav"; d.mFrames = frames; d.mRhythmFrames = rframes;
    for (int s = 0; s < AFX_N_FS; ++s) { const int n = s < AFX_N_FS_MAIN ? frames : rframes; d.mFramedScalars[s].resize(n); for (int i = 0; i < n; ++i) d.mFramedScalars[s][i] = 0.001 * i + s; }
    for (int v = 0; v < AFX_N_FV; ++v) { const size_t n = (size_t)frames * kFramedVectorBands[v]; d.mFramedVectors[v].resize(n); for (size_t i = 0; i < n; ++i) d.mFramedVectors[v][i] = 1e-3 * (double)i; }
    for (int s = 0; s < AFX_N_SERIES; ++s) for (int k = 0; k < AFX_N_STATS; ++k) d.mStats[s][k] = s + 0.01 * k;
    // the packed image of that row
    std::vector<unsigned char> packed, one; std::vector<uint32_t> off;
    for (int s = 0; s < AFX_N_FS; ++s) { off.push_back((uint32_t)packed.size()); PackVR(one, d.mFramedScalars[s].data(), d.mFramedScalars[s].size()); packed.insert(packed.end(), one.begin(), one.end()); }
    int series = AFX_N_FS;
    for (int v = 0; v < AFX_N_FV; ++v) {
      const int nb = kFramedVectorBands[v];
      off.push_back((uint32_t)packed.size()); PackVVR(one, d.mFramedVectors[v].data(), (size_t)frames, (size_t)nb); packed.insert(packed.end(), one.begin(), one.end());
      for (int k = 0; k < AFX_N_STATS; ++k) {
        double col[28]; for (int b = 0; b < nb; ++b) col[b] = d.mStats[series + b][k];
        off.push_back((uint32_t)packed.size()); PackVR(one, col, (size_t)nb); packed.insert(packed.end(), one.begin(), one.end());
      }
      series += nb;
    }
    off.push_back((uint32_t)packed.size());
    afx_file_result r; memset(&r, 0, sizeof(r));
    r.n_frames = frames; r.n_rhythm_frames = rframes; r.header = d.mHeader; r.stats = &d.mStats[0][0];
    r.packed = packed.data(); r.packed_off = off.data();

    std::vector<std::unique_ptr<TSqliteSampleDescriptorPool>> pools;
    std::vector<std::string> names;
    for (int k = 0; k < std::max(1, shards); ++k) {
      names.push_back(k == 0 ? std::string(db) : std::string(db) + "." + std::to_string(k));
      pools.emplace_back(new TSqliteSampleDescriptorPool());
      if (!pools.back()->Open(names.back())) return -1.0;
      if (mode & 4) { if (!pools.back()->BeginDirectLoad()) return -3.0; }
      else if (mode & 1) pools.back()->BeginBulkLoad();
    }
    const auto t0 = std::chrono::steady_clock::now();
    std::atomic<bool> dup_accepted(false);
    std::vector<std::thread> th;
    for (size_t k = 0; k < pools.size(); ++k) th.emplace_back([&, k]() {
      TSqliteSampleDescriptorPool& pool = *pools[k];
      TSampleDescriptors mine = d;
      char name[64];
      int in_txn = 0;
      for (int i = (int)k; i < n_rows; i += (int)pools.size()) {
        if (bulk > 1 && in_txn == 0) pool.BeginBulk();
        snprintf(name, sizeof(name), "/nonexistent/f%07d.wav", i);
        if (mode & 2) pool.InsertPackedSample(name, "wav", r);
        else { mine.mFileName = name; pool.InsertSample(name, mine); }
        if ((mode & 8) && i % 3 == 0) { snprintf(name, sizeof(name), "/nonexistent/bad%07d.wav", i); pool.InsertFailedSample(name, "Sample failed to load: test"); }
        if (bulk > 1 && ++in_txn == bulk) { pool.EndBulk(); in_txn = 0; }
      }
      if ((mode & 16) && k == 0 && n_rows > 0) {           // the first row's name once more: the later row replaces the earlier one on every path
        snprintf(name, sizeof(name), "/nonexistent/f%07d.wav", 0);
        try { pool.InsertFailedSample(name, "again"); dup_accepted = true; } catch (const std::exception&) {}
      }
      if (in_txn) pool.EndBulk();
      if (mode & 5) pool.EndBulkLoad();
      pool.Close();
    });
    for (auto& t : th) t.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (bytes) { *bytes = 0; for (const auto& n : names) { struct stat st; if (stat(n.c_str(), &st) == 0) *bytes += (long long)st.st_size; } }
    if ((mode & 16) && !dup_accepted) return -4.0;
    return secs;
  } catch (const std::exception&) { return -2.0; }
}

// append shard databases to `db` (TSqliteSampleDescriptorPool::MergeFrom); returns rows merged or < 0
int afxh_merge_shards(const char* db, const char* const* shards, int n_shards, int delete_shards)
{
  try {
    TSqliteSampleDescriptorPool pool;
    if (!pool.Open(db)) return -1;
    return pool.MergeFrom(std::vector<std::string>(shards, shards + n_shards), delete_shards != 0);
  } catch (const std::exception&) { return -2; }
}

// one long file through the part path (TGpuSampleAnalyser::AnalyzeInParts) into the pool
int afxh_extract_one_in_parts(const char* db, const char* filename, int hop, const int* devices, int n_devices, int n_parts)
{
  try {
    TSqliteSampleDescriptorPool pool;
    if (!pool.Open(db)) return -1;
    TGpuSampleAnalyser an(44100, 2048, hop, std::vector<int>(devices, devices + n_devices), 1);
    try {
      const TSampleDescriptors d = an.AnalyzeInParts(filename, n_parts);
      pool.InsertSample(filename, d);
    } catch (const std::exception& e) {
      pool.InsertFailedSample(filename, std::string("Sample failed to analyse: ") + e.what());
    }
    return 0;
  } catch (const std::exception&) { return -2; }
}

}  // extern "C"

// test hook: n failed rows with file names of `name_len` characters through a direct load (long names make the index records
// spill into overflow pages); returns 0, -1 when the pool declines the direct load, -2 on an exception
extern "C" int afxh_direct_failed_rows(const char* db, int n, int name_len)
{
  try {
    afec::TSqliteSampleDescriptorPool pool;
    if (!pool.Open(db)) return -1;
    if (!pool.BeginDirectLoad()) return -1;
    for (int i = 0; i < n; ++i) {
      char tail[32]; snprintf(tail, sizeof(tail), "%07d.wav", (i * 7919) % n);      // not in sorted order
      std::string name = "/nonexistent/" + std::string((size_t)std::max(0, name_len - 24), (char)('a' + i % 3)) + tail;
      pool.InsertFailedSample(name, "Sample failed to load: test");
    }
    pool.EndBulkLoad();
    pool.Close();
    return 0;
  } catch (const std::exception&) { return -2; }
}

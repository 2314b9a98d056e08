// Small C entry points over the host classes, for language bindings and tests (ctypes).
#include "afx_host.h"

#include <chrono>
#include <cstdio>
#include <algorithm>
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

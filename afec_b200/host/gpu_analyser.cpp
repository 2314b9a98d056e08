// TGpuSampleAnalyser: the reference's TSampleAnalyser entry points (Export/SampleAnalyser.h:33-63;
// SampleAnalyser.cpp:345-416) over the C ABI of libafec_b200.so.
//
// Each "slot" owns one afx context (stream + device buffers) and one pinned PCM ring slot.  ExtractBatch
// starts one thread per slot; a thread repeatedly claims the next chunk of files, decodes it into its
// pinned slot, runs upload -> compute -> download on its own stream and hands the results to the pool.
// With two or more slots per device the decode / H2D of one chunk overlaps the kernels of another.
#include "afx_host.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

namespace afec {

struct TGpuSampleAnalyser::Slot {
  afx_ctx* ctx = nullptr;
  int device = 0;
  unsigned char* pinned = nullptr; size_t pinned_cap = 0;
  ~Slot() { if (ctx) { if (pinned) afx_host_free(ctx, pinned); afx_destroy(ctx); } }
  bool reserve(size_t bytes) {
    if (bytes <= pinned_cap) return true;
    if (pinned) afx_host_free(ctx, pinned);
    pinned = nullptr; pinned_cap = 0;
    void* p = nullptr;
    const size_t want = bytes + bytes / 4 + 4096;
    if (afx_host_alloc(ctx, want, &p) != AFX_OK) return false;
    pinned = (unsigned char*)p; pinned_cap = want;
    return true;
  }
};

TGpuSampleAnalyser::TGpuSampleAnalyser(int SampleRate, int FftFrameSize, int HopFrameSize,
                                       const std::vector<int>& Devices, int SlotsPerDevice)
  : mSampleRate(SampleRate), mFftFrameSize(FftFrameSize), mHopFrameSize(HopFrameSize)
{
  if (Devices.empty() || SlotsPerDevice < 1) throw TReadableException("TGpuSampleAnalyser: no devices");
  for (int d : Devices) for (int s = 0; s < SlotsPerDevice; ++s) {
    std::unique_ptr<Slot> slot(new Slot());
    afx_config cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.device = d; cfg.sample_rate = SampleRate; cfg.fft_size = FftFrameSize; cfg.hop_size = HopFrameSize;
    cfg.features = AFX_FEAT_ALL;
    slot->device = d;
    if (afx_create(&cfg, &slot->ctx) != AFX_OK)
      throw TReadableException(std::string("TGpuSampleAnalyser: ") + afx_last_error(nullptr));
    mSlots.push_back(std::move(slot));
  }
  mNumDevices = (int)Devices.size();
}

TGpuSampleAnalyser::~TGpuSampleAnalyser() {}

static void describe(const TDecodedAudio& a, const void* pcm, afx_file& f)
{
  memset(&f, 0, sizeof(f));
  f.pcm = pcm; f.nframes = a.mFrames; f.channels = a.mChannels; f.src_rate = a.mSampleRate;
  f.format = a.mFormat; f.bit_depth = a.mBitDepth; f.file_size = a.mFileSize;
}

static const char* file_status_message(int status)
{
  switch (status) {                            // SampleAnalyser.cpp:472-482
    case AFX_FILE_BAD_CHANNELS: return "Unsupported audio file channel layout: Supporting mono, stereo, 3.0, 5.0, 5.1 and 7.1 audio files only.";
    case AFX_FILE_EMPTY: return "Sample file is empty, probably failed to read.";
    case AFX_FILE_UNSUPPORTED: return "Unsupported sample rate, sample format or file length.";
    default: return "Unknown error";
  }
}

TSampleDescriptors TGpuSampleAnalyser::AnalyzeInParts(const std::string& FileName, int NumParts) const
{
  TDecodedAudio audio;
  ReadWaveFile(FileName, audio);
  return AnalyzeDecodedInParts(FileName, audio, NumParts);
}

TSampleDescriptors TGpuSampleAnalyser::AnalyzeDecodedInParts(const std::string& FileName, const TDecodedAudio& audio, int NumParts) const
{
  if (audio.mChannels < 1 || audio.mChannels > 8) throw TReadableException(file_status_message(AFX_FILE_BAD_CHANNELS));
  if (audio.mFrames == 0) throw TReadableException(file_status_message(AFX_FILE_EMPTY));
  std::lock_guard<std::mutex> lock(mSingleLock);
  // one slot per device first (slots are laid out device-major), so parts spread over the GPUs
  std::vector<Slot*> slots;
  const int per_dev = (int)mSlots.size() / mNumDevices;
  for (int d = 0; d < mNumDevices; ++d) slots.push_back(mSlots[(size_t)d * per_dev].get());
  const int n_parts = NumParts > 0 ? NumParts : (int)slots.size();
  std::vector<afx_part> parts((size_t)n_parts);
  if (afx_part_plan(mSampleRate, audio.mFrames, audio.mSampleRate, n_parts, parts.data()) != AFX_OK)
    throw TReadableException("afx_part_plan failed");
  afx_file whole; describe(audio, nullptr, whole);
  const size_t frame_bytes = (size_t)audio.mChannels * (audio.mFormat == AFX_PCM_I16 ? 2 : 4);
  std::vector<afx_partjob*> jobs((size_t)n_parts, nullptr);
  std::vector<afx_part_sums> sums((size_t)n_parts);
  std::vector<std::string> errors((size_t)n_parts);
  struct Closer { std::vector<afx_partjob*>& j; ~Closer() { for (auto* p : j) if (p) afx_part_close(p); } } closer{ jobs };
  {
    // phase A carries the bulk (H2D copy, downmix, resample): one thread per part so the GPUs work side by side
    std::vector<std::thread> th;
    for (int p = 0; p < n_parts; ++p) th.emplace_back([&, p]() {
      Slot* S = slots[(size_t)p % slots.size()];
      const unsigned char* slice = audio.mBytes.data() + (size_t)parts[p].src_begin * frame_bytes;
      if (afx_part_open(S->ctx, &whole, &parts[p], slice, &jobs[p]) != AFX_OK || afx_part_peak(jobs[p], &sums[p]) != AFX_OK)
        errors[p] = afx_last_error(S->ctx);
    });
    for (auto& t : th) t.join();
  }
  for (const auto& e : errors) if (!e.empty()) throw TReadableException(e);
  auto combine = [&]() { afx_part_sums g; afx_part_sums_init(&g); for (const auto& s : sums) afx_part_sums_merge(&g, &s); return g; };
  afx_part_sums g = combine();
  for (int p = 0; p < n_parts; ++p)
    if (afx_part_trim(jobs[p], &g, &sums[p]) != AFX_OK) throw TReadableException(afx_last_error(slots[(size_t)p % slots.size()]->ctx));
  g = combine();
  for (int p = 0; p < n_parts; ++p)
    if (afx_part_effective(jobs[p], &g, &sums[p]) != AFX_OK) throw TReadableException(afx_last_error(slots[(size_t)p % slots.size()]->ctx));
  g = combine();
  Slot& S0 = *slots[0];
  int64_t begin = 0, count = 0;
  afx_part_window(S0.ctx, &whole, &g, &begin, &count);
  std::vector<float> window((size_t)std::max<int64_t>(count, 1), 0.0f);
  int64_t got = 0;
  for (int p = 0; p < n_parts; ++p) { const int64_t r = afx_part_read(jobs[p], begin, count, window.data()); if (r > 0) got += r; }
  if (got != count) throw TReadableException("long file: incomplete analysis window");
  afx_batch* b = nullptr;
  if (afx_analyze_conditioned(S0.ctx, &whole, &g, window.data(), begin, count, &b) != AFX_OK) throw TReadableException(afx_last_error(S0.ctx));
  afx_file_result r;
  afx_batch_result(b, 0, &r);
  TSampleDescriptors out;
  out.mFileName = FileName; out.mFileType = ExtractFileExtension(FileName);
  out.Assign(r);
  afx_batch_free(b);
  return out;
}

TSampleDescriptors TGpuSampleAnalyser::Analyze(const std::string& FileName) const
{
  TDecodedAudio audio;
  ReadWaveFile(FileName, audio);               // throws with the loader's message
  if (mNumDevices > 1 && mLongFileBytes && audio.mBytes.size() >= mLongFileBytes) return AnalyzeDecodedInParts(FileName, audio, 0);
  std::lock_guard<std::mutex> lock(mSingleLock);
  Slot& S = *mSlots[0];
  afx_file f; describe(audio, audio.mBytes.data(), f);
  afx_batch* b = nullptr;
  if (afx_analyze(S.ctx, &f, 1, &b) != AFX_OK) throw TReadableException(afx_last_error(S.ctx));
  afx_file_result r;
  afx_batch_result(b, 0, &r);
  if (r.status != AFX_FILE_OK) { const std::string m = file_status_message(r.status); afx_batch_free(b); throw TReadableException(m); }
  TSampleDescriptors out;
  out.mFileName = FileName; out.mFileType = ExtractFileExtension(FileName);
  out.Assign(r);
  afx_batch_free(b);
  return out;
}

void TGpuSampleAnalyser::Extract(const std::string& FileName, TSampleDescriptorPool* pPool, std::mutex& PoolLock) const
{
  TDecodedAudio audio;
  try { ReadWaveFile(FileName, audio); }
  catch (const std::exception& e) {
    const std::lock_guard<std::mutex> lock(PoolLock);
    pPool->InsertFailedSample(FileName, std::string("Sample failed to load: ") + e.what());
    return;
  }
  if (audio.mChannels < 1 || audio.mChannels > 8 || audio.mFrames == 0) {    // load-time rejections, SA.cpp:472-482
    const std::lock_guard<std::mutex> lock(PoolLock);
    pPool->InsertFailedSample(FileName, std::string("Sample failed to load: ") +
      file_status_message(audio.mFrames == 0 ? AFX_FILE_EMPTY : AFX_FILE_BAD_CHANNELS));
    return;
  }
  TSampleDescriptors results;
  try {
    if (mNumDevices > 1 && mLongFileBytes && audio.mBytes.size() >= mLongFileBytes) {
      results = AnalyzeDecodedInParts(FileName, audio, 0);
      const std::lock_guard<std::mutex> lock(PoolLock);
      pPool->InsertSample(FileName, results);
      return;
    }
    std::lock_guard<std::mutex> lock(mSingleLock);
    Slot& S = *mSlots[0];
    afx_file f; describe(audio, audio.mBytes.data(), f);
    afx_batch* b = nullptr;
    if (afx_analyze(S.ctx, &f, 1, &b) != AFX_OK) throw TReadableException(afx_last_error(S.ctx));
    afx_file_result r; afx_batch_result(b, 0, &r);
    if (r.status != AFX_FILE_OK) { const std::string m = file_status_message(r.status); afx_batch_free(b); throw TReadableException(m); }
    results.mFileName = FileName; results.mFileType = ExtractFileExtension(FileName);
    results.Assign(r);
    afx_batch_free(b);
  } catch (const std::exception& e) {
    const std::lock_guard<std::mutex> lock(PoolLock);
    pPool->InsertFailedSample(FileName, std::string("Sample failed to analyse: ") + e.what());
    return;
  }
  const std::lock_guard<std::mutex> lock(PoolLock);
  pPool->InsertSample(FileName, results);
}

int TGpuSampleAnalyser::ExtractBatch(const std::vector<std::string>& FileNames, TSampleDescriptorPool* pPool,
                                     std::mutex& PoolLock, TProgress* pProgress, const volatile bool* pAbort) const
{
  const auto t0 = std::chrono::steady_clock::now();
  // chunks by file size on disk (a cheap upper bound of the decoded PCM for 16-bit files)
  struct Chunk { size_t first, count; };
  std::vector<Chunk> chunks;
  {
    size_t first = 0, bytes = 0;
    for (size_t i = 0; i < FileNames.size(); ++i) {
      struct { int64_t size; } s; TDecodedAudio probe; (void)probe;
      FILE* f = fopen(FileNames[i].c_str(), "rb"); s.size = 0;
      if (f) { fseek(f, 0, SEEK_END); s.size = ftell(f); fclose(f); }
      if (i > first && (bytes + (size_t)s.size > mMaxBatchBytes || (int)(i - first) >= mMaxBatchFiles)) {
        chunks.push_back({ first, i - first }); first = i; bytes = 0;
      }
      bytes += (size_t)s.size;
    }
    if (first < FileNames.size()) chunks.push_back({ first, FileNames.size() - first });
  }
  std::atomic<size_t> next(0);
  std::atomic<long long> failed(0), frames(0), rframes(0), files(0);
  std::mutex stat_lock; double audio_s = 0.0;

  auto worker = [&](Slot* S) {
    std::vector<TDecodedAudio> audio;
    std::vector<std::string> load_error;
    std::vector<afx_file> descr;
    std::vector<int> index;                   // batch position -> file position inside the chunk
    for (;;) {
      if (pAbort && *pAbort) return;
      const size_t ci = next.fetch_add(1);
      if (ci >= chunks.size()) return;
      const Chunk c = chunks[ci];
      audio.assign(c.count, TDecodedAudio()); load_error.assign(c.count, std::string());
      size_t total = 0;
      for (size_t k = 0; k < c.count; ++k) {
        try {
          ReadWaveFile(FileNames[c.first + k], audio[k]);
          if (audio[k].mChannels > 8) load_error[k] = file_status_message(AFX_FILE_BAD_CHANNELS);
        } catch (const std::exception& e) { load_error[k] = e.what(); if (load_error[k].empty()) load_error[k] = "Audio file failed to load: Unknown error"; }
        if (load_error[k].empty()) total += (audio[k].mBytes.size() + 15) & ~(size_t)15;
      }
      // stage the chunk's PCM contiguously in the pinned slot: one H2D copy for the whole chunk
      descr.clear(); index.clear();
      std::string batch_error;
      afx_batch* b = nullptr;
      if (!S->reserve(total + 16)) batch_error = "out of pinned host memory";
      else {
        size_t off = 0;
        for (size_t k = 0; k < c.count; ++k) {
          if (!load_error[k].empty()) continue;
          memcpy(S->pinned + off, audio[k].mBytes.data(), audio[k].mBytes.size());
          afx_file f; describe(audio[k], S->pinned + off, f);
          descr.push_back(f); index.push_back((int)k);
          off += audio[k].mBytes.size();
          // keep files back to back (the library merges host-contiguous files into one copy); int16 / float32
          // alignment is preserved because every file's byte count is a multiple of its sample size
          std::vector<unsigned char>().swap(audio[k].mBytes);
        }
        if (!descr.empty() && afx_analyze(S->ctx, descr.data(), (int32_t)descr.size(), &b) != AFX_OK)
          batch_error = afx_last_error(S->ctx);      // a CUDA failure fails this batch's files only
      }
      // hand the chunk to the pool in file order
      TSampleDescriptors results;
      const std::lock_guard<std::mutex> lock(PoolLock);
      pPool->BeginBulk();
      size_t bi = 0; double chunk_audio = 0.0;
      for (size_t k = 0; k < c.count; ++k) {
        const std::string& name = FileNames[c.first + k];
        try {
          if (!load_error[k].empty()) { pPool->InsertFailedSample(name, "Sample failed to load: " + load_error[k]); ++failed; continue; }
          if (!batch_error.empty() || !b) { pPool->InsertFailedSample(name, "Sample failed to analyse: " + batch_error); ++failed; ++bi; continue; }
          afx_file_result r; afx_batch_result(b, (int32_t)bi, &r); ++bi;
          if (r.status != AFX_FILE_OK) { pPool->InsertFailedSample(name, std::string("Sample failed to load: ") + file_status_message(r.status)); ++failed; continue; }
          results.mFileName = name; results.mFileType = ExtractFileExtension(name);
          results.Assign(r);
          pPool->InsertSample(name, results);
          frames += r.n_frames; rframes += r.n_rhythm_frames; ++files;
          chunk_audio += results.mHeader[1];
        } catch (const std::exception&) { ++failed; }
      }
      pPool->EndBulk();
      if (b) afx_batch_free(b);
      { std::lock_guard<std::mutex> sl(stat_lock); audio_s += chunk_audio; }
    }
  };

  std::vector<std::thread> threads;
  for (auto& s : mSlots) threads.emplace_back(worker, s.get());
  for (auto& t : threads) t.join();
  if (pProgress) {
    pProgress->mFiles = files; pProgress->mFailed = failed; pProgress->mMainFrames = frames; pProgress->mRhythmFrames = rframes;
    pProgress->mAudioSeconds = audio_s;
    pProgress->mSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return (int)failed;
}

}  // namespace afec

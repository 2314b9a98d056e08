// TGpuSampleAnalyser: the reference's TSampleAnalyser entry points (Export/SampleAnalyser.h:33-63;
// SampleAnalyser.cpp:345-416) over the C ABI of libafec_b200.so.
//
// Each "slot" owns one afx context (stream + device buffers) and one pinned PCM ring slot.  ExtractBatch runs a
// three-stage pipeline over the slots: decode threads fill pinned slots, one GPU thread per slot analyses them, one
// sink thread inserts the rows (see ExtractBatch).
#include "afx_host.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <sys/stat.h>
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

static std::unique_ptr<TGpuSampleAnalyser::Slot> make_slot(int Device, int SampleRate, int FftFrameSize, int HopFrameSize, unsigned Features);

TGpuSampleAnalyser::TGpuSampleAnalyser(int SampleRate, int FftFrameSize, int HopFrameSize,
                                       const std::vector<int>& Devices, int SlotsPerDevice, bool PackRowsOnDevice)
  : mSampleRate(SampleRate), mFftFrameSize(FftFrameSize), mHopFrameSize(HopFrameSize), mPackRows(PackRowsOnDevice)
{
  if (Devices.empty() || SlotsPerDevice < 1) throw TReadableException("TGpuSampleAnalyser: no devices");
  // ring slots: with PackRowsOnDevice their contexts also pack every row's msgpack BLOBs on the GPU (the batched
  // extractor then downloads rows, not arrays); slot order is device-major
  for (int d : Devices) for (int s = 0; s < SlotsPerDevice; ++s)
    mSlots.push_back(make_slot(d, SampleRate, FftFrameSize, HopFrameSize, AFX_FEAT_ALL | (PackRowsOnDevice ? AFX_FEAT_PACK : 0u)));
  mNumDevices = (int)Devices.size();
  mDevices = Devices;
}

std::unique_ptr<TGpuSampleAnalyser::Slot> make_slot(int Device, int SampleRate, int FftFrameSize, int HopFrameSize, unsigned Features)
{
  std::unique_ptr<TGpuSampleAnalyser::Slot> slot(new TGpuSampleAnalyser::Slot());
  afx_config cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.device = Device; cfg.sample_rate = SampleRate; cfg.fft_size = FftFrameSize; cfg.hop_size = HopFrameSize;
  cfg.features = Features;
  slot->device = Device;
  if (afx_create(&cfg, &slot->ctx) != AFX_OK)
    throw TReadableException(std::string("TGpuSampleAnalyser: ") + afx_last_error(nullptr));
  return slot;
}

// the single-file entry points (Analyze / Extract / the last step of AnalyzeInParts) return descriptor ARRAYS: they run
// on their own context, created on first use
TGpuSampleAnalyser::Slot& TGpuSampleAnalyser::SingleSlot() const
{
  if (!mSingle) mSingle = make_slot(mDevices[0], mSampleRate, mFftFrameSize, mHopFrameSize, AFX_FEAT_ALL);
  return *mSingle;
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
  ReadAudioFile(FileName, audio);
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
  const size_t frame_bytes = (size_t)audio.mChannels * (size_t)afx_pcm_bytes(audio.mFormat);
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
  Slot& S0 = SingleSlot();
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
  ReadAudioFile(FileName, audio);               // throws with the loader's message
  if (mNumDevices > 1 && mLongFileBytes && audio.mBytes.size() >= mLongFileBytes) return AnalyzeDecodedInParts(FileName, audio, 0);
  std::lock_guard<std::mutex> lock(mSingleLock);
  Slot& S = SingleSlot();
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
  try { ReadAudioFile(FileName, audio); }
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
    Slot& S = SingleSlot();
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

// The batched extractor as a three-stage pipeline over a ring of slots (one afx context + one pinned buffer each):
//
//   decode threads  claim the next chunk of files, take a FREE slot, probe the headers (sizes, formats) and read every
//                   file's sample bytes straight into the slot's pinned buffer, back to back      -> READY
//   GPU threads     one per slot: afx_analyze (one H2D copy of the chunk, the kernels, one D2H block)   -> DONE
//   sink thread     rows of finished chunks go to the pool one by one (PoolLock is held per ROW, a bulk
//                   transaction per chunk), then the slot's batch is freed                               -> FREE
//
// The decode threads never wait for the GPU (only for a free slot), the GPU never waits for sqlite: with the reference's
// one-file-at-a-time loop (Crawler.cpp:706-728) decode, analysis and insert of a file are serial on one thread.
int TGpuSampleAnalyser::ExtractBatch(const std::vector<std::string>& FileNames, TSampleDescriptorPool* pPool,
                                     std::mutex& PoolLock, TProgress* pProgress, const volatile bool* pAbort) const
{
  std::vector<TSampleDescriptorPool*> pools(1, pPool);
  std::vector<std::mutex*> locks(1, &PoolLock);
  return ExtractBatchSharded(FileNames, pools, locks, pProgress, pAbort);
}

// One sink thread per pool: every finished chunk goes to the pool whose sink thread takes it first, so N pools (N
// afec-ll.db shard files, merged or attached afterwards) are written side by side -- one sqlite writer moves 0.3-0.7 GB/s.
int TGpuSampleAnalyser::ExtractBatchSharded(const std::vector<std::string>& FileNames, const std::vector<TSampleDescriptorPool*>& Pools,
                                            const std::vector<std::mutex*>& PoolLocks, TProgress* pProgress, const volatile bool* pAbort) const
{
  if (Pools.empty() || Pools.size() != PoolLocks.size()) throw TReadableException("ExtractBatchSharded: one lock per pool");
  const auto t0 = std::chrono::steady_clock::now();
  // chunks by file size on disk (an upper bound of the raw PCM bytes)
  struct Chunk { size_t first, count; };
  std::vector<Chunk> chunks;
  {
    size_t first = 0, bytes = 0;
    for (size_t i = 0; i < FileNames.size(); ++i) {
      struct stat st; const size_t size = (stat(FileNames[i].c_str(), &st) == 0) ? (size_t)st.st_size : 0;
      if (i > first && (bytes + size > mMaxBatchBytes || (int)(i - first) >= mMaxBatchFiles)) {
        chunks.push_back({ first, i - first }); first = i; bytes = 0;
      }
      bytes += size;
    }
    if (first < FileNames.size()) chunks.push_back({ first, FileNames.size() - first });
  }

  enum { kFree = 0, kFilling, kReady, kDone };
  struct Work {                               // per slot
    int state = kFree;
    Chunk chunk{ 0, 0 };
    std::vector<TAudioInfo> info;
    std::vector<std::string> load_error;
    std::vector<afx_file> descr;
    std::vector<int> index;                   // batch position -> file position inside the chunk
    std::string batch_error;
    afx_batch* batch = nullptr;
  };
  const size_t S = mSlots.size();
  std::vector<Work> work(S);
  std::mutex mu; std::condition_variable cv;  // guards every Work::state, `done_order`, `decoders_left`, `workers_left`
  std::deque<size_t> done_order;
  std::atomic<size_t> next(0);
  int decoders_left = 0, workers_left = (int)S;
  std::atomic<long long> failed(0), frames(0), rframes(0), files(0);
  std::mutex stat_mu; double audio_s = 0.0;
  int sinks_left = (int)Pools.size(); (void)sinks_left;
  auto aborted = [&]() { return pAbort && *pAbort; };

  auto decoder = [&]() {
    for (;;) {
      const size_t ci = aborted() ? chunks.size() : next.fetch_add(1);
      if (ci >= chunks.size()) break;
      size_t si = S;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { for (size_t k = 0; k < S; ++k) if (work[k].state == kFree) { si = k; return true; } return false; });
        work[si].state = kFilling;
      }
      Work& W = work[si]; Slot* Sl = mSlots[si].get();
      W.chunk = chunks[ci];
      const size_t n = W.chunk.count;
      W.info.assign(n, TAudioInfo()); W.load_error.assign(n, std::string()); W.descr.clear(); W.index.clear(); W.batch_error.clear(); W.batch = nullptr;
      size_t total = 0;
      for (size_t k = 0; k < n; ++k) {
        try {
          ProbeAudioFile(FileNames[W.chunk.first + k], W.info[k]);
          if (W.info[k].mChannels > 8) W.load_error[k] = file_status_message(AFX_FILE_BAD_CHANNELS);
          else if (W.info[k].mFrames == 0) W.load_error[k] = file_status_message(AFX_FILE_EMPTY);
        } catch (const std::exception& e) { W.load_error[k] = e.what(); if (W.load_error[k].empty()) W.load_error[k] = "Audio file failed to load: Unknown error"; }
        if (W.load_error[k].empty()) total += (W.info[k].mDataBytes + 15) & ~(size_t)15;
      }
      if (!Sl->reserve(total + 16)) W.batch_error = "out of pinned host memory";
      else {
        // files back to back in the pinned slot: one H2D copy for the whole chunk (the library merges host-contiguous
        // files; every file starts on a 4-byte boundary so 16- and 32-bit samples stay aligned)
        size_t off = 0;
        for (size_t k = 0; k < n; ++k) {
          if (!W.load_error[k].empty()) continue;
          try { ReadAudioData(W.info[k], Sl->pinned + off); }
          catch (const std::exception& e) { W.load_error[k] = e.what(); continue; }
          afx_file f; memset(&f, 0, sizeof(f));
          f.pcm = Sl->pinned + off; f.nframes = W.info[k].mFrames; f.channels = W.info[k].mChannels; f.src_rate = W.info[k].mSampleRate;
          f.format = W.info[k].mFormat; f.bit_depth = W.info[k].mBitDepth; f.file_size = W.info[k].mFileSize;
          W.descr.push_back(f); W.index.push_back((int)k);
          off += (W.info[k].mDataBytes + 3) & ~(size_t)3;
        }
      }
      { std::lock_guard<std::mutex> lk(mu); W.state = kReady; }
      cv.notify_all();
    }
    { std::lock_guard<std::mutex> lk(mu); --decoders_left; }
    cv.notify_all();
  };

  auto gpu_worker = [&](size_t si) {
    Work& W = work[si]; Slot* Sl = mSlots[si].get();
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { return W.state == kReady || (decoders_left == 0 && W.state == kFree); });
        if (W.state != kReady) break;
      }
      if (W.batch_error.empty() && !W.descr.empty()) {
        // rows packed on the device come back without the framed arrays (afx_batch_download_rows)
        bool ok = afx_batch_create(Sl->ctx, W.descr.data(), (int32_t)W.descr.size(), &W.batch) == AFX_OK;
        ok = ok && afx_batch_upload(W.batch) == AFX_OK && afx_batch_compute(W.batch) == AFX_OK &&
             (mPackRows ? afx_batch_download_rows(W.batch) : afx_batch_download(W.batch)) == AFX_OK && afx_batch_sync(W.batch) == AFX_OK;
        if (!ok) {                                    // a CUDA failure fails this chunk's files only
          W.batch_error = afx_last_error(Sl->ctx);
          if (W.batch) { afx_batch_free(W.batch); W.batch = nullptr; }
        }
      }
      { std::lock_guard<std::mutex> lk(mu); W.state = kDone; done_order.push_back(si); }
      cv.notify_all();
    }
    { std::lock_guard<std::mutex> lk(mu); --workers_left; }
    cv.notify_all();
  };

  auto sink = [&](size_t pi) {
    TSampleDescriptorPool* pPool = Pools[pi];
    std::mutex& PoolLock = *PoolLocks[pi];
    const bool packed = mPackRows && pPool->AcceptsPackedSamples();
    TSampleDescriptors results;
    for (;;) {
      size_t si;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { return !done_order.empty() || workers_left == 0; });
        if (done_order.empty()) break;
        si = done_order.front(); done_order.pop_front();
      }
      Work& W = work[si];
      { const std::lock_guard<std::mutex> lock(PoolLock); pPool->BeginBulk(); }
      size_t bi = 0; double chunk_audio = 0.0;
      for (size_t k = 0; k < W.chunk.count; ++k) {
        const std::string& name = FileNames[W.chunk.first + k];
        try {
          if (!W.load_error[k].empty()) {
            const std::lock_guard<std::mutex> lock(PoolLock);
            pPool->InsertFailedSample(name, "Sample failed to load: " + W.load_error[k]); ++failed; continue;
          }
          if (!W.batch_error.empty() || !W.batch) {
            const std::lock_guard<std::mutex> lock(PoolLock);
            pPool->InsertFailedSample(name, "Sample failed to analyse: " + W.batch_error); ++failed; ++bi; continue;
          }
          afx_file_result r; afx_batch_result(W.batch, (int32_t)bi, &r); ++bi;
          if (r.status != AFX_FILE_OK) {
            const std::lock_guard<std::mutex> lock(PoolLock);
            pPool->InsertFailedSample(name, std::string("Sample failed to load: ") + file_status_message(r.status)); ++failed; continue;
          }
          if (packed) {                        // the BLOBs were packed on the GPU: bind them as they are
            const std::lock_guard<std::mutex> lock(PoolLock); pPool->InsertPackedSample(name, ExtractFileExtension(name), r);
          } else if (!r.fs[0]) {
            throw TReadableException("the pool takes no packed rows: build the analyser with PackRowsOnDevice = false");
          } else {
            results.mFileName = name; results.mFileType = ExtractFileExtension(name);
            results.Assign(r);                 // outside the lock: only the insert itself is serialised
            const std::lock_guard<std::mutex> lock(PoolLock); pPool->InsertSample(name, results);
          }
          frames += r.n_frames; rframes += r.n_rhythm_frames; ++files;
          chunk_audio += r.header[1];
        } catch (const std::exception& e) {
          // SampleAnalyser.cpp:372-408: whatever goes wrong with a file ends as a failed row, never as a missing one
          ++failed;
          try { const std::lock_guard<std::mutex> lock(PoolLock); pPool->InsertFailedSample(name, std::string("Sample failed to analyse: ") + e.what()); }
          catch (const std::exception&) {}
        }
      }
      { const std::lock_guard<std::mutex> lock(PoolLock); pPool->EndBulk(); }
      if (W.batch) { afx_batch_free(W.batch); W.batch = nullptr; }
      { std::lock_guard<std::mutex> sl(stat_mu); audio_s += chunk_audio; }
      { std::lock_guard<std::mutex> lk(mu); W.state = kFree; }
      cv.notify_all();
    }
  };

  const int n_dec = std::max(1, std::min<int>(mDecodeThreads > 0 ? mDecodeThreads : (int)S, (int)std::max<size_t>(1, chunks.size())));
  decoders_left = n_dec;
  std::vector<std::thread> threads;
  for (int d = 0; d < n_dec; ++d) threads.emplace_back(decoder);
  for (size_t si = 0; si < S; ++si) threads.emplace_back(gpu_worker, si);
  for (size_t pi = 0; pi < Pools.size(); ++pi) threads.emplace_back(sink, pi);
  for (auto& t : threads) t.join();
  if (pProgress) {
    pProgress->mFiles = files; pProgress->mFailed = failed; pProgress->mMainFrames = frames; pProgress->mRhythmFrames = rframes;
    pProgress->mAudioSeconds = audio_s;
    pProgress->mSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return (int)failed;
}

}  // namespace afec

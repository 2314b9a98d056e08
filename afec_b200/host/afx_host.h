// Host side of the B200 descriptor path: the reference's extractor interface over libafec_b200.so.
//
// Mirrors (names, argument meaning, error behaviour), file:line relative to the reference root:
//   TSampleAnalyser            Source/Crawler/FeatureExtraction/Export/SampleAnalyser.h:28-64
//   TSampleDescriptors         Export/SampleDescriptors.h:396-466   (low-level set only)
//   TSampleDescriptorPool      Export/SampleDescriptorPool.h:20-81
//   TSqliteSampleDescriptorPool Export/SqliteSampleDescriptorPool.h:21-86 (afec-ll.db writer)
// File decoding stays on the host (WAV here; the reference's FLAC / Ogg / MP3 decoders are outside this
// path) and the descriptor sink stays sqlite; everything between them runs on the GPU through the C ABI
// in include/afec_b200.h.  There is no CPU implementation of the analysis in this library.
#ifndef AFX_HOST_H_
#define AFX_HOST_H_

#include <cstdint>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/afec_b200.h"

namespace afec {

// the reference throws TReadableException(what); the adapter keeps the messages
struct TReadableException : public std::runtime_error {
  explicit TReadableException(const std::string& what) : std::runtime_error(what) {}
};

// ---- names and order of the low-level descriptor set (SampleDescriptors.cpp:154-203) ----------------
extern const char* const kHeaderNames[23];           // scalar descriptors held in afx_file_result.header
extern const char* const kFramedScalarNames[AFX_N_FS];
extern const char* const kFramedVectorNames[AFX_N_FV];
extern const int kFramedVectorBands[AFX_N_FV];
extern const char* const kStatNames[AFX_N_STATS];     // SampleDescriptors.h:186-210

// Values of one analysed file.  Same content as the reference's TSampleDescriptors low-level members;
// stored as flat arrays in the order of the C ABI result.
struct TSampleDescriptors {
  std::string mFileName;
  std::string mFileType;                 // file extension, SampleAnalyser.cpp:736
  double mHeader[AFX_N_HEADER] = { 0 };  // file_size .. rhythm_final_tempo_confidence (+ conditioning outputs)
  int mFrames = 0, mRhythmFrames = 0;
  std::vector<double> mFramedScalars[AFX_N_FS];      // [series][frame]
  std::vector<double> mFramedVectors[AFX_N_FV];      // [descriptor][frame * nbands + band]
  double mStats[AFX_N_SERIES][AFX_N_STATS] = { { 0 } };
  void Assign(const afx_file_result& r);
};

// ---- the sink -------------------------------------------------------------------------------------
class TSampleDescriptorPool {
public:
  virtual ~TSampleDescriptorPool() {}
  virtual void InsertSample(const std::string& FileName, const TSampleDescriptors& Results) = 0;
  virtual void InsertFailedSample(const std::string& FileName, const std::string& Reason) = 0;
  // optional: group the following inserts into one transaction (content identical to one per file)
  virtual void BeginBulk() {}
  virtual void EndBulk() {}
  // optional: rows whose BLOBs were packed on the GPU (afx_file_result.packed, AFX_FEAT_PACK): a pool that accepts them
  // binds the bytes as they are; the default unpacks nothing and is never called
  virtual bool AcceptsPackedSamples() const { return false; }
  virtual void InsertPackedSample(const std::string& FileName, const std::string& FileType, const afx_file_result& Result)
  { (void)FileName; (void)FileType; (void)Result; throw TReadableException("this pool takes no packed rows"); }
};

// afec-ll.db: schema, pragmas, msgpack BLOB encoding and status strings of the reference
// (SqliteSampleDescriptorPool.cpp:1116-1350, 1551-1733; Database.cpp:337-351)
class TSqliteSampleDescriptorPool : public TSampleDescriptorPool {
public:
  enum { kCurrentVersion = 2 };          // Export/SqliteSampleDescriptorPool.h:58
  TSqliteSampleDescriptorPool();
  ~TSqliteSampleDescriptorPool() override;
  bool Open(const std::string& DatabaseName, bool ReadOnly = false);
  void Close();
  void SetBasePath(const std::string& BasePath);     // filenames below it are stored relative, '/'-separated
  std::string BasePath() const { return mBasePath; }
  bool IsEmpty() const;
  int NumberOfSamples() const;
  std::vector<std::pair<std::string, int>> SampleModificationDates() const;   // (abs filename, modtime)
  void InsertSample(const std::string& FileName, const TSampleDescriptors& Results) override;
  void InsertFailedSample(const std::string& FileName, const std::string& Reason) override;
  void RemoveSample(const std::string& FileName);
  void RemoveSamples(const std::vector<std::string>& FileNames);
  void BeginBulk() override;
  void EndBulk() override;
  bool AcceptsPackedSamples() const override { return true; }
  void InsertPackedSample(const std::string& FileName, const std::string& FileType, const afx_file_result& Result) override;
  // journal-less filling of an EMPTY database (returns false and changes nothing otherwise); EndBulkLoad / Close restore
  // the reference's WAL / NORMAL pragmas
  bool BeginBulkLoad();
  void EndBulkLoad();
  // the same without sqlite in the data path (direct_db_writer.h): rows are written into the file in sqlite's format
  bool BeginDirectLoad();
  // append the rows of other afec-ll.db files (shards written side by side); returns the number of rows taken
  int MergeFrom(const std::vector<std::string>& ShardFiles, bool DeleteShards);
  static std::vector<std::string> ColumnNamesAndTypes();   // "name TYPE" in table order (461 entries)
private:
  void PrepareInsert();
  void InsertRow(const std::string& FileName, const std::string& FileType, const double* Header, const double (*Stats)[AFX_N_STATS],
                 const unsigned char* const* BlobPtr, const int* BlobLen);
  struct Impl;
  std::unique_ptr<Impl> mImpl;
  std::string mBasePath;
  std::string RelativeFilenamePath(const std::string& FileName) const;
};

// msgpack encoding of VR / VVR values (msgpack-c 2.1.5 pack_array / pack_double: 0x9X | 0xdc | 0xdd
// headers, 0xcb + 8 bytes big endian per value)
void PackVR(std::vector<unsigned char>& out, const double* values, size_t n);
void PackVVR(std::vector<unsigned char>& out, const double* values, size_t frames, size_t bands);

// ---- decoded audio --------------------------------------------------------------------------------
// Where a file's samples are and what they look like.  The bytes go to the GPU as they sit in the file (AFX_PCM_* raw
// formats, converted on the device); only 64-bit float data is converted on the host (mHostConvert).
struct TAudioInfo {
  std::string mFileName;
  int64_t mFrames = 0;
  int mChannels = 0, mSampleRate = 0, mBitDepth = 0, mFormat = AFX_PCM_I16;
  int64_t mFileSize = 0, mDataOffset = 0;
  int mFileBytesPerSample = 0, mHostConvert = 0;
  size_t mDataBytes = 0;                 // bytes ReadAudioData() writes: frames x channels x afx_pcm_bytes(mFormat)
};
struct TDecodedAudio {
  std::vector<unsigned char> mBytes;     // interleaved frames in format mFormat
  int64_t mFrames = 0;
  int mChannels = 0, mSampleRate = 0, mBitDepth = 0, mFormat = AFX_PCM_I16;
  int64_t mFileSize = 0;
};
// RIFF / WAVE (PCM 8 / 16 / 24 / 32, IEEE float 32 / 64, WAVE_FORMAT_EXTENSIBLE; WaveFile.cpp:372-407) and AIFF / AIFC
// (PCM 8 / 16 / 24 / 32 big endian, "sowt" little endian, "fl32" / "fl64"; AifFile.cpp:150-372, 436-480) by file extension.
// Throw TReadableException with the reference's messages.  ProbeAudioFile reads headers only; ReadAudioData reads the
// samples into caller memory (mDataBytes bytes -- e.g. a pinned ring slot), zero-filling what the file does not deliver.
void ProbeAudioFile(const std::string& FileName, TAudioInfo& Out);
void ReadAudioData(const TAudioInfo& Info, unsigned char* Dst);
void ReadAudioFile(const std::string& FileName, TDecodedAudio& Out);
void ReadWaveFile(const std::string& FileName, TDecodedAudio& Out);     // == ReadAudioFile (kept for callers of the first version)
bool IsSupportedAudioFileExtension(const std::string& FileName);
int ModificationStatTime(const std::string& FileName);
std::string ExtractFileExtension(const std::string& FileName);

// ---- the analyser -----------------------------------------------------------------------------------
class TGpuSampleAnalyser {
public:
  // TSampleAnalyser(SampleRate, FftFrameSize, HopFrameSize), Export/SampleAnalyser.h:33-36 (+ device list)
  TGpuSampleAnalyser(int SampleRate, int FftFrameSize, int HopFrameSize,
                     const std::vector<int>& Devices = std::vector<int>(1, 0), int SlotsPerDevice = 3, bool PackRowsOnDevice = true);
  ~TGpuSampleAnalyser();

  // Export/SampleAnalyser.h:54-56: analyse one file; throws TReadableException on load / analysis failure
  TSampleDescriptors Analyze(const std::string& FileName) const;
  // Export/SampleAnalyser.h:58-63: analyse and write into the pool; per-file failures become
  // InsertFailedSample rows ("Sample failed to load: ..." / "Sample failed to analyse: ..."), never throws for them
  void Extract(const std::string& FileName, TSampleDescriptorPool* pPool, std::mutex& PoolLock) const;
  // the batched form the Crawler's worker loop (Crawler.cpp:706-728) maps to on a GPU: files are decoded by
  // the slot threads into pinned ring slots, analysed batch by batch, and inserted in file order per batch.
  // Returns the number of files that failed.  `Abort` may be set from a signal handler (Crawler.cpp:69-73).
  struct TProgress { int64_t mFiles = 0, mFailed = 0, mMainFrames = 0, mRhythmFrames = 0; double mAudioSeconds = 0, mSeconds = 0; };
  int ExtractBatch(const std::vector<std::string>& FileNames, TSampleDescriptorPool* pPool, std::mutex& PoolLock,
                   TProgress* pProgress = nullptr, const volatile bool* pAbort = nullptr) const;
  // the same with several pools written side by side (one sink thread per pool; a chunk's rows all go to one pool)
  int ExtractBatchSharded(const std::vector<std::string>& FileNames, const std::vector<TSampleDescriptorPool*>& Pools,
                          const std::vector<std::mutex*>& PoolLocks, TProgress* pProgress = nullptr, const volatile bool* pAbort = nullptr) const;

  // Long files (BASELINE config 5): the decoded file is cut into NumParts sample-range parts (0 = one per device),
  // part p is conditioned on slot p % slots -- its own GPU when the analyser was built over several devices --
  // the per-file reductions are combined here on the host between the three phases, and the <= 20 s the reference
  // analyses run on slot 0.  Same values as Analyze() (include/afec_b200.h, "long files conditioned in parts").
  TSampleDescriptors AnalyzeInParts(const std::string& FileName, int NumParts = 0) const;
  // Analyze() / Extract() route files with at least this many decoded PCM bytes through AnalyzeInParts when the
  // analyser drives more than one device (0 disables)
  void SetLongFileBytes(size_t Bytes) { mLongFileBytes = Bytes; }

  int SampleRate() const { return mSampleRate; }
  int FftFrameSize() const { return mFftFrameSize; }
  int HopFrameSize() const { return mHopFrameSize; }
  void SetMaxBatchBytes(size_t Bytes) { mMaxBatchBytes = Bytes; }
  void SetMaxBatchFiles(int Files) { mMaxBatchFiles = Files; }
  // decode threads of ExtractBatch (0 = one per slot); the reference decodes on hardware_concurrency threads (Crawler.cpp:680-681)
  void SetDecodeThreads(int Threads) { mDecodeThreads = Threads; }

  struct Slot;
private:
  int mSampleRate, mFftFrameSize, mHopFrameSize;
  bool mPackRows = true;
  std::vector<int> mDevices;
  mutable std::unique_ptr<Slot> mSingle;
  Slot& SingleSlot() const;
  size_t mMaxBatchBytes = (size_t)256 << 20;
  int mMaxBatchFiles = 2048;
  int mDecodeThreads = 0;
  size_t mLongFileBytes = (size_t)512 << 20;
  int mNumDevices = 1;
  TSampleDescriptors AnalyzeDecodedInParts(const std::string& FileName, const TDecodedAudio& Audio, int NumParts) const;
  std::vector<std::unique_ptr<Slot>> mSlots;
  mutable std::mutex mSingleLock;        // serialises Analyze()/Extract() on slot 0
};

}  // namespace afec
#endif

// TEST INFRASTRUCTURE -- not part of the product path.
//
// Driver for the UNMODIFIED reference (emuell/AFEC) low-level descriptor path.
// Links against objects compiled from /root/reference by oracle/build_ref.sh and
// calls the reference's public API only:
//   TSampleAnalyser::Analyze  (Source/Crawler/FeatureExtraction/Export/SampleAnalyser.h:54-56)
//   TSampleAnalyser::Extract  (SampleAnalyser.h:58-63, as Crawler.cpp:716-726 does)
//   TSqliteSampleDescriptorPool (Export/SqliteSampleDescriptorPool.h:21-86)
//
//   afec_ref dump  <hop> <out.bin> <wav>...       flat binary dump ("AFXD" layout, see
//                                                  afec_b200/layout.py) of every low-level value
//   afec_ref dumphl <hop> <out.bin> <wav>...      per file the "AFXD" record followed by an "AFXH" record: the high-level
//                                                  descriptors that need no classification model (SampleAnalyser.cpp:1232-1606)
//                                                  and the classification feature vector (SampleClassificationDescriptors.cpp)
//   afec_ref db    <hop> <out.db>  <wav>...       golden afec-ll.db through the reference sink
//   afec_ref bench <hop> <threads> <reps> <out.db|-> <wav>...
//                                                  times Extract() like the Crawler's thread pool
//
// Nothing here is copied from the reference; it is a caller.

#include "CoreTypes/Export/Str.h"
#include "CoreTypes/Export/Log.h"
#include "CoreTypes/Export/File.h"
#include "CoreTypes/Export/Directory.h"
#include "AudioTypes/Export/AudioTypesInit.h"
#include "CoreFileFormats/Export/CoreFileFormatsInit.h"
#include "FeatureExtraction/Export/FeatureExtractionInit.h"
#include "FeatureExtraction/Export/SampleAnalyser.h"
#include "FeatureExtraction/Export/SampleDescriptors.h"
#include "FeatureExtraction/Export/SqliteSampleDescriptorPool.h"
#include "FeatureExtraction/Export/SampleClassificationDescriptors.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace TProductDescription {
  TString ProductName() { return "AfecRefOracle"; }
  TString ProductVendorName() { return "AFEC"; }
  TString ProductProjectsLocation() { return "Crawler/XCrawler"; }
  int MajorVersion() { return 1; }
  int MinorVersion() { return 0; }
  int RevisionVersion() { return 0; }
  TString AlphaOrBetaVersionString() { return ""; }
  TDate ExpirationDate() { return TDate(); }
  TString SupportEMailAddress() { return ""; }
  TString ProductHomeURL() { return ""; }
  TString CopyrightString() { return ""; }
}

#include "CoreTypes/Export/MainEntry.h"

// -------------------------------------------------------------------------------------------------

static void put_i32(FILE* f, int v) { fwrite(&v, 4, 1, f); }
static void put_f64(FILE* f, double v) { fwrite(&v, 8, 1, f); }

static void put_framed(FILE* f, const TSampleDescriptors::TFramedScalarData& d)
{
  for (int i = 0; i < d.mValues.Size(); ++i) put_f64(f, d.mValues[i]);
}

static void put_stats(FILE* f, const TSampleDescriptors::TFramedScalarData& d)
{
  const double s[13] = { d.mMin, d.mMax, d.mMedian, d.mMean, d.mGeometricMean, d.mVariance,
    d.mCentroid, d.mSpread, d.mSkewness, d.mKurtosis, d.mFlatness, d.mDMean, d.mDVariance };
  fwrite(s, 8, 13, f);
}

template <size_t N>
static void put_framed_vec(FILE* f, const TSampleDescriptors::TFramedVectorData<N>& d)
{
  for (int i = 0; i < d.mValues.Size(); ++i)
    for (int b = 0; b < (int)N; ++b) put_f64(f, d.mValues[i][b]);
}

template <size_t N>
static void put_stats_vec(FILE* f, const TSampleDescriptors::TFramedVectorData<N>& d)
{
  // band-major: 13 stats per band, same stat order as the scalar series
  for (int b = 0; b < (int)N; ++b) {
    const double s[13] = { d.mMin[b], d.mMax[b], d.mMedian[b], d.mMean[b], d.mGeometricMean[b],
      d.mVariance[b], d.mCentroid[b], d.mSpread[b], d.mSkewness[b], d.mKurtosis[b],
      d.mFlatness[b], d.mDMean[b], d.mDVariance[b] };
    fwrite(s, 8, 13, f);
  }
}

// One record per file. Layout documented in afec_b200/layout.py (AFXD v1).
static void dump_one(FILE* f, const TSampleDescriptors& R, int status)
{
  fwrite("AFXD", 1, 4, f);
  put_i32(f, status);
  if (status != 0) return;

  const int F = R.mAmplitudeSilence.mValues.Size();
  const int Fr = R.mRhythmComplexOnsets.mValues.Size();
  put_i32(f, F);
  put_i32(f, Fr);

  // 32 header scalars
  double H[32]; memset(H, 0, sizeof(H));
  H[0] = R.mFileSize.mValue; H[1] = R.mFileLength.mValue; H[2] = R.mFileSampleRate.mValue;
  H[3] = R.mFileChannelCount.mValue; H[4] = R.mFileBitDepth.mValue;
  H[5] = R.mEffectiveLength48dB.mValue; H[6] = R.mEffectiveLength24dB.mValue;
  H[7] = R.mEffectiveLength12dB.mValue; H[8] = R.mAnalyzationOffset.mValue;
  H[9] = R.mRhythmComplexOnsetCount.mValue; H[10] = R.mRhythmComplexOnsetContrast.mValue;
  H[11] = R.mRhythmComplexOnsetFrequencyMean.mValue; H[12] = R.mRhythmComplexOnsetStrength.mValue;
  H[13] = R.mRhythmComplexTempo.mValue; H[14] = R.mRhythmComplexTempoConfidence.mValue;
  H[15] = R.mRhythmPercussiveOnsetCount.mValue; H[16] = R.mRhythmPercussiveOnsetContrast.mValue;
  H[17] = R.mRhythmPercussiveOnsetFrequencyMean.mValue; H[18] = R.mRhythmPercussiveOnsetStrength.mValue;
  H[19] = R.mRhythmPercussiveTempo.mValue; H[20] = R.mRhythmPercussiveTempoConfidence.mValue;
  H[21] = R.mRhythmFinalTempo.mValue; H[22] = R.mRhythmFinalTempoConfidence.mValue;
  fwrite(H, 8, 32, f);

  const TSampleDescriptors::TFramedScalarData* FS[24] = {
    &R.mAmplitudeSilence, &R.mAmplitudePeak, &R.mAmplitudeRms, &R.mAmplitudeEnvelope,
    &R.mSpectralRms, &R.mSpectralCentroid, &R.mSpectralRolloff, &R.mSpectralSpread,
    &R.mSpectralSkewness, &R.mSpectralKurtosis, &R.mSpectralFlatness, &R.mSpectralInharmonicity,
    &R.mSpectralComplexity, &R.mSpectralContrast, &R.mSpectralFlux, &R.mF0, &R.mF0Confidence,
    &R.mFailSafeF0, &R.mTristimulus1, &R.mTristimulus2, &R.mTristimulus3, &R.mAutoCorrelation,
    &R.mRhythmComplexOnsets, &R.mRhythmPercussiveOnsets };
  for (int s = 0; s < 24; ++s) put_framed(f, *FS[s]);

  put_framed_vec(f, R.mSpectralRmsBands);
  put_framed_vec(f, R.mSpectralFlatnessBands);
  put_framed_vec(f, R.mSpectralFluxBands);
  put_framed_vec(f, R.mSpectralComplexityBands);
  put_framed_vec(f, R.mSpectralContrastBands);
  put_framed_vec(f, R.mSpectrumBands);
  put_framed_vec(f, R.mCepstrumBands);

  for (int s = 0; s < 24; ++s) put_stats(f, *FS[s]);
  put_stats_vec(f, R.mSpectralRmsBands);
  put_stats_vec(f, R.mSpectralFlatnessBands);
  put_stats_vec(f, R.mSpectralFluxBands);
  put_stats_vec(f, R.mSpectralComplexityBands);
  put_stats_vec(f, R.mSpectralContrastBands);
  put_stats_vec(f, R.mSpectrumBands);
  put_stats_vec(f, R.mCepstrumBands);
}

// High-level record (layout documented in afec_b200/layout.py, AFXH v1)
static void dump_highlevel(FILE* f, const TSampleDescriptors& R, int status)
{
  fwrite("AFXH", 1, 4, f);
  put_i32(f, status);
  if (status != 0) return;
  const TSampleClassificationDescriptors Features(R);
  const int F = R.mHighLevelPitch.mValues.Size();
  put_i32(f, F);
  put_i32(f, (int)Features.mFeatures.size());
  const double S[16] = { R.mHighLevelBaseNote.mValue, R.mHighLevelBaseNoteConfidence.mValue, R.mHighLevelPeakDb.mValue,
    R.mHighLevelRmsDb.mValue, R.mHighLevelBpm.mValue, R.mHighLevelBpmConfidence.mValue, R.mHighLevelBrightness.mValue,
    R.mHighLevelNoisiness.mValue, R.mHighLevelHarmonicity.mValue, R.mHighLevelSpectralFlatness.mValue,
    R.mHighLevelSpectralFlux.mValue, R.mHighLevelSpectralComplexity.mValue, R.mHighLevelSpectralContrast.mValue,
    R.mHighLevelSpectralInharmonicity.mValue, R.mHighLevelPitchConfidence.mValue, 0.0 };
  fwrite(S, 8, 16, f);
  for (int i = 0; i < F; ++i) put_f64(f, R.mHighLevelPitch.mValues[i]);
  for (int i = 0; i < F; ++i) put_f64(f, R.mHighLevelPeak.mValues[i]);
  for (int i = 0; i < R.mHighLevelSpectrumSignature.mValues.Size(); ++i)
    for (int b = 0; b < R.mHighLevelSpectrumSignature.mValues[i].Size(); ++b) put_f64(f, R.mHighLevelSpectrumSignature.mValues[i][b]);
  for (size_t i = 0; i < Features.mFeatures.size(); ++i) put_f64(f, Features.mFeatures[i]);
}

// -------------------------------------------------------------------------------------------------

static int usage()
{
  fprintf(stderr, "usage: afec_ref dump <hop> <out.bin> <wav>...\n"
                  "       afec_ref dumphl <hop> <out.bin> <wav>...\n"
                  "       afec_ref db <hop> <out.db> <wav>...\n"
                  "       afec_ref bench <hop> <threads> <reps> <out.db|-> <wav>...\n");
  return 2;
}

int gMain(const TList<TString>& Args)
{
  M__DisableFloatingPointAssertions

  std::vector<std::string> A;
  for (int i = 0; i < Args.Size(); ++i) A.push_back(Args[i].StdCString());
  // Args may or may not hold argv[0]; normalise so that A[0] is the mode
  if (!A.empty() && A[0] != "dump" && A[0] != "dumphl" && A[0] != "db" && A[0] != "bench") A.erase(A.begin());
  if (A.size() < 4) return usage();

  AudioTypesInit();
  CoreFileFormatsInit();
  FeatureExtractionInit();

  const std::string Mode = A[0];
  const int Hop = atoi(A[1].c_str());
  TSampleAnalyser Analyser(44100, 2048, Hop);

  if (Mode == "dump")
  {
    FILE* f = fopen(A[2].c_str(), "wb");
    if (!f) { perror("fopen"); return 1; }
    for (size_t i = 3; i < A.size(); ++i)
    {
      try {
        TSampleDescriptors R = Analyser.Analyze(
          TString(A[i].c_str(), TString::kFileSystemEncoding), TSampleDescriptors::kLowLevelDescriptors);
        dump_one(f, R, 0);
      }
      catch (const std::exception& e) {
        fprintf(stderr, "analyze failed for %s: %s\n", A[i].c_str(), e.what());
        TSampleDescriptors Empty;
        dump_one(f, Empty, 1);
      }
    }
    fclose(f);
    return 0;
  }
  else if (Mode == "dumphl")
  {
    FILE* f = fopen(A[2].c_str(), "wb");
    if (!f) { perror("fopen"); return 1; }
    for (size_t i = 3; i < A.size(); ++i)
    {
      try {
        TSampleDescriptors R = Analyser.Analyze(
          TString(A[i].c_str(), TString::kFileSystemEncoding), TSampleDescriptors::kHighLevelDescriptors);
        dump_one(f, R, 0);
        dump_highlevel(f, R, 0);
      }
      catch (const std::exception& e) {
        fprintf(stderr, "analyze failed for %s: %s\n", A[i].c_str(), e.what());
        TSampleDescriptors Empty;
        dump_one(f, Empty, 1);
        dump_highlevel(f, Empty, 1);
      }
    }
    fclose(f);
    return 0;
  }
  else if (Mode == "db")
  {
    TSqliteSampleDescriptorPool Pool(TSampleDescriptors::kLowLevelDescriptors);
    if (!Pool.Open(TString(A[2].c_str(), TString::kFileSystemEncoding))) {
      fprintf(stderr, "failed to open db\n"); return 1;
    }
    std::mutex Lock;
    for (size_t i = 3; i < A.size(); ++i)
      Analyser.Extract(TString(A[i].c_str(), TString::kFileSystemEncoding), &Pool, Lock);
    return 0;
  }
  else if (Mode == "bench")
  {
    if (A.size() < 6) return usage();
    const int Threads = atoi(A[2].c_str());
    const int Reps = atoi(A[3].c_str());
    const std::string DbPath = A[4];
    std::vector<std::string> Files(A.begin() + 5, A.end());

    TOwnerPtr<TSqliteSampleDescriptorPool> pPool;
    if (DbPath != "-") {
      pPool = TOwnerPtr<TSqliteSampleDescriptorPool>(
        new TSqliteSampleDescriptorPool(TSampleDescriptors::kLowLevelDescriptors));
      if (!pPool->Open(TString(DbPath.c_str(), TString::kFileSystemEncoding))) {
        fprintf(stderr, "failed to open db\n"); return 1;
      }
    }
    std::mutex Lock;
    const size_t Total = Files.size() * (size_t)Reps;
    std::atomic<size_t> Next(0);
    std::atomic<long long> Frames(0);
    const auto T0 = std::chrono::steady_clock::now();
    std::vector<std::thread> Pool;
    for (int t = 0; t < Threads; ++t) {
      Pool.emplace_back([&]() {
        M__DisableFloatingPointAssertions
        for (;;) {
          const size_t i = Next.fetch_add(1);
          if (i >= Total) break;
          const TString Name(Files[i % Files.size()].c_str(), TString::kFileSystemEncoding);
          if (pPool) {
            Analyser.Extract(Name, pPool, Lock);   // exactly what Crawler.cpp:716-726 runs
          } else {
            TSampleDescriptors R = Analyser.Analyze(Name, TSampleDescriptors::kLowLevelDescriptors);
            Frames += R.mAmplitudeSilence.mValues.Size();
          }
        }
      });
    }
    for (auto& t : Pool) t.join();
    const double Secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count();
    printf("{\"files\": %zu, \"threads\": %d, \"seconds\": %.6f, \"frames\": %lld}\n",
      Total, Threads, Secs, (long long)Frames.load());
    return 0;
  }
  return usage();
}

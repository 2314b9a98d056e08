// Names / order of the low-level descriptor set and the msgpack BLOB encoding.
// Reference: SampleDescriptors.cpp:20-136 (names), :154-203 (order), SampleDescriptors.h:186-230, 327-355
// (statistic suffixes), SqliteSampleDescriptorPool.cpp:601-713 (msgpack).
#include "afx_host.h"

#include <cstring>

namespace afec {

const char* const kHeaderNames[23] = {
  "file_size", "file_length", "file_sample_rate", "file_channel_count", "file_bit_depth",
  "effectve_length_48dB", "effectve_length_24dB", "effectve_length_12dB", "analyzation_offset",
  "rhythm_complex_onset_count", "rhythm_complex_onset_contrast", "rhythm_complex_onset_frequency_mean",
  "rhythm_complex_onset_strength", "rhythm_complex_tempo", "rhythm_complex_tempo_confidence",
  "rhythm_percussive_onset_count", "rhythm_percussive_onset_contrast", "rhythm_percussive_onset_frequency_mean",
  "rhythm_percussive_onset_strength", "rhythm_percussive_tempo", "rhythm_percussive_tempo_confidence",
  "rhythm_final_tempo", "rhythm_final_tempo_confidence" };

const char* const kFramedScalarNames[AFX_N_FS] = {
  "amplitude_silence", "amplitude_peak", "amplitude_rms", "amplitude_envelope",
  "spectral_rms", "spectral_centroid", "spectral_rolloff", "spectral_spread", "spectral_skewness",
  "spectral_kurtosis", "spectral_flatness", "spectral_inharmonicity", "spectral_complexity",
  "spectral_contrast", "spectral_flux", "f0", "f0_confidence", "failsafe_f0",
  "tristimulus1", "tristimulus2", "tristimulus3", "auto_correlation",
  "rhythm_complex_onsets", "rhythm_percussive_onsets" };

const char* const kFramedVectorNames[AFX_N_FV] = {
  "spectral_rms_bands", "spectral_flatness_bands", "spectral_flux_bands", "spectral_complexity_bands",
  "spectral_contrast_bands", "frequency_bands", "cepstrum_bands" };
const int kFramedVectorBands[AFX_N_FV] = { 14, 14, 14, 14, 14, 28, 14 };

const char* const kStatNames[AFX_N_STATS] = { "min", "max", "median", "mean", "gmean", "variance", "centroid",
  "spread", "skewness", "kurtosis", "flatness", "dmean", "dvariance" };

void TSampleDescriptors::Assign(const afx_file_result& r)
{
  mFrames = r.n_frames; mRhythmFrames = r.n_rhythm_frames;
  if (r.header) memcpy(mHeader, r.header, sizeof(mHeader));
  for (int s = 0; s < AFX_N_FS; ++s) {
    const int n = (s < AFX_N_FS_MAIN) ? mFrames : mRhythmFrames;
    if (r.fs[s]) mFramedScalars[s].assign(r.fs[s], r.fs[s] + n); else mFramedScalars[s].assign(n, 0.0);
  }
  for (int v = 0; v < AFX_N_FV; ++v) {
    const size_t n = (size_t)mFrames * kFramedVectorBands[v];
    if (r.fv[v]) mFramedVectors[v].assign(r.fv[v], r.fv[v] + n); else mFramedVectors[v].assign(n, 0.0);
  }
  if (r.stats) memcpy(mStats, r.stats, sizeof(mStats)); else memset(mStats, 0, sizeof(mStats));
}

// ---- msgpack --------------------------------------------------------------------------------------
static inline void pack_array_header(std::vector<unsigned char>& out, size_t n)
{
  if (n < 16) out.push_back((unsigned char)(0x90u | n));
  else if (n < 65536) { out.push_back(0xdc); out.push_back((unsigned char)(n >> 8)); out.push_back((unsigned char)n); }
  else { out.push_back(0xdd); out.push_back((unsigned char)(n >> 24)); out.push_back((unsigned char)(n >> 16)); out.push_back((unsigned char)(n >> 8)); out.push_back((unsigned char)n); }
}
static inline void pack_doubles(std::vector<unsigned char>& out, const double* v, size_t n)
{
  size_t o = out.size();
  out.resize(o + 9 * n);
  unsigned char* p = out.data() + o;
  for (size_t i = 0; i < n; ++i) {
    uint64_t u; memcpy(&u, v + i, 8);
    u = __builtin_bswap64(u);
    *p++ = 0xcb; memcpy(p, &u, 8); p += 8;
  }
}
void PackVR(std::vector<unsigned char>& out, const double* values, size_t n)
{
  out.clear(); out.reserve(5 + 9 * n);
  pack_array_header(out, n);
  pack_doubles(out, values, n);
}
void PackVVR(std::vector<unsigned char>& out, const double* values, size_t frames, size_t bands)
{
  out.clear(); out.reserve(5 + frames * (3 + 9 * bands));
  pack_array_header(out, frames);
  for (size_t f = 0; f < frames; ++f) { pack_array_header(out, bands); pack_doubles(out, values + f * bands, bands); }
}

}  // namespace afec

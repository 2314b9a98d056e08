// K8: high-level derivations that need no classification model, and the classification feature vector.
//
// Reference: TSampleAnalyser::AnalyzeHighLevelDescriptors, base note .. pitch / peak
// (Source/Crawler/FeatureExtraction/Source/SampleAnalyser.cpp:1232-1606; helpers :139-155, :420-439; aubio
// mathutils.c:535-546; TMath::Quantize, CoreTypes/Export/InlineMath.inl:581-640; TAudioMath::LinToDb(float),
// AudioTypes/Export/AudioMath.inl:38-54) and TSampleClassificationDescriptors
// (Source/SampleClassificationDescriptors.cpp:39-43, 60-63, 404-560) -- the input vector of the LightGBM models,
// whose evaluation stays on the host (north_star).
//
// Everything here is a reduction or a gather over the low-level arrays the batch already holds in HBM, so the
// `--level high` run of the reference needs no second pass over the audio: one CTA per file walks the file's
// series segments (frame_off / F) with coalesced loads; sums go through block reductions, the median of the
// confident pitches through the radix select of afx_select.cuh, the "last confident pitch" recurrence through a
// max-scan of indices.
#include "afx_select.cuh"
#include "../../include/afec_b200.h"

#define HT 128
#define HL_NTIME 48

__constant__ int c_hl_bands[14] = { 0, 1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25 };      // SampleAnalyser.cpp:1462-1465
__constant__ int c_hl_time[HL_NTIME] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24,
  25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 64, 128, 256, 512 };   // SampleClassificationDescriptors.cpp:39-43
__constant__ int c_hl_pick[7] = { 0, 1, 3, 5, 10, 11, 12 };      // min, max, mean, variance, flatness, dmean, dvariance (:118-142)

__device__ __forceinline__ double hl_freqtomidi(double freq)     // mathutils.c:535-546, smpl_t = double
{
  if (freq < 2. || freq > 100000.) return 0.;
  double midi = freq / 6.875;
  midi = log(midi) / 0.69314718055995;
  midi *= 12;
  midi -= 3;
  return midi;
}
__device__ __forceinline__ double hl_lin_to_db_f(float v)        // AudioMath.inl:38-54
{
  if (v == 1.0f) return 0.0;
  if (v > 1e-12f) return (double)(float)(log((double)v) * (20.0 / 2.302585092994045684));
  return -200.0;
}
__device__ __forceinline__ double hl_cubic(double ym1, double y0, double y1, double y2, double pos)   // SampleAnalyser.cpp:139-155
{
  const double x = pos - floor(pos), xx = x * x, xxx = xx * x;
  const double a = -0.5 * xxx + xx - 0.5 * x, b = 1.5 * xxx - 2.5 * xx + 1.0, c = -1.5 * xxx + 2.0 * xx + 0.5 * x, d = 0.5 * xxx - 0.5 * xx;
  return a * ym1 + b * y0 + c * y1 + d * y2;
}
// one merged, compressed band of the 28 frequency bands of a frame (SampleAnalyser.cpp:1476-1489)
__device__ __forceinline__ double hl_merged_band(const double* __restrict__ bands28, int b)
{
  const int s = (b >= 1) ? c_hl_bands[b - 1] + 1 : 0, e = c_hl_bands[b];
  double m = 0.0;
  for (int sb = s; sb <= e; ++sb) m += bands28[sb];
  m /= (double)(e - s + 1);
  return pow(m * 1.25, 1.0 / 6.0);
}
__device__ __forceinline__ double hl_clamp01(double w) { w = w < 1.0 ? w : 1.0; return w > 0.0 ? w : 0.0; }

// block-wide max / min of one double, sum of ints (HT threads); results broadcast
__device__ __forceinline__ double hl_block_max(double v, double* scr) { return block_max(v, scr); }
__device__ __forceinline__ double hl_block_min(double v, double* scr) { return -block_max(-v, scr); }

__global__ void __launch_bounds__(HT) k_highlevel(AfxBatchDev B, AfxParams P, AfxHighLevelDev O)
{
  __shared__ double scr[16 * 32];
  __shared__ int iscr[64];
  __shared__ int hist[256];
  __shared__ int ctl[4];
  const int tid = threadIdx.x;
  const int fi = blockIdx.x;
  const AfxFile f = B.files[fi];
  double* __restrict__ hl = O.scalars + (size_t)fi * AFX_N_HL;
  if (f.status != 0) { if (tid < AFX_N_HL) hl[tid] = 0.0; if (tid == 0) O.status[fi] = 0; return; }
  const AfxState st = B.state[fi];
  const int F = st.F;
  const size_t TF = (size_t)B.TF;
  auto fsp = [&](int s) { return B.fs + (size_t)s * TF + f.frame_off; };
  const double* __restrict__ sil = fsp(FS_AMP_SILENCE);
  const double* __restrict__ f0 = fsp(FS_F0);
  const double* __restrict__ conf = fsp(FS_F0_CONF);
  const double* __restrict__ bands28 = B.fv + (size_t)FV_BANDS28 * TF + (size_t)f.frame_off * 28;
  const double* __restrict__ header = B.header + (size_t)fi * AFX_N_HEADER;
  const double* __restrict__ stats = B.stats + (size_t)fi * AFX_N_SERIES * AFX_N_STATS;
  double* pitch = O.pitch + f.frame_off;                         // also the scratch of the confident-pitch list (not __restrict__: written and re-read)

  // ---- pass 1: sums / extrema over the audible frames (SampleAnalyser.cpp:420-439, 867) --------------------------
  double a[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };               // n, conf, rolloff, flatness, autocorr, flux, complexity, contrast, inharmonicity, -
  double cmax = -1.0e308, fmin_ = 1.0e308, fmax_ = -1.0e308;
  {
    const double* __restrict__ roll = fsp(FS_SPEC_ROLLOFF); const double* __restrict__ cent = fsp(FS_SPEC_CENTROID);
    const double* __restrict__ flat = fsp(FS_SPEC_FLATNESS); const double* __restrict__ ac = fsp(FS_AUTOCORR);
    const double* __restrict__ flux = fsp(FS_SPEC_FLUX); const double* __restrict__ cplx = fsp(FS_SPEC_COMPLEXITY);
    const double* __restrict__ contr = fsp(FS_SPEC_CONTRAST); const double* __restrict__ inh = fsp(FS_SPEC_INHARM);
    for (int i = tid; i < F; i += HT) {
      if (sil[i] == 0.0) {
        a[0] += 1.0; a[1] += conf[i]; a[2] += roll[i]; a[3] += flat[i]; a[4] += ac[i]; a[5] += flux[i]; a[6] += cplx[i]; a[7] += contr[i]; a[8] += inh[i];
        cmax = fmax(cmax, cent[i]); fmin_ = fmin(fmin_, flat[i]); fmax_ = fmax(fmax_, flat[i]);
      }
    }
  }
  block_sum<10>(a, scr);
  cmax = hl_block_max(cmax, scr); fmin_ = hl_block_min(fmin_, scr); fmax_ = hl_block_max(fmax_, scr);
  const int na = (int)a[0];
  const double dna = (double)(na > 0 ? na : 1);
  const double apcm = na ? a[1] / dna : 0.0;                                            // :1256-1260 (Mean of one value is the value)
  const double thr = (apcm >= 0.8) ? 0.8 : (apcm >= 0.5) ? 0.5 : 0.2;                   // :1262-1277
  const double fcut = (double)(P.sr / 4);
  auto confident = [&](int i) { const double hz = f0[i]; return conf[i] > thr && hz > 20.0 && hz < fcut; };

  // ---- confident pitches, compacted in frame order (:1281-1292): contiguous runs per thread + exclusive scan --------
  const int per = (F + HT - 1) / HT, i0 = tid * per, i1 = min(F, i0 + per);
  int cnt = 0;
  for (int i = i0; i < i1; ++i) cnt += confident(i) ? 1 : 0;
  int inc = cnt;
  {
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int pv = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += pv; }
    if (lane == 31) iscr[wid] = inc;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < wid; ++w) base += iscr[w];
    inc += base;
    if (tid == HT - 1) iscr[8] = inc;
    __syncthreads();
  }
  const int nc = iscr[8];
  {
    int w = inc - cnt;
    for (int i = i0; i < i1; ++i) if (confident(i)) pitch[w++] = f0[i];
  }
  __syncthreads();
  double base_note = -1.0, base_conf = 0.0;
  if (nc > 0) {                                                                         // block-uniform
    const double hz = (nc >= 2) ? block_select(pitch, nc, (nc - 1) / 2, hist, ctl) : pitch[0];     // Statistics.cpp:316-413 (lower median)
    if (hz > 20.0 && hz < fcut) base_note = hl_freqtomidi(hz);
    if (base_note > 0.0) {                                                              // :1307-1325
      double s[1] = { 0.0 };
      for (int i = tid; i < nc; i += HT) s[0] += fabs(base_note - hl_freqtomidi(pitch[i]));
      block_sum<1>(s, scr);
      const double mean = (nc >= 2) ? s[0] / (double)nc : s[0];
      double v[1] = { 0.0 };
      for (int i = tid; i < nc; i += HT) { const double d = fabs(base_note - hl_freqtomidi(pitch[i])) - mean; v[0] += d * d; }
      block_sum<1>(v, scr);
      const double sd = (nc >= 2) ? sqrt(v[0] / (double)nc) : 0.0;                      // Statistics.cpp:275-292, 304-307
      const double q = sd / 6.0;
      base_conf = apcm * (1.0 - (q < 1.0 ? q : 1.0));
    }
  }
  __syncthreads();                                         // the confident-pitch list is dead: `pitch` becomes the output series

  // ---- scalars --------------------------------------------------------------------------------------------------
  if (tid == 0) {
    hl[0] = base_note; hl[1] = base_conf;
    hl[2] = hl_lin_to_db_f((float)header[H_PEAK]);                                      // :1336-1339, mPeakValue / mRmsValue are floats
    hl[3] = hl_lin_to_db_f((float)header[H_RMS]);
    double v = header[H_FINAL_TEMPO];                                                   // :1345-1349, InlineMath.inl:625-637
    if (v > 0.0) v += 0.25; else v -= 0.25;
    hl[4] = (double)((double)(int)(v / 0.5) * 0.5);
    hl[5] = header[H_FINAL_TEMPO_CONF];
    double bright = 0.0, noisy = 0.0, harm = 0.0;
    const double flat_mean = na ? a[3] / dna : 0.0;
    if (na) {
      bright = pow(hl_clamp01(hl_freqtomidi(a[2] / dna) / 128.0 * 0.7 + hl_freqtomidi(cmax) / 128.0 * 0.3), 4.0);      // :1355-1383
      noisy = pow(hl_clamp01((1.0 - fmin_) * 0.2 + (1.0 - flat_mean) * 0.6 + (1.0 - fmax_) * 0.2), 2.0);              // :1387-1414
      const double x = 1.5 * (a[4] / dna), y = 2.0 * apcm;
      harm = pow(hl_clamp01((x < 1.0 ? x : 1.0) * 0.4 + (y < 1.0 ? y : 1.0) * 0.3 + flat_mean * 0.3), 2.0);           // :1419-1446
    }
    hl[6] = bright; hl[7] = noisy; hl[8] = harm;
    hl[9] = flat_mean; hl[10] = na ? a[5] / dna : 0.0; hl[11] = na ? a[6] / dna : 0.0; hl[12] = na ? a[7] / dna : 0.0;  // :1527-1554
    hl[13] = na ? a[8] / dna : 0.0; hl[14] = apcm; hl[15] = 0.0;
  }

  // ---- spectrum signature: 64 cubic-resampled frames of the 14 merged bands (:1449-1521) ----------------------------
  {
    double* __restrict__ sig = O.signature + (size_t)fi * (64 * 14);
    const double step = (double)F / 64.0;
    for (int o = tid; o < 64 * 14; o += HT) {
      const int i = o / 14, j = o % 14;
      double pos = 0.0;
      for (int q = 0; q < i; ++q) pos += step;                      // the reference accumulates CurrentPos (:1497-1519)
      const int ip = (int)pos, im1 = max(0, ip - 1), ip1 = min(F - 1, ip + 1), ip2 = min(F - 1, ip + 2);
      sig[o] = hl_cubic(hl_merged_band(bands28 + (size_t)im1 * 28, j), hl_merged_band(bands28 + (size_t)ip * 28, j),
                        hl_merged_band(bands28 + (size_t)ip1 * 28, j), hl_merged_band(bands28 + (size_t)ip2 * 28, j), pos);
    }
  }

  // ---- pitch series with fall-back to the last confident pitch (:1558-1598) ---------------------------------------------
  {
    // look-ahead: first audible + confident frame among 0 .. max(1, F / 4)
    int first = 0x7fffffff;
    if (F > 1) { const int lim = max(1, F / 4); for (int i = tid; i <= lim; i += HT) if (sil[i] == 0.0 && confident(i)) { first = i; break; } }
    first = block_min_i(first, iscr + 16);
    const double start = (first != 0x7fffffff) ? f0[first] : 0.0;
    // last valid index at or before every frame: per-thread runs + max-scan of the runs' last valid index
    int last = -1;
    for (int i = i0; i < i1; ++i) if (sil[i] == 0.0 && confident(i)) last = i;
    int run = last;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int pv = __shfl_up_sync(0xffffffffu, run, o); if (lane >= o) run = max(run, pv); }
    __syncthreads();
    if (lane == 31) iscr[32 + wid] = run;
    __syncthreads();
    int before = -1;                                                // last valid index in the runs of the threads before this one
    for (int w = 0; w < wid; ++w) before = max(before, iscr[32 + w]);
    { const int pv = __shfl_up_sync(0xffffffffu, run, 1); if (lane > 0) before = max(before, pv); }
    int cur = before;
    for (int i = i0; i < i1; ++i) {
      if (sil[i] == 0.0 && confident(i)) cur = i;
      pitch[i] = hl_freqtomidi(cur >= 0 ? f0[cur] : start);
    }
  }

  // ---- classification features (SampleClassificationDescriptors.cpp:404-560) ----------------------------------------------
  {
    double* __restrict__ feat = O.features + (size_t)fi * AFX_HL_FEATURES;
    const double* __restrict__ pad = O.silence_pad;
    const int tser[6] = { FS_SPEC_RMS, FS_SPEC_FLATNESS, FS_SPEC_FLUX, FS_SPEC_CONTRAST, FS_SPEC_COMPLEXITY, FS_F0_CONF };
    const int vbase[6] = { 24, 38, 52, 66, 80, 122 };               // stats rows: rms / flatness / flux / complexity / contrast sub-bands, cepstrum
    int bad = 0;
    for (int n = tid; n < AFX_HL_FEATURES; n += HT) {
      double v;
      int r = n;
      if (r < 14 * HL_NTIME) { const int b = r / HL_NTIME, tf = c_hl_time[r % HL_NTIME]; v = (tf < F) ? hl_merged_band(bands28 + (size_t)tf * 28, b) : pad[b]; }
      else if ((r -= 14 * HL_NTIME) < 6 * HL_NTIME) { const int k = r / HL_NTIME, tf = c_hl_time[r % HL_NTIME]; v = (tf < F) ? fsp(tser[k])[tf] : pad[14 + k]; }
      else if ((r -= 6 * HL_NTIME) < 6 * 7) v = stats[tser[r / 7] * AFX_N_STATS + c_hl_pick[r % 7]];
      else if ((r -= 6 * 7) < 6 * 14 * 7) { const int k = r / 98, b = (r % 98) / 7; v = stats[(vbase[k] + b) * AFX_N_STATS + c_hl_pick[r % 7]]; }
      else if ((r -= 6 * 14 * 7) < HL_NTIME) { const int tf = c_hl_time[r]; v = (tf < F) ? fsp(FS_AMP_RMS)[tf] : pad[20]; }
      else if ((r -= HL_NTIME) < 7) v = stats[FS_AMP_RMS * AFX_N_STATS + c_hl_pick[r]];
      else if ((r -= 7) < 7) v = stats[FS_AMP_SILENCE * AFX_N_STATS + c_hl_pick[r]];
      else if ((r -= 7) < 7) {
        const int hidx[7] = { H_RC_TEMPO_CONF, H_RP_TEMPO_CONF, H_RC_CONTRAST, H_RP_CONTRAST, H_RC_STRENGTH, H_RP_STRENGTH, H_EFF12 };
        v = header[hidx[r]];
      } else v = stats[FS_SPEC_RMS * AFX_N_STATS + 3];              // padding to the time-series width: spectral_rms mean (:536-547)
      feat[n] = v;
      if (v != v || fabs(v) > 1.7976931348623157e308) bad = 1;      // NaN / Inf: the reference throws, the file fails to analyse
    }
    bad = block_sum_i(bad, iscr + 48);
    if (tid == 0) O.status[fi] = bad ? 1 : 0;
  }
}

void afx_launch_highlevel(const AfxParams& P, const AfxBatchDev& B, const AfxHighLevelDev& O, cudaStream_t s, long long* launches)
{
  if (B.n_files <= 0) return;
  k_highlevel<<<B.n_files, HT, 0, s>>>(B, P, O); ++*launches;
}

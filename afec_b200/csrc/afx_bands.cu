// K4: band projections of the magnitude spectrum: a selection kernel (one warp per sub-band role and frame)
// and an epilogue kernel (one lane per sub-band); no sort, no block-wide barrier.
//
//   * 14 sub-bands (SampleAnalyser.cpp:2067-2260): rms, flatness (dB scaled), flux (Pearson correlation
//     with the previous frame), complexity (strict local maxima above 0.25 x band max) and contrast
//     -(peakMean / valleyMean)^(1 / ln(mean)) from the sorted band; spectral_contrast = mean of the 14
//   * 28 "frequency bands" (SampleAnalyser.cpp:2007-2048): sum of squared magnitudes
//   * 14 cepstrum bands (SampleAnalyser.cpp:2052-2063; LibXtract vector.c:350-391): 14 triangular mel
//     filters -> log -> unnormalised DCT-II, each filter evaluated on its non-zero support only (the
//     filters cover bins 1..358: they are laid over 512 of the 1024 bins -- quirk).
//
// Each sub-band belongs to one warp that keeps the band in registers.  The reference sorts every band to
// average its lowest / highest 30 %; a sum over the k smallest values only needs the k-th order statistic
// v:  sum = sum_{x < v} x + (k - #{x < v}) v  (ties carry the same value, so the result is that of the
// sort).  Order statistics come from an exact MSB-first bisection on the 64-bit patterns of the (non-negative)
// magnitudes with 32-bit integer compares and warp vote/reduce -- the FP64 pipe is left to the arithmetic.
// Bands of <= 32 bins rank their elements against each other with shuffles instead.
#include "afx_common.cuh"
#include <algorithm>

#define BT 256

__device__ __forceinline__ double warp_sum_d(double v) { return warp_sum(v); }

// raw per-band sums handed from the band's warp to the epilogue lane
struct BandRaw { double s1, s2, s11, s12, s22, ls, x0, lo_sum, hi_sum, cplx; };   // 10 doubles
// ls / x0: for n >= 2 the band's log-sum travels as (product of the mantissas, sum of the exponents) and the ONE log per
// band is taken by the epilogue lane; for n == 1 x0 is the band's only value (TStatistics::GeometricMean returns it)
#define BR_STRIDE 154      // doubles per frame: 14 x BandRaw + 14 mel energies

// per-band epilogue (one lane per band, all 14 in lock step so the pow / log / exp chains run once)
__device__ __forceinline__ double band_write(AfxBatchDev& B, size_t TF, int slot, int b, int n, int nei, const BandRaw& r)
{
  const double dn = (double)n;
  const double mean = (n >= 2) ? r.s1 / dn : r.s1;          // TStatistics::Mean, Statistics.cpp:249-266
  const double gmean = (n >= 2) ? exp((log(r.ls) + r.x0 * 0.693147180559945309417) / dn) : r.x0;    // TStatistics::GeometricMean :417-455
  const size_t o = (size_t)slot * 14 + b;
  B.fv[(size_t)FV_RMS * TF + o] = sqrt(r.s11 / dn);
  B.fv[(size_t)FV_FLATNESS * TF + o] = flatness_db(mean, gmean);
  const double m1 = r.s1 / dn, m2 = r.s2 / dn;
  const double den2 = (r.s11 - m1 * m1 * dn) * (r.s22 - m2 * m2 * dn);
  const double num = r.s12 - (m1 * m2 * dn);
  B.fv[(size_t)FV_FLUX * TF + o] = (fabs(den2) > (double)1e-12f) ? num / sqrt(den2) : 0.0;
  B.fv[(size_t)FV_COMPLEXITY * TF + o] = r.cplx;
  const double valley = r.lo_sum / nei + 1e-30, peak = r.hi_sum / nei + 1e-30;      // SampleAnalyser.cpp:2199-2232
  const double c = -1.0 * pow(peak / valley, 1.0 / log(mean + 1e-30));
  B.fv[(size_t)FV_CONTRAST * TF + o] = c;
  return c;
}

// One sub-band on one warp, C = ceil(n / 32) elements per lane.
template <int C>
__device__ __forceinline__ void subband(const AfxParams& P, int b, const double* __restrict__ g,
                                        const double* __restrict__ gp, int lane, BandRaw* raw)
{
  const int s0 = P.band14_start[b], n = P.band14_n[b], nei = P.band14_nei[b];
  double x[C]; unsigned hi[C], lo[C];
  double s1 = 0, s2 = 0, s11 = 0, s12 = 0, s22 = 0, mx = 0.0, mant = 1.0; int ex = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int k = lane + 32 * c;
    const bool valid = k < n;
    const double xv = valid ? g[s0 + k] : 0.0, yv = valid ? gp[s0 + k] : 0.0;
    x[c] = xv;
    const unsigned long long u = valid ? (unsigned long long)__double_as_longlong(xv) : 0xffffffffffffffffull;
    hi[c] = (unsigned)(u >> 32); lo[c] = (unsigned)u;
    if (valid) {
      s12 += xv * yv; s1 += xv; s11 += xv * xv; s2 += yv; s22 += yv * yv;
      mx = fmax(mx, xv);
      const double v = fabs(xv) + 1e-20;               // Statistics.cpp:417-455: product with the exponents peeled off
      const int hw = __double2hiint(v);                // (C <= 9 factors >= 1/2 per lane, >= 2^-288 per warp: no rescue needed)
      ex += ((hw >> 20) & 0x7ff) - 1022;
      mant *= __hiloint2double((hw & 0x800fffff) | 0x3fe00000, __double2loint(v));
    }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2); s11 = warp_sum(s11); s12 = warp_sum(s12); s22 = warp_sum(s22);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mant *= __shfl_xor_sync(0xffffffffu, mant, o);
  ex = __reduce_add_sync(0xffffffffu, ex);
  mx = warp_max(mx);
  const double thr = mx * 0.25;
  int cplx = 0;
  if (thr > 0.0) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int k = lane + 32 * c, q = s0 + k;
      if (k < n && x[c] > thr && q > 0 && q < AFX_NBIN - 1 && x[c] > g[q - 1] && x[c] > g[q + 1]) ++cplx;
    }
  }
  cplx = __reduce_add_sync(0xffffffffu, cplx);

  double lo_sum = 0.0, hi_sum = 0.0;
  if (C == 1) {
    // rank by shuffles: #less and #less-or-equal of every element; the element whose interval holds rank r is the r-th
    int less = 0, leq = 0;
    for (int j = 0; j < n; ++j) {
      const unsigned oh = __shfl_sync(0xffffffffu, hi[0], j), ol = __shfl_sync(0xffffffffu, lo[0], j);
      const bool lt = (oh < hi[0]) || (oh == hi[0] && ol < lo[0]);
      const bool eq = (oh == hi[0]) && (ol == lo[0]);
      less += lt ? 1 : 0; leq += (lt || eq) ? 1 : 0;
    }
    const bool valid = lane < n;
    const int r1 = nei - 1, r2 = n - nei;
    const unsigned m1 = __ballot_sync(0xffffffffu, valid && less <= r1 && r1 < leq);
    const unsigned m2 = __ballot_sync(0xffffffffu, valid && less <= r2 && r2 < leq);
    const int l1 = __ffs(m1) - 1, l2 = __ffs(m2) - 1;
    const double v1 = __shfl_sync(0xffffffffu, x[0], l1), v2 = __shfl_sync(0xffffffffu, x[0], l2);
    const int less1 = __shfl_sync(0xffffffffu, less, l1), leq2 = __shfl_sync(0xffffffffu, leq, l2);
    lo_sum = warp_sum((valid && x[0] < v1) ? x[0] : 0.0) + (double)(nei - less1) * v1;
    hi_sum = warp_sum((valid && x[0] > v2) ? x[0] : 0.0) + (double)(nei - (n - leq2)) * v2;
  } else {
    // Exact MSB-first bisection on the 64-bit patterns.  Select 1 looks for a threshold with exactly `nei`
    // elements below it (then the sum of those elements IS the sum of the nei smallest), select 2 for one with
    // exactly n - nei below it; each stops as soon as a trial splits the band that way, which takes about
    // log2(spread / gap) steps once the bits common to the whole band are skipped.  Only when equal values
    // straddle the split does a select run to the last bit; it then ends on the order statistic v of rank
    // r (0-based) and the sum is  sum_{x < v} x + (k - #{x < v}) v.
    const int r1 = nei - 1, r2 = n - nei;
    unsigned mnh = 0xffffffffu, mxh = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) if (lane + 32 * c < n) { mnh = min(mnh, hi[c]); mxh = max(mxh, hi[c]); }
    mnh = __reduce_min_sync(0xffffffffu, mnh); mxh = __reduce_max_sync(0xffffffffu, mxh);
    const int top = 31 - __clz((mnh ^ mxh) | 1u);          // highest bit in which the high words differ (0 if equal)
    const unsigned common = (top >= 31) ? 0u : (mxh & ~((2u << top) - 1u));
    unsigned p1h = common, p2h = common, p1l = 0, p2l = 0;
    bool done1 = false, done2 = false;                       // exact split found: threshold = (t?h, t?l)
    unsigned t1h = 0, t1l = 0, t2h = 0, t2l = 0;
    for (int bit = top; bit >= 0 && !(done1 && done2); --bit) {
      const unsigned a1 = p1h | (1u << bit), a2 = p2h | (1u << bit);
      int c1 = 0, c2 = 0;
#pragma unroll
      for (int c = 0; c < C; ++c) { c1 += (hi[c] < a1) ? 1 : 0; c2 += (hi[c] < a2) ? 1 : 0; }
      c1 = __reduce_add_sync(0xffffffffu, c1); c2 = __reduce_add_sync(0xffffffffu, c2);
      if (!done1) { if (c1 == nei) { done1 = true; t1h = a1; t1l = 0; } else if (c1 <= r1) p1h = a1; }
      if (!done2) { if (c2 == r2) { done2 = true; t2h = a2; t2l = 0; } else if (c2 <= r2) p2h = a2; }
    }
    if (!(done1 && done2)) {
      int b1 = 0, b2 = 0;
#pragma unroll
      for (int c = 0; c < C; ++c) { b1 += (hi[c] < p1h) ? 1 : 0; b2 += (hi[c] < p2h) ? 1 : 0; }
      b1 = __reduce_add_sync(0xffffffffu, b1); b2 = __reduce_add_sync(0xffffffffu, b2);
      for (int bit = 31; bit >= 0 && !(done1 && done2); --bit) {
        const unsigned a1 = p1l | (1u << bit), a2 = p2l | (1u << bit);
        int c1 = 0, c2 = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          c1 += (hi[c] == p1h && lo[c] < a1) ? 1 : 0;
          c2 += (hi[c] == p2h && lo[c] < a2) ? 1 : 0;
        }
        c1 = b1 + __reduce_add_sync(0xffffffffu, c1); c2 = b2 + __reduce_add_sync(0xffffffffu, c2);
        if (!done1) { if (c1 == nei) { done1 = true; t1h = p1h; t1l = a1; } else if (c1 <= r1) p1l = a1; }
        if (!done2) { if (c2 == r2) { done2 = true; t2h = p2h; t2l = a2; } else if (c2 <= r2) p2l = a2; }
      }
    }
    // thresholds as 64-bit patterns: an exact split, else the order statistic itself
    const unsigned long long T1 = done1 ? (((unsigned long long)t1h << 32) | t1l) : (((unsigned long long)p1h << 32) | p1l);
    const unsigned long long T2 = done2 ? (((unsigned long long)t2h << 32) | t2l) : (((unsigned long long)p2h << 32) | p2l);
    int nl = 0, nge = 0, ngt = 0; double sl = 0.0, sge = 0.0, sgt = 0.0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const unsigned long long u = ((unsigned long long)hi[c] << 32) | lo[c];
      const bool valid = (lane + 32 * c) < n;
      if (valid && u < T1) { ++nl; sl += x[c]; }
      if (valid && u >= T2) { ++nge; sge += x[c]; }
      if (valid && u > T2) { ++ngt; sgt += x[c]; }
    }
    nl = __reduce_add_sync(0xffffffffu, nl); nge = __reduce_add_sync(0xffffffffu, nge); ngt = __reduce_add_sync(0xffffffffu, ngt);
    sl = warp_sum(sl); sge = warp_sum(sge); sgt = warp_sum(sgt);
    lo_sum = done1 ? sl : sl + (double)(nei - nl) * __longlong_as_double((long long)T1);
    hi_sum = done2 ? sge : sgt + (double)(nei - ngt) * __longlong_as_double((long long)T2);
    (void)nge;
  }
  const double x0 = __shfl_sync(0xffffffffu, x[0], 0);
  if (lane == 0) {
    BandRaw& r = raw[b];
    r.s1 = s1; r.s2 = s2; r.s11 = s11; r.s12 = s12; r.s22 = s22; r.ls = mant; r.x0 = (n >= 2) ? (double)ex : x0; r.lo_sum = lo_sum; r.hi_sum = hi_sum; r.cplx = (double)cplx;
  }
}

__device__ __forceinline__ void mel_energy(const AfxParams& P, int q, const double* __restrict__ g, int lane, double* lg)
{
  const double* row = P.t.mel + (size_t)q * AFX_NBIN;
  double e = 0.0;
  for (int k = P.mel_lo[q] + lane; k <= P.mel_hi[q]; k += 32) e += g[k] * __ldg(row + k);
  e = warp_sum(e);
  if (lane == 0) lg[q] = e;
}

__device__ __forceinline__ void bands28(AfxBatchDev& B, const AfxParams& P, size_t TF, int slot, int b0, int b1,
                                        const double* __restrict__ g, int lane)
{
  for (int b = b0; b < b1; ++b) {
    double s = 0.0;
    for (int k = P.band28_s[b] + lane; k < P.band28_e[b]; k += 32) { const double m = g[k]; s += m * m; }
    s = warp_sum(s);
    if (lane == 0) B.fv[(size_t)FV_BANDS28 * TF + (size_t)slot * 28 + b] = s;
  }
}

// Phase A: grid (frames / 8, roles).  All warps of a CTA run the SAME role (same code, same duration) on 8
// different frames and never synchronise; the raw sums go to a per-frame scratch record in global memory.
// Roles 4..7 own the large sub-bands: their loads are issued in one batch at the top of subband().  Roles 0..3 run
// many short loops over the row (nine small sub-bands, mel filters, the 28 bands): there every warp first copies its
// frame's magnitude row to shared memory with all loads in flight at once -- as dependent global loads those loops
// were 44 % long-scoreboard stalls.
template <int MINB>
__global__ void __launch_bounds__(BT, MINB) k_bands_a_big(AfxBatchDev B, AfxParams P)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rel = blockIdx.x * 8 + wid;
  if (rel >= B.g_slots) return;
  const int slot = B.slot0 + rel;
  const int fi = B.slot_file[slot];
  const int t = slot - B.files[fi].frame_off;
  if (B.files[fi].status != 0 || t >= B.state[fi].F) return;
  const double* __restrict__ g = B.mag + (size_t)rel * AFX_NBIN;
  const double* __restrict__ gp = (t > 0) ? g - AFX_NBIN : g;              // SampleAnalyser.cpp:936-940
  BandRaw* raw = reinterpret_cast<BandRaw*>(B.bandraw + (size_t)rel * BR_STRIDE);
  double* lg = B.bandraw + (size_t)rel * BR_STRIDE + 140;
  switch (blockIdx.y) {
    case 3: subband<9>(P, 13, g, gp, lane, raw); break;
    case 2: subband<5>(P, 12, g, gp, lane, raw); break;
    case 1: subband<3>(P, 11, g, gp, lane, raw); mel_energy(P, 12, g, lane, lg); mel_energy(P, 13, g, lane, lg); break;
    default: subband<2>(P, 10, g, gp, lane, raw); subband<2>(P, 9, g, gp, lane, raw); break;
  }
}

// One launch per role; NSTAGE = how much of the row the role reads (role 0: sub-bands 0..1 + the lower 14 of the 28
// bands end below bin 128; roles 2 / 3: sub-bands 5..8 and the mel filters end below 384; role 1 takes the upper 14 bands
// up to bin 1024).  A short stage leaves room for more CTAs per SM -- these roles are latency bound.
template <int NSTAGE>
__global__ void __launch_bounds__(BT) k_bands_a_small(AfxBatchDev B, AfxParams P, int role)
{
  extern __shared__ __align__(16) double srow[];            // [8][NSTAGE]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rel = blockIdx.x * 8 + wid;
  if (rel >= B.g_slots) return;
  const int slot = B.slot0 + rel;
  const int fi = B.slot_file[slot];
  const int t = slot - B.files[fi].frame_off;
  if (B.files[fi].status != 0 || t >= B.state[fi].F) return;
  const size_t TF = (size_t)B.TF;
  const double* __restrict__ gg = B.mag + (size_t)rel * AFX_NBIN;
  const double* __restrict__ gp = (t > 0) ? gg - AFX_NBIN : gg;            // SampleAnalyser.cpp:936-940
  double* g = srow + wid * NSTAGE;
  {
    const double2* src = reinterpret_cast<const double2*>(gg);
    double2 tmp[NSTAGE / 64];
#pragma unroll
    for (int q = 0; q < NSTAGE / 64; ++q) tmp[q] = src[lane + 32 * q];
#pragma unroll
    for (int q = 0; q < NSTAGE / 64; ++q) reinterpret_cast<double2*>(g)[lane + 32 * q] = tmp[q];
  }
  __syncwarp();
  BandRaw* raw = reinterpret_cast<BandRaw*>(B.bandraw + (size_t)rel * BR_STRIDE);
  double* lg = B.bandraw + (size_t)rel * BR_STRIDE + 140;
  // roles 0..3: the nine bands of <= 32 bins (one code instance, looped), the mel energies and the 28 bands
  const int first = (role == 3) ? 7 : (role == 2) ? 5 : (role == 1) ? 2 : 0;
  const int last = (role == 3) ? 8 : (role == 2) ? 6 : (role == 1) ? 4 : 1;
  for (int b = last; b >= first; --b) subband<1>(P, b, g, gp, lane, raw);
  if (role >= 2) { for (int q = (role == 3 ? 8 : 0); q < (role == 3 ? 12 : 8); ++q) mel_energy(P, q, g, lane, lg); }
  else bands28(B, P, TF, slot, role == 1 ? 14 : 0, role == 1 ? 28 : 14, g, lane);
}

// Phase B: 16 lanes per frame; lane j < 14 finishes sub-band j (the pow / log / exp chains run on full warps)
// and takes the log of mel energy j; the DCT (vector.c:372-391) and the mean contrast go through shuffles.
__global__ void __launch_bounds__(BT) k_bands_b(AfxBatchDev B, AfxParams P)
{
  const int j = threadIdx.x & 15;
  const int rel = (blockIdx.x * BT + threadIdx.x) >> 4;
  const bool in_range = rel < B.g_slots;
  const int slot = B.slot0 + (in_range ? rel : 0);
  const int fi = B.slot_file[slot];
  const AfxFile f = B.files[fi];
  const int t = slot - f.frame_off;
  const bool live = in_range && f.status == 0 && t < B.state[fi].F;
  const size_t TF = (size_t)B.TF;
  double c = 0.0, lg = 0.0;
  if (live && j < 14) {
    const BandRaw r = reinterpret_cast<const BandRaw*>(B.bandraw + (size_t)rel * BR_STRIDE)[j];
    c = band_write(B, TF, slot, j, P.band14_n[j], P.band14_nei[j], r);
    const double e = B.bandraw[(size_t)rel * BR_STRIDE + 140 + j];
    lg = log(e < 2e-42 ? 2e-42 : e);                       // XTRACT_LOG_LIMIT
  }
  double a = 0.0, csum = 0.0;
#pragma unroll
  for (int m = 0; m < 14; ++m) {
    const double lm = __shfl_sync(0xffffffffu, lg, m, 16);
    const double cm = __shfl_sync(0xffffffffu, c, m, 16);
    // separate rn multiply and add in the reference's m order: an all-equal input (silent frame) then cancels to
    // the same last-bit residue as vector.c:381-386 instead of a different one
    if (j < 14) a = __dadd_rn(a, __dmul_rn(lm, __ldg(P.t.dct + j * 14 + m)));
    csum += cm;
  }
  if (live && j < 14) B.fv[(size_t)FV_CEPSTRUM * TF + (size_t)slot * 14 + j] = a;
  if (live && j == 0) B.fs[(size_t)FS_SPEC_CONTRAST * TF + slot] = csum / 14.0;
}

void afx_launch_bands(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  // 6 CTAs per SM (40 registers, a few spilled words): these roles wait on global loads, residency beats registers
  // (measured 4 / 5 / 6 CTAs: 4.98 / 4.82 / 4.74 ms per 248k frames)
  k_bands_a_big<6><<<dim3((B.g_slots + 7) / 8, 4), BT, 0, s>>>(B, P); ++*launches;
  cudaFuncSetAttribute(k_bands_a_small<AFX_NBIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * AFX_NBIN * (int)sizeof(double));   // per device, see afx_pitch.cu
  for (int role = 0; role < 4; ++role) {
    // highest bin the role reads (+1: the complexity test looks at a bin's neighbours)
    const int first = (role == 3) ? 7 : (role == 2) ? 5 : (role == 1) ? 2 : 0, last = (role == 3) ? 8 : (role == 2) ? 6 : (role == 1) ? 4 : 1;
    int need = 0;
    for (int b = first; b <= last; ++b) need = std::max(need, P.band14_start[b] + P.band14_n[b] + 2);
    if (role >= 2) { for (int q = (role == 3 ? 8 : 0); q < (role == 3 ? 12 : 8); ++q) need = std::max(need, P.mel_hi[q] + 1); }
    else for (int b = (role == 1 ? 14 : 0); b < (role == 1 ? 28 : 14); ++b) need = std::max(need, P.band28_e[b]);
    const dim3 grid((B.g_slots + 7) / 8);
    if (need <= 128) k_bands_a_small<128><<<grid, BT, 8 * 128 * sizeof(double), s>>>(B, P, role);
    else if (need <= 384) k_bands_a_small<384><<<grid, BT, 8 * 384 * sizeof(double), s>>>(B, P, role);
    else k_bands_a_small<AFX_NBIN><<<grid, BT, 8 * AFX_NBIN * sizeof(double), s>>>(B, P, role);
    ++*launches;
  }
  k_bands_b<<<(B.g_slots * 16 + BT - 1) / BT, BT, 0, s>>>(B, P); ++*launches;
}

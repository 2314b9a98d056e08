// K4: band projections of the magnitude spectrum.
//
//   * 14 sub-bands (SampleAnalyser.cpp:2067-2260): rms, flatness (dB scaled), flux (Pearson correlation
//     with the previous frame), complexity (strict local maxima above 0.25 x band max) and contrast
//     -(peakMean / valleyMean)^(1 / ln(mean)) from the sorted band; spectral_contrast = mean of the 14
//   * 28 "frequency bands" (SampleAnalyser.cpp:2007-2048): sum of squared magnitudes
//   * 14 cepstrum bands (SampleAnalyser.cpp:2052-2063; LibXtract vector.c:350-391): 14 triangular mel
//     filters -> log -> unnormalised DCT-II, each filter evaluated on its non-zero support only (the
//     filters cover bins 1..358: they are laid over 512 of the 1024 bins -- quirk).
//
// The reference sorts every band to average its lowest / highest 30 %; a sum over the k smallest values only needs
// the k-th order statistic v:  sum = sum_{x < v} x + (k - #{x < v}) v  (ties carry the same value, so the result is
// that of the sort).  For the five large bands the order statistics come from an exact MSB-first bisection on the
// 64-bit patterns of the (non-negative) magnitudes with 32-bit integer compares and warp vote / reduce; the nine
// bands of <= 32 bins are ranked per lane.  See "Round-2 schedule" below for how the work is laid over the GPU.
#include "afx_common.cuh"
#include <algorithm>

#define BT 256

// raw per-band sums handed to the per-band epilogue
struct BandRaw { double s1, s2, s11, s12, s22, ls, x0, lo_sum, hi_sum, cplx; };   // 10 doubles
// ls / x0: for n >= 2 the band's log-sum travels as (product of the mantissas, sum of the exponents) and ONE log per
// band is taken in the epilogue; for n == 1 x0 is the band's only value (TStatistics::GeometricMean returns it)

// per-band epilogue
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

// ---------------------------------------------------------------------------------------------------------
// Round-2 schedule: two kernels instead of seven launches.
//
//   k_bands_select  warp per (frame, LARGE sub-band: the five with more than 32 bins): only what needs the whole band
//                   at once -- the order statistics behind the contrast (exact bisection, as subband<C> above) and the
//                   peak count against 0.25 x band max.  3 doubles per band go to the per-frame scratch record.
//   k_bands_lane    LANE per frame, 32 consecutive frame slots per warp: everything that is a running sum over the bins
//                   of a frame -- the five correlation sums, log-sum and maximum of all 14 sub-bands, the 28 frequency
//                   bands, the 14 mel energies -- walks the row once, bin by bin, with NO cross-lane reduction at all
//                   (the first schedule spent most of its instructions in warp reductions of 7 sums x 14 bands); the
//                   nine sub-bands of <= 32 bins keep their values in a per-lane array and are ranked there; the
//                   per-band epilogue (log / exp / pow chains) and the DCT run on all 32 lanes.  The magnitude rows
//                   reach the lanes through a transposed shared-memory tile (32 bins x 33 frames: the extra column is
//                   the frame before the warp's first, for the flux of lane 0), loaded with coalesced 256-byte reads.
//                   Sums run in bin order -- the reference's own order.
#define BR2_STRIDE 16       // doubles per frame: 5 large bands x (lo_sum, hi_sum, complexity) + pad
#define BIG0 9              // first large sub-band (41, 61, 96, 148, 287 bins)

// order statistics + complexity of one large sub-band on one warp (C = ceil(n / 32) elements per lane)
template <int C>
__device__ __forceinline__ void subband_select(const AfxParams& P, int b, const double* __restrict__ g, int lane, double* __restrict__ out3)
{
  const int s0 = P.band14_start[b], n = P.band14_n[b], nei = P.band14_nei[b];
  double x[C]; unsigned hi[C], lo[C];
  double mx = 0.0;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int k = lane + 32 * c;
    const bool valid = k < n;
    const double xv = valid ? g[s0 + k] : 0.0;
    x[c] = xv;
    const unsigned long long u = valid ? (unsigned long long)__double_as_longlong(xv) : 0xffffffffffffffffull;
    hi[c] = (unsigned)(u >> 32); lo[c] = (unsigned)u;
    mx = fmax(mx, xv);
  }
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
  // exact MSB-first bisection on the 64-bit patterns (see subband<C>)
  const int r1 = nei - 1, r2 = n - nei;
  unsigned mnh = 0xffffffffu, mxh = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) if (lane + 32 * c < n) { mnh = min(mnh, hi[c]); mxh = max(mxh, hi[c]); }
  mnh = __reduce_min_sync(0xffffffffu, mnh); mxh = __reduce_max_sync(0xffffffffu, mxh);
  const int top = 31 - __clz((mnh ^ mxh) | 1u);
  const unsigned common = (top >= 31) ? 0u : (mxh & ~((2u << top) - 1u));
  unsigned p1h = common, p2h = common, p1l = 0, p2l = 0;
  bool done1 = false, done2 = false;
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
  const unsigned long long T1 = done1 ? (((unsigned long long)t1h << 32) | t1l) : (((unsigned long long)p1h << 32) | p1l);
  const unsigned long long T2 = done2 ? (((unsigned long long)t2h << 32) | t2l) : (((unsigned long long)p2h << 32) | p2l);
  int nl = 0, ngt = 0; double sl = 0.0, sge = 0.0, sgt = 0.0;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const unsigned long long u = ((unsigned long long)hi[c] << 32) | lo[c];
    const bool valid = (lane + 32 * c) < n;
    if (valid && u < T1) { ++nl; sl += x[c]; }
    if (valid && u >= T2) sge += x[c];
    if (valid && u > T2) { ++ngt; sgt += x[c]; }
  }
  nl = __reduce_add_sync(0xffffffffu, nl); ngt = __reduce_add_sync(0xffffffffu, ngt);
  sl = warp_sum(sl); sge = warp_sum(sge); sgt = warp_sum(sgt);
  if (lane == 0) {
    out3[0] = done1 ? sl : sl + (double)(nei - nl) * __longlong_as_double((long long)T1);
    out3[1] = done2 ? sge : sgt + (double)(nei - ngt) * __longlong_as_double((long long)T2);
    out3[2] = (double)cplx;
  }
}

__global__ void __launch_bounds__(BT, 6) k_bands_select(AfxBatchDev B, AfxParams P)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rel = blockIdx.x * 8 + wid;
  if (rel >= B.g_slots) return;
  const int slot = B.slot0 + rel;
  const int fi = B.slot_file[slot];
  const int t = slot - B.files[fi].frame_off;
  if (B.files[fi].status != 0 || t >= B.state[fi].F) return;
  const double* __restrict__ g = B.mag + (size_t)rel * AFX_NBIN;
  double* out = B.bandraw + (size_t)rel * BR2_STRIDE;
  switch (blockIdx.y) {                         // heaviest band first in launch order (grid y is the slow index)
    case 0: subband_select<9>(P, 13, g, lane, out + 12); break;
    case 1: subband_select<5>(P, 12, g, lane, out + 9); break;
    case 2: subband_select<3>(P, 11, g, lane, out + 6); break;
    default: subband_select<2>(P, 10, g, lane, out + 3); subband_select<2>(P, 9, g, lane, out); break;
  }
}

#define BLW 4               // warps per CTA of k_bands_lane (128 frame slots)
__global__ void __launch_bounds__(BLW * 32) k_bands_lane(AfxBatchDev B, AfxParams P)
{
  __shared__ double tiles[BLW][32][33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rel0 = (blockIdx.x * BLW + wid) * 32;
  if (rel0 >= B.g_slots) return;                // warp-uniform
  double (*tile)[33] = tiles[wid];
  const int rel = rel0 + lane;
  const bool in_range = rel < B.g_slots;
  const int slot = B.slot0 + (in_range ? rel : rel0);
  const int fi = B.slot_file[slot];
  const int t = slot - B.files[fi].frame_off;
  const bool live = in_range && B.files[fi].status == 0 && t < B.state[fi].F;
  const bool has_prev = t > 0;                  // SampleAnalyser.cpp:936-940: a file's first frame correlates with itself
  const size_t TF = (size_t)B.TF;
  const double* __restrict__ mag = B.mag;
  const int last_row = B.g_slots - 1;

  // running state of the walk over the bins
  int b = 0, bs = P.band14_start[0], be = bs + P.band14_n[0];
  double s1 = 0, s2 = 0, s11 = 0, s12 = 0, s22 = 0, mx = 0, mant = 1.0; int ex = 0;
  double vals[34];                              // a small band with its two neighbours: vals[0] = x[bs - 1], vals[1 + i] = x[bs + i]
  double xprev = 0.0;
  int b28 = 0; double acc28 = 0.0;
  double mel[14];
#pragma unroll
  for (int q = 0; q < 14; ++q) mel[q] = 0.0;
  int qlo = 0;                                  // first mel filter whose support has not ended
  double csum = 0.0;

  for (int kb = 0; kb < AFX_NBIN; kb += 32) {
    // the tile: rows rel0 - 1 .. rel0 + 31, bins kb .. kb + 31; lane = bin on the way in, lane = frame on the way out
#pragma unroll 11
    for (int c = 0; c < 33; ++c) {
      const int rc = min(max(rel0 - 1 + c, 0), last_row);
      tile[lane][c] = mag[(size_t)rc * AFX_NBIN + kb + lane];
    }
    __syncwarp();
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
      const int k = kb + j;                     // uniform
      const double x = tile[j][lane + 1];
      const double y = has_prev ? tile[j][lane] : x;
      // ---- 28 frequency bands (SampleAnalyser.cpp:2007-2048) ----
      while (b28 < 28 && k >= P.band28_e[b28]) {
        if (live) B.fv[(size_t)FV_BANDS28 * TF + (size_t)slot * 28 + b28] = acc28;
        acc28 = 0.0; ++b28;
      }
      if (b28 < 28 && k >= P.band28_s[b28]) acc28 = fma(x, x, acc28);
      // ---- mel energies on the filters' supports (vector.c:350-391) ----
      while (qlo < 14 && k > P.mel_hi[qlo]) ++qlo;
      for (int q = qlo; q < 14 && P.mel_lo[q] <= k; ++q) mel[q] = fma(x, __ldg(P.t.mel + (size_t)q * AFX_NBIN + k), mel[q]);
      // ---- 14 sub-bands (SampleAnalyser.cpp:2067-2260) ----
      if (b < 14 && k == be) {
        // band b is complete and x is its right neighbour
        const int n = be - bs, nei = P.band14_nei[b];
        BandRaw r;
        r.s1 = s1; r.s2 = s2; r.s11 = s11; r.s12 = s12; r.s22 = s22; r.ls = mant; r.x0 = (n >= 2) ? (double)ex : vals[1];
        if (n <= 32) {
          vals[n + 1] = x;
          const double thr = 0.25 * mx;
          int cplx = 0;
          double lo_sum = 0.0, hi_sum = 0.0;
          for (int i = 1; i <= n; ++i) {
            const double vi = vals[i];
            if (thr > 0.0 && vi > thr && vi > vals[i - 1] && vi > vals[i + 1]) ++cplx;
            int rank = 0;                         // position of vals[i] in the sorted band (ties in index order)
            for (int m = 1; m <= n; ++m) { const double vm = vals[m]; rank += (vm < vi || (vm == vi && m < i)) ? 1 : 0; }
            if (rank < nei) lo_sum += vi;
            if (rank >= n - nei) hi_sum += vi;
          }
          r.lo_sum = lo_sum; r.hi_sum = hi_sum; r.cplx = (double)cplx;
        } else {
          const double* br = B.bandraw + (size_t)(in_range ? rel : rel0) * BR2_STRIDE + (b - BIG0) * 3;
          r.lo_sum = br[0]; r.hi_sum = br[1]; r.cplx = br[2];
        }
        if (live) csum += band_write(B, TF, slot, b, n, nei, r);
        ++b;
        if (b < 14) { bs = P.band14_start[b]; be = bs + P.band14_n[b]; }
        s1 = s2 = s11 = s12 = s22 = 0.0; mx = 0.0; mant = 1.0; ex = 0;
      }
      if (b < 14 && k >= bs) {
        if (k == bs) vals[0] = xprev;
        s12 = fma(x, y, s12); s1 += x; s11 = fma(x, x, s11); s2 += y; s22 = fma(y, y, s22);
        mx = fmax(mx, x);
        const double v = fabs(x) + 1e-20;            // Statistics.cpp:417-455: product with the exponents peeled off
        const int hw = __double2hiint(v);            // (<= 287 factors >= 1/2: the mantissa product cannot underflow)
        ex += ((hw >> 20) & 0x7ff) - 1022;
        mant *= __hiloint2double((hw & 0x800fffff) | 0x3fe00000, __double2loint(v));
        if (k - bs < 32) vals[k - bs + 1] = x;
      }
      xprev = x;
    }
    __syncwarp();
  }
  while (b28 < 28) {                             // bands that run to the end of the row
    if (live) B.fv[(size_t)FV_BANDS28 * TF + (size_t)slot * 28 + b28] = acc28;
    acc28 = 0.0; ++b28;
  }
  if (!live) return;
  // ---- cepstrum: log of the mel energies, unnormalised DCT-II in the reference's order (vector.c:372-391) ----
  double lg[14];
#pragma unroll
  for (int q = 0; q < 14; ++q) lg[q] = log(mel[q] < 2e-42 ? 2e-42 : mel[q]);     // XTRACT_LOG_LIMIT
#pragma unroll 1
  for (int j = 0; j < 14; ++j) {
    double a = 0.0;
#pragma unroll
    for (int m = 0; m < 14; ++m) a = __dadd_rn(a, __dmul_rn(lg[m], __ldg(P.t.dct + j * 14 + m)));
    B.fv[(size_t)FV_CEPSTRUM * TF + (size_t)slot * 14 + j] = a;
  }
  B.fs[(size_t)FS_SPEC_CONTRAST * TF + slot] = csum / 14.0;
}

void afx_launch_bands(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  k_bands_select<<<dim3((B.g_slots + 7) / 8, 4), BT, 0, s>>>(B, P); ++*launches;
  k_bands_lane<<<(B.g_slots + BLW * 32 - 1) / (BLW * 32), BLW * 32, 0, s>>>(B, P); ++*launches;
}

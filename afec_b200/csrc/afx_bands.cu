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
//                   at once -- the order statistics behind the contrast (exact radix select, subband_select<C> below) and the
//                   peak count against 0.25 x band max.  3 doubles per band go to the per-frame scratch record.
//   k_bands_lane    LANE per frame, 32 consecutive frame slots per warp: everything that is a running sum over the bins
//                   of a frame -- the five correlation sums, log-sum and maximum of all 14 sub-bands, the 28 frequency
//                   bands, the 14 mel energies -- walks the row once, bin by bin, with NO cross-lane reduction at all
//                   (the first schedule spent most of its instructions in warp reductions of 7 sums x 14 bands); the
//                   nine sub-bands of <= 32 bins keep their values in a per-lane array and are ranked there; the
//                   per-band epilogue (log / exp / pow chains) runs on all 32 lanes, and a mel filter that ends goes
//                   straight into the cepstrum sums.  The bin axis is cut by the host into runs inside which nothing
//                   changes (AfxBandSeg), so the inner loop has no control flow.  The magnitude rows reach the lanes through
//                   transposed 16-bin shared-memory tiles (16 bins x 33 frames: the extra column is the frame before the
//                   warp's first, for the flux of lane 0), double buffered with cp.async: the next tile and its mel
//                   weights land while the warp walks the current one.  Sums run in bin order -- the reference's own order.
#define BR2_STRIDE 16       // doubles per frame: 5 large bands x (lo_sum, hi_sum, complexity) + pad
#define BIG0 9              // first large sub-band (41, 61, 96, 148, 287 bins)

// order statistics + complexity of one large sub-band on one warp (C = ceil(n / 32) elements per lane).
//
// Selection is an exact most-significant-digit radix select on the 64-bit patterns of the (non-negative) magnitudes,
// 8 bits per pass, histograms in shared memory: the bits common to the whole band are skipped, the first pass is shared
// by the two selections (the nei smallest / the nei largest), and a selection stops as soon as a bucket boundary splits
// the band exactly -- after one or two passes for almost every frame (the bisection this replaces took ~14 passes of
// one bit each).  `hist` = 2 x 256 ints of the warp.
template <int C>
__device__ __forceinline__ void subband_select(const AfxParams& P, int b, const double* __restrict__ g, int lane, int* hist, double* __restrict__ out3)
{
  const int s0 = P.band14_start[b], n = P.band14_n[b], nei = P.band14_nei[b];
  double x[C]; unsigned long long u[C];
  double mx = 0.0;
  unsigned long long kmin = 0xffffffffffffffffull, kmax = 0ull;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int k = lane + 32 * c;
    const bool valid = k < n;
    const double xv = valid ? g[s0 + k] : 0.0;
    x[c] = xv;
    u[c] = valid ? (unsigned long long)__double_as_longlong(xv) : 0xffffffffffffffffull;   // invalid slots sort last and are never counted
    if (valid) { kmin = min(kmin, u[c]); kmax = max(kmax, u[c]); }
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
  {
    unsigned lo32 = (unsigned)kmin, hi32 = (unsigned)(kmin >> 32);
    hi32 = __reduce_min_sync(0xffffffffu, hi32);
    lo32 = __reduce_min_sync(0xffffffffu, ((unsigned)(kmin >> 32) == hi32) ? lo32 : 0xffffffffu);
    kmin = ((unsigned long long)hi32 << 32) | lo32;
    lo32 = (unsigned)kmax; hi32 = (unsigned)(kmax >> 32);
    hi32 = __reduce_max_sync(0xffffffffu, hi32);
    lo32 = __reduce_max_sync(0xffffffffu, ((unsigned)(kmax >> 32) == hi32) ? lo32 : 0u);
    kmax = ((unsigned long long)hi32 << 32) | lo32;
  }
  // Two selections run side by side: s = 0 takes the need[0] = nei smallest, s = 1 separates the need[1] = n - nei smallest
  // from the nei largest.  State per selection: the digits fixed so far (prefix under pmask) and `below` = how many elements
  // are smaller than every remaining candidate.  A selection ends either on a key boundary with exactly need[s] elements
  // under it (exact), or -- equal values straddle the split -- on the order statistic itself after the last digit.
  unsigned long long prefix[2], pmask[2], bound[2] = { 0ull, 0ull };
  int below[2] = { 0, 0 };
  const int need[2] = { nei, n - nei };
  bool done[2] = { false, false }, exact[2] = { false, false };
  const unsigned long long diff = kmin ^ kmax;
  int lo = 0, w = 0;                                  // current digit: bits [lo, lo + w)
  if (diff == 0ull) { done[0] = done[1] = true; bound[0] = bound[1] = kmin; prefix[0] = prefix[1] = pmask[0] = pmask[1] = 0ull; }
  else {
    const int top = 63 - __clzll((long long)diff);    // highest bit in which the band's keys differ
    lo = max(0, top - 7); w = top - lo + 1;
    const unsigned long long hm = (top >= 63) ? 0ull : (~0ull << (top + 1));
    prefix[0] = prefix[1] = kmin & hm; pmask[0] = pmask[1] = hm;
  }
  while (!(done[0] && done[1])) {
    const bool shared = !done[0] && !done[1] && prefix[0] == prefix[1];     // same candidates: one histogram serves both
    for (int q = lane; q < 512; q += 32) hist[q] = 0;
    __syncwarp();
    const unsigned dm = (1u << w) - 1u;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (lane + 32 * c < n) {
        const int d = (int)((unsigned)(u[c] >> lo) & dm);
        if (!done[0] && (u[c] & pmask[0]) == prefix[0]) atomicAdd(&hist[d], 1);
        if (!shared && !done[1] && (u[c] & pmask[1]) == prefix[1]) atomicAdd(&hist[256 + d], 1);
      }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (done[s]) continue;                            // warp-uniform
      const int* h = hist + ((s == 1 && !shared) ? 256 : 0);
      int cnt[8]; int tot = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { cnt[q] = h[lane * 8 + q]; tot += cnt[q]; }
      int inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int pv = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += pv; }
      const int k = need[s] - below[s];                 // the k smallest candidates are taken, 1 <= k <= candidates
      int run = inc - tot, digit = -1, under = 0, inb = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { if (digit < 0 && cnt[q] > 0 && k > run && k <= run + cnt[q]) { digit = lane * 8 + q; under = run; inb = cnt[q]; } run += cnt[q]; }
      const int src = __ffs(__ballot_sync(0xffffffffu, digit >= 0)) - 1;
      digit = __shfl_sync(0xffffffffu, digit, src); under = __shfl_sync(0xffffffffu, under, src); inb = __shfl_sync(0xffffffffu, inb, src);
      const unsigned long long dk = prefix[s] | ((unsigned long long)digit << lo);
      if (under + inb == k) {                           // the bucket ends exactly at the split: everything up to it is taken
        done[s] = true; exact[s] = true; bound[s] = dk + (1ull << lo);      // keys < bound: exactly need[s] elements (finite keys: no wrap)
      } else if (lo == 0) {                             // last digit: dk is the order statistic, equal values straddle the split
        done[s] = true; bound[s] = dk;
      } else {
        below[s] += under; prefix[s] = dk; pmask[s] |= (unsigned long long)dm << lo;
      }
    }
    const int nlo = max(0, lo - 8);
    w = lo - nlo; lo = nlo;
    __syncwarp();                                       // histogram reads above are over before the next pass clears it
  }
  // sums over the band with the two bounds
  int nlt0 = 0, ngt1 = 0; double slt0 = 0.0, sge1 = 0.0, sgt1 = 0.0;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    if (lane + 32 * c < n) {
      if (u[c] < bound[0]) { ++nlt0; slt0 += x[c]; }
      if (u[c] >= bound[1]) sge1 += x[c];
      if (u[c] > bound[1]) { ++ngt1; sgt1 += x[c]; }
    }
  }
  nlt0 = __reduce_add_sync(0xffffffffu, nlt0); ngt1 = __reduce_add_sync(0xffffffffu, ngt1);
  slt0 = warp_sum(slt0); sge1 = warp_sum(sge1); sgt1 = warp_sum(sgt1);
  if (lane == 0) {
    out3[0] = exact[0] ? slt0 : slt0 + (double)(nei - nlt0) * __longlong_as_double((long long)bound[0]);
    out3[1] = exact[1] ? sge1 : sgt1 + (double)(nei - ngt1) * __longlong_as_double((long long)bound[1]);
    out3[2] = (double)cplx;
  }
}

__global__ void __launch_bounds__(BT, 6) k_bands_select(AfxBatchDev B, AfxParams P)
{
  __shared__ int hists[BT / 32][512];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int* hist = hists[wid];
  const int rel = blockIdx.x * 8 + wid;
  if (rel >= B.g_slots) return;
  const int slot = B.slot0 + rel;
  const int fi = B.slot_file[slot];
  const int t = slot - B.files[fi].frame_off;
  if (B.files[fi].status != 0 || t >= B.state[fi].F) return;
  const double* __restrict__ g = B.mag + (size_t)rel * AFX_NBIN;
  double* out = B.bandraw + (size_t)rel * BR2_STRIDE;
  switch (blockIdx.y) {                         // heaviest band first in launch order (grid y is the slow index)
    case 0: subband_select<9>(P, 13, g, lane, hist, out + 12); break;
    case 1: subband_select<5>(P, 12, g, lane, hist, out + 9); break;
    case 2: subband_select<3>(P, 11, g, lane, hist, out + 6); break;
    default: subband_select<2>(P, 10, g, lane, hist, out + 3); subband_select<2>(P, 9, g, lane, hist, out); break;
  }
}

#define BLW 4               // warps per CTA of k_bands_lane (128 frame slots)

// finish sub-band b of one frame (lane-local): order statistics / peak count of a small band from the lane's value array
// (vals[0] and vals[n + 1] are the band's neighbours), of a large band from k_bands_select's record; epilogue; returns the contrast
__device__ __forceinline__ double lane_finish_band(AfxBatchDev& B, const AfxParams& P, size_t TF, int slot, int rel, int b, bool live,
                                                   const double* vals, BandRaw& r, double mx)
{
  const int n = P.band14_n[b], nei = P.band14_nei[b];
  if (n <= 32) {
    const double thr = 0.25 * mx;
    int cplx = 0;
    double lo_sum = 0.0, hi_sum = 0.0;
    for (int i = 1; i <= n; ++i) {
      const double vi = vals[i];
      if (thr > 0.0 && vi > thr && vi > vals[i - 1] && vi > vals[i + 1]) ++cplx;
      int rank = 0;                             // position of vals[i] in the sorted band (ties in index order)
      for (int m = 1; m <= n; ++m) { const double vm = vals[m]; rank += (vm < vi || (vm == vi && m < i)) ? 1 : 0; }
      if (rank < nei) lo_sum += vi;
      if (rank >= n - nei) hi_sum += vi;
    }
    r.lo_sum = lo_sum; r.hi_sum = hi_sum; r.cplx = (double)cplx;
  } else {
    const double* br = B.bandraw + (size_t)rel * BR2_STRIDE + (b - BIG0) * 3;
    r.lo_sum = br[0]; r.hi_sum = br[1]; r.cplx = br[2];
  }
  r.x0 = (n >= 2) ? r.x0 : vals[1];
  return live ? band_write(B, TF, slot, b, n, nei, r) : 0.0;
}

__global__ void __launch_bounds__(BLW * 32, 4) k_bands_lane(AfxBatchDev B, AfxParams P)
{
  // two tiles of 16 bins per warp: the next 16 bins arrive (cp.async, 8 bytes per lane and row, transposed on the way in) while the
  // warp walks the current ones -- with one tile the walk stood still for a global round trip once per tile
  // (ncu: 42 % of the stall samples on the loads' scoreboard)
  __shared__ double tiles[BLW][2][16][33];
  __shared__ double2 melw[BLW][2][16];          // the tile's mel weights (per bin: the two filters that can cover it)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rel0 = (blockIdx.x * BLW + wid) * 32;
  if (rel0 >= B.g_slots) return;                // warp-uniform
  double (*tile)[33] = tiles[wid][0];
  const double2* mw = melw[wid][0];
  const int rel_raw = rel0 + lane;
  const bool in_range = rel_raw < B.g_slots;
  const int rel = in_range ? rel_raw : rel0;
  const int slot = B.slot0 + rel;
  const int fi = B.slot_file[slot];
  const int t = slot - B.files[fi].frame_off;
  const bool live = in_range && B.files[fi].status == 0 && t < B.state[fi].F;
  const bool has_prev = t > 0;                  // SampleAnalyser.cpp:936-940: a file's first frame correlates with itself
  const size_t TF = (size_t)B.TF;
  const double* __restrict__ mag = B.mag;
  const int last_row = B.g_slots - 1;
  // rows rel0 - 1 .. rel0 + 31, bins k0 .. k0 + 15 -> tiles[wid][buf]; half a warp per row on the way in (lane = bin),
  // lane = frame on the way out
  auto fetch_tile = [&](int k0, int buf) {
    const int hb = lane >> 4, b = lane & 15;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(&tiles[wid][buf][b][hb]);
#pragma unroll
    for (int c = 0; c < 17; ++c) {
      const int r = 2 * c + hb;
      if (r < 33) {
        const int rc = min(max(rel0 - 1 + r, 0), last_row);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst + 16 * c), "l"(mag + (size_t)rc * AFX_NBIN + k0 + b) : "memory");
      }
    }
    if (lane < 16) {
      const unsigned wd = (unsigned)__cvta_generic_to_shared(&melw[wid][buf][lane]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(wd), "l"(P.t.mel_ab + k0 + lane) : "memory");
    }
  };
  fetch_tile(0, 0);

  // state of the walk.  Everything that changes along the bin axis is constant inside a segment (AfxBandSeg, built by
  // afx_create), so the per-bin loop below is straight arithmetic.
  double s1 = 0, s2 = 0, s11 = 0, s12 = 0, s22 = 0, mx = 0, mant = 1.0; int ex = 0, nv = 0;
  double vals[34];                              // a small band with its two neighbours: vals[0] = x[start - 1], vals[1 + i] = x[start + i]
  double xprev = 0.0, a28 = 0.0, csum = 0.0;
  // mel energies: at most two consecutive filters are open at a time -- they live in two registers by parity; a filter that
  // ends goes straight into the cepstrum sums (log, then its column of the DCT), in filter order as the reference sums them
  double melA = 0.0, melB = 0.0, cep[14];
#pragma unroll
  for (int j = 0; j < 14; ++j) cep[j] = 0.0;
  int next_q = 0;                               // next filter the cepstrum is waiting for (uniform)
  auto cep_add = [&](int q, double energy) {    // vector.c:372-391; filters without support come in with energy 0
    for (; next_q <= q; ++next_q) {
      const double en = (next_q == q) ? energy : 0.0;
      const double lg = log(en < 2e-42 ? 2e-42 : en);               // XTRACT_LOG_LIMIT
#pragma unroll
      for (int j = 0; j < 14; ++j) cep[j] = __dadd_rn(cep[j], __dmul_rn(lg, __ldg(P.t.dct + j * 14 + next_q)));
    }
  };
  int pending = -1;                             // sub-band waiting for its right neighbour
  BandRaw pend; double pend_mx = 0.0;
  pend.s1 = pend.s2 = pend.s11 = pend.s12 = pend.s22 = pend.ls = pend.x0 = pend.lo_sum = pend.hi_sum = pend.cplx = 0.0;

  const AfxBandSeg* __restrict__ segs = P.t.band_segs;
  const int nseg = P.t.n_band_segs;
#pragma unroll 1
  for (int si = 0; si < nseg; ++si) {
    const AfxBandSeg sg = segs[si];             // uniform
    const int k0 = sg.k0, k1 = sg.k1;
    if ((k0 & 15) == 0) {                       // the segments are cut at every 16 bins: a new tile starts here
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();                             // every lane's part has landed, and every lane is done with the tile before
      tile = tiles[wid][(k0 >> 4) & 1];
      mw = melw[wid][(k0 >> 4) & 1];
      if (k0 + 16 < AFX_NBIN) fetch_tile(k0 + 16, ((k0 >> 4) + 1) & 1);
    }
    const bool in14 = sg.b14 >= 0, in28 = sg.b28 >= 0;
    const bool small14 = in14 && P.band14_n[sg.b14] <= 32;
    double e0 = 0.0, e1 = 0.0;
    if (pending >= 0) {                         // the first bin of this segment is the right neighbour of the band that just ended
      vals[nv + 1] = tile[k0 & 15][lane + 1];
      csum += lane_finish_band(B, P, TF, slot, rel, pending, live, vals, pend, pend_mx);
      pending = -1;
    }
    if (sg.start14) { s1 = s2 = s11 = s12 = s22 = 0.0; mx = 0.0; mant = 1.0; ex = 0; nv = 0; vals[0] = xprev; }
    double x = xprev;
#pragma unroll 2
    for (int k = k0; k < k1; ++k) {
      x = tile[k & 15][lane + 1];
      const double y = has_prev ? tile[k & 15][lane] : x;
      if (in14) {                               // SampleAnalyser.cpp:2067-2260
        s12 = fma(x, y, s12); s1 += x; s11 = fma(x, x, s11); s2 += y; s22 = fma(y, y, s22);
        mx = fmax(mx, x);
        const double v = fabs(x) + 1e-20;            // Statistics.cpp:417-455: product with the exponents peeled off
        const int hw = __double2hiint(v);            // (<= 287 factors >= 1/2: the mantissa product cannot underflow)
        ex += ((hw >> 20) & 0x7ff) - 1022;
        mant *= __hiloint2double((hw & 0x800fffff) | 0x3fe00000, __double2loint(v));
        if (small14) vals[++nv] = x;
      }
      if (in28) a28 = fma(x, x, a28);           // SampleAnalyser.cpp:2007-2048
      const double2 wab = mw[k & 15];            // vector.c:350-391, on the filters' supports (weights outside are 0 and unused)
      if (sg.nq > 0) e0 = fma(x, wab.x, e0);
      if (sg.nq > 1) e1 = fma(x, wab.y, e1);
    }
    xprev = x;
    if (sg.nq > 0) { if (sg.q0 & 1) melB += e0; else melA += e0; }
    if (sg.nq > 1) { if (sg.q0 & 1) melA += e1; else melB += e1; }
    if (sg.fin & 1) { if (sg.q0 & 1) { cep_add(sg.q0, melB); melB = 0.0; } else { cep_add(sg.q0, melA); melA = 0.0; } }
    if (sg.fin & 2) { if (sg.q0 & 1) { cep_add(sg.q0 + 1, melA); melA = 0.0; } else { cep_add(sg.q0 + 1, melB); melB = 0.0; } }
    if (sg.end28) { if (live) B.fv[(size_t)FV_BANDS28 * TF + (size_t)slot * 28 + sg.b28] = a28; a28 = 0.0; }
    if (sg.end14) {
      pend.s1 = s1; pend.s2 = s2; pend.s11 = s11; pend.s12 = s12; pend.s22 = s22; pend.ls = mant; pend.x0 = (double)ex; pend_mx = mx;
      pending = sg.b14;
    }
  }
  if (pending >= 0) { vals[nv + 1] = 0.0; csum += lane_finish_band(B, P, TF, slot, rel, pending, live, vals, pend, pend_mx); }
  if (!live) return;
  for (int b = 0; b < 28; ++b)                   // bands that lie beyond the spectrum hold no bins (SampleAnalyser.cpp:2026-2045)
    if (P.band28_e[b] <= P.band28_s[b]) B.fv[(size_t)FV_BANDS28 * TF + (size_t)slot * 28 + b] = 0.0;
  // ---- cepstrum: log of the mel energies, unnormalised DCT-II in the reference's order (vector.c:372-391); filters that
  // never ended (no support) enter with the log limit ----
  cep_add(13, 0.0);
#pragma unroll
  for (int j = 0; j < 14; ++j) B.fv[(size_t)FV_CEPSTRUM * TF + (size_t)slot * 14 + j] = cep[j];
  B.fs[(size_t)FS_SPEC_CONTRAST * TF + slot] = csum / 14.0;
}

void afx_launch_bands(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  k_bands_select<<<dim3((B.g_slots + 7) / 8, 4), BT, 0, s>>>(B, P); ++*launches;
  k_bands_lane<<<(B.g_slots + BLW * 32 - 1) / (BLW * 32), BLW * 32, 0, s>>>(B, P); ++*launches;
}

// K5: per-file temporal aggregation -- TStatistics::Calc (Statistics.cpp:12-90) over each of the 136
// series of a file (24 framed scalars, 5 x 14 sub-band series, 28 frequency bands, 14 cepstrum bands;
// SampleAnalyser.cpp:2402-2412, SampleDescriptors.h:212-230, 327-355).
//
// A segmented reduction: one warp per (file, series); the series are segments of the batch-wide
// descriptor arrays (frame_off / rframe_off give the segment start, F / Fr its length).  A series of up to SCACHE values is
// read ONCE into the warp's shared-memory row (the vector series are strided in global memory: a sector per value) and
// every pass below runs on that copy; longer ones (the two onset series) stay in global memory.  Thirteen
// statistics per segment: min, max, lower median (Statistics.cpp:316-413, here an MSD radix select on order-preserving
// 64-bit keys that skips the bytes all keys share and stops once a bucket holds one key -- typically 2-3 passes instead
// of 8), mean, geometric mean, variance (/n), the reference's
// index-weighted centroid / spread, its "value minus centroid over spread" skewness / kurtosis,
// flatness = gmean / mean, and mean / variance of |x[i+1] - x[i]|.
#include "afx_common.cuh"
#include "../../include/afec_b200.h"

#define SW 8            // warps per CTA
#define SCACHE 1024     // values of a series kept in shared memory (frame series: <= 862 at hop 1024, the reference's hop)

__device__ __forceinline__ unsigned long long order_key(double x)
{
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k)
{
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// series the configured feature groups produce (the others hold no data -- their arrays may not even be allocated -- and
// get all-zero statistics)
__device__ __forceinline__ bool series_enabled(int s, unsigned feat)
{
  if (s < 4) return feat & AFX_FEAT_AMPLITUDE;
  if (s <= FS_SPEC_FLATNESS || s == FS_SPEC_INHARM || s == FS_SPEC_FLUX || (s >= FS_TRISTIM1 && s <= FS_TRISTIM3)) return feat & AFX_FEAT_SPECTRAL;
  if (s == FS_SPEC_COMPLEXITY) return feat & AFX_FEAT_PEAKS;
  if (s == FS_SPEC_CONTRAST) return feat & AFX_FEAT_BANDS;
  if (s >= FS_F0 && s <= FS_F0_FAILSAFE) return feat & AFX_FEAT_PITCH;
  if (s == FS_AUTOCORR) return feat & AFX_FEAT_AUTOCORR;
  if (s < AFX_N_FS) return feat & AFX_FEAT_RHYTHM;
  return feat & AFX_FEAT_BANDS;
}

// the thirteen statistics of one series of n > 1 values; ld(i) reads value i (shared-memory copy or global memory)
template <class Ld>
__device__ __forceinline__ void series_stats(Ld ld, int n, int lane, int* h, double (&r)[AFX_N_STATS])
{
    // ---- pass 1 -------------------------------------------------------------------------------------
    double sum = 0, sj = 0, sd = 0, mn = 1.0e308, mx = -1.0e308, mant = 1.0; int ex = 0;
    unsigned long long kmin = ~0ull, kmax = 0ull;
    for (int i = lane; i < n; i += 32) {
      const double v = ld(i);
      sum += v; sj += (double)i * v; mn = fmin(mn, v); mx = fmax(mx, v);
      const unsigned long long key = order_key(v);
      kmin = key < kmin ? key : kmin; kmax = key > kmax ? key : kmax;
      mul_frexp(mant, ex, fabs(v) + 1e-20);
      if (i + 1 < n) sd += fabs(ld(i + 1) - v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
      kmin = a < kmin ? a : kmin; kmax = b > kmax ? b : kmax;
    }
    double ls = log(mant) + (double)ex * 0.693147180559945309417;
    sum = warp_sum(sum); sj = warp_sum(sj); sd = warp_sum(sd); ls = warp_sum(ls);
    mn = -warp_max(-mn); mx = warp_max(mx);
    const double dn = (double)n;
    const double mean = sum / dn;
    const double gmean = exp(ls / dn);
    const double cen = (sum == 0.0) ? 0.0 : sj / sum;
    const double dmean = (n > 2) ? sd / (double)(n - 1) : 0.0;
    // ---- pass 2 -------------------------------------------------------------------------------------
    double var = 0, sp = 0, dvar = 0;
    for (int i = lane; i < n; i += 32) {
      const double v = ld(i);
      var += (v - mean) * (v - mean);
      const double d = (double)i - cen; sp += d * d * v;
      if (n > 2 && i + 1 < n) { const double q = fabs(ld(i + 1) - v) - dmean; dvar += q * q; }
    }
    var = warp_sum(var); sp = warp_sum(sp); dvar = warp_sum(dvar);
    const double spread = (sum == 0.0) ? 0.0 : sp / sum;
    // ---- pass 3 -------------------------------------------------------------------------------------
    double sk = 0, ku = 0;
    const bool have = fabs(spread) > (double)1e-12f;
    if (have) for (int i = lane; i < n; i += 32) {
      const double d = (ld(i) - cen) / spread; const double d2 = d * d;
      sk += d2 * d; ku += d2 * d2;
    }
    sk = warp_sum(sk); ku = warp_sum(ku);
    // ---- lower median: MSD radix select of rank (n-1)/2 ------------------------------------------------
    // every key lies in [kmin, kmax]: the leading bytes those two share are settled
    unsigned long long prefix = 0ull, pmask = 0ull;
    int k = (n - 1) / 2;
        int byte = 7;
    {
      const unsigned long long diff = kmin ^ kmax;
      const int lead = diff ? (__clzll((long long)diff) >> 3) : 8;
      byte = 7 - lead;
      if (lead) { pmask = (lead == 8) ? ~0ull : (~0ull << (8 * (8 - lead))); prefix = kmin & pmask; }
    }
    for (; byte >= 0; --byte) {
      for (int q = lane; q < 256; q += 32) h[q] = 0;
      __syncwarp();
      const int sh = byte * 8;
      for (int i = lane; i < n; i += 32) {
        const unsigned long long key = order_key(ld(i));
        if ((key & pmask) == prefix) atomicAdd(&h[(int)((key >> sh) & 0xff)], 1);
      }
      __syncwarp();
      // each lane owns 8 consecutive buckets
      int c[8]; int tot = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { c[q] = h[lane * 8 + q]; tot += c[q]; }
      int inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int pv = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += pv; }
      const int excl = inc - tot;
      int digit = -1, newk = 0, inb = 0;
      if (k >= excl && k < inc) {
        int run = excl;
#pragma unroll
        for (int q = 0; q < 8; ++q) { if (digit < 0 && k < run + c[q]) { digit = lane * 8 + q; newk = k - run; inb = c[q]; } run += c[q]; }
      }
      const unsigned ball = __ballot_sync(0xffffffffu, digit >= 0);
      const int src = __ffs(ball) - 1;
      digit = __shfl_sync(0xffffffffu, digit, src); newk = __shfl_sync(0xffffffffu, newk, src); inb = __shfl_sync(0xffffffffu, inb, src);
      prefix |= ((unsigned long long)digit) << sh; pmask |= 0xffull << sh; k = newk;
      __syncwarp();
      if (inb == 1 && byte > 0) {                     // one key left under this prefix: it is the median
        unsigned long long found = 0ull;
        for (int i = lane; i < n; i += 32) {
          const unsigned long long key = order_key(ld(i));
          if ((key & pmask) == prefix) found = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) found |= __shfl_xor_sync(0xffffffffu, found, o);
        prefix = found;
        break;
      }
    }
    r[0] = mn; r[1] = mx; r[2] = key_to_double(prefix); r[3] = mean; r[4] = gmean; r[5] = var / dn;
    r[6] = cen; r[7] = spread; r[8] = have ? sk / dn : 0.0; r[9] = have ? ku / dn - 3.0 : 0.0;
    r[10] = (mean == 0.0) ? 0.0 : gmean / mean;
    r[11] = dmean; r[12] = (n > 2) ? dvar / (double)(n - 1) : 0.0;
}

__global__ void __launch_bounds__(SW * 32) k_stats(AfxBatchDev B, AfxParams P, unsigned features)
{
  extern __shared__ __align__(16) unsigned char st_smem[];
  double (*cache)[SCACHE] = reinterpret_cast<double (*)[SCACHE]>(st_smem);
  int (*hist)[256] = reinterpret_cast<int (*)[256]>(st_smem + sizeof(double) * SW * SCACHE);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * SW + wid;
  // the two onset series of a file are eight times longer than its frame series: they come first in the grid, eight to a CTA,
  // so that no CTA holds its shared memory for one long warp while seven short ones are done
  constexpr int NR = AFX_N_FS - AFX_N_FS_MAIN;
  const long long n_long = (long long)B.n_files * NR;
  int fi, s;
  if (gw < n_long) { fi = (int)(gw / NR); s = AFX_N_FS_MAIN + (int)(gw % NR); }
  else {
    const long long g2 = gw - n_long;
    fi = (int)(g2 / (AFX_N_SERIES - NR)); s = (int)(g2 % (AFX_N_SERIES - NR));
    if (s >= AFX_N_FS_MAIN) s += NR;
  }
  if (fi >= B.n_files) return;
  const AfxFile f = B.files[fi];
  double* out = B.stats + ((size_t)fi * AFX_N_SERIES + s) * AFX_N_STATS;
  if (f.status != 0 || !series_enabled(s, features)) { if (lane < AFX_N_STATS) out[lane] = 0.0; return; }
  const AfxState* st = B.state + fi;
  const size_t TF = (size_t)B.TF;
  const double* x; int n, stride;
  if (s < AFX_N_FS_MAIN) { x = B.fs + (size_t)s * TF + f.frame_off; n = st->F; stride = 1; }
  else if (s < AFX_N_FS) { x = B.fsr + (size_t)(s - AFX_N_FS_MAIN) * B.TFr + f.rframe_off; n = st->Fr; stride = 1; }
  else {
    int b = s - AFX_N_FS, off, nb;
    if (b < 70) { off = (b / 14) * 14; nb = 14; b = b % 14; }
    else if (b < 98) { off = FV_BANDS28; nb = 28; b -= 70; }
    else { off = FV_CEPSTRUM; nb = 14; b -= 98; }
    x = B.fv + (size_t)off * TF + (size_t)f.frame_off * nb + b; n = st->F; stride = nb;
  }
  double r[AFX_N_STATS];
#pragma unroll
  for (int k = 0; k < AFX_N_STATS; ++k) r[k] = 0.0;
  if (n == 1) { const double v = x[0]; r[0] = v; r[1] = v; r[3] = v; }
  if (n > 1) {
    if (n <= SCACHE) {
      double* c = cache[wid];
      for (int i = lane; i < n; i += 32) c[i] = x[(size_t)i * stride];
      __syncwarp();
      series_stats([c](int i) { return c[i]; }, n, lane, hist[wid], r);
    } else {
      series_stats([x, stride](int i) { return x[(size_t)i * stride]; }, n, lane, hist[wid], r);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < AFX_N_STATS; ++q) out[q] = r[q];
  }
}

void afx_launch_stats(const AfxParams& P, const AfxBatchDev& B, unsigned features, cudaStream_t s, long long* launches)
{
  if (B.n_files <= 0) return;
  const long long warps = (long long)B.n_files * AFX_N_SERIES;
  const int smem = (int)(sizeof(double) * SW * SCACHE + sizeof(int) * SW * 256);
  cudaFuncSetAttribute(k_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device (one process may drive several)
  k_stats<<<(unsigned)((warps + SW - 1) / SW), SW * 32, smem, s>>>(B, P, features); ++*launches;
}

// K6a: auto_correlation -- TSampleAnalyser::CalcAutoCorrelation (SampleAnalyser.cpp:2312-2398) with
// TAutocorrelation::Calc (Source/Crawler/FeatureExtraction/Source/Autocorrelation.cpp:62-104).
//
// Per main frame: from the frame start (looking ahead over the REST OF THE FILE, not just the frame)
// find the first rising sample pair within 1024 samples, then the next rising pair at least 0.8 ms
// (35 samples) later -> period; correlate the 12 ms (529 samples) window that starts at the first pair
// with itself for lags 0..528 (R[i] = sum_{j < 529-i} x[j] x[j+i]), normalise by R[0] and report the
// largest coefficient at lags >= period / 2.
//
// One WARP per frame, 8 frames per CTA, no block-wide barrier.  The two searches read the signal
// straight from global memory in 32-sample steps (they nearly always stop in the first step); only the
// 529-sample window goes to shared memory.  The 140k multiply-adds per frame are register tiled: a lane
// owns 9 consecutive lags and slides a 9-sample window along j, so every pair of shared-memory loads
// feeds 9 DFMAs (the FP64 pipe, not shared memory, is the limit); lag groups g and G-1-g are paired so
// that all lanes carry the same number of products.
#include "afx_common.cuh"

#define AW 8                // warps (frames) per CTA
#define AL 9                // lags per lane task
#define AC_MAXW 544         // >= ac_width (529) + AL, multiple of 8

// returns max(R[i]) over the group's lags i with lo <= i < width; the group holding lag 0 also reports R[0]
__device__ __forceinline__ double ac_group(const double* __restrict__ x, int width, int g, int lo, double& r0)
{
  const int i0 = g * AL;
  const int nj = width - i0;            // products of the group's first lag; later lags read the zero padding
  double acc[AL], w[AL];
#pragma unroll
  for (int q = 0; q < AL; ++q) { acc[q] = 0.0; w[q] = x[i0 + q]; }
  for (int j = 0; j < nj; ++j) {
    const double a = x[j];
#pragma unroll
    for (int q = 0; q < AL; ++q) acc[q] = fma(a, w[q], acc[q]);
#pragma unroll
    for (int q = 0; q < AL - 1; ++q) w[q] = w[q + 1];
    w[AL - 1] = x[j + i0 + AL];
  }
  if (g == 0) r0 = acc[0];
  double best = 0.0;
#pragma unroll
  for (int q = 0; q < AL; ++q) if (i0 + q < width && i0 + q >= lo) best = fmax(best, acc[q]);
  return best;
}

__global__ void __launch_bounds__(AW * 32) k_autocorr(AfxBatchDev B, AfxParams P)
{
  __shared__ double xs[AW][AC_MAXW + 2 * AL + 8];

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rel = blockIdx.x * AW + wid;
  if (rel >= B.g_slots) return;                    // warp-uniform; only warp-level sync below
  const int slot = B.slot0 + rel;
  const int fi = B.slot_file[slot];
  const AfxFile f = B.files[fi];
  const AfxState st = B.state[fi];
  const int t = slot - f.frame_off;
  if (f.status != 0 || t >= st.F) return;
  const int n0 = t * P.H;
  const float* __restrict__ mono = B.mono + f.mono_off;
  int remaining = st.len - n0;                                   // SampleAnalyser.cpp:943
  const int max_seek = P.N / 2;

  // first rising pair (SampleAnalyser.cpp:2331-2341)
  int start = 0;
  {
    const int lim = min(remaining, max_seek) - 1;
    for (int base = 0; base < lim; base += 32) {
      const int i = base + lane;
      const bool hit = (i < lim) && (mdata(mono, st, n0 + i + 1) > mdata(mono, st, n0 + i));
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) { start = base + __ffs(m) - 1; remaining -= start; break; }
    }
  }
  // next rising pair at least min_period later (SampleAnalyser.cpp:2344-2356)
  const int seek_off = min(remaining, P.ac_min_period);
  int period = seek_off;
  {
    const int lim = min(remaining - seek_off, max_seek) - 1;
    const int o = n0 + start + seek_off;
    for (int base = 0; base < lim; base += 32) {
      const int i = base + lane;
      const bool hit = (i < lim) && (mdata(mono, st, o + i + 1) > mdata(mono, st, o + i));
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) { period = seek_off + base + __ffs(m) - 1; break; }
    }
  }
  double* out = B.fs + (size_t)FS_AUTOCORR * B.TF + slot;
  if (!remaining || period >= remaining) { if (lane == 0) *out = 0.0; return; }   // :2361-2365

  const int width = min(remaining, P.ac_width);
  double* x = xs[wid];
  for (int k = lane; k < AC_MAXW + 2 * AL + 8; k += 32) x[k] = (k < width) ? mdata(mono, st, n0 + start + k) : 0.0;
  __syncwarp();

  const int G = (width + AL - 1) / AL;             // lag groups; the last one may be partial (zero padded)
  const int lo = period / 2;
  double r0 = 0.0, best = 0.0;                     // the result is floored at 0 (Autocorrelation.cpp:96-103)
  for (int g = lane; g < (G + 1) / 2; g += 32) {
    best = fmax(best, ac_group(x, width, g, lo, r0));
    const int g2 = G - 1 - g;
    if (g2 != g) best = fmax(best, ac_group(x, width, g2, lo, r0));
  }
  best = warp_max(best);
  r0 = __shfl_sync(0xffffffffu, r0, 0);            // lane 0 owns group 0
  // normalisation by R[0] > 0 is monotonic, so max_i (R[i] / R[0]) == (max_i R[i]) / R[0] exactly
  if (lane == 0) *out = (r0 != 0) ? best / r0 : best;
}

void afx_launch_autocorr(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  k_autocorr<<<(B.g_slots + AW - 1) / AW, AW * 32, 0, s>>>(B, P); ++*launches;
}

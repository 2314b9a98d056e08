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
// feeds 9 DFMAs (the FP64 pipe, not shared memory, is the limit); see ac_rounds for how the triangle of
// lag x sample work is spread over the lanes.
#include "afx_common.cuh"
#include <cstdlib>

#define AW 4                // warps (frames) per CTA
#define AL 9                // lags per lane task
#define AC_MAXW 544         // >= ac_width (529) + AL, multiple of 8
#define AC_XS 696           // window + zero padding: a round reads up to width + 75; idle lanes read zeros from AC_ZERO on
#define AC_ZERO 544

// The lag groups g (lags 9g .. 9g+8, width - 9g products each) form a triangle of work.  Four lanes share a group
// (each takes a quarter of its j range and slides its own 9-sample window) and the warp walks the groups 8 at a
// time: inside a round every lane runs the same number of steps -- a quarter of the round's longest group, the
// shorter ones run into the zero padding -- so the warp never waits for one long lane.  88 % of the lane-steps carry
// products (62 % when a lane owned whole groups).  Returns max(R[i]) over lags lo <= i < width; r0 = R[0].
__device__ __forceinline__ double ac_rounds(const double* __restrict__ x, int width, int G, int lane, int lo, double& r0)
{
  const int sub = lane & 3, gl = lane >> 2;
  double best = 0.0;
  for (int gbase = 0; gbase < G; gbase += 8) {
    const int g = gbase + gl, i0 = AL * g;
    const bool act = g < G;
    const int len = (width - AL * gbase + 3) >> 2;   // uniform across the warp
    const int j0 = sub * len;
    const double* __restrict__ xa = x + j0;
    const double* __restrict__ xw = x + (act ? j0 + i0 : AC_ZERO);
    double acc[AL], w[AL];
#pragma unroll
    for (int q = 0; q < AL; ++q) { acc[q] = 0.0; w[q] = xw[q]; }
    for (int jj = 0; jj < len; ++jj) {
      const double a = xa[jj];
#pragma unroll
      for (int q = 0; q < AL; ++q) acc[q] = fma(a, w[q], acc[q]);
#pragma unroll
      for (int q = 0; q < AL - 1; ++q) w[q] = w[q + 1];
      w[AL - 1] = xw[jj + AL];
    }
#pragma unroll
    for (int q = 0; q < AL; ++q) {
      acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 1);
      acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 2);
      if (act && i0 + q < width && i0 + q >= lo) best = fmax(best, acc[q]);
    }
    if (gbase == 0) r0 = acc[0];                     // lanes 0..3 hold R[0]
  }
  return best;
}

// FP32 form of the same tiling (round 2, "precision demotion"): R[i] / R[0] does not depend on the file's scale, so the
// window is taken as the RAW float32 mono samples (exact), a lane owns 17 lags (two shared-memory loads feed 17 FFMAs on
// the pipe that is twice as wide as the FP64 one) and a round's partial sums -- at most 133 products each -- stay in FP32;
// they are widened to FP64 before the four lanes of a group and the rounds are combined.  Rounding analysis: the error of
// a partial sum is ~ eps sqrt(n) |s| / 3 ~ 2e-7 |s| with |s| <= R[0] / 4, i.e. <= 1e-7 of R[0], against a tolerance of
// 1e-6 + 1e-4 |v| on v = R[i] / R[0] in [0, 1]; nothing downstream of this value is a threshold (it is a maximum over lags,
// then temporal statistics).  Measured against the FP64 kernel: profiles/README.md.  AFX_AUTOCORR_FP64=1 keeps the FP64 form.
#define ALF 17              // lags per lane task (odd: the lag groups and the quarters of a group then start on different banks)
#define ACF_XS 768          // floats: window + zero padding
#define ACF_ZERO 592        // idle lanes read zeros from here (a round reads at most 136 + 34 samples)
__device__ __forceinline__ double ac_rounds_f32(const float* __restrict__ x, int width, int G, int lane, int lo, double& r0)
{
  const int sub = lane & 3, gl = lane >> 2;
  double best = 0.0;
  for (int gbase = 0; gbase < G; gbase += 8) {
    const int g = gbase + gl, i0 = ALF * g;
    const bool act = g < G;
    // a quarter of the round's longest group, rounded up to the unroll depth: steps past a group's last product read the
    // zero padding behind the window (x[j + i] = 0 for j + i >= width), so a longer walk adds nothing
    const int len = ((((width - ALF * gbase + 3) >> 2) + ALF - 1) / ALF) * ALF;    // uniform across the warp
    const int j0 = sub * len;
    const float* __restrict__ xa = x + j0;
    const float* __restrict__ xw = x + (act ? j0 + i0 : ACF_ZERO);
    float acc[ALF], w[ALF];
#pragma unroll
    for (int q = 0; q < ALF; ++q) { acc[q] = 0.0f; w[q] = xw[q]; }
    // ALF steps per trip with the window registers addressed modulo ALF: register u holds x[.. + u] until step u has
    // used it, then takes the sample ALF further on -- the window slides without a single register move.  (With 16 lags
    // per lane every group started on bank 0 or 16: 24 ns per frame instead of 12.)
    for (int jj = 0; jj < len; jj += ALF) {
#pragma unroll
      for (int u = 0; u < ALF; ++u) {
        const float a = xa[jj + u];
#pragma unroll
        for (int q = 0; q < ALF; ++q) acc[q] = fmaf(a, w[(q + u) % ALF], acc[q]);
        w[u] = xw[jj + u + ALF];
      }
    }
#pragma unroll
    for (int q = 0; q < ALF; ++q) {
      double d = (double)acc[q];
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      if (act && i0 + q < width && i0 + q >= lo) best = fmax(best, d);
      if (gbase == 0 && q == 0) r0 = d;              // lanes 0..3 hold R[0]
    }
  }
  return best;
}

template <bool F32>
__global__ void __launch_bounds__(AW * 32) k_autocorr(AfxBatchDev B, AfxParams P)
{
  __shared__ double xs[F32 ? 1 : AW][AC_XS];
  __shared__ float xf[F32 ? AW : 1][ACF_XS];

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

  // The two searches compare neighbouring samples of the conditioned signal, mdata(i) = raw(i) * fs with fs > 0: the product
  // of a float32 and a double cannot merge two different floats, so the RAW samples compare the same way -- no conversion.
  // A warp looks at 128 samples per step (four independent loads per lane in flight): frames in digital silence have no
  // rising pair and walk the whole 1024-sample limit, one dependent global round trip per step -- with 32 samples per step
  // the searches were 40 % of this kernel's time (ncu, round 2).
  auto raw = [&](int i) -> float {
    const int j = i - st.start_off;
    return (j >= 0 && j < st.audible) ? __ldg(mono + st.lead + j) : 0.0f;
  };
  auto first_rise = [&](int o, int lim) -> int {             // smallest i in [0, lim) with raw(o + i + 1) > raw(o + i), else -1
    for (int base = 0; base < lim; base += 128) {
      float a[4], b[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { const int i = base + 32 * k + lane; a[k] = (i < lim) ? raw(o + i) : 0.0f; b[k] = (i < lim) ? raw(o + i + 1) : 0.0f; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned m = __ballot_sync(0xffffffffu, (base + 32 * k + lane < lim) && b[k] > a[k]);
        if (m) return base + 32 * k + __ffs(m) - 1;
      }
    }
    return -1;
  };
  // first rising pair (SampleAnalyser.cpp:2331-2341)
  int start = 0;
  {
    const int hit = first_rise(n0, min(remaining, max_seek) - 1);
    if (hit >= 0) { start = hit; remaining -= start; }
  }
  // next rising pair at least min_period later (SampleAnalyser.cpp:2344-2356)
  const int seek_off = min(remaining, P.ac_min_period);
  int period = seek_off;
  {
    const int hit = first_rise(n0 + start + seek_off, min(remaining - seek_off, max_seek) - 1);
    if (hit >= 0) period = seek_off + hit;
  }
  double* out = B.fs + (size_t)FS_AUTOCORR * B.TF + slot;
  if (!remaining || period >= remaining) { if (lane == 0) *out = 0.0; return; }   // :2361-2365

  const int width = min(remaining, P.ac_width);
  const int lo = period / 2;
  double r0 = 0.0, best;                           // the result is floored at 0 (Autocorrelation.cpp:96-103)
  if (F32) {
    // raw mono samples of the conditioned window: mdata() without its scale (trim and padding as there)
    float* x = xf[wid];
    for (int k = lane; k < ACF_XS; k += 32) {
      const int j = n0 + start + k - st.start_off;
      x[k] = (k < width && j >= 0 && j < st.audible) ? __ldg(mono + st.lead + j) : 0.0f;
    }
    __syncwarp();
    best = ac_rounds_f32(x, width, (width + ALF - 1) / ALF, lane, lo, r0);
  } else {
    double* x = xs[wid];
    for (int k = lane; k < AC_XS; k += 32) x[k] = (k < width) ? mdata(mono, st, n0 + start + k) : 0.0;
    __syncwarp();
    best = ac_rounds(x, width, (width + AL - 1) / AL, lane, lo, r0);   // lag groups; the last one may be partial (zero padded)
  }
  best = warp_max(best);
  r0 = __shfl_sync(0xffffffffu, r0, 0);            // lane 0 owns group 0
  // normalisation by R[0] > 0 is monotonic, so max_i (R[i] / R[0]) == (max_i R[i]) / R[0] exactly
  if (lane == 0) *out = (r0 != 0) ? best / r0 : best;
}

void afx_launch_autocorr(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  static const bool fp64 = [] { const char* e = getenv("AFX_AUTOCORR_FP64"); return e && atoi(e) != 0; }();
  if (fp64) k_autocorr<false><<<(B.g_slots + AW - 1) / AW, AW * 32, 0, s>>>(B, P);
  else k_autocorr<true><<<(B.g_slots + AW - 1) / AW, AW * 32, 0, s>>>(B, P);
  ++*launches;
}

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
#include "afx_fft16.cuh"
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

// -------------------------------------------------------------------------------------------------
// FFT form (round 2, default): the 529 x 529 / 2 products are a linear correlation, and since the FP32 demotion nothing ties
// the kernel to the reference's summation order any more, so R = IFFT(|FFT(x)|^2) on the window zero-padded to 1024:
//   z[m] = x[2m] + i x[2m+1]  ->  512-point complex FFT  ->  bin-pair unpack to 2 X[k]  ->  P[k] = |2 X[k]|^2
//   V[k] = (E - D s_k, -D c_k),  V[512-k] = (E + D s_k, -D c_k)   with E = P[k] + P[512-k], D = P[k] - P[512-k],
//   (c_k, s_k) = (cos, sin)(2 pi k / 1024)                            (the real-output inverse as ONE more forward transform)
//   y = FFT512(V):  R[2m] = Re y[m],  R[2m+1] = -Im y[m]             (x 4096: the scale cancels in R[i] / R[0])
// The circular correlation of length 1024 is the linear one for lags < 1024 - (width - 1) = 496; the 33 lags above that
// (at most 33 products each) are summed directly.  ~60 kflop per frame instead of 280 k; FP32 error of R[i] / R[0] measured
// against FP64 sums: <= 2.1e-7 absolute, <= 0.17 of the tolerance 1e-6 + 1e-4 |v| (numpy emulation over 4000 windows incl.
// DC offsets and quiet files, profiles/README.md).  AFX_AUTOCORR_DIRECT=1 keeps the direct FP32 sums, AFX_AUTOCORR_FP64=1
// the FP64 ones.  Transform: radix 16 x 16 x 2 (afx_fft16.cuh's scheme in float2), one warp per frame, 16 points per lane;
// the first pass knows that inputs 9..15 of every butterfly are zero padding, the last pass of the second transform keeps
// its outputs in registers (only lags < 496 are wanted) and takes the maximum right there.
__device__ __forceinline__ float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 c_mul(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ void c_r4(float2& a0, float2& a1, float2& a2, float2& a3)
{
  const float2 s02 = c_add(a0, a2), d02 = c_sub(a0, a2), s13 = c_add(a1, a3), d13 = c_sub(a1, a3);
  const float2 md = make_float2(d13.y, -d13.x);          // -i (a1 - a3)
  a0 = c_add(s02, s13); a1 = c_add(d02, md); a2 = c_sub(s02, s13); a3 = c_sub(d02, md);
}
#define CF_C1 0.92387953251128673848f
#define CF_S1 0.38268343236508978178f
#define CF_H 0.70710678118654752440f
__device__ __forceinline__ float2 c_w16_1(float2 a) { return make_float2(fmaf(a.x, CF_C1, a.y * CF_S1), fmaf(a.y, CF_C1, -(a.x * CF_S1))); }
__device__ __forceinline__ float2 c_w16_2(float2 a) { return make_float2((a.x + a.y) * CF_H, (a.y - a.x) * CF_H); }
__device__ __forceinline__ float2 c_w16_3(float2 a) { return make_float2(fmaf(a.x, CF_S1, a.y * CF_C1), fmaf(a.y, CF_S1, -(a.x * CF_C1))); }
__device__ __forceinline__ float2 c_w16_4(float2 a) { return make_float2(a.y, -a.x); }
__device__ __forceinline__ float2 c_w16_6(float2 a) { return make_float2((a.y - a.x) * CF_H, -(a.x + a.y) * CF_H); }
// twiddles + row transforms of the 4 x 4 decomposition (the column transforms come first and differ, see below)
__device__ __forceinline__ void c_dft16_rows(float2 (&v)[16])
{
  v[5] = c_w16_1(v[5]);  v[9] = c_w16_2(v[9]);   v[13] = c_w16_3(v[13]);
  v[6] = c_w16_2(v[6]);  v[10] = c_w16_4(v[10]); v[14] = c_w16_6(v[14]);
  v[7] = c_w16_3(v[7]);  v[11] = c_w16_6(v[11]);
  { const float2 a = v[15]; v[15] = make_float2(-fmaf(a.x, CF_C1, a.y * CF_S1), fmaf(a.x, CF_S1, -(a.y * CF_C1))); }   // W16^9
  c_r4(v[0], v[1], v[2], v[3]);
  c_r4(v[4], v[5], v[6], v[7]);
  c_r4(v[8], v[9], v[10], v[11]);
  c_r4(v[12], v[13], v[14], v[15]);
}
// 16-point forward DFT in registers; X[m + 4 n] ends up in v[4 m + n] (FFT_REG16)
__device__ __forceinline__ void c_dft16(float2 (&v)[16])
{
  c_r4(v[0], v[4], v[8], v[12]);
  c_r4(v[1], v[5], v[9], v[13]);
  c_r4(v[2], v[6], v[10], v[14]);
  c_r4(v[3], v[7], v[11], v[15]);
  c_dft16_rows(v);
}
// the same with inputs 9..15 known to be zero: column 0 holds (v0, v4, v8, 0), columns 1..3 hold (v_c, v_c+4, 0, 0)
__device__ __forceinline__ void c_dft16_pruned(float2 (&v)[16])
{
  {
    const float2 s02 = c_add(v[0], v[8]), d02 = c_sub(v[0], v[8]), a1 = v[4], md = make_float2(a1.y, -a1.x);
    v[0] = c_add(s02, a1); v[4] = c_add(d02, md); v[8] = c_sub(s02, a1); v[12] = c_sub(d02, md);
  }
#pragma unroll
  for (int c = 1; c < 4; ++c) {
    const float2 a0 = v[c], a1 = v[c + 4], md = make_float2(a1.y, -a1.x);
    v[c] = c_add(a0, a1); v[c + 4] = c_add(a0, md); v[c + 8] = c_sub(a0, a1); v[c + 12] = c_sub(a0, md);
  }
  c_dft16_rows(v);
}
#define ACT_T2 0            // [15][16] exp(-2 pi i r k / 256), r = 1..15   (AfxTables::ac_tw, float2)
#define ACT_T3 240          // [256]    exp(-2 pi i j / 512)
#define ACT_UN 496          // [257]    exp(-2 pi i k / 1024)
#define ACB 544             // float2 per frame: 512 + one pad per 16 (FFT_PHYS)
#define ACX 576             // staged window floats per frame (zero from `width` on; the packed pairs read up to 575)
#define AC_LIN 496          // lags below this come out of the 1024-point circular correlation unaliased

// passes 1 and 2 of the 512-point transform (16 x 16); the radix-2 pass that follows is left to the caller
template <bool PRUNED>
__device__ __forceinline__ void acf_fft_p12(float2 (&v)[16], float2* __restrict__ buf, const float2* __restrict__ tw, int lane)
{
  if (PRUNED) c_dft16_pruned(v); else c_dft16(v);
#pragma unroll
  for (int q = 0; q < 16; ++q) buf[lane * 17 + q] = v[FFT_REG16(q)];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = buf[FFT_PHYS(lane + 32 * r)];
  __syncwarp();
  const int k = lane & 15;
#pragma unroll
  for (int r = 1; r < 16; ++r) v[r] = c_mul(v[r], __ldg(tw + ACT_T2 + (r - 1) * 16 + k));
  c_dft16(v);
  const int base = (lane - k) * 16 + k;
#pragma unroll
  for (int q = 0; q < 16; ++q) buf[FFT_PHYS(base + q * 16)] = v[FFT_REG16(q)];
  __syncwarp();
}

// x: the staged window (ACX floats, zero from `width` on); returns max(R[i]) over lags lo <= i < width (floored at 0), r0 = R[0]
__device__ __forceinline__ float ac_fft(const float* __restrict__ x, float2* __restrict__ buf, const float2* __restrict__ tw,
                                        int width, int lane, int lo, float& r0)
{
  float2 v[16];
#pragma unroll
  for (int r = 0; r < 9; ++r) v[r] = reinterpret_cast<const float2*>(x)[lane + 32 * r];
#pragma unroll
  for (int r = 9; r < 16; ++r) v[r] = make_float2(0.f, 0.f);
  acf_fft_p12<true>(v, buf, tw, lane);
  // pass 3 (radix 2), in place: butterfly i owns slots i and i + 256
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const int i = lane + 32 * b;
    const float2 a = buf[FFT_PHYS(i)], c = c_mul(buf[FFT_PHYS(i + 256)], __ldg(tw + ACT_T3 + i));
    buf[FFT_PHYS(i)] = c_add(a, c); buf[FFT_PHYS(i + 256)] = c_sub(a, c);
  }
  __syncwarp();
  // unpack to the power spectrum and pack the inverse's input, in place: the pair (k, 512 - k) belongs to one lane
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    const int k = (c < 8) ? lane + 32 * c : 256;
    if (c == 8 && lane != 0) break;
    const float2 zk = buf[FFT_PHYS(k)], zr = buf[FFT_PHYS((512 - k) & 511)];
    const float2 E = make_float2(zk.x + zr.x, zk.y - zr.y);          // Z[k] + conj(Z[512 - k])
    const float2 O = make_float2(zk.y + zr.y, zr.x - zk.x);          // (Z[k] - conj(Z[512 - k])) / i
    const float2 w = __ldg(tw + ACT_UN + k);                         // (cos, -sin)(2 pi k / 1024)
    const float2 T = c_mul(w, O);
    const float2 Xa = c_add(E, T), Xb = c_sub(E, T);                 // 2 X[k], conj(2 X[512 - k])
    const float Pa = fmaf(Xa.x, Xa.x, Xa.y * Xa.y), Pb = fmaf(Xb.x, Xb.x, Xb.y * Xb.y);
    const float Es = Pa + Pb, D = Pa - Pb, Ds = D * -w.y, Dc = D * w.x;
    buf[FFT_PHYS(k)] = make_float2(Es - Ds, -Dc);
    if (k > 0 && k < 256) buf[FFT_PHYS(512 - k)] = make_float2(Es + Ds, -Dc);
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = buf[FFT_PHYS(lane + 32 * r)];
  __syncwarp();
  acf_fft_p12<false>(v, buf, tw, lane);
  // last pass: only y[i], i < 248, is wanted (lags 2 i and 2 i + 1 below AC_LIN): the maximum is taken from the registers
  const int wlim = min(width, AC_LIN);
  float best = 0.f;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const int i = lane + 32 * b;
    const float2 a = buf[FFT_PHYS(i)], c = c_mul(buf[FFT_PHYS(i + 256)], __ldg(tw + ACT_T3 + i));
    const float re = a.x + c.x, im = -(a.y + c.y);
    if (b == 0) r0 = re;                                             // lane 0: R[0]
    if (2 * i >= lo && 2 * i < wlim) best = fmaxf(best, re);
    if (2 * i + 1 >= lo && 2 * i + 1 < wlim) best = fmaxf(best, im);
  }
  // the aliased lags, directly: lag AC_LIN + lane (and 528 on lane 0); x is zero from `width` on, so the sums end by themselves
  if (width > AC_LIN) {
    float acc = 0.f;
    const float* __restrict__ xl = x + AC_LIN + lane;
#pragma unroll 11
    for (int j = 0; j < 33; ++j) acc = fmaf(x[j], xl[j], acc);
    const float scale = 4096.f;                                      // the transforms' R carries 2^2 (unpack) x 1024 (no 1 / N)
    if (AC_LIN + lane >= lo && AC_LIN + lane < width) best = fmaxf(best, acc * scale);
    if (lane == 0 && 528 >= lo && 528 < width) best = fmaxf(best, x[0] * x[528] * scale);
  }
  return best;
}

// MODE 0: FP64 direct sums, 1: FP32 direct sums, 2: FP32 FFT form
template <int MODE>
__global__ void __launch_bounds__(AW * 32) k_autocorr(AfxBatchDev B, AfxParams P)
{
  __shared__ double xs[MODE == 0 ? AW : 1][AC_XS];
  __shared__ float xf[MODE == 0 ? 1 : AW][MODE == 2 ? ACX : ACF_XS];
  __shared__ float2 fb[MODE == 2 ? AW : 1][MODE == 2 ? ACB : 1];

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
  if (MODE == 2) {
    float* x = xf[wid];
    for (int k = lane; k < ACX; k += 32) {
      const int j = n0 + start + k - st.start_off;
      x[k] = (k < width && j >= 0 && j < st.audible) ? __ldg(mono + st.lead + j) : 0.0f;
    }
    __syncwarp();
    float r0f = 0.f;
    best = (double)ac_fft(x, fb[wid], P.t.ac_tw, width, lane, lo, r0f);
    r0 = (double)r0f;
  } else if (MODE == 1) {
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
  static const int mode = [] {
    const char* e = getenv("AFX_AUTOCORR_FP64"); if (e && atoi(e) != 0) return 0;
    e = getenv("AFX_AUTOCORR_DIRECT"); return (e && atoi(e) != 0) ? 1 : 2;
  }();
  const int grid = (B.g_slots + AW - 1) / AW;
  if (mode == 0) k_autocorr<0><<<grid, AW * 32, 0, s>>>(B, P);
  else if (mode == 1) k_autocorr<1><<<grid, AW * 32, 0, s>>>(B, P);
  else k_autocorr<2><<<grid, AW * 32, 0, s>>>(B, P);
  ++*launches;
}

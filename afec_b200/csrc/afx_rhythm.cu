// K7: rhythm -- the 512 / 128 onset front end and the per-file back end.
//
// Reference: TSampleAnalyser::AnalyzeLowLevelDescriptors rhythm part (SampleAnalyser.cpp:983-1048),
// TRhythmTracker (Source/Crawler/FeatureExtraction/Source/RhythmTracker.cpp:46-117 front end,
// :121-660 back end), TOnsetFftProcessor / TOnsetDetector (Source/Core/AudioTypes/Source/
// OnsetDetector.cpp:96-243, 371-590), TCannyWindow (CannyWindow.cpp:27-80) and aubio's beat tracker
// (3rdParty/Aubio/Dist/src/tempo/beattracking.c:59-110, 126-262, 286-410, 424-441; mathutils.c:652-666).
//
// The reference walks the rhythm frames of a file one by one, but only three things really are
// sequential: the per-bin whitening peak memory, the onset min-gap state machine and the valley
// tracking of the contrast measure.  Everything else is a function of a few neighbouring frames and is
// computed frame-parallel over the whole batch:
//
//   k_rhythm_polar   (warp per frame)  Hann(512) x frame -> 512-pt real FFT (FP64) -> float32 polar row
//                                      [mag 0..254 | dc | phase 0..254] (bins 0..254 + Re[0]; quirk: the
//                                      "Nyquist" slot of the reference is Im[0] == 0)
//   k_rhythm_whiten  (thread per file x bin, sequential over frames) adaptive-max whitening, in place
//   k_rhythm_odf     (warp per 8 frames) rectified complex-domain onset function from rows t, t-1, t-2 (carried in
//                                      registers); k_rhythm_power (thread per frame) the power onset function;
//                                      float32 operation order of the reference
//   k_rhythm_median  (warp per frame)  running median of the last 69 ODF values -> post = odf - median
//   k_rhythm_back    (CTA per file)    min-gap peak picker -> onset series, onset count, Canny
//                                      sharpening + z-score, peak strength / frequency / contrast, beat
//                                      tracker (ACF, comb filterbank, Rayleigh weighting), tempo heuristics
#include "afx_fft16.cuh"
#include "afx_select.cuh"
#include "../../include/afec_b200.h"

#define RPW 8              // frames (warps) per CTA in the frame-parallel kernels
#define MEDSPAN 69         // int(44100 * 0.2 / 128 + 0.5), OnsetDetector.cpp:280-282
#define BT_THREADS 256

// OnsetDetector.cpp:19-23 (float32; explicit rn intrinsics keep nvcc from contracting to FMA)
__device__ __forceinline__ float phase_rewrap(float p)
{
  const float pi = (float)3.14159265358979323846, twopi = (float)6.2831853071795864769252867665590,
    inv2pi = (float)0.15915494309189533576888376337251;
  if (p > -pi && p < pi) return p;
  const float k = __fadd_rn(1.f, floorf(__fmul_rn(__fsub_rn(-pi, p), inv2pi)));
  return __fadd_rn(p, __fmul_rn(twopi, k));
}

// atan2 for the float32 phase column.  The reference computes atan2 in double and rounds to float
// (OnsetDetector.cpp:136-155); the result only has to be right to well below a float ulp (2.4e-7 at pi), so an
// octant reduction to |w| <= tan(pi/8), one reciprocal-seeded division (MUFU.RCP64H + one Newton step: ~1e-13 of the
// quotient) and atan(w) = w - w^3 Q(w^2) with Q the degree-6 interpolant of (w - atan w) / w^3 at the Chebyshev nodes of
// [0, tan^2(pi/8)] (error of atan < 1.2e-12; the Taylor series needs w^21 for 6e-11) replace libdevice's fully rounded
// double atan2 -- a fifth of its FP64 work.  The FP64 constants sit in constant memory: as literals every use costs
// two UMOVs, and this kernel is issue bound.
__constant__ double c_at[16] = { 0.04043224825887161, -0.07135325122330678, 0.09028983500350463, -0.11107495135714474,
                                 0.14285612511387016, -0.1999999891728858, 0.3333333333144073, 0.0, 0.0, 0.0,
                                 0.41421356237309503, 0.78539816339744830962, 1.57079632679489661923,
                                 3.14159265358979323846, 1.0, 0.5 };
__device__ __forceinline__ double atan2_phase(double y, double x)
{
  const double ax = fabs(x), ay = fabs(y);
  const bool sw = ay > ax;
  const double mx = sw ? ay : ax, mn = sw ? ax : ay;
  const bool hi = mn > c_at[10] * mx;                       // beyond pi/8: atan(z) = pi/4 + atan((z - 1) / (z + 1))
  const double num = hi ? mn - mx : mn;
  double den = hi ? mn + mx : mx;
  den = (mx == 0.0) ? c_at[14] : den;                       // atan2(0, 0) = 0: num is 0 there
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));   // den is 0 or far above the subnormals (float32 samples)
  r = fma(fma(-den, r, c_at[14]), r, r);
  const double w = num * r;
  const double w2 = w * w;
  double p = c_at[0];
#pragma unroll
  for (int i = 1; i < 7; ++i) p = fma(p, w2, c_at[i]);
  r = fma(-(p * w2), w, w);                                 // w - w^3 Q(w^2)
  if (hi) r += c_at[11];
  if (sw) r = c_at[12] - r;
  if (x < 0.0) r = c_at[13] - r;
  return __hiloint2double(__double2hiint(r) | (__double2hiint(y) & 0x80000000), __double2loint(r));   // copysign(r, y), r >= 0
}
// sqrt of a squared magnitude to ~2^-44 relative (MUFU.RSQ64H seed + one Newton step); the result is rounded to float.
// s is exactly 0 or far above the subnormals (squares of sums of float32 samples)
__device__ __forceinline__ double sqrt_mag_r(double s)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  y = (s > 0.0) ? y : 0.0;
  const double r = s * y;
  return fma(fma(-r, r, s), c_at[15] * y, r);
}

// -------------------------------------------------------------------------------------------------
// 16 threads per rhythm frame (256-point packed transform = radix 16 x 16 in registers), two frames per warp.
// The real unpack works on bin pairs (k, 256 - k):  2 X[k] = E + T,  2 X[256 - k] = conj(E - T)  with
// E = Z[k] + conj(Z[256 - k]),  T = W^k (Z[k] - conj(Z[256 - k])) / i;  the window table is pre-halved so that E + T
// is X[k] itself.  The mirror of bin 0 is the Nyquist bin, which the polar row does not hold (quirk,
// OnsetDetector.cpp:136-155), so that slot takes bin 128.
#define PF 8                // frames per CTA (4 warps)
__device__ __forceinline__ void polar_store(float* __restrict__ row, int k, double2 X)
{
  row[k] = (float)sqrt_mag_r(fma(X.x, X.x, X.y * X.y));                           // OnsetDetector.cpp:136-155
  row[256 + k] = (float)atan2_phase(X.y, X.x);
}
// windowed frame of one 16-thread group: v[r] = packed samples (2m, 2m + 1), m = ht + 16 r   (OnsetDetector.cpp:119-120)
__device__ __forceinline__ void polar_load(double2 (&v)[16], const float* __restrict__ mono, const AfxState& st, int n0, bool live, int ht,
                                           const double2* __restrict__ win2)
{
  const int j0 = n0 - st.start_off;                 // frame start relative to the first audible sample
  const float* __restrict__ src = mono + st.lead + j0;
  if (live && j0 >= 0 && j0 + AFX_RFFT <= st.audible) {     // whole frame inside the audible span (the common case)
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int m = ht + 16 * r;
      const double2 w = __ldg(win2 + m);
      v[r] = make_double2(w.x * ((double)__ldg(src + 2 * m) * st.fs), w.y * ((double)__ldg(src + 2 * m + 1) * st.fs));
    }
  } else {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int m = ht + 16 * r;
      const double2 w = __ldg(win2 + m);
      const double x0 = live ? mdata(mono, st, n0 + 2 * m) : 0.0, x1 = live ? mdata(mono, st, n0 + 2 * m + 1) : 0.0;
      v[r] = make_double2(w.x * x0, w.y * x1);
    }
  }
}
// real unpack of the packed transform in buf + polar conversion into row[0..254] (mag) | row[255] (dc) | row[256..510] (phase)
__device__ __forceinline__ void polar_unpack(float* __restrict__ row, const double2* __restrict__ buf, int ht, const double2* __restrict__ tw512)
{
#pragma unroll 2                                             // the body holds two inlined atan2 + sqrt: keep the code small
  for (int c = 0; c < 8; ++c) {
    const int k = ht + 16 * c;                               // 0..127, mirror 256 - k
    const double2 zk = buf[FFT_PHYS(k)], zc = buf[FFT_PHYS((256 - k) & 255)];
    const double2 E = make_double2(zk.x + zc.x, zk.y - zc.y);
    const double2 O = make_double2(zk.y + zc.y, zc.x - zk.x);
    const double2 T = f_mul(__ldg(tw512 + k), O);
    double2 Xa = f_add(E, T);
    double2 Xb = make_double2(E.x - T.x, T.y - E.y);        // conj(E - T)
    int kb = 256 - k;
    if (c == 0 && ht == 0) {
      Xa.y = 0.0; row[255] = (float)Xa.x;                    // mDC, OnsetDetector.cpp:146
      const double2 z = buf[FFT_PHYS(128)];
      Xb = make_double2(2.0 * z.x, -2.0 * z.y); kb = 128;
    }
    if (kb < AFX_RBINS) polar_store(row, kb, Xb);
    polar_store(row, k, Xa);
  }
}
// A CTA walks PI consecutive groups of PF frames: the slot -> file -> state chain of dependent loads in front of a frame's
// samples (two levels, ~1.5 k cycles: a tenth of a one-group CTA's life, ncu) is fetched for the NEXT group while the
// current one is transformed.
#define PI 4
struct PolarMeta { int fi, t; bool live; int start_off, audible, lead; long long mono_off; double fs; };
__device__ __forceinline__ PolarMeta polar_meta(const AfxBatchDev& B, int rel)
{
  PolarMeta m;
  const bool in_range = rel < B.g_rslots;
  const int slot = B.rslot0 + (in_range ? rel : 0);
  m.fi = B.rslot_file[slot];
  const AfxFile* __restrict__ fp = B.files + m.fi;
  const AfxState* __restrict__ sp = B.state + m.fi;
  m.t = slot - fp->rframe_off;
  m.live = in_range && fp->status == 0 && m.t < sp->Fr;     // both halves of a warp run the transform (warp-wide sync)
  m.start_off = sp->start_off; m.audible = sp->audible; m.lead = sp->lead; m.fs = sp->fs; m.mono_off = fp->mono_off;
  return m;
}
__global__ void __launch_bounds__(PF * 16, 5) k_rhythm_polar(AfxBatchDev B, AfxParams P)
{
  __shared__ double2 sbuf[PF][256 + 16];
  const int h = threadIdx.x >> 4, ht = threadIdx.x & 15;
  double2* buf = sbuf[h];
  int rel = blockIdx.x * (PF * PI) + h;
  PolarMeta nx = polar_meta(B, rel);
#pragma unroll 1
  for (int it = 0; it < PI; ++it, rel += PF) {
    if (rel - h >= B.g_rslots) break;                          // uniform: the whole group lies past the end
    const PolarMeta m = nx;
    AfxState st;                                               // the fields polar_load / mdata read
    st.start_off = m.start_off; st.audible = m.audible; st.lead = m.lead; st.fs = m.fs;
    double2 v[16];
    polar_load(v, B.mono + m.mono_off, st, m.t * AFX_RHOP, m.live, ht, reinterpret_cast<const double2*>(P.t.rwindow));
    if (it + 1 < PI) nx = polar_meta(B, rel + PF);
    fft16_run<256>(v, buf, FftTw{ P.t.fft_t2, nullptr }, ht, FftSyncWarp());
    if (m.live) polar_unpack(B.rpolar + (size_t)rel * AFX_RROW, buf, ht, P.t.tw512);
    __syncwarp();                                              // the warp's two buffers are read before the next transform writes them
  }
}

// -------------------------------------------------------------------------------------------------
// adaptive-max whitening (OnsetDetector.cpp:193-243): psp is a per-bin recurrence over the file's frames
__global__ void __launch_bounds__(256) k_rhythm_whiten(AfxBatchDev B, AfxParams P)
{
  const int fi = B.file_order[B.file0 + blockIdx.x];
  const AfxFile f = B.files[fi];
  if (f.status != 0) return;
  const int Fr = B.state[fi].Fr;
  if (Fr <= 0) return;
  float* col = B.rpolar + (size_t)(f.rframe_off - B.rslot0) * AFX_RROW + threadIdx.x;   // 0..254 bins, 255 dc
  const double relax = (double)P.r_relax, wfloor = (double)0.1f;
  double psp = 0.0;
  constexpr int D = 8;                        // rows in flight per thread: the recurrence is latency bound otherwise
  float nxt[D];
#pragma unroll
  for (int q = 0; q < D; ++q) nxt[q] = (q < Fr) ? col[(size_t)q * AFX_RROW] : 0.0f;
  for (int t0 = 0; t0 < Fr; t0 += D) {
    float cur[D];
#pragma unroll
    for (int q = 0; q < D; ++q) { cur[q] = nxt[q]; nxt[q] = (t0 + D + q < Fr) ? col[(size_t)(t0 + D + q) * AFX_RROW] : 0.0f; }
#pragma unroll
    for (int q = 0; q < D; ++q) {
      if (t0 + q < Fr) {
        const float v = cur[q];
        double a = (double)fabsf(v);
        if (a < psp) a = __dadd_rn(a, __dmul_rn(__dsub_rn(psp, a), relax));
        psp = a;
        col[(size_t)(t0 + q) * AFX_RROW] = __fdiv_rn(v, (float)(wfloor > psp ? wfloor : psp));
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// onset functions (OnsetDetector.cpp:371-547): kFunctionRComplex (k_rhythm_odf) and kFunctionPower (k_rhythm_power).
// A warp walks OB consecutive rhythm frame slots.  The complex-domain function of frame t needs |mag| of t-1 and the
// phases of t-1 and t-2: they stay in registers from the previous steps (two warm-up rows in front of the run), so
// every polar row is read once, unconditionally -- loads of the next rows are in flight while a frame is evaluated.
#define OB 8                // frames per warp
#define OW 8                // warps per CTA
// one bin of the rectified complex-domain function (OnsetDetector.cpp:440-470): cur / pmv = whitened |mag| of this and the
// previous frame, ph / yp / yp2 = phases of this and the previous two frames
__device__ __forceinline__ float odf_complex_bin(float cur, float pmv, float ph, float yp, float yp2, bool has_prev)
{
  const float ypd = has_prev ? phase_rewrap(__fsub_rn(yp, yp2)) : 0.0f;
  const float pred = __fadd_rn(yp, ypd);
  const float dev = __fsub_rn(pred, ph);
  const float cs = cosf(phase_rewrap(dev));
  const float q = __fsub_rn(__fadd_rn(__fmul_rn(pmv, pmv), __fmul_rn(cur, cur)), __fmul_rn(__fmul_rn(pmv, cur), cs));
  return sqrtf(q);
}
// the power function of one whitened row: float32 sum of the squared bins in bin order (the reference's rounding, :388-396)
__device__ __forceinline__ float odf_power_row(const float4* __restrict__ row, float dc)
{
  float v = __fadd_rn(__fmul_rn(0.0f, 0.0f), __fmul_rn(dc, dc));            // nyq^2 + dc^2
#pragma unroll 8
  for (int i = 0; i < 63; ++i) {
    const float4 q = row[i];
    v = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(v, __fmul_rn(q.x, q.x)), __fmul_rn(q.y, q.y)), __fmul_rn(q.z, q.z)), __fmul_rn(q.w, q.w));
  }
  { const float4 q = row[63]; v = __fadd_rn(__fadd_rn(__fadd_rn(v, __fmul_rn(q.x, q.x)), __fmul_rn(q.y, q.y)), __fmul_rn(q.z, q.z)); }
  return v;
}
__global__ void __launch_bounds__(OW * 32) k_rhythm_odf(AfxBatchDev B, AfxParams P)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rel0 = (blockIdx.x * OW + wid) * OB;
  if (rel0 >= B.g_rslots) return;
  float pm[8], ph1[8], ph2[8];                   // |mag| of the previous row, phases of the previous two rows
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int i = lane + 32 * c;
    const float* r1 = B.rpolar + (size_t)(rel0 - 1) * AFX_RROW;
    const float* r2 = B.rpolar + (size_t)(rel0 - 2) * AFX_RROW;
    pm[c] = (rel0 >= 1) ? fabsf(r1[i]) : 0.0f;
    ph1[c] = (rel0 >= 1) ? r1[256 + i] : 0.0f;
    ph2[c] = (rel0 >= 2) ? r2[256 + i] : 0.0f;
  }
#pragma unroll 1                                 // one copy of the body: eight were an instruction-cache problem
  for (int k = 0; k < OB; ++k) {
    const int rel = rel0 + k;
    if (rel >= B.g_rslots) break;                // warp-uniform
    const int slot = B.rslot0 + rel;
    const int fi = B.rslot_file[slot];
    const int t = slot - B.files[fi].rframe_off;
    if (B.files[fi].status != 0 || t >= B.state[fi].Fr) continue;     // warp-uniform; the next live frame has t == 0
    const float* r0 = B.rpolar + (size_t)rel * AFX_RROW;
    float m[8], ph[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { m[c] = r0[lane + 32 * c]; ph[c] = r0[256 + lane + 32 * c]; }
    double total = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int i = lane + 32 * c;
      const float cur = fabsf(m[c]);
      const float pmv = (t >= 1) ? pm[c] : 0.0f;
      if (i < AFX_RBINS && cur > 0.01f && !(cur < pmv))
        total += (double)odf_complex_bin(cur, pmv, ph[c], (t >= 1) ? ph1[c] : 0.0f, (t >= 2) ? ph2[c] : 0.0f, t >= 1);
      pm[c] = cur; ph2[c] = ph1[c]; ph1[c] = ph[c];
    }
    total = warp_sum(total);
    if (lane == 0) B.rodf[slot] = __fmul_rn((float)total, P.r_norm_complex);
  }
}

// One thread per frame adds its row; a warp's 16-byte loads touch 32 rows, both halves of every sector get used (L1).
__global__ void __launch_bounds__(128) k_rhythm_power(AfxBatchDev B, AfxParams P)
{
  const int rel = blockIdx.x * 128 + threadIdx.x;
  if (rel >= B.g_rslots) return;
  const int slot = B.rslot0 + rel;
  const int fi = B.rslot_file[slot];
  const int t = slot - B.files[fi].rframe_off;
  if (B.files[fi].status != 0 || t >= B.state[fi].Fr) return;
  const float v = odf_power_row(reinterpret_cast<const float4*>(B.rpolar + (size_t)rel * AFX_RROW), B.rpolar[(size_t)rel * AFX_RROW + 255]);
  B.rodf[(size_t)B.TFr + slot] = __fmul_rn(v, P.r_norm_power);
}

// -------------------------------------------------------------------------------------------------
// Whitening + both onset functions as ONE producer / consumer pipeline per file (round 2, the default for launch groups
// with at least two files per SM; the three kernels above remain for smaller groups).  The polar rows are read once and
// never rewritten -- 4 KB of DRAM traffic per rhythm frame (polar write + this read) instead of 7:
//   producers (8 warps, a thread per column 0..254 | dc) walk the file's frames in order with the peak memory in a
//     register, RP_H rows of loads in flight, and put the whitened magnitudes into a shared-memory ring of two halves of
//     RP_H frames; row 0 of a half repeats the last frame of the half before it (the complex-domain function of frame t
//     needs |mag| of frame t - 1);
//   consumers: 8 warps evaluate the complex-domain function of one frame each (phases of frames t, t-1, t-2 straight from
//     global memory, requested before the wait for the producers), a ninth warp runs the in-order float32 power sums,
//     one lane per frame.
// The halves are handed over with named barriers (bar.arrive / bar.sync), one hand-over per RP_H frames, as in
// k_peaks_pipe.  Values are those of the split kernels bit for bit (same device functions, same order of every sum).
#define RP_P 8
#define RP_C 7
#define RP_H 7
#define RP_ROWF 260         // floats per ring row: 256 + 4 (16-byte aligned rows on different bank groups for the lane-per-row power sums)
#define RP_THREADS ((RP_P + RP_C + 1) * 32)
// barrier ids as immediates (a register id makes ptxas reserve all 16 barriers of the CTA)
template <int ID> __device__ __forceinline__ void rp_bar_sync_i() { asm volatile("bar.sync %0, %1;" :: "n"(ID), "n"(RP_THREADS) : "memory"); }
template <int ID> __device__ __forceinline__ void rp_bar_arrive_i() { asm volatile("bar.arrive %0, %1;" :: "n"(ID), "n"(RP_THREADS) : "memory"); }
__device__ __forceinline__ void rp_full_sync(int h) { if (h) rp_bar_sync_i<2>(); else rp_bar_sync_i<1>(); }
__device__ __forceinline__ void rp_full_arrive(int h) { if (h) rp_bar_arrive_i<2>(); else rp_bar_arrive_i<1>(); }
__device__ __forceinline__ void rp_free_sync(int h) { if (h) rp_bar_sync_i<4>(); else rp_bar_sync_i<3>(); }
__device__ __forceinline__ void rp_free_arrive(int h) { if (h) rp_bar_arrive_i<4>(); else rp_bar_arrive_i<3>(); }
__global__ void __launch_bounds__(RP_THREADS, 3) k_rhythm_pipe(AfxBatchDev B, AfxParams P)
{
  __shared__ __align__(16) float ring[2][RP_H + 1][RP_ROWF];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int fi = B.file_order[B.file0 + blockIdx.x];
  const AfxFile f = B.files[fi];
  if (f.status != 0) return;
  const int Fr = B.state[fi].Fr;
  if (Fr <= 0) return;
  const float* __restrict__ rows = B.rpolar + (size_t)(f.rframe_off - B.rslot0) * AFX_RROW;
  const int nb = (Fr + RP_H - 1) / RP_H;                      // batches of RP_H frames; batch b uses ring half b & 1
  // barriers: 1 + h = "half h is full", 3 + h = "half h is free again"
  if (wid < RP_P) {
    const int tp = threadIdx.x;                                // column: bins 0..254, 255 = dc
    const float* __restrict__ col = rows + tp;
    const double relax = (double)P.r_relax, wfloor = (double)0.1f;
    double psp = 0.0;
    float lastw = 0.0f;
    float nxt[RP_H];
#pragma unroll
    for (int q = 0; q < RP_H; ++q) nxt[q] = (q < Fr) ? __ldg(col + (size_t)q * AFX_RROW) : 0.0f;
    for (int b = 0; b < nb; ++b) {
      const int h = b & 1, t0 = b * RP_H;
      if (b >= 2) rp_free_sync(h);                             // the consumers are done with batch b - 2
      ring[h][0][tp] = lastw;
#pragma unroll
      for (int q = 0; q < RP_H; ++q) {
        const float v = nxt[q];
        nxt[q] = (t0 + RP_H + q < Fr) ? __ldg(col + (size_t)(t0 + RP_H + q) * AFX_RROW) : 0.0f;
        if (t0 + q < Fr) {                                     // OnsetDetector.cpp:193-243
          double a = (double)fabsf(v);
          if (a < psp) a = __dadd_rn(a, __dmul_rn(__dsub_rn(psp, a), relax));
          psp = a;
          lastw = __fdiv_rn(v, (float)(wfloor > psp ? wfloor : psp));
          ring[h][q + 1][tp] = lastw;
        }
      }
      rp_full_arrive(h);
    }
  } else if (wid < RP_P + RP_C) {
    const int cw = wid - RP_P;
    float* __restrict__ odf_c = B.rodf + f.rframe_off;
    for (int b = 0; b < nb; ++b) {
      const int h = b & 1, t = b * RP_H + cw;
      float ph[8], ph1[8], ph2[8];
      if (t < Fr) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int i = 256 + lane + 32 * c;
          ph[c] = __ldg(rows + (size_t)t * AFX_RROW + i);
          ph1[c] = (t >= 1) ? __ldg(rows + (size_t)(t - 1) * AFX_RROW + i) : 0.0f;
          ph2[c] = (t >= 2) ? __ldg(rows + (size_t)(t - 2) * AFX_RROW + i) : 0.0f;
        }
      }
      rp_full_sync(h);
      if (t < Fr) {
        const float* r0 = ring[h][cw + 1];
        const float* r1 = ring[h][cw];
        double total = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int i = lane + 32 * c;
          const float cur = fabsf(r0[i]);
          const float pmv = (t >= 1) ? fabsf(r1[i]) : 0.0f;
          if (i < AFX_RBINS && cur > 0.01f && !(cur < pmv))
            total += (double)odf_complex_bin(cur, pmv, ph[c], ph1[c], ph2[c], t >= 1);
        }
        total = warp_sum(total);
        if (lane == 0) odf_c[t] = __fmul_rn((float)total, P.r_norm_complex);
      }
      if (b + 2 < nb) rp_free_arrive(h);
    }
  } else {
    float* __restrict__ odf_p = B.rodf + (size_t)B.TFr + f.rframe_off;
    for (int b = 0; b < nb; ++b) {
      const int h = b & 1, t = b * RP_H + lane;
      rp_full_sync(h);
      if (lane < RP_H && t < Fr) {
        const float* r0 = ring[h][lane + 1];
        odf_p[t] = __fmul_rn(odf_power_row(reinterpret_cast<const float4*>(r0), r0[255]), P.r_norm_power);
      }
      if (b + 2 < nb) rp_free_arrive(h);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// The fused front end: ONE kernel, one CTA per file (longest first), the polar rows never leave shared memory.
// A CTA walks its file in spans of RF_S rhythm frames:
//   1. 16 threads per frame: window, 256-point packed FFT, real unpack, polar conversion -> float32 row in the ring
//   2. whitening (OnsetDetector.cpp:193-243): every thread owns two of the 256 columns and walks the span's rows in
//      frame order; the per-bin peak memories live in registers for the whole file
//   3. onset functions: warps 0..2 evaluate the complex-domain function of the span's frames (one frame per warp and
//      turn), warp 3 runs the in-order float32 power sums, one lane per frame
// The row ring holds RF_S + 2 rows (frame t at position t mod (RF_S + 2)): the two frames before a span stay where the
// previous span left them.  Values are those of the split kernels above bit for bit (same device functions, same order
// of every sum); the split path remains for launch groups with too few files to fill the GPU with one CTA per file.
#define RF_S 8
#define RF_THREADS (RF_S * 16)
#define RF_RING (RF_S + 2)
#define RF_ROW 516          // floats per ring row: 512 + 4 (rows start on different banks for the lane-per-row power sums)
#define RF_SMEM (RF_S * (256 + 16) * 16 + RF_RING * RF_ROW * 4)
__global__ void __launch_bounds__(RF_THREADS, 4) k_rhythm_front(AfxBatchDev B, AfxParams P)
{
  extern __shared__ __align__(16) unsigned char rf_smem[];
  double2* sbuf = reinterpret_cast<double2*>(rf_smem);
  float* ring = reinterpret_cast<float*>(rf_smem + RF_S * (256 + 16) * 16);
  const int tid = threadIdx.x, h = tid >> 4, ht = tid & 15, lane = tid & 31, wid = tid >> 5;
  const int fi = B.file_order[B.file0 + blockIdx.x];
  const AfxFile* __restrict__ fp = B.files + fi;
  if (fp->status != 0) return;
  const AfxState st = B.state[fi];
  const int Fr = st.Fr;
  if (Fr <= 0) return;
  const float* __restrict__ mono = B.mono + fp->mono_off;
  const double2* __restrict__ win2 = reinterpret_cast<const double2*>(P.t.rwindow);
  double2* buf = sbuf + h * (256 + 16);
  float* __restrict__ odf_c = B.rodf + fp->rframe_off;
  float* __restrict__ odf_p = B.rodf + (size_t)B.TFr + fp->rframe_off;
  const double relax = (double)P.r_relax, wfloor = (double)0.1f;
  double psp0 = 0.0, psp1 = 0.0;                            // peak memories of columns tid and tid + 128 (255 = dc)
  for (int t0 = 0; t0 < Fr; t0 += RF_S) {
    const int nlive = min(RF_S, Fr - t0);
    // ---- 1. transform (both halves of a warp run it: warp-wide sync inside)
    {
      const bool live = h < nlive;
      double2 v[16];
      polar_load(v, mono, st, (t0 + h) * AFX_RHOP, live, ht, win2);
      fft16_run<256>(v, buf, FftTw{ P.t.fft_t2, nullptr }, ht, FftSyncWarp());
      __syncthreads();                                       // the previous span's onset functions are done with the ring
      if (live) polar_unpack(ring + ((t0 + h) % RF_RING) * RF_ROW, buf, ht, P.t.tw512);
    }
    __syncthreads();
    // ---- 2. whitening, in place
    for (int k = 0; k < nlive; ++k) {
      float* row = ring + ((t0 + k) % RF_RING) * RF_ROW;
      const float v0 = row[tid], v1 = row[tid + 128];
      double a0 = (double)fabsf(v0), a1 = (double)fabsf(v1);
      if (a0 < psp0) a0 = __dadd_rn(a0, __dmul_rn(__dsub_rn(psp0, a0), relax));
      if (a1 < psp1) a1 = __dadd_rn(a1, __dmul_rn(__dsub_rn(psp1, a1), relax));
      psp0 = a0; psp1 = a1;
      row[tid] = __fdiv_rn(v0, (float)(wfloor > psp0 ? wfloor : psp0));
      row[tid + 128] = __fdiv_rn(v1, (float)(wfloor > psp1 ? wfloor : psp1));
    }
    __syncthreads();
    // ---- 3. onset functions
    if (wid < 3) {
#pragma unroll 1
      for (int k = wid; k < nlive; k += 3) {
        const int t = t0 + k;
        const float* r0 = ring + (t % RF_RING) * RF_ROW;
        const float* r1 = ring + ((t + RF_RING - 1) % RF_RING) * RF_ROW;
        const float* r2 = ring + ((t + RF_RING - 2) % RF_RING) * RF_ROW;
        double total = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int i = lane + 32 * c;
          const float cur = fabsf(r0[i]);
          const float pmv = (t >= 1) ? fabsf(r1[i]) : 0.0f;
          if (i < AFX_RBINS && cur > 0.01f && !(cur < pmv))
            total += (double)odf_complex_bin(cur, pmv, r0[256 + i], (t >= 1) ? r1[256 + i] : 0.0f, (t >= 2) ? r2[256 + i] : 0.0f, t >= 1);
        }
        total = warp_sum(total);
        if (lane == 0) odf_c[t] = __fmul_rn((float)total, P.r_norm_complex);
      }
    } else if (lane < nlive) {
      const int t = t0 + lane;
      const float* r0 = ring + (t % RF_RING) * RF_ROW;
      odf_p[t] = __fmul_rn(odf_power_row(reinterpret_cast<const float4*>(r0), r0[255]), P.r_norm_power);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// median removal (OnsetDetector.cpp:551-575): post = odf[t] - median(odf[t-68 .. t]), zeros before the file.
// The windows of consecutive frames differ by one value, so a thread walks a chunk of frames with the sorted
// window held in 69 REGISTERS and replaces the outgoing value by the incoming one with two branch-free
// compare/select sweeps (static indices only): ~300 FP32 ops per frame instead of 69 x 69 comparisons.
// A chunk that starts inside a file first replays the 68 frames before it.
#define ML 128             // frames per chunk
#define MED_INF __int_as_float(0x7f800000)

__device__ __forceinline__ void med_replace(float (&s)[MEDSPAN], float out, float in)
{
  // drop one copy of `out`: everything at or after its first position moves down by one
#pragma unroll
  for (int i = 0; i < MEDSPAN - 1; ++i) s[i] = (s[i] < out) ? s[i] : s[i + 1];
  s[MEDSPAN - 1] = MED_INF;
  // insert `in`: s'[i] = max(r[i-1], min(r[i], in))
#pragma unroll
  for (int i = MEDSPAN - 1; i > 0; --i) s[i] = fmaxf(s[i - 1], fminf(s[i], in));
  s[0] = fminf(s[0], in);
}

__global__ void __launch_bounds__(64) k_rhythm_median(AfxBatchDev B)
{
  const int nchunks = (B.g_rslots + ML - 1) / ML;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= 2 * nchunks) return;
  const int ty = gid & 1, chunk = gid >> 1;
  const float* __restrict__ odf = B.rodf + (size_t)ty * B.TFr;
  float* __restrict__ post = B.rpost + (size_t)ty * B.TFr;
  int slot = B.rslot0 + chunk * ML;
  const int s_end = min(slot + ML, B.rslot0 + B.g_rslots);
  int fi = B.rslot_file[slot];
  float s[MEDSPAN];
  while (slot < s_end && fi < B.n_files) {
    const AfxFile f = B.files[fi];
    const int Fr = (f.status == 0) ? B.state[fi].Fr : 0;
    const int file_end = f.rframe_off + f.rframe_cap;
    int t = slot - f.rframe_off;
    if (t >= 0 && t < Fr) {
      const float* x = odf + f.rframe_off;
      float* y = post + f.rframe_off;
      bool primed = false;
      if (t == 0) {
#pragma unroll
        for (int i = 0; i < MEDSPAN; ++i) s[i] = 0.0f;         // the detector's history starts as zeros
        primed = true;
      } else {
#pragma unroll
        for (int i = 0; i < MEDSPAN; ++i) s[i] = MED_INF;
        for (int j = t - (MEDSPAN - 1); j < t; ++j) med_replace(s, MED_INF, j >= 0 ? x[j] : 0.0f);
      }
      for (; t < Fr && slot < s_end; ++t, ++slot) {
        const float out = primed ? ((t >= MEDSPAN) ? x[t - MEDSPAN] : 0.0f) : MED_INF;
        primed = true;
        const float in = x[t];
        med_replace(s, out, in);
        y[t] = __fsub_rn(in, s[(MEDSPAN - 1) / 2]);
      }
    }
    if (slot < file_end) slot = file_end;      // unused capacity slots of this file
    ++fi;
  }
}

// -------------------------------------------------------------------------------------------------
// back end helpers
// TAudioMath::SamplesToMs / MsToSamples, AudioMath.inl:127-137 (float32)
__device__ __forceinline__ float r_samples_to_ms(int sr, int samples) { return __fdiv_rn((float)samples, __fdiv_rn((float)sr, 1000.0f)); }
__device__ __forceinline__ int r_ms_to_samples(int sr, float ms)
{
  const float v = __fmul_rn(__fdiv_rn((float)sr, 1000.0f), ms);
  return (int)__fadd_rn(v, signbit(v) ? -0.5f : 0.5f);
}

// RhythmTracker.cpp:502-555
__device__ double r_guess_beats(double min_bpm, double dur)
{
  const double beat = 60.0 / min_bpm, bar = 4.0 * beat;
  if (dur < beat) return 0.0;
  if (dur < bar) {
    for (int div = 2; div >= 1; div /= 2) { const double d = (4.0 / (double)div) * beat; if (4 % div == 0 && dur < d) return (double)(float)(4.0 / (double)div); }
    return 4.0;
  }
  for (int bars = 1; bars <= 8; bars *= 2) { const double nb = 4.0 * bars; if (dur / nb < beat) return (double)(float)nb; }
  return 0.0;
}

// RhythmTracker.cpp:559-603
__device__ double r_onset_match_conf(int sr, const double* raw, int n, double off_s, double nbeats, double tempo, double thr)
{
  const int off = r_ms_to_samples(sr, (float)(off_s * 1000));
  const double spb = 60.0 / tempo * sr;
  const int range = (int)(spb / 32) / AFX_RHOP;
  double strength = 0;
  for (int i = 0; i < nbeats * 2; ++i) {
    const int t = (int)(i * spb / 2.0) + off;
    const int idx = ((t + AFX_RHOP / 2) / AFX_RHOP);
    double peak = 0.0;
    for (int j = idx - range; j < idx + range; ++j) if (j >= 0 && j < n) peak = peak > raw[j] ? peak : raw[j];
    if (peak >= thr) strength += 1.0;
  }
  const double v = strength / (nbeats * 2) * 2.0;
  return v < 1.0 ? v : 1.0;
}

// block-wide sum / max of one double (blockDim.x == BT_THREADS); result broadcast
// nine consecutive lags of the beat tracker's autocorrelation (mathutils.c:652-666) on one thread: a sliding
// 9-sample register window turns every pair of shared-memory loads into 9 DFMAs (same tiling as k_autocorr).
// x must be readable (zero) up to index n + 17.
#define RB_AL 9
__device__ __noinline__ void rb_acf_group(const double* __restrict__ x, int n, int g, double* __restrict__ acf)
{
  const int i0 = g * RB_AL;
  const int nj = n - i0;
  double acc[RB_AL], w[RB_AL];
#pragma unroll
  for (int q = 0; q < RB_AL; ++q) { acc[q] = 0.0; w[q] = x[i0 + q]; }
  for (int j = 0; j < nj; ++j) {
    const double a = x[j];
#pragma unroll
    for (int q = 0; q < RB_AL; ++q) acc[q] = fma(a, w[q], acc[q]);
#pragma unroll
    for (int q = 0; q < RB_AL - 1; ++q) w[q] = w[q + 1];
    w[RB_AL - 1] = x[j + i0 + RB_AL];
  }
#pragma unroll
  for (int q = 0; q < RB_AL; ++q) if (i0 + q < n) acf[i0 + q] = acc[q] / (double)(n - i0 - q);
}

__device__ __forceinline__ double rb_sum(double v, double* scr) { double a[1] = { v }; block_sum<1>(a, scr); return a[0]; }

__global__ void __launch_bounds__(BT_THREADS) k_rhythm_back(AfxBatchDev B, AfxParams P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scr[64];
  __shared__ int hist[256];
  __shared__ int ctl[4];
  __shared__ double res[2][2];      // [type][tempo, confidence]

  const int tid = threadIdx.x;
  const int fi = B.file_order[B.file0 + blockIdx.x];
  const AfxFile f = B.files[fi];
  if (f.status != 0) return;
  const int n = B.state[fi].Fr;
  if (n <= 0) return;
  const int cap = B.max_fr;
  double* s = reinterpret_cast<double*>(smem_raw);                        // [cap] sharpened onsets
  unsigned char* lm = reinterpret_cast<unsigned char*>(s + cap + 32);     // [cap] local-maximum flags (s is zero padded)
  unsigned* bits = reinterpret_cast<unsigned*>(lm + ((cap + 15) & ~15));  // [(cap + 31) / 32] candidate mask
  const int nw = (n + 31) >> 5;
  double* H = B.header + (size_t)fi * AFX_N_HEADER;
  double* gs = B.scratch + (size_t)(f.rframe_off - B.rslot0) * 4;
  const int fcap = f.rframe_cap;
  const size_t TFr = (size_t)B.TFr;

  for (int ty = 0; ty < 2; ++ty) {
    const float* post = B.rpost + (size_t)ty * TFr + f.rframe_off;
    double* raw = B.fsr + (size_t)ty * TFr + f.rframe_off;
    double* sharp = gs + (size_t)ty * fcap;
    double* acf = gs + (size_t)2 * fcap;
    double* acfout = gs + (size_t)3 * fcap;
    const float thr_f = ty ? 0.8f : 0.2f;                                 // RhythmTracker.cpp:17-40
    const double thr_d = ty ? 0.8 : 0.2;
    const int mingap = ty ? 41 : 21;                                      // int(44100 * {0.12, 0.06} / 128 + .5)

    // ---- min-gap peak picker (OnsetDetector.cpp:577-590): candidates in parallel, gaps sequentially ----
    for (int w = tid; w < nw; w += BT_THREADS) {
      unsigned m = 0;
      for (int q = 0; q < 32; ++q) {
        const int t = w * 32 + q;
        if (t < n) {
          const float p = post[t], pp = (t > 0) ? post[t - 1] : 0.0f;
          if (p > thr_f && pp <= thr_f) m |= 1u << q;
          raw[t] = 0.0;
        }
      }
      bits[w] = m;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      while (t < n) {
        int w = t >> 5;
        unsigned m = bits[w] & (0xffffffffu << (t & 31));
        while (!m && ++w < nw) m = bits[w];
        if (!m) break;
        t = w * 32 + __ffs(m) - 1;
        raw[t] = (double)post[t];                                        // RhythmTracker.cpp:105-115
        t += mingap + 1;
      }
    }
    __syncthreads();

    // ---- onset count (RhythmTracker.cpp:121-134) and Canny convolution (CannyWindow.cpp:50-66) ----
    int cnt = 0;
    for (int i = tid; i < n; i += BT_THREADS) {
      cnt += (raw[i] > thr_d) ? 1 : 0;
      double sum = 0.0;
      for (int sh = -12; sh < 12; ++sh) { const int j = i + sh; if (j >= 0 && j < n) sum = __dadd_rn(sum, __dmul_rn(raw[j], P.canny[sh + 12])); }
      s[i] = sum;
    }
    cnt = block_sum_i(cnt, hist);
    __syncthreads();
    // z-score, rectified (CannyWindow.cpp:68-80)
    double a = 0.0;
    for (int i = tid; i < n; i += BT_THREADS) a += s[i];
    const double mean = (n >= 2) ? rb_sum(a, scr) / (double)n : s[0];
    a = 0.0;
    for (int i = tid; i < n; i += BT_THREADS) { const double d = s[i] - mean; a += d * d; }
    const double var = (n >= 2) ? rb_sum(a, scr) / (double)n : 0.0;
    __syncthreads();
    if (var > 0.0) {
      const double sd = sqrt(var);
      for (int i = tid; i < n; i += BT_THREADS) { const double v = (s[i] - mean) / sd; s[i] = v > 0.0 ? v : 0.0; }
    }
    __syncthreads();
    for (int i = tid; i < n; i += BT_THREADS) sharp[i] = s[i];

    // ---- peaks of the sharpened function (RhythmTracker.cpp:623-659): strength, frequency ----
    a = 0.0; int np = 0; double tot = 0.0;
    for (int i = tid; i < n; i += BT_THREADS) {
      const double v = s[i];
      tot += v;
      bool is_max = true;
      const int lo = max(0, i - 24), hi = min(n - 1, i + 24);
      for (int j = lo; j <= hi; ++j) if (s[j] > v) { is_max = false; break; }
      lm[i] = is_max ? 1 : 0;
      if (v > 0.1 && is_max) { a += v; ++np; }
    }
    const double psum = rb_sum(a, scr);
    np = block_sum_i(np, hist);
    const double total_mean = (n >= 2) ? rb_sum(tot, scr) / (double)n : s[0];

    // ---- contrast (RhythmTracker.cpp:392-480): 85th percentile threshold, peaks vs preceding valleys ----
    const double pthr = block_select(s, n, (int)(85.0 / 100.0 * (n - 1)), hist, ctl);
    // The reference walks the function once, tracking the running minimum ("valley") since the last accepted peak:
    //     if (v < vval) { vpos = i; vval = v; }   if (v >= pthr && local max) { ps += v; vs += s[vpos]; ++pc; vval = v; }
    // One warp walks it 32 samples at a time instead (a single thread spent half of this kernel here): the samples after a
    // peak form a segment whose level is that peak's value; a segmented min-scan gives every lane the first position of
    // the minimum of its segment so far, a valley is "found" where that minimum is below the level, and a peak whose segment
    // found none inherits the position from the last segment that did.
    if (tid < 32) {
      const int lane = tid;
      double ps = 0, vs = 0; int pc = 0;
      double vval = pthr; int vpos = 0;                           // carried between the 32-sample steps (uniform)
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const bool valid = i < n;
        const double v = valid ? s[i] : __longlong_as_double(0x7ff0000000000000ll);
        const bool pk = valid && !(v < pthr) && lm[i];
        const unsigned pkm = __ballot_sync(0xffffffffu, pk);
        const unsigned heads = (pkm << 1) | 1u;                    // a segment starts at lane 0 and after every peak
        const int h = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));      // head of this lane's segment
        const double vprev = __shfl_up_sync(0xffffffffu, v, 1);
        const double lvl_h = (lane == 0) ? vval : vprev;          // level of a segment that starts at this lane
        const double lvl = __shfl_sync(0xffffffffu, lvl_h, h);
        double mv = v; int mp = i; bool fl = (heads >> lane) & 1u;   // inclusive segmented (min, first position) scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double ov = __shfl_up_sync(0xffffffffu, mv, o);
          const int op = __shfl_up_sync(0xffffffffu, mp, o);
          const bool of = __shfl_up_sync(0xffffffffu, fl, o);
          if (lane >= o && !fl) { if (ov <= mv) { mv = ov; mp = op; } fl = of; }
        }
        const bool found = mv < lvl;
        const unsigned fm = __ballot_sync(0xffffffffu, pk && found);         // peaks whose segment found a valley of its own
        // valley position in force at this lane's peak: its own segment's, else the last found one before it, else the carry
        const unsigned below = fm & (0xffffffffu >> (31 - lane));
        const int src = below ? 31 - __clz(below) : 0;
        const int psrc = __shfl_sync(0xffffffffu, mp, src);
        const int myv = below ? psrc : vpos;
        if (pk) { ps += v; vs += s[myv]; }
        pc += __popc(pkm);
        // carry: after a peak in lane 31 the level is its value and the valley in force stays; else the open segment's state
        const bool pk31 = (pkm >> 31) & 1u;
        const double v31 = __shfl_sync(0xffffffffu, v, 31), lvl31 = __shfl_sync(0xffffffffu, lvl, 31), mv31 = __shfl_sync(0xffffffffu, mv, 31);
        const int mp31 = __shfl_sync(0xffffffffu, mp, 31), myv31 = __shfl_sync(0xffffffffu, myv, 31);
        const bool found31 = __shfl_sync(0xffffffffu, found, 31);
        if (pk31) { vval = v31; vpos = myv31; }
        else {
          const int lastf = fm ? __shfl_sync(0xffffffffu, mp, 31 - __clz(fm)) : vpos;     // (fm has no bit 31 here)
          vval = found31 ? mv31 : lvl31;
          vpos = found31 ? mp31 : lastf;
        }
      }
      ps = warp_sum(ps); vs = warp_sum(vs);
      if (lane == 0) {
      const double pmean = pc ? ps / pc : 0.0, vmean = (pc ? vs / pc : 0.0) + 0.0001;
      double* Hh = H + H_RC_COUNT + 6 * ty;
      Hh[0] = (double)cnt;
      Hh[1] = (pmean != 0.0) ? -1.0 * pow(pmean / vmean, 1.0 / log(total_mean + 0.0001)) : 0.0;
      Hh[2] = (double)np / (double)n * (double)AFX_RHOP / (double)AFX_RFFT;        // SA.cpp:1012-1016
      double st = np ? (psum / np) / 4.0 : 0.0;                                      // SA.cpp:1018-1027
      Hh[3] = st < 0.0 ? 0.0 : (st > 1.0 ? 1.0 : st);
      Hh[4] = 0.0; Hh[5] = 0.0;
      res[ty][0] = 0.0; res[ty][1] = 0.0;
      }
    }
    __syncthreads();

    // ---- tempo: one fresh aubio beat tracker pass over the whole vector (RhythmTracker.cpp:155-230) ----
    if (cnt >= 4) {
      const int winlen = n, laglen = winlen / 4;
      // autocorrelation (mathutils.c:652-666): register-tiled groups of 9 lags, groups g and G-1-g paired for balance
      for (int i = n + tid; i < n + 32; i += BT_THREADS) s[i] = 0.0;
      __syncthreads();
      {
        const int G = (winlen + RB_AL - 1) / RB_AL;
        for (int g = tid; g < (G + 1) / 2; g += BT_THREADS) {
          rb_acf_group(s, winlen, g, acf);
          const int g2 = G - 1 - g;
          if (g2 != g) rb_acf_group(s, winlen, g2, acf);
        }
      }
      __syncthreads();
      // comb filterbank (beattracking.c:167-177) + Rayleigh weighting (:104-107, 180)
      const double rp_d = 60. * (double)P.sr / 120. / (double)AFX_RHOP;
      double lsum = 0.0, lmax = 0.0;
      for (int i = tid; i < laglen; i += BT_THREADS) {
        double v = 0.0;
        if (i >= 1 && i < laglen - 1)
          for (int aa = 1; aa <= 4; ++aa) for (int b = 1; b < 2 * aa; ++b) v += acf[i * aa + b - 1] * 1. / (2. * aa - 1.);
        v *= ((double)(i + 1.) / (rp_d * rp_d)) * exp((-((double)(i + 1.) * (double)(i + 1.)) / (2. * rp_d * rp_d)));
        acfout[i] = v;
        lsum += v; lmax = fmax(lmax, v);
      }
      const double asum = rb_sum(lsum, scr);
      const double gmax = block_max(lmax, scr);           // >= 0: the reference's running maximum starts at 0
      int li = -1;
      for (int i = tid; i < laglen; i += BT_THREADS) if (acfout[i] == gmax) li = i;   // ties -> last index (mathutils.c:267-283)
      const int maxi_raw = -block_min_i(-li, hist);
      __syncthreads();
      if (tid == 0) {
        const int maxi = maxi_raw < 0 ? 0 : maxi_raw;
        double rp;
        if (maxi > 0 && maxi < laglen - 1) {
          const double s0 = acfout[maxi - 1], s1 = acfout[maxi], s2 = acfout[maxi + 1];
          rp = maxi + .5 * (s0 - s2) / (s0 - 2. * s1 + s2);
        } else rp = (double)(unsigned)rp_d;
        double bp = rp;                                     // beattracking.c:286-410, first call: gp == 0
        while (0 < bp && bp < 25) bp = bp * 2;
        double tempo = 0.0;
        if (bp != 0) tempo = 60. / (((double)AFX_RHOP * bp) / (double)P.sr);
        double conf = 0.0;                                  // beattracking.c:432-441, mathutils.c:508-517
        if (asum != 0.) {
          double qm = 0.;
          if (!(rp >= laglen || rp < 0.)) {
            const unsigned idx = (unsigned)(rp - .5) + 1;
            if ((double)idx == rp) qm = acfout[idx];
            else { const double x0 = acfout[idx - 1], x1 = acfout[idx], x2 = acfout[idx + 1]; qm = x1 - .25 * (x0 - x2) * (rp - idx); }
          }
          conf = qm / asum;
        }
        conf = conf * 16.0; conf = conf < 0.0 ? 0.0 : (conf > 1.0 ? 1.0 : conf);
        if (tempo < 20.0 || tempo > 300.0) { tempo = 0.0; conf = 0.0; }
        else { while (tempo < 80.0) tempo *= 2.0; while (tempo >= 200.0) tempo /= 2.0; }
        H[H_RC_COUNT + 6 * ty + 4] = tempo; H[H_RC_COUNT + 6 * ty + 5] = conf;
        res[ty][0] = tempo; res[ty][1] = conf;
      }
      __syncthreads();
    }
  }

  // ---- final tempo: the more confident onset type + duration heuristics (RhythmTracker.cpp:234-325) ----
  if (tid == 0) {
    const int w = (res[1][1] > res[0][1]) ? 1 : 0;
    const double tempo_in = res[w][0], conf_in = res[w][1];
    double tempo = 0.0, conf = 0.0;
    if (tempo_in != 0) {
      const double* sharp = gs + (size_t)w * fcap;
      const double* raw = B.fsr + (size_t)w * TFr + f.rframe_off;
      const double thr = w == 0 ? 0.2 : 0.8;
      const double dur_s = (double)__fdiv_rn(r_samples_to_ms(f.src_rate, f.nframes_src), 1000.0f);   // SA.cpp:1031-1040
      const double off_s = (double)__fdiv_rn(r_samples_to_ms(f.src_rate, B.state[fi].data_offset), 1000.0f);
      tempo = tempo_in; conf = conf_in;
      const double spb = 60.0 / tempo * P.sr;
      int last = n - 1;
      while (last > 0 && sharp[last] < 0.1) --last;
      const double ns = (double)(last * AFX_RHOP);
      if (ns < spb * 3) { tempo = 0.0; conf = 0.0; }
      else {
        const double nb = r_guess_beats(80, dur_s);
        if (nb >= 4 && nb <= 16) {
          const double gbpm = nb / (dur_s / 60);
          const double delay = (double)r_samples_to_ms(P.sr, AFX_RHOP / 2) / 1000.0;
          const double gc = r_onset_match_conf(P.sr, raw, n, off_s + delay, nb, gbpm, thr);
          if ((gc > 0.5) || (conf < 0.1 && gc > 0.1) || (conf < 0.5 && fabs(gbpm - tempo) < 10)) { tempo = gbpm; conf = 0.5 > gc ? 0.5 : gc; }
        }
      }
    }
    H[H_FINAL_TEMPO] = tempo; H[H_FINAL_TEMPO_CONF] = conf;
  }
}

void afx_launch_rhythm(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches)
{
  if (B.g_files <= 0 || B.g_rslots <= 0) return;
  const int cap = B.max_fr;
  const int smem_back = (cap + 32) * 8 + ((cap + 15) & ~15) + ((cap + 31) / 32) * 4 + 16;
  cudaFuncSetAttribute(k_rhythm_back, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);   // per device, see afx_pitch.cu
  if (B.rhythm_fused) {     // one CTA per file: chosen per launch group by afx_batch_compute (enough files to fill the GPU that way)
    cudaFuncSetAttribute(k_rhythm_front, cudaFuncAttributeMaxDynamicSharedMemorySize, RF_SMEM);
    k_rhythm_front<<<B.g_files, RF_THREADS, RF_SMEM, s>>>(B, P); ++*launches;
  } else {
    const int fb = (B.g_rslots + OW * OB - 1) / (OW * OB);
    // the pipeline needs a CTA (a file) per resident slot to be worth it; AFX_RHYTHM_PIPE=0 / 1 forces the choice (afx_create)
    const int pipe = B.rhythm_pipe;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    k_rhythm_polar<<<(B.g_rslots + PF * PI - 1) / (PF * PI), PF * 16, 0, s>>>(B, P); ++*launches;
    if (pipe == 1 || (pipe < 0 && B.g_files >= 2 * sms)) {
      k_rhythm_pipe<<<B.g_files, RP_THREADS, 0, s>>>(B, P); ++*launches;
    } else {
      k_rhythm_whiten<<<B.g_files, 256, 0, s>>>(B, P); ++*launches;
      k_rhythm_odf<<<fb, OW * 32, 0, s>>>(B, P); ++*launches;
      k_rhythm_power<<<(B.g_rslots + 127) / 128, 128, 0, s>>>(B, P); ++*launches;
    }
  }
  { const int nchunks = (B.g_rslots + ML - 1) / ML; k_rhythm_median<<<(2 * nchunks + 63) / 64, 64, 0, s>>>(B); ++*launches; }
  k_rhythm_back<<<B.g_files, BT_THREADS, smem_back, s>>>(B, P); ++*launches;
}

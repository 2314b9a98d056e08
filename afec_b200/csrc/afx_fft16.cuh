// Register-blocked FP64 FFTs for the main (1024 complex), pitch (2048 complex) and rhythm (256 complex)
// transforms: Stockham autosort with radix-16 butterflies held in registers, so a 2048-point transform
// makes 3 trips through shared memory instead of the 6 of a radix-4 kernel, and the FP64 pipe -- the
// scarce unit on this path -- only sees the butterfly arithmetic.
//
//   N = NT * 16 points, NT cooperating threads, each thread owns 16 points.
//   input  : v[r] = x[tid + r * NT]             (r = 0..15)   -- callers load straight from global memory
//   output : X[j] at buf[FFT_PHYS(j)]           natural order, followed by a group sync
//   buf    : N + N/16 double2 (one pad element per 16: keeps the stride-16 stores of the first pass and
//            the stride-NT loads of the later passes on distinct banks)
//   tw     : per-pass twiddle tables stored in the order the threads read them (FftTw), so a warp's twiddle
//            load is one or two contiguous lines instead of up to 32 scattered ones
//   sync   : functor synchronising the NT threads (__syncthreads, a named barrier or __syncwarp)
//
// A single buffer is enough: every pass first pulls its 16 inputs into registers, syncs, then writes.
// Forward transform only (exp(-i)); the inverse the pitch kernel needs is conj(FFT(conj(.))).
#pragma once
#include "afx_common.cuh"

#define FFT_PHYS(i) ((i) + ((i) >> 4))

__device__ __forceinline__ double2 f_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 f_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 f_mul(double2 a, double2 b)
{
  return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}

// forward radix-4 butterfly in place: (a0..a3) -> (X0..X3)
__device__ __forceinline__ void f_r4(double2& a0, double2& a1, double2& a2, double2& a3)
{
  const double2 s02 = f_add(a0, a2), d02 = f_sub(a0, a2), s13 = f_add(a1, a3), d13 = f_sub(a1, a3);
  const double2 md = make_double2(d13.y, -d13.x);          // -i (a1 - a3)
  a0 = f_add(s02, s13); a1 = f_add(d02, md); a2 = f_sub(s02, s13); a3 = f_sub(d02, md);
}

// multiply by exp(-2 pi i e / 16) for the exponents the 4 x 4 decomposition needs
#define FFT_C1 0.92387953251128673848
#define FFT_S1 0.38268343236508978178
#define FFT_H 0.70710678118654752440
__device__ __forceinline__ double2 f_w16_1(double2 a) { return make_double2(fma(a.x, FFT_C1, a.y * FFT_S1), fma(a.y, FFT_C1, -(a.x * FFT_S1))); }
__device__ __forceinline__ double2 f_w16_2(double2 a) { return make_double2((a.x + a.y) * FFT_H, (a.y - a.x) * FFT_H); }
__device__ __forceinline__ double2 f_w16_3(double2 a) { return make_double2(fma(a.x, FFT_S1, a.y * FFT_C1), fma(a.y, FFT_S1, -(a.x * FFT_C1))); }
__device__ __forceinline__ double2 f_w16_4(double2 a) { return make_double2(a.y, -a.x); }
__device__ __forceinline__ double2 f_w16_6(double2 a) { return make_double2((a.y - a.x) * FFT_H, -(a.x + a.y) * FFT_H); }

// 16-point forward DFT in registers.  On return X[m + 4 n] sits in v[4 m + n]  (see FFT_REG16).
#define FFT_REG16(q) (4 * ((q) & 3) + ((q) >> 2))
__device__ __forceinline__ void f_dft16(double2 (&v)[16])
{
  f_r4(v[0], v[4], v[8], v[12]);
  f_r4(v[1], v[5], v[9], v[13]);
  f_r4(v[2], v[6], v[10], v[14]);
  f_r4(v[3], v[7], v[11], v[15]);
  // v[c + 4 m] = a[c][m]; twiddle by W16^(c m)
  v[5] = f_w16_1(v[5]);  v[9] = f_w16_2(v[9]);   v[13] = f_w16_3(v[13]);
  v[6] = f_w16_2(v[6]);  v[10] = f_w16_4(v[10]); v[14] = f_w16_6(v[14]);
  v[7] = f_w16_3(v[7]);  v[11] = f_w16_6(v[11]);
  { // W16^9 = (-c1, +s1):  (a.x + i a.y)(-c1 + i s1) = (-a.x c1 - a.y s1) + i (a.x s1 - a.y c1)
    const double2 a = v[15];
    v[15] = make_double2(-fma(a.x, FFT_C1, a.y * FFT_S1), fma(a.x, FFT_S1, -(a.y * FFT_C1)));
  }
  f_r4(v[0], v[1], v[2], v[3]);
  f_r4(v[4], v[5], v[6], v[7]);
  f_r4(v[8], v[9], v[10], v[11]);
  f_r4(v[12], v[13], v[14], v[15]);
}

// 8-point forward DFT of v[o .. o+7] in place; on return X[m] = v[o + 2m], X[m + 4] = v[o + 2m + 1]
#define FFT_REG8(q) (2 * ((q) & 3) + ((q) >> 2))
template <int O>
__device__ __forceinline__ void f_dft8(double2 (&v)[16])
{
  f_r4(v[O + 0], v[O + 2], v[O + 4], v[O + 6]);
  f_r4(v[O + 1], v[O + 3], v[O + 5], v[O + 7]);
  // v[O + c + 2 m] = a[c][m]; twiddle the odd column by W8^m
  v[O + 3] = f_w16_2(v[O + 3]); v[O + 5] = f_w16_4(v[O + 5]); v[O + 7] = f_w16_6(v[O + 7]);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const double2 a = v[O + 2 * m], b = v[O + 2 * m + 1];
    v[O + 2 * m] = f_add(a, b); v[O + 2 * m + 1] = f_sub(a, b);
  }
}

struct FftSyncBlock { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
struct FftSyncWarp { __device__ __forceinline__ void operator()() const { __syncwarp(); } };
template <int NT>
struct FftSyncNamed {       // NT threads (a multiple of 32) sharing barrier `id` (1..15)
  int id;
  __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(NT) : "memory"); }
};

struct FftTw {
  const double2* t2;     // [15][16]     exp(-2 pi i r k / 256)
  const double2* t3;     // [R - 1][256] exp(-2 pi i r j / N), R = 4 (N = 1024) or 8 (N = 2048); unused for N = 256
};

// TW_SMEM: the twiddle tables were copied to shared memory by the caller (plain loads instead of ld.global.nc)
template <bool TW_SMEM>
__device__ __forceinline__ double2 fft_tw_load(const double2* p) { return TW_SMEM ? *p : __ldg(p); }

// T3STEP: the N = 1024 third pass may read its twiddles from the N = 2048 table (exp(-2 pi i r j / 1024) is row 2r of
// exp(-2 pi i r j / 2048)): T3STEP = 2 with tw.t3 pointing at the 2048 table
//
// T2_POWERS: the second pass loads four of its fifteen twiddles and multiplies the rest together.
// T3_POWERS (N = 1024 only): the third pass derives its twiddles w^2, w^3 from w (see there).
// HALF_OPT (N = 1024 only): with `half` set at run time the last pass keeps only outputs 0..511 and stores X[j] at
// buf[FFT_PAD8(j)] -- a layout in which a thread can then read 8 consecutive outputs without bank conflicts (the pitch kernel
// needs just the first half of its inverse transform, 16 consecutive lags per thread).
#define FFT_PAD8(i) ((i) + ((i) >> 3))
template <int N, class Sync, bool TW_SMEM = false, int T3STEP = 1, bool HALF_OPT = false, bool T3_POWERS = false, bool T2_POWERS = false>
__device__ __forceinline__ void fft16_run(double2 (&v)[16], double2* __restrict__ buf, const FftTw tw, int tid, Sync sync, bool half = false)
{
  constexpr int NT = N / 16;
  static_assert(N == 256 || N == 1024 || N == 2048, "supported sizes");
  // ---- pass 1: radix 16, p = 1 (no twiddles); thread tid writes the 16 contiguous outputs of butterfly tid
  f_dft16(v);
#pragma unroll
  for (int q = 0; q < 16; ++q) buf[tid * 17 + q] = v[FFT_REG16(q)];      // FFT_PHYS(tid * 16 + q)
  sync();
  // ---- pass 2: radix 16, p = 16
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = buf[FFT_PHYS(tid + r * NT)];
  sync();
  {
    const int k = tid & 15;
    if (T2_POWERS) {                       // w^1, w^2, w^4, w^8 from the table, the other eleven powers as products (at most three deep)
      double2 w[16];
      w[1] = fft_tw_load<TW_SMEM>(tw.t2 + 0 * 16 + k); w[2] = fft_tw_load<TW_SMEM>(tw.t2 + 1 * 16 + k);
      w[4] = fft_tw_load<TW_SMEM>(tw.t2 + 3 * 16 + k); w[8] = fft_tw_load<TW_SMEM>(tw.t2 + 7 * 16 + k);
      w[3] = f_mul(w[1], w[2]); w[5] = f_mul(w[1], w[4]); w[6] = f_mul(w[2], w[4]); w[7] = f_mul(w[3], w[4]);
#pragma unroll
      for (int r = 9; r < 16; ++r) w[r] = f_mul(w[8], w[r - 8]);
#pragma unroll
      for (int r = 1; r < 16; ++r) v[r] = f_mul(v[r], w[r]);
    } else {
#pragma unroll
      for (int r = 1; r < 16; ++r) v[r] = f_mul(v[r], fft_tw_load<TW_SMEM>(tw.t2 + (r - 1) * 16 + k));
    }
    f_dft16(v);
    const int base = (tid - k) * 16 + k;
#pragma unroll
    for (int q = 0; q < 16; ++q) buf[FFT_PHYS(base + q * 16)] = v[FFT_REG16(q)];
  }
  sync();
  if (N == 256) return;
  // ---- pass 3: radix 4 (N = 1024, four butterflies per thread) or radix 8 (N = 2048, two per thread), p = 256
  if (N == 1024) {
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int r = 0; r < 4; ++r) v[4 * b + r] = buf[FFT_PHYS(tid + NT * b + r * 256)];
    sync();
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = tid + NT * b;          // k = i (i < p = 256)
      if (T3_POWERS) {                     // w, w^2, w^3 from ONE table load (two products, ~2 ulp): the kernel is short of shared-memory bandwidth
        const double2 w1 = fft_tw_load<TW_SMEM>(tw.t3 + (T3STEP - 1) * 256 + i);
        const double2 w2 = f_mul(w1, w1), w3 = f_mul(w2, w1);
        v[4 * b + 1] = f_mul(v[4 * b + 1], w1); v[4 * b + 2] = f_mul(v[4 * b + 2], w2); v[4 * b + 3] = f_mul(v[4 * b + 3], w3);
      } else {
#pragma unroll
        for (int r = 1; r < 4; ++r) v[4 * b + r] = f_mul(v[4 * b + r], fft_tw_load<TW_SMEM>(tw.t3 + (r * T3STEP - 1) * 256 + i));
      }
      f_r4(v[4 * b], v[4 * b + 1], v[4 * b + 2], v[4 * b + 3]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (HALF_OPT && half) { if (q < 2) buf[FFT_PAD8(i + q * 256)] = v[4 * b + q]; }
        else buf[FFT_PHYS(i + q * 256)] = v[4 * b + q];
      }
    }
    sync();
  } else {
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int r = 0; r < 8; ++r) v[8 * b + r] = buf[FFT_PHYS(tid + NT * b + r * 256)];
    sync();
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = tid + NT * b;
#pragma unroll
      for (int r = 1; r < 8; ++r) v[8 * b + r] = f_mul(v[8 * b + r], fft_tw_load<TW_SMEM>(tw.t3 + (r - 1) * 256 + i));
    }
    f_dft8<0>(v);
    f_dft8<8>(v);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = tid + NT * b;
#pragma unroll
      for (int q = 0; q < 8; ++q) buf[FFT_PHYS(i + q * 256)] = v[8 * b + FFT_REG8(q)];
    }
    sync();
  }
}

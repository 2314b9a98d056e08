// Shared-memory Stockham radix-4 FFTs (FP64) used by the main (2048 real), pitch (2048 complex
// pair) and rhythm (512 real) kernels.
//
// The reference runs Ooura's cdft on a zero-imaginary 2048-point buffer
// (Source/Core/AudioTypes/Source/Fourier.cpp:219-274).  Here a real frame is packed as 1024
// complex points (even samples -> re, odd -> im), transformed with five radix-4 Stockham passes
// (auto-sorting: no bit reversal) and unpacked into the half spectrum.  Twiddles come from a
// device table exp(-2 pi i k / 2048) built on the host in double precision.
#pragma once
#include "afx_common.cuh"

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// One radix-4 Stockham pass of an N-point forward (exp(-i)) transform.
//   src, dst : N complex values in shared memory (distinct buffers)
//   p        : current sub-transform length (1, 4, 16, ...)
//   tw       : table exp(-2 pi i k / TWN), TWN a multiple of N
// Thread `tid` of a group of `nthreads` handles butterflies tid, tid + nthreads, ...
template <int N, int TWN>
__device__ __forceinline__ void stockham_r4_pass(const double2* __restrict__ src, double2* __restrict__ dst,
                                                 int p, const double2* __restrict__ tw, int tid, int nthreads)
{
  constexpr int T = N / 4;
  for (int i = tid; i < T; i += nthreads) {
    const int k = i & (p - 1);
    double2 u0 = src[i], u1 = src[i + T], u2 = src[i + 2 * T], u3 = src[i + 3 * T];
    if (p > 1) {
      const int step = (TWN / 4) / p;           // twiddle index stride: exp(-2 pi i k / (4p))
      const double2 w1 = __ldg(tw + k * step), w2 = __ldg(tw + 2 * k * step), w3 = __ldg(tw + 3 * k * step);
      u1 = cmul(u1, w1); u2 = cmul(u2, w2); u3 = cmul(u3, w3);
    }
    const double2 v0 = cadd(u0, u2), v1 = cadd(u1, u3), v2 = csub(u0, u2);
    const double2 d = csub(u1, u3);
    const double2 v3 = make_double2(d.y, -d.x);   // (u1 - u3) * (-i)
    const int j = ((i - k) << 2) + k;
    dst[j] = cadd(v0, v1);
    dst[j + p] = cadd(v2, v3);
    dst[j + 2 * p] = csub(v0, v1);
    dst[j + 3 * p] = csub(v2, v3);
  }
}

// radix-2 pass (used when log4 does not divide: 256 = 4^4 -> none needed; kept for 512 complex)
template <int N, int TWN>
__device__ __forceinline__ void stockham_r2_pass(const double2* __restrict__ src, double2* __restrict__ dst,
                                                 int p, const double2* __restrict__ tw, int tid, int nthreads)
{
  constexpr int T = N / 2;
  for (int i = tid; i < T; i += nthreads) {
    const int k = i & (p - 1);
    double2 u0 = src[i], u1 = src[i + T];
    if (p > 1) u1 = cmul(u1, __ldg(tw + k * ((TWN / 2) / p)));
    const int j = ((i - k) << 1) + k;
    dst[j] = cadd(u0, u1);
    dst[j + p] = csub(u0, u1);
  }
}

// Full forward transform of N = 4^m complex points held in `a`; result ends in the returned buffer.
// `sync` is __syncthreads() for block-wide groups or __syncwarp() for warp-sized groups.
template <int N, int TWN, bool WARP_SYNC>
__device__ __forceinline__ double2* fft_pow4(double2* a, double2* b, const double2* __restrict__ tw, int tid, int nthreads)
{
  double2* src = a; double2* dst = b;
#pragma unroll 1
  for (int p = 1; p < N; p <<= 2) {
    stockham_r4_pass<N, TWN>(src, dst, p, tw, tid, nthreads);
    if (WARP_SYNC) __syncwarp(); else __syncthreads();
    double2* t = src; src = dst; dst = t;
  }
  return src;
}

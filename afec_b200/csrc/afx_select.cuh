// Block-wide order statistics of doubles by an 8-pass most-significant-digit radix select (shared by the rhythm back end
// and the high-level derivations): the k-th smallest of s[0..n) without sorting, exact for ties.
#pragma once
#include "afx_common.cuh"

__device__ __forceinline__ unsigned long long rb_order_key(double x)
{
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double rb_key_to_double(unsigned long long k)
{
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// k-th smallest (0-based) of s[0..n) by an 8-pass MSD radix select; all threads call, result broadcast
static __device__ double block_select(const double* s, int n, int k, int* hist, int* ctl)
{
  unsigned long long prefix = 0ull, pmask = 0ull;
  const int tid = threadIdx.x, lane = tid & 31;
  for (int byte = 7; byte >= 0; --byte) {
    for (int q = tid; q < 256; q += blockDim.x) hist[q] = 0;
    __syncthreads();
    const int sh = byte * 8;
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned long long key = rb_order_key(s[i]);
      if ((key & pmask) == prefix) atomicAdd(&hist[(int)((key >> sh) & 0xff)], 1);
    }
    __syncthreads();
    if (tid < 32) {
      int c[8]; int tot = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { c[q] = hist[lane * 8 + q]; tot += c[q]; }
      int inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int pv = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += pv; }
      const int excl = inc - tot;
      if (k >= excl && k < inc) {
        int run = excl, digit = -1, newk = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { if (digit < 0 && k < run + c[q]) { digit = lane * 8 + q; newk = k - run; } run += c[q]; }
        ctl[0] = digit; ctl[1] = newk;
      }
    }
    __syncthreads();
    prefix |= ((unsigned long long)ctl[0]) << sh; pmask |= 0xffull << sh; k = ctl[1];
    __syncthreads();
  }
  return rb_key_to_double(prefix);
}


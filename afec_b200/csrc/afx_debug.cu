// Parity / known-answer hooks (tests only; not on the product path): run the register-blocked FFT core of
// afx_fft16.cuh on caller data so that it can be checked against an independent FFT, like the reference's
// own TAudioTypesTest::Fourier (Source/Core/AudioTypes/Test/TestFourier.cpp:12-84).
#include "afx_fft16.cuh"
#include "../../include/afec_b200.h"

template <int N>
__global__ void __launch_bounds__(N / 16) k_debug_fft(const double2* __restrict__ in, double2* __restrict__ out,
                                                      FftTw tw)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* buf = reinterpret_cast<double2*>(smem_raw);
  const int tid = threadIdx.x;
  const double2* x = in + (size_t)blockIdx.x * N;
  double2 v[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = x[tid + r * (N / 16)];
  if (N == 256) fft16_run<N>(v, buf, tw, tid, FftSyncWarp());
  else fft16_run<N>(v, buf, tw, tid, FftSyncBlock());
  for (int j = tid; j < N; j += N / 16) out[(size_t)blockIdx.x * N + j] = buf[FFT_PHYS(j)];
}

int afx_debug_fft_launch(int n, int batch, const double2* in, double2* out, const AfxTables& T, cudaStream_t s)
{
  const int smem = (n + n / 16) * (int)sizeof(double2);
  if (n == 256) k_debug_fft<256><<<batch, 16, smem, s>>>(in, out, FftTw{ T.fft_t2, nullptr });
  else if (n == 1024) k_debug_fft<1024><<<batch, 64, smem, s>>>(in, out, FftTw{ T.fft_t2, T.fft_t3_1024 });
  else if (n == 2048) k_debug_fft<2048><<<batch, 128, smem, s>>>(in, out, FftTw{ T.fft_t2, T.fft_t3_2048 });
  else return -1;
  return 0;
}

// K6a: auto_correlation -- TSampleAnalyser::CalcAutoCorrelation (SampleAnalyser.cpp:2312-2398) with
// TAutocorrelation::Calc (Source/Crawler/FeatureExtraction/Source/Autocorrelation.cpp:62-104).
//
// Per main frame: from the frame start (looking ahead over the REST OF THE FILE, not just the frame)
// find the first rising sample pair within 1024 samples, then the next rising pair at least 0.8 ms
// (35 samples) later -> period; correlate the 12 ms (529 samples) window that starts at the first pair
// with itself for lags 0..528 (R[i] = sum_{j < 529-i} x[j] x[j+i]), normalise by R[0] and report the
// largest coefficient at lags >= period / 2.
//
// One CTA of 288 threads per frame; lags i and width-1-i are paired so every thread sums ~width+1
// products.  The window (<= 1024 + 35 + 1024 + 1 samples are searched, 529 correlated) sits in
// shared memory as FP64.
#include "afx_common.cuh"

#define AT 288
#define AC_SPAN (1024 + 64 + 1024 + 8)

__global__ void __launch_bounds__(AT) k_autocorr(AfxBatchDev B, AfxParams P)
{
  __shared__ double xs[AC_SPAN];
  __shared__ double R[544];
  __shared__ int iscr[32];
  __shared__ double dscr[32];
  __shared__ int s_file;

  const int tid = threadIdx.x;
  const int slot = B.slot0 + blockIdx.x;
  if (tid == 0) s_file = find_file_by_frame(B.files, B.n_files, slot);
  __syncthreads();
  const int fi = s_file;
  const AfxFile f = B.files[fi];
  const AfxState st = B.state[fi];
  const int t = slot - f.frame_off;
  if (f.status != 0 || t >= st.F) return;
  const int n0 = t * P.H;
  const float* __restrict__ mono = B.mono + f.mono_off;
  int remaining = st.len - n0;                                   // SampleAnalyser.cpp:943
  const int span = min(remaining, AC_SPAN);
  for (int k = tid; k < span; k += AT) xs[k] = mdata(mono, st, n0 + k);
  __syncthreads();

  const int max_seek = P.N / 2;
  // first rising pair (SampleAnalyser.cpp:2331-2341)
  int cand = 0x7fffffff;
  { const int lim = min(remaining, max_seek) - 1;
    for (int i = tid; i < lim; i += AT) if (xs[i + 1] > xs[i]) { cand = i; break; } }
  cand = block_min_i(cand, iscr);
  int start = 0;
  if (cand != 0x7fffffff) { start = cand; remaining -= cand; }
  // next rising pair at least min_period later (SampleAnalyser.cpp:2344-2356)
  const int seek_off = min(remaining, P.ac_min_period);
  int cand2 = 0x7fffffff;
  { const int lim = min(remaining - seek_off, max_seek) - 1;
    for (int i = tid; i < lim; i += AT) if (xs[start + seek_off + i + 1] > xs[start + seek_off + i]) { cand2 = i; break; } }
  cand2 = block_min_i(cand2, iscr);
  const int period = (cand2 != 0x7fffffff) ? seek_off + cand2 : seek_off;
  double* out = B.fs + (size_t)FS_AUTOCORR * B.TF + slot;
  if (!remaining || period >= remaining) { if (tid == 0) *out = 0.0; return; }   // :2361-2365

  const int width = min(remaining, P.ac_width);
  const double* x = xs + start;
  // lags tid and width-1-tid
  for (int lag = tid; lag < (width + 1) / 2; lag += AT) {
    const int lag2 = width - 1 - lag;
    double a = 0.0, b = 0.0;
    const int na = width - lag, nb = width - lag2;
    for (int j = 0; j < na; ++j) a = fma(x[j], x[j + lag], a);
    if (lag2 != lag) for (int j = 0; j < nb; ++j) b = fma(x[j], x[j + lag2], b);
    R[lag] = a;
    if (lag2 != lag) R[lag2] = b;
  }
  __syncthreads();
  const double r0 = R[0];
  double best = 0.0;
  for (int i = period / 2 + tid; i < width; i += AT) {
    const double v = (r0 != 0) ? R[i] / r0 : R[i];
    best = fmax(best, v);
  }
  best = block_max(best, dscr);
  if (tid == 0) *out = best;
}

void afx_launch_autocorr(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  k_autocorr<<<B.g_slots, AT, 0, s>>>(B, P); ++*launches;
}

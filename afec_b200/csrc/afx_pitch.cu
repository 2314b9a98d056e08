// K6b: fundamental frequency -- aubio "yinfast" as the reference drives it
// (SampleAnalyser.cpp:876-917; aubio pitch.c:399-407, 450-462; pitchyinfast.c:81-176;
//  mathutils.c:250-258, 494-506, 606-615).
//
// Per main frame (2048 samples, W = 1024):
//   sq[tau]  = sum_{j<W} x[j+tau]^2 + sum_{j<W} x[j]^2            (prefix sums of squares)
//   r[tau]   = sum_{m<W} x[m] x[m+tau]                            (cross-correlation through FFTs)
//   yin[tau] = sq[tau] - r[tau]         QUIRK: aubio's Ooura back end scales the inverse rdft by 1/n
//                                       instead of 2/n (fft.c:462-476), so the reference subtracts r, not 2r
//   cumulative-mean normalisation, first tau' in [2, W-4] with yin'[tau'] < 0.75 and
//   yin'[tau'] < yin'[tau'+1] (else the LAST global minimum), parabolic refinement, f0 = sr / period,
//   f0 = 0 when the 2048-sample level is below -48 dB; confidence = clip((1 - yin'[(uint)period]) / 0.25).
//   failsafe_f0 = f0 if f0 > 0 and confidence > 0.2, else sr/N * centroid(mag[0..1023]) for audible hops.
//
// One CTA of 256 threads per frame.  The zero-padded first half a and the full frame b are transformed
// together as z = a + i b (one 2048-point complex FFT), split into A and B, and conj(A) B goes back
// through the same forward kernel (r = Re FFT(conj(P)) / N).  64 KB of ping-pong FFT buffers + 16 KB of
// prefix sums in dynamic shared memory.
#include "afx_fft.cuh"

#define YT 256
#define YN 2048
#define YW 1024

// exclusive prefix sum across the block of one double per thread (256 threads); returns prefix, total in *tot
__device__ __forceinline__ double block_scan_excl(double v, double* scratch, double* tot)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double p = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += p; }
  __syncthreads();
  if (lane == 31) scratch[wid] = inc;
  __syncthreads();
  double base = 0.0, total = 0.0;
  for (int w = 0; w < (YT >> 5); ++w) { const double s = scratch[w]; if (w < wid) base += s; total += s; }
  if (tot) *tot = total;
  return base + inc - v;
}

__device__ __forceinline__ double2* fft2048(double2* a, double2* b, const double2* __restrict__ tw, int tid)
{
  stockham_r2_pass<YN, YN>(a, b, 1, tw, tid, YT);
  __syncthreads();
  double2* src = b; double2* dst = a;
#pragma unroll 1
  for (int p = 2; p < YN; p <<= 2) {
    stockham_r4_pass<YN, YN>(src, dst, p, tw, tid, YT);
    __syncthreads();
    double2* t = src; src = dst; dst = t;
  }
  return src;
}

__global__ void __launch_bounds__(YT) k_pitch(AfxBatchDev B, AfxParams P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* bufA = reinterpret_cast<double2*>(smem_raw);
  double2* bufB = bufA + YN;
  double* S = reinterpret_cast<double*>(bufB + YN);      // [2049] prefix sums of squares
  __shared__ double scratch[32];
  __shared__ int iscr[32];
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

  // ---- load 8 consecutive samples per thread, prefix sums of squares, pack z = a + i b -----------
  double x[8]; double loc = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) { x[q] = mdata(mono, st, n0 + 8 * tid + q); loc += x[q] * x[q]; }
  double total;
  double pre = block_scan_excl(loc, scratch, &total);
  if (tid == 0) S[0] = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    pre += x[q] * x[q];
    S[8 * tid + q + 1] = pre;
    const int m = 8 * tid + q;
    bufA[m] = make_double2(m < YW ? x[q] : 0.0, x[q]);
  }
  __syncthreads();

  // ---- forward FFT, split, conj(conj(A) B), forward FFT again --------------------------------------
  double2* Z = fft2048(bufA, bufB, P.t.tw2048, tid);
  double2* O = (Z == bufA) ? bufB : bufA;
  for (int k = tid; k < YN; k += YT) {
    const double2 zk = Z[k], zn = cconj(Z[(YN - k) & (YN - 1)]);
    const double2 A = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y + zn.y));
    const double2 D = make_double2(0.5 * (zk.x - zn.x), 0.5 * (zk.y - zn.y));
    const double2 Bc = make_double2(D.y, -D.x);                  // D / i
    const double2 Pk = cmul(cconj(A), Bc);
    O[k] = cconj(Pk);
  }
  __syncthreads();
  double2* Rr = fft2048(O, Z, P.t.tw2048, tid);
  double* yin = reinterpret_cast<double*>(Rr == bufA ? bufB : bufA);   // the free buffer

  // ---- difference function, cumulative-mean normalisation -----------------------------------------
  double y[4]; double ysum = 0.0;
  const double sW = S[YW];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int tau = 4 * tid + q;
    const double sq = (S[tau + YW] - S[tau]) + sW;
    y[q] = sq - Rr[tau].x * (1.0 / YN);
    if (tau >= 1) ysum += y[q];
  }
  __syncthreads();        // everyone done reading Rr before yin (aliases the other buffer: safe) -- keeps phases tidy
  double run = block_scan_excl(ysum, scratch, nullptr);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int tau = 4 * tid + q;
    double v;
    if (tau == 0) v = 1.0;
    else { run += y[q]; v = (run != 0.0) ? y[q] * ((double)tau / run) : 1.0; }
    y[q] = v;
    yin[tau] = v;
  }
  __syncthreads();

  // ---- first dip below the tolerance, else the last global minimum ---------------------------------
  int cand = 0x7fffffff;
#pragma unroll
  for (int q = 3; q >= 0; --q) {
    const int p = 4 * tid + q;
    if (p >= 2 && p <= YW - 4 && y[q] < 0.75 && y[q] < yin[p + 1]) cand = p;
  }
  cand = block_min_i(cand, iscr);
  int pos;
  if (cand != 0x7fffffff) pos = cand;
  else {
    // argmin with ties -> last index (mathutils.c:250-258)
    double mv = y[0]; int mi = 4 * tid;
#pragma unroll
    for (int q = 1; q < 4; ++q) if (!(mv < y[q])) { mv = y[q]; mi = 4 * tid + q; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, mv, o); const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (ov < mv || (ov == mv && oi > mi)) { mv = ov; mi = oi; }
    }
    __syncthreads();
    if ((tid & 31) == 0) { scratch[tid >> 5] = mv; iscr[tid >> 5] = mi; }
    __syncthreads();
    mv = scratch[0]; mi = iscr[0];
    for (int w = 1; w < (YT >> 5); ++w) { const double ov = scratch[w]; const int oi = iscr[w]; if (ov < mv || (ov == mv && oi > mi)) { mv = ov; mi = oi; } }
    pos = mi;
  }
  if (tid == 0) {
    double period;
    if (pos == 0 || pos == YW - 1) period = (double)pos;          // mathutils.c:494-506
    else { const double s0 = yin[pos - 1], s1 = yin[pos], s2 = yin[pos + 1]; period = pos + .5 * (s0 - s2) / (s0 - 2. * s1 + s2); }
    unsigned peak_pos = 0;
    if (period == period && period >= 0.0 && period < (double)YW) peak_pos = (unsigned)period;
    double pitch = (period > 0.0) ? (double)P.sr / (period + 0.) : 0.0;                  // pitch.c:450-462
    const bool silent_frame = (10.0 * log10(S[YN] / (double)YN) < -48.0);                // pitch.c:399-407
    if (silent_frame) pitch = 0.0;
    double conf = (1.0 - yin[peak_pos]) / 0.25;                                          // SA.cpp:887-889
    conf = conf < 0.0 ? 0.0 : (conf > 1.0 ? 1.0 : conf);
    double fsafe = 0.0;                                                                  // SA.cpp:897-916
    if (pitch > 0.0 && conf > 0.2) fsafe = pitch;
    else {
      const bool silent_hop = (10.0 * log10(S[P.H] / (double)P.H) < -48.0);
      if (!silent_hop) { const double c = B.cent_full[slot]; fsafe = (double)P.sr / (double)P.N * (c > 0.0 ? c : 0.0); }
    }
    const size_t TF = (size_t)B.TF;
    B.fs[(size_t)FS_F0 * TF + slot] = pitch;
    B.fs[(size_t)FS_F0_CONF * TF + slot] = conf;
    B.fs[(size_t)FS_F0_FAILSAFE * TF + slot] = fsafe;
  }
}

void afx_launch_pitch(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  static bool attr_set = false;
  const int smem = 2 * YN * (int)sizeof(double2) + (YN + 8) * (int)sizeof(double);
  if (!attr_set) { cudaFuncSetAttribute(k_pitch, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr_set = true; }
  k_pitch<<<B.g_slots, YT, smem, s>>>(B, P); ++*launches;
}

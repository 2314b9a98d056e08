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
// One CTA of 128 threads per frame.  The zero-padded first half a and the full frame b are transformed
// together as z = a + i b by ONE 2048-point complex FFT (register-blocked radix 16 x 16 x 8, afx_fft16.cuh),
// split into A and B, and conj(conj(A) B) goes through the same forward transform (r = Re FFT(conj(P)) / N).
// Shared memory: one padded 2048-point FFT buffer (34 KB), prefix sums of squares (17 KB), yin' (9 KB).
#include "afx_fft16.cuh"

#define YT 128
#define YN 2048
#define YW 1024
#define PAD16(i) ((i) + ((i) >> 4))
#define PAD8(i) ((i) + ((i) >> 3))

// exclusive prefix sum across the block of one double per thread (YT threads); returns prefix, total in *tot
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
#pragma unroll
  for (int w = 0; w < (YT >> 5); ++w) { const double s = scratch[w]; if (w < wid) base += s; total += s; }
  if (tot) *tot = total;
  return base + inc - v;
}

__global__ void __launch_bounds__(YT, 3) k_pitch(AfxBatchDev B, AfxParams P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* buf = reinterpret_cast<double2*>(smem_raw);                 // [2048 + 128]
  double* S = reinterpret_cast<double*>(buf + YN + YN / 16);           // [PAD16(2048) + 1] prefix sums of squares
  double* yin = S + (YN + YN / 16 + 8);                                // [PAD8(1024)]
  __shared__ double scratch[32];
  __shared__ int iscr[32];

  const int tid = threadIdx.x;
  const int slot = B.slot0 + blockIdx.x;
  const int fi = B.slot_file[slot];
  const AfxFile f = B.files[fi];
  const AfxState st = B.state[fi];
  const int t = slot - f.frame_off;
  if (f.status != 0 || t >= st.F) return;
  const int n0 = t * P.H;
  const float* __restrict__ mono = B.mono + f.mono_off;
  FftSyncBlock sync;

  // ---- one coalesced pass over the frame: z = a + i b in the FFT's strided order, squares to shared memory ----
  double2 v[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int m = tid + YT * r;
    const double xv = mdata(mono, st, n0 + m);
    v[r] = make_double2(m < YW ? xv : 0.0, xv);
    S[PAD16(m)] = xv * xv;
  }
  __syncthreads();
  // ---- prefix sums of squares: 16 consecutive samples per thread (padded -> conflict free) -----------------
  {
    double q2[16]; double loc = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) { q2[q] = S[PAD16(16 * tid + q)]; loc += q2[q]; }
    double pre = block_scan_excl(loc, scratch, nullptr);      // its first barrier: every square has been read
    if (tid == 0) S[0] = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) { pre += q2[q]; S[PAD16(16 * tid + q + 1)] = pre; }
  }
  const FftTw ftw = { P.t.fft_t2, P.t.fft_t3_2048 };
  fft16_run<YN>(v, buf, ftw, tid, sync);
  // split into A (transform of a) and B (of b), O = conj(conj(A) B); every thread builds its own 16 inputs
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int k = tid + YT * r;
    const double2 zk = buf[FFT_PHYS(k)], zc = buf[FFT_PHYS((YN - k) & (YN - 1))];
    const double2 zn = make_double2(zc.x, -zc.y);
    const double2 A = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y + zn.y));
    const double2 D = make_double2(0.5 * (zk.x - zn.x), 0.5 * (zk.y - zn.y));
    const double2 Bc = make_double2(D.y, -D.x);                  // D / i
    const double2 Pk = f_mul(make_double2(A.x, -A.y), Bc);
    v[r] = make_double2(Pk.x, -Pk.y);
  }
  __syncthreads();                                               // all reads of buf done before it is rewritten
  fft16_run<YN>(v, buf, ftw, tid, sync);

  // ---- difference function (elementwise, tau = tid + 128 c) ------------------------------------------
  const double sW = S[PAD16(YW)];
#pragma unroll
  for (int c = 0; c < YW / YT; ++c) {
    const int tau = tid + YT * c;
    const double sq = (S[PAD16(tau + YW)] - S[PAD16(tau)]) + sW;
    yin[PAD8(tau)] = sq - buf[FFT_PHYS(tau)].x * (1.0 / YN);
  }
  __syncthreads();
  // ---- cumulative-mean normalisation: 8 consecutive tau per thread ---------------------------------------
  double y[8]; double ysum = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) { y[q] = yin[PAD8(8 * tid + q)]; if (8 * tid + q >= 1) ysum += y[q]; }
  double run = block_scan_excl(ysum, scratch, nullptr);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int tau = 8 * tid + q;
    double vv;
    if (tau == 0) vv = 1.0;
    else { run += y[q]; vv = (run != 0.0) ? y[q] * ((double)tau / run) : 1.0; }
    y[q] = vv;
    yin[PAD8(tau)] = vv;
  }
  __syncthreads();

  // ---- first dip below the tolerance, else the last global minimum ---------------------------------
  int cand = 0x7fffffff;
#pragma unroll
  for (int q = 7; q >= 0; --q) {
    const int p = 8 * tid + q;
    const double nxt = (q < 7) ? y[q + 1] : yin[PAD8(min(p + 1, YW - 1))];
    if (p >= 2 && p <= YW - 4 && y[q] < 0.75 && y[q] < nxt) cand = p;
  }
  cand = block_min_i(cand, iscr);
  int pos;
  if (cand != 0x7fffffff) pos = cand;
  else {
    // argmin with ties -> last index (mathutils.c:250-258)
    double mv = y[0]; int mi = 8 * tid;
#pragma unroll
    for (int q = 1; q < 8; ++q) if (!(mv < y[q])) { mv = y[q]; mi = 8 * tid + q; }
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
    else { const double s0 = yin[PAD8(pos - 1)], s1 = yin[PAD8(pos)], s2 = yin[PAD8(pos + 1)]; period = pos + .5 * (s0 - s2) / (s0 - 2. * s1 + s2); }
    unsigned peak_pos = 0;
    if (period == period && period >= 0.0 && period < (double)YW) peak_pos = (unsigned)period;
    double pitch = (period > 0.0) ? (double)P.sr / (period + 0.) : 0.0;                  // pitch.c:450-462
    const bool silent_frame = (S[PAD16(YN)] / (double)YN) < AFX_SILENCE_LEVEL;         // pitch.c:399-407
    if (silent_frame) pitch = 0.0;
    double conf = (1.0 - yin[PAD8(peak_pos)]) / 0.25;                                    // SA.cpp:887-889
    conf = conf < 0.0 ? 0.0 : (conf > 1.0 ? 1.0 : conf);
    double fsafe = 0.0;                                                                  // SA.cpp:897-916
    if (pitch > 0.0 && conf > 0.2) fsafe = pitch;
    else {
      const bool silent_hop = (S[PAD16(P.H)] / (double)P.H) < AFX_SILENCE_LEVEL;
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
  const int smem = (YN + YN / 16) * (int)sizeof(double2) + (YN + YN / 16 + 8) * (int)sizeof(double) + (YW + YW / 8 + 8) * (int)sizeof(double);
  // per launch: function attributes are per device, and one process may drive several devices
  cudaFuncSetAttribute(k_pitch, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_pitch, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  k_pitch<<<B.g_slots, YT, smem, s>>>(B, P); ++*launches;
}

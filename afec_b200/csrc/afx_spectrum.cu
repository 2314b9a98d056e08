// K2 + K3: per main frame -- window, 2048-point real FFT, magnitude spectrum, the spectral scalars
// on the analysis window (bins 1..738) and the time-domain amplitude features of the hop slice.
//
// Reference: SampleAnalyser.cpp:814-847 (window, FFT / N, magnitude), :865-873 + :1760-1804
// (silence, amplitude peak / rms / envelope), :1808-1933 (spectral rms, centroid, spread, skewness,
// kurtosis, rolloff, flatness, flux) with TStatistics (Statistics.cpp:459-638) and LibXtract
// (scalar.c:472-493, 624-636).
//
// 64 threads per frame, 4 frames per CTA; a frame group synchronises on its own named barrier.  The real
// frame is packed as 1024 complex points (even samples -> re, odd -> im) and goes through the register-blocked
// radix 16 x 16 x 4 transform of afx_fft16.cuh; every thread then owns bins tid, tid + 64, ... for the
// magnitude and the order-free sums, and 12 consecutive analysis bins for the rolloff prefix sums.
#include "afx_fft16.cuh"
#include "../../include/afec_b200.h"

#define SG 64               // threads per frame
#define SF 4                // frames per CTA

// sum of K doubles over the 64 threads of a frame group; xch = K * 2 doubles of the group's shared scratch.
// Two barriers: the scratch is free for reuse on return.
template <int K, class Sync>
__device__ __forceinline__ void group_sum(double (&v)[K], double* xch, int gt, Sync sync)
{
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  if ((gt & 31) == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) xch[k * 2 + (gt >> 5)] = v[k];
  }
  sync();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = xch[k * 2] + xch[k * 2 + 1];
  sync();
}
template <class Sync>
__device__ __forceinline__ double group_max(double v, double* xch, int gt, Sync sync)
{
  v = warp_max(v);
  if ((gt & 31) == 0) xch[gt >> 5] = v;
  sync();
  v = fmax(xch[0], xch[1]);
  sync();
  return v;
}

__global__ void __launch_bounds__(SG * SF, 3) k_spectrum(AfxBatchDev B, AfxParams P, unsigned features)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int g = threadIdx.x / SG, gt = threadIdx.x % SG, lane = gt & 31, gw = gt >> 5;
  double2* buf = reinterpret_cast<double2*>(smem_raw) + g * (AFX_NBIN + AFX_NBIN / 16);   // FFT buffer, later mag[1024]
  double* xch = reinterpret_cast<double*>(reinterpret_cast<double2*>(smem_raw) + SF * (AFX_NBIN + AFX_NBIN / 16)) + g * 16;

  const int rel = blockIdx.x * SF + g;
  if (rel >= B.g_slots) return;                      // group-uniform; only the group's named barrier is used below
  const int slot = B.slot0 + rel;
  const int fi = B.slot_file[slot];
  const AfxFile f = B.files[fi];
  const AfxState st = B.state[fi];
  const int t = slot - f.frame_off;
  if (f.status != 0 || t >= st.F) return;
  const int n0 = t * P.H;
  const float* __restrict__ mono = B.mono + f.mono_off;
  const double2* __restrict__ win2 = reinterpret_cast<const double2*>(P.t.window);
  const int TF = B.TF;
  FftSyncNamed<SG> sync{ 1 + g };

  // ---- amplitude features of the hop slice (SA.cpp:865-873): H / 64 consecutive samples per thread --------
  if (features & AFX_FEAT_AMPLITUDE) {
    const int per = P.H / SG;                         // 4, 8, 16 or 32 (hop is a multiple of 256)
    const double c = P.env_coef;
    // one-pole envelope (Envelopes.inl:14-18) as a scan of affine maps s -> A s + Bv
    double A = 1.0, Bv = 0.0, e_hop = 0.0, pk_hop = 0.0;
    for (int q = 0; q < per; ++q) {
      const double xv = mdata(mono, st, n0 + gt * per + q), a = fabs(xv);
      e_hop += xv * xv; pk_hop = fmax(pk_hop, a);
      Bv = a + c * (Bv - a);
      A *= c;
    }
    double sA = A, sB = Bv;                           // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double pA = __shfl_up_sync(0xffffffffu, sA, o), pB = __shfl_up_sync(0xffffffffu, sB, o);
      if (lane >= o) { sB = sA * pB + sB; sA = sA * pA; }
    }
    if (lane == 31 && gw == 0) { xch[4] = sA; xch[5] = sB; }
    double ev[1] = { e_hop };
    group_sum<1>(ev, xch, gt, sync);                  // also publishes xch[4..5] (first barrier inside)
    double s_in = (gw == 1) ? xch[5] : 0.0;           // state entering the second warp = first warp's map applied to 0
    const double pA = __shfl_up_sync(0xffffffffu, sA, 1), pB = __shfl_up_sync(0xffffffffu, sB, 1);
    if (lane > 0) s_in = pA * s_in + pB;
    double env = s_in, emax = 0.0;
    for (int q = 0; q < per; ++q) { const double a = fabs(mdata(mono, st, n0 + gt * per + q)); env = a + c * (env - a); emax = fmax(emax, env); }
    const double pk = group_max(pk_hop, xch, gt, sync);
    emax = group_max(emax, xch, gt, sync);
    if (gt == 0) {
      const double level = ev[0] / (double)P.H;
      B.fs[(size_t)FS_AMP_SILENCE * TF + slot] = (level < AFX_SILENCE_LEVEL) ? 1.0 : 0.0;   // mathutils.c:606-615
      B.fs[(size_t)FS_AMP_PEAK * TF + slot] = pk;
      const double r = sqrt(level);
      B.fs[(size_t)FS_AMP_RMS * TF + slot] = (r != r) ? 0.0 : r;
      B.fs[(size_t)FS_AMP_ENV * TF + slot] = emax;
    }
  }

  // ---- load, window, pack (even -> re, odd -> im) in the FFT's strided order; transform ---------------------
  double2 v[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int m = gt + SG * r;
    const double2 w = __ldg(win2 + m);
    v[r] = make_double2(mdata(mono, st, n0 + 2 * m) * w.x, mdata(mono, st, n0 + 2 * m + 1) * w.y);
  }
  fft16_run<AFX_NBIN>(v, buf, FftTw{ P.t.fft_t2, P.t.fft_t3_1024 }, gt, sync);

  // ---- real unpack + magnitude / N for bins gt + 64 c (Fourier.cpp:266-271, AudioMath.cpp:497-504) ---------
  double m16[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int k = gt + SG * c;
    const double2 zk = buf[FFT_PHYS(k)], zc = buf[FFT_PHYS((AFX_NBIN - k) & (AFX_NBIN - 1))];
    const double2 zm = make_double2(zc.x, -zc.y);
    const double2 E = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y + zm.y));
    const double2 D = make_double2(0.5 * (zk.x - zm.x), 0.5 * (zk.y - zm.y));
    const double2 O = make_double2(D.y, -D.x);                 // D / i
    const double2 X = f_add(E, f_mul(__ldg(P.t.tw2048 + k), O));
    m16[c] = sqrt(X.x * X.x + X.y * X.y) * (1.0 / AFX_NFFT);
  }
  double* gmag = B.mag + (size_t)(slot - B.slot0) * AFX_NBIN;
  sync();                                            // everyone has read Z before buf becomes the magnitude array
  double* mag = reinterpret_cast<double*>(buf);
#pragma unroll
  for (int c = 0; c < 16; ++c) { const int k = gt + SG * c; mag[k] = m16[c]; gmag[k] = m16[c]; }

  // ---- order-free sums over the analysis window (bins first_bin .. first_bin + nbins - 1) and all bins -----
  const int nb = P.nbins, fb = P.first_bin;
  double acc[6] = { 0, 0, 0, 0, 0, 0 };        // S1, S2, SJ, log-sum, full S, full SJ
  double mant = 1.0; int ex = 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int k = gt + SG * c, j = k - fb;
    const double m = m16[c];
    acc[4] += m; acc[5] += (double)k * m;
    if (j >= 0 && j < nb) {
      acc[0] += m; acc[1] += m * m; acc[2] += (double)j * m;
      mul_frexp_pos(mant, ex, fabs(m) + 1e-20);               // Statistics.cpp:417-455
    }
  }
  acc[3] = log(mant) + (double)ex * 0.693147180559945309417;
  group_sum<6>(acc, xch, gt, sync);                  // (its first barrier also publishes mag[])
  const double S1 = acc[0];
  const double cen = (S1 == 0.0) ? 0.0 : acc[2] / S1;                          // Statistics.cpp:459-477
  double sp[1] = { 0.0 };
#pragma unroll
  for (int c = 0; c < 16; ++c) { const int j = gt + SG * c - fb; if (j >= 0 && j < nb) { const double d = (double)j - cen; sp[0] += d * d * m16[c]; } }
  group_sum<1>(sp, xch, gt, sync);
  const double spread = (S1 == 0.0) ? 0.0 : sp[0] / S1;                        // Statistics.cpp:486-506
  double sk[2] = { 0.0, 0.0 };
  const bool have_sk = fabs(spread) > (double)1e-12f;                          // Statistics.cpp:510-554
  if (have_sk) {
#pragma unroll
    for (int c = 0; c < 16; ++c) { const int j = gt + SG * c - fb; if (j >= 0 && j < nb) { const double d = (m16[c] - cen) / spread; const double d2 = d * d; sk[0] += d2 * d; sk[1] += d2 * d2; } }
  }
  group_sum<2>(sk, xch, gt, sync);

  // ---- rolloff (LibXtract scalar.c:472-493): count of prefixes below 85 % of the total; 12 bins per thread ---
  {
    const double pivot = S1 * (85.0 / 100.0);
    const int j0 = 12 * gt;
    double m12[12], loc = 0.0;
#pragma unroll
    for (int q = 0; q < 12; ++q) { m12[q] = (j0 + q < nb) ? mag[fb + j0 + q] : 0.0; loc += m12[q]; }
    double inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double pv = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += pv; }
    if (lane == 31 && gw == 0) xch[0] = inc;
    sync();
    double pre = ((gw == 1) ? xch[0] : 0.0) + inc - loc;      // exclusive prefix = sum of bins before j0
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 12; ++q) if (j0 + q < nb) { cnt += (pre < pivot) ? 1 : 0; pre += m12[q]; }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    int* ix = reinterpret_cast<int*>(xch + 2);
    if (lane == 0) ix[gw] = cnt;
    sync();
    if (gt == 0) {
      const double r = (double)(ix[0] + ix[1]) * (double)(P.sr / (P.N / 2));   // SA.cpp:1892: 44100 / 1024 = 43
      B.fs[(size_t)FS_SPEC_ROLLOFF * TF + slot] = r;
    }
  }

  if (gt == 0) {
    const double n = (double)nb;
    const double rms = sqrt(acc[1] / n);
    B.fs[(size_t)FS_SPEC_RMS * TF + slot] = (rms != rms) ? 0.0 : rms;
    B.fs[(size_t)FS_SPEC_CENTROID * TF + slot] = cen;
    B.fs[(size_t)FS_SPEC_SPREAD * TF + slot] = spread;
    B.fs[(size_t)FS_SPEC_SKEW * TF + slot] = have_sk ? sk[0] / n : 0.0;
    B.fs[(size_t)FS_SPEC_KURT * TF + slot] = have_sk ? sk[1] / n - 3.0 : 0.0;
    const double mean = S1 / n, gmean = exp(acc[3] / n);
    const double fl = flatness_db(mean, gmean);
    B.fs[(size_t)FS_SPEC_FLATNESS * TF + slot] = (fl != fl) ? 0.0 : fl;
    B.cent_full[slot] = (acc[4] == 0.0) ? 0.0 : acc[5] / acc[4];
    // degenerate in the reference (see oracle/afec_oracle.c, "harmonic spectrum"): always 0
    B.fs[(size_t)FS_SPEC_INHARM * TF + slot] = 0.0;
    B.fs[(size_t)FS_TRISTIM1 * TF + slot] = 0.0;
    B.fs[(size_t)FS_TRISTIM2 * TF + slot] = 0.0;
    B.fs[(size_t)FS_TRISTIM3 * TF + slot] = 0.0;
  }
}

// spectral flux = Pearson correlation with the previous frame's spectrum (first frame: itself),
// Statistics.cpp:578-638, SA.cpp:936-940, 1919-1933.  One warp per frame, 8 frames per CTA.
__global__ void __launch_bounds__(256) k_flux(AfxBatchDev B, AfxParams P)
{
  const int lane = threadIdx.x & 31;
  const int slot = B.slot0 + blockIdx.x * 8 + (threadIdx.x >> 5);
  if (slot >= B.slot0 + B.g_slots) return;
  const int fi = B.slot_file[slot];
  const AfxFile f = B.files[fi];
  const int t = slot - f.frame_off;
  if (f.status != 0 || t >= B.state[fi].F) return;
  const double* a = B.mag + (size_t)(slot - B.slot0) * AFX_NBIN + P.first_bin;
  const double* b = (t > 0) ? a - AFX_NBIN : a;
  double s1 = 0, s2 = 0, s11 = 0, s12 = 0, s22 = 0;
  for (int j = lane; j < P.nbins; j += 32) {
    const double x = a[j], y = b[j];
    s12 += x * y; s1 += x; s11 += x * x; s2 += y; s22 += y * y;
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2); s11 = warp_sum(s11); s12 = warp_sum(s12); s22 = warp_sum(s22);
  if (lane == 0) {
    const double n = (double)P.nbins;
    s1 = s1 / n; s2 = s2 / n;
    const double den2 = (s11 - s1 * s1 * n) * (s22 - s2 * s2 * n);
    const double num = s12 - (s1 * s2 * n);
    B.fs[(size_t)FS_SPEC_FLUX * B.TF + slot] = (fabs(den2) > (double)1e-12f) ? num / sqrt(den2) : 0.0;
  }
}

void afx_launch_spectrum(const AfxParams& P, const AfxBatchDev& B, unsigned features, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  const int smem = SF * (AFX_NBIN + AFX_NBIN / 16) * (int)sizeof(double2) + SF * 16 * (int)sizeof(double);
  cudaFuncSetAttribute(k_spectrum, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device, see afx_pitch.cu
  cudaFuncSetAttribute(k_spectrum, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  k_spectrum<<<(B.g_slots + SF - 1) / SF, SG * SF, smem, s>>>(B, P, features); ++*launches;
  k_flux<<<(B.g_slots + 7) / 8, 256, 0, s>>>(B, P); ++*launches;
}

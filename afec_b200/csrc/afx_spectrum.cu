// K2 + K3: per main frame -- window, 2048-point real FFT, magnitude spectrum, the spectral scalars
// on the analysis window (bins 1..738) and the time-domain amplitude features of the hop slice.
//
// Reference: SampleAnalyser.cpp:814-847 (window, FFT / N, magnitude), :865-873 + :1760-1804
// (silence, amplitude peak / rms / envelope), :1808-1933 (spectral rms, centroid, spread, skewness,
// kurtosis, rolloff, flatness, flux) with TStatistics (Statistics.cpp:459-638) and LibXtract
// (scalar.c:472-493, 624-636).
//
// One CTA of 256 threads per frame.  Shared memory: two 1024-point complex ping-pong buffers
// (32 KB); the magnitude spectrum aliases the buffer the last FFT pass did not write.
#include "afx_fft.cuh"
#include "../../include/afec_b200.h"

#define ST 256

__global__ void __launch_bounds__(ST) k_spectrum(AfxBatchDev B, AfxParams P, unsigned features)
{
  __shared__ double2 bufA[AFX_NBIN];
  __shared__ double2 bufB[AFX_NBIN];
  __shared__ double scratch[8 * 32];
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
  const double* __restrict__ win = P.t.window;
  const int TF = B.TF;

  // ---- load, window, pack (even -> re, odd -> im); hop-slice energy and peak on the way -------
  double e_hop = 0.0, pk_hop = 0.0;
  for (int m = tid; m < AFX_NBIN; m += ST) {
    const double x0 = mdata(mono, st, n0 + 2 * m), x1 = mdata(mono, st, n0 + 2 * m + 1);
    bufA[m] = make_double2(x0 * __ldg(win + 2 * m), x1 * __ldg(win + 2 * m + 1));
    if (2 * m < P.H) { e_hop += x0 * x0 + x1 * x1; pk_hop = fmax(pk_hop, fmax(fabs(x0), fabs(x1))); }
  }
  __syncthreads();

  // ---- amplitude features of the hop slice (SA.cpp:865-873) ------------------------------------
  if (features & AFX_FEAT_AMPLITUDE) {
    double v[1] = { e_hop };
    block_sum<1>(v, scratch);
    const double pk = block_max(pk_hop, scratch);
    // one-pole envelope (Envelopes.inl:14-18) as a scan of affine maps s -> A s + B
    const int per = P.H / ST;                 // samples per thread (hop is a multiple of 256)
    const double c = P.env_coef;
    double xs[8];
    double A = 1.0, Bv = 0.0;
    for (int q = 0; q < per; ++q) {
      xs[q] = fabs(mdata(mono, st, n0 + tid * per + q));
      Bv = xs[q] + c * (Bv - xs[q]);
      A *= c;
    }
    // inclusive scan over threads
    const int lane = tid & 31, wid = tid >> 5;
    double sA = A, sB = Bv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double pA = __shfl_up_sync(0xffffffffu, sA, o), pB = __shfl_up_sync(0xffffffffu, sB, o);
      if (lane >= o) { sB = sA * pB + sB; sA = sA * pA; }
    }
    __syncthreads();
    if (lane == 31) { scratch[wid] = sA; scratch[32 + wid] = sB; }
    __syncthreads();
    // state entering this warp = composition of all previous warps applied to 0
    double s_in = 0.0;
    for (int w = 0; w < wid; ++w) s_in = scratch[w] * s_in + scratch[32 + w];
    // state entering this thread
    const double pA = __shfl_up_sync(0xffffffffu, sA, 1), pB = __shfl_up_sync(0xffffffffu, sB, 1);
    if (lane > 0) s_in = pA * s_in + pB;
    double env = s_in, emax = 0.0;
    for (int q = 0; q < per; ++q) { env = xs[q] + c * (env - xs[q]); emax = fmax(emax, env); }
    emax = block_max(emax, scratch);
    if (tid == 0) {
      const double level = v[0] / (double)P.H;
      B.fs[(size_t)FS_AMP_SILENCE * TF + slot] = (10.0 * log10(level) < -48.0) ? 1.0 : 0.0;   // mathutils.c:606-615
      B.fs[(size_t)FS_AMP_PEAK * TF + slot] = pk;
      const double r = sqrt(level);
      B.fs[(size_t)FS_AMP_RMS * TF + slot] = (r != r) ? 0.0 : r;
      B.fs[(size_t)FS_AMP_ENV * TF + slot] = emax;
    }
  }

  // ---- FFT (1024 complex, 5 radix-4 passes) + real unpack + magnitude / N -----------------------
  double2* Z = fft_pow4<AFX_NBIN, AFX_NFFT, false>(bufA, bufB, P.t.tw2048, tid, ST);
  double* mag = reinterpret_cast<double*>(Z == bufA ? bufB : bufA);
  double* gmag = B.mag + (size_t)(slot - B.slot0) * AFX_NBIN;
  for (int k = tid; k < AFX_NBIN; k += ST) {
    const double2 zk = Z[k], zm = cconj(Z[(AFX_NBIN - k) & (AFX_NBIN - 1)]);
    const double2 E = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y + zm.y));
    const double2 D = make_double2(0.5 * (zk.x - zm.x), 0.5 * (zk.y - zm.y));
    const double2 O = make_double2(D.y, -D.x);                 // D / i
    const double2 X = cadd(E, cmul(__ldg(P.t.tw2048 + k), O));
    const double m = sqrt(X.x * X.x + X.y * X.y) * (1.0 / AFX_NFFT);   // Fourier.cpp:266-271, AudioMath.cpp:497-504
    mag[k] = m;
    gmag[k] = m;
  }
  __syncthreads();

  // ---- spectral scalars on bins first_bin .. first_bin + nbins - 1 -----------------------------------
  const int nb = P.nbins, fb = P.first_bin;
  const int j0 = 3 * tid;                       // three consecutive analysis bins per thread
  double m3[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) m3[q] = (j0 + q < nb) ? mag[fb + j0 + q] : 0.0;

  double acc[6] = { 0, 0, 0, 0, 0, 0 };        // S1, S2, SJ, log-sum, full S, full SJ
  double mant = 1.0; int ex = 0;
#pragma unroll
  for (int q = 0; q < 3; ++q) if (j0 + q < nb) {
    acc[0] += m3[q]; acc[1] += m3[q] * m3[q]; acc[2] += (double)(j0 + q) * m3[q];
    mul_frexp(mant, ex, fabs(m3[q]) + 1e-20);            // Statistics.cpp:417-455
  }
  acc[3] = log(mant) + (double)ex * 0.693147180559945309417;
  for (int k = tid; k < AFX_NBIN; k += ST) { acc[4] += mag[k]; acc[5] += (double)k * mag[k]; }
  block_sum<6>(acc, scratch);
  const double S1 = acc[0];
  const double cen = (S1 == 0.0) ? 0.0 : acc[2] / S1;                          // Statistics.cpp:459-477

  double sp[1] = { 0.0 };
#pragma unroll
  for (int q = 0; q < 3; ++q) if (j0 + q < nb) { const double d = (double)(j0 + q) - cen; sp[0] += d * d * m3[q]; }
  block_sum<1>(sp, scratch);
  const double spread = (S1 == 0.0) ? 0.0 : sp[0] / S1;                        // Statistics.cpp:486-506

  double sk[2] = { 0.0, 0.0 };
  const bool have_sk = fabs(spread) > (double)1e-12f;                          // Statistics.cpp:510-554
  if (have_sk) {
#pragma unroll
    for (int q = 0; q < 3; ++q) if (j0 + q < nb) { const double d = (m3[q] - cen) / spread; const double d2 = d * d; sk[0] += d2 * d; sk[1] += d2 * d2; }
  }
  block_sum<2>(sk, scratch);

  // rolloff (LibXtract scalar.c:472-493): count of prefixes below 85 % of the total
  const double pivot = S1 * (85.0 / 100.0);
  double loc = m3[0] + m3[1] + m3[2];
  {
    const int lane = tid & 31, wid = tid >> 5;
    double inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double pv = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += pv; }
    __syncthreads();
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    double base = 0.0;
    for (int w = 0; w < wid; ++w) base += scratch[w];
    double pre = base + inc - loc;              // exclusive prefix = sum of bins before j0
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) if (j0 + q < nb) { cnt += (pre < pivot) ? 1 : 0; pre += m3[q]; }
    int* iscr = reinterpret_cast<int*>(scratch + 64);
    cnt = block_sum_i(cnt, iscr);
    if (tid == 0) {
      const double r = (double)cnt * (double)(P.sr / (P.N / 2));              // SA.cpp:1892: 44100 / 1024 = 43
      B.fs[(size_t)FS_SPEC_ROLLOFF * TF + slot] = r;
    }
  }

  if (tid == 0) {
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
  const int fi = find_file_by_frame(B.files, B.n_files, slot);
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
  k_spectrum<<<B.g_slots, ST, 0, s>>>(B, P, features); ++*launches;
  k_flux<<<(B.g_slots + 7) / 8, 256, 0, s>>>(B, P); ++*launches;
}

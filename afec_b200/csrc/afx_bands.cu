// K4: band projections of the magnitude spectrum, one CTA (256 threads) per main frame.
//
//   * 14 sub-bands (SampleAnalyser.cpp:2067-2260): rms, flatness (dB scaled), flux (Pearson correlation
//     with the previous frame), complexity (strict local maxima above 0.25 x band max) and contrast
//     -(peakMean / valleyMean)^(1 / ln(mean)) from the sorted band; spectral_contrast = mean of the 14
//   * 28 "frequency bands" (SampleAnalyser.cpp:2007-2048): sum of squared magnitudes
//   * 14 cepstrum bands (SampleAnalyser.cpp:2052-2063; LibXtract vector.c:350-391): 14 triangular mel
//     filters -> log -> unnormalised DCT-II.  The filters only cover bins 0..359 (quirk: they are laid
//     over 512 of the 1024 bins), so the dense 14 x 1024 contraction is evaluated on its support.
//
// The in-band sort is ONE block-wide bitonic sort of all 1024 bins keyed by (band id, value): bins
// outside the 14 bands carry id 15 and sink to the end, band b ends up sorted at [start_b - first_bin, ..).
#include "afx_common.cuh"

#define BT 256
#define MEL_SUPPORT 360

__global__ void __launch_bounds__(BT) k_bands(AfxBatchDev B, AfxParams P)
{
  __shared__ double mag[AFX_NBIN];
  __shared__ double prev[AFX_NBIN];
  __shared__ double srt[AFX_NBIN];
  __shared__ unsigned char sid[AFX_NBIN];
  __shared__ double lg[16];
  __shared__ double contrast[16];
  __shared__ double bmean[16];
  __shared__ int s_file;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int slot = B.slot0 + blockIdx.x;
  if (tid == 0) s_file = find_file_by_frame(B.files, B.n_files, slot);
  __syncthreads();
  const int fi = s_file;
  const AfxFile f = B.files[fi];
  const int t = slot - f.frame_off;
  if (f.status != 0 || t >= B.state[fi].F) return;
  const size_t TF = (size_t)B.TF;
  const double* g = B.mag + (size_t)(slot - B.slot0) * AFX_NBIN;
  const double* gp = (t > 0) ? g - AFX_NBIN : g;              // SampleAnalyser.cpp:936-940
  for (int k = tid; k < AFX_NBIN; k += BT) {
    const double m = g[k];
    mag[k] = m; prev[k] = gp[k]; srt[k] = m; sid[k] = 15;
  }
  __syncthreads();
  if (tid < 14) { const int s = P.band14_start[tid], n = P.band14_n[tid]; for (int k = 0; k < n; ++k) sid[s + k] = (unsigned char)tid; }

  // ---- 28 frequency bands: warp w takes bands w, w+8, ... ----------------------------------------
  for (int b = wid; b < 28; b += 8) {
    double s = 0.0;
    for (int k = P.band28_s[b] + lane; k < P.band28_e[b]; k += 32) s += mag[k] * mag[k];
    s = warp_sum(s);
    if (lane == 0) B.fv[(size_t)FV_BANDS28 * TF + (size_t)slot * 28 + b] = s;
  }
  // ---- mel filter energies -> log ------------------------------------------------------------------
  for (int q = wid; q < 14; q += 8) {
    const double* row = P.t.mel + (size_t)q * AFX_NBIN;
    double e = 0.0;
    for (int k = lane; k < MEL_SUPPORT; k += 32) e += mag[k] * __ldg(row + k);
    e = warp_sum(e);
    if (lane == 0) lg[q] = log(e < 2e-42 ? 2e-42 : e);        // XTRACT_LOG_LIMIT
  }
  // ---- 14 sub-bands: sums, flux, complexity --------------------------------------------------------
  for (int b = wid; b < 14; b += 8) {
    const int s0 = P.band14_start[b], n = P.band14_n[b];
    double s1 = 0, s2 = 0, s11 = 0, s12 = 0, s22 = 0, mx = 0.0, mant = 1.0; int ex = 0;
    for (int k = lane; k < n; k += 32) {
      const double x = mag[s0 + k], y = prev[s0 + k];
      s12 += x * y; s1 += x; s11 += x * x; s2 += y; s22 += y * y;
      mx = fmax(mx, x);
      mul_frexp(mant, ex, fabs(x) + 1e-20);
    }
    double ls = log(mant) + (double)ex * 0.693147180559945309417;
    s1 = warp_sum(s1); s2 = warp_sum(s2); s11 = warp_sum(s11); s12 = warp_sum(s12); s22 = warp_sum(s22);
    ls = warp_sum(ls); mx = warp_max(mx);
    const double thr = mx * 0.25;
    int cplx = 0;
    if (thr > 0.0) for (int k = lane; k < n; k += 32) {
      const int q = s0 + k;
      const double x = mag[q];
      if (x > thr && q > 0 && q < AFX_NBIN - 1 && x > mag[q - 1] && x > mag[q + 1]) ++cplx;
    }
    cplx = warp_sum_i(cplx);
    if (lane == 0) {
      const double dn = (double)n;
      const double mean = (n >= 2) ? s1 / dn : s1;            // TStatistics::Mean, Statistics.cpp:249-266
      const double gmean = (n >= 2) ? exp(ls / dn) : mag[s0]; // TStatistics::GeometricMean :417-455
      bmean[b] = mean;
      const size_t o = (size_t)slot * 14 + b;
      B.fv[(size_t)FV_RMS * TF + o] = sqrt(s11 / dn);
      B.fv[(size_t)FV_FLATNESS * TF + o] = flatness_db(mean, gmean);
      const double m1 = s1 / dn, m2 = s2 / dn;
      const double den2 = (s11 - m1 * m1 * dn) * (s22 - m2 * m2 * dn);
      const double num = s12 - (m1 * m2 * dn);
      B.fv[(size_t)FV_FLUX * TF + o] = (fabs(den2) > (double)1e-12f) ? num / sqrt(den2) : 0.0;
      B.fv[(size_t)FV_COMPLEXITY * TF + o] = (double)cplx;
    }
  }
  __syncthreads();
  // ---- DCT of the log mel energies (vector.c:372-391) ----------------------------------------------
  if (tid < 14) {
    double a = 0.0;
    for (int m = 0; m < 14; ++m) a += lg[m] * __ldg(P.t.dct + tid * 14 + m);
    B.fv[(size_t)FV_CEPSTRUM * TF + (size_t)slot * 14 + tid] = a;
  }
  // ---- bitonic sort of (band id, value), ascending ---------------------------------------------------
  for (int k2 = 2; k2 <= AFX_NBIN; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int q = tid; q < AFX_NBIN / 2; q += BT) {
        const int a = ((q & ~(j - 1)) << 1) | (q & (j - 1));   // lower index of the pair
        const int c = a | j;
        const bool up = ((a & k2) == 0);
        const unsigned char ia = sid[a], ic = sid[c];
        const double va = srt[a], vc = srt[c];
        const bool gt = (ia > ic) || (ia == ic && va > vc);
        if (gt == up) { srt[a] = vc; srt[c] = va; sid[a] = ic; sid[c] = ia; }
      }
      __syncthreads();
    }
  }
  // ---- contrast (SampleAnalyser.cpp:2199-2232) --------------------------------------------------------
  for (int b = wid; b < 14; b += 8) {
    const int pos = P.band14_start[b] - P.first_bin, n = P.band14_n[b], nei = P.band14_nei[b];
    double lo = 0.0, hi = 0.0;
    for (int k = lane; k < nei && k < n; k += 32) lo += srt[pos + k];
    for (int k = lane; k < nei; k += 32) hi += srt[pos + n - 1 - k];
    lo = warp_sum(lo); hi = warp_sum(hi);
    if (lane == 0) {
      const double valley = lo / nei + 1e-30, peak = hi / nei + 1e-30;
      const double c = -1.0 * pow(peak / valley, 1.0 / log(bmean[b] + 1e-30));
      contrast[b] = c;
      B.fv[(size_t)FV_CONTRAST * TF + (size_t)slot * 14 + b] = c;
    }
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int b = 0; b < 14; ++b) s += contrast[b];
    B.fs[(size_t)FS_SPEC_CONTRAST * TF + slot] = s / 14.0;
  }
}

void afx_launch_bands(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  k_bands<<<B.g_slots, BT, 0, s>>>(B, P); ++*launches;
}

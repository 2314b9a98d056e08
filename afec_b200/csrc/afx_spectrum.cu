// K2 + K3: per main frame -- window, 2048-point real FFT, magnitude spectrum, the spectral scalars
// on the analysis window (bins 1..738) and the time-domain amplitude features of the hop slice.
//
// Reference: SampleAnalyser.cpp:814-847 (window, FFT / N, magnitude), :865-873 + :1760-1804
// (silence, amplitude peak / rms / envelope), :1808-1933 (spectral rms, centroid, spread, skewness,
// kurtosis, rolloff, flatness, flux) with TStatistics (Statistics.cpp:459-638) and LibXtract
// (scalar.c:472-493, 624-636).
//
// 64 threads per frame; a frame group synchronises on its own named barrier.  The real frame is packed as 1024
// complex points (even samples -> re, odd -> im) and goes through the register-blocked radix 16 x 16 x 4 transform
// of afx_fft16.cuh; see k_spectrum below for the persistent layout, the bin-pair unpack and the fused statistics.
#include "afx_fft16.cuh"
#include "../../include/afec_b200.h"
#include <algorithm>
#include <cstdlib>

#define SG 64               // threads per frame

// sqrt of a squared magnitude to ~2^-44 relative: MUFU.RSQ64H seed (2^-22) + one Newton step.  The full IEEE sqrt
// costs twice the FP64 instructions and its last 8 bits are far below what the FFT's own rounding leaves intact.
// s is exactly 0 or far above the subnormals the seed instruction flushes (squares of sums of float32 samples times
// the window: >= 1e-112), so a select on s > 0 replaces the range branch.
__device__ __forceinline__ double sqrt_mag(double s)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  y = (s > 0.0) ? y : 0.0;
  const double r = s * y;
  return fma(fma(-r, r, s), 0.5 * y, r);
}

// Pearson correlation from the window sums of two spectra (Statistics.cpp:604-638); explicit rounding steps so that
// both kernels that call it produce the same bits
__device__ __forceinline__ double flux_from_sums(double S1, double S2, double S1p, double S2p, double S12, double n)
{
  const double s1 = __ddiv_rn(S1, n), s2 = __ddiv_rn(S1p, n);
  const double da = __dsub_rn(S2, __dmul_rn(__dmul_rn(s1, s1), n)), db = __dsub_rn(S2p, __dmul_rn(__dmul_rn(s2, s2), n));
  const double den2 = __dmul_rn(da, db);
  const double num = __dsub_rn(S12, __dmul_rn(__dmul_rn(s1, s2), n));
  return (fabs(den2) > (double)1e-12f) ? __ddiv_rn(num, __dsqrt_rn(den2)) : 0.0;
}

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

// Amplitude features of the hop slice x[n0 .. n0 + H) (SA.cpp:865-873, 1760-1804; mathutils.c:345-357, 606-615;
// Envelopes.inl:14-18): silence flag, peak, rms, and the maximum of a one-pole envelope follower that starts from 0
// at every frame.  PER = H / 64 consecutive samples per thread, loaded once and kept in registers; the follower
// s -> a + c (s - a) is an affine map per sample, so the 64 threads compose their maps with a scan and then replay
// their own samples from the incoming state.
template <int PER, class Sync>
__device__ __forceinline__ void amp_features(const AfxBatchDev& B, const AfxParams& P, const float* __restrict__ mono, const AfxState& st,
                                             int n0, int slot, int gt, int per, double* xch, Sync sync)
{
  const int np = (PER == 32) ? per : PER;             // PER = 32 also serves the other hop sizes (12, 20, 24, 28 samples per thread)
  const int lane = gt & 31, gw = gt >> 5;
  const double c = P.env_coef;
  float xf[PER];
  {
    const int j0 = n0 + gt * np - st.start_off;
    const float* __restrict__ src = mono + st.lead + j0;
    if (j0 >= 0 && j0 + np <= st.audible) {
#pragma unroll
      for (int q = 0; q < PER; ++q) xf[q] = (q < np) ? __ldg(src + q) : 0.0f;
    } else {
#pragma unroll
      for (int q = 0; q < PER; ++q) xf[q] = (q < np && j0 + q >= 0 && j0 + q < st.audible) ? __ldg(src + q) : 0.0f;
    }
  }
  double A = 1.0, Bv = 0.0, e_hop = 0.0, pk_hop = 0.0;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    if (q < np) {
      const double xv = (double)xf[q] * st.fs, a = fabs(xv);     // mdata()
      e_hop += xv * xv; pk_hop = fmax(pk_hop, a);
      Bv = a + c * (Bv - a);
      A *= c;
    }
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
#pragma unroll
  for (int q = 0; q < PER; ++q) if (q < np) { const double a = fabs((double)xf[q] * st.fs); env = a + c * (env - a); emax = fmax(emax, env); }
  // one exchange for both maxima
  pk_hop = warp_max(pk_hop); emax = warp_max(emax);
  if (lane == 0) { xch[8 + gw] = pk_hop; xch[10 + gw] = emax; }
  sync();
  if (gt == 0) {
    const size_t TF = (size_t)B.TF;
    const double level = ev[0] / (double)P.H;
    B.fs[(size_t)FS_AMP_SILENCE * TF + slot] = (level < AFX_SILENCE_LEVEL) ? 1.0 : 0.0;   // mathutils.c:606-615
    B.fs[(size_t)FS_AMP_PEAK * TF + slot] = fmax(xch[8], xch[9]);
    const double r = sqrt(level);
    B.fs[(size_t)FS_AMP_RMS * TF + slot] = (r != r) ? 0.0 : r;
    B.fs[(size_t)FS_AMP_ENV * TF + slot] = fmax(xch[10], xch[11]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent kernel: one CTA per SM, NG frame groups of 64 threads, every group walks frame slots it claims from a
// global counter (SCH at a time).  What this buys over one CTA per 4 frames (the first form of this kernel):
//   * the window, the two FFT twiddle tables and the real-unpack twiddles (40 KB) live in shared memory for the
//     whole kernel -- with 3 x 70 KB of FFT buffers per SM only 28 KB of L1 are left and the tables used to miss
//     half the time (ncu: 53 % L1 hit rate on global loads, long-scoreboard the top stall);
//   * one CTA per SM leaves 96+ registers per thread: the 16 complex points and the 16 magnitudes of a thread
//     stay in registers (the 80-register form spilled 110 loads/stores per frame to local memory);
//   * the real unpack works on bin pairs (k, 1024 - k): X[k] = E + T, X[1024 - k] = conj(E - T) share E and
//     T = W^k O, which halves the unpack arithmetic and the twiddle table (k <= 512).
// A thread owns bins gt + 64 c (c = 0..7) and their mirrors 1024 - (gt + 64 c); the mirror of bin 0 would be the
// Nyquist bin, which the magnitude spectrum does not hold (AudioMath.cpp:497-504), so that slot takes bin 512.
#define SCH 16              // frame slots per claim

template <int NG>
struct SpecSmem {
  static constexpr int BUF = AFX_NBIN + AFX_NBIN / 16;                 // double2 per group
  static constexpr size_t o_t2 = (size_t)NG * BUF * sizeof(double2);  // [15][16]
  static constexpr size_t o_t3 = o_t2 + 240 * sizeof(double2);        // [3][256]
  static constexpr size_t o_tw = o_t3 + 768 * sizeof(double2);        // exp(-2 pi i k / 2048), k = 0..512 (padded to 528)
  static constexpr size_t o_win = o_tw + 528 * sizeof(double2);       // [2048] Hann * 2
  static constexpr size_t o_xch = o_win + AFX_NFFT * sizeof(double);  // [NG][24] reduction scratch
  static constexpr size_t o_claim = o_xch + (size_t)NG * 24 * sizeof(double);   // [NG][2] claimed slot (double buffered)
  static constexpr size_t bytes = o_claim + (size_t)NG * 2 * sizeof(int);
};

// FULLC: also reduce the full-band sums (centroid of mag[0..1023], the pitch kernel's fail-safe f0) and carry the
// amplitude features' code; the instantiation without it serves contexts that ask for neither (BASELINE config 2)
template <int NG, bool FULLC>
__global__ void __launch_bounds__(SG * NG, 1) k_spectrum(AfxBatchDev B, AfxParams P, unsigned features, unsigned int* __restrict__ work_ctr)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using L = SpecSmem<NG>;
  const int g = threadIdx.x / SG, gt = threadIdx.x % SG, lane = gt & 31, gw = gt >> 5;
  double2* buf = reinterpret_cast<double2*>(smem_raw) + g * L::BUF;   // FFT buffer, later mag[1024]
  double2* s_t2 = reinterpret_cast<double2*>(smem_raw + L::o_t2);
  double2* s_t3 = reinterpret_cast<double2*>(smem_raw + L::o_t3);
  double2* s_tw = reinterpret_cast<double2*>(smem_raw + L::o_tw);
  double2* s_win2 = reinterpret_cast<double2*>(smem_raw + L::o_win);
  double* xch = reinterpret_cast<double*>(smem_raw + L::o_xch) + g * 24;
  volatile int* claim = reinterpret_cast<int*>(smem_raw + L::o_claim) + g * 2;

  for (int i = threadIdx.x; i < 240; i += SG * NG) s_t2[i] = __ldg(P.t.fft_t2 + i);
  for (int i = threadIdx.x; i < 768; i += SG * NG) s_t3[i] = __ldg(P.t.fft_t3_1024 + i);
  for (int i = threadIdx.x; i <= 512; i += SG * NG) s_tw[i] = __ldg(P.t.tw2048 + i);
  // the shared-memory copy of the window carries the magnitude scale 0.5 / N = 2^-12 of the pair unpack (a power of
  // two: the spectrum is the same to the bit, one multiply per bin less)
  for (int i = threadIdx.x; i < AFX_NBIN; i += SG * NG) {
    const double2 w = __ldg(reinterpret_cast<const double2*>(P.t.window) + i);
    s_win2[i] = make_double2(w.x * (0.5 / AFX_NFFT), w.y * (0.5 / AFX_NFFT));
  }
  __syncthreads();

  const int TF = B.TF;
  FftSyncNamed<SG> sync{ 1 + g };
  for (int it = 0;; ++it) {
    if (gt == 0) claim[it & 1] = (int)atomicAdd(work_ctr, (unsigned)SCH);
    sync();
    const int rel0 = claim[it & 1];
    if (rel0 >= B.g_slots) break;                      // group-uniform
    const int rel1 = min(rel0 + SCH, B.g_slots);
    double S1p = 0.0, S2p = 0.0;                       // window sums of the previous frame of this chunk (flux)
    for (int rel = rel0; rel < rel1; ++rel) {
  const int slot = B.slot0 + rel;
  const int fi = B.slot_file[slot];
  const AfxFile* __restrict__ fp = B.files + fi;
  const AfxState st = B.state[fi];
  const int t = slot - fp->frame_off;
  if (fp->status != 0 || t >= st.F) continue;
  const int n0 = t * P.H;
  const float* __restrict__ mono = B.mono + fp->mono_off;

  // ---- amplitude features of the hop slice (SA.cpp:865-873): H / 64 consecutive samples per thread --------
  if (FULLC && (features & AFX_FEAT_AMPLITUDE)) {
    const int per = P.H / SG;                         // 4, 8, ..., 32 (hop is a multiple of 256)
    if (per == 16) amp_features<16>(B, P, mono, st, n0, slot, gt, per, xch, sync);
    else if (per == 8) amp_features<8>(B, P, mono, st, n0, slot, gt, per, xch, sync);
    else amp_features<32>(B, P, mono, st, n0, slot, gt, per, xch, sync);         // 4, 12, 20, 24, 28 or 32 samples per thread
  }

  // ---- load, window, pack (even -> re, odd -> im) in the FFT's strided order; transform ---------------------
  double2 v[16];
  {
    const int j0 = n0 - st.start_off;                 // frame start relative to the first audible sample
    const float* __restrict__ src = mono + st.lead + j0;
    if (j0 >= 0 && j0 + AFX_NFFT <= st.audible) {     // whole frame inside the audible span (the common case)
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int m = gt + SG * r;
        const double2 w = s_win2[m];
        v[r] = make_double2(((double)__ldg(src + 2 * m) * st.fs) * w.x, ((double)__ldg(src + 2 * m + 1) * st.fs) * w.y);
      }
    } else {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int m = gt + SG * r;
        const double2 w = s_win2[m];
        v[r] = make_double2(mdata(mono, st, n0 + 2 * m) * w.x, mdata(mono, st, n0 + 2 * m + 1) * w.y);
      }
    }
  }
  fft16_run<AFX_NBIN, FftSyncNamed<SG>, true>(v, buf, FftTw{ s_t2, s_t3 }, gt, sync);

  // ---- real unpack on bin pairs + magnitude / N (Fourier.cpp:266-271, AudioMath.cpp:497-504) ----------------
  // with Zk = Z[k], Zc = conj(Z[1024 - k]):  2 E = Zk + Zc,  2 O = (Zk - Zc) / i,  2 X[k] = 2E + W^k 2O,
  // 2 X[1024 - k] = conj(2E - W^k 2O); the halves are folded into the final scale
  double m16[16];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int k = gt + SG * c;
    const double2 zk = buf[FFT_PHYS(k)], zc = buf[FFT_PHYS((AFX_NBIN - k) & (AFX_NBIN - 1))];
    const double2 E = make_double2(zk.x + zc.x, zk.y - zc.y);
    const double2 O = make_double2(zk.y + zc.y, zc.x - zk.x);
    const double2 T = f_mul(s_tw[k], O);
    const double2 Xa = f_add(E, T), Xb = f_sub(E, T);
    m16[c] = sqrt_mag(Xa.x * Xa.x + Xa.y * Xa.y);
    m16[8 + c] = sqrt_mag(Xb.x * Xb.x + Xb.y * Xb.y);
  }
  if (gt == 0) { const double2 z = buf[FFT_PHYS(AFX_NBIN / 2)]; m16[8] = 2.0 * sqrt(z.x * z.x + z.y * z.y); }
  const int kmir0 = (gt == 0) ? AFX_NBIN / 2 : AFX_NBIN - gt;          // bin held by m16[8]
  double* gmag = B.mag + (size_t)(slot - B.slot0) * AFX_NBIN;
  sync();                                            // everyone has read Z before buf becomes the magnitude array
  double* mag = reinterpret_cast<double*>(buf);
  double acc[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };     // S1, S2, S3, S4, SJ, log-sum over the analysis window; sum m * m_prev; full-band S, SJ
  {
    const double dgt = (double)gt;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int k = gt + SG * c, km = (c == 0) ? kmir0 : AFX_NBIN - k;
      const double dk = dgt + (double)(SG * c), dkm = (c == 0) ? (double)kmir0 : (double)AFX_NBIN - dk;
      mag[k] = m16[c]; gmag[k] = m16[c];
      mag[km] = m16[8 + c]; gmag[km] = m16[8 + c];
      if (FULLC) {
        acc[7] += m16[c] + m16[8 + c];
        acc[8] = fma(dk, m16[c], fma(dkm, m16[8 + c], acc[8]));
      }
    }
  }
  sync();                                            // mag[] is complete

  // ---- one pass over the analysis window (bins first_bin .. first_bin + nbins - 1, nbins <= 768): thread gt owns
  // window bins j = gt + 64 c for the power sums and the RB consecutive bins RB gt .. RB gt + RB - 1 for the rolloff.
  // RB = 13 (not 12): an odd count of doubles between the lanes' runs spreads them over 16 bank pairs (2-way, the
  // minimum for 64-bit loads); 12 put all 32 lanes on 4 (8-way) ------------------------------------------------------
  constexpr int nb = AFX_WIN_BINS, fb = AFX_WIN_FIRST;      // == P.nbins, P.first_bin (checked by afx_create)
  const double dj0 = (double)gt;
  // spectral flux (Statistics.cpp:578-638, SA.cpp:936-940, 1919-1933) = Pearson correlation with the previous
  // frame's window.  Inside a chunk the previous row is the one this group wrote last iteration (read back through
  // L2; its S1 / S2 are carried); a file's first frame correlates with itself; the first slot of a chunk is left to
  // k_flux, which runs after this kernel over those slots only.
  const bool flux_here = rel > rel0 || t == 0;
  const bool flux_prev = rel > rel0 && t > 0;
  constexpr int RB = 13;
  double mj[12], pj[12], m12[RB], loc = 0.0;
#pragma unroll
  for (int c = 0; c < 12; ++c) pj[c] = (flux_prev && gt + SG * c < nb) ? __ldcg(gmag - AFX_NBIN + fb + gt + SG * c) : 0.0;
#pragma unroll
  for (int c = 0; c < 12; ++c) mj[c] = (gt + SG * c < nb) ? mag[fb + gt + SG * c] : 0.0;
  {
    double mant = 1.0; int ex = 0;
#pragma unroll
    for (int c = 0; c < 12; ++c) {
      const double m = mj[c], m2 = m * m, jm = (dj0 + (double)(SG * c)) * m;
      acc[0] += m; acc[1] = fma(m, m, acc[1]); acc[2] = fma(m2, m, acc[2]); acc[3] = fma(m2, m2, acc[3]); acc[4] += jm;
      acc[6] = fma(m, pj[c], acc[6]);
      if (gt + SG * c < nb) {                                                  // Statistics.cpp:417-455: product with the exponents peeled off
        const double v = fabs(m) + 1e-20;
        const int hi = __double2hiint(v);
        ex += ((hi >> 20) & 0x7ff) - 1022;
        mant *= __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(v));
      }
    }
    // log of the product = log(product of the 64 threads' mantissas) + ln 2 x (sum of their exponents): the mantissa
    // product cannot underflow (12 factors >= 1/2 per thread: >= 2^-768 over the group), so ONE log per frame (thread 0)
    // replaces one per thread; the exponent sum rides in the FP64 reduction (exact)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mant *= __shfl_xor_sync(0xffffffffu, mant, o);
    if (lane == 0) xch[20 + gw] = mant;              // published by group_sum's first barrier
    acc[5] = (double)ex;
  }
#pragma unroll
  for (int q = 0; q < RB; ++q) { m12[q] = (RB * gt + q < nb) ? mag[fb + RB * gt + q] : 0.0; loc += m12[q]; }
  double inc = loc;                                  // rolloff: inclusive scan of the 12-bin sums inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double pv = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += pv; }
  if (lane == 31 && gw == 0) xch[18] = inc;          // published by group_sum's first barrier
  if (FULLC) group_sum<9>(acc, xch, gt, sync);
  else group_sum<7>(reinterpret_cast<double (&)[7]>(acc), xch, gt, sync);
  const double S1 = acc[0];
  // 1 / S1 by reciprocal seed + two Newton steps (every thread needs the centroid; S1 is 0 or >= 1e-56: sums of
  // magnitudes of float32-derived spectra), good to an ulp or two -- the IEEE division costs five times the instructions
  double rS1;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rS1) : "d"(S1));
  rS1 = fma(fma(-S1, rS1, 1.0), rS1, rS1);
  rS1 = fma(fma(-S1, rS1, 1.0), rS1, rS1);
  const double cen = (S1 == 0.0) ? 0.0 : acc[4] * rS1;                          // Statistics.cpp:459-477
  double sp[1] = { 0.0 };
#pragma unroll
  for (int c = 0; c < 12; ++c) { const double d = (dj0 + (double)(SG * c)) - cen; sp[0] = fma(d * d, mj[c], sp[0]); }
  {
    // rolloff (LibXtract scalar.c:472-493): count of prefixes below 85 % of the total
    const double pivot = S1 * (85.0 / 100.0);
    double pre = ((gw == 1) ? xch[18] : 0.0) + inc - loc;     // exclusive prefix = sum of the window bins before RB gt
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < RB; ++q) if (RB * gt + q < nb) { cnt += (pre < pivot) ? 1 : 0; pre += m12[q]; }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) reinterpret_cast<int*>(xch + 19)[gw] = cnt;                // published by the next group_sum
  }
  group_sum<1>(sp, xch, gt, sync);
  if (gt == 0) {
    const int* ix = reinterpret_cast<const int*>(xch + 19);
    const double n = (double)nb;
    B.fs[(size_t)FS_SPEC_ROLLOFF * TF + slot] = (double)(ix[0] + ix[1]) * (double)(P.sr / (P.N / 2));   // SA.cpp:1892: 44100 / 1024 = 43
    const double spread = (S1 == 0.0) ? 0.0 : sp[0] * rS1;                     // Statistics.cpp:486-506
    // skewness / kurtosis (Statistics.cpp:510-554): (1/n) sum ((m_j - cen) / spread)^p from the power sums S1..S4
    // (binomial expansion; the terms cannot cancel to nothing because cen >= 1 > m_j except for spectra
    // concentrated in the first two window bins, where all other bins contribute -cen each)
    const bool have_sk = fabs(spread) > (double)1e-12f;
    const double S2 = acc[1], S3 = acc[2], S4 = acc[3];
    const double c2 = cen * cen, c3 = c2 * cen, c4 = c2 * c2;
    const double mu3 = ((S3 - 3.0 * cen * S2) + 3.0 * c2 * S1) - n * c3;
    const double mu4 = (((S4 - 4.0 * cen * S3) + 6.0 * c2 * S2) - 4.0 * c3 * S1) + n * c4;
    const double inv3 = 1.0 / (spread * spread * spread * n);                  // two divisions for both moments
    const double rms = sqrt(S2 * (1.0 / AFX_WIN_BINS));
    B.fs[(size_t)FS_SPEC_RMS * TF + slot] = (rms != rms) ? 0.0 : rms;
    B.fs[(size_t)FS_SPEC_CENTROID * TF + slot] = cen;
    B.fs[(size_t)FS_SPEC_SPREAD * TF + slot] = spread;
    B.fs[(size_t)FS_SPEC_SKEW * TF + slot] = have_sk ? mu3 * inv3 : 0.0;
    B.fs[(size_t)FS_SPEC_KURT * TF + slot] = have_sk ? (mu4 * inv3) / spread - 3.0 : 0.0;
  }
  // the other half of the frame's scalar tail runs on the first thread of the group's SECOND warp, next to the one above (one
  // thread doing both was a tenth of this kernel's time)
  if (gt == 32) {
    const double n = (double)nb;
    const double S2 = acc[1];
    // flatness (SA.cpp:129-133, 1898-1913): min(LinToDb(gmean / mean) / -60, 1) with gmean = exp(log-sum / n), taken in
    // the log domain: log(gmean / mean) = log-sum / n - log(mean) (one log instead of exp + division + log).
    // LinToDb's branches: ratio <= 1e-12f -> -200 dB; an all-zero window has ratio 0 (Statistics.cpp:568-571) -> 1.0
    const double lsum = log(xch[20] * xch[21]) + acc[5] * 0.693147180559945309417;
    const double mean = S1 * (1.0 / AFX_WIN_BINS);
    double fl = 1.0;
    if (mean != 0.0) {
      const double lg = lsum * (1.0 / AFX_WIN_BINS) - log(mean);
      const double db = (lg > -27.631021119924352) ? lg * (20.0 / 2.302585092994045684) : -200.0;   // log((double)1e-12f)
      fl = fmin(db / -60.0, 1.0);
    }
    B.fs[(size_t)FS_SPEC_FLATNESS * TF + slot] = (fl != fl) ? 0.0 : fl;
    if (FULLC) B.cent_full[slot] = (acc[7] == 0.0) ? 0.0 : acc[8] / acc[7];
    if (flux_here)
      B.fs[(size_t)FS_SPEC_FLUX * TF + slot] = flux_from_sums(S1, S2, flux_prev ? S1p : S1, flux_prev ? S2p : S2, flux_prev ? acc[6] : S2, n);
    // degenerate in the reference (see oracle/afec_oracle.c, "harmonic spectrum"): always 0
    B.fs[(size_t)FS_SPEC_INHARM * TF + slot] = 0.0;
    B.fs[(size_t)FS_TRISTIM1 * TF + slot] = 0.0;
    B.fs[(size_t)FS_TRISTIM2 * TF + slot] = 0.0;
    B.fs[(size_t)FS_TRISTIM3 * TF + slot] = 0.0;
  }
  S1p = S1; S2p = acc[1];
  // no barrier here: the next frame's first shared-memory writes (xch[0..1, 4..5], the FFT buffer) touch nothing
  // that is still read after the last group_sum (only xch[19], by thread 0, and xch[20..21], by thread 32 -- rewritten only
  // behind the next frame's FFT barriers)
    }
  }
}

// spectral flux = Pearson correlation with the previous frame's spectrum (first frame: itself),
// Statistics.cpp:578-638, SA.cpp:936-940, 1919-1933.  TPF = 64 threads per frame (two warps, as k_spectrum's frame
// groups), 256 / TPF frames per CTA, with the summation order of k_spectrum's fused flux (window bin j = gt + TPF c in
// c order, xor butterfly, warp 0 + warp 1), so a frame's flux does not depend on which of the two kernels produced it.
// `stride`: the persistent spectrum kernels leave the first slot of every chunk of `stride` slots to this kernel.
template <int TPF>
__global__ void __launch_bounds__(256) k_flux(AfxBatchDev B, AfxParams P, int stride)
{
  constexpr int FPC = 256 / TPF, NC = 768 / TPF;
  __shared__ double xs[FPC][5][2];
  const int g = threadIdx.x / TPF, gt = threadIdx.x % TPF, lane = gt & 31, gw = gt >> 5;
  const int slot = B.slot0 + (blockIdx.x * FPC + g) * stride;
  bool live = slot < B.slot0 + B.g_slots;
  int t = 0;
  if (live) {
    const int fi = B.slot_file[slot];
    t = slot - B.files[fi].frame_off;
    live = B.files[fi].status == 0 && t < B.state[fi].F;
  }
  double v[5] = { 0, 0, 0, 0, 0 };             // S1, S2, S1 prev, S2 prev, S12
  if (live) {
    const double* a = B.mag + (size_t)(slot - B.slot0) * AFX_NBIN + P.first_bin;
    const double* b = (t > 0) ? a - AFX_NBIN : a;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int j = gt + TPF * c;
      const double x = (j < P.nbins) ? a[j] : 0.0, y = (j < P.nbins) ? b[j] : 0.0;
      v[0] += x; v[1] = fma(x, x, v[1]); v[2] += y; v[3] = fma(y, y, v[3]); v[4] = fma(x, y, v[4]);
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) v[k] = warp_sum(v[k]);
  if (TPF == 64) {
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) xs[g][k][gw] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 5; ++k) v[k] = xs[g][k][0] + xs[g][k][1];
  }
  if (live && gt == 0) B.fs[(size_t)FS_SPEC_FLUX * B.TF + slot] = flux_from_sums(v[0], v[1], v[2], v[3], v[4], (double)P.nbins);
}

template <int NG, bool FULLC>
static void launch_spectrum_p(const AfxParams& P, const AfxBatchDev& B, unsigned features, cudaStream_t s, int sms)
{
  const int smem = (int)SpecSmem<NG>::bytes;
  cudaFuncSetAttribute(k_spectrum<NG, FULLC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device, see afx_pitch.cu
  const int groups_needed = (B.g_slots + SCH - 1) / SCH;
  const int grid = std::max(1, std::min(sms, (groups_needed + NG - 1) / NG));
  cudaMemsetAsync(P.t.work_ctr, 0, sizeof(unsigned int), s);
  k_spectrum<NG, FULLC><<<grid, SG * NG, smem, s>>>(B, P, features, P.t.work_ctr);
}

void afx_launch_spectrum(const AfxParams& P, const AfxBatchDev& B, unsigned features, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (features & (AFX_FEAT_PITCH | AFX_FEAT_AMPLITUDE)) launch_spectrum_p<10, true>(P, B, features, s, sms);
  else launch_spectrum_p<10, false>(P, B, features, s, sms);
  ++*launches;
  const int stride = SCH, nflux = (B.g_slots + stride - 1) / stride;
  k_flux<64><<<(nflux + 3) / 4, 256, 0, s>>>(B, P, stride); ++*launches;
}

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
// 128 threads per frame.  The zero-padded first half a and the full frame b are transformed
// together as z = a + i b by ONE 2048-point complex FFT (register-blocked radix 16 x 16 x 8, afx_fft16.cuh),
// split into A and B; the real correlation r = IFFT(conj(A) B) comes back through a HALF-size (1024-point) transform.
// Shared memory: one padded 2048-point FFT buffer (34 KB) and yin' (9 KB).  The prefix sums of squares live in the FFT
// buffer before the first transform: the windowed square sums sq[tau] they are needed for go to the yin array up front.
#include "afx_fft16.cuh"
#include <algorithm>
#include <cstdlib>

#define YT 128
#define YN 2048
#define YW 1024
#define YCH 4               // frame slots per claim
#define PAD16(i) ((i) + ((i) >> 4))
#define PAD8(i) ((i) + ((i) >> 3))

// Persistent form (see afx_spectrum.cu): one CTA per SM, NG frame groups of 128 threads with a named barrier each,
// frame slots claimed from a global counter.  The FFT twiddle tables (32 KB) stay in shared memory, and a group
// PREFETCHES its next frame while it works on the current one: the slot -> file -> state -> samples chain of
// dependent global loads was 39 % of the stall samples of the one-CTA-per-frame form (ncu, long scoreboard).
template <int NG>
struct PitchSmem {
  static constexpr int BUF = YN + YN / 16;                             // double2 per group
  static constexpr int YIN = YW + YW / 8 + 8;                          // doubles per group
  static constexpr size_t group_bytes = (size_t)BUF * sizeof(double2) + (size_t)YIN * sizeof(double) + 40 * sizeof(double);
  static constexpr size_t o_t2 = (size_t)NG * group_bytes;            // [15][16]
  static constexpr size_t o_t3 = o_t2 + 240 * sizeof(double2);        // [7][256]
  static constexpr size_t bytes = o_t3 + 7 * 256 * sizeof(double2);
};

// exclusive prefix sum over the 128 threads of a group; scratch: 4 doubles
template <class Sync>
__device__ __forceinline__ double group_scan_excl(double v, double* scratch, int tid, Sync sync)
{
  const int lane = tid & 31, wid = tid >> 5;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double p = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += p; }
  sync();
  if (lane == 31) scratch[wid] = inc;
  sync();
  double base = 0.0;
#pragma unroll
  for (int w = 0; w < (YT >> 5) - 1; ++w) { const double sv = scratch[w]; if (w < wid) base += sv; }
  return base + inc - v;
}

struct PitchNext {          // what a group knows about the frame it will work on next
  int slot;                 // global slot, -1 = none
  bool live;
  double fs;
  float x[16];              // raw mono samples tid + 128 r of the frame (0 outside the audible span)
};

template <int NG>
__global__ void __launch_bounds__(YT * NG, 1) k_pitch(AfxBatchDev B, AfxParams P, unsigned int* __restrict__ work_ctr)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using L = PitchSmem<NG>;
  const int g = threadIdx.x / YT, tid = threadIdx.x % YT;
  unsigned char* gbase = smem_raw + (size_t)g * L::group_bytes;
  double2* buf = reinterpret_cast<double2*>(gbase);                    // [2048 + 128]
  double* S = reinterpret_cast<double*>(buf);                          // [PAD16(2048) + 1] prefix sums of squares (before the FFTs)
  double2* hbuf = buf;                                                 // the half-size inverse runs in the first half of buf
  double* yin = reinterpret_cast<double*>(buf + L::BUF);               // [PAD8(1024)]
  double* scratch = yin + L::YIN;                                      // [8] scans / argmin
  double* level = scratch + 8;                                         // [2] sum of squares of the frame / of the hop
  int* iscr = reinterpret_cast<int*>(scratch + 12);                    // [8] argmin / first dip
  volatile int* claim = reinterpret_cast<int*>(scratch + 20);          // [2] claimed chunk (double buffered)
  double2* s_t2 = reinterpret_cast<double2*>(smem_raw + L::o_t2);
  double2* s_t3 = reinterpret_cast<double2*>(smem_raw + L::o_t3);
  for (int i = threadIdx.x; i < 240; i += YT * NG) s_t2[i] = __ldg(P.t.fft_t2 + i);
  for (int i = threadIdx.x; i < 7 * 256; i += YT * NG) s_t3[i] = __ldg(P.t.fft_t3_2048 + i);
  __syncthreads();
  FftSyncNamed<YT> sync{ 1 + g };
  const FftTw ftw = { s_t2, s_t3 };
  const size_t TF = (size_t)B.TF;

  // metadata + samples of frame slot `rel` (relative to the launch group) into nx
  auto fetch = [&](int rel, PitchNext& nx) {
    nx.slot = -1; nx.live = false; nx.fs = 0.0;
    if (rel < 0 || rel >= B.g_slots) return;
    const int slot = B.slot0 + rel;
    nx.slot = slot;
    const int fi = B.slot_file[slot];
    const AfxFile* __restrict__ fp = B.files + fi;
    const AfxState* __restrict__ sp = B.state + fi;
    const int t = slot - fp->frame_off;
    if (fp->status != 0 || t >= sp->F) return;
    nx.live = true; nx.fs = sp->fs;
    const int j0 = t * P.H - sp->start_off, audible = sp->audible;
    const float* __restrict__ src = B.mono + fp->mono_off + sp->lead + j0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int m = tid + YT * r, j = j0 + m;
      nx.x[r] = (j >= 0 && j < audible) ? __ldg(src + m) : 0.0f;
    }
  };

  int it = 0;
  if (tid == 0) claim[0] = (int)atomicAdd(work_ctr, (unsigned)YCH);
  sync();
  int rel = claim[0], rel_end = rel + YCH;
  PitchNext cur;
#pragma unroll
  for (int r = 0; r < 16; ++r) cur.x[r] = 0.0f;
  fetch(rel, cur);
  while (rel < B.g_slots) {
    // the slot after this one: next in the chunk, else the head of a freshly claimed chunk
    int nrel = rel + 1, nrel_end = rel_end;
    if (nrel >= rel_end) {
      ++it;
      if (tid == 0) claim[it & 1] = (int)atomicAdd(work_ctr, (unsigned)YCH);
      sync();
      nrel = claim[it & 1]; nrel_end = nrel + YCH;
    }
    if (!cur.live) {                                 // group-uniform
      fetch(nrel, cur); rel = nrel; rel_end = nrel_end;
      continue;
    }
    const int slot = cur.slot;

    // ---- z = a + i b in the FFT's strided order, squares to shared memory --------------------------------
    double2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int m = tid + YT * r;
      const double xv = (double)cur.x[r] * cur.fs;              // mdata(): SA.cpp:712-718
      v[r] = make_double2(m < YW ? xv : 0.0, xv);
      S[PAD16(m)] = xv * xv;
    }
    sync();
    // ---- prefix sums of squares: 16 consecutive samples per thread (padded -> conflict free) -----------------
    {
      double q2[16]; double loc = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) { q2[q] = S[PAD16(16 * tid + q)]; loc += q2[q]; }
      double pre = group_scan_excl(loc, scratch, tid, sync);     // its first barrier: every square has been read
      if (tid == 0) S[0] = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) { pre += q2[q]; S[PAD16(16 * tid + q + 1)] = pre; }
    }
    sync();
    // windowed square sums sq[tau] = sum_{j<W} x[j+tau]^2 + sum_{j<W} x[j]^2 -> yin[] (the difference function subtracts r later)
    {
      const double sW = S[PAD16(YW)];
#pragma unroll
      for (int c = 0; c < YW / YT; ++c) {
        const int tau = tid + YT * c;
        yin[PAD8(tau)] = (S[PAD16(tau + YW)] - S[PAD16(tau)]) + sW;
      }
      if (tid == 0) { level[0] = S[PAD16(YN)]; level[1] = S[PAD16(P.H)]; }
    }
    sync();                                                      // S is dead: the FFT buffer takes its place
    fft16_run<YN, FftSyncNamed<YT>, true>(v, buf, ftw, tid, sync);
    // ---- r = IFFT_2048(P), P[k] = conj(A[k]) B[k] with A, B the transforms of a and b split out of Z.  r is real, so
    // the inverse runs at HALF size (Hermitian P): with z[m] = r[2m] + i r[2m+1],
    //   z = IFFT_1024(Zc),  Zc[k] = (P[k] + conj(P[1024-k])) / 2 + i W^-k (P[k] - conj(P[1024-k])) / 2,  W = exp(-2 pi i / 2048).
    // All 128 threads build Zc pairwise (k, 1024 - k share P[k] and P[1024-k]) IN PLACE: a pair reads Z at
    // {k, 1024-k, 1024+k, 2048-k} -- no other pair touches these -- and overwrites slots k and 1024-k.  64 threads then
    // run the 1024-point transform (as conj(FFT(conj(Zc))) / 1024) in the first half of the buffer -- a quarter of the
    // shared-memory traffic of the full-size inverse, which is what bounds this kernel.
    {
      auto pk = [&](int k, int kc) {                              // P[k] from Z[k] and Z[kc], kc = (2048 - k) mod 2048
        const double2 z1 = buf[FFT_PHYS(k)], z2 = buf[FFT_PHYS(kc)];
        // 2 A and 2 B: the halves (and those of e1 / d1 below) are powers of two and go into the final scale 1 / (8 W)
        const double2 A = make_double2(z1.x + z2.x, z1.y - z2.y);
        const double2 Bc = make_double2(z1.y + z2.y, z2.x - z1.x);                            // (z1 - conj(z2)) / i
        return f_mul(make_double2(A.x, -A.y), Bc);
      };
      auto pair = [&](int k) {                                    // 0 <= k <= 512
        const double2 p1 = pk(k, (YN - k) & (YN - 1)), p2 = pk(YW - k, YW + k);
        const double2 w = __ldg(P.t.tw2048 + k);                  // W^k; W^-k = conj
        const double2 e1 = make_double2(p1.x + p2.x, p1.y - p2.y);                           // P[k] + conj(P[k'])   (x 8: see pk)
        const double2 d1 = make_double2(p1.x - p2.x, p1.y + p2.y);                           // P[k] - conj(P[k'])
        const double2 o1 = f_mul(make_double2(w.x, -w.y), d1);                                 // W^-k d1
        const double2 zc1 = make_double2(e1.x - o1.y, e1.y + o1.x);                            // e1 + i o1
        hbuf[FFT_PHYS(k)] = make_double2(zc1.x, -zc1.y);                                       // conj(Zc[k])
        if (k > 0) {
          // Zc[k'] = conj(e1) + i (-W^k) (-conj(d1)) = conj(e1) + i W^k conj(d1)
          const double2 o2 = f_mul(w, make_double2(d1.x, -d1.y));
          const double2 zc2 = make_double2(e1.x - o2.y, -e1.y + o2.x);
          hbuf[FFT_PHYS(YW - k)] = make_double2(zc2.x, -zc2.y);
        }
      };
#pragma unroll
      for (int c = 0; c < 4; ++c) pair(tid + YT * c);
      if (tid == 0) pair(YW / 2);
    }
    sync();
    if (tid < 64) {
      FftSyncNamed<64> hsync{ 8 + g };
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = hbuf[FFT_PHYS(tid + 64 * r)];
      hsync();                                                     // every input is in registers before hbuf is rewritten
      fft16_run<YW, FftSyncNamed<64>, true, 2>(v, hbuf, ftw, tid, hsync);
    }
    sync();

    // ---- the next frame's loads go out now and land while this frame's search runs -------------------------
    PitchNext nxt;
    fetch(nrel, nxt);

    // ---- difference function (elementwise, tau = tid + 128 c) ------------------------------------------
#pragma unroll
    for (int c = 0; c < YW / YT; ++c) {
      const int tau = tid + YT * c;
      const double2 f = hbuf[FFT_PHYS(tau >> 1)];                // z[m] = conj(F[m]) / 1024: r[2m] = Re, r[2m+1] = Im
      yin[PAD8(tau)] = yin[PAD8(tau)] - ((tau & 1) ? -f.y : f.x) * (0.125 / YW);
    }
    sync();
    // ---- cumulative-mean normalisation: 8 consecutive tau per thread ---------------------------------------
    double y[8]; double ysum = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { y[q] = yin[PAD8(8 * tid + q)]; if (8 * tid + q >= 1) ysum += y[q]; }
    double run = group_scan_excl(ysum, scratch, tid, sync);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int tau = 8 * tid + q;
      double vv;
      if (tau == 0) vv = 1.0;
      else { run += y[q]; vv = (run != 0.0) ? y[q] * ((double)tau / run) : 1.0; }
      y[q] = vv;
      yin[PAD8(tau)] = vv;
    }
    sync();

    // ---- first dip below the tolerance, else the last global minimum ---------------------------------
    int cand = 0x7fffffff;
#pragma unroll
    for (int q = 7; q >= 0; --q) {
      const int p = 8 * tid + q;
      const double nx1 = (q < 7) ? y[q + 1] : yin[PAD8(min(p + 1, YW - 1))];
      if (p >= 2 && p <= YW - 4 && y[q] < 0.75 && y[q] < nx1) cand = p;
    }
    // argmin with ties -> last index (mathutils.c:250-258), evaluated alongside: one exchange serves both
    double mv = y[0]; int mi = 8 * tid;
#pragma unroll
    for (int q = 1; q < 8; ++q) if (!(mv < y[q])) { mv = y[q]; mi = 8 * tid + q; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
      const double ov = __shfl_xor_sync(0xffffffffu, mv, o); const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (ov < mv || (ov == mv && oi > mi)) { mv = ov; mi = oi; }
    }
    if ((tid & 31) == 0) { scratch[4 + (tid >> 5)] = mv; iscr[tid >> 5] = mi; iscr[4 + (tid >> 5)] = cand; }
    sync();
    if (tid == 0) {
      int pos;
      cand = min(min(iscr[4], iscr[5]), min(iscr[6], iscr[7]));
      if (cand != 0x7fffffff) pos = cand;
      else {
        mv = scratch[4]; mi = iscr[0];
        for (int w = 1; w < (YT >> 5); ++w) { const double ov = scratch[4 + w]; const int oi = iscr[w]; if (ov < mv || (ov == mv && oi > mi)) { mv = ov; mi = oi; } }
        pos = mi;
      }
      double period;
      if (pos == 0 || pos == YW - 1) period = (double)pos;          // mathutils.c:494-506
      else { const double s0 = yin[PAD8(pos - 1)], s1 = yin[PAD8(pos)], s2 = yin[PAD8(pos + 1)]; period = pos + .5 * (s0 - s2) / (s0 - 2. * s1 + s2); }
      unsigned peak_pos = 0;
      if (period == period && period >= 0.0 && period < (double)YW) peak_pos = (unsigned)period;
      double pitch = (period > 0.0) ? (double)P.sr / (period + 0.) : 0.0;                  // pitch.c:450-462
      const bool silent_frame = (level[0] / (double)YN) < AFX_SILENCE_LEVEL;             // pitch.c:399-407
      if (silent_frame) pitch = 0.0;
      double conf = (1.0 - yin[PAD8(peak_pos)]) / 0.25;                                    // SA.cpp:887-889
      conf = conf < 0.0 ? 0.0 : (conf > 1.0 ? 1.0 : conf);
      double fsafe = 0.0;                                                                  // SA.cpp:897-916
      if (pitch > 0.0 && conf > 0.2) fsafe = pitch;
      else {
        const bool silent_hop = (level[1] / (double)P.H) < AFX_SILENCE_LEVEL;
        if (!silent_hop) { const double c = B.cent_full[slot]; fsafe = (double)P.sr / (double)P.N * (c > 0.0 ? c : 0.0); }
      }
      B.fs[(size_t)FS_F0 * TF + slot] = pitch;
      B.fs[(size_t)FS_F0_CONF * TF + slot] = conf;
      B.fs[(size_t)FS_F0_FAILSAFE * TF + slot] = fsafe;
    }
    sync();                                          // thread 0 is done with yin / level / scratch before the next frame
    cur = nxt; rel = nrel; rel_end = nrel_end;
  }
}

// =====================================================================================================================
// Hop == W == 1024 form (round 2): consecutive frames share a block.  With the frame split into its two 1024-sample blocks
// (A | B) and F_X = FFT_2048(X zero padded), FFT(a) = F_A and FFT(frame)[k] = F_A[k] + (-1)^k F_B[k], so
//     P[k] = conj(F_A[k]) (F_A[k] + (-1)^k F_B[k])
// needs ONE new transform per frame -- of a real, half-empty 2048-sequence, i.e. a 1024-point complex FFT of the packed
// pairs -- because this frame's B is the next frame's A.  Everything is a 1024-point transform on 64 threads, so a frame
// belongs to a group of TWO warps that are busy in every phase (the 2048-point form above parks half of its four warps
// during the inverse).  F_A is carried in REGISTERS: the thread that combines bins k and 1024 - k of one frame combines
// them in the next as well.  The sums of squares are block-local prefix sums I_X[j] = sum_{i<=j} x_i^2, also carried:
//     sq[tau] = (T_A - I_A[tau-1]) + I_B[tau-1] + T_A,   T_A = I_A[1023]
// A group starts a run of consecutive slots of one file with an extra transform of block A.
//
// The kernel is bound by SHARED-MEMORY BANDWIDTH (ncu: 63 % of the wavefront rate in the first version of this form, 70 % in
// the 2048-point kernel), so the layouts are chosen to move every value once: the block lands in shared memory through
// cp.async (no staging registers, no spills) in a padded layout that serves both the FFT's strided pairs and 16 consecutive
// samples per thread; prefix sums are made in registers and written once; the difference function, its cumulative mean
// and the normalisation run in registers on 16 consecutive lags per thread straight from I_A, I_B and the inverse
// transform's output, which the last FFT pass stores -- first half only -- in a layout made for that read.
#define HT 64
#ifndef PITCH_T2P
#define PITCH_T2P false        // measured: the eleven extra products cost more than the eleven table loads they save (15.2 vs 14.7 ns per frame)
#endif
#define HCH 16              // frame slots per claim: one extra transform per claimed run
#define PADX(m) ((m) + 4 * ((m) >> 4))      // raw block: 16 consecutive floats per thread at a 20-float stride

// tau / run for the cumulative-mean normalisation: reciprocal seed + Newton steps on the quotient (~1 ulp; the reference's
// own quotient is one rounding of the same value, and nothing here is bit-exact against its FFT anyway).  den is a sum of
// differences of squared float32-derived samples: zero (handled by the caller) or far inside the normal range.
__device__ __forceinline__ double yin_div(double num, double den)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  r = fma(fma(-den, r, 1.0), r, r);
  const double q = num * r;
  return fma(fma(-den, q, num), r, q);
}

template <int NG>
struct PitchHopSmem {
  static constexpr int BUF = YW + YW / 16;                             // double2 per group (FFT_PHYS; FFT_PAD8(511) fits too)
  static constexpr int YIN = YW + YW / 16 + 8;                         // doubles: PAD16 layout (yin', I_A, I_B)
  static constexpr int XIN = YW + YW / 4;                              // floats: PADX layout
  static constexpr size_t group_bytes = (size_t)BUF * sizeof(double2) + 2 * (size_t)YIN * sizeof(double) + (size_t)XIN * sizeof(float) + 48 * sizeof(double);
  static constexpr size_t o_t2 = (size_t)NG * group_bytes;            // [15][16]
  static constexpr size_t o_t3 = o_t2 + 240 * sizeof(double2);        // [3][256]
  static constexpr size_t o_tw = o_t3 + 3 * 256 * sizeof(double2);    // [513] exp(-2 pi i k / 2048)
  static constexpr size_t bytes = o_tw + 520 * sizeof(double2);
};

struct HopNext {            // what a group knows about the frame it will work on next
  int slot, fi;             // global slot (-1 = none) and its file
  bool live;
  double fs;
};

// global -> shared without a register in between (the block a group will need next is in flight while it still works with
// all of its registers); n = 0 fills with zeros
__device__ __forceinline__ void hop_cp4(float* dst, const float* src, bool pred)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void hop_cp8(float* dst, const float* src)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void hop_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int NG>
__global__ void __launch_bounds__(HT * NG, 1) k_pitch_hop(AfxBatchDev B, AfxParams P, unsigned int* __restrict__ work_ctr)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using L = PitchHopSmem<NG>;
  const int g = threadIdx.x / HT, tid = threadIdx.x % HT;
  unsigned char* gbase = smem_raw + (size_t)g * L::group_bytes;
  double2* buf = reinterpret_cast<double2*>(gbase);                    // packed spectrum of the new block, then Zc, then r (first half)
  double* yin = reinterpret_cast<double*>(buf);                        // [PAD16(1024)] yin': takes the place of r once every thread holds its lags
  double* I0 = reinterpret_cast<double*>(buf + L::BUF);                // [2][PAD16(1024)] block-local inclusive sums of squares:
                                                                       //   I0 + pp * YIN: the carried block (A), the other one: the new block
  float* xin = reinterpret_cast<float*>(I0 + 2 * L::YIN);              // [PADX(1024)] raw samples of the block to transform next (cp.async target)
  double* scratch = reinterpret_cast<double*>(xin + L::XIN);           // [8] scans / argmin
  int* iscr = reinterpret_cast<int*>(scratch + 12);                    // [8] argmin / first dip
  volatile int* claim = reinterpret_cast<int*>(scratch + 20);          // [1] claimed chunk
  double2* s_t2 = reinterpret_cast<double2*>(smem_raw + L::o_t2);
  double2* s_t3 = reinterpret_cast<double2*>(smem_raw + L::o_t3);
  double2* s_tw = reinterpret_cast<double2*>(smem_raw + L::o_tw);
  for (int i = threadIdx.x; i < 240; i += HT * NG) s_t2[i] = __ldg(P.t.fft_t2 + i);
  for (int i = threadIdx.x; i < 3 * 256; i += HT * NG) s_t3[i] = __ldg(P.t.fft_t3_1024 + i);
  for (int i = threadIdx.x; i <= 512; i += HT * NG) s_tw[i] = __ldg(P.t.tw2048 + i);
  __syncthreads();
  FftSyncNamed<HT> sync{ 1 + g };
  const FftTw ftw = { s_t2, s_t3 };
  const int lane = tid & 31, wid = tid >> 5;

  // block `blk` of the frame in nx -> xin, asynchronously; a thread copies samples 2 tid + 128 r (+ 1)
  auto load_block = [&](const HopNext& nx, int blk) {
    const AfxFile* __restrict__ fp = B.files + nx.fi;
    const AfxState* __restrict__ sp = B.state + nx.fi;
    const int t = nx.slot - fp->frame_off;
    const int jb = t * YW - sp->start_off + YW * blk, audible = sp->audible;       // first sample of the block (group-uniform)
    const size_t o = (size_t)fp->mono_off + sp->lead + jb;
    const float* __restrict__ src = B.mono + o + 2 * tid;
    float* dst = xin + PADX(2 * tid);                                                // PADX(2 tid + 128 r) = PADX(2 tid) + 160 r
    if (jb >= 0 && jb + YW <= audible && (o & 1) == 0) {                             // inside the audible span, pairs aligned
#pragma unroll
      for (int r = 0; r < 8; ++r) hop_cp8(dst + 160 * r, src + 128 * r);
    } else {
      const int j0 = jb + 2 * tid;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int j = j0 + 128 * r;
        const bool p0 = j >= 0 && j < audible, p1 = j + 1 >= 0 && j + 1 < audible;
        hop_cp4(dst + 160 * r, p0 ? src + 128 * r : B.mono, p0);
        hop_cp4(dst + 160 * r + 1, p1 ? src + 128 * r + 1 : B.mono, p1);
      }
    }
  };
  auto fetch = [&](int rel, HopNext& nx) {
    nx.slot = -1; nx.fi = -1; nx.live = false; nx.fs = 0.0;
    if (rel < 0 || rel >= B.g_slots) return;
    const int slot = B.slot0 + rel;
    nx.slot = slot;
    const int fi = B.slot_file[slot];
    nx.fi = fi;
    const AfxFile* __restrict__ fp = B.files + fi;
    const AfxState* __restrict__ sp = B.state + fi;
    if (fp->status != 0 || slot - fp->frame_off >= sp->F) return;
    nx.live = true; nx.fs = sp->fs;
    load_block(nx, 1);
  };

  double2 FA[16];           // 2 F_A: FA[2c] = bin k = tid + 64 c, FA[2c+1] = bin 1024 - k; thread 0, c = 0: (F[0], F[1024]) and F[512]
#pragma unroll
  for (int q = 0; q < 16; ++q) FA[q] = make_double2(0.0, 0.0);
  int carried = -1;         // slot whose FIRST block FA and I0[pp] describe
  int pp = 0;

  // loop state is kept small (the transforms run with F_A and 16 points in registers): chunks are aligned, so the end of the
  // claimed run follows from `rel`; the next slot is claimed only where it is needed (before the prefetch)
  if (tid == 0) claim[0] = (int)atomicAdd(work_ctr, (unsigned)HCH);
  sync();
  int rel = claim[0];
  sync();
  HopNext cur;
  fetch(rel, cur);
  auto next_rel = [&](int r) {                       // slot after r: next in the chunk, else the head of a freshly claimed chunk
    int n = r + 1;
    if ((n & (HCH - 1)) == 0) {
      if (tid == 0) claim[0] = (int)atomicAdd(work_ctr, (unsigned)HCH);
      sync();
      n = claim[0];
      sync();                                        // everyone has read the claim before it can be overwritten
    }
    return n;
  };
  while (rel < B.g_slots) {
    if (!cur.live) {                                 // group-uniform
      rel = next_rel(rel);
      fetch(rel, cur);
      continue;
    }
    const int slot = cur.slot;
    const double fs = cur.fs;

    // Three transforms at most, ONE instance of the FFT code (the instruction cache is shared by the groups, each in another
    // phase): step 0 (only at the head of a run) block A -> I_A, F_A; step 1 block B -> I_B, F_B, P, Zc; step 2 the half-size
    // inverse.  The block of steps 0 / 1 arrives in xin (a second block fetched ahead is dropped at the head of a run and
    // read again).
    int step = (carried == slot) ? 1 : 0;
    if (step == 0) { hop_cp_wait(); sync(); load_block(cur, 0); }   // (the copy issued ahead targets the same words: let it land, and be read by nobody, first)
#pragma unroll 1
    for (;; ++step) {
      double2 v[16];
      if (step < 2) {
        hop_cp_wait();
        sync();                                                      // the whole block is in xin
        // ---- block-local inclusive sums of squares: 16 consecutive samples per thread, made in registers ----
        {
          double* __restrict__ In = I0 + (pp ^ 1) * L::YIN;
          double pr[16];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 x4 = *reinterpret_cast<const float4*>(xin + 20 * tid + 4 * q4);     // PADX(16 tid + 4 q4)
            const double d0 = (double)x4.x * fs, d1 = (double)x4.y * fs, d2 = (double)x4.z * fs, d3 = (double)x4.w * fs;   // mdata(): SA.cpp:712-718
            double a = d0 * d0;
            pr[4 * q4] = a; a = fma(d1, d1, a); pr[4 * q4 + 1] = a; a = fma(d2, d2, a); pr[4 * q4 + 2] = a; a = fma(d3, d3, a); pr[4 * q4 + 3] = a;
          }
          const double loc = (pr[3] + pr[7]) + (pr[11] + pr[15]);
          double inc = loc;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const double pv = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += pv; }
          if (lane == 31 && wid == 0) scratch[0] = inc;
          sync();
          double off = inc - loc + (wid ? scratch[0] : 0.0);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) In[17 * tid + 4 * q4 + q] = pr[4 * q4 + q] + off;   // PAD16(16 tid + q)
            off += pr[4 * q4 + 3];
          }
        }
        // ---- packed block: z[m] = x[2m] + i x[2m+1], zero padded ----
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float2 xb = *reinterpret_cast<const float2*>(xin + PADX(2 * tid) + 160 * r);
          v[r] = make_double2((double)xb.x * fs, (double)xb.y * fs);
          v[r + 8] = make_double2(0.0, 0.0);
        }
      } else {
        // ---- r through the half-size inverse (as conj(FFT(conj(Zc))) / 1024) ----
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = buf[FFT_PHYS(tid + HT * r)];
        sync();                                                      // every input is in registers before buf is rewritten
      }
      fft16_run<YW, FftSyncNamed<HT>, true, 1, true, true, PITCH_T2P>(v, buf, ftw, tid, sync, step == 2);
      if (step == 2) break;
      // ---- unpack bins k and 1024 - k of F_B; step 1: P, then Zc of the half-size inverse, in place ----
      // (r real => the inverse of the Hermitian P runs at half size: with z[m] = r[2m] + i r[2m+1], z = IFFT_1024(Zc),
      //  Zc[k] = (P[k] + conj(P[1024-k])) / 2 + i W^-k (P[k] - conj(P[1024-k])) / 2, W = exp(-2 pi i / 2048); every halving
      //  here and in the unpacking is left out and folded into the final scale 1 / (8 W))
      const double2 wt = s_tw[tid];                                                       // W^k = W^tid W^(64 c): one load, seven products
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int k = tid + HT * c;
        const double sgn = (tid & 1) ? -1.0 : 1.0;                                        // (-1)^k for every k = tid + 64 c
        if (c == 0 && tid == 0) {
          const double2 z0 = buf[0], zh = buf[FFT_PHYS(YW / 2)];
          const double2 fb0 = make_double2(2.0 * (z0.x + z0.y), 2.0 * (z0.x - z0.y));     // 2 F_B[0], 2 F_B[1024] (real)
          const double2 fbh = make_double2(2.0 * zh.x, -2.0 * zh.y);                      // 2 F_B[512] = 2 conj(Z[512])
          if (step == 1) {
            const double p0 = FA[0].x * (FA[0].x + fb0.x), pn = FA[0].y * (FA[0].y + fb0.y);
            const double2 ph = f_mul(make_double2(FA[1].x, -FA[1].y), f_add(FA[1], fbh));
            buf[0] = make_double2(p0 + pn, -(p0 - pn));                                   // conj(Zc[0])
            buf[FFT_PHYS(YW / 2)] = make_double2(2.0 * ph.x, 2.0 * ph.y);                 // conj(Zc[512]) = 2 P[512]
          }
          FA[0] = fb0; FA[1] = fbh;
        } else {
          const double2 z1 = buf[FFT_PHYS(k)], z2 = buf[FFT_PHYS(YW - k)];
          constexpr double kC[8] = { 1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440,
                                     0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785 };     // cos(2 pi 64 c / 2048)
          constexpr double kS[8] = { 0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474, 0.70710678118654752440,
                                     0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913 };     // sin(2 pi 64 c / 2048)
          const double2 w = (c == 0) ? wt : f_mul(wt, make_double2(kC[c], -kS[c]));       // W^k; W^-k = conj
          const double2 e = make_double2(z1.x + z2.x, z1.y - z2.y);                       // Z[k] + conj(Z[1024-k])
          const double2 d = make_double2(z1.x - z2.x, z1.y + z2.y);                       // Z[k] - conj(Z[1024-k])
          const double2 o = f_mul(w, d);
          const double2 fbk = make_double2(e.x + o.y, e.y - o.x);                         // 2 F_B[k]      = e - i o
          const double2 fbc = make_double2(e.x - o.y, -(e.y + o.x));                      // 2 F_B[1024-k] = conj(e + i o)
          if (step == 1) {
            const double2 xk = make_double2(fma(sgn, fbk.x, FA[2 * c].x), fma(sgn, fbk.y, FA[2 * c].y));
            const double2 xc = make_double2(fma(sgn, fbc.x, FA[2 * c + 1].x), fma(sgn, fbc.y, FA[2 * c + 1].y));
            const double2 p1 = f_mul(make_double2(FA[2 * c].x, -FA[2 * c].y), xk);
            const double2 p2 = f_mul(make_double2(FA[2 * c + 1].x, -FA[2 * c + 1].y), xc);
            const double2 e1 = make_double2(p1.x + p2.x, p1.y - p2.y);                    // P[k] + conj(P[1024-k])
            const double2 d1 = make_double2(p1.x - p2.x, p1.y + p2.y);                    // P[k] - conj(P[1024-k])
            const double2 o1 = f_mul(make_double2(w.x, -w.y), d1);                        // W^-k d1
            buf[FFT_PHYS(k)] = make_double2(e1.x - o1.y, -(e1.y + o1.x));                 // conj(e1 + i o1)
            const double2 o2 = f_mul(w, make_double2(d1.x, -d1.y));                       // Zc[1024-k] = conj(e1) + i W^k conj(d1)
            buf[FFT_PHYS(YW - k)] = make_double2(e1.x - o2.y, -(-e1.y + o2.x));
          }
          FA[2 * c] = fbk; FA[2 * c + 1] = fbc;
        }
      }
      sync();
      if (step == 0) { pp ^= 1; load_block(cur, 1); }                // block A is now the carried block; xin is free (read before the FFT's barriers)
    }
    carried = -1;

    // ---- the next frame's block goes out now and lands while this frame's search runs -------------------------
    const int nrel = next_rel(rel);
    HopNext nxt;
    fetch(nrel, nxt);
    if (nxt.live && nxt.slot == slot + 1 && nxt.fi == B.slot_file[slot]) carried = nxt.slot;   // its first block is this frame's second

    // ---- difference function and its cumulative mean on 16 consecutive lags per thread, in registers:
    //      yin[tau] = sq[tau] - r[tau],  sq[tau] = (T_A - I_A[tau-1]) + I_B[tau-1] + T_A,  r[2m] = Re z[m], r[2m+1] = Im z[m],
    //      z[m] = conj(F[m]) / 1024 with F at buf[FFT_PAD8(m)]
    const double* __restrict__ IA = I0 + pp * L::YIN;
    const double* __restrict__ IB = I0 + (pp ^ 1) * L::YIN;
    const double TA = IA[PAD16(YW - 1)];
    double y[16]; double ysum = 0.0;
#pragma unroll
    for (int q2 = 0; q2 < 8; ++q2) {
      const double2 f = buf[9 * tid + q2];                         // FFT_PAD8(8 tid + q2)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int q = 2 * q2 + e, tau = 16 * tid + q;
        double sq;
        if (q == 0) {                                              // tau - 1 belongs to the previous thread's 16 (or is -1)
          const int tm = (tid > 0) ? PAD16(16 * tid - 1) : 0;
          const double ia = IA[tm], ib = IB[tm];
          sq = (tid > 0) ? ((TA - ia) + ib) + TA : TA + TA;
        } else sq = ((TA - IA[17 * tid + q - 1]) + IB[17 * tid + q - 1]) + TA;
        y[q] = fma(e ? f.y : f.x, e ? (0.125 / YW) : -(0.125 / YW), sq);
        if (tau >= 1) ysum += y[q];
      }
    }
    double run;
    {
      double inc = ysum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const double pv = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += pv; }
      if (lane == 31 && wid == 0) scratch[1] = inc;
      sync();
      run = inc - ysum + (wid ? scratch[1] : 0.0);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int tau = 16 * tid + q;
      double vv;
      if (tau == 0) vv = 1.0;
      else { run += y[q]; vv = (run != 0.0) ? y[q] * yin_div((double)tau, run) : 1.0; }
      y[q] = vv;
      yin[PAD16(tau)] = vv;
    }
    sync();

    // ---- first dip below the tolerance, else the last global minimum ---------------------------------
    int cand = 0x7fffffff;
#pragma unroll
    for (int q = 15; q >= 0; --q) {
      const int p = 16 * tid + q;
      const double nx1 = (q < 15) ? y[q + 1] : yin[PAD16(min(p + 1, YW - 1))];
      if (p >= 2 && p <= YW - 4 && y[q] < 0.75 && y[q] < nx1) cand = p;
    }
    // argmin with ties -> last index (mathutils.c:250-258), evaluated alongside: one exchange serves both
    double mv = y[0]; int mi = 16 * tid;
#pragma unroll
    for (int q = 1; q < 16; ++q) if (!(mv < y[q])) { mv = y[q]; mi = 16 * tid + q; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
      const double ov = __shfl_xor_sync(0xffffffffu, mv, o); const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (ov < mv || (ov == mv && oi > mi)) { mv = ov; mi = oi; }
    }
    if (lane == 0) { scratch[4 + wid] = mv; iscr[wid] = mi; iscr[4 + wid] = cand; }
    sync();
    if (tid == 0) {
      int pos;
      cand = min(iscr[4], iscr[5]);
      if (cand != 0x7fffffff) pos = cand;
      else {
        mv = scratch[4]; mi = iscr[0];
        { const double ov = scratch[5]; const int oi = iscr[1]; if (ov < mv || (ov == mv && oi > mi)) { mv = ov; mi = oi; } }
        pos = mi;
      }
      double period;
      if (pos == 0 || pos == YW - 1) period = (double)pos;          // mathutils.c:494-506
      else { const double s0 = yin[PAD16(pos - 1)], s1 = yin[PAD16(pos)], s2 = yin[PAD16(pos + 1)]; period = pos + .5 * (s0 - s2) / (s0 - 2. * s1 + s2); }
      unsigned peak_pos = 0;
      if (period == period && period >= 0.0 && period < (double)YW) peak_pos = (unsigned)period;
      double pitch = (period > 0.0) ? (double)P.sr / (period + 0.) : 0.0;                  // pitch.c:450-462
      const double lvl_hop = TA, lvl_frame = TA + IB[PAD16(YW - 1)];                       // the hop is the first block
      const bool silent_frame = (lvl_frame / (double)YN) < AFX_SILENCE_LEVEL;            // pitch.c:399-407
      if (silent_frame) pitch = 0.0;
      double conf = (1.0 - yin[PAD16(peak_pos)]) / 0.25;                                   // SA.cpp:887-889
      conf = conf < 0.0 ? 0.0 : (conf > 1.0 ? 1.0 : conf);
      double fsafe = 0.0;                                                                  // SA.cpp:897-916
      if (pitch > 0.0 && conf > 0.2) fsafe = pitch;
      else {
        const bool silent_hop = (lvl_hop / (double)YW) < AFX_SILENCE_LEVEL;
        if (!silent_hop) { const double c = B.cent_full[slot]; fsafe = (double)P.sr / (double)P.N * (c > 0.0 ? c : 0.0); }
      }
      B.fs[(size_t)FS_F0 * (size_t)B.TF + slot] = pitch;
      B.fs[(size_t)FS_F0_CONF * (size_t)B.TF + slot] = conf;
      B.fs[(size_t)FS_F0_FAILSAFE * (size_t)B.TF + slot] = fsafe;
    }
    sync();                                          // thread 0 is done with yin / I / scratch before the next frame
    pp ^= 1;                                         // this frame's second block is the carried block of the next
    cur = nxt; rel = nrel;
  }
}

template <int NG>
static void launch_pitch_hop_t(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s)
{
  const int smem = (int)PitchHopSmem<NG>::bytes;
  cudaFuncSetAttribute(k_pitch_hop<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int chunks = (B.g_slots + HCH - 1) / HCH;
  const int grid = std::max(1, std::min(sms, (chunks + NG - 1) / NG));
  unsigned int* ctr = P.t.work_ctr + 1;
  cudaMemsetAsync(ctr, 0, sizeof(unsigned int), s);
  k_pitch_hop<NG><<<grid, HT * NG, smem, s>>>(B, P, ctr);
}

template <int NG>
static void launch_pitch_t(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s)
{
  const int smem = (int)PitchSmem<NG>::bytes;
  // per launch: function attributes are per device, and one process may drive several devices
  cudaFuncSetAttribute(k_pitch<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int chunks = (B.g_slots + YCH - 1) / YCH;
  const int grid = std::max(1, std::min(sms, (chunks + NG - 1) / NG));
  unsigned int* ctr = P.t.work_ctr + 1;
  cudaMemsetAsync(ctr, 0, sizeof(unsigned int), s);
  k_pitch<NG><<<grid, YT * NG, smem, s>>>(B, P, ctr);
}

void afx_launch_pitch(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  // hop == 1024 (the reference's and BASELINE's hop): the block-sharing form; AFX_PITCH_GENERIC=1 keeps the general one
  static const int ng = [] { const char* e = getenv("AFX_PITCH_NG"); return e ? atoi(e) : 4; }();
  if (P.H == YW && P.N == YN && !B.pitch_generic) { if (ng == 4) launch_pitch_hop_t<4>(P, B, s); else launch_pitch_hop_t<5>(P, B, s); }
  else launch_pitch_t<3>(P, B, s);      // 3 groups: 168 registers per thread, no spills (4 groups at 128 registers measured 6 % slower)
  ++*launches;
}

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
  launch_pitch_t<3>(P, B, s);      // 3 groups: 168 registers per thread, no spills (4 groups at 128 registers measured 6 % slower)
  ++*launches;
}

// K1: PCM conditioning -- the tail of TSampleAnalyser::LoadSample
// (Source/Crawler/FeatureExtraction/Source/SampleAnalyser.cpp:531-719) for a whole batch.
//
//   k_downmix   : interleaved raw PCM (8 / 16 / 24 / 32-bit integer, float32; either byte order) -> float32 mono: the
//                 decoders' sample conversion (SampleConverter.h:392-518), then (sum of channels) * (1/C) in float32
//                 exactly as SA.cpp:535-548, fused with the peak / sum-of-squares reduction
//                 (SA.cpp:612-631) for files that need no resampling
//   k_resample  : libresample HQ restatement for files with src_rate != 44100 (SA.cpp:563-607)
//   k_reduce    : peak / sum-of-squares for resampled files
//   k_amp       : Amplification, FinalScaling (SA.cpp:630-637)
//   k_trim      : first / last sample above the -48 dB floor (SA.cpp:648-669)
//   k_layout    : lead / trail / padding / frame counts (SA.cpp:681-701, 760-764, 814, 991)
//   k_eff       : effective length at -48 / -24 / -12 dB (SA.cpp:1715-1756)
//   k_header    : file properties and scalar outputs (SA.cpp:734-754)
//
// All passes stream each sample once, coalesced; they are HBM-bound (bytes in DESIGN.md).
#include "afx_common.cuh"
#include "../../include/afec_b200.h"

#define CHUNK 8192          // samples per CTA
#define CT 256

// one sample of a raw interleaved PCM stream as the float32 in 16-bit range the reference's decoders produce
// (CoreFileFormats/Export/SampleConverter.h:392-518); idx counts samples (frame * channels + channel)
__device__ __forceinline__ float pcm_sample(const unsigned char* __restrict__ base, long long idx, int format)
{
  switch (format) {
    case AFX_PCM_I16: return (float)reinterpret_cast<const short*>(base)[idx];
    case AFX_PCM_F32: return reinterpret_cast<const float*>(base)[idx];
    case AFX_PCM_U8: return (float)(((int)base[idx] - 128) << 8);
    case AFX_PCM_I8: return (float)((int)(signed char)base[idx] << 8);
    case AFX_PCM_I16BE: { const unsigned char* p = base + 2 * idx; return (float)(short)((p[0] << 8) | p[1]); }
    case AFX_PCM_I24: case AFX_PCM_I24BE: {
      const unsigned char* p = base + 3 * idx;
      const unsigned u = (format == AFX_PCM_I24) ? ((unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16))
                                                 : ((unsigned)p[2] | ((unsigned)p[1] << 8) | ((unsigned)p[0] << 16));
      return (float)((double)(int)(u << 8) * 32768.0 / 2147483648.0);
    }
    case AFX_PCM_I32: case AFX_PCM_I32BE: {
      unsigned u = reinterpret_cast<const unsigned*>(base)[idx];
      if (format == AFX_PCM_I32BE) u = __byte_perm(u, 0, 0x0123);
      const float v = (float)((double)(int)u * 32768.0 / 2147483648.0);
      return fmaxf(-32768.0f, fminf(32767.0f, v));
    }
    default: {   // AFX_PCM_F32U / AFX_PCM_F32UBE
      unsigned u = reinterpret_cast<const unsigned*>(base)[idx];
      if (format == AFX_PCM_F32UBE) u = __byte_perm(u, 0, 0x0123);
      const double d = (double)__uint_as_float(u) * 32768.0;
      return (float)(d < -32768.0 ? -32768.0 : (d > 32767.0 ? 32767.0 : d));
    }
  }
}

__device__ __forceinline__ float load_mono(const unsigned char* __restrict__ pcm, const AfxFile& f, int i)
{
  // SA.cpp:535-548: dst = ch0; dst += ch[c] (c = 1..C-1); dst *= 1.0f / C    -- all float32
  const int C = f.channels;
  const unsigned char* base = pcm + f.pcm_off;
  const long long i0 = (long long)i * C;
  float acc = pcm_sample(base, i0, f.format);
  for (int c = 1; c < C; ++c) acc = __fadd_rn(acc, pcm_sample(base, i0 + c, f.format));
  if (C > 1) acc = __fmul_rn(acc, __fdiv_rn(1.0f, (float)C));
  return acc;
}

__device__ __forceinline__ void reduce_peak_sumsq(float amax, double ssq, AfxState* st, double* scratch)
{
  double v[1] = { ssq };
  block_sum<1>(v, scratch);
  const double m = block_max((double)amax, scratch);
  if (threadIdx.x == 0) {
    atomicMax(&st->maxabs_bits, __float_as_uint((float)m));
    atomicAdd(&st->sumsq, v[0]);
  }
}

__global__ void __launch_bounds__(CT) k_downmix(AfxBatchDev B, const int* __restrict__ chunk_file,
                                               const int* __restrict__ chunk_start, int analysis_rate)
{
  __shared__ double scratch[64];
  const int fi = chunk_file[blockIdx.x];
  const AfxFile f = B.files[fi];
  if (f.status != 0) return;
  const int start = chunk_start[blockIdx.x];
  const int end = min(start + CHUNK, f.src_end);
  const bool resampled = (f.src_rate != analysis_rate);
  float* dst = resampled ? (B.mono_src + f.src_off) : (B.mono + f.mono_off);
  float amax = 0.0f;
  double ssq = 0.0;
  if (f.channels == 1 && f.format == AFX_PCM_I16 && (f.pcm_off & 1) == 0) {
    // the common case (mono 16-bit): 4 samples per thread from aligned 8-byte loads, 16-byte stores (mono offsets are multiples
    // of 4).  Files packed back to back in one upload start at any even byte: the group is then cut out of TWO aligned words
    // (the neighbour thread reads the second one as its first: one DRAM pass either way; the words past a file's end belong
    // to the next file or to the slack afx_batch_upload leaves behind the buffer).  The per-sample path below took 2.7 x as
    // long on the bench's packed corpus.
    const unsigned mis = (unsigned)(f.pcm_off & 7);
    const unsigned long long* __restrict__ src8 = reinterpret_cast<const unsigned long long*>(B.pcm + (f.pcm_off - mis));
    const unsigned sh = 8u * mis;
    float4* __restrict__ dst4 = reinterpret_cast<float4*>(dst);
    const int g_end = end >> 2;
    for (int g = (start >> 2) + threadIdx.x; g < g_end; g += CT) {
      unsigned long long w = __ldg(src8 + g);
      if (mis) { const unsigned long long hi = __ldg(src8 + g + 1); w = (w >> sh) | (hi << (64u - sh)); }
      const float4 v = make_float4((float)(short)(w & 0xffffull), (float)(short)((w >> 16) & 0xffffull), (float)(short)((w >> 32) & 0xffffull),
                                   (float)(short)(w >> 48));
      dst4[g] = v;
      amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      const double q0 = (double)(v.x / 32768.0f), q1 = (double)(v.y / 32768.0f), q2 = (double)(v.z / 32768.0f), q3 = (double)(v.w / 32768.0f);
      ssq += q0 * q0; ssq += q1 * q1; ssq += q2 * q2; ssq += q3 * q3;
    }
    for (int i = (g_end << 2) + threadIdx.x; i < end; i += CT) {     // tail of the last chunk
      const float v = load_mono(B.pcm, f, i);
      dst[i] = v;
      amax = fmaxf(amax, fabsf(v));
      const double q = (double)(v / 32768.0f);
      ssq += q * q;
    }
  } else if (f.channels == 2 && f.format == AFX_PCM_I16 && (f.pcm_off & 1) == 0) {
    // stereo 16-bit, the usual sample-library file: 4 frames (16 bytes) per thread the same way; (l + r) * (1 / 2) in float32
    // is what load_mono computes (SA.cpp:535-548)
    const unsigned mis = (unsigned)(f.pcm_off & 7);
    const unsigned long long* __restrict__ src8 = reinterpret_cast<const unsigned long long*>(B.pcm + (f.pcm_off - mis));
    const unsigned sh = 8u * mis;
    float4* __restrict__ dst4 = reinterpret_cast<float4*>(dst);
    const float half = __fdiv_rn(1.0f, 2.0f);
    const int g_end = end >> 2;
    for (int g = (start >> 2) + threadIdx.x; g < g_end; g += CT) {
      unsigned long long w0 = __ldg(src8 + 2 * g), w1 = __ldg(src8 + 2 * g + 1);
      if (mis) { const unsigned long long w2 = __ldg(src8 + 2 * g + 2); w0 = (w0 >> sh) | (w1 << (64u - sh)); w1 = (w1 >> sh) | (w2 << (64u - sh)); }
      float4 v;
      v.x = __fmul_rn(__fadd_rn((float)(short)(w0 & 0xffffull), (float)(short)((w0 >> 16) & 0xffffull)), half);
      v.y = __fmul_rn(__fadd_rn((float)(short)((w0 >> 32) & 0xffffull), (float)(short)(w0 >> 48)), half);
      v.z = __fmul_rn(__fadd_rn((float)(short)(w1 & 0xffffull), (float)(short)((w1 >> 16) & 0xffffull)), half);
      v.w = __fmul_rn(__fadd_rn((float)(short)((w1 >> 32) & 0xffffull), (float)(short)(w1 >> 48)), half);
      dst4[g] = v;
      amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      const double q0 = (double)(v.x / 32768.0f), q1 = (double)(v.y / 32768.0f), q2 = (double)(v.z / 32768.0f), q3 = (double)(v.w / 32768.0f);
      ssq += q0 * q0; ssq += q1 * q1; ssq += q2 * q2; ssq += q3 * q3;
    }
    for (int i = (g_end << 2) + threadIdx.x; i < end; i += CT) {     // tail of the last chunk
      const float v = load_mono(B.pcm, f, i);
      dst[i] = v;
      amax = fmaxf(amax, fabsf(v));
      const double q = (double)(v / 32768.0f);
      ssq += q * q;
    }
  } else {
    for (int i = start + threadIdx.x; i < end; i += CT) {
      const float v = load_mono(B.pcm, f, i);
      dst[i] = v;
      amax = fmaxf(amax, fabsf(v));
      const double q = (double)(v / 32768.0f);
      ssq += q * q;
    }
  }
  if (!resampled) reduce_peak_sumsq(amax, ssq, B.state + fi, scratch);
}

// peak / sum of squares over the analysis-rate signal of resampled files (chunks over f.n)
__global__ void __launch_bounds__(CT) k_reduce(AfxBatchDev B, const int* __restrict__ chunk_file,
                                              const int* __restrict__ chunk_start)
{
  __shared__ double scratch[64];
  const int fi = chunk_file[blockIdx.x];
  const AfxFile f = B.files[fi];
  if (f.status != 0) return;
  const int start = chunk_start[blockIdx.x];
  const int end = min(start + CHUNK, f.dst_end);
  const float* src = B.mono + f.mono_off;
  float amax = 0.0f;
  double ssq = 0.0;
  for (int i = start + threadIdx.x; i < end; i += CT) {
    const float v = src[i];
    amax = fmaxf(amax, fabsf(v));
    const double q = (double)(v / 32768.0f);
    ssq += q * q;
  }
  reduce_peak_sumsq(amax, ssq, B.state + fi, scratch);
}

// ---- libresample HQ restatement (3rdParty/Resample/Dist/src) -------------------------------------
// One CTA per libresample block (<= 4096 source samples, resample.c:230-300), one thread per output sample.
// The output time stamps (a double accumulator advanced by repeated "+= 1/factor", resamplesubs.c:97-119) are
// data independent; the host replays that recurrence once per distinct (rate, length) pair and uploads, per
// block, the first output index, the input offset and every 64th time stamp.  A thread reproduces its own
// stamp from the nearest checkpoint by the same repeated additions (<= 63).
//
// Coefficients.  Each output sums ~2 x 17 x max(1, speed) products with coefficients picked from the 69632-entry
// Kaiser-sinc wing at indices (int)(frac(t) * dh + j * dh) that the reference forms by repeated double additions.
// Done literally that is ~15 instructions per tap (double add, convert, bounds checks, two scattered loads) and
// the kernel is issue bound.  For a rational rate ratio p / q the fractional parts of the stamps repeat every q
// outputs, so the first q outputs of a block build, in shared memory, one row of coefficients per residue
// k mod q, together with the row's stamp fraction and its MARGIN: the smallest distance of any of its running
// indices ho_j to an integer.  A later output with the same residue has a fraction that differs by ~1e-13
// (accumulated rounding of the stamps).  Its index sequence is then provably the row's when
//     |frac' - frac| * dh + 2e-9  <  margin
// (the two ho sequences start |frac' - frac| * dh (+ 1 ulp) apart and each of the <= 80 additions adds at most
// half an ulp(2^17) = 7.3e-12 of divergence), and the output is a plain dot product of the row with the staged
// source span: two shared loads, one multiply, one add per tap, in the reference's order.  Otherwise -- about
// once per 10^9 taps, and for rows with a tiny margin -- the output walks the exact index sequence against the
// table.  Results are bit-identical to the literal evaluation.  Rates whose rows do not fit in shared memory
// (q * taps too large, e.g. 44101 Hz) take the literal path.
#define RS_NWING (4096 * 34 / 2)
#define RS_THREADS 512

struct RsRow { double lph; double margin; int nl, nr; int h0l, h0r; };   // margin < 0: never trust the row

// literal evaluation (filterkit.c:115-215) against the staged source span xs[] (indexed like the block's X[])
__device__ __forceinline__ float rs_output_exact(double t, double factor, double dh, float lpscl, const float* __restrict__ imp,
                                                 const float* __restrict__ xs)
{
  const double fl = floor(t);
  const double lph = t - fl, rph = 1.0 - lph;
  const int xi = (int)fl;
  float vl = 0.0f, vr = 0.0f;
  if (factor >= 1) {
    { double ph = lph * 4096.0; int h = (int)ph; int x = xi;
      while (h < RS_NWING) { vl = __fadd_rn(vl, __fmul_rn(__ldg(imp + h), xs[x])); h += 4096; x -= 1; } }
    { double ph = rph * 4096.0; int h = (int)ph; int x = xi + 1; const int end = RS_NWING - 1;
      if (ph == 0) h += 4096;
      while (h < end) { vr = __fadd_rn(vr, __fmul_rn(__ldg(imp + h), xs[x])); h += 4096; x += 1; } }
  } else {
    { double ho = lph * dh; int x = xi;
      while ((int)ho < RS_NWING) { vl = __fadd_rn(vl, __fmul_rn(__ldg(imp + (int)ho), xs[x])); ho += dh; x -= 1; } }
    { double ho = rph * dh; int x = xi + 1; const int end = RS_NWING - 1;
      if (rph == 0) ho += dh;
      while ((int)ho < end) { vr = __fadd_rn(vr, __fmul_rn(__ldg(imp + (int)ho), xs[x])); ho += dh; x += 1; } }
  }
  return __fmul_rn(__fadd_rn(vl, vr), lpscl);
}

__device__ __forceinline__ double rs_int_dist(double v) { const double f = v - floor(v); return fmin(f, 1.0 - f); }

// coefficient row of one stamp: the literal index sequences, their tap counts and the margin
__device__ __forceinline__ void rs_build_row(double t, double factor, double dh, const float* __restrict__ imp,
                                             float* __restrict__ cl, float* __restrict__ cr, int tpw, RsRow* meta)
{
  const double fl = floor(t);
  const double lph = t - fl, rph = 1.0 - lph;
  RsRow m; m.lph = lph; m.margin = 1.0; m.nl = 0; m.nr = 0; m.h0l = 0; m.h0r = 0;
  if (factor >= 1) {
    // integer index steps: the row applies to every stamp with the same two starting indices (checked exactly)
    { double ph = lph * 4096.0; int h = (int)ph; m.h0l = h;
      while (h < RS_NWING && m.nl < tpw) { cl[m.nl++] = __ldg(imp + h); h += 4096; }
      if (h < RS_NWING) m.margin = -1.0; }
    { double ph = rph * 4096.0; int h = (int)ph; const int end = RS_NWING - 1; m.h0r = h;
      if (ph == 0) h += 4096;
      while (h < end && m.nr < tpw) { cr[m.nr++] = __ldg(imp + h); h += 4096; }
      if (h < end) m.margin = -1.0; }
  } else {
    { double ho = lph * dh;
      while ((int)ho < RS_NWING && m.nl < tpw) { m.margin = fmin(m.margin, rs_int_dist(ho)); cl[m.nl++] = __ldg(imp + (int)ho); ho += dh; }
      if ((int)ho < RS_NWING) m.margin = -1.0; else m.margin = fmin(m.margin, rs_int_dist(ho)); }
    { double ho = rph * dh; const int end = RS_NWING - 1;
      if (rph == 0) ho += dh;
      while ((int)ho < end && m.nr < tpw) { m.margin = fmin(m.margin, rs_int_dist(ho)); cr[m.nr++] = __ldg(imp + (int)ho); ho += dh; }
      if ((int)ho < end) m.margin = -1.0; else m.margin = fmin(m.margin, rs_int_dist(ho)); }
  }
  *meta = m;
}

__global__ void __launch_bounds__(RS_THREADS) k_resample(AfxBatchDev B, AfxTables T, const RsBlock* __restrict__ blocks,
                                                         const int* __restrict__ blk_file, const double* __restrict__ times,
                                                         int n_blocks, int analysis_rate, int smem_bytes)
{
  extern __shared__ __align__(16) unsigned char rs_smem[];
  const int bi = blockIdx.x;
  if (bi >= n_blocks) return;
  const RsBlock rb = blocks[bi];
  const AfxFile f = B.files[blk_file[bi]];
  const double speed = (double)f.src_rate / (double)analysis_rate;   // SA.cpp:563
  const double factor = 1.0 / speed;                                  // SA.cpp:584-590
  const double dt = 1.0 / factor;                                     // resamplesubs.c:44, 90
  const float* __restrict__ src = B.mono_src + f.src_off;
  float* __restrict__ dst = B.mono + f.mono_off;
  const float* __restrict__ imp = T.rs_imp;
  const int nsrc = f.nframes_src;
  float lpscl = 1.0f;
  if (factor < 1) lpscl = (float)((double)lpscl * factor);
  double dh = factor * 4096.0; if (dh > 4096.0) dh = 4096.0;
  const double* chk = times + rb.chk_off;

  // rate ratio p / q in lowest terms: the stamps' fractional parts repeat every q outputs
  int ga = f.src_rate, gb = analysis_rate;
  while (gb) { const int r = ga % gb; ga = gb; gb = r; }
  const int q = analysis_rate / ga;
  const int tpw = (int)(17.0 * (speed > 1.0 ? speed : 1.0)) + 3;     // taps per wing, upper bound
  const int rl = (2 * tpw) | 1;                                       // floats per row (odd: rows start on distinct banks)
  const int span = rb.span;
  const size_t o_meta = ((size_t)span * 4 + 15) & ~(size_t)15, o_rows = o_meta + (size_t)q * sizeof(RsRow);
  float* xs = reinterpret_cast<float*>(rs_smem);                      // [span] source samples X[0 .. span)
  RsRow* meta = reinterpret_cast<RsRow*>(rs_smem + o_meta);           // [q]
  float* rows = reinterpret_cast<float*>(rs_smem + o_rows);           // [q][rl]: left wing at 0, right wing at tpw
  const bool cached = o_rows + (size_t)q * rl * 4 <= (size_t)smem_bytes;
  const bool staged = (size_t)span * 4 <= (size_t)smem_bytes;

  if (staged) {
    for (int i = threadIdx.x; i < span; i += RS_THREADS) { const long long x = rb.in0 + i; xs[i] = (x >= 0 && x < nsrc) ? src[x] : 0.0f; }
  }
  if (cached) {
    for (int k = threadIdx.x; k < q && k < rb.nout; k += RS_THREADS) {
      double t = chk[k >> 6];
      for (int a = 0; a < (k & 63); ++a) t = __dadd_rn(t, dt);
      rs_build_row(t, factor, dh, imp, rows + (size_t)k * rl, rows + (size_t)k * rl + tpw, tpw, meta + k);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < rb.nout; k += RS_THREADS) {
    const int o = rb.out0 + k;
    if (o >= f.n) break;
    double t = chk[k >> 6];
    for (int a = 0; a < (k & 63); ++a) t = __dadd_rn(t, dt);         // the block's own repeated additions
    float v;
    bool done = false;
    if (cached) {
      const int r = k % q;
      const RsRow m = meta[r];
      const double fl = floor(t), lph = t - fl;
      bool same;
      if (factor >= 1) {
        const double rph = 1.0 - lph;
        same = m.margin >= 0.0 && (int)(lph * 4096.0) == m.h0l && (int)(rph * 4096.0) == m.h0r && ((rph * 4096.0 == 0) == ((1.0 - m.lph) * 4096.0 == 0));
      } else {
        same = (fabs(lph - m.lph) * dh + 2e-9 < m.margin) && ((lph == 0.0) == (m.lph == 0.0));
      }
      if (same) {
        const float* __restrict__ cl = rows + (size_t)r * rl;
        const float* __restrict__ cr = cl + tpw;
        const int xi = (int)fl;
        float vl = 0.0f, vr = 0.0f;
        const float* __restrict__ xl = xs + xi;          // left wing walks down from X[xi], right wing up from X[xi + 1]
        const float* __restrict__ xr = xs + xi + 1;
        int j = 0;
        for (; j + 4 <= m.nl; j += 4) {
          vl = __fadd_rn(vl, __fmul_rn(cl[j], xl[-j]));         vl = __fadd_rn(vl, __fmul_rn(cl[j + 1], xl[-j - 1]));
          vl = __fadd_rn(vl, __fmul_rn(cl[j + 2], xl[-j - 2])); vl = __fadd_rn(vl, __fmul_rn(cl[j + 3], xl[-j - 3]));
        }
        for (; j < m.nl; ++j) vl = __fadd_rn(vl, __fmul_rn(cl[j], xl[-j]));
        for (j = 0; j + 4 <= m.nr; j += 4) {
          vr = __fadd_rn(vr, __fmul_rn(cr[j], xr[j]));         vr = __fadd_rn(vr, __fmul_rn(cr[j + 1], xr[j + 1]));
          vr = __fadd_rn(vr, __fmul_rn(cr[j + 2], xr[j + 2])); vr = __fadd_rn(vr, __fmul_rn(cr[j + 3], xr[j + 3]));
        }
        for (; j < m.nr; ++j) vr = __fadd_rn(vr, __fmul_rn(cr[j], xr[j]));
        v = __fmul_rn(__fadd_rn(vl, vr), lpscl);
        done = true;
      }
    }
    if (!done) {
      if (staged) v = rs_output_exact(t, factor, dh, lpscl, imp, xs);
      else {
        // span larger than shared memory (never for the libresample block sizes): direct loads with bounds checks
        const double fl = floor(t); const double lph = t - fl, rph = 1.0 - lph; const long long xi = rb.in0 + (long long)fl;
        float vl = 0.0f, vr = 0.0f;
        if (factor >= 1) {
          { double ph = lph * 4096.0; int h = (int)ph; long long x = xi;
            while (h < RS_NWING) { const float sv = (x >= 0 && x < nsrc) ? src[x] : 0.0f; vl = __fadd_rn(vl, __fmul_rn(imp[h], sv)); h += 4096; x -= 1; } }
          { double ph = rph * 4096.0; int h = (int)ph; long long x = xi + 1; const int end = RS_NWING - 1;
            if (ph == 0) h += 4096;
            while (h < end) { const float sv = (x >= 0 && x < nsrc) ? src[x] : 0.0f; vr = __fadd_rn(vr, __fmul_rn(imp[h], sv)); h += 4096; x += 1; } }
        } else {
          { double ho = lph * dh; long long x = xi;
            while ((int)ho < RS_NWING) { const float sv = (x >= 0 && x < nsrc) ? src[x] : 0.0f; vl = __fadd_rn(vl, __fmul_rn(imp[(int)ho], sv)); ho += dh; x -= 1; } }
          { double ho = rph * dh; long long x = xi + 1; const int end = RS_NWING - 1;
            if (rph == 0) ho += dh;
            while ((int)ho < end) { const float sv = (x >= 0 && x < nsrc) ? src[x] : 0.0f; vr = __fadd_rn(vr, __fmul_rn(imp[(int)ho], sv)); ho += dh; x += 1; } }
        }
        v = __fmul_rn(__fadd_rn(vl, vr), lpscl);
      }
    }
    dst[o] = v;
  }
}

// ---- Amplification / FinalScaling -------------------------------------------------------------
// smallest non-negative float t with |(double)t * scale| > floor: the double-precision test of the reference is monotone in
// |t|, so the scans over the samples can compare float magnitudes against this instead of converting every sample
__device__ float first_float_above(double scale, double floor_)
{
  if (!(scale > 0.0)) return __int_as_float(0x7f800000);
  float t = (float)(floor_ / scale);
  if (!(t >= 0.0f)) t = 0.0f;
  for (int it = 0; it < 64 && t > 0.0f; ++it) {                       // down while the value below still passes
    const float below = __int_as_float(__float_as_int(t) - 1);
    if (fabs((double)below * scale) > floor_) t = below; else break;
  }
  for (int it = 0; it < 64 && !(fabs((double)t * scale) > floor_); ++it) t = __int_as_float(__float_as_int(t) + 1);   // up until it passes
  return t;
}

__global__ void k_amp(AfxBatchDev B, AfxParams P)
{
  const int fi = blockIdx.x * blockDim.x + threadIdx.x;
  if (fi >= B.n_files) return;
  AfxState* st = B.state + fi;
  const double maxamp = (double)__uint_as_float(st->maxabs_bits);
  const double amp = (maxamp > (double)1e-12f) ? 32768.0 / maxamp : 1.0;   // SA.cpp:636-637
  st->amp = amp;
  st->fs = amp / 32768.0;                                                  // SA.cpp:712
  st->thr_trim = first_float_above(amp, P.silence_floor_amp);              // fabs(amp * x) > floor (SA.cpp:648-669)
  for (int k = 0; k < 3; ++k) st->thr_eff[k] = first_float_above(st->fs, P.eff_floor[k]);   // fabs(x * fs) > floor (SA.cpp:1715-1756)
}

// EFF: the thresholds of the effective-length scan (SA.cpp:1715-1756) are tested in the same pass over the samples (one
// read of the mono signal instead of two); indices stay raw here, k_layout maps them
template <bool EFF>
__global__ void __launch_bounds__(CT) k_trim(AfxBatchDev B, const int* __restrict__ chunk_file,
                                            const int* __restrict__ chunk_start, double floor_amp, AfxParams P)
{
  __shared__ int scratch[32];
  const int fi = chunk_file[blockIdx.x];
  const AfxFile f = B.files[fi];
  if (f.status != 0) return;
  const int start = chunk_start[blockIdx.x];
  if (start >= f.n) return;
  const int end = min(start + CHUNK, f.dst_end);
  const float* src = B.mono + f.mono_off;
  const float tt = B.state[fi].thr_trim;
  const float t0 = B.state[fi].thr_eff[0], t1 = B.state[fi].thr_eff[1], t2 = B.state[fi].thr_eff[2];
  // a thread visits samples start + tid + it * CT, it = 0..31, in ascending order: one bit per visit and threshold, first and
  // last hit from the masks afterwards (two instructions per sample and threshold; the pass stays bound by the one read of
  // the signal)
  static_assert(CHUNK / CT == 32, "one mask bit per sample of a thread");
  // 16-byte loads: a thread visits the 4-sample groups tid + it * CT, it = 0..7 (mono offsets and chunk starts are multiples
  // of 4; the last group of a file may reach into the padding behind it -- those lanes are masked by i < end)
  unsigned mt = 0u, me[3] = { 0u, 0u, 0u };
  const float4* __restrict__ src4 = reinterpret_cast<const float4*>(src + start);
  const int nrem = end - start;
  const bool aligned = (reinterpret_cast<size_t>(src4) & 15) == 0;
#pragma unroll
  for (int it = 0; it < CHUNK / CT / 4; ++it) {
    const int g = threadIdx.x + it * CT;
    if (4 * g < nrem) {
      float4 v;
      if (aligned) v = src4[g];
      else {                                                     // a part's range (afx_part.cu) may start anywhere
        const float* q = src + start + 4 * g;
        v.x = q[0]; v.y = (4 * g + 1 < nrem) ? q[1] : 0.0f; v.z = (4 * g + 2 < nrem) ? q[2] : 0.0f; v.w = (4 * g + 3 < nrem) ? q[3] : 0.0f;
      }
      const float a4[4] = { fabsf(v.x), fabsf(v.y), fabsf(v.z), fabsf(v.w) };
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = (4 * g + j < nrem) ? a4[j] : -1.0f;      // (thresholds are >= 0: -1 passes none)
        const unsigned bit = 1u << (4 * it + j);
        if (a >= tt) mt |= bit;
        if (EFF) {
          if (a >= t0) me[0] |= bit;
          if (a >= t1) me[1] |= bit;
          if (a >= t2) me[2] |= bit;
        }
      }
    }
  }
  // bit b of a mask <-> sample start + 4 (tid + (b >> 2) CT) + (b & 3): ascending in b
  auto pos = [&](int b) { return start + 4 * ((int)threadIdx.x + (b >> 2) * CT) + (b & 3); };
  int first = mt ? pos(__ffs(mt) - 1) : 0x7fffffff, last = mt ? pos(31 - __clz(mt)) : -1;
  int ef[3], el[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { ef[k] = me[k] ? pos(__ffs(me[k]) - 1) : 0x7fffffff; el[k] = me[k] ? pos(31 - __clz(me[k])) : -1; }
  first = block_min_i(first, scratch);
  last = -block_min_i(-last, scratch);
  if (threadIdx.x == 0) {
    if (first != 0x7fffffff) atomicMin(&B.state[fi].first, first);
    if (last >= 0) atomicMax(&B.state[fi].last, last);
  }
  if (EFF) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int a = block_min_i(ef[k], scratch);
      const int b = -block_min_i(-el[k], scratch);
      if (threadIdx.x == 0) {
        if (a != 0x7fffffff) atomicMin(&B.state[fi].effraw_first[k], a);
        if (b >= 0) atomicMax(&B.state[fi].effraw_last[k], b);
      }
    }
  }
}

__global__ void k_layout(AfxBatchDev B, AfxParams P, int fused_eff)
{
  const int fi = blockIdx.x * blockDim.x + threadIdx.x;
  if (fi >= B.n_files) return;
  const AfxFile f = B.files[fi];
  AfxState* st = B.state + fi;
  if (f.status != 0) { st->F = 0; st->Fr = 0; st->len = 0; st->audible = 0; return; }
  const int n = f.n;
  // SA.cpp:651-669: lead = first sample above the floor (n if none); the trailing scan stops at lead
  const int first = st->first < n ? st->first : n;
  const int lead = first;
  int trail = 0;
  if (lead < n) { const int last = st->last > lead ? st->last : lead; trail = n - 1 - last; }
  const int audible = n - lead - trail;
  int end_off = 0, start_off = 0;                                          // SA.cpp:685-696
  if ((audible % P.N) < P.N / 2) end_off += P.N / 2;
  if (audible + end_off < P.N) start_off = P.N - audible - end_off;
  st->lead = lead; st->audible = audible; st->start_off = start_off;
  st->len = audible + start_off + end_off;
  st->data_offset = -lead + start_off;                                     // SA.cpp:701
  const int L = st->len < P.analysis_cap ? st->len : P.analysis_cap;       // SA.cpp:760-764
  st->L = L;
  int F = (L >= P.N) ? (L - P.N) / P.H + 1 : 0;
  int Fr = (L >= AFX_RFFT) ? (L - AFX_RFFT) / AFX_RHOP + 1 : 0;
  if (F > f.frame_cap) F = f.frame_cap;      // cannot happen (caps are upper bounds); keeps writes in range
  if (Fr > f.rframe_cap) Fr = f.rframe_cap;
  st->F = F; st->Fr = Fr;
  for (int k = 0; k < 3; ++k) { st->eff_first[k] = 0x7fffffff; st->eff_last[k] = -1; }
  if (fused_eff) {
    // a sample above an effective-length floor is above the trim floor too -- except on a rounding boundary of the two
    // expressions (the -48 dB floors are the same level); then, and only then, the audible span is scanned again
    bool inside = true;
    for (int k = 0; k < 3; ++k)
      if (st->effraw_last[k] >= 0 && (st->effraw_first[k] < lead || st->effraw_last[k] >= lead + audible)) inside = false;
    if (inside) {
      for (int k = 0; k < 3; ++k)
        if (st->effraw_last[k] >= 0) { st->eff_first[k] = st->effraw_first[k] - lead + start_off; st->eff_last[k] = st->effraw_last[k] - lead + start_off; }
    } else st->eff_rescan = 1;
  }
  if (B.inject && f.inject >= 0) {
    const AfxInject in = B.inject[f.inject];
    for (int k = 0; k < 3; ++k) { st->eff_first[k] = in.eff_first[k]; st->eff_last[k] = in.eff_last[k]; }
  }
}

__global__ void __launch_bounds__(CT) k_eff(AfxBatchDev B, const int* __restrict__ chunk_file,
                                           const int* __restrict__ chunk_start, AfxParams P)
{
  __shared__ int scratch[32];
  const int fi = chunk_file[blockIdx.x];
  const AfxFile f = B.files[fi];
  if (f.status != 0) return;
  const AfxState st = B.state[fi];
  // chunk over the mono index space; only the audible region [lead, lead + audible) maps into mData
  int start = chunk_start[blockIdx.x];
  int end = min(start + CHUNK, f.dst_end);
  start = max(start, st.lead); end = min(end, st.lead + st.audible);
  if (start >= end) return;      // uniform per block
  const float* src = B.mono + f.mono_off;
  int first[3] = { 0x7fffffff, 0x7fffffff, 0x7fffffff }, last[3] = { -1, -1, -1 };
  for (int i = start + threadIdx.x; i < end; i += CT) {
    const double v = fabs((double)src[i] * st.fs);
    const int idx = i - st.lead + st.start_off;
#pragma unroll
    for (int k = 0; k < 3; ++k) if (v > P.eff_floor[k]) { first[k] = min(first[k], idx); last[k] = max(last[k], idx); }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int a = block_min_i(first[k], scratch);
    const int b = -block_min_i(-last[k], scratch);
    if (threadIdx.x == 0) {
      if (a != 0x7fffffff) atomicMin(&B.state[fi].eff_first[k], a);
      if (b >= 0) atomicMax(&B.state[fi].eff_last[k], b);
    }
  }
}

// TAudioMath::SamplesToMs, AudioTypes/Export/AudioMath.inl:134-137 (float32 math)
__device__ __forceinline__ float samples_to_ms(int sr, int samples)
{
  return __fdiv_rn((float)samples, __fdiv_rn((float)sr, 1000.0f));
}

// the exact rescan k_layout may ask for (see there): one CTA per file, a no-op for all but pathological files
__global__ void __launch_bounds__(CT) k_eff_fix(AfxBatchDev B, AfxParams P)
{
  __shared__ int scratch[32];
  const int fi = blockIdx.x;
  const AfxFile f = B.files[fi];
  const AfxState st = B.state[fi];
  if (f.status != 0 || !st.eff_rescan || (B.inject && f.inject >= 0)) return;
  const float* src = B.mono + f.mono_off;
  int first[3] = { 0x7fffffff, 0x7fffffff, 0x7fffffff }, last[3] = { -1, -1, -1 };
  for (int i = st.lead + threadIdx.x; i < st.lead + st.audible; i += CT) {
    const double v = fabs((double)src[i] * st.fs);
    const int idx = i - st.lead + st.start_off;
#pragma unroll
    for (int k = 0; k < 3; ++k) if (v > P.eff_floor[k]) { first[k] = min(first[k], idx); last[k] = max(last[k], idx); }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int a = block_min_i(first[k], scratch);
    const int b = -block_min_i(-last[k], scratch);
    if (threadIdx.x == 0) { B.state[fi].eff_first[k] = a; B.state[fi].eff_last[k] = b; }
  }
}

__global__ void k_header(AfxBatchDev B, AfxParams P)
{
  const int fi = blockIdx.x * blockDim.x + threadIdx.x;
  if (fi >= B.n_files) return;
  const AfxFile f = B.files[fi];
  const AfxState st = B.state[fi];
  double* H = B.header + (size_t)fi * AFX_N_HEADER;
  for (int k = 0; k < AFX_N_HEADER; ++k) H[k] = 0.0;
  if (f.status != 0) return;
  H[H_FILE_SIZE] = (double)f.file_size;
  H[H_FILE_LENGTH] = (double)samples_to_ms(f.src_rate, f.nframes_src) / 1000.0;   // SA.cpp:741-742
  H[H_FILE_RATE] = f.src_rate; H[H_FILE_CHANNELS] = f.channels; H[H_FILE_BITS] = f.bit_depth;
  H[H_ANALYZATION_OFFSET] = (double)samples_to_ms(P.sr, st.data_offset) / 1000.0; // SA.cpp:748-749
  for (int k = 0; k < 3; ++k) {                                                   // SA.cpp:1731-1754
    const int len = st.len;
    const int first = st.eff_first[k] < len ? st.eff_first[k] : len;
    int trail = 0;
    if (first < len) { const int last = st.eff_last[k] > first ? st.eff_last[k] : first; trail = len - 1 - last; }
    H[H_EFF48 + k] = (double)samples_to_ms(P.sr, len - first - trail) / 1000.0;
  }
  const double maxamp = (double)__uint_as_float(st.maxabs_bits);
  H[H_PEAK] = (double)(float)fmin(1.0, maxamp / 32768.0);                         // SA.cpp:634
  H[H_RMS] = (double)(float)fmin(1.0, sqrt(st.sumsq / (double)f.n));              // SA.cpp:618-619
  H[H_DATA_OFFSET] = st.data_offset; H[H_DATA_LEN] = st.len;
}

// slot -> file maps: every frame-level kernel starts from its slot index, one coalesced table look-up replaces a
// 14-step dependent binary search over the file table
__global__ void __launch_bounds__(256) k_slotmap(AfxBatchDev B, int* __restrict__ slot_file, int* __restrict__ rslot_file)
{
  const int fi = blockIdx.x;
  const AfxFile f = B.files[fi];
  for (int i = threadIdx.x; i < f.frame_cap; i += 256) slot_file[f.frame_off + i] = fi;
  for (int i = threadIdx.x; i < f.rframe_cap; i += 256) rslot_file[f.rframe_off + i] = fi;
}

__global__ void k_state_init(AfxBatchDev B)
{
  const int fi = blockIdx.x * blockDim.x + threadIdx.x;
  if (fi >= B.n_files) return;
  AfxState* st = B.state + fi;
  st->maxabs_bits = 0u; st->sumsq = 0.0; st->first = 0x7fffffff; st->last = -1;
  for (int k = 0; k < 3; ++k) { st->effraw_first[k] = 0x7fffffff; st->effraw_last[k] = -1; }
  st->eff_rescan = 0;
  const int inj = B.files[fi].inject;
  if (B.inject && inj >= 0) {
    const AfxInject in = B.inject[inj];
    st->maxabs_bits = in.maxabs_bits; st->sumsq = in.sumsq; st->first = in.first; st->last = in.last;
  }
}

// debug / parity: materialise mData of one file
__global__ void k_materialise(const float* mono, const AfxState* st, double* out, int len)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) out[i] = mdata(mono, *st, i);
}
void afx_launch_materialise(const float* mono, const AfxState* st, double* out, int len, cudaStream_t s)
{
  k_materialise<<<(len + 255) / 256, 256, 0, s>>>(mono, st, out, len);
}

static void launch_resample(const AfxParams& P, const AfxBatchDev& B, const AfxCondPlan& C, cudaStream_t s)
{
  int smem = C.rs_smem_bytes > 0 ? C.rs_smem_bytes : 64 * 1024;
  if (smem > 200 * 1024) smem = 200 * 1024;
  cudaFuncSetAttribute(k_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   // per device, see afx_pitch.cu
  k_resample<<<C.n_rs_blocks, RS_THREADS, smem, s>>>(B, P.t, C.rs_blocks, C.rs_blk_file, C.rs_times, C.n_rs_blocks, P.sr, smem);
}

// long files conditioned in parts (afx_part.cu): the same passes, phase by phase, over one part's ranges
void afx_launch_part_reduce(const AfxParams& P, const AfxBatchDev& B, const AfxCondPlan& C, cudaStream_t s)
{
  k_state_init<<<1, 128, 0, s>>>(B);
  if (C.n_src_chunks > 0) k_downmix<<<C.n_src_chunks, CT, 0, s>>>(B, C.src_chunk_file, C.src_chunk_start, P.sr);
  if (C.n_rs_blocks > 0) launch_resample(P, B, C, s);
  if (C.n_rs_chunks > 0) k_reduce<<<C.n_rs_chunks, CT, 0, s>>>(B, C.rs_chunk_file, C.rs_chunk_start);
}
void afx_launch_part_trim(const AfxParams& P, const AfxBatchDev& B, const AfxCondPlan& C, cudaStream_t s)
{
  k_amp<<<1, 128, 0, s>>>(B, P);
  if (C.n_dst_chunks > 0) k_trim<false><<<C.n_dst_chunks, CT, 0, s>>>(B, C.dst_chunk_file, C.dst_chunk_start, P.silence_floor_amp, P);
}
void afx_launch_part_eff(const AfxParams& P, const AfxBatchDev& B, const AfxCondPlan& C, cudaStream_t s)
{
  k_layout<<<1, 128, 0, s>>>(B, P, 0);
  if (C.n_dst_chunks > 0) k_eff<<<C.n_dst_chunks, CT, 0, s>>>(B, C.dst_chunk_file, C.dst_chunk_start, P);
}

void afx_launch_condition_plan(const AfxParams& P, const AfxBatchDev& B, const AfxCondPlan& C, cudaStream_t s, long long* launches)
{
  if (B.n_files <= 0) return;
  const int fb = (B.n_files + 127) / 128;
  k_state_init<<<fb, 128, 0, s>>>(B); ++*launches;
  k_slotmap<<<B.n_files, 256, 0, s>>>(B, const_cast<int*>(B.slot_file), const_cast<int*>(B.rslot_file)); ++*launches;
  if (C.n_src_chunks > 0) { k_downmix<<<C.n_src_chunks, CT, 0, s>>>(B, C.src_chunk_file, C.src_chunk_start, P.sr); ++*launches; }
  if (C.n_rs_blocks > 0) {
    launch_resample(P, B, C, s); ++*launches;
    k_reduce<<<C.n_rs_chunks, CT, 0, s>>>(B, C.rs_chunk_file, C.rs_chunk_start); ++*launches;
  }
  k_amp<<<fb, 128, 0, s>>>(B, P); ++*launches;
  if (C.n_dst_chunks > 0) { k_trim<true><<<C.n_dst_chunks, CT, 0, s>>>(B, C.dst_chunk_file, C.dst_chunk_start, P.silence_floor_amp, P); ++*launches; }
  k_layout<<<fb, 128, 0, s>>>(B, P, 1); ++*launches;
  k_eff_fix<<<B.n_files, CT, 0, s>>>(B, P); ++*launches;
  k_header<<<fb, 128, 0, s>>>(B, P); ++*launches;
}

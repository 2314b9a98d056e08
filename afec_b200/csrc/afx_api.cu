// C ABI of libafec_b200.so (include/afec_b200.h): context, per-batch planning, staging and the
// kernel schedule.  Host code here only builds data-independent tables and plans (window, mel
// filters, twiddles, chunk / frame-slot tables); every sample-dependent operation runs on the GPU.
#include "afx_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#define CHUNK 8192

static thread_local std::string g_create_error;

static int fail(afx_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess);
int afx_fail(afx_ctx* c, int code, const char* what, cudaError_t e) { return fail(c, code, what, e); }
static int fail(afx_ctx* c, int code, const char* what, cudaError_t e)
{
  char buf[512];
  if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  else snprintf(buf, sizeof(buf), "%s", what);
  if (c) c->error = buf; else g_create_error = buf;
  return code;
}
#define CK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, AFX_ERR_CUDA, what, e_); } while (0)

// reference rounding helpers (CoreTypes/Export/InlineMath.inl:758-761, 823-826)
static int d2i_round(double v) { return (int)(v + (std::signbit(v) ? -0.5 : 0.5)); }
int afx_reference_round(double v) { return d2i_round(v); }
static int f2i_round(float v) { return (int)(v + (std::signbit(v) ? -0.5f : 0.5f)); }
static int ms_to_samples(int sr, float ms) { return f2i_round((float)sr / 1000.0f * ms); }
static double db_to_lin(double v) { if (v == 0.0) return 1.0; if (v > -200.0) return std::exp(v * (std::log(10.0) / 20.0)); return 0.0; }

// LibXtract init.c:237-378, equal gain, called as (N/2 bins, nyquist = sr/2, 20..15500 Hz, 14 filters)
static void build_mel(std::vector<double>& tab, int nbins, double nyquist, double fmin, double fmax, int nf)
{
  tab.assign((size_t)nf * nbins, 0.0);
  const int M = nbins >> 1;
  std::vector<double> mel_peak(nf + 2), lin_peak(nf + 2); std::vector<int> fft_peak(nf + 2);
  const double mel_max = 1127 * std::log(1 + fmax / 700), mel_min = 1127 * std::log(1 + fmin / 700);
  const double bw = (mel_max - mel_min) / nf;
  mel_peak[0] = mel_min; lin_peak[0] = fmin; fft_peak[0] = (int)(lin_peak[0] / nyquist * M);
  for (int n = 1; n < nf + 2; ++n) {
    mel_peak[n] = mel_peak[n - 1] + bw;
    lin_peak[n] = 700 * (std::exp(mel_peak[n] / 1127) - 1);
    fft_peak[n] = (int)(lin_peak[n] / nyquist * M);
  }
  int i = 0;
  for (int n = 0; n < nf; ++n) {
    double* row = tab.data() + (size_t)n * nbins;
    double inc = (n == 0) ? 1.0 / fft_peak[n] : 1.0 / (fft_peak[n] - fft_peak[n - 1]);
    double val = 0;
    for (; i <= fft_peak[n]; ++i) { row[i] = val; val += inc; }
    inc = 1.0 / (fft_peak[n + 1] - fft_peak[n]);
    val = 0;
    for (i = fft_peak[n + 1]; i > fft_peak[n]; --i) { row[i] = val; val += inc; }
  }
}

// extension tables (afx_ext.cu): 40 equal-gain triangular mel filters over ALL `nbins` bins (LibXtract's construction,
// init.c:237-378, without its halving of the bin range) and the 12 chroma classes
static void build_ext_weights(std::vector<float>& w, int nbins, double sr, int nfft)
{
  const int nmel = 40, nchr = 12, nout = nmel + nchr;
  w.assign((size_t)nout * nbins, 0.0f);
  const double nyquist = sr / 2.0, fmin = 20.0, fmax = 15500.0;
  std::vector<double> lin(nmel + 2); std::vector<int> peak(nmel + 2);
  const double mel_max = 1127 * std::log(1 + fmax / 700), mel_min = 1127 * std::log(1 + fmin / 700), bw = (mel_max - mel_min) / nmel;
  for (int n = 0; n < nmel + 2; ++n) {
    const double mel = mel_min + bw * n;
    lin[n] = (n == 0) ? fmin : 700 * (std::exp(mel / 1127) - 1);
    peak[n] = (int)(lin[n] / nyquist * nbins);
  }
  for (int n = 0; n < nmel; ++n) {              // filter n: rises peak[n] .. peak[n + 1], falls to peak[n + 2]
    float* row = w.data() + (size_t)n * nbins;
    const int p0 = peak[n], p1 = peak[n + 1], p2 = peak[n + 2];
    for (int k = p0; k <= p1 && k < nbins; ++k) row[k] = (p1 > p0) ? (float)((double)(k - p0) / (p1 - p0)) : 1.0f;
    for (int k = p1 + 1; k <= p2 && k < nbins; ++k) row[k] = (p2 > p1) ? (float)((double)(p2 - k) / (p2 - p1)) : 0.0f;
  }
  for (int k = 1; k < nbins; ++k) {             // chroma: a bin is shared linearly between its two nearest semitones
    const double f = (double)k * sr / nfft;
    if (f < 65.40639132514966 || f > 8372.018089619156) continue;
    const double pitch = 69.0 + 12.0 * std::log2(f / 440.0), lower = std::floor(pitch), frac = pitch - lower;
    const int c0 = (((int)lower % 12) + 12) % 12, c1 = (c0 + 1) % 12;
    w[(size_t)(nmel + c0) * nbins + k] += (float)(1.0 - frac);
    w[(size_t)(nmel + c1) * nbins + k] += (float)frac;
  }
}

// libresample filterkit.c:67-113 + resample.c:113-116 (float copy of one Kaiser-windowed sinc wing)
static double rs_izero(double x)
{
  double sum = 1, u = 1, halfx = x / 2.0; int n = 1;
  do { double t = halfx / (double)n; n += 1; t *= t; u *= t; sum += u; } while (u >= 1E-21 * sum);
  return sum;
}
static void build_rs_wing(std::vector<float>& imp)
{
  const int nwing = 4096 * 34 / 2;
  std::vector<double> c(nwing);
  const double frq = 0.5 * 0.90, beta = 6, pi = 3.14159265358979232846;
  c[0] = 2.0 * frq;
  for (int i = 1; i < nwing; ++i) { const double t = pi * (double)i / 4096.0; c[i] = std::sin(2.0 * t * frq) / t; }
  const double ibeta = 1.0 / rs_izero(beta), inm1 = 1.0 / ((double)(nwing - 1));
  for (int i = 1; i < nwing; ++i) {
    const double t = (double)i * inm1; double t1 = 1.0 - t * t; t1 = (t1 < 0 ? 0 : t1);
    c[i] *= rs_izero(beta * std::sqrt(t1)) * ibeta;
  }
  imp.resize(nwing);
  for (int i = 0; i < nwing; ++i) imp[i] = (float)c[i];
}

// resample.c:80-164 (open, high quality) + :170-337 (process, lastFlag = 1) + resamplesubs.c:30-123: the
// block structure and the time accumulator do not depend on the samples, so they are replayed here once per
// (source rate, length) and the filter sums run on the GPU (k_resample).
//
// The accumulator `t += dt` (resamplesubs.c:97-119) is replayed without walking every output sample: while t stays
// inside one binade [2^e, 2^(e+1)) every double is a multiple of u = ulp(t), so fl(t + dt) = t + c with the
// CONSTANT c = dt rounded to the grid u (round to nearest; an exact tie falls back to stepping) -- t advances
// linearly and the k-th stamp is t + k c exactly.  Only the addition that crosses into the next binade (about 7
// per 4096-sample block) is made in hardware.  Integers in units of 2^-80 keep the bookkeeping exact.  The result
// is bit-identical to the step-by-step replay (rs_advance_stepwise; tests/test_longfile_host.py), at O(blocks)
// instead of O(samples): 86 k blocks instead of 159 M additions for an hour of 96 kHz audio.
typedef __int128 i128;
#define RS_UNIT 80
static inline i128 rs_to_fixed(double v) { return (i128)std::ldexp(v, RS_UNIT); }
static inline double rs_from_fixed(i128 v) { return std::ldexp((double)v, -RS_UNIT); }

// one block: stamps of outputs 0.. while t < end_time; pushes every 64th stamp; returns the count, t ends as the
// accumulator after the last addition
static int rs_advance_stepwise(double& t, double dt, double end_time, std::vector<double>& chk)
{
  int nout = 0;
  while (t < end_time) { if ((nout & 63) == 0) chk.push_back(t); ++nout; t += dt; }
  return nout;
}

static int rs_advance(double& t, double dt, double end_time, std::vector<double>& chk)
{
  if (!(dt >= 0x1p-20 && dt < 0x1p12 && t >= 1.0 && end_time < 0x1p40)) return rs_advance_stepwise(t, dt, end_time, chk);
  const i128 DT = rs_to_fixed(dt), E = rs_to_fixed(end_time);
  long long nout = 0;
  while (t < end_time) {
    int e; std::frexp(t, &e);                          // t in [2^(e-1), 2^e)
    const i128 u = (i128)1 << (RS_UNIT + e - 53);      // ulp(t)
    const i128 B = (i128)1 << (RS_UNIT + e);           // upper end of the binade
    const i128 T = rs_to_fixed(t);
    const i128 lo = DT & (u - 1), c = DT - lo + ((2 * lo > u) ? u : 0);   // u is a power of two
    long long k = 0;
    if (2 * lo != u && c > 0) {
      // outputs j = 0..k-1 at T + j c: each must satisfy the loop test (T + j c < E) and its addition must stay
      // inside the binade (T + j c + DT < B)
      const i128 r1 = E - T, r2 = B - DT - T;
      const i128 r = r1 < r2 ? r1 : r2;
      if (r > 0) {                                      // k = ceil(r / c): a double estimate, corrected exactly
        k = (long long)((double)r / (double)c);
        while ((i128)k * c < r) ++k;
        while (k > 0 && (i128)(k - 1) * c >= r) --k;
      }
    }
    if (k > 0) {
      const long long m0 = (nout + 63) & ~63LL;
      if (m0 < nout + k) {
        // stamps 64 outputs apart differ by 64 c: a multiple of u that keeps the sum inside the binade -> exact in double
        double v = rs_from_fixed(T + (i128)(m0 - nout) * c);
        const double step = rs_from_fixed(64 * c);
        for (long long m = m0; m < nout + k; m += 64) { chk.push_back(v); v += step; }
      }
      t = rs_from_fixed(T + (i128)k * c);
      nout += k;
    } else {
      if ((nout & 63) == 0) chk.push_back(t);          // the crossing (or tie) step, in hardware
      ++nout; t += dt;
    }
  }
  return (int)nout;
}

static std::shared_ptr<RsShape> rs_plan(int in_len, int src_rate, int sr, int out_len, bool stepwise = false)
{
  auto sh = std::make_shared<RsShape>();
  const double speed = (double)src_rate / (double)sr, factor = 1.0 / speed;
  const int nmult = 35;
  const double inv = 1.0 / factor;
  const unsigned xoff = (unsigned)(((nmult + 1) / 2.0) * (inv > 1.0 ? inv : 1.0) + 10);
  const unsigned xsize = (2 * xoff + 10 > 4096) ? 2 * xoff + 10 : 4096;
  unsigned xread = xoff;
  double time = (double)xoff;
  const double dt = 1.0 / factor;
  int used = 0, outc = 0;
  long long in0 = -(long long)xoff;
  sh->chk.reserve((size_t)out_len / 64 + (size_t)in_len / 2048 + 64);
  for (;;) {
    int len = (int)(xsize - xread);
    if (len >= in_len - used) len = in_len - used;
    used += len; xread += len;
    const int nx = (used == in_len) ? (int)(xread - xoff) : (int)(xread - 2 * xoff);
    if (nx <= 0) break;
    RsBlock rb; rb.out0 = outc; rb.in0 = in0; rb.chk_off = (long long)sh->chk.size();
    double t = time; const double end_time = t + nx;
    const int nout = stepwise ? rs_advance_stepwise(t, dt, end_time, sh->chk) : rs_advance(t, dt, end_time, sh->chk);
    time = t;
    time -= nx; unsigned xp = xoff + nx;
    const unsigned ncreep = (unsigned)((int)time - (int)xoff);
    if (ncreep) { time -= ncreep; xp += ncreep; }
    const unsigned shift = xp - xoff;
    const unsigned nreuse = xread - shift;
    int ncopy = out_len - outc; if (ncopy > nout) ncopy = nout;
    rb.nout = ncopy;
    rb.span = nx + 2 * (int)xoff + 2; rb.pad = 0;
    if (ncopy > 0) { sh->blocks.push_back(rb); sh->span.push_back(rb.span); sh->max_span = std::max(sh->max_span, rb.span); }
    outc += ncopy;
    in0 += shift; xread = nreuse;
    if (ncopy < nout) break;
  }
  sh->produced = outc;
  return sh;
}

// shared memory k_resample wants for one rate: the block's source span + one (index, coefficient) row per residue
int afx_rs_smem_need(int sr, int src_rate, int span)
{
  int a = src_rate, b = sr;
  while (b) { const int r = a % b; a = b; b = r; }
  const long long q = sr / a;
  const double speed = (double)src_rate / (double)sr;
  const int tpw = (int)(17.0 * (speed > 1.0 ? speed : 1.0)) + 3, rl = (2 * tpw) | 1;
  const long long need = (((long long)span * 4 + 15) & ~15LL) + q * 32 + q * rl * 4;   // span + RsRow[q] + float[q][rl]
  return need <= 200 * 1024 ? (int)need : (int)std::min<long long>(200 * 1024, (((long long)span * 4 + 15) & ~15LL));
}

// test hook: 0 when the closed-form replay equals the step-by-step one bit for bit (blocks, spans, every checkpoint)
extern "C" int afx_debug_rs_plan_check(int32_t sample_rate, int64_t nframes, int32_t src_rate)
{
  if (sample_rate <= 0 || src_rate <= 0 || nframes <= 0 || nframes > 0x7fffffffLL) return AFX_ERR_ARG;
  const double speed = (double)src_rate / (double)sample_rate;
  int n = d2i_round((double)(int)nframes / speed); if (n < 1) n = 1;
  auto a = rs_plan((int)nframes, src_rate, sample_rate, n, false), b = rs_plan((int)nframes, src_rate, sample_rate, n, true);
  if (a->produced != b->produced || a->blocks.size() != b->blocks.size() || a->chk.size() != b->chk.size() || a->span != b->span) return 1;
  for (size_t i = 0; i < a->blocks.size(); ++i)
    if (a->blocks[i].out0 != b->blocks[i].out0 || a->blocks[i].nout != b->blocks[i].nout || a->blocks[i].in0 != b->blocks[i].in0 ||
        a->blocks[i].chk_off != b->blocks[i].chk_off) return 2;
  if (!a->chk.empty() && memcmp(a->chk.data(), b->chk.data(), a->chk.size() * 8) != 0) return 3;
  return 0;
}

extern "C" int afx_abi_version(void) { return AFX_ABI_VERSION; }

extern "C" int afx_pcm_bytes(int32_t format)
{
  switch (format) {
    case AFX_PCM_U8: case AFX_PCM_I8: return 1;
    case AFX_PCM_I16: case AFX_PCM_I16BE: return 2;
    case AFX_PCM_I24: case AFX_PCM_I24BE: return 3;
    case AFX_PCM_F32: case AFX_PCM_I32: case AFX_PCM_I32BE: case AFX_PCM_F32U: case AFX_PCM_F32UBE: return 4;
    default: return 0;
  }
}

extern "C" const char* afx_last_error(const afx_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

extern "C" int afx_create(const afx_config* cfg, afx_ctx** out)
{
  afx_ctx* ctx = nullptr;
  if (!cfg || !out) return fail(nullptr, AFX_ERR_ARG, "afx_create: null argument");
  *out = nullptr;
  if (cfg->sample_rate != 44100 || cfg->fft_size != 2048)
    return fail(nullptr, AFX_ERR_ARG, "afx_create: only sample_rate 44100 / fft_size 2048 are supported (Crawler.cpp:41-43)");
  if (cfg->hop_size < 256 || cfg->hop_size > 2048 || (cfg->hop_size % 256) != 0)
    return fail(nullptr, AFX_ERR_ARG, "afx_create: hop_size must be a multiple of 256 in [256, 2048]");
  if ((cfg->features & AFX_FEAT_PACK) && (cfg->features & AFX_FEAT_ALL) != AFX_FEAT_ALL)
    return fail(nullptr, AFX_ERR_ARG, "afx_create: AFX_FEAT_PACK packs the rows of the full low-level set (AFX_FEAT_ALL | AFX_FEAT_PACK)");
  if ((cfg->features & AFX_FEAT_EXT_MELCHROMA) && !(cfg->features & AFX_FEAT_SPECTRAL))
    return fail(nullptr, AFX_ERR_ARG, "afx_create: AFX_FEAT_EXT_MELCHROMA contracts the magnitude spectra (needs AFX_FEAT_SPECTRAL)");
  if ((cfg->features & AFX_FEAT_HIGHLEVEL) && (cfg->features & AFX_FEAT_ALL) != AFX_FEAT_ALL)
    return fail(nullptr, AFX_ERR_ARG, "afx_create: AFX_FEAT_HIGHLEVEL derives from the full low-level set (AFX_FEAT_ALL | AFX_FEAT_HIGHLEVEL)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) return fail(nullptr, AFX_ERR_CUDA, "afx_create: no CUDA device (no CPU fallback exists)", e);
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, AFX_ERR_ARG, "afx_create: bad device ordinal");
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) return fail(nullptr, AFX_ERR_CUDA, "cudaSetDevice", e);

  ctx = new afx_ctx();
  ctx->cfg = *cfg; ctx->device = cfg->device;
  ctx->debug_times = getenv("AFX_DEBUG_KERNEL_TIMES") && atoi(getenv("AFX_DEBUG_KERNEL_TIMES")) != 0;
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete ctx; return fail(nullptr, AFX_ERR_CUDA, "cudaStreamCreate", e); }
  for (int i = 0; i < 3 && e == cudaSuccess; ++i) { e = cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking); if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming); }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_spec, cudaEventDisableTiming);
  if (e != cudaSuccess) { afx_destroy(ctx); return fail(nullptr, AFX_ERR_CUDA, "cudaStreamCreate(side)", e); }
  // MEASURED (round 2, full workload): the four kernel chains on four streams make a step 5 % SLOWER than one stream (606 vs
  // 577 ms): every frame kernel fills the GPU on its own, and side by side they only share caches and shared memory.  One
  // stream is the default; AFX_MULTI_STREAM=1 brings the fork / join back.
  ctx->multi_stream = !ctx->debug_times && getenv("AFX_MULTI_STREAM") && atoi(getenv("AFX_MULTI_STREAM")) != 0
                      && !(getenv("AFX_SINGLE_STREAM") && atoi(getenv("AFX_SINGLE_STREAM")) != 0);
  // ... and for the same reason the computes of the contexts of one device (the adapter's slots) run one after the other, in the
  // order they were enqueued, while their copies overlap the other contexts' kernels (AFX_COMPUTE_CHAIN=0: independent)
  ctx->compute_chain = !(getenv("AFX_COMPUTE_CHAIN") && atoi(getenv("AFX_COMPUTE_CHAIN")) == 0);

  AfxParams& P = ctx->P;
  memset(&P, 0, sizeof(P));
  const int sr = cfg->sample_rate, N = cfg->fft_size, H = cfg->hop_size;
  P.sr = sr; P.N = N; P.H = H;
  // SampleAnalyser.cpp:171-175 -- FrequenciesPerBin is an integer division (= 21)
  const double fpb = (double)(sr / N);
  P.first_bin = d2i_round(20.0 / fpb);
  const int last_bin = d2i_round(15500.0 / fpb);
  P.nbins = last_bin - P.first_bin + 1;
  if (P.first_bin != AFX_WIN_FIRST || P.nbins != AFX_WIN_BINS) { afx_destroy(ctx); return fail(nullptr, AFX_ERR_ARG, "afx_create: analysis window differs from the compiled-in one"); }
  static const double b14[14] = { 50.0, 100.0, 200.0, 400.0, 630.0, 920.0, 1270.0, 1720.0, 2320.0, 3150.0, 4400.0, 6400.0, 9500.0, 15500.0 };
  static const double b28[28] = { 50.0, 100.0, 150.0, 200.0, 300.0, 400.0, 510.0, 630.0, 770.0, 920.0, 1080.0, 1270.0, 1480.0, 1720.0,
    2000.0, 2320.0, 2700.0, 3150.0, 3700.0, 4400.0, 5300.0, 6400.0, 7700.0, 9500.0, 12000.0, 15500.0, 19000.0, 22050.0 };
  { // SampleAnalyser.cpp:2084-2100, 2134-2241: bins are consumed cumulatively from first_bin
    int cur = P.first_bin;
    for (int b = 0; b < 14; ++b) {
      const int s = (b == 0) ? P.first_bin : d2i_round(b14[b - 1] / fpb);
      const int en = d2i_round(b14[b] / fpb);
      int nb = en - s + 1; if (nb > N / 2 - cur) nb = N / 2 - cur;
      P.band14_start[b] = cur; P.band14_n[b] = nb;
      int nei = (int)(0.3 * nb); if (nei < 1) nei = 1;
      P.band14_nei[b] = nei;
      cur += nb;
    }
  }
  for (int b = 0; b < 28; ++b) {   // SampleAnalyser.cpp:2026-2045
    int s = d2i_round((b == 0) ? (double)P.first_bin : b28[b - 1] / fpb);
    int en = d2i_round(b28[b] / fpb); if (en > N / 2) en = N / 2;
    if (s >= N / 2) { s = 0; en = 0; }
    P.band28_s[b] = s; P.band28_e[b] = en;
  }
  P.wh_decay = std::pow(0.001, (double)((float)H / (float)sr) / 22.0);       // awhitening.c:84-86, SA.cpp:44
  P.env_coef = std::pow(0.01, (1000.0 / (8.0 * (double)sr)));                // Envelopes.cpp:66-69, SA.cpp:69
  P.silence_floor_amp = (double)32768.0f * db_to_lin(-48.0);                 // SA.cpp:648-649
  P.eff_floor[0] = db_to_lin(-48.0); P.eff_floor[1] = db_to_lin(-24.0); P.eff_floor[2] = db_to_lin(-12.0);
  P.analysis_cap = ms_to_samples(sr, 1000 * 20);                             // SA.cpp:37, 760-761
  P.ac_min_period = ms_to_samples(sr, 0.8f); P.ac_width = ms_to_samples(sr, 12.0f);   // SA.cpp:2318-2324

  // rhythm constants: OnsetDetector.cpp:106-111 (relax 25 s), :296, :318 (normalisation), CannyWindow.cpp:27-48
  P.r_relax = (float)(std::exp((-2.30258509 * (float)AFX_RHOP) / (25.0f * (float)sr)));
  P.r_norm_power = 2560.f / (float)((AFX_RBINS + 2) * AFX_RFFT);
  P.r_norm_complex = (float)(231.70475 / std::pow((double)AFX_RFFT, 1.5));
  for (int i = -12; i < 13; ++i) P.canny[i + 12] = (double)i / (16.0 * 16.0) * std::exp(-1.0 * (i * i) / (2.0 * 16.0 * 16.0));
  if (getenv("AFX_GROUP_FRAMES")) ctx->group_frames = std::max(1LL, atoll(getenv("AFX_GROUP_FRAMES")));
  if (getenv("AFX_GROUP_RFRAMES")) ctx->group_rframes = std::max(1LL, atoll(getenv("AFX_GROUP_RFRAMES")));
  if (getenv("AFX_PITCH_GENERIC")) ctx->pitch_generic = atoi(getenv("AFX_PITCH_GENERIC")) != 0;
  if (getenv("AFX_RHYTHM_PIPE")) ctx->rhythm_pipe = atoi(getenv("AFX_RHYTHM_PIPE")) != 0 ? 1 : 0;
  if (getenv("AFX_RHYTHM_FUSED")) ctx->rhythm_fused_min = atoi(getenv("AFX_RHYTHM_FUSED")) != 0 ? 0 : 0x7fffffff;

  // ---- constant tables ----
  std::vector<double> window(N), rwindow(AFX_RFFT), mel, dct(14 * 14);
  std::vector<double> tw2048(2 * 2048), tw512(2 * 512);
  std::vector<float> imp;
  const double pi = 3.14159265358979323846;
  for (int n = 0; n < N; ++n) window[n] = (0.5 * (1.0 - std::cos(2.0 * pi * (double)n / (double)(N - 1)))) * 2.0;   // window.c:67-76, SA.cpp:178-181
  for (int i = 0; i < AFX_RFFT; ++i) rwindow[i] = 0.5 * (0.5 * (1.0 - std::cos(6.2831853071795864769252867665590 * (double)i * (1.0 / (double)(AFX_RFFT - 1)))));  // Fourier.cpp:545-551, pre-halved (exact): k_rhythm_polar's pair unpack yields 2 X[k]
  build_mel(mel, N / 2, sr / 2, 20.0, 15500.0, 14);
  for (int q = 0; q < 14; ++q) {
    int lo = N / 2, hi = -1;
    for (int k = 0; k < N / 2; ++k) if (mel[(size_t)q * (N / 2) + k] != 0.0) { if (k < lo) lo = k; hi = k; }
    P.mel_lo[q] = lo; P.mel_hi[q] = hi;
  }
  for (int n = 0; n < 14; ++n) for (int m = 1; m <= 14; ++m) dct[n * 14 + (m - 1)] = std::cos(pi * (n / (double)14) * (m - 0.5));   // vector.c:372-391
  for (int k = 0; k < 2048; ++k) { tw2048[2 * k] = std::cos(-2.0 * pi * k / 2048.0); tw2048[2 * k + 1] = std::sin(-2.0 * pi * k / 2048.0); }
  for (int k = 0; k < 512; ++k) { tw512[2 * k] = std::cos(-2.0 * pi * k / 512.0); tw512[2 * k + 1] = std::sin(-2.0 * pi * k / 512.0); }
  build_rs_wing(imp);
  // FFT pass twiddles in access order (afx_fft16.cuh)
  std::vector<double> ft2(2 * 15 * 16), ft3a(2 * 3 * 256), ft3b(2 * 7 * 256);
  for (int r = 1; r < 16; ++r) for (int k = 0; k < 16; ++k) { const double a = -2.0 * pi * (double)(r * k) / 256.0; ft2[2 * ((r - 1) * 16 + k)] = std::cos(a); ft2[2 * ((r - 1) * 16 + k) + 1] = std::sin(a); }
  for (int r = 1; r < 4; ++r) for (int j = 0; j < 256; ++j) { const double a = -2.0 * pi * (double)(r * j) / 1024.0; ft3a[2 * ((r - 1) * 256 + j)] = std::cos(a); ft3a[2 * ((r - 1) * 256 + j) + 1] = std::sin(a); }
  for (int r = 1; r < 8; ++r) for (int j = 0; j < 256; ++j) { const double a = -2.0 * pi * (double)(r * j) / 2048.0; ft3b[2 * ((r - 1) * 256 + j)] = std::cos(a); ft3b[2 * ((r - 1) * 256 + j) + 1] = std::sin(a); }

  // float32 twiddles of the autocorrelation's transforms (afx_autocorr.cu): pass 2, pass 3 (N = 512), bin-pair unpack (N = 1024)
  std::vector<float> actw(2 * (240 + 256 + 257));
  for (int r = 1; r < 16; ++r) for (int k = 0; k < 16; ++k) { const double a = -2.0 * pi * (double)(r * k) / 256.0; actw[2 * ((r - 1) * 16 + k)] = (float)std::cos(a); actw[2 * ((r - 1) * 16 + k) + 1] = (float)std::sin(a); }
  for (int j = 0; j < 256; ++j) { const double a = -2.0 * pi * (double)j / 512.0; actw[2 * (240 + j)] = (float)std::cos(a); actw[2 * (240 + j) + 1] = (float)std::sin(a); }
  for (int k = 0; k <= 256; ++k) { const double a = -2.0 * pi * (double)k / 1024.0; actw[2 * (496 + k)] = (float)std::cos(a); actw[2 * (496 + k) + 1] = (float)std::sin(a); }

  size_t off = 0;
  auto place = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_actw = place(actw.size() * 4);
  const size_t o_win = place(window.size() * 8), o_rwin = place(rwindow.size() * 8), o_mel = place(mel.size() * 8),
    o_dct = place(dct.size() * 8), o_tw = place(tw2048.size() * 8), o_tw5 = place(tw512.size() * 8), o_imp = place(imp.size() * 4),
    o_ft2 = place(ft2.size() * 8), o_ft3a = place(ft3a.size() * 8), o_ft3b = place(ft3b.size() * 8), o_ctr = place(64 * 4),
    o_pad = place(32 * 8);
  // the walk of k_bands_lane: cut the bin axis at every sub-band / frequency-band / mel-support edge and every 16 bins (its tile)
  std::vector<AfxBandSeg> segs;
  {
    std::vector<int> cuts;
    for (int k = 0; k <= N / 2; k += 16) cuts.push_back(k);
    for (int b = 0; b < 14; ++b) { cuts.push_back(P.band14_start[b]); cuts.push_back(P.band14_start[b] + P.band14_n[b]); }
    for (int b = 0; b < 28; ++b) { cuts.push_back(P.band28_s[b]); cuts.push_back(P.band28_e[b]); }
    for (int q = 0; q < 14; ++q) if (P.mel_hi[q] >= P.mel_lo[q]) { cuts.push_back(P.mel_lo[q]); cuts.push_back(P.mel_hi[q] + 1); }
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    for (size_t i = 0; i + 1 < cuts.size(); ++i) {
      const int k0 = cuts[i], k1 = cuts[i + 1];
      if (k0 < 0 || k1 > N / 2 || k0 >= k1) continue;
      AfxBandSeg sg; memset(&sg, 0, sizeof(sg));
      sg.k0 = (short)k0; sg.k1 = (short)k1; sg.b14 = sg.b28 = sg.q0 = -1;
      for (int b = 0; b < 14; ++b) if (k0 >= P.band14_start[b] && k0 < P.band14_start[b] + P.band14_n[b]) {
        sg.b14 = (signed char)b; sg.start14 = (k0 == P.band14_start[b]); sg.end14 = (k1 == P.band14_start[b] + P.band14_n[b]);
      }
      for (int b = 0; b < 28; ++b) if (k0 >= P.band28_s[b] && k0 < P.band28_e[b]) { sg.b28 = (signed char)b; sg.end28 = (k1 == P.band28_e[b]); }
      for (int q = 0; q < 14; ++q) if (P.mel_hi[q] >= P.mel_lo[q] && k0 >= P.mel_lo[q] && k0 <= P.mel_hi[q]) { if (sg.q0 < 0) sg.q0 = (signed char)q; ++sg.nq; }
      if (sg.nq > 0 && k1 == P.mel_hi[sg.q0] + 1) sg.fin |= 1;
      if (sg.nq > 1 && k1 == P.mel_hi[sg.q0 + 1] + 1) sg.fin |= 2;
      if (sg.nq > 2 || (sg.nq == 2 && !(k0 >= P.mel_lo[sg.q0 + 1] && k0 <= P.mel_hi[sg.q0 + 1]))) {
        afx_destroy(ctx); return fail(nullptr, AFX_ERR_ARG, "afx_create: more than two mel filters overlap (unexpected filter table)");
      }
      segs.push_back(sg);
    }
  }
  std::vector<double> mel_ab(2 * (size_t)(N / 2), 0.0);
  for (const auto& sg : segs)
    for (int k = sg.k0; k < sg.k1; ++k) {
      if (sg.nq > 0) mel_ab[2 * k] = mel[(size_t)sg.q0 * (N / 2) + k];
      if (sg.nq > 1) mel_ab[2 * k + 1] = mel[(size_t)(sg.q0 + 1) * (N / 2) + k];
    }
  std::vector<float> extw, extw_k; std::vector<double> extdct(13 * 40);
  build_ext_weights(extw, N / 2, (double)sr, N);
  extw_k.resize(extw.size());
  for (int o = 0; o < 52; ++o) for (int k = 0; k < N / 2; ++k) extw_k[(size_t)k * 52 + o] = extw[(size_t)o * (N / 2) + k];
  for (int n = 0; n < 13; ++n) for (int m = 0; m < 40; ++m) extdct[n * 40 + m] = std::cos(pi * n * (m + 0.5) / 40.0);
  const size_t o_segs = place(segs.size() * sizeof(AfxBandSeg));
  const size_t o_melab = place(mel_ab.size() * 8);
  const size_t o_extw = place(extw.size() * 4), o_extwk = place(extw_k.size() * 4), o_extdct = place(extdct.size() * 8);
  e = ctx->tables.reserve(off);
  if (e != cudaSuccess) { afx_destroy(ctx); return fail(nullptr, AFX_ERR_CUDA, "cudaMalloc(tables)", e); }
  unsigned char* base = (unsigned char*)ctx->tables.p;
  cudaMemcpy(base + o_win, window.data(), window.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_rwin, rwindow.data(), rwindow.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_mel, mel.data(), mel.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_dct, dct.data(), dct.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_tw, tw2048.data(), tw2048.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_tw5, tw512.data(), tw512.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_ft2, ft2.data(), ft2.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_ft3a, ft3a.data(), ft3a.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_ft3b, ft3b.data(), ft3b.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_actw, actw.data(), actw.size() * 4, cudaMemcpyHostToDevice);
  P.t.ac_tw = (const float2*)(base + o_actw);
  e = cudaMemcpy(base + o_imp, imp.data(), imp.size() * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { afx_destroy(ctx); return fail(nullptr, AFX_ERR_CUDA, "cudaMemcpy(tables)", e); }
  P.t.window = (const double*)(base + o_win); P.t.rwindow = (const double*)(base + o_rwin);
  P.t.mel = (const double*)(base + o_mel); P.t.dct = (const double*)(base + o_dct);
  cudaMemcpy(base + o_melab, mel_ab.data(), mel_ab.size() * 8, cudaMemcpyHostToDevice);
  P.t.mel_ab = (const double2*)(base + o_melab);
  P.t.tw2048 = (const double2*)(base + o_tw); P.t.tw512 = (const double2*)(base + o_tw5);
  P.t.rs_imp = (const float*)(base + o_imp);
  P.t.work_ctr = (unsigned int*)(base + o_ctr);
  P.t.fft_t2 = (const double2*)(base + o_ft2); P.t.fft_t3_1024 = (const double2*)(base + o_ft3a); P.t.fft_t3_2048 = (const double2*)(base + o_ft3b);

  P.t.hl_pad = (double*)(base + o_pad);
  cudaMemcpy(base + o_segs, segs.data(), segs.size() * sizeof(AfxBandSeg), cudaMemcpyHostToDevice);
  P.t.band_segs = (const AfxBandSeg*)(base + o_segs); P.t.n_band_segs = (int)segs.size();
  cudaMemcpy(base + o_extw, extw.data(), extw.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_extwk, extw_k.data(), extw_k.size() * 4, cudaMemcpyHostToDevice);
  e = cudaMemcpy(base + o_extdct, extdct.data(), extdct.size() * 8, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { afx_destroy(ctx); return fail(nullptr, AFX_ERR_CUDA, "cudaMemcpy(ext tables)", e); }
  memset(&ctx->ext_tables, 0, sizeof(ctx->ext_tables));
  ctx->ext_tables.w_nmajor = (const float*)(base + o_extw); ctx->ext_tables.w_kmajor = (const float*)(base + o_extwk);
  ctx->ext_tables.dct = (const double*)(base + o_extdct);
  ctx->ext_tensor = getenv("AFX_EXT_TENSOR") && atoi(getenv("AFX_EXT_TENSOR")) != 0;
  ctx->max_frame_cap = (P.analysis_cap - AFX_RFFT) / AFX_RHOP + 2;
  ctx->zeros.assign(ctx->max_frame_cap, 0.0);
  if (cfg->features & AFX_FEAT_HIGHLEVEL) {
    // The classification features pad short files with the LAST frame of a silent 0.5-s sample
    // (SampleClassificationDescriptors.cpp:330-368 analyses one at start-up; so does this context, through its own kernels)
    std::vector<int16_t> silence(sr / 2, 0);
    afx_file f; memset(&f, 0, sizeof(f));
    f.pcm = silence.data(); f.nframes = (int64_t)silence.size(); f.channels = 1; f.src_rate = sr; f.format = AFX_PCM_I16; f.bit_depth = 16;
    afx_batch* b = nullptr;
    afx_file_result r;
    if (afx_analyze(ctx, &f, 1, &b) != AFX_OK || afx_batch_result(b, 0, &r) != AFX_OK || r.status != AFX_FILE_OK || r.n_frames < 1) {
      g_create_error = "afx_create: the silent reference sample failed to analyse: " + ctx->error;
      if (b) afx_batch_free(b);
      afx_destroy(ctx);
      return AFX_ERR_CUDA;
    }
    const int last = r.n_frames - 1;
    double pad[21];
    for (int k = 0; k < 14; ++k) pad[k] = r.fv[5][(size_t)last * 28 + k];
    static const int series[7] = { FS_SPEC_RMS, FS_SPEC_FLATNESS, FS_SPEC_FLUX, FS_SPEC_CONTRAST, FS_SPEC_COMPLEXITY, FS_F0_CONF, FS_AMP_RMS };
    for (int k = 0; k < 7; ++k) pad[14 + k] = r.fs[series[k]][last];
    afx_batch_free(b);
    e = cudaMemcpy(P.t.hl_pad, pad, sizeof(pad), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { afx_destroy(ctx); return fail(nullptr, AFX_ERR_CUDA, "cudaMemcpy(hl_pad)", e); }
    ctx->hl_pad_ready = true;
  }
  *out = ctx;
  return AFX_OK;
}

// end of the last compute enqueued on each device, by any context: the next compute on that device waits for it
static std::mutex g_chain_mu;
static cudaEvent_t g_chain[64] = { nullptr };

extern "C" void afx_destroy(afx_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  {
    std::lock_guard<std::mutex> lk(g_chain_mu);
    if (ctx->device >= 0 && ctx->device < 64 && g_chain[ctx->device] == ctx->ev_chain) g_chain[ctx->device] = nullptr;   // (its work is done: synchronised above)
  }
  if (ctx->ev_chain) cudaEventDestroy(ctx->ev_chain);
  for (int i = 0; i < 3; ++i) { if (ctx->side[i]) { cudaStreamSynchronize(ctx->side[i]); cudaStreamDestroy(ctx->side[i]); } if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]); }
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_spec) cudaEventDestroy(ctx->ev_spec);
  DevBuf* bufs[] = { &ctx->tables, &ctx->d_pcm, &ctx->d_mono, &ctx->d_mono_src, &ctx->d_files, &ctx->d_state, &ctx->d_mag, &ctx->d_cent,
    &ctx->d_fs, &ctx->d_fsr, &ctx->d_fv, &ctx->d_rpolar, &ctx->d_rodf, &ctx->d_rpost, &ctx->d_bandraw, &ctx->d_slotmap, &ctx->d_stats, &ctx->d_header, &ctx->d_plan, &ctx->d_scratch,
    &ctx->d_hl, &ctx->d_hl_pitch, &ctx->d_hl_sig, &ctx->d_hl_feat, &ctx->d_hl_status, &ctx->d_pack, &ctx->d_pack_off, &ctx->d_pack_file_off, &ctx->d_ext_mfcc, &ctx->d_ext_chroma, &ctx->d_ext_idx };
  for (DevBuf* b : bufs) b->release();
  ctx->h_results_cache.release(); ctx->h_plan_cache.release(); ctx->h_pack_cache.release();
  for (auto& pb : ctx->part_pool) pb.release();
  delete ctx;
}

extern "C" int afx_trim(afx_ctx* ctx)
{
  if (!ctx) return AFX_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamSynchronize(ctx->side[i]);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->copy_stream);
  if (e != cudaSuccess) return fail(ctx, AFX_ERR_CUDA, "afx_trim", e);
  if (ctx->live) return fail(ctx, AFX_ERR_STATE, "afx_trim: a batch of this context is still alive");
  DevBuf* bufs[] = { &ctx->d_pcm, &ctx->d_mono, &ctx->d_mono_src, &ctx->d_files, &ctx->d_state, &ctx->d_mag, &ctx->d_cent,
    &ctx->d_fs, &ctx->d_fsr, &ctx->d_fv, &ctx->d_rpolar, &ctx->d_rodf, &ctx->d_rpost, &ctx->d_bandraw, &ctx->d_slotmap, &ctx->d_stats, &ctx->d_header, &ctx->d_plan, &ctx->d_scratch,
    &ctx->d_hl, &ctx->d_hl_pitch, &ctx->d_hl_sig, &ctx->d_hl_feat, &ctx->d_hl_status, &ctx->d_pack, &ctx->d_pack_off, &ctx->d_pack_file_off, &ctx->d_ext_mfcc, &ctx->d_ext_chroma, &ctx->d_ext_idx };
  for (DevBuf* b : bufs) b->release();
  for (auto& pb : ctx->part_pool) pb.release();
  ctx->part_pool.clear();
  return AFX_OK;
}

extern "C" int afx_host_alloc(afx_ctx* ctx, uint64_t bytes, void** out)
{
  if (!ctx || !out) return AFX_ERR_ARG;
  cudaSetDevice(ctx->device);
  CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault), "cudaHostAlloc");
  return AFX_OK;
}
extern "C" int afx_host_free(afx_ctx* ctx, void* p)
{
  if (!ctx) return AFX_ERR_ARG;
  if (p) CK(cudaFreeHost(p), "cudaFreeHost");
  return AFX_OK;
}

// ---- batch planning ------------------------------------------------------------------------------
// process-wide cache of the data-independent libresample replays, keyed by (analysis rate, source rate, length)
std::shared_ptr<RsShape> afx_rs_shape(int sr, int in_len, int src_rate, int out_len)
{
  // least-recently-used eviction; a batch keeps the shapes it uses alive itself (afx_batch::rs_shapes), so an entry
  // dropped here mid-batch is only dropped from the cache
  struct Entry { std::shared_ptr<RsShape> shape; unsigned long long used; };
  static std::mutex mu;
  static std::map<std::tuple<int, int, int>, Entry> cache;
  static unsigned long long tick = 0;
  const auto key = std::make_tuple(sr, src_rate, in_len);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { it->second.used = ++tick; return it->second.shape; }
  }
  std::shared_ptr<RsShape> made = rs_plan(in_len, src_rate, sr, out_len);   // planned outside the lock: slot threads plan in parallel
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { it->second.used = ++tick; return it->second.shape; }
  if (cache.size() >= 512) {
    auto victim = cache.begin();
    for (auto c = cache.begin(); c != cache.end(); ++c) if (c->second.used < victim->second.used) victim = c;
    cache.erase(victim);
  }
  cache.emplace(key, Entry{ made, ++tick });
  return made;
}

extern "C" int afx_batch_create(afx_ctx* ctx, const afx_file* files, int32_t n_files, afx_batch** out)
{
  return afx_batch_create_impl(ctx, files, n_files, nullptr, out);
}

int afx_batch_create_impl(afx_ctx* ctx, const afx_file* files, int32_t n_files, const AfxCondInput* cond, afx_batch** out)
{
  if (!ctx || !out || (n_files > 0 && !files) || n_files < 0 || (cond && n_files != 1)) return fail(ctx, AFX_ERR_ARG, "afx_batch_create: bad arguments");
  *out = nullptr;
  const AfxParams& P = ctx->P;
  afx_batch* b = new afx_batch();
  b->ctx = ctx; b->n_files = n_files;
  b->in.assign(files, files + n_files);
  b->files.resize(n_files);
  size_t pcm_off = 0; long long mono_off = 0, src_off = 0; long long tf = 0, tfr = 0;
  const unsigned char* run_end = nullptr;
  std::map<std::pair<int, int>, long long> shape_pool;   // (source rate, source length) -> offset of that shape's checkpoints in rs_chk
  afx_batch::Group g = { 0, 0, 0, 0, 0, 0 };
  auto close_group = [&]() {
    if (g.nfiles > 0) {
      b->groups.push_back(g);
      b->max_gslots = std::max(b->max_gslots, g.nslots); b->max_grslots = std::max(b->max_grslots, g.nrslots);
    }
  };
  for (int i = 0; i < n_files; ++i) {
    const afx_file& f = files[i];
    AfxFile& d = b->files[i];
    memset(&d, 0, sizeof(d));
    d.channels = f.channels; d.src_rate = f.src_rate; d.format = f.format; d.bit_depth = f.bit_depth; d.file_size = f.file_size;
    d.status = AFX_FILE_OK;
    if (f.channels < 1 || f.channels > 8) d.status = AFX_FILE_BAD_CHANNELS;          // SA.cpp:472-477
    else if (f.nframes <= 0 || (!f.pcm && !cond)) d.status = AFX_FILE_EMPTY;          // SA.cpp:479-482
    // a description the kernels cannot take fails that file only (the reference fails files one by one, SA.cpp:372-408)
    else if (f.nframes * f.channels > 0x7fffffffLL || f.src_rate <= 0 || afx_pcm_bytes(f.format) == 0) d.status = AFX_FILE_UNSUPPORTED;
    d.frame_off = (int)tf; d.rframe_off = (int)tfr; d.mono_off = mono_off; d.src_off = src_off; d.pcm_off = (long long)pcm_off;
    if (d.status != AFX_FILE_OK) { if (g.nfiles > 0) ++g.nfiles; else { g.file0 = i; g.nfiles = 1; g.slot0 = (int)tf; g.rslot0 = (int)tfr; } continue; }
    d.nframes_src = (int)f.nframes;
    const double speed = (double)f.src_rate / (double)P.sr;                            // SA.cpp:563-573
    d.n = d.nframes_src;
    if (speed != 1.0) { int nn = d2i_round(d.nframes_src / speed); d.n = nn < 1 ? 1 : nn; }
    d.src_end = d.nframes_src; d.dst_end = d.n; d.inject = -1;
    // upper bounds for the conditioned length: len <= max(n + N/2, N)
    long long lmax = std::max<long long>((long long)d.n + P.N / 2, P.N);
    if (lmax > P.analysis_cap) lmax = P.analysis_cap;
    d.frame_cap = (int)((lmax - P.N) / P.H + 1);
    d.rframe_cap = (int)((lmax - AFX_RFFT) / AFX_RHOP + 1);
    if (g.nfiles > 0 && (g.nslots + d.frame_cap > ctx->group_frames || g.nrslots + d.rframe_cap > ctx->group_rframes)) {
      close_group();
      g.nfiles = 0;
    }
    if (g.nfiles == 0) { g.file0 = i; g.slot0 = (int)tf; g.rslot0 = (int)tfr; g.nslots = 0; g.nrslots = 0; }
    ++g.nfiles; g.nslots += d.frame_cap; g.nrslots += d.rframe_cap;
    b->max_fr = std::max(b->max_fr, d.rframe_cap);
    tf += d.frame_cap; tfr += d.rframe_cap;
    if (cond) {
      // conditioned elsewhere: the mono window goes straight into the mono buffer, addressed by its global sample index
      if ((!cond->mono && cond->mono_count > 0) || cond->mono_count < 0 || cond->mono_begin < 0 || cond->mono_begin + cond->mono_count > d.n) {
        delete b; return fail(ctx, AFX_ERR_ARG, "afx_analyze_conditioned: bad mono window");
      }
      if (cond->mono_count > 0) b->runs.push_back({ (const unsigned char*)cond->mono, (size_t)mono_off * 4, (size_t)cond->mono_count * 4, true });
      d.mono_off = mono_off - cond->mono_begin;
      mono_off += (cond->mono_count + 3) & ~3LL;
      d.inject = (int)b->inject.size();
      b->inject.push_back(cond->inj);
      continue;
    }
    // PCM packing: keep host-contiguous files contiguous on the device so they move in one copy
    const size_t bps = (size_t)afx_pcm_bytes(f.format);
    const size_t align = (bps == 3) ? 1 : bps;             // the kernels read 2- and 4-byte samples with aligned loads, 3-byte ones bytewise
    const size_t bytes = (size_t)f.nframes * f.channels * bps;
    const unsigned char* hp = (const unsigned char*)f.pcm;
    if (run_end == hp && (pcm_off % align) == 0 && !b->runs.empty()) {
      b->runs.back().bytes += bytes;
    } else {
      pcm_off = (pcm_off + 15) & ~(size_t)15;
      b->runs.push_back({ hp, pcm_off, bytes, false });
    }
    d.pcm_off = (long long)pcm_off;
    pcm_off += bytes; run_end = hp + bytes;
    mono_off += ((long long)d.n + 3) & ~3LL;
    d.mono_off = mono_off - (((long long)d.n + 3) & ~3LL);
    if (speed != 1.0) {
      d.src_off = src_off; src_off += ((long long)d.nframes_src + 3) & ~3LL;
      std::shared_ptr<RsShape> sh = afx_rs_shape(P.sr, d.nframes_src, f.src_rate, d.n);
      const auto skey = std::make_pair(f.src_rate, d.nframes_src);
      auto pit = shape_pool.find(skey);
      if (pit == shape_pool.end()) {
        pit = shape_pool.emplace(skey, (long long)b->rs_chk.size()).first;
        b->rs_chk.insert(b->rs_chk.end(), sh->chk.begin(), sh->chk.end());
        b->rs_shapes.push_back(sh);
      }
      for (RsBlock rb : sh->blocks) { rb.chk_off += pit->second; b->rs_blocks.push_back(rb); b->rs_blk_file.push_back(i); }
      if (!sh->blocks.empty()) b->rs_smem = std::max(b->rs_smem, afx_rs_smem_need(P.sr, f.src_rate, sh->max_span));
      if (sh->produced < d.n) b->rs_tails.push_back({ d.mono_off + sh->produced, (long long)d.n - sh->produced });
    }
    for (int s = 0; s < d.nframes_src; s += CHUNK) { b->src_chunk_file.push_back(i); b->src_chunk_start.push_back(s); }
    for (int s = 0; s < d.n; s += CHUNK) {
      b->dst_chunk_file.push_back(i); b->dst_chunk_start.push_back(s);
      if (speed != 1.0) { b->rs_chunk_file.push_back(i); b->rs_chunk_start.push_back(s); }
    }
    if (tf > 0x7fffffffLL || tfr > 0x7fffffffLL) { delete b; return fail(ctx, AFX_ERR_ARG, "afx_batch_create: batch too large (frame slots overflow int32)"); }
  }
  close_group();
  // per-file kernels (whitening recurrences, rhythm back end) take a group's files longest first: their run time grows
  // with the file's frame count (the beat tracker's with its square), and a long file started last is the launch's tail
  b->file_order.resize(n_files);
  for (int i = 0; i < n_files; ++i) b->file_order[i] = i;
  for (const auto& gg : b->groups)
    std::stable_sort(b->file_order.begin() + gg.file0, b->file_order.begin() + gg.file0 + gg.nfiles,
                     [&](int x, int y) { return b->files[x].rframe_cap > b->files[y].rframe_cap; });
  b->pcm_bytes = pcm_off; b->mono_samples = mono_off; b->mono_src_samples = src_off;
  b->TF = (int)tf; b->TFr = (int)tfr;

  // host result layout
  size_t o = 0;
  b->o_header = o; o += (size_t)n_files * AFX_N_HEADER;
  b->o_stats = o; o += (size_t)n_files * AFX_N_SERIES * AFX_N_STATS;
  b->o_fs = o; o += (size_t)AFX_N_FS_MAIN * b->TF;
  b->o_fsr = o; o += (size_t)2 * b->TFr;
  b->o_fv = o; o += (size_t)AFX_FV_STRIDE * b->TF;
  b->o_state = o; o += ((size_t)n_files * sizeof(AfxState) + 7) / 8;
  if (ctx->cfg.features & AFX_FEAT_EXT_MELCHROMA) {
    b->o_ext_mfcc = o; o += (size_t)b->TF * 13;
    b->o_ext_chroma = o; o += (size_t)b->TF * 12;
    b->o_ext_idx = o; o += (size_t)b->TF;
  }
  if (ctx->cfg.features & AFX_FEAT_HIGHLEVEL) {
    b->o_hl = o; o += (size_t)n_files * AFX_N_HL;
    b->o_hl_sig = o; o += (size_t)n_files * AFX_HL_SIGNATURE;
    b->o_hl_feat = o; o += (size_t)n_files * AFX_HL_FEATURES;
    b->o_hl_pitch = o; o += (size_t)b->TF;
    b->o_hl_status = o; o += ((size_t)n_files * sizeof(int) + 7) / 8;
  }
  b->total_doubles = o;
  if (ctx->cfg.features & AFX_FEAT_PACK) {
    b->pack_file_off.resize(n_files);
    size_t po = 0;
    for (int i = 0; i < n_files; ++i) {
      b->pack_file_off[i] = po;
      if (b->files[i].status == AFX_FILE_OK) po += (afx_pack_region_bytes(b->files[i].frame_cap, b->files[i].rframe_cap) + 15) & ~(size_t)15;
    }
    b->pack_bytes = po;
  }
  *out = b;
  return AFX_OK;
}

static void take_cached(PinBuf& dst, PinBuf& cache, size_t bytes)
{
  if (cache.p && cache.cap >= bytes) { dst = cache; cache.p = nullptr; cache.cap = 0; }
}
static void give_back(PinBuf& src, PinBuf& cache)
{
  if (!src.p) return;
  if (!cache.p || cache.cap < src.cap) { cache.release(); cache = src; src.p = nullptr; src.cap = 0; }
  else src.release();
}

extern "C" int afx_batch_upload(afx_batch* b)
{
  if (!b) return AFX_ERR_ARG;
  afx_ctx* ctx = b->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  // the device buffers belong to the context: a second batch uploaded while another one's results still live in them
  // would overwrite (or reallocate) what the first one downloads -- one uploaded batch per context at a time
  if (ctx->live && ctx->live != b)
    return fail(ctx, AFX_ERR_STATE, "afx_batch_upload: another batch of this context is still alive (afx_batch_free it first, or use one context per batch in flight)");
  ctx->live = b;
  cudaSetDevice(ctx->device);
  const int n = b->n_files;
  if (!b->ev[0]) for (int i = 0; i < 6; ++i) CK(cudaEventCreate(&b->ev[i]), "cudaEventCreate");

  // device buffers
  const size_t TF = (size_t)b->TF, TFr = (size_t)b->TFr;
  const unsigned feat = ctx->cfg.features;
  CK(ctx->d_pcm.reserve(b->pcm_bytes + 16), "cudaMalloc(pcm)");
  CK(ctx->d_mono.reserve((size_t)(b->mono_samples + 8) * 4), "cudaMalloc(mono)");
  if (b->mono_src_samples) CK(ctx->d_mono_src.reserve((size_t)(b->mono_src_samples + 8) * 4), "cudaMalloc(mono_src)");
  CK(ctx->d_files.reserve((size_t)(n + 1) * sizeof(AfxFile)), "cudaMalloc(files)");
  CK(ctx->d_state.reserve((size_t)(n + 1) * sizeof(AfxState)), "cudaMalloc(state)");
  const size_t GS = (size_t)b->max_gslots, GR = (size_t)b->max_grslots;
  CK(ctx->d_mag.reserve((GS + 1) * AFX_NBIN * 8), "cudaMalloc(mag)");
  CK(ctx->d_cent.reserve((TF + 1) * 8), "cudaMalloc(cent)");
  CK(ctx->d_slotmap.reserve((TF + TFr + 2) * 4), "cudaMalloc(slotmap)");
  CK(ctx->d_fs.reserve((TF + 1) * AFX_N_FS_MAIN * 8), "cudaMalloc(fs)");
  CK(ctx->d_fsr.reserve((TFr + 1) * 2 * 8), "cudaMalloc(fsr)");
  if (feat & AFX_FEAT_BANDS) {
    CK(ctx->d_fv.reserve((TF + 1) * AFX_FV_STRIDE * 8), "cudaMalloc(fv)");
    CK(ctx->d_bandraw.reserve((GS + 1) * 16 * 8), "cudaMalloc(bandraw)");   // afx_bands.cu BR2_STRIDE
  }
  if (feat & AFX_FEAT_RHYTHM) {
    size_t GRs = 0;                                  // polar rows only exist for the groups that take the split rhythm kernels
    for (const auto& g : b->groups) if (!ctx->rhythm_fused(g.nfiles)) GRs = std::max(GRs, (size_t)g.nrslots);
    if (GRs) CK(ctx->d_rpolar.reserve((GRs + 1) * AFX_RROW * 4), "cudaMalloc(rpolar)");
    CK(ctx->d_rodf.reserve((TFr + 1) * 2 * 4), "cudaMalloc(rodf)");
    CK(ctx->d_rpost.reserve((TFr + 1) * 2 * 4), "cudaMalloc(rpost)");
    CK(ctx->d_scratch.reserve((GR + 1) * 4 * 8 + 1024), "cudaMalloc(scratch)");
  }
  if (feat & AFX_FEAT_HIGHLEVEL) {
    CK(ctx->d_hl.reserve((size_t)(n + 1) * AFX_N_HL * 8), "cudaMalloc(hl)");
    CK(ctx->d_hl_sig.reserve((size_t)(n + 1) * AFX_HL_SIGNATURE * 8), "cudaMalloc(hl signature)");
    CK(ctx->d_hl_feat.reserve((size_t)(n + 1) * AFX_HL_FEATURES * 8), "cudaMalloc(hl features)");
    CK(ctx->d_hl_pitch.reserve((TF + 1) * 8), "cudaMalloc(hl pitch)");
    CK(ctx->d_hl_status.reserve((size_t)(n + 1) * 4), "cudaMalloc(hl status)");
  }
  if (feat & AFX_FEAT_EXT_MELCHROMA) {
    CK(ctx->d_ext_mfcc.reserve((TF + 1) * 13 * 8), "cudaMalloc(ext mfcc)");
    CK(ctx->d_ext_chroma.reserve((TF + 1) * 12 * 8), "cudaMalloc(ext chroma)");
    CK(ctx->d_ext_idx.reserve((TF + 1) * 8), "cudaMalloc(ext chroma index)");
  }
  if (feat & AFX_FEAT_PACK) {
    CK(ctx->d_pack.reserve(b->pack_bytes + 64), "cudaMalloc(pack)");
    CK(ctx->d_pack_off.reserve((size_t)(n + 1) * (AFX_N_BLOBS + 1) * 4), "cudaMalloc(pack offsets)");
    CK(ctx->d_pack_file_off.reserve((size_t)(n + 1) * 8), "cudaMalloc(pack file offsets)");
  }
  CK(ctx->d_stats.reserve((size_t)(n + 1) * AFX_N_SERIES * AFX_N_STATS * 8), "cudaMalloc(stats)");
  CK(ctx->d_header.reserve((size_t)(n + 1) * AFX_N_HEADER * 8), "cudaMalloc(header)");

  // plan tables -> one pinned block -> one H2D copy
  const size_t nsc = b->src_chunk_file.size(), ndc = b->dst_chunk_file.size(), nrc = b->rs_chunk_file.size();
  size_t po = 0;
  auto place = [&](size_t bytes) { size_t o = po; po += (bytes + 255) & ~(size_t)255; return o; };
  const size_t p_files = place((size_t)n * sizeof(AfxFile));
  const size_t p_scf = place(nsc * 4), p_scs = place(nsc * 4), p_dcf = place(ndc * 4), p_dcs = place(ndc * 4),
    p_rcf = place(nrc * 4), p_rcs = place(nrc * 4);
  const size_t nrb = b->rs_blocks.size(), nck = b->rs_chk.size();
  const size_t p_rb = place(nrb * sizeof(RsBlock)), p_rbf = place(nrb * 4), p_chk = place(nck * 8);
  const size_t p_inj = place(b->inject.size() * sizeof(AfxInject));
  const size_t p_order = place((size_t)n * 4);
  const size_t p_pack = place(b->pack_file_off.size() * 8);
  take_cached(b->h_plan, ctx->h_plan_cache, po);
  CK(b->h_plan.reserve(po + 256), "cudaHostAlloc(plan)");
  CK(ctx->d_plan.reserve(po + 256), "cudaMalloc(plan)");
  unsigned char* hp = (unsigned char*)b->h_plan.p;
  if (n) memcpy(hp + p_files, b->files.data(), (size_t)n * sizeof(AfxFile));
  if (nsc) { memcpy(hp + p_scf, b->src_chunk_file.data(), nsc * 4); memcpy(hp + p_scs, b->src_chunk_start.data(), nsc * 4); }
  if (ndc) { memcpy(hp + p_dcf, b->dst_chunk_file.data(), ndc * 4); memcpy(hp + p_dcs, b->dst_chunk_start.data(), ndc * 4); }
  if (nrc) { memcpy(hp + p_rcf, b->rs_chunk_file.data(), nrc * 4); memcpy(hp + p_rcs, b->rs_chunk_start.data(), nrc * 4); }
  if (nrb) { memcpy(hp + p_rb, b->rs_blocks.data(), nrb * sizeof(RsBlock)); memcpy(hp + p_rbf, b->rs_blk_file.data(), nrb * 4); }
  if (nck) memcpy(hp + p_chk, b->rs_chk.data(), nck * 8);
  if (!b->inject.empty()) memcpy(hp + p_inj, b->inject.data(), b->inject.size() * sizeof(AfxInject));
  if (n) memcpy(hp + p_order, b->file_order.data(), (size_t)n * 4);
  if (!b->pack_file_off.empty()) memcpy(hp + p_pack, b->pack_file_off.data(), b->pack_file_off.size() * 8);

  CK(cudaEventRecord(b->ev[0], ctx->stream), "cudaEventRecord");
  CK(cudaMemcpyAsync(ctx->d_plan.p, hp, po, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(plan)");
  b->h2d_bytes = (long long)po;
  for (const auto& r : b->runs) {
    unsigned char* dst = (unsigned char*)(r.to_mono ? ctx->d_mono.p : ctx->d_pcm.p) + r.dev_off;
    CK(cudaMemcpyAsync(dst, r.host, r.bytes, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(pcm)");
    b->h2d_bytes += (long long)r.bytes;
  }
  CK(cudaEventRecord(b->ev[1], ctx->stream), "cudaEventRecord");

  unsigned char* dp = (unsigned char*)ctx->d_plan.p;
  AfxBatchDev& D = b->dev;
  memset(&D, 0, sizeof(D));
  D.n_files = n; D.TF = b->TF; D.TFr = b->TFr;
  D.pcm = (const unsigned char*)ctx->d_pcm.p; D.mono = (float*)ctx->d_mono.p; D.mono_src = (float*)ctx->d_mono_src.p;
  D.files = (const AfxFile*)(dp + p_files); D.state = (AfxState*)ctx->d_state.p;
  D.inject = b->inject.empty() ? nullptr : (const AfxInject*)(dp + p_inj);
  D.file_order = (const int*)(dp + p_order);
  D.mag = (double*)ctx->d_mag.p; D.cent_full = (double*)ctx->d_cent.p; D.fs = (double*)ctx->d_fs.p; D.fsr = (double*)ctx->d_fsr.p;
  D.fv = (double*)ctx->d_fv.p; D.rpolar = (float*)ctx->d_rpolar.p; D.rodf = (float*)ctx->d_rodf.p; D.rpost = (float*)ctx->d_rpost.p; D.bandraw = (double*)ctx->d_bandraw.p;
  D.slot_file = (const int*)ctx->d_slotmap.p; D.rslot_file = (const int*)ctx->d_slotmap.p + TF + 1;
  D.max_fr = b->max_fr;
  D.stats = (double*)ctx->d_stats.p; D.header = (double*)ctx->d_header.p; D.scratch = (double*)ctx->d_scratch.p;
  b->ext = ctx->ext_tables;
  b->ext.mfcc = (double*)ctx->d_ext_mfcc.p; b->ext.chroma = (double*)ctx->d_ext_chroma.p; b->ext.chroma_index = (double*)ctx->d_ext_idx.p;
  b->pack.packed = (unsigned char*)ctx->d_pack.p; b->pack.blob_off = (unsigned*)ctx->d_pack_off.p;
  b->pack.file_off = (const unsigned long long*)(dp + p_pack);
  b->hl.scalars = (double*)ctx->d_hl.p; b->hl.pitch = (double*)ctx->d_hl_pitch.p; b->hl.signature = (double*)ctx->d_hl_sig.p;
  b->hl.features = (double*)ctx->d_hl_feat.p; b->hl.status = (int*)ctx->d_hl_status.p; b->hl.silence_pad = ctx->P.t.hl_pad;
  AfxCondPlan& C = b->cond;
  memset(&C, 0, sizeof(C));
  C.src_chunk_file = (const int*)(dp + p_scf); C.src_chunk_start = (const int*)(dp + p_scs); C.n_src_chunks = (int)nsc;
  C.dst_chunk_file = (const int*)(dp + p_dcf); C.dst_chunk_start = (const int*)(dp + p_dcs); C.n_dst_chunks = (int)ndc;
  C.rs_chunk_file = (const int*)(dp + p_rcf); C.rs_chunk_start = (const int*)(dp + p_rcs); C.n_rs_chunks = (int)nrc;
  C.rs_blocks = (const RsBlock*)(dp + p_rb); C.rs_blk_file = (const int*)(dp + p_rbf); C.rs_times = (const double*)(dp + p_chk);
  C.n_rs_blocks = (int)nrb; C.rs_smem_bytes = b->rs_smem;
  b->uploaded = true;
  return AFX_OK;
}

static void ktime_begin(afx_batch* b, const char* name)
{
  if (!b->ctx->debug_times) return;
  KernelTime k; k.name = name;
  cudaEventCreate(&k.a); cudaEventCreate(&k.b);
  cudaEventRecord(k.a, b->ctx->stream);
  b->ktimes.push_back(k);
}
static void ktime_end(afx_batch* b)
{
  if (!b->ctx->debug_times) return;
  cudaEventRecord(b->ktimes.back().b, b->ctx->stream);
}

extern "C" int afx_batch_compute(afx_batch* b)
{
  if (!b) return AFX_ERR_ARG;
  afx_ctx* ctx = b->ctx;
  if (!b->uploaded) return fail(ctx, AFX_ERR_STATE, "afx_batch_compute: upload first");
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  const unsigned feat = ctx->cfg.features;
  for (auto& k : b->ktimes) { cudaEventDestroy(k.a); cudaEventDestroy(k.b); }
  b->ktimes.clear();
  b->launches = 0;
  std::unique_lock<std::mutex> chain_lk;
  if (ctx->compute_chain && ctx->device >= 0 && ctx->device < 64) {
    chain_lk = std::unique_lock<std::mutex>(g_chain_mu);            // held while this compute is enqueued: chain order = enqueue order
    if (g_chain[ctx->device] && g_chain[ctx->device] != ctx->ev_chain) CK(cudaStreamWaitEvent(ctx->stream, g_chain[ctx->device], 0), "cudaStreamWaitEvent(chain)");
  }
  CK(cudaEventRecord(b->ev[2], ctx->stream), "cudaEventRecord");
  for (const auto& t : b->rs_tails)   // samples past what libresample delivers stay 0 (SA.cpp:579-580: zero-initialised buffer)
    CK(cudaMemsetAsync((float*)ctx->d_mono.p + t.off, 0, (size_t)t.count * 4, ctx->stream), "cudaMemsetAsync(mono tail)");
  ktime_begin(b, "condition"); afx_launch_condition_plan(ctx->P, b->dev, b->cond, ctx->stream, &b->launches); ktime_end(b);
  const bool ms = ctx->multi_stream;
  cudaStream_t s_main = ctx->stream;
  cudaStream_t s_pitch = ms ? ctx->side[0] : s_main, s_ac = ms ? ctx->side[1] : s_main, s_rhythm = ms ? ctx->side[2] : s_main;
  if (ms) {   // the side chains start once the conditioning is done
    CK(cudaEventRecord(ctx->ev_fork, s_main), "cudaEventRecord");
    for (int i = 0; i < 3; ++i) CK(cudaStreamWaitEvent(ctx->side[i], ctx->ev_fork, 0), "cudaStreamWaitEvent");
  }
  for (const auto& g : b->groups) {
    AfxBatchDev D = b->dev;
    D.file0 = g.file0; D.g_files = g.nfiles; D.slot0 = g.slot0; D.g_slots = g.nslots; D.rslot0 = g.rslot0; D.g_rslots = g.nrslots;
    D.rhythm_fused = ctx->rhythm_fused(g.nfiles) ? 1 : 0;
    D.rhythm_pipe = ctx->rhythm_pipe;
    D.pitch_generic = ctx->pitch_generic ? 1 : 0;
    // every chain keeps to its own stream across groups, so the reuse of a chain's group scratch stays ordered
#ifdef AFX_HAVE_AUTOCORR
    if (feat & AFX_FEAT_AUTOCORR) { ktime_begin(b, "autocorr"); afx_launch_autocorr(ctx->P, D, s_ac, &b->launches); ktime_end(b); }
#endif
#ifdef AFX_HAVE_RHYTHM
    if (feat & AFX_FEAT_RHYTHM) { ktime_begin(b, "rhythm"); afx_launch_rhythm(ctx->P, D, s_rhythm, &b->launches); ktime_end(b); }
#endif
    if (feat & (AFX_FEAT_SPECTRAL | AFX_FEAT_AMPLITUDE | AFX_FEAT_PEAKS | AFX_FEAT_BANDS | AFX_FEAT_PITCH)) {
      ktime_begin(b, "spectrum"); afx_launch_spectrum(ctx->P, D, feat, s_main, &b->launches); ktime_end(b);
    }
    if (feat & AFX_FEAT_EXT_MELCHROMA) { ktime_begin(b, "ext"); afx_launch_ext(D, b->ext, ctx->ext_tensor, s_main, &b->launches); ktime_end(b); }
#ifdef AFX_HAVE_PITCH
    if (feat & AFX_FEAT_PITCH) {            // needs the spectrum's full-band centroid (fail-safe f0)
      if (ms) { CK(cudaEventRecord(ctx->ev_spec, s_main), "cudaEventRecord"); CK(cudaStreamWaitEvent(s_pitch, ctx->ev_spec, 0), "cudaStreamWaitEvent"); }
      ktime_begin(b, "pitch"); afx_launch_pitch(ctx->P, D, s_pitch, &b->launches); ktime_end(b);
    }
#endif
#ifdef AFX_HAVE_BANDS
    if (feat & AFX_FEAT_BANDS) { ktime_begin(b, "bands"); afx_launch_bands(ctx->P, D, s_main, &b->launches); ktime_end(b); }
#endif
    // peaks last among the users of the magnitude rows: its whitening pass overwrites them in place
#ifdef AFX_HAVE_PEAKS
    if (feat & AFX_FEAT_PEAKS) { ktime_begin(b, "peaks"); afx_launch_peaks(ctx->P, D, s_main, &b->launches); ktime_end(b); }
#endif
  }
  if (ms) {   // join before the statistics pass reads every series
    for (int i = 0; i < 3; ++i) { CK(cudaEventRecord(ctx->ev_join[i], ctx->side[i]), "cudaEventRecord"); CK(cudaStreamWaitEvent(s_main, ctx->ev_join[i], 0), "cudaStreamWaitEvent"); }
  }
#ifdef AFX_HAVE_STATS
  if (feat & AFX_FEAT_STATS) { ktime_begin(b, "stats"); afx_launch_stats(ctx->P, b->dev, feat, ctx->stream, &b->launches); ktime_end(b); }
#endif
  // the high-level stage reads the finished series and statistics (skipped while afx_create computes the silence pad)
  if ((feat & AFX_FEAT_HIGHLEVEL) && ctx->hl_pad_ready) { ktime_begin(b, "highlevel"); afx_launch_highlevel(ctx->P, b->dev, b->hl, ctx->stream, &b->launches); ktime_end(b); }
  if ((feat & AFX_FEAT_PACK) && b->n_files > 0) { ktime_begin(b, "pack"); afx_launch_pack(b->dev, b->pack, ctx->stream, &b->launches); ktime_end(b); }
  CK(cudaEventRecord(b->ev[3], ctx->stream), "cudaEventRecord");
  if (chain_lk.owns_lock()) { CK(cudaEventRecord(ctx->ev_chain, ctx->stream), "cudaEventRecord(chain)"); g_chain[ctx->device] = ctx->ev_chain; }
  CK(cudaGetLastError(), "kernel launch");
  b->computed = true;
  return AFX_OK;
}

// fs rows produced by a feature mask, as [begin, end) ranges
static void fs_row_ranges(unsigned feat, std::vector<std::pair<int, int>>& out)
{
  bool have[AFX_N_FS_MAIN] = { false };
  if (feat & AFX_FEAT_AMPLITUDE) for (int r = 0; r < 4; ++r) have[r] = true;
  if (feat & AFX_FEAT_SPECTRAL) { for (int r = 4; r <= 10; ++r) have[r] = true; have[FS_SPEC_FLUX] = true; }
  if (feat & AFX_FEAT_PEAKS) have[FS_SPEC_COMPLEXITY] = true;
  if (feat & AFX_FEAT_BANDS) have[FS_SPEC_CONTRAST] = true;
  if (feat & AFX_FEAT_PITCH) { have[FS_F0] = have[FS_F0_CONF] = have[FS_F0_FAILSAFE] = true; }
  if (feat & AFX_FEAT_AUTOCORR) have[FS_AUTOCORR] = true;
  int r = 0;
  while (r < AFX_N_FS_MAIN) {
    if (!have[r]) { ++r; continue; }
    int e = r; while (e < AFX_N_FS_MAIN && have[e]) ++e;
    out.push_back({ r, e }); r = e;
  }
}

static int batch_download(afx_batch* b, bool arrays);
extern "C" int afx_batch_download(afx_batch* b) { return batch_download(b, true); }
extern "C" int afx_batch_download_rows(afx_batch* b)
{
  if (b && !(b->ctx->cfg.features & AFX_FEAT_PACK)) return fail(b->ctx, AFX_ERR_STATE, "afx_batch_download_rows: the context was not created with AFX_FEAT_PACK");
  return batch_download(b, false);
}

static int batch_download(afx_batch* b, bool arrays)
{
  if (!b) return AFX_ERR_ARG;
  afx_ctx* ctx = b->ctx;
  if (!b->computed) return fail(ctx, AFX_ERR_STATE, "afx_batch_download: compute first");
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  const unsigned feat = ctx->cfg.features;
  const size_t bytes = b->total_doubles * 8 + 64;
  if (!b->h_results.p) take_cached(b->h_results, ctx->h_results_cache, bytes);
  CK(b->h_results.reserve(bytes), "cudaHostAlloc(results)");
  double* H = (double*)b->h_results.p;
  const int n = b->n_files; const size_t TF = (size_t)b->TF, TFr = (size_t)b->TFr;
  b->d2h_bytes = 0;
  CK(cudaEventRecord(b->ev[4], ctx->stream), "cudaEventRecord");
  auto cp = [&](void* dst, const void* src, size_t nbytes) -> cudaError_t {
    if (!nbytes) return cudaSuccess;
    b->d2h_bytes += (long long)nbytes;
    return cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx->stream);
  };
  CK(cp(H + b->o_header, b->dev.header, (size_t)n * AFX_N_HEADER * 8), "D2H header");
  CK(cp(H + b->o_state, b->dev.state, (size_t)n * sizeof(AfxState)), "D2H state");
  if (feat & AFX_FEAT_STATS) CK(cp(H + b->o_stats, b->dev.stats, (size_t)n * AFX_N_SERIES * AFX_N_STATS * 8), "D2H stats");
  b->arrays_downloaded = arrays;
  std::vector<std::pair<int, int>> rr; fs_row_ranges(feat, rr);
  if (arrays) for (auto& r : rr) CK(cp(H + b->o_fs + (size_t)r.first * TF, b->dev.fs + (size_t)r.first * TF, (size_t)(r.second - r.first) * TF * 8), "D2H fs");
  if (arrays && (feat & AFX_FEAT_RHYTHM)) CK(cp(H + b->o_fsr, b->dev.fsr, 2 * TFr * 8), "D2H fsr");
  if (arrays && (feat & AFX_FEAT_BANDS)) CK(cp(H + b->o_fv, b->dev.fv, TF * AFX_FV_STRIDE * 8), "D2H fv");
  if (feat & AFX_FEAT_EXT_MELCHROMA) {
    CK(cp(H + b->o_ext_mfcc, b->ext.mfcc, TF * 13 * 8), "D2H ext mfcc");
    CK(cp(H + b->o_ext_chroma, b->ext.chroma, TF * 12 * 8), "D2H ext chroma");
    CK(cp(H + b->o_ext_idx, b->ext.chroma_index, TF * 8), "D2H ext chroma index");
  }
  if ((feat & AFX_FEAT_PACK) && n > 0) {
    const size_t tab = (size_t)n * (AFX_N_BLOBS + 1) * 4;
    if (!b->h_pack.p) take_cached(b->h_pack, ctx->h_pack_cache, b->pack_bytes + tab + 64);
    CK(b->h_pack.reserve(b->pack_bytes + tab + 64), "cudaHostAlloc(pack)");
    CK(cp(b->h_pack.p, b->pack.packed, b->pack_bytes), "D2H pack");
    CK(cp((unsigned char*)b->h_pack.p + b->pack_bytes, b->pack.blob_off, tab), "D2H pack offsets");
  }
  if ((feat & AFX_FEAT_HIGHLEVEL) && ctx->hl_pad_ready) {
    CK(cp(H + b->o_hl, b->hl.scalars, (size_t)n * AFX_N_HL * 8), "D2H hl");
    CK(cp(H + b->o_hl_sig, b->hl.signature, (size_t)n * AFX_HL_SIGNATURE * 8), "D2H hl signature");
    CK(cp(H + b->o_hl_feat, b->hl.features, (size_t)n * AFX_HL_FEATURES * 8), "D2H hl features");
    CK(cp(H + b->o_hl_pitch, b->hl.pitch, TF * 8), "D2H hl pitch");
    CK(cp(H + b->o_hl_status, b->hl.status, (size_t)n * sizeof(int)), "D2H hl status");
  }
  CK(cudaEventRecord(b->ev[5], ctx->stream), "cudaEventRecord");
  b->downloaded = true;
  return AFX_OK;
}

extern "C" int afx_batch_sync(afx_batch* b)
{
  if (!b) return AFX_ERR_ARG;
  afx_ctx* ctx = b->ctx;
  cudaSetDevice(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  return AFX_OK;
}

extern "C" int afx_analyze(afx_ctx* ctx, const afx_file* files, int32_t n_files, afx_batch** out)
{
  int rc = afx_batch_create(ctx, files, n_files, out);
  if (rc != AFX_OK) return rc;
  afx_batch* b = *out;
  if ((rc = afx_batch_upload(b)) != AFX_OK || (rc = afx_batch_compute(b)) != AFX_OK ||
      (rc = afx_batch_download(b)) != AFX_OK || (rc = afx_batch_sync(b)) != AFX_OK) {
    afx_batch_free(b); *out = nullptr; return rc;
  }
  return AFX_OK;
}

extern "C" int afx_batch_result(const afx_batch* b, int32_t i, afx_file_result* out)
{
  if (!b || !out || i < 0 || i >= b->n_files) return AFX_ERR_ARG;
  if (!b->downloaded) return fail(b->ctx, AFX_ERR_STATE, "afx_batch_result: download + sync first");
  memset(out, 0, sizeof(*out));
  const AfxFile& f = b->files[i];
  out->status = f.status;
  const double* H = (const double*)b->h_results.p;
  out->header = H + b->o_header + (size_t)i * AFX_N_HEADER;
  if (f.status != AFX_FILE_OK) return AFX_OK;
  const AfxState* st = reinterpret_cast<const AfxState*>(H + b->o_state) + i;
  out->n_frames = st->F; out->n_rhythm_frames = st->Fr;
  const unsigned feat = b->ctx->cfg.features;
  if ((feat & AFX_FEAT_PACK) && b->h_pack.p) {
    out->packed = (const unsigned char*)b->h_pack.p + b->pack_file_off[i];
    out->packed_off = reinterpret_cast<const uint32_t*>((const unsigned char*)b->h_pack.p + b->pack_bytes) + (size_t)i * (AFX_N_BLOBS + 1);
  }
  if (feat & AFX_FEAT_STATS) out->stats = H + b->o_stats + (size_t)i * AFX_N_SERIES * AFX_N_STATS;
  if (b->arrays_downloaded) {
    std::vector<std::pair<int, int>> rr; fs_row_ranges(feat, rr);
    const size_t TF = (size_t)b->TF, TFr = (size_t)b->TFr;
    for (auto& r : rr) for (int s = r.first; s < r.second; ++s) out->fs[s] = H + b->o_fs + (size_t)s * TF + f.frame_off;
    if (feat & AFX_FEAT_SPECTRAL) {   // constant-zero series (degenerate in the reference)
      out->fs[FS_SPEC_INHARM] = out->fs[FS_TRISTIM1] = out->fs[FS_TRISTIM2] = out->fs[FS_TRISTIM3] = b->ctx->zeros.data();
    }
    if (feat & AFX_FEAT_RHYTHM) for (int s = 0; s < 2; ++s) out->fs[AFX_N_FS_MAIN + s] = H + b->o_fsr + (size_t)s * TFr + f.rframe_off;
    if (feat & AFX_FEAT_BANDS) {
      static const int offs[AFX_N_FV] = { FV_RMS, FV_FLATNESS, FV_FLUX, FV_COMPLEXITY, FV_CONTRAST, FV_BANDS28, FV_CEPSTRUM };
      static const int nbv[AFX_N_FV] = { 14, 14, 14, 14, 14, 28, 14 };
      for (int v = 0; v < AFX_N_FV; ++v) out->fv[v] = H + b->o_fv + (size_t)offs[v] * TF + (size_t)f.frame_off * nbv[v];
    }
  }
  if (feat & AFX_FEAT_EXT_MELCHROMA) {
    out->ext_mfcc = H + b->o_ext_mfcc + (size_t)f.frame_off * 13;
    out->ext_chroma = H + b->o_ext_chroma + (size_t)f.frame_off * 12;
    out->ext_chroma_index = H + b->o_ext_idx + f.frame_off;
  }
  if ((feat & AFX_FEAT_HIGHLEVEL) && b->ctx->hl_pad_ready) {
    out->highlevel = H + b->o_hl + (size_t)i * AFX_N_HL;
    out->hl_signature = H + b->o_hl_sig + (size_t)i * AFX_HL_SIGNATURE;
    out->hl_features = H + b->o_hl_feat + (size_t)i * AFX_HL_FEATURES;
    out->hl_pitch = H + b->o_hl_pitch + f.frame_off;
    out->hl_status = reinterpret_cast<const int*>(H + b->o_hl_status)[i];
  }
  return AFX_OK;
}

extern "C" void afx_batch_free(afx_batch* b)
{
  if (!b) return;
  afx_ctx* ctx = b->ctx;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    give_back(b->h_results, ctx->h_results_cache);
    give_back(b->h_plan, ctx->h_plan_cache);
    give_back(b->h_pack, ctx->h_pack_cache);
    if (ctx->live == b) ctx->live = nullptr;
  }
  for (int i = 0; i < 6; ++i) if (b->ev[i]) cudaEventDestroy(b->ev[i]);
  for (auto& k : b->ktimes) { cudaEventDestroy(k.a); cudaEventDestroy(k.b); }
  delete b;
}

extern "C" int afx_batch_timings(const afx_batch* b, float* up, float* comp, float* down)
{
  if (!b) return AFX_ERR_ARG;
  cudaSetDevice(b->ctx->device);
  float v;
  if (up) { *up = 0; if (b->uploaded && cudaEventElapsedTime(&v, b->ev[0], b->ev[1]) == cudaSuccess) *up = v; }
  if (comp) { *comp = 0; if (b->computed && cudaEventElapsedTime(&v, b->ev[2], b->ev[3]) == cudaSuccess) *comp = v; }
  if (down) { *down = 0; if (b->downloaded && cudaEventElapsedTime(&v, b->ev[4], b->ev[5]) == cudaSuccess) *down = v; }
  return AFX_OK;
}

extern "C" int afx_batch_counters(const afx_batch* b, int64_t* launches, int64_t* h2d, int64_t* d2h, int64_t* frames, int64_t* rframes)
{
  if (!b) return AFX_ERR_ARG;
  if (launches) *launches = b->launches;
  if (h2d) *h2d = b->h2d_bytes;
  if (d2h) *d2h = b->d2h_bytes;
  long long F = 0, Fr = 0;
  if (b->downloaded) {
    const AfxState* st = reinterpret_cast<const AfxState*>((const double*)b->h_results.p + b->o_state);
    for (int i = 0; i < b->n_files; ++i) if (b->files[i].status == 0) { F += st[i].F; Fr += st[i].Fr; }
  }
  if (frames) *frames = F;
  if (rframes) *rframes = Fr;
  return AFX_OK;
}

extern "C" int afx_batch_kernel_times(const afx_batch* b, const char** names, float* ms, int32_t cap)
{
  if (!b) return AFX_ERR_ARG;
  cudaSetDevice(b->ctx->device);
  int n = 0;
  std::vector<std::pair<const char*, float>> agg;     // one entry per kernel group, summed over the launch groups
  for (const auto& k : b->ktimes) {
    float v = 0; cudaEventElapsedTime(&v, k.a, k.b);
    bool found = false;
    for (auto& a : agg) if (!strcmp(a.first, k.name)) { a.second += v; found = true; break; }
    if (!found) agg.push_back({ k.name, v });
  }
  for (const auto& a : agg) {
    if (n >= cap) break;
    if (names) names[n] = a.first;
    if (ms) ms[n] = a.second;
    ++n;
  }
  return n;
}

extern "C" int64_t afx_batch_conditioned(const afx_batch* b, int32_t i, double* out, int64_t cap)
{
  if (!b || i < 0 || i >= b->n_files) return AFX_ERR_ARG;
  afx_ctx* ctx = b->ctx;
  if (!b->computed) return fail(ctx, AFX_ERR_STATE, "afx_batch_conditioned: compute first");
  cudaSetDevice(ctx->device);
  AfxState st;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return AFX_ERR_CUDA;
  if (cudaMemcpy(&st, b->dev.state + i, sizeof(st), cudaMemcpyDeviceToHost) != cudaSuccess) return AFX_ERR_CUDA;
  if (b->files[i].status != 0) return 0;
  if (!out || cap < st.len) return st.len;
  double* d = nullptr;
  if (cudaMalloc(&d, (size_t)st.len * 8) != cudaSuccess) return AFX_ERR_NOMEM;
  afx_launch_materialise(b->dev.mono + b->files[i].mono_off, b->dev.state + i, d, st.len, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaMemcpy(out, d, (size_t)st.len * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return st.len;
}

// ---- test hook: the FFT core on caller data -------------------------------------------------------------
int afx_debug_fft_launch(int n, int batch, const double2* in, double2* out, const AfxTables& T, cudaStream_t s);
extern "C" int afx_debug_fft(afx_ctx* ctx, int32_t n, int32_t batch, const double* in, double* out)
{
  if (!ctx || !in || !out || batch <= 0 || (n != 256 && n != 1024 && n != 2048)) return fail(ctx, AFX_ERR_ARG, "afx_debug_fft: bad arguments");
  cudaSetDevice(ctx->device);
  const size_t bytes = (size_t)n * batch * 16;
  double2 *di = nullptr, *dout = nullptr;
  CK(cudaMalloc(&di, bytes), "cudaMalloc"); CK(cudaMalloc(&dout, bytes), "cudaMalloc");
  cudaMemcpy(di, in, bytes, cudaMemcpyHostToDevice);
  afx_debug_fft_launch(n, batch, di, dout, ctx->P.t, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost);
  cudaFree(di); cudaFree(dout);
  if (e != cudaSuccess) return fail(ctx, AFX_ERR_CUDA, "afx_debug_fft", e);
  return AFX_OK;
}

// ---- FP64 FMA peak (roofline denominator) -----------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters)
{
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}

extern "C" int afx_measure_fp64_peak(afx_ctx* ctx, double* tflops)
{
  if (!ctx || !tflops) return AFX_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, ctx->device), "cudaGetDeviceProperties");
  double* d = nullptr;
  CK(cudaMalloc(&d, 8), "cudaMalloc");
  const int blocks = prop.multiProcessorCount * 8, iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a, ctx->stream);
    k_fp64_peak<<<blocks, 256, 0, ctx->stream>>>(d, iters);
    cudaEventRecord(b, ctx->stream);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double fl = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
    if (rep > 0 && ms > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
  *tflops = best;
  return AFX_OK;
}

/*
 * TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
 *
 * afec_oracle.c: plain-C, single-threaded CPU restatement of the reference's
 * (emuell/AFEC) low-level descriptor hot path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load it; the CUDA library never does.
 *
 * Parity pinning: the reference's tests pin only TStatistics
 * (FeatureExtraction/Test/TestStatistics.cpp:16-113, checked in
 * tests/test_oracle_kats.py).  Every descriptor value is pinned against the
 * UNMODIFIED reference compiled from /root/reference (oracle/_ref/afec_ref,
 * recipe oracle/build_ref.sh) through tests/test_oracle_vs_reference.py and the
 * committed fixtures under tests/golden/.
 *
 * Each function cites the reference file:line it restates.  Paths are relative
 * to the reference root; "SA.cpp" = Source/Crawler/FeatureExtraction/Source/SampleAnalyser.cpp.
 * No reference source text is reproduced: algorithms are re-expressed.
 *
 * The output record layout ("AFXD") is described in afec_b200/layout.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AFX_N_HEADER 32
#define AFX_N_FS 24
#define AFX_N_FS_MAIN 22
#define AFX_N_STATS 13
#define AFX_N_SERIES 136
#define AFX_FV_BANDS 112

#define NB14 14
#define NB28 28
#define NCEP 14

/* InlineMath.h:32 -- a float literal promoted to double in every comparison */
static const double kEps = (double)1e-12f;
static const double kPi = 3.1415926535897932384626433832795;

/* ------------------------------------------------------------------------------------------- */
/* small helpers                                                                               */

/* TMath::d2iRound (CoreTypes/Export/InlineMath.inl:823-826): sign taken from the sign bit */
static int d2i_round(double v) { return (int)(v + (signbit(v) ? -0.5 : 0.5)); }
/* TMath::f2iRound (InlineMath.inl:758-761) */
static int f2i_round(float v) { return (int)(v + (signbit(v) ? -0.5f : 0.5f)); }
/* TAudioMath::MsToSamples / SamplesToMs (AudioTypes/Export/AudioMath.inl:127-137): float math */
static int ms_to_samples(int sr, float ms) { return f2i_round((float)sr / 1000.0f * ms); }
static float samples_to_ms(int sr, int samples) { return (float)samples / ((float)sr / 1000.0f); }

/* TAudioMath::LinToDb(double) (AudioMath.inl:55-70) */
static double lin_to_db(double v)
{
  if (v == 1.0) return 0.0;
  if (v > kEps) return log(v) * (20.0 / log(10.0));
  return -200.0;
}
/* TAudioMath::DbToLin(double) (AudioMath.inl:107-123) */
static double db_to_lin(double v)
{
  if (v == 0.0) return 1.0;
  if (v > -200.0) return exp(v * (log(10.0) / 20.0));
  return 0.0;
}

/* ------------------------------------------------------------------------------------------- */
/* FFT: iterative radix-2, double.  sign=+1 matches ooura_cdft(.., isgn=1, ..) as called by     */
/* TFftTransformComplex::ForwardInplace (AudioTypes/Source/Fourier.cpp:219-274):                */
/*   X[k] = sum_j x[j] exp(+2 pi i j k / n)                                                    */

/* A second, mathematically identical transform (decimation in frequency: butterflies first, bit reversal last), so   */
/* that tests can show which outputs of the PATH are decided by the FFT's last-bit rounding -- the situation of the     */
/* reference itself built with another FFT back end (Fourier.cpp selects IPP / vDSP / Ooura per platform).              */
static int g_fft_variant = 0;
void afxo_set_fft_variant(int v) { g_fft_variant = v; }

static void fft_c2c_dif(double* re, double* im, int n, int sign)
{
  int i, j, len;
  for (len = n; len >= 2; len >>= 1) {
    const int half = len >> 1;
    for (i = 0; i < n; i += len) {
      for (j = 0; j < half; ++j) {
        const double ang = sign * 2.0 * kPi * (double)j / (double)len;
        const double wr = cos(ang), wi = sin(ang);
        const double ar = re[i + j], ai = im[i + j], br = re[i + j + half], bi = im[i + j + half];
        const double dr = ar - br, di = ai - bi;
        re[i + j] = ar + br; im[i + j] = ai + bi;
        re[i + j + half] = dr * wr - di * wi; im[i + j + half] = dr * wi + di * wr;
      }
    }
  }
  for (i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      double t = re[i]; re[i] = re[j]; re[j] = t;
      t = im[i]; im[i] = im[j]; im[j] = t;
    }
  }
}

static void fft_c2c(double* re, double* im, int n, int sign)
{
  int i, j, len;
  if (g_fft_variant) { fft_c2c_dif(re, im, n, sign); return; }
  for (i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      double t = re[i]; re[i] = re[j]; re[j] = t;
      t = im[i]; im[i] = im[j]; im[j] = t;
    }
  }
  for (len = 2; len <= n; len <<= 1) {
    const int half = len >> 1;
    for (i = 0; i < n; i += len) {
      for (j = 0; j < half; ++j) {
        const double ang = sign * 2.0 * kPi * (double)j / (double)len;
        const double wr = cos(ang), wi = sin(ang);
        const double ur = re[i + j], ui = im[i + j];
        const double vr = re[i + j + half] * wr - im[i + j + half] * wi;
        const double vi = re[i + j + half] * wi + im[i + j + half] * wr;
        re[i + j] = ur + vr; im[i + j] = ui + vi;
        re[i + j + half] = ur - vr; im[i + j + half] = ui - vi;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------- */
/* TStatistics (FeatureExtraction/Source/Statistics.cpp)                                       */

static double st_sum(const double* x, int n) { double s = 0; for (int i = 0; i < n; ++i) s += x[i]; return s; }

/* Statistics.cpp:249-266 */
static double st_mean(const double* x, int n)
{
  if (n >= 2) return st_sum(x, n) / (double)n;
  return n == 1 ? x[0] : 0.0;
}
/* Statistics.cpp:275-300 */
static double st_variance(const double* x, int n, double mean)
{
  if (n < 2) return 0.0;
  double r = 0; for (int i = 0; i < n; ++i) r += (x[i] - mean) * (x[i] - mean);
  return r / n;
}
static int cmp_double(const void* a, const void* b)
{
  const double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}
/* Statistics.cpp:316-413: quick-select of element (n-1)/2 == the lower median */
static double st_median(const double* x, int n)
{
  if (n >= 2) {
    double* t = (double*)malloc(sizeof(double) * n);
    memcpy(t, x, sizeof(double) * n);
    qsort(t, n, sizeof(double), cmp_double);
    const double m = t[(n - 1) / 2];
    free(t);
    return m;
  }
  return n == 1 ? x[0] : 0.0;
}
/* Statistics.cpp:417-455: running product, folded into a log sum when it leaves [1e-64, 1e64] */
static double st_gmean(const double* x, int n)
{
  if (n >= 2) {
    double sumlog = 0.0, prod = 1.0;
    for (int i = 0; i < n; ++i) {
      prod *= (fabs(x[i]) + 1e-20);
      if (prod > 1.e64 || prod < 1.e-64) { sumlog += log(prod); prod = 1.0; }
    }
    return exp((sumlog + log(prod)) / (double)n);
  }
  return n == 1 ? x[0] : 0.0;
}
/* Statistics.cpp:459-477 */
static double st_centroid(const double* x, int n)
{
  const double s = st_sum(x, n);
  if (s == 0.0) return 0.0;
  double sc = 0; for (int j = 0; j < n; ++j) sc += (double)j * x[j];
  return sc / s;
}
/* Statistics.cpp:486-506 */
static double st_spread(const double* x, int n, double c)
{
  const double s = st_sum(x, n);
  if (s == 0.0) return 0.0;
  double sc = 0; for (int j = 0; j < n; ++j) { const double t = j - c; sc += t * t * x[j]; }
  return sc / s;
}
/* Statistics.cpp:510-528: value minus centroid over spread, cubed; summed from the back */
static double st_skewness(const double* x, int n, double c, double sp)
{
  if (!n || fabs(sp) <= kEps) return 0.0;
  double r = 0; for (int i = n - 1; i >= 0; --i) { const double t = (x[i] - c) / sp; r += t * t * t; }
  return r / n;
}
/* Statistics.cpp:532-554 */
static double st_kurtosis(const double* x, int n, double c, double sp)
{
  if (!n || fabs(sp) <= kEps) return 0.0;
  double r = 0;
  for (int i = n - 1; i >= 0; --i) { const double t = (x[i] - c) / sp; const double tt = t * t; r += tt * tt; }
  return r / n - 3.0;
}
/* Statistics.cpp:565-574 */
static double st_flatness2(double mean, double gmean) { return mean == 0.0 ? 0.0 : gmean / mean; }
static double st_flatness(const double* x, int n) { return st_flatness2(st_mean(x, n), st_gmean(x, n)); }
/* Statistics.cpp:604-638 Pearson correlation; 578-600 "flux" == correlation */
static double st_correlation(const double* a, const double* b, int n)
{
  if (!n) return 0.0;
  double s1 = 0, s2 = 0, s11 = 0, s12 = 0, s22 = 0;
  for (int i = 0; i < n; ++i) {
    s12 = s12 + a[i] * b[i]; s1 = s1 + a[i]; s11 = s11 + a[i] * a[i];
    s2 = s2 + b[i]; s22 = s22 + b[i] * b[i];
  }
  s1 = s1 / n; s2 = s2 / n;
  const double den2 = (s11 - s1 * s1 * n) * (s22 - s2 * s2 * n);
  const double num = s12 - (s1 * s2 * n);
  if (fabs(den2) > kEps) return num / sqrt(den2);
  return 0.0;
}

/* Statistics.cpp:140-232.  Walks down / up / along plateaus; returns the number of peak
 * entries written to bins[] (bins may repeat in degenerate tails, as in the reference). */
static int st_peaks(const double* a, int n, double thr, int* bins, double* vals)
{
  int cnt = 0;
  if (n <= 2) return 0;
  int i = 0;
  if (a[0] > a[1] && a[0] > thr) { bins[cnt] = 0; vals[cnt++] = a[0]; }
  for (;;) {
    while (i + 1 < n - 1 && a[i] >= a[i + 1]) i++;
    while (i + 1 < n - 1 && a[i] < a[i + 1]) i++;
    int j = i;
    while (j + 1 < n - 1 && a[j] == a[j + 1]) j++;
    if (j + 1 < n - 1 && a[j + 1] < a[j] && a[j] > thr) {
      if (j != i) { bins[cnt] = (i + j) / 2; vals[cnt++] = a[i]; }
      else { bins[cnt] = j; vals[cnt++] = a[j]; }
    }
    i = j;
    if (i + 1 >= n - 1) {
      if (i == n - 2 && a[i - 1] < a[i] && a[i + 1] < a[i] && a[i] > thr) { bins[cnt] = i; vals[cnt++] = a[i]; }
      break;
    }
  }
  if (a[n - 1] > a[n - 2] && a[n - 1] > thr) { bins[cnt] = n - 1; vals[cnt++] = a[n - 1]; }
  return cnt;
}

/* TStatistics::Calc (Statistics.cpp:12-90).  out[13] must be zero-initialised by the caller:
 * for n == 1 only min/max/mean (+ zero variance/dmean/dvariance) are assigned. */
static void st_calc13(const double* x, int n, double* out)
{
  if (n > 1) {
    double mn = x[0], mx = x[0];
    for (int i = 1; i < n; ++i) { mn = x[i] < mn ? x[i] : mn; mx = x[i] > mx ? x[i] : mx; }
    out[0] = mn; out[1] = mx;
    out[2] = st_median(x, n);
    out[3] = st_mean(x, n);
    out[4] = st_gmean(x, n);
    out[5] = st_variance(x, n, out[3]);
    out[6] = st_centroid(x, n);
    out[7] = st_spread(x, n, out[6]);
    out[8] = st_skewness(x, n, out[6], out[7]);
    out[9] = st_kurtosis(x, n, out[6], out[7]);
    out[10] = st_flatness2(out[3], out[4]);
    if (n > 2) {
      double* d = (double*)malloc(sizeof(double) * (n - 1));
      for (int i = 0; i < n - 1; ++i) d[i] = fabs(x[i + 1] - x[i]);
      out[11] = st_mean(d, n - 1);
      out[12] = st_variance(d, n - 1, out[11]);
      free(d);
    } else { out[11] = 0; out[12] = 0; }
  } else if (n > 0) {
    out[0] = x[0]; out[1] = x[0]; out[3] = x[0]; out[5] = 0; out[11] = 0; out[12] = 0;
  } else {
    out[0] = 0; out[1] = 0; out[3] = 0; out[5] = 0; out[11] = 0; out[12] = 0;
  }
}

/* exported for the KAT tests */
void afxo_stats13(const double* x, int n, double* out13) { memset(out13, 0, 13 * sizeof(double)); st_calc13(x, n, out13); }
int afxo_peaks(const double* a, int n, double thr, int* bins, double* vals) { return st_peaks(a, n, thr, bins, vals); }
double afxo_variance(const double* x, int n) { return st_variance(x, n, st_mean(x, n)); }
double afxo_centroid(const double* x, int n) { return st_centroid(x, n); }
double afxo_spread(const double* x, int n) { return st_spread(x, n, st_centroid(x, n)); }
double afxo_skewness(const double* x, int n) { const double c = st_centroid(x, n); return st_skewness(x, n, c, st_spread(x, n, c)); }
double afxo_kurtosis(const double* x, int n) { const double c = st_centroid(x, n); return st_kurtosis(x, n, c, st_spread(x, n, c)); }
double afxo_flatness(const double* x, int n) { return st_flatness(x, n); }
double afxo_flux(const double* a, const double* b, int n) { return st_correlation(a, b, n); }
void afxo_fft(double* re, double* im, int n, int sign) { fft_c2c(re, im, n, sign); }

/* ------------------------------------------------------------------------------------------- */
/* libresample restatement (3rdParty/Resample/Dist/src)                                        */

#define RS_NPC 4096

/* filterkit.c:67-81 */
static double rs_izero(double x)
{
  double sum = 1, u = 1, halfx = x / 2.0; int n = 1;
  do { double t = halfx / (double)n; n += 1; t *= t; u *= t; sum += u; } while (u >= 1E-21 * sum);
  return sum;
}
/* filterkit.c:84-113 Kaiser-windowed sinc, one wing; resample.c:113-127 float copy */
static float* rs_design(int nwing)
{
  double* c = (double*)malloc(sizeof(double) * nwing);
  const double frq = 0.5 * 0.90, beta = 6;
  c[0] = 2.0 * frq;
  for (int i = 1; i < nwing; ++i) { const double t = 3.14159265358979232846 /* resample_defs.h:30, sic */ * (double)i / (double)RS_NPC; c[i] = sin(2.0 * t * frq) / t; }
  const double ibeta = 1.0 / rs_izero(beta), inm1 = 1.0 / ((double)(nwing - 1));
  for (int i = 1; i < nwing; ++i) {
    const double t = (double)i * inm1; double t1 = 1.0 - t * t; t1 = (t1 < 0 ? 0 : t1);
    c[i] *= rs_izero(beta * sqrt(t1)) * ibeta;
  }
  float* imp = (float*)malloc(sizeof(float) * nwing);
  for (int i = 0; i < nwing; ++i) imp[i] = (float)c[i];
  free(c);
  return imp;
}
/* filterkit.c:115-164 (no coefficient interpolation) */
static float rs_filter_up(const float* imp, int nwing, const float* xp, double ph, int inc)
{
  ph *= RS_NPC;
  float v = 0.0f;
  int h = (int)ph, end = nwing;
  if (inc == 1) { end--; if (ph == 0) h += RS_NPC; }
  while (h < end) { float t = imp[h]; t *= *xp; v += t; h += RS_NPC; xp += inc; }
  return v;
}
/* filterkit.c:166-215 (no coefficient interpolation) */
static float rs_filter_ud(const float* imp, int nwing, const float* xp, double ph, int inc, double dhb)
{
  float v = 0.0f;
  double ho = ph * dhb;
  int end = nwing;
  if (inc == 1) { end--; if (ph == 0) ho += dhb; }
  while ((int)ho < end) { float t = imp[(int)ho]; t *= *xp; v += t; ho += dhb; xp += inc; }
  return v;
}

/* resample.c:80-164 (open, high quality) + :170-337 (process, lastFlag = 1, whole buffer) +
 * resamplesubs.c:30-123 (block kernels).  Writes at most out_len samples; returns #written. */
static int rs_resample(const float* in, int in_len, double factor, float* out, int out_len)
{
  const int nmult = 35, nwing = RS_NPC * (nmult - 1) / 2;
  float* imp = rs_design(nwing);
  const double inv = 1.0 / factor;
  const unsigned xoff = (unsigned)(((nmult + 1) / 2.0) * (inv > 1.0 ? inv : 1.0) + 10);
  const unsigned xsize = (2 * xoff + 10 > 4096) ? 2 * xoff + 10 : 4096;
  float* X = (float*)malloc(sizeof(float) * (xsize + xoff));
  const int ysize = (int)(((double)xsize) * factor + 2.0);
  float* Y = (float*)malloc(sizeof(float) * (ysize + 16));
  unsigned xp = xoff, xread = xoff;
  for (unsigned i = 0; i < xoff; ++i) X[i] = 0;
  double time = (double)xoff;
  float lpscl = 1.0f;
  if (factor < 1) lpscl = (float)(lpscl * factor);
  const double dt = 1.0 / factor;
  double dh = factor * RS_NPC; if (dh > RS_NPC) dh = RS_NPC;
  int used = 0, outc = 0;
  for (;;) {
    int len = (int)(xsize - xread);
    if (len >= in_len - used) len = in_len - used;
    for (int i = 0; i < len; ++i) X[xread + i] = in[used + i];
    used += len; xread += len;
    int nx;
    if (used == in_len) { nx = (int)(xread - xoff); for (unsigned i = 0; i < xoff; ++i) X[xread + i] = 0; }
    else nx = (int)(xread - 2 * xoff);
    if (nx <= 0) break;
    /* one block */
    int nout = 0;
    {
      double t = time; const double end_time = t + nx;
      while (t < end_time) {
        const double lph = t - floor(t), rph = 1.0 - lph;
        const float* p = &X[(int)t];
        float v;
        if (factor >= 1) { v = rs_filter_up(imp, nwing, p, lph, -1); v += rs_filter_up(imp, nwing, p + 1, rph, 1); }
        else { v = rs_filter_ud(imp, nwing, p, lph, -1, dh); v += rs_filter_ud(imp, nwing, p + 1, rph, 1, dh); }
        v *= lpscl;
        Y[nout++] = v;
        t += dt;
      }
      time = t;
    }
    time -= nx; xp += nx;
    const unsigned ncreep = (unsigned)((int)time - (int)xoff);
    if (ncreep) { time -= ncreep; xp += ncreep; }
    const unsigned nreuse = xread - (xp - xoff);
    for (unsigned i = 0; i < nreuse; ++i) X[i] = X[i + (xp - xoff)];
    xread = nreuse; xp = xoff;
    int ncopy = out_len - outc; if (ncopy > nout) ncopy = nout;
    for (int i = 0; i < ncopy; ++i) out[outc + i] = Y[i];
    outc += ncopy;
    if (ncopy < nout) break;  /* output buffer full */
  }
  free(imp); free(X); free(Y);
  return outc;
}

/* ------------------------------------------------------------------------------------------- */
/* analyser constants (TSampleAnalyser ctor, SA.cpp:162-198)                                   */

typedef struct {
  int sr, N, H;
  int first_bin, last_bin, nbins;
  double* window;          /* Hann * 2 */
  double* mel;             /* [14][N/2] */
  int band14_n[NB14];
  int band28_s[NB28], band28_e[NB28];
  double wh_decay;         /* awhitening.c:84-86 */
  double env_coef;         /* Envelopes.cpp:66-69 */
  double silence_floor;    /* SA.cpp:648-649 */
} analyser_t;

static const double kBands14[NB14] = { 50.0, 100.0, 200.0, 400.0, 630.0, 920.0, 1270.0, 1720.0, 2320.0,
  3150.0, 4400.0, 6400.0, 9500.0, 15500.0 };
static const double kBands28[NB28] = { 50.0, 100.0, 150.0, 200.0, 300.0, 400.0, 510.0, 630.0, 770.0, 920.0,
  1080.0, 1270.0, 1480.0, 1720.0, 2000.0, 2320.0, 2700.0, 3150.0, 3700.0, 4400.0, 5300.0, 6400.0, 7700.0,
  9500.0, 12000.0, 15500.0, 19000.0, 22050.0 };

/* LibXtract init.c:237-378 (equal gain), called as xtract_init_mfcc(N/2, sr/2, EQUAL_GAIN, 20, 15500, 14) */
static void init_mel(double* tab, int nbins, double nyquist, double fmin, double fmax, int nf)
{
  const int M = nbins >> 1;
  double mel_peak[NCEP + 2], lin_peak[NCEP + 2]; int fft_peak[NCEP + 2];
  const double mel_max = 1127 * log(1 + fmax / 700), mel_min = 1127 * log(1 + fmin / 700);
  const double bw = (mel_max - mel_min) / nf;
  mel_peak[0] = mel_min; lin_peak[0] = fmin; fft_peak[0] = (int)(lin_peak[0] / nyquist * M);
  for (int n = 1; n < nf + 2; ++n) {
    mel_peak[n] = mel_peak[n - 1] + bw;
    lin_peak[n] = 700 * (exp(mel_peak[n] / 1127) - 1);
    fft_peak[n] = (int)(lin_peak[n] / nyquist * M);
  }
  int i = 0;
  for (int n = 0; n < nf; ++n) {
    double* row = tab + (size_t)n * nbins;
    double inc = (n == 0) ? 1.0 / fft_peak[n] : 1.0 / (fft_peak[n] - fft_peak[n - 1]);
    double val = 0;
    for (int k = 0; k < i; ++k) row[k] = 0.0;
    for (; i <= fft_peak[n]; ++i) { row[i] = val; val += inc; }
    inc = 1.0 / (fft_peak[n + 1] - fft_peak[n]);
    val = 0;
    const int next = fft_peak[n + 1];
    for (i = next; i > fft_peak[n]; --i) { row[i] = val; val += inc; }
    for (int k = next + 1; k < nbins; ++k) row[k] = 0.0;
  }
}

static void analyser_init(analyser_t* a, int sr, int N, int H)
{
  a->sr = sr; a->N = N; a->H = H;
  const double fpb = (double)(sr / N);                      /* integer division: 21 (SA.cpp:171) */
  a->first_bin = d2i_round(20.0 / fpb);
  a->last_bin = d2i_round(15500.0 / fpb);
  a->nbins = a->last_bin - a->first_bin + 1;
  a->window = (double*)malloc(sizeof(double) * N);
  for (int n = 0; n < N; ++n)                               /* LibXtract window.c:67-76, then x2 */
    a->window[n] = (0.5 * (1.0 - cos(2.0 * M_PI * (double)n / (double)(N - 1)))) * 2.0;
  a->mel = (double*)malloc(sizeof(double) * NCEP * (N / 2));
  init_mel(a->mel, N / 2, sr / 2, 20.0, 15500.0, NCEP);
  for (int b = 0; b < NB14; ++b) {                          /* SA.cpp:2090-2100 */
    const int s = (b == 0) ? a->first_bin : d2i_round(kBands14[b - 1] / fpb);
    const int e = d2i_round(kBands14[b] / fpb);
    a->band14_n[b] = e - s + 1;
  }
  for (int b = 0; b < NB28; ++b) {                          /* SA.cpp:2026-2045 */
    int s = d2i_round((b == 0) ? (double)a->first_bin : kBands28[b - 1] / fpb);
    int e = d2i_round(kBands28[b] / fpb);
    if (e > N / 2) e = N / 2;
    if (s >= N / 2) { s = 0; e = 0; }
    a->band28_s[b] = s; a->band28_e[b] = e;
  }
  a->wh_decay = pow(0.001, (double)((float)H / (float)sr) / 22.0);
  a->env_coef = pow(0.01, (1000.0 / (8.0 * (double)sr)));
  a->silence_floor = (double)32768.0f * db_to_lin(-48.0);
}
static void analyser_free(analyser_t* a) { free(a->window); free(a->mel); }

/* ------------------------------------------------------------------------------------------- */
/* PCM conditioning (tail of TSampleAnalyser::LoadSample, SA.cpp:531-719)                       */

typedef struct {
  double* data; int len;     /* mData */
  int offset;                /* mDataOffset */
  float peak, rms;
} sample_t;

/* in: planar float32 channels in 16-bit range.  ch[0] is overwritten (as in the reference). */
static int condition(const analyser_t* a, float** ch, int nch, int n, int src_rate, sample_t* s)
{
  const float scale = 32768.0f;
  float* mono = ch[0];
  if (nch > 1) {                                            /* SA.cpp:535-548 */
    const float k = 1.0f / (float)nch;
    for (int i = 0; i < n; ++i) {
      for (int c = 1; c < nch; ++c) mono[i] += ch[c][i];
      mono[i] *= k;
    }
  }
  float* buf = mono; int owned = 0;
  const double speed = (double)src_rate / (double)a->sr;
  if (speed != 1.0) {                                       /* SA.cpp:563-607 */
    int nn = d2i_round(n / speed); if (nn < 1) nn = 1;
    float* r = (float*)calloc(nn, sizeof(float));
    rs_resample(mono, n, 1.0 / speed, r, nn);
    buf = r; n = nn; owned = 1;
  }
  double rms = 0.0;                                         /* SA.cpp:612-619 */
  for (int i = 0; i < n; ++i) { const double v = buf[i] / scale; rms += v * v; }
  rms = sqrt(rms / (1 * n));
  s->rms = (float)(rms < 1.0 ? rms : 1.0);
  float mn = buf[0], mx = buf[0];                           /* SA.cpp:624-637 */
  for (int i = 0; i < n; ++i) { if (buf[i] < mn) mn = buf[i]; if (buf[i] > mx) mx = buf[i]; }
  const double maxamp = (double)(fabsf(mn) > fabsf(mx) ? fabsf(mn) : fabsf(mx));
  { const double p = maxamp / scale; s->peak = (float)(p < 1.0 ? p : 1.0); }
  const double amp = (maxamp > kEps) ? scale / maxamp : 1.0;
  int lead = 0;                                             /* SA.cpp:651-669 */
  for (int f = 0; f < n; ++f, ++lead) if (fabs(amp * buf[f]) > a->silence_floor) break;
  int trail = 0;
  for (int f = n - 1; f > lead; --f, ++trail) if (fabs(amp * buf[f]) > a->silence_floor) break;
  const int audible = n - lead - trail;                     /* SA.cpp:681-718 */
  int end_off = 0, start_off = 0;
  if ((audible % a->N) < a->N / 2) end_off += a->N / 2;
  if (audible + end_off < a->N) start_off = a->N - audible - end_off;
  s->len = audible + start_off + end_off;
  s->data = (double*)calloc(s->len, sizeof(double));
  s->offset = -lead + start_off;
  const double fs = (double)amp / scale;
  for (int i = 0; i < audible; ++i) s->data[i + start_off] = buf[i + lead] * fs;
  if (owned) free(buf);
  return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* per-frame features                                                                          */

/* aubio mathutils.c:345-357, 606-615: mean square -> 10 log10 < threshold */
static int is_silent(const double* x, int n, double thr_db)
{
  double e = 0; for (int j = 0; j < n; ++j) e += x[j] * x[j];
  return (10. * log10(e / n)) < thr_db;
}

/* LibXtract scalar.c:624-636 (summed from the back) */
static double xt_rms(const double* x, int n)
{
  double r = 0; for (int i = n - 1; i >= 0; --i) r += x[i] * x[i];
  return sqrt(r / (double)n);
}

/* aubio mathutils.c:494-506 */
static double quad_peak_pos(const double* x, unsigned len, unsigned pos)
{
  if (pos == 0 || pos == len - 1) return pos;
  const double s0 = x[pos - 1], s1 = x[pos], s2 = x[pos + 1];
  return pos + .5 * (s0 - s2) / (s0 - 2. * s1 + s2);
}

/* aubio pitch.c:399-407, 450-462 + pitchyinfast.c:81-176 (yinfast, tol 0.75, silence -48 dB).
 * The reference gets r(tau) through three real FFTs; the restatement uses one complex FFT pair
 * with the same mathematical result: r(tau) = sum_{m<W} x[m] x[m+tau]. */
static void yin_f0(const analyser_t* a, const double* x, double* f0, double* conf)
{
  const int B = a->N, W = B / 2;
  double* sq = (double*)malloc(sizeof(double) * W);
  double* yin = (double*)malloc(sizeof(double) * W);
  double* re = (double*)calloc(B, sizeof(double)); double* im = (double*)calloc(B, sizeof(double));
  double* kr = (double*)calloc(B, sizeof(double)); double* ki = (double*)calloc(B, sizeof(double));
  /* running sums of squares (pitchyinfast.c:104-112) */
  double s0 = 0; for (int j = 0; j < W; ++j) s0 += x[j] * x[j];
  sq[0] = s0;
  for (int t = 1; t < W; ++t) { double v = sq[t - 1]; v -= x[t - 1] * x[t - 1]; v += x[W + t - 1] * x[W + t - 1]; sq[t] = v; }
  for (int t = 0; t < W; ++t) sq[t] += s0;
  /* cross-correlation via FFT: kernel = first W samples reversed, placed at 1..W */
  for (int j = 0; j < B; ++j) re[j] = x[j];
  for (int j = 0; j < W; ++j) kr[1 + j] = x[W - 1 - j];
  fft_c2c(re, im, B, -1); fft_c2c(kr, ki, B, -1);
  for (int j = 0; j < B; ++j) { const double pr = re[j] * kr[j] - im[j] * ki[j], pi = re[j] * ki[j] + im[j] * kr[j]; re[j] = pr; im[j] = pi; }
  fft_c2c(re, im, B, +1);
  /* QUIRK: aubio's Ooura back end scales the inverse rdft by 1/n (fft.c:462-476) although Ooura's
   * rdft(-1) returns (n/2) x, so the correlation it hands back is HALF the true r(tau).  The
   * reference's "difference function" is therefore sqdiff - r, not sqdiff - 2 r. */
  for (int t = 0; t < W; ++t) yin[t] = sq[t] - 2. * (0.5 * (re[t + W] / B));
  /* cumulative mean normalisation + first dip (pitchyinfast.c:150-167) */
  double period = 0; int found = 0; double tmp2 = 0;
  yin[0] = 1.;
  for (int t = 1; t < W; ++t) {
    tmp2 += yin[t];
    if (tmp2 != 0) yin[t] *= t / tmp2; else yin[t] = 1.;
    const int p = t - 3;
    if (t > 4 && yin[p] < 0.75 && yin[p] < yin[p + 1]) { period = quad_peak_pos(yin, W, p); found = 1; break; }
  }
  if (!found) {                                             /* mathutils.c:250-258: last minimum wins ties */
    unsigned pos = 0; double m = yin[0];
    for (int j = 0; j < W; ++j) { if (!(m < yin[j])) pos = j; m = (m < yin[j]) ? m : yin[j]; }
    period = quad_peak_pos(yin, W, pos);
  }
  unsigned peak_pos = 0;
  if (period == period && period >= 0) peak_pos = (unsigned)period;
  double pitch = (period > 0) ? a->sr / (period + 0.) : 0.;
  if (is_silent(x, B, -48.0)) pitch = 0.;
  *f0 = pitch;
  double c = (1. - yin[peak_pos]) / 0.25;                   /* SA.cpp:887-889 */
  *conf = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);
  free(sq); free(yin); free(re); free(im); free(kr); free(ki);
}

/* SA.cpp:2312-2398 + Autocorrelation.cpp:62-104 */
static double auto_correlation(const analyser_t* a, const double* x, int remaining)
{
  const int min_period = ms_to_samples(a->sr, 0.8f), seek_w = ms_to_samples(a->sr, 12.0f);
  const int max_seek = a->N / 2;
  const double* start = x;
  { const int lim = (remaining < max_seek ? remaining : max_seek) - 1;
    for (int i = 0; i < lim; ++i) if (x[i + 1] > x[i]) { start = x + i; remaining -= i; break; } }
  const int seek_off = remaining < min_period ? remaining : min_period;
  const double* end = start + seek_off;
  { const int r = remaining - seek_off; const int lim = (r < max_seek ? r : max_seek) - 1;
    for (int i = 0; i < lim; ++i) if (start[seek_off + i + 1] > start[seek_off + i]) { end = start + seek_off + i; break; } }
  const int period = (int)(end - start);
  if (!remaining || period >= remaining) return 0.0;
  const int width = remaining < seek_w ? remaining : seek_w;
  double* r = (double*)malloc(sizeof(double) * width);
  for (int i = 0; i < width; ++i) { double s = 0; for (int j = 0; j < width - i; ++j) s += start[j] * start[j + i]; r[i] = s; }
  if (r[0] != 0) for (int i = width - 1; i >= 0; --i) r[i] /= r[0];
  double best = 0.0;
  for (int i = period / 2; i < width; ++i) best = best > r[i] ? best : r[i];
  free(r);
  return best;
}

/* SA.cpp:129-133 */
static double flatness_db(const double* x, int n)
{
  const double v = lin_to_db(st_flatness(x, n)) / -60.0;
  return v < 1.0 ? v : 1.0;
}

/* ------------------------------------------------------------------------------------------- */
/* rhythm front end (AudioTypes/Source/OnsetDetector.cpp, RhythmTracker.cpp:46-117)             */

#define R_FFT 512
#define R_HOP 128
#define R_BINS 255

static float phase_rewrap(float p)   /* OnsetDetector.cpp:19-23 */
{
  const float pi = (float)kPi, twopi = (float)6.2831853071795864769252867665590,
    inv2pi = (float)0.15915494309189533576888376337251;
  return (p > -pi && p < pi) ? p : p + twopi * (1.f + floorf((-pi - p) * inv2pi));
}

typedef struct {
  int type;            /* 0 = rectified complex, 1 = power */
  float thresh, norm; int medspan, mingap, gapleft;
  float odfvals[128]; float other[R_BINS * 3];
  float post, postprev;
} onset_det_t;

static void onset_det_init(onset_det_t* d, int type, float sr, float thresh, float medspan_s, float mingap_s)
{
  memset(d, 0, sizeof(*d));
  d->type = type; d->thresh = thresh;
  d->medspan = (int)((sr * medspan_s) / (float)R_HOP + 0.5f);   /* OnsetDetector.cpp:280-282 */
  if (d->medspan < 3) d->medspan = 3;
  d->mingap = (int)((sr * mingap_s) / (float)R_HOP + 0.5f);     /* :290 */
  if (type == 1) d->norm = 2560.f / (float)((R_BINS + 2) * R_FFT);              /* :296 */
  else d->norm = (float)(231.70475 / pow((double)R_FFT, 1.5));                   /* :318 */
}

/* OnsetDetector.cpp:371-547 (kFunctionRComplex / kFunctionPower) + :551-590 */
static int onset_det_process(onset_det_t* d, const float* mag, const float* phase, float dc, float nyq)
{
  memmove(d->odfvals + 1, d->odfvals, (d->medspan - 1) * sizeof(float));
  if (d->type == 1) {
    float v = (nyq * nyq) + (dc * dc);
    for (int i = 0; i < R_BINS; ++i) { const float m = mag[i]; v += m * m; }
    d->odfvals[0] = v;
  } else {
    double total = 0.0;
    for (int i = 0; i < R_BINS; ++i) {
      const float cur = fabsf(mag[i]);
      const float pm = d->other[3 * i], yp = d->other[3 * i + 1], ypd = d->other[3 * i + 2];
      if (cur > 0.01f && !(cur < pm)) {
        const float pred = yp + ypd;
        float dev = pred - phase[i];
        dev = sqrtf(pm * pm + cur * cur - pm * cur * cosf(phase_rewrap(dev)));
        total += dev;
      }
    }
    for (int i = 0; i < R_BINS; ++i) {
      d->other[3 * i] = fabsf(mag[i]);
      float diff = phase[i] - d->other[3 * i + 1];
      d->other[3 * i + 1] = phase[i];
      d->other[3 * i + 2] = phase_rewrap(diff);
    }
    d->odfvals[0] = (float)total;
  }
  d->odfvals[0] *= d->norm;
  /* median removal + gap state machine */
  d->postprev = d->post;
  float sorted[128];
  memcpy(sorted, d->odfvals, d->medspan * sizeof(float));
  for (int i = 1; i < d->medspan; ++i) { float v = sorted[i]; int j = i - 1; while (j >= 0 && sorted[j] > v) { sorted[j + 1] = sorted[j]; --j; } sorted[j + 1] = v; }
  const float median = (d->medspan & 1) ? sorted[(d->medspan - 1) >> 1]
    : ((sorted[d->medspan >> 1] + sorted[(d->medspan >> 1) - 1]) * 0.5f);
  d->post = d->odfvals[0] - median;
  int detected;
  if (d->gapleft != 0) { d->gapleft--; detected = 0; }
  else { detected = (d->post > d->thresh) && (d->postprev <= d->thresh); if (detected) d->gapleft = d->mingap; }
  return detected;
}

/* per-file rhythm front end: fills onsets[2][Fr] */
static void rhythm_front(const analyser_t* a, const double* data, int L, int Fr, double* on_c, double* on_p)
{
  double win[R_FFT], re[R_FFT], im[R_FFT];
  for (int i = 0; i < R_FFT; ++i) win[i] = 0.5 * (1.0 - cos(6.2831853071795864769252867665590 * (double)i * (1.0 / (double)(R_FFT - 1))));
  double psp[R_BINS + 2]; memset(psp, 0, sizeof(psp));
  const float relax = (float)(exp((-2.30258509 * (float)R_HOP) / (25.0f * (float)a->sr)));
  const float wfloor = 0.1f;
  onset_det_t dc_, dp_;
  onset_det_init(&dc_, 0, (float)a->sr, 0.2f, 0.2f, 0.06f);
  onset_det_init(&dp_, 1, (float)a->sr, 0.8f, 0.2f, 0.12f);
  float mag[R_BINS], ph[R_BINS];
  (void)L;
  for (int t = 0; t < Fr; ++t) {
    const double* x = data + (size_t)t * R_HOP;
    for (int i = 0; i < R_FFT; ++i) { re[i] = win[i] * x[i]; im[i] = 0.0; }
    fft_c2c(re, im, R_FFT, +1);
    for (int i = 0; i < R_BINS; ++i) { mag[i] = (float)sqrt(re[i] * re[i] + im[i] * im[i]); ph[i] = (float)atan2(im[i], re[i]); }
    float dc = (float)re[0], nyq = (float)im[0];
    /* adaptive-max whitening (OnsetDetector.cpp:193-243) */
    { double v = fabsf(dc), o = psp[0]; if (v < o) v = v + (o - v) * relax; psp[0] = v; }
    { double v = fabsf(nyq), o = psp[1 + R_BINS]; if (v < o) v = v + (o - v) * relax; psp[1 + R_BINS] = v; }
    for (int i = 0; i < R_BINS; ++i) { double v = fabsf(mag[i]), o = psp[1 + i]; if (v < o) v = v + (o - v) * relax; psp[1 + i] = v; }
    dc /= (float)((double)wfloor > psp[0] ? (double)wfloor : psp[0]);
    nyq /= (float)((double)wfloor > psp[1 + R_BINS] ? (double)wfloor : psp[1 + R_BINS]);
    for (int i = 0; i < R_BINS; ++i) mag[i] /= (float)((double)wfloor > psp[1 + i] ? (double)wfloor : psp[1 + i]);
    on_c[t] = onset_det_process(&dc_, mag, ph, dc, nyq) ? (double)dc_.post : 0.0;
    on_p[t] = onset_det_process(&dp_, mag, ph, dc, nyq) ? (double)dp_.post : 0.0;
  }
}

/* ------------------------------------------------------------------------------------------- */
/* rhythm back end (RhythmTracker.cpp:121-660, CannyWindow.cpp, aubio beattracking.c)            */

/* CannyWindow.cpp:27-80: taps -12..+11 of n/s^2 exp(-n^2/(2 s^2)), then z-score rectified */
static void canny_sharpen(double* x, int n)
{
  const int L = 12; const double s2 = 16.0 * 16.0;
  double w[2 * 12 + 1];
  for (int i = -L; i < L + 1; ++i) w[i + L] = (double)i / s2 * exp(-1.0 * (i * i) / (2.0 * s2));
  double* t = (double*)malloc(sizeof(double) * n);
  for (int i = 0; i < n; ++i) {
    double sum = 0.0;
    for (int sh = -L; sh < L; ++sh) if (i + sh >= 0 && i + sh < n) sum += x[i + sh] * w[sh + L];
    t[i] = sum;
  }
  memcpy(x, t, sizeof(double) * n); free(t);
  const double mean = st_mean(x, n), var = st_variance(x, n, mean);
  if (var > 0.0) { const double sd = sqrt(var); for (int i = 0; i < n; ++i) { const double v = (x[i] - mean) / sd; x[i] = v > 0.0 ? v : 0.0; } }
}

/* RhythmTracker.cpp:623-659 */
static int calc_peaks(const double* o, int n, double* peaks)
{
  int cnt = 0;
  for (int f = 0; f < n; ++f) {
    if (o[f] <= 0.1) continue;
    int ok = 1;
    for (int i = -24; i <= 24; ++i) if (f + i >= 0 && f + i < n && o[f + i] > o[f]) { ok = 0; break; }
    if (ok) peaks[cnt++] = o[f];
  }
  return cnt;
}

/* aubio mathutils.c:508-517 */
static double quad_peak_mag(const double* x, unsigned len, double pos)
{
  if (pos >= len || pos < 0.) return 0.;
  const unsigned index = (unsigned)(pos - .5) + 1;
  if ((double)index == pos) return x[index];
  const double x0 = x[index - 1], x1 = x[index], x2 = x[index + 1];
  return x1 - .25 * (x0 - x2) * (pos - index);
}

/* One fresh aubio beat tracker run over the whole onset vector:
 * beattracking.c:59-110 (init), :126-262 (do; only the period branch influences bpm/confidence),
 * :286-410 (checkstate, first call), :424-441 (bpm, confidence); RhythmTracker.cpp:155-230. */
static double calc_tempo(const analyser_t* a, const double* sharp, int n, int onset_count, double* confidence)
{
  if (onset_count < 4) { *confidence = 0.0; return 0.0; }
  const unsigned winlen = n, laglen = winlen / 4;
  const double rayparam_d = 60. * a->sr / 120. / R_HOP;
  const unsigned rayparam = (unsigned)rayparam_d;
  double* acf = (double*)calloc(winlen, sizeof(double));
  double* acfout = (double*)calloc(laglen ? laglen : 1, sizeof(double));
  for (unsigned i = 0; i < winlen; ++i) {                   /* mathutils.c:652-666 */
    double t = 0.; for (unsigned j = i; j < winlen; ++j) t += sharp[j - i] * sharp[j];
    acf[i] = t / (double)(winlen - i);
  }
  if (laglen >= 2)
    for (unsigned i = 1; i < laglen - 1; ++i)
      for (unsigned aa = 1; aa <= 4; ++aa)
        for (unsigned b = 1; b < 2 * aa; ++b)
          acfout[i] += acf[i * aa + b - 1] * 1. / (2. * aa - 1.);
  for (unsigned i = 0; i < laglen; ++i)                      /* Rayleigh weighting (:104-107) */
    acfout[i] *= ((double)(i + 1.) / (rayparam_d * rayparam_d)) * exp((-((double)(i + 1.) * (double)(i + 1.)) / (2. * rayparam_d * rayparam_d)));
  unsigned maxi = 0; { double m = 0.0;                     /* mathutils.c:267-283: ties go to the last index */
    for (unsigned j = 0; j < laglen; ++j) { if (!(m > acfout[j])) maxi = j; m = (m > acfout[j]) ? m : acfout[j]; } }
  double rp;
  if (maxi > 0 && maxi < laglen - 1) rp = quad_peak_pos(acfout, laglen, maxi); else rp = rayparam;
  double bp = rp;                                           /* checkstate: initial state, gp = 0 */
  while (0 < bp && bp < 25) bp = bp * 2;
  double tempo = 0.0;
  if (bp != 0) tempo = 60. / ((R_HOP * bp) / (double)a->sr);
  double conf = 0.0;
  { double s = 0; for (unsigned j = 0; j < laglen; ++j) s += acfout[j]; if (s != 0.) conf = quad_peak_mag(acfout, laglen, rp) / s; }
  conf = conf * 16.0; conf = conf < 0.0 ? 0.0 : (conf > 1.0 ? 1.0 : conf);
  free(acf); free(acfout);
  if (tempo < 20.0 || tempo > 300.0) { *confidence = 0.0; return 0.0; }
  while (tempo < 80.0) tempo *= 2.0;
  while (tempo >= 200.0) tempo /= 2.0;
  *confidence = conf;
  return tempo;
}

/* RhythmTracker.cpp:502-555 */
static double guess_beats(double min_bpm, double max_bpm, double dur)
{
  (void)max_bpm;
  const double beat = 60.0 / min_bpm, bar = 4.0 * beat;
  if (dur < beat) return 0.0f;
  else if (dur < bar) {
    for (int div = 2; div >= 1; div /= 2) {
      const double d = ((4.0 / (double)div) * beat);
      if (4 % div == 0 && dur < d) return (float)(4.0 / (double)div);
    }
    return 4.0;
  } else {
    for (int bars = 1; bars <= 8; bars *= 2) { const double nb = 4.0 * bars; if (dur / nb < beat) return (float)nb; }
  }
  return 0.0f;
}

/* RhythmTracker.cpp:559-603 */
static double onset_match_conf(const analyser_t* a, const double* raw, int n, double off_s, double nbeats, double tempo, double thr)
{
  const int off = ms_to_samples(a->sr, (float)(off_s * 1000));
  const double spb = 60.0 / tempo * a->sr;
  const int range = (int)(spb / 32) / R_HOP;
  double strength = 0;
  for (int i = 0; i < nbeats * 2; ++i) {
    const int t = (int)(i * spb / 2.0) + off;
    const int idx = ((t + R_HOP / 2) / R_HOP);
    double peak = 0.0;
    for (int j = idx - range; j < idx + range; ++j) if (j >= 0 && j < n) peak = peak > raw[j] ? peak : raw[j];
    if (peak >= thr) strength += 1.0;
  }
  const double v = strength / (nbeats * 2) * 2.0;
  return v < 1.0 ? v : 1.0;
}

/* RhythmTracker.cpp:234-325 */
static double tempo_heuristics(const analyser_t* a, double* confidence, double tempo_in, double conf_in,
  double dur_s, double off_s, const double* sharp, const double* raw, int n, double thr)
{
  if (tempo_in == 0) { *confidence = 0.0; return 0.0; }
  double tempo = tempo_in; *confidence = conf_in;
  const double spb = 60.0 / tempo * a->sr;
  int last = n - 1;
  while (last > 0 && sharp[last] < 0.1) --last;
  const double ns = last * R_HOP;
  if (ns < spb * 3) { *confidence = 0.0; return 0.0; }
  const double nb = guess_beats(80, 180, dur_s);
  if (nb >= 4 && nb <= 16) {
    const double gbpm = nb / (dur_s / 60);
    const double delay = samples_to_ms(a->sr, R_HOP / 2) / 1000.0;
    const double gc = onset_match_conf(a, raw, n, off_s + delay, nb, gbpm, thr);
    if ((gc > 0.5) || (*confidence < 0.1 && gc > 0.1) || (*confidence < 0.5 && fabs(gbpm - tempo) < 10)) {
      tempo = gbpm; *confidence = 0.5 > gc ? 0.5 : gc;
    }
  }
  return tempo;
}

/* RhythmTracker.cpp:392-480 */
static double rhythm_contrast(const double* o, int n)
{
  double* sorted = (double*)malloc(sizeof(double) * n);
  memcpy(sorted, o, sizeof(double) * n); qsort(sorted, n, sizeof(double), cmp_double);
  const double thr = sorted[(int)(85.0 / 100.0 * (n - 1))];
  free(sorted);
  double psum = 0, vsum = 0; int pc = 0;
  int vpos = 0; double vval = thr;
  for (int i = 0; i < n; ++i) {
    if (o[i] < vval) { vpos = i; vval = o[i]; }
    if (o[i] < thr) continue;
    int ok = 1;
    for (int j = -24; j <= 24; ++j) if (i + j >= 0 && i + j < n && o[i + j] > o[i]) { ok = 0; break; }
    if (ok) { psum += o[i]; vsum += o[vpos]; pc++; vval = o[i]; }
  }
  const double total_mean = st_mean(o, n);
  /* TStatistics::Mean: n>=2 -> sum/n, n==1 -> the value, n==0 -> 0 */
  const double pmean = pc ? psum / pc : 0.0;
  const double vmean = (pc ? vsum / pc : 0.0) + 0.0001;
  if (pmean != 0.0) return -1.0 * pow(pmean / vmean, 1.0 / log(total_mean + 0.0001));
  return 0.0;
}

/* ------------------------------------------------------------------------------------------- */
/* the whole path                                                                              */

static long record_doubles(int F, int Fr)
{
  return AFX_N_HEADER + (long)AFX_N_FS_MAIN * F + 2L * Fr + (long)AFX_FV_BANDS * F + AFX_N_SERIES * AFX_N_STATS;
}

/* frame counts for a conditioned length (SA.cpp:760-764, 814, 991) */
static void frame_counts(const analyser_t* a, int len, int* F, int* Fr, int* L)
{
  const int cap = ms_to_samples(a->sr, 1000 * 20);
  *L = len < cap ? len : cap;
  *F = (*L >= a->N) ? (*L - a->N) / a->H + 1 : 0;
  *Fr = (*L >= R_FFT) ? (*L - R_FFT) / R_HOP + 1 : 0;
}

/*
 * pcm        : planar float32 [channels][nframes], 16-bit range (what the reference decoders hand to
 *              LoadSample); contents are not modified.
 * returns    : number of doubles written to out (AFXD record body: header, fs, fv, stats), or
 *              -1 bad arguments (SA.cpp:472-482), -2 out too small (needed count in *needed).
 */
long afxo_analyze(const float* pcm, int channels, int nframes, int src_rate,
                  int sample_rate, int fft_size, int hop_size,
                  int file_size, int bit_depth,
                  double* out, long out_cap, int* outF, int* outFr, long* needed)
{
  if (channels < 1 || channels > 8 || nframes <= 0) return -1;
  analyser_t A; analyser_init(&A, sample_rate, fft_size, hop_size);
  float** ch = (float**)malloc(sizeof(float*) * channels);
  for (int c = 0; c < channels; ++c) {
    ch[c] = (float*)malloc(sizeof(float) * nframes);
    memcpy(ch[c], pcm + (size_t)c * nframes, sizeof(float) * nframes);
  }
  sample_t S; condition(&A, ch, channels, nframes, src_rate, &S);
  for (int c = 0; c < channels; ++c) free(ch[c]);
  free(ch);

  int F, Fr, L; frame_counts(&A, S.len, &F, &Fr, &L);
  if (outF) *outF = F;
  if (outFr) *outFr = Fr;
  const long need = record_doubles(F, Fr);
  if (needed) *needed = need;
  if (need > out_cap) { free(S.data); analyser_free(&A); return -2; }
  memset(out, 0, sizeof(double) * need);

  double* H = out;
  double* fs[AFX_N_FS];
  { double* p = out + AFX_N_HEADER; for (int s = 0; s < AFX_N_FS; ++s) { fs[s] = p; p += (s < AFX_N_FS_MAIN) ? F : Fr; } }
  double* fv[7]; static const int fvn[7] = { 14, 14, 14, 14, 14, 28, 14 };
  { double* p = fs[AFX_N_FS - 1] + Fr; for (int v = 0; v < 7; ++v) { fv[v] = p; p += (long)F * fvn[v]; } }
  double* stats = fv[6] + (long)F * 14;

  /* ---- header (SA.cpp:734-754) ---- */
  H[0] = file_size;
  H[1] = samples_to_ms(src_rate, nframes) / 1000.0;
  H[2] = src_rate; H[3] = channels; H[4] = bit_depth;
  H[8] = samples_to_ms(A.sr, S.offset) / 1000.0;
  H[23] = S.peak; H[24] = S.rms; H[25] = S.offset; H[26] = S.len;
  { /* effective length (SA.cpp:1715-1756) */
    const double floors[3] = { db_to_lin(-48.0), db_to_lin(-24.0), db_to_lin(-12.0) };
    for (int s = 0; s < 3; ++s) {
      int lead = 0; for (int f = 0; f < S.len; ++f, ++lead) if (fabs(S.data[f]) > floors[s]) break;
      int trail = 0; for (int f = S.len - 1; f > lead; --f, ++trail) if (fabs(S.data[f]) > floors[s]) break;
      H[5 + s] = samples_to_ms(A.sr, S.len - lead - trail) / 1000.0;
    }
  }

  /* ---- main frame loop (SA.cpp:814-976) ---- */
  const int N = A.N, NB = N / 2;
  double* re = (double*)malloc(sizeof(double) * N); double* im = (double*)malloc(sizeof(double) * N);
  double* mag = (double*)calloc(NB, sizeof(double)); double* last = (double*)calloc(NB, sizeof(double));
  double* wh = (double*)calloc(NB + 1, sizeof(double)); double* whp = (double*)malloc(sizeof(double) * (NB + 1));
  char* ispeak = (char*)malloc(NB); int* pbins = (int*)malloc(sizeof(int) * (NB + 8)); double* pvals = (double*)malloc(sizeof(double) * (NB + 8));
  double* sorted = (double*)malloc(sizeof(double) * NB);
  for (int i = 0; i <= NB; ++i) whp[i] = 1.e-4;               /* awhitening.c:111-116 */
  const double scale = (double)(1.0f / (float)N);

  for (int fi = 0; fi < F; ++fi) {
    const int n = fi * A.H;
    const double* x = S.data + n;
    /* window, FFT (/N), magnitude (SA.cpp:826-846) */
    for (int i = 0; i < N; ++i) { re[i] = x[i] * A.window[i]; im[i] = 0.0; }
    fft_c2c(re, im, N, +1);
    for (int k = 0; k < NB; ++k) { const double r = re[k] * scale, q = im[k] * scale; mag[k] = sqrt(r * r + q * q); }
    /* whitening (awhitening.c:43-52) over NB+1 entries; entry NB of the spectrum is 0 */
    for (int i = 0; i <= NB; ++i) {
      const double v = (i < NB) ? mag[i] : 0.0;
      double t = A.wh_decay * whp[i]; if (t < 1.e-4) t = 1.e-4;
      whp[i] = v > t ? v : t;
      wh[i] = v / whp[i];
    }
    /* peak spectrum (SA.cpp:95-123) */
    double wmax = wh[0]; for (int i = 1; i < NB; ++i) wmax = wh[i] > wmax ? wh[i] : wmax;
    const int npk = st_peaks(wh, NB, 0.25 * wmax, pbins, pvals);
    memset(ispeak, 0, NB);
    for (int i = 0; i < npk; ++i) if (pvals[i] != 0.0) ispeak[pbins[i]] = 1;
    /* silence + amplitude (SA.cpp:865-873, 1760-1804) */
    const int silent = is_silent(x, A.H, -48.0);
    fs[0][fi] = silent ? 1.0 : 0.0;
    { double mn = x[0], mx = x[0]; for (int i = 0; i < A.H; ++i) { mn = x[i] < mn ? x[i] : mn; mx = x[i] > mx ? x[i] : mx; }
      fs[1][fi] = fabs(mn) > fabs(mx) ? fabs(mn) : fabs(mx); }
    { const double r = xt_rms(x, A.H); fs[2][fi] = (r != r) ? 0.0 : r; }
    { double env = 0.0, fmax = 0.0; for (int i = 0; i < A.H; ++i) { const double in = fabs(x[i]); env = in + A.env_coef * (env - in); fmax = fmax > env ? fmax : env; }
      fs[3][fi] = fmax; }
    /* F0 (SA.cpp:876-917) */
    double f0, conf, fsafe = 0.0;
    yin_f0(&A, x, &f0, &conf);
    fs[15][fi] = f0; fs[16][fi] = conf;
    if (f0 > 0.0 && conf > 0.2) fsafe = f0;
    else if (!silent) { const double c = st_centroid(mag, NB); fsafe = (double)A.sr / (double)N * (c > 0.0 ? c : 0.0); }
    fs[17][fi] = fsafe;
    /* harmonic spectrum / inharmonicity / tristimulus: the "frequency" half handed to LibXtract is all
     * zeros (SA.cpp:111-116, 844-846), so vector.c:545-577 and scalar.c:302-415, 638-661 yield 0. */
    fs[11][fi] = 0.0; fs[18][fi] = 0.0; fs[19][fi] = 0.0; fs[20][fi] = 0.0;
    if (fi == 0) memcpy(last, mag, sizeof(double) * NB);
    /* autocorrelation over the rest of the file (SA.cpp:943-944) */
    fs[21][fi] = auto_correlation(&A, x, S.len - n);
    /* spectral scalars on the analysis window (SA.cpp:1808-1947) */
    const double* m = mag + A.first_bin; const double* lm = last + A.first_bin; const int nb = A.nbins;
    { const double r = xt_rms(m, nb); fs[4][fi] = (r != r) ? 0.0 : r; }
    const double cen = st_centroid(m, nb), spr = st_spread(m, nb, cen);
    fs[5][fi] = cen; fs[7][fi] = spr;
    fs[8][fi] = st_skewness(m, nb, cen, spr); fs[9][fi] = st_kurtosis(m, nb, cen, spr);
    { /* LibXtract scalar.c:472-493 */
      double pivot = 0, t = 0; int k;
      for (k = nb - 1; k >= 0; --k) pivot += m[k];
      pivot *= 85.0 / 100.0;
      for (k = 0; t < pivot; k++) t += m[k];
      const double r = k * (double)(A.sr / (N / 2));
      fs[6][fi] = (r != r) ? 0.0 : r; }
    { const double f = flatness_db(m, nb); fs[10][fi] = (f != f) ? 0.0 : f; }
    fs[14][fi] = st_correlation(m, lm, nb);
    { int c = 0; for (int k = 0; k < nb; ++k) c += ispeak[A.first_bin + k]; fs[12][fi] = c; }
    /* 14 sub-band features (SA.cpp:2067-2260) */
    {
      memcpy(sorted, mag, sizeof(double) * NB);
      int cur = A.first_bin; double csum = 0.0;
      for (int b = 0; b < NB14; ++b) {
        int nbin = A.band14_n[b]; if (nbin > NB - cur) nbin = NB - cur;
        const double bmean = st_mean(mag + cur, nbin);
        double r = 0.0; for (int i = 0; i < nbin; ++i) r += mag[cur + i] * mag[cur + i];
        r = sqrt(r / (double)nbin);
        const double fl = flatness_db(mag + cur, nbin);
        const double fx = st_correlation(mag + cur, last + cur, nbin);
        double thr = 0.0; for (int i = 0; i < nbin; ++i) thr = thr > mag[cur + i] ? thr : mag[cur + i];
        thr *= 0.25;
        double cplx = 0;
        if (thr > 0.0) for (int i = 0; i < nbin; ++i) { const int q = cur + i;
          if (mag[q] > thr && q > 0 && q < NB - 1 && mag[q] > mag[q - 1] && mag[q] > mag[q + 1]) ++cplx; }
        qsort(sorted + cur, nbin, sizeof(double), cmp_double);
        int nei = (int)(0.3 * nbin); if (nei < 1) nei = 1;
        double sum = 0; for (int i = 0; i < nei && i < nbin; ++i) sum += sorted[cur + i];
        const double valley = sum / nei + 1e-30;
        sum = 0; for (int i = nbin; i > nbin - nei; --i) sum += sorted[cur + i - 1];
        const double peak = sum / nei + 1e-30;
        const double contrast = -1.0 * pow(peak / valley, 1.0 / log(bmean + 1e-30));
        fv[0][fi * 14 + b] = r; fv[1][fi * 14 + b] = fl; fv[2][fi * 14 + b] = fx;
        fv[3][fi * 14 + b] = cplx; fv[4][fi * 14 + b] = contrast;
        csum += contrast; cur += nbin;
      }
      fs[13][fi] = csum / NB14;
    }
    /* 28 frequency bands (SA.cpp:2007-2048) */
    for (int b = 0; b < NB28; ++b) { double s = 0.0; for (int k = A.band28_s[b]; k < A.band28_e[b]; ++k) s += mag[k] * mag[k]; fv[5][fi * 28 + b] = s; }
    /* MFCC (LibXtract vector.c:350-391) */
    { double lg[NCEP];
      for (int f = 0; f < NCEP; ++f) { double e = 0.0; const double* row = A.mel + (size_t)f * NB; for (int k = 0; k < NB; ++k) e += mag[k] * row[k];
        lg[f] = log(e < 2e-42 ? 2e-42 : e); }
      for (int q = 0; q < NCEP; ++q) { double t = 0; for (int mm = 1; mm <= NCEP; ++mm) t += lg[mm - 1] * cos(M_PI * (q / (double)NCEP) * (mm - 0.5)); fv[6][fi * 14 + q] = t; } }
    memcpy(last, mag, sizeof(double) * NB);
  }
  free(re); free(im); free(mag); free(last); free(wh); free(whp); free(ispeak); free(pbins); free(pvals); free(sorted);

  /* ---- rhythm (SA.cpp:983-1048) ---- */
  if (Fr > 0) {
    rhythm_front(&A, S.data, L, Fr, fs[22], fs[23]);
    const double dur_s = (double)(samples_to_ms(src_rate, nframes) / 1000);
    const double off_s = (double)(samples_to_ms(src_rate, S.offset) / 1000);
    double tempo[2], tconf[2];
    double* sharp[2];
    for (int t = 0; t < 2; ++t) {
      const double* raw = fs[22 + t]; const double thr = t == 0 ? 0.2 : 0.8;
      double* Hh = H + 9 + 6 * t;
      int count = 0; for (int i = 0; i < Fr; ++i) if (raw[i] > thr) ++count;
      sharp[t] = (double*)malloc(sizeof(double) * Fr); memcpy(sharp[t], raw, sizeof(double) * Fr);
      canny_sharpen(sharp[t], Fr);
      double* peaks = (double*)malloc(sizeof(double) * Fr); const int np = calc_peaks(sharp[t], Fr, peaks);
      Hh[0] = count;
      Hh[1] = rhythm_contrast(sharp[t], Fr);
      Hh[2] = (double)np / (double)Fr * (double)R_HOP / (double)R_FFT;
      if (np) { const double pm = st_mean(peaks, np) / 4.0; Hh[3] = pm < 0.0 ? 0.0 : (pm > 1.0 ? 1.0 : pm); } else Hh[3] = 0.0;
      tempo[t] = calc_tempo(&A, sharp[t], Fr, count, &tconf[t]);
      Hh[4] = tempo[t]; Hh[5] = tconf[t];
      free(peaks);
    }
    const int w = (tconf[1] > tconf[0]) ? 1 : 0;
    double fc = 0.0;
    H[21] = tempo_heuristics(&A, &fc, tempo[w], tconf[w], dur_s, off_s, sharp[w], fs[22 + w], Fr, w == 0 ? 0.2 : 0.8);
    H[22] = fc;
    free(sharp[0]); free(sharp[1]);
  }

  /* ---- statistics (SA.cpp:2402-2412, SampleDescriptors.h:212-230, 327-355) ---- */
  for (int s = 0; s < AFX_N_FS; ++s) st_calc13(fs[s], (s < AFX_N_FS_MAIN) ? F : Fr, stats + (long)s * AFX_N_STATS);
  { long so = AFX_N_FS; double* col = (double*)malloc(sizeof(double) * (F > 0 ? F : 1));
    for (int v = 0; v < 7; ++v) for (int b = 0; b < fvn[v]; ++b, ++so) {
      for (int i = 0; i < F; ++i) col[i] = fv[v][(long)i * fvn[v] + b];
      st_calc13(col, F, stats + so * AFX_N_STATS);
    }
    free(col); }

  free(S.data); analyser_free(&A);
  return need;
}

/* constants exported for the CUDA library's table tests */
void afxo_tables(int sample_rate, int fft_size, int hop_size, double* window, double* mel, int* band14_n, int* band28_se)
{
  analyser_t A; analyser_init(&A, sample_rate, fft_size, hop_size);
  if (window) memcpy(window, A.window, sizeof(double) * fft_size);
  if (mel) memcpy(mel, A.mel, sizeof(double) * NCEP * (fft_size / 2));
  if (band14_n) memcpy(band14_n, A.band14_n, sizeof(int) * NB14);
  if (band28_se) for (int b = 0; b < NB28; ++b) { band28_se[2 * b] = A.band28_s[b]; band28_se[2 * b + 1] = A.band28_e[b]; }
  analyser_free(&A);
}

/* conditioning only (for K1 parity): returns conditioned length; data may be NULL to query */
int afxo_condition(const float* pcm, int channels, int nframes, int src_rate, int sample_rate, int fft_size,
                   double* data, int data_cap, int* offset, float* peak, float* rms)
{
  if (channels < 1 || channels > 8 || nframes <= 0) return -1;
  analyser_t A; analyser_init(&A, sample_rate, fft_size, fft_size / 2);
  float** ch = (float**)malloc(sizeof(float*) * channels);
  for (int c = 0; c < channels; ++c) { ch[c] = (float*)malloc(sizeof(float) * nframes); memcpy(ch[c], pcm + (size_t)c * nframes, sizeof(float) * nframes); }
  sample_t S; condition(&A, ch, channels, nframes, src_rate, &S);
  for (int c = 0; c < channels; ++c) free(ch[c]);
  free(ch);
  if (offset) *offset = S.offset;
  if (peak) *peak = S.peak;
  if (rms) *rms = S.rms;
  const int len = S.len;
  if (data && data_cap >= len) memcpy(data, S.data, sizeof(double) * len);
  free(S.data); analyser_free(&A);
  return len;
}

/* single-function exports used by the KAT tests (TestStatistics.cpp:16-113) */
double afxo_sum(const double* x, int n) { return st_sum(x, n); }
double afxo_mean(const double* x, int n) { return st_mean(x, n); }
double afxo_median(const double* x, int n) { return st_median(x, n); }
double afxo_gmean(const double* x, int n) { return st_gmean(x, n); }
double afxo_min(const double* x, int n) { double m = n > 0 ? x[0] : 0.0; for (int i = 1; i < n; ++i) m = x[i] < m ? x[i] : m; return m; }
double afxo_max(const double* x, int n) { double m = n > 0 ? x[0] : 0.0; for (int i = 1; i < n; ++i) m = x[i] > m ? x[i] : m; return m; }

/* =========================================================================================== */
/* High-level derivations that need no classification model (SampleAnalyser.cpp:1232-1606) and   */
/* the classification feature vector (SampleClassificationDescriptors.cpp:330-560), from one     */
/* low-level record laid out as afxo_analyze writes it (header[32], 22 x fs[F], 2 x fs[Fr],      */
/* 7 x fv[F][nb], stats[136][13]).                                                               */
/*   hl[16]: base_note, base_note_confidence, peak_db, rms_db, bpm, bpm_confidence, brightness,  */
/*           noisiness, harmonicity, spectral_flatness, spectral_flux, spectral_complexity,      */
/*           spectral_contrast, spectral_inharmonicity, pitch_confidence, 0                      */
/*   pitch[F] (MIDI notes), signature[64][14], features[AFXO_NFEAT]                              */
/*   pad[21]: the reference pads short files with the LAST frame of a silent 0.5-s sample        */
/*           analysed at hop 1024 (SampleClassificationDescriptors.cpp:330-368): frequency_bands */
/*           [0..13] of that frame, then spectral_rms, spectral_flatness, spectral_flux,         */
/*           spectral_contrast, spectral_complexity, f0_confidence, amplitude_rms                */
/* Returns 0, or -1 when a feature is NaN / Inf (the reference throws: the file fails to analyse) */

#define HL_NSIG_FRAMES 64
#define HL_NSIG_BANDS 14
#define HL_NTIME 48
#define AFXO_NFEAT 1680

static const int kHlBands[HL_NSIG_BANDS] = { 0, 1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25 };   /* SampleAnalyser.cpp:1462-1465 */
static const int kHlTimeSeries[HL_NTIME] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,
  25,26,27,28,29,30,31,32,33,34,35,36,37,38,39,40,41,42,43, 64,128,256,512 };                    /* SampleClassificationDescriptors.cpp:39-43 */

static double hl_freqtomidi(double freq)      /* aubio mathutils.c:535-546 (smpl_t = double) */
{
  if (freq < 2. || freq > 100000.) return 0.;
  double midi = freq / 6.875;
  midi = log(midi) / 0.69314718055995;
  midi *= 12;
  midi -= 3;
  return midi;
}
static double hl_lin_to_db_f(float v)          /* TAudioMath::LinToDb(float), AudioMath.inl:38-54 */
{
  if (v == 1.0f) return 0.0f;
  if (v > 1e-12f) return (float)(log((double)v) * (20.0 / log(10.0)));
  return -200.0f;
}
static double hl_cubic(double ym1, double y0, double y1, double y2, double pos)   /* SampleAnalyser.cpp:139-155 */
{
  const double x = pos - floor(pos), xx = x * x, xxx = xx * x;
  const double a = -0.5 * xxx + xx - 0.5 * x, b = 1.5 * xxx - 2.5 * xx + 1.0, c = -1.5 * xxx + 2.0 * xx + 0.5 * x, d = 0.5 * xxx - 0.5 * xx;
  return a * ym1 + b * y0 + c * y1 + d * y2;
}
static double hl_merged_band(const double* bands28_frame, int b)                 /* SampleAnalyser.cpp:1476-1489 */
{
  const int s = (b - 1) >= 0 ? kHlBands[b - 1] + 1 : 0, e = kHlBands[b];
  double m = 0.0;
  for (int sb = s; sb <= e; ++sb) m += bands28_frame[sb];
  m /= (double)(e - s + 1);
  return pow(m * 1.25, 1.0 / 6.0);
}
static int hl_push(double* f, int* n, double v) { f[(*n)++] = v; return (isnan(v) || isinf(v)) ? 1 : 0; }

int afxo_highlevel(const double* rec, int F, int Fr, int sample_rate, float peak_value, float rms_value, const double* pad,
                   double* hl, double* pitch, double* signature, double* features)
{
  const double* header = rec;
  const double* fs[24];
  const double* p = rec + 32;
  for (int s = 0; s < 24; ++s) { fs[s] = p; p += (s < 22) ? F : Fr; }
  static const int nbv[7] = { 14, 14, 14, 14, 14, 28, 14 };
  const double* fv[7];
  for (int v = 0; v < 7; ++v) { fv[v] = p; p += (size_t)F * nbv[v]; }
  const double* stats = p;
  enum { S_SIL = 0, S_PEAK = 1, S_ARMS = 2, S_RMS = 4, S_CENT = 5, S_ROLL = 6, S_FLAT = 10, S_INH = 11, S_CPLX = 12, S_CONTR = 13, S_FLUX = 14,
         S_F0 = 15, S_CONF = 16, S_AC = 21 };
  double* tmp = (double*)malloc(sizeof(double) * (F + 1) * 2);
  double* tmp2 = tmp + F + 1;
  memset(hl, 0, sizeof(double) * 16);

  /* audible frames = !IsSilentFrame (SampleAnalyser.cpp:867) */
  int na = 0;
  #define AUDIBLE(i) (fs[S_SIL][i] == 0.0)
  #define GATHER(series, out, n) do { n = 0; for (int i_ = 0; i_ < F; ++i_) if (AUDIBLE(i_)) out[n++] = fs[series][i_]; } while (0)
  GATHER(S_CONF, tmp, na);
  const double apcm = na ? st_mean(tmp, na) : 0.0;                              /* :1256-1260 */
  const double thr = (apcm >= 0.8) ? 0.8 : (apcm >= 0.5) ? 0.5 : 0.2;           /* :1262-1277 */
  const double fmax = (double)(sample_rate / 4);
  #define CONFIDENT(i) (fs[S_CONF][i] > thr && fs[S_F0][i] > 20 && fs[S_F0][i] < fmax)

  /* base note (:1279-1330) */
  double base_note = -1.0;
  int nc = 0;
  for (int i = 0; i < F; ++i) if (CONFIDENT(i)) tmp[nc++] = fs[S_F0][i];
  if (nc) { const double hz = st_median(tmp, nc); if (hz > 20 && hz < fmax) base_note = hl_freqtomidi(hz); }
  hl[0] = base_note;
  if (base_note > 0.0) {
    for (int i = 0; i < nc; ++i) tmp2[i] = fabs(base_note - hl_freqtomidi(tmp[i]));
    const double sd = sqrt(st_variance(tmp2, nc, st_mean(tmp2, nc)));
    const double q = sd / 6.0;
    hl[1] = apcm * (1.0 - (q < 1.0 ? q : 1.0));
  }
  hl[2] = hl_lin_to_db_f(peak_value); hl[3] = hl_lin_to_db_f(rms_value);        /* :1336-1339 */
  { double v = header[21];                                                      /* bpm: :1345-1349, TMath::Quantize(0.5, kRoundToNearest) */
    if (v > 0.0) v += 0.25; else v -= 0.25;
    v = (double)((double)(int)(v / 0.5) * 0.5);
    hl[4] = v; hl[5] = header[22]; }
  /* brightness (:1355-1383) */
  { int n1, n2; GATHER(S_ROLL, tmp, n1); GATHER(S_CENT, tmp2, n2);
    if (n1 && n2) {
      const double rm = st_mean(tmp, n1); double cm = tmp2[0]; for (int i = 1; i < n2; ++i) cm = tmp2[i] > cm ? tmp2[i] : cm;
      double w = hl_freqtomidi(rm) / 128.0 * 0.7 + hl_freqtomidi(cm) / 128.0 * 0.3;
      w = w < 1.0 ? w : 1.0; w = w > 0.0 ? w : 0.0;
      hl[6] = pow(w, 4.0);
    } }
  /* noisiness (:1387-1414), harmonicity (:1419-1446) */
  { int n; GATHER(S_FLAT, tmp, n);
    double flat_mean = 0.0;
    if (n) {
      double mn = tmp[0], mx = tmp[0]; for (int i = 1; i < n; ++i) { mn = tmp[i] < mn ? tmp[i] : mn; mx = tmp[i] > mx ? tmp[i] : mx; }
      flat_mean = st_mean(tmp, n);
      double w = (1.0 - mn) * 0.2 + (1.0 - flat_mean) * 0.6 + (1.0 - mx) * 0.2;
      w = w < 1.0 ? w : 1.0; w = w > 0.0 ? w : 0.0;
      hl[7] = pow(w, 2.0);
    }
    hl[9] = n ? flat_mean : 0.0;                                                /* :1538-1539 */
    int m; GATHER(S_AC, tmp, m);
    if (m) {
      const double acm = st_mean(tmp, m);
      const double a = 1.5 * acm, b = 2.0 * apcm;
      double w = (a < 1.0 ? a : 1.0) * 0.4 + (b < 1.0 ? b : 1.0) * 0.3 + flat_mean * 0.3;
      w = w < 1.0 ? w : 1.0; w = w > 0.0 ? w : 0.0;
      hl[8] = pow(w, 2.0);
    } }
  { int n;                                                                      /* :1527-1554 */
    GATHER(S_FLUX, tmp, n); hl[10] = n ? st_mean(tmp, n) : 0.0;
    GATHER(S_CPLX, tmp, n); hl[11] = n ? st_mean(tmp, n) : 0.0;
    GATHER(S_CONTR, tmp, n); hl[12] = n ? st_mean(tmp, n) : 0.0;
    GATHER(S_INH, tmp, n); hl[13] = n ? st_mean(tmp, n) : 0.0; }
  hl[14] = apcm;                                                                /* :1604 */

  /* spectrum signature (:1449-1521) */
  { double* scaled = (double*)malloc(sizeof(double) * (size_t)F * HL_NSIG_BANDS);
    for (int f = 0; f < F; ++f) for (int b = 0; b < HL_NSIG_BANDS; ++b) scaled[f * HL_NSIG_BANDS + b] = hl_merged_band(fv[5] + (size_t)f * 28, b);
    const double step = (double)F / HL_NSIG_FRAMES;
    double pos = 0;
    for (int i = 0; i < HL_NSIG_FRAMES; ++i) {
      const int ipos = (int)pos, im1 = ipos - 1 > 0 ? ipos - 1 : 0, i1 = ipos + 1 < F - 1 ? ipos + 1 : F - 1, i2 = ipos + 2 < F - 1 ? ipos + 2 : F - 1;
      for (int j = 0; j < HL_NSIG_BANDS; ++j)
        signature[i * HL_NSIG_BANDS + j] = hl_cubic(scaled[im1 * HL_NSIG_BANDS + j], scaled[ipos * HL_NSIG_BANDS + j],
                                                   scaled[i1 * HL_NSIG_BANDS + j], scaled[i2 * HL_NSIG_BANDS + j], pos);
      pos += step;
    }
    free(scaled); }

  /* pitch (:1558-1598) */
  { double last = 0.0;
    if (F > 1) { const int lim = (F / 4 > 1) ? F / 4 : 1; for (int i = 0; i <= lim; ++i) if (AUDIBLE(i) && CONFIDENT(i)) { last = fs[S_F0][i]; break; } }
    for (int i = 0; i < F; ++i) {
      if (AUDIBLE(i) && CONFIDENT(i)) last = fs[S_F0][i];
      pitch[i] = hl_freqtomidi(last);
    } }

  /* classification features (SampleClassificationDescriptors.cpp:404-560) */
  int n = 0, bad = 0;
  for (int b = 0; b < HL_NSIG_BANDS; ++b) for (int i = 0; i < HL_NTIME; ++i) {
    const int tf = kHlTimeSeries[i];
    bad |= hl_push(features, &n, tf < F ? hl_merged_band(fv[5] + (size_t)tf * 28, b) : pad[b]);
  }
  static const int tser[6] = { S_RMS, S_FLAT, S_FLUX, S_CONTR, S_CPLX, S_CONF };
  static const int pick[7] = { 0, 1, 3, 5, 10, 11, 12 };     /* min, max, mean, variance, flatness, dmean, dvariance */
  for (int k = 0; k < 6; ++k) for (int i = 0; i < HL_NTIME; ++i) { const int tf = kHlTimeSeries[i]; bad |= hl_push(features, &n, tf < F ? fs[tser[k]][tf] : pad[14 + k]); }
  for (int k = 0; k < 6; ++k) for (int q = 0; q < 7; ++q) bad |= hl_push(features, &n, stats[tser[k] * 13 + pick[q]]);
  static const int vbase[6] = { 24, 38, 52, 66, 80, 122 };   /* rms, flatness, flux, complexity, contrast sub-bands; cepstrum */
  for (int k = 0; k < 6; ++k) for (int b = 0; b < 14; ++b) for (int q = 0; q < 7; ++q) bad |= hl_push(features, &n, stats[(vbase[k] + b) * 13 + pick[q]]);
  for (int i = 0; i < HL_NTIME; ++i) { const int tf = kHlTimeSeries[i]; bad |= hl_push(features, &n, tf < F ? fs[S_ARMS][tf] : pad[20]); }
  for (int q = 0; q < 7; ++q) bad |= hl_push(features, &n, stats[S_ARMS * 13 + pick[q]]);
  for (int q = 0; q < 7; ++q) bad |= hl_push(features, &n, stats[S_SIL * 13 + pick[q]]);
  bad |= hl_push(features, &n, header[14]); bad |= hl_push(features, &n, header[20]);   /* tempo confidences */
  bad |= hl_push(features, &n, header[10]); bad |= hl_push(features, &n, header[16]);   /* onset contrasts */
  bad |= hl_push(features, &n, header[12]); bad |= hl_push(features, &n, header[18]);   /* onset strengths */
  bad |= hl_push(features, &n, header[7]);                                              /* effectve_length_12dB */
  while (n % HL_NTIME) bad |= hl_push(features, &n, stats[S_RMS * 13 + 3]);             /* padding: spectral_rms mean */
  free(tmp);
  return (n != AFXO_NFEAT) ? -2 : (bad ? -1 : 0);
  #undef AUDIBLE
  #undef GATHER
  #undef CONFIDENT
}

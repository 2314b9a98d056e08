// Shared device-side types and helpers of libafec_b200 (sm_100a).
//
// Data layout in HBM for one batch (see DESIGN.md "Data layout"):
//   pcm      raw interleaved PCM of every file, packed in submission order
//   mono     float32 mono signal at the analysis rate (after downmix / resample), per file
//   files    AfxFile[n_files]       host-built descriptor table
//   state    AfxState[n_files]      device-computed conditioning result
//   mag      double [TF][1024]      magnitude spectra of all main frames (TF = sum of frame caps)
//   fs       double [22][TF]        framed scalars on the main grid (series-major -> coalesced)
//   fsr      double [2][TFr]        onset series on the rhythm grid
//   fv       double 7 x [TF][nb]    framed vectors (5 x 14 sub-bands, 28 bands, 14 cepstrum), one array each
//   stats    double [n_files][136][13]
//   header   double [n_files][32]
// The conditioned signal mData (SampleAnalyser.cpp:698-718) is never materialised: frames read
// the float32 mono buffer through mdata() which applies trim, padding and normalisation.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define AFX_NFFT 2048
#define AFX_NBIN 1024
#define AFX_RFFT 512
#define AFX_RHOP 128
#define AFX_RBINS 255
#define AFX_RROW 512        // floats per rhythm frame row: mag[0..255], dc, nyq, pad | phase[256..511]
#define AFX_FV_STRIDE 112
#define AFX_N_BLOBS 122      // BLOB columns of an afec-ll.db row: 24 VR + 7 x (1 VVR + 13 VR)
#define AFX_N_HL 16
#define AFX_HL_SIGNATURE (64 * 14)
#define AFX_HL_FEATURES 1680
// the analysis window of the spectral statistics: bins round(20 / 21) .. round(15500 / 21) with the reference's integer
// FrequenciesPerBin = 44100 / 2048 = 21 (SampleAnalyser.cpp:171-175).  afx_create only accepts that rate / size and checks
// that its table agrees, so the frame kernels may treat the bounds as compile-time constants.
#define AFX_WIN_FIRST 1
#define AFX_WIN_BINS 738

// fs series ids (order of afec_b200/layout.py FRAMED_SCALARS)
enum {
  FS_AMP_SILENCE = 0, FS_AMP_PEAK, FS_AMP_RMS, FS_AMP_ENV,
  FS_SPEC_RMS, FS_SPEC_CENTROID, FS_SPEC_ROLLOFF, FS_SPEC_SPREAD, FS_SPEC_SKEW, FS_SPEC_KURT,
  FS_SPEC_FLATNESS, FS_SPEC_INHARM, FS_SPEC_COMPLEXITY, FS_SPEC_CONTRAST, FS_SPEC_FLUX,
  FS_F0, FS_F0_CONF, FS_F0_FAILSAFE, FS_TRISTIM1, FS_TRISTIM2, FS_TRISTIM3, FS_AUTOCORR,
  FS_ONSETS_COMPLEX, FS_ONSETS_PERC
};
// fv: seven arrays [TF][nbands]; array v starts at fv + FV_x * TF (FV_x = cumulative band count)
enum { FV_RMS = 0, FV_FLATNESS = 14, FV_FLUX = 28, FV_COMPLEXITY = 42, FV_CONTRAST = 56, FV_BANDS28 = 70, FV_CEPSTRUM = 98 };
// header slots
enum {
  H_FILE_SIZE = 0, H_FILE_LENGTH, H_FILE_RATE, H_FILE_CHANNELS, H_FILE_BITS,
  H_EFF48, H_EFF24, H_EFF12, H_ANALYZATION_OFFSET,
  H_RC_COUNT, H_RC_CONTRAST, H_RC_FREQ, H_RC_STRENGTH, H_RC_TEMPO, H_RC_TEMPO_CONF,
  H_RP_COUNT, H_RP_CONTRAST, H_RP_FREQ, H_RP_STRENGTH, H_RP_TEMPO, H_RP_TEMPO_CONF,
  H_FINAL_TEMPO, H_FINAL_TEMPO_CONF, H_PEAK, H_RMS, H_DATA_OFFSET, H_DATA_LEN
};

struct AfxFile {            // host-built, one per file
  long long pcm_off;        // byte offset of the file's interleaved PCM in the device pcm buffer
  long long mono_off;       // sample offset into the mono buffer
  long long src_off;        // sample offset into the pre-resample mono buffer (resampled files only)
  int nframes_src;          // frames per channel as decoded
  int n;                    // samples at the analysis rate (== nframes_src unless resampled)
  int channels, src_rate, format, bit_depth;
  long long file_size;
  int frame_off, frame_cap; // main frame slots [frame_off, frame_off + frame_cap)
  int rframe_off, rframe_cap;
  int status;               // AFX_FILE_*
  int inject;               // >= 0: conditioning reductions come from AfxBatchDev::inject[inject] (long files conditioned in parts)
  int src_end, dst_end;     // the conditioning passes stop here (== nframes_src / n, or the end of one part's range)
};

struct AfxInject {          // conditioning reductions of a whole file made elsewhere (afx_part.cu)
  unsigned int maxabs_bits;
  int first, last;
  int eff_first[3], eff_last[3];
  int pad;
  double sumsq;
};

struct AfxState {           // device-computed, one per file (SampleAnalyser.cpp:610-718)
  unsigned int maxabs_bits; // max |x| as float bits (non-negative floats order like ints)
  int first, last;          // first / last sample above the -48 dB floor (n / -1 when none)
  int eff_first[3], eff_last[3];
  double sumsq;             // sum (x/32768)^2
  double fs;                // FinalScaling = Amplification / 32768
  double amp;
  int lead, audible, start_off, len;
  int L, F, Fr, data_offset;
  // batch path: the effective-length scan rides in the trim pass (raw mono indices); k_layout maps them into the conditioned
  // signal, and asks for an exact rescan of the audible span in the (rounding-boundary) case that one lies outside it
  int effraw_first[3], effraw_last[3];
  int eff_rescan, pad_;
  // smallest float32 magnitudes that pass the trim floor / the three effective-length floors (k_amp): the scans compare floats
  float thr_trim, thr_eff[3];
};

// One run of consecutive bins inside which nothing changes for k_bands_lane (afx_bands.cu): the sub-band, the frequency
// band and the (at most two) mel filters that cover it are constant, and the run does not cross a 32-bin tile.
struct AfxBandSeg {
  short k0, k1;             // bins [k0, k1)
  signed char b14, b28;     // sub-band / frequency band of the run, -1 = none
  signed char q0, nq;       // first mel filter covering the run and how many (0..2)
  unsigned char start14, end14, end28, fin;   // the run starts / ends its sub-band, ends its frequency band; fin bit 0 / 1: mel filter q0 / q0 + 1 ends with the run
};

struct AfxTables {          // per-context constant tables in device memory
  const double* window;     // [2048] Hann * 2
  const double2* tw2048;    // [2048] exp(-2 pi i k / 2048)
  const double2* tw512;     // [512]  exp(-2 pi i k / 512)
  // FFT pass twiddles laid out in access order (coalesced): afx_fft16.cuh
  const double2* fft_t2;        // [15][16]  exp(-2 pi i r k / 256),  r = 1..15
  const double2* fft_t3_1024;   // [3][256]  exp(-2 pi i r j / 1024), r = 1..3
  const double2* fft_t3_2048;   // [7][256]  exp(-2 pi i r j / 2048), r = 1..7
  const double* rwindow;    // [512] rhythm Hann x 0.5 (see k_rhythm_polar)
  const float2* ac_tw;      // [240 + 256 + 257] float32 twiddles of the autocorrelation's 512-point transforms (afx_autocorr.cu, ACT_*)
  const double* mel;        // [14][1024]
  const double2* mel_ab;    // [1024] per bin: weights of the (at most two) mel filters covering it, in filter order (k_bands_lane)
  const double* dct;        // [14][14] cos(pi n/14 (m+0.5)), row n
  const float* rs_imp;      // [69632] resampler wing
  unsigned int* work_ctr;   // [64] work-claim counters of the persistent kernels (zeroed on the launching stream)
  double* hl_pad;           // [21] last-frame values of a silent sample, written by afx_create (AFX_FEAT_HIGHLEVEL)
  const AfxBandSeg* band_segs;   // [n_band_segs] the walk of k_bands_lane over the 1024 bins
  int n_band_segs;
};

struct AfxParams {
  int sr, N, H;
  int first_bin, nbins;     // 1, 738
  int band14_start[14], band14_n[14], band14_nei[14];
  int band28_s[28], band28_e[28];
  int mel_lo[14], mel_hi[14];   // non-zero support [lo, hi] of each mel filter row
  double wh_decay, env_coef, silence_floor_amp;   // silence floor in 16-bit range (SA.cpp:648-649)
  double eff_floor[3];
  int analysis_cap;         // 882000
  int ac_min_period, ac_width;   // 35, 529
  // rhythm front / back end (OnsetDetector.cpp:106-111, 280-322; RhythmTracker.cpp:17-40; CannyWindow.cpp:27-80)
  float r_relax, r_norm_complex, r_norm_power;
  double canny[25];         // taps -12 .. +12 (the convolution uses -12 .. +11)
  AfxTables t;
};

// ---------------------------------------------------------------------------------------------
// the conditioned signal (reference: TSampleData::mData) as a function of the mono buffer
__device__ __forceinline__ double mdata(const float* __restrict__ mono, const AfxState& st, int i)
{
  const int j = i - st.start_off;
  return (j >= 0 && j < st.audible) ? (double)__ldg(mono + st.lead + j) * st.fs : 0.0;
}

// locate the file owning global slot `slot` in a prefix table off[] (off[i] = first slot of file i,
// file i owns [off_i, off_i + cap_i)); binary search over AfxFile entries
__device__ __forceinline__ int find_file_by_frame(const AfxFile* __restrict__ files, int n_files, int slot)
{
  int lo = 0, hi = n_files - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (files[mid].frame_off <= slot) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__device__ __forceinline__ int find_file_by_rframe(const AfxFile* __restrict__ files, int n_files, int slot)
{
  int lo = 0, hi = n_files - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (files[mid].rframe_off <= slot) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions (all threads of the block must call; blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// K simultaneous block-wide sums; scratch must hold K * 32 doubles.  Result broadcast to all threads.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* scratch)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();   // scratch may still be read from a previous call
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) scratch[k * 32 + wid] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double t = (lane < nw) ? scratch[k * 32 + lane] : 0.0;
    v[k] = warp_sum(t);
  }
}
__device__ __forceinline__ double block_max(double v, double* scratch)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double t = (lane < nw) ? scratch[lane] : -1.0e308;
  return warp_max(t);
}
__device__ __forceinline__ int block_sum_i(int v, int* scratch)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum_i(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  int t = (lane < nw) ? scratch[lane] : 0;
  return warp_sum_i(t);
}
__device__ __forceinline__ int block_min_i(int v, int* scratch)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  int t = (lane < nw) ? scratch[lane] : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t = min(t, __shfl_xor_sync(0xffffffffu, t, o));
  return t;
}

// aubio_silence_detection (mathutils.c:345-357, 606-615): 10 log10(mean square) < -48 dB.  log10 is monotonic,
// so the test is a comparison of the mean square with 10^-4.8 (saves a serial double log10 per frame)
#define AFX_SILENCE_LEVEL 1.5848931924611134e-05

// TAudioMath::LinToDb(double), AudioTypes/Export/AudioMath.inl:55-70 (MEpsilon is a float literal)
__device__ __forceinline__ double lin_to_db(double v)
{
  if (v == 1.0) return 0.0;
  if (v > (double)1e-12f) return log(v) * (20.0 / 2.302585092994045684);
  return -200.0;
}
// SFlatnessDb, SampleAnalyser.cpp:129-133, from a mean and a geometric mean
__device__ __forceinline__ double flatness_db(double mean, double gmean)
{
  const double fl = (mean == 0.0) ? 0.0 : gmean / mean;
  return fmin(lin_to_db(fl) / -60.0, 1.0);
}

// split a positive double into mantissa in [0.5, 1) and exponent, for overflow-free products
__device__ __forceinline__ void mul_frexp(double& mant, int& ex, double v)
{
  int e;
  mant *= frexp(v, &e);
  ex += e;
  if (mant < 0x1p-512) { mant *= 0x1p512; ex -= 512; }
}

// same for a NORMAL positive double (callers add 1e-20 first, Statistics.cpp:424): the exponent is peeled off
// with integer ops on the high word instead of the general frexp()
__device__ __forceinline__ void mul_frexp_pos(double& mant, int& ex, double v)
{
  const int hi = __double2hiint(v), lo = __double2loint(v);
  ex += ((hi >> 20) & 0x7ff) - 1022;
  mant *= __hiloint2double((hi & 0x800fffff) | 0x3fe00000, lo);
  if (mant < 0x1p-512) { mant *= 0x1p512; ex -= 512; }
}

// kernel launchers (one per translation unit); all asynchronous on `s`
struct AfxBatchDev {
  int n_files, TF, TFr;
  // the frame-level kernels run group by group (whole files) so that the per-frame scratch (mag, rpolar,
  // scratch) is bounded: this launch covers files [file0, file0 + g_files), main frame slots
  // [slot0, slot0 + g_slots) and rhythm frame slots [rslot0, rslot0 + g_rslots)
  int file0, g_files, slot0, g_slots, rslot0, g_rslots;
  const unsigned char* pcm;
  float* mono;
  float* mono_src;
  const AfxFile* files;
  const AfxInject* inject;  // null unless a file of the batch was conditioned in parts
  AfxState* state;
  const int* file_order;  // [n_files] per launch group [file0, file0 + g_files): the group's files, longest first -- CTA i of a per-file kernel takes file_order[file0 + i]
  const int* slot_file;   // [TF]  main frame slot -> file index (built on the device by k_slotmap)
  const int* rslot_file;  // [TFr] rhythm frame slot -> file index
  double* mag;        // [g_slots][1024]   (group scratch, indexed by slot - slot0)
  double* cent_full;  // [TF] centroid of mag[0..1023] (failsafe f0)
  double* bandraw;    // [g_slots][16] order statistics + peak counts of the five large sub-bands (group scratch)
  double* fs;         // [22][TF]
  double* fsr;        // [2][TFr]
  double* fv;         // 7 arrays [TF][nb], array v at fv + FV_x * TF
  float* rpolar;      // [g_rslots][512]   (group scratch, indexed by rslot - rslot0)
  float* rodf;        // [2][TFr] raw onset functions
  float* rpost;       // [2][TFr] onset functions minus their running median
  double* stats;      // [n_files][136][13]
  double* header;     // [n_files][32]
  double* scratch;    // [g_rslots][4] rhythm back-end workspace (group scratch)
  int max_fr;         // largest rhythm frame capacity of any file in the batch
  int pitch_generic;  // AFX_PITCH_GENERIC=1: the general 2048-point pitch kernel also at hop 1024 (else the block-sharing form)
  int rhythm_pipe;    // k_rhythm_pipe for this launch group: -1 = by group size, 0 = no, 1 = yes
  int rhythm_fused;   // this launch group's rhythm front end runs as ONE kernel with a CTA per file (k_rhythm_front); else the split kernels over rpolar
};

struct AfxHighLevelDev {   // outputs of k_highlevel (afx_highlevel.cu), batch-wide
  double* scalars;          // [n_files][AFX_N_HL]
  double* pitch;            // [TF] MIDI notes on the main frame grid (frame_off / F segments)
  double* signature;        // [n_files][64][14]
  double* features;         // [n_files][AFX_HL_FEATURES]
  int* status;              // [n_files] 1: a classification feature is NaN / Inf (the reference fails the file)
  const double* silence_pad;// [21] last-frame values of a silent sample (SampleClassificationDescriptors.cpp:330-368)
};

struct AfxExtDev {          // the mel-40 / MFCC-13 / chroma-12 extension (afx_ext.cu)
  const float* w_kmajor;    // [1024][52] weights, bin-major: 40 mel filters then 12 chroma classes
  const float* w_nmajor;    // [52][1024] the same, output-major (K-major rows for the tensor-core operands)
  const double* dct;        // [13][40] cos(pi n (m + 1/2) / 40)
  double* mfcc;             // [TF][13]
  double* chroma;           // [TF][12]
  double* chroma_index;     // [TF]
};
void afx_launch_ext(const AfxBatchDev& B, const AfxExtDev& X, bool tensor, cudaStream_t s, long long* launches);

struct AfxPackDev {         // outputs of k_pack (afx_pack.cu)
  unsigned char* packed;    // every file's msgpack BLOB images, file i at file_off[i]
  const unsigned long long* file_off;   // [n_files] host-built from the frame-slot capacities
  unsigned* blob_off;       // [n_files][AFX_N_BLOBS + 1] offsets of the blobs inside a file's region
};
size_t afx_pack_region_bytes(int F, int Fr);
void afx_launch_pack(const AfxBatchDev& B, const AfxPackDev& O, cudaStream_t s, long long* launches);

struct RsBlock { int out0; int nout; long long in0; long long chk_off; int span; int pad; };   // in0: source index of X[0] (may be negative); chk_off: first time checkpoint; span: X[0 .. span) covers every sample the block's filter sums read
struct AfxCondPlan {       // device arrays built by the host for one batch
  const int* src_chunk_file; const int* src_chunk_start; int n_src_chunks;   // chunks over source frames
  const int* dst_chunk_file; const int* dst_chunk_start; int n_dst_chunks;   // chunks over analysis-rate samples
  const int* rs_chunk_file; const int* rs_chunk_start; int n_rs_chunks;      // same, resampled files only
  const RsBlock* rs_blocks; const int* rs_blk_file; const double* rs_times; int n_rs_blocks;   // rs_times: every 64th output time stamp
  int rs_smem_bytes;         // shared memory the resampler wants for this batch (source span + coefficient rows of the largest rate)
};
void afx_launch_condition_plan(const AfxParams& P, const AfxBatchDev& B, const AfxCondPlan& C, cudaStream_t s, long long* launches);
void afx_launch_part_reduce(const AfxParams& P, const AfxBatchDev& B, const AfxCondPlan& C, cudaStream_t s);
void afx_launch_part_trim(const AfxParams& P, const AfxBatchDev& B, const AfxCondPlan& C, cudaStream_t s);
void afx_launch_part_eff(const AfxParams& P, const AfxBatchDev& B, const AfxCondPlan& C, cudaStream_t s);
void afx_launch_materialise(const float* mono, const AfxState* st, double* out, int len, cudaStream_t s);
void afx_launch_spectrum(const AfxParams& P, const AfxBatchDev& B, unsigned features, cudaStream_t s, long long* launches);
void afx_launch_peaks(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches);
void afx_launch_bands(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches);
void afx_launch_pitch(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches);
void afx_launch_autocorr(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches);
void afx_launch_rhythm(const AfxParams& P, const AfxBatchDev& B, cudaStream_t s, long long* launches);
void afx_launch_stats(const AfxParams& P, const AfxBatchDev& B, unsigned features, cudaStream_t s, long long* launches);
void afx_launch_highlevel(const AfxParams& P, const AfxBatchDev& B, const AfxHighLevelDev& O, cudaStream_t s, long long* launches);

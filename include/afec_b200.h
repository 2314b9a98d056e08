/*
 * afec_b200.h -- C ABI of libafec_b200.so, the B200 (sm_100a) implementation of AFEC's
 * low-level descriptor hot path.
 *
 * What it replaces in the reference (emuell/AFEC), file:line relative to the reference root:
 *   - the PCM conditioning tail of TSampleAnalyser::LoadSample
 *       Source/Crawler/FeatureExtraction/Source/SampleAnalyser.cpp:531-719
 *   - TSampleAnalyser::AnalyzeLowLevelDescriptors           SampleAnalyser.cpp:723-1066
 *   - every Calc* helper it calls                            SampleAnalyser.cpp:1715-2412
 *   - TStatistics::Calc for the per-file temporal statistics Statistics.cpp:12-90
 * The caller keeps file decoding (TAudioFile / TAudioStream::ReadSamples) and the descriptor
 * sink (TSampleDescriptorPool::InsertSample): it hands decoded PCM in, and gets the values of
 * TSampleDescriptors' low-level members back (Export/SampleDescriptors.h:396-466).
 *
 * There is no CPU fallback: every entry point that computes fails with AFX_ERR_CUDA when no
 * usable device is present.
 *
 * Plain C: pointers and sizes only.  All functions return AFX_OK (0) or a negative error code;
 * afx_last_error() gives the message of the last failure on that context.
 */
#ifndef AFEC_B200_H_
#define AFEC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFX_ABI_VERSION 3

/* ---- error codes ------------------------------------------------------------------------ */
#define AFX_OK 0
#define AFX_ERR_ARG (-1)        /* bad argument */
#define AFX_ERR_CUDA (-2)       /* CUDA runtime failure / no device: the batch's files failed to analyse */
#define AFX_ERR_NOMEM (-3)
#define AFX_ERR_STATE (-4)      /* call sequence violated */

/* per-file status (afx_file_result.status) -- mirrors the reference's load-time rejections,
 * SampleAnalyser.cpp:472-482 */
#define AFX_FILE_OK 0
#define AFX_FILE_BAD_CHANNELS 1 /* "Unsupported audio file channel layout" (channels outside 1..8) */
#define AFX_FILE_EMPTY 2        /* "Sample file is empty, probably failed to read." */
#define AFX_FILE_UNSUPPORTED 3  /* src_rate <= 0, unknown format, or more than INT32_MAX samples (the reference counts samples in
                                   an int, SampleAnalyser.cpp:484): this file fails, the rest of the batch is analysed */

/* ---- PCM formats -------------------------------------------------------------------------- */
/* All are INTERLEAVED frames, uploaded as they sit in the file and converted ON THE DEVICE to the float32 in 16-bit
 * range (+-32768) the reference's decoders hand LoadSample (Source/Core/CoreFileFormats/Export/SampleConverter.h:392-518
 * -- the conversions below restate those, value for value): raw bytes cross PCIe, no host conversion pass. */
#define AFX_PCM_I16 0     /* little-endian int16: (float)v                                   SampleConverter.h:446-449 */
#define AFX_PCM_F32 1     /* float32 already in 16-bit range, taken as is */
#define AFX_PCM_U8 2      /* unsigned 8-bit (WAV): (float)((v - 128) << 8)                    :392-395 */
#define AFX_PCM_I24 3     /* packed 3-byte little-endian: (float)((v24 << 8) * 32768.0 / 2^31) :473-487 */
#define AFX_PCM_I32 4     /* little-endian int32: clamp((float)(v * 32768.0 / 2^31))          :514-518 */
#define AFX_PCM_F32U 5    /* little-endian float32 in unit range (WAV IEEE float): (float)clamp(v * 32768.0), WaveFile.cpp */
#define AFX_PCM_I8 6      /* signed 8-bit (AIFF): (float)(v << 8)                             :410-413 */
#define AFX_PCM_I16BE 7   /* big-endian forms of the above (AIFF) */
#define AFX_PCM_I24BE 8
#define AFX_PCM_I32BE 9
#define AFX_PCM_F32UBE 10
#define AFX_PCM_FORMATS 11
/* bytes per sample of a format (0: unknown format) */
int afx_pcm_bytes(int32_t format);

/* ---- feature groups (afx_config.features) --------------------------------------------- */
#define AFX_FEAT_SPECTRAL (1u << 0)  /* window+FFT+magnitude, spectral rms/centroid/spread/skew/kurt/rolloff/flatness/flux */
#define AFX_FEAT_AMPLITUDE (1u << 1) /* amplitude silence/peak/rms/envelope */
#define AFX_FEAT_PEAKS (1u << 2)     /* whitening + peak spectrum -> spectral_complexity */
#define AFX_FEAT_BANDS (1u << 3)     /* 14 sub-band features, spectral_contrast, 28 frequency bands, 14 cepstrum bands */
#define AFX_FEAT_PITCH (1u << 4)     /* f0, f0_confidence, failsafe_f0 (YIN fast) */
#define AFX_FEAT_AUTOCORR (1u << 5)  /* auto_correlation */
#define AFX_FEAT_RHYTHM (1u << 6)    /* onset functions + rhythm scalars */
#define AFX_FEAT_STATS (1u << 7)     /* 13 statistics per series */
#define AFX_FEAT_ALL 0xFFu           /* the full low-level set (`Crawler --level low`) */
/* On top of AFX_FEAT_ALL: the high-level derivations that need no classification model (base note, loudness, bpm,
 * brightness / noisiness / harmonicity, spectrum signature, audible-frame spectral means, pitch series;
 * SampleAnalyser.cpp:1232-1606) and the classification feature vector the LightGBM models take
 * (SampleClassificationDescriptors.cpp:404-560).  Model evaluation itself stays on the host. */
#define AFX_FEAT_HIGHLEVEL (1u << 8)
/* The sink's BLOB images packed on the GPU: every VR / VVR column of an afec-ll.db row as the msgpack bytes the reference
 * stores (SqliteSampleDescriptorPool.cpp:601-713, 890-954), in column order, ready to bind (afx_file_result.packed).
 * afx_batch_download_rows() then copies back only what a sink needs -- the packed rows, headers and statistics -- and
 * leaves the framed arrays on the device (fs / fv stay NULL). */
#define AFX_FEAT_PACK (1u << 9)
/* Extension (BASELINE configs[2]; the reference has no counterpart): per main frame MFCC-13 over 40 mel
 * filters and a 12-class chroma vector with its arg-max, as one [frames x 1024] x [1024 x 52] contraction of the
 * magnitude spectra (afec_b200/csrc/afx_ext.cu; definition and CPU restatement: tests/ext_reference.py).  Needs
 * AFX_FEAT_SPECTRAL.  The contraction runs as an FP32 FMA tile, or -- AFX_EXT_TENSOR=1 in the environment when the
 * context is created -- as 3xTF32 on the tensor cores (tcgen05). */
#define AFX_FEAT_EXT_MELCHROMA (1u << 10)

typedef struct afx_ctx afx_ctx;
typedef struct afx_batch afx_batch;

typedef struct afx_config {
  int32_t device;       /* CUDA device ordinal */
  int32_t sample_rate;  /* TSampleAnalyser ctor args, Export/SampleAnalyser.h:33-36; Crawler.cpp:41-43 */
  int32_t fft_size;     /*   only 44100 / 2048 are supported (the Crawler's compile-time values) */
  int32_t hop_size;     /*   any hop with fft_size % hop == 0 (1024 default, 512 in BASELINE config 2) */
  uint32_t features;    /* AFX_FEAT_* mask; AFX_FEAT_ALL for a drop-in afec-ll.db */
  uint32_t reserved;
} afx_config;

/* one decoded file, as TSampleAnalyser::LoadSample sees it after ReadSamples (SampleAnalyser.cpp:484-528) */
typedef struct afx_file {
  const void* pcm;      /* interleaved frames; may be pageable or from afx_host_alloc (pinned) */
  int64_t nframes;      /* sample frames per channel */
  int32_t channels;
  int32_t src_rate;     /* != sample_rate -> resampled like libresample HQ (SampleAnalyser.cpp:563-607) */
  int32_t format;       /* AFX_PCM_* */
  int32_t bit_depth;    /* file property only (file_bit_depth_R) */
  int64_t file_size;    /* file property only (file_size_R) */
} afx_file;

/* number of entries of afx_file_result.header / series; order = Descriptors(kLowLevelDescriptors),
 * SampleDescriptors.cpp:154-203.  Names are listed in afec_b200/layout.py and host/afx_names.h. */
#define AFX_N_HEADER 32
#define AFX_N_FS 24          /* framed scalars; the first AFX_N_FS_MAIN run on the main frame grid */
#define AFX_N_FS_MAIN 22
#define AFX_N_FV 7           /* framed vectors: 5 x 14 sub-bands, 28 frequency bands, 14 cepstrum bands */
#define AFX_N_STATS 13
#define AFX_N_SERIES 136     /* 24 + 5*14 + 28 + 14 */
#define AFX_N_BLOBS 122      /* BLOB columns of a row: 22 framed scalars + 2 onset series (VR), then per framed vector its VVR and its
                                13 per-band statistic VRs -- the order of the table's columns (SqliteSampleDescriptorPool.cpp:1313-1350) */
#define AFX_N_HL 16          /* base_note, base_note_confidence, peak_db, rms_db, bpm, bpm_confidence, brightness, noisiness,
                                harmonicity, spectral_flatness, spectral_flux, spectral_complexity, spectral_contrast,
                                spectral_inharmonicity, pitch_confidence, reserved (names: afec_b200/layout.py HL_SCALARS) */
#define AFX_HL_SIGNATURE_FRAMES 64   /* TSampleDescriptors::kNumberOfHighLevelSpectrumBandFrames, SampleDescriptors.h:503 */
#define AFX_HL_SIGNATURE_BANDS 14
#define AFX_HL_FEATURES 1680         /* 35 rows of the 48-entry time series (SampleClassificationDescriptors.cpp:39-43, 536-547) */

typedef struct afx_file_result {
  int32_t status;            /* AFX_FILE_* */
  int32_t n_frames;          /* F : main frames (fft_size / hop_size grid) */
  int32_t n_rhythm_frames;   /* Fr: 512 / 128 grid */
  int32_t hl_status;         /* AFX_FEAT_HIGHLEVEL: 1 when a classification feature is NaN / Inf ("Invalid Descriptor array
                                value", SampleClassificationDescriptors.cpp:88-91: the file fails to analyse at --level high) */
  const double* header;      /* [AFX_N_HEADER] scalars */
  const double* fs[AFX_N_FS];/* fs[s][frame]; s < 22 -> n_frames values, s = 22, 23 -> n_rhythm_frames */
  const double* fv[AFX_N_FV];/* fv[v][frame * nbands(v) + band] */
  const double* stats;       /* [AFX_N_SERIES][AFX_N_STATS] */
  /* AFX_FEAT_HIGHLEVEL only (NULL otherwise) */
  const double* highlevel;   /* [AFX_N_HL] */
  const double* hl_pitch;    /* [n_frames] MIDI notes ("pitch"); "peak" is fs[1] (amplitude_peak), SampleAnalyser.cpp:1609 */
  const double* hl_signature;/* [AFX_HL_SIGNATURE_FRAMES][AFX_HL_SIGNATURE_BANDS] */
  const double* hl_features; /* [AFX_HL_FEATURES] */
  /* AFX_FEAT_EXT_MELCHROMA only (NULL otherwise) */
  const double* ext_mfcc;        /* [n_frames][13] */
  const double* ext_chroma;      /* [n_frames][12], max-normalised */
  const double* ext_chroma_index;/* [n_frames] arg-max class 0..11 (C = 0), integer-valued */
  /* AFX_FEAT_PACK only (NULL otherwise) */
  const unsigned char* packed;   /* the row's AFX_N_BLOBS msgpack blobs back to back */
  const uint32_t* packed_off;    /* [AFX_N_BLOBS + 1] byte offsets into packed: blob k = [packed_off[k], packed_off[k + 1]) */
} afx_file_result;

/* ---- context ------------------------------------------------------------------------------- */
int afx_abi_version(void);
int afx_create(const afx_config* cfg, afx_ctx** out);
void afx_destroy(afx_ctx* ctx);
const char* afx_last_error(const afx_ctx* ctx);   /* ctx may be NULL: last afx_create failure */
/* Give the context's device buffers back to the driver (they grow again on demand; constant tables, streams and
 * pinned host memory stay).  The reference has no counterpart: TSampleAnalyser::Extract frees its temporaries per
 * file (SampleAnalyser.cpp:372-408); here buffers are kept across batches and this is the explicit release between
 * crawls.  Waits for the context's streams; no batch of the context may be alive. */
int afx_trim(afx_ctx* ctx);

/* pinned host memory for decode threads (ring slots); H2D copies from it are asynchronous */
int afx_host_alloc(afx_ctx* ctx, uint64_t bytes, void** out);
int afx_host_free(afx_ctx* ctx, void* p);

/* ---- batches ------------------------------------------------------------------------------ */
/* The stages may be called one by one (upload / compute / download are asynchronous on the
 * context's stream), or all at once through afx_analyze().  PCM memory must stay valid until
 * afx_batch_upload() has been followed by afx_batch_sync() (or afx_analyze() returned).
 * The device buffers belong to the context: ONE batch per context may be between afx_batch_upload() and
 * afx_batch_free() at a time (a second upload returns AFX_ERR_STATE); batches in flight side by side need one
 * context each, which is how the host adapter's slots overlap copies with kernels. */
int afx_batch_create(afx_ctx* ctx, const afx_file* files, int32_t n_files, afx_batch** out);
int afx_batch_upload(afx_batch* b);     /* host -> device copies of the PCM + file table */
int afx_batch_compute(afx_batch* b);    /* all kernels of the configured feature set */
int afx_batch_download(afx_batch* b);   /* device -> pinned host copies of every result array */
int afx_batch_download_rows(afx_batch* b);  /* AFX_FEAT_PACK: packed rows + headers + statistics only (instead of afx_batch_download) */
int afx_batch_sync(afx_batch* b);       /* wait for everything issued so far */
int afx_analyze(afx_ctx* ctx, const afx_file* files, int32_t n_files, afx_batch** out); /* all of the above */
int afx_batch_result(const afx_batch* b, int32_t file_index, afx_file_result* out);
void afx_batch_free(afx_batch* b);

/* ---- long files conditioned in parts ------------------------------------------------------------- */
/* One long file (BASELINE config 5: 1-hour 96 kHz stereo) is cut into sample-range parts; each part is
 * downmixed / resampled / reduced by its own context (its own GPU), and the per-file reductions that
 * TSampleAnalyser::LoadSample makes over the whole file are combined BY THE CALLER between three phases
 * (a handful of scalars: host reduce, or an all-reduce when the parts live in different processes):
 *   afx_part_peak       max |x|, sum of squares            SampleAnalyser.cpp:612-631   combine: max, sum
 *   afx_part_trim       first / last sample above -48 dB   SampleAnalyser.cpp:636-669   combine: min, max
 *   afx_part_effective  -48 / -24 / -12 dB effective spans  SampleAnalyser.cpp:1715-1756 combine: min, max
 * afx_part_sums_merge() is that combine: merging the outputs of ALL parts of a phase gives the global sums that
 * the next phase takes as input (fields a phase does not compute are passed through so that they merge back to
 * themselves).  The analysis itself reads at most the 20 s behind the trim point
 * (SampleAnalyser.cpp:37, 760-764): afx_part_window() names those samples, afx_part_read() hands out the ones a
 * part owns, and afx_analyze_conditioned() runs the regular kernel schedule on them.  Results are those of
 * afx_analyze() on the whole file (bit-identical, except that the sum of squares adds in a different order). */
typedef struct afx_part {
  int64_t src_begin, src_end;   /* source frames the part needs, filter halo included: pcm_slice = frames [begin, end) */
  int64_t out_begin, out_end;   /* analysis-rate samples the part produces */
} afx_part;

typedef struct afx_part_sums {
  float maxabs;                 /* max |x|, 16-bit range */
  int32_t reserved;
  double sumsq;                 /* sum (x / 32768)^2 */
  int64_t first, last;          /* analysis-rate sample indices; INT64_MAX / -1 when nothing is above the floor */
  int64_t eff_first[3], eff_last[3];   /* conditioned-signal (mData) indices at -48 / -24 / -12 dB */
} afx_part_sums;

typedef struct afx_partjob afx_partjob;

/* even split of a file into n_parts (cut at libresample block boundaries when src_rate != sample_rate);
 * host-only arithmetic: needs no context and no device */
int afx_part_plan(int32_t sample_rate, int64_t nframes, int32_t src_rate, int32_t n_parts, afx_part* out);
void afx_part_sums_init(afx_part_sums* s);
void afx_part_sums_merge(afx_part_sums* acc, const afx_part_sums* other);
/* `whole` describes the whole file (nframes = all frames; its pcm member is ignored), pcm_slice holds the
 * interleaved frames [part->src_begin, part->src_end) and must stay valid until afx_part_peak() returned (the
 * host -> device copy is asynchronous) */
int afx_part_open(afx_ctx* ctx, const afx_file* whole, const afx_part* part, const void* pcm_slice, afx_partjob** out);
int afx_part_peak(afx_partjob* j, afx_part_sums* out);
int afx_part_trim(afx_partjob* j, const afx_part_sums* global, afx_part_sums* out);
int afx_part_effective(afx_partjob* j, const afx_part_sums* global, afx_part_sums* out);
int afx_part_window(afx_ctx* ctx, const afx_file* whole, const afx_part_sums* global, int64_t* begin, int64_t* count);
/* copies the samples of [begin, begin + count) this part owns to dst + (index - begin); returns how many */
int64_t afx_part_read(afx_partjob* j, int64_t begin, int64_t count, float* dst);
void afx_part_close(afx_partjob* j);
/* mono = analysis-rate samples [mono_begin, mono_begin + mono_count) of the file (at least afx_part_window's) */
int afx_analyze_conditioned(afx_ctx* ctx, const afx_file* whole, const afx_part_sums* global, const float* mono,
                            int64_t mono_begin, int64_t mono_count, afx_batch** out);

/* ---- measurement helpers ---------------------------------------------------------------------- */
/* device time (ms, CUDA events on the context's stream) of the last upload / compute / download */
int afx_batch_timings(const afx_batch* b, float* upload_ms, float* compute_ms, float* download_ms);
/* number of kernels launched by the last afx_batch_compute() and bytes moved by upload / download */
int afx_batch_counters(const afx_batch* b, int64_t* kernel_launches, int64_t* h2d_bytes, int64_t* d2h_bytes,
                       int64_t* main_frames, int64_t* rhythm_frames);
/* per-kernel device time of the last compute (only when the context was created with
 * AFX_DEBUG_KERNEL_TIMES=1 in the environment); returns the number of entries written */
int afx_batch_kernel_times(const afx_batch* b, const char** names, float* ms, int32_t cap);

/* measured FP64 FMA throughput of the device (TFLOP/s), the denominator of the compute roofline
 * (MEASURED_PEAKS.json only holds HBM and bf16 tensor peaks) */
int afx_measure_fp64_peak(afx_ctx* ctx, double* tflops);

/* debugging / parity: copy the conditioned signal (the reference's TSampleData::mData,
 * SampleAnalyser.cpp:698-718) of one file back to the host.  Returns its length. */
int64_t afx_batch_conditioned(const afx_batch* b, int32_t file_index, double* out, int64_t cap);

/* debugging / parity: forward complex FFT (exp(-i), unscaled) of `batch` interleaved re/im sequences of
 * n = 256, 1024 or 2048 points through the kernels' own FFT core -- the known-answer hook that plays the
 * role of the reference's TAudioTypesTest::Fourier (Source/Core/AudioTypes/Test/TestFourier.cpp:12-84). */
int afx_debug_fft(afx_ctx* ctx, int32_t n, int32_t batch, const double* in, double* out);

/* debugging / parity: 0 when the closed-form replay of libresample's time accumulator (resamplesubs.c:97-119)
 * used to plan the resampler kernel equals the step-by-step replay bit for bit; host arithmetic only */
int afx_debug_rs_plan_check(int32_t sample_rate, int64_t nframes, int32_t src_rate);

#ifdef __cplusplus
}
#endif
#endif /* AFEC_B200_H_ */

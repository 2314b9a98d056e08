// Host-side internals shared by the translation units that implement the C ABI (afx_api.cu, afx_part.cu):
// growable buffers, the context and the batch.  Not part of the public interface (include/afec_b200.h).
#pragma once
#include "afx_common.cuh"
#include "../../include/afec_b200.h"

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

// growable device / pinned buffers ------------------------------------------------------------
struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// data-independent replay of libresample's block / time bookkeeping for one (rate, length) pair
struct RsShape {
  std::vector<RsBlock> blocks;   // chk_off relative to chk
  std::vector<double> chk;       // every 64th output time stamp of each block
  std::vector<int> span;         // per block: source samples [in0, in0 + span) cover everything its filter sums read
  int produced = 0;              // output samples libresample delivers (<= the requested count)
  int max_span = 0;
};

// device + pinned buffers of one part job (afx_part.cu); recycled through afx_ctx::part_pool so that a stream of
// long files does not pay cudaMalloc / cudaFree per part
struct PartBufs {
  DevBuf pcm, mono_src, mono, tab, state;
  PinBuf htab;
  void release() { pcm.release(); mono_src.release(); mono.release(); tab.release(); state.release(); htab.release(); }
};

struct afx_ctx {
  afx_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  // side streams: the pitch, autocorrelation and rhythm chains only depend on the conditioned signal (pitch also on
  // the spectrum's centroid), so they run beside the spectrum -> bands -> peaks chain and fill each other's idle pipes
  cudaStream_t side[3] = { nullptr, nullptr, nullptr };
  cudaStream_t copy_stream = nullptr;   // H2D copies of part jobs: a later part uploads while an earlier one computes
  cudaEvent_t ev_fork = nullptr, ev_spec = nullptr, ev_join[3] = { nullptr, nullptr, nullptr };
  cudaEvent_t ev_chain = nullptr;   // end of this context's last compute (see g_chain in afx_api.cu)
  bool compute_chain = true;
  bool multi_stream = false;
  AfxParams P;
  DevBuf tables;                      // all constant tables in one allocation
  DevBuf d_pcm, d_mono, d_mono_src, d_files, d_state, d_mag, d_cent, d_fs, d_fsr, d_fv, d_rpolar, d_rodf, d_rpost, d_bandraw, d_slotmap,
         d_stats, d_header, d_plan, d_scratch, d_hl, d_hl_pitch, d_hl_sig, d_hl_feat, d_hl_status, d_pack, d_pack_off, d_pack_file_off, d_ext_mfcc, d_ext_chroma, d_ext_idx;
  bool ext_tensor = false;            // the extension's contraction runs on the tensor cores (AFX_EXT_TENSOR=1)
  AfxExtDev ext_tables;               // weight / DCT tables of the extension (inside `tables`)
  bool hl_pad_ready = false;          // the silence pad table (tables: hl_pad) has been computed by afx_create
  long long group_frames = 1572864, group_rframes = 12582912;   // per-launch scratch bound: 12 GB mag, 24 GB rpolar (allocated by need). Large groups matter to the per-file kernels: 4x the files in flight took 17 % off the rhythm chain
  PinBuf h_results_cache, h_plan_cache, h_pack_cache;   // recycled between batches
  std::vector<PartBufs> part_pool;        // recycled between part jobs
  std::vector<double> zeros;          // backing store of the all-zero series
  std::string error;
  bool debug_times = false;
  std::mutex mu;
  int max_frame_cap = 0;
  // launch groups with at least this many files run the fused rhythm front end (one CTA per file).  MEASURED SLOWER than
  // the split kernels on the full workload (158 vs 135 ms per 50.7 M rhythm frames, profiles/README.md: with 16 warps per
  // SM every phase of the fused kernel is latency bound on its own), so it is off unless AFX_RHYTHM_FUSED=1 asks for it
  // (AFX_RHYTHM_FUSED=0 / 1 forces never / always: the parity test compares the two schedules)
  int rhythm_fused_min = 0x7fffffff;
  bool pitch_generic = false;
  // whitening + onset functions as one producer / consumer pipeline per file (k_rhythm_pipe): -1 = for launch groups with at
  // least two files per SM, 0 / 1 = never / always (AFX_RHYTHM_PIPE; the parity test compares the schedules)
  int rhythm_pipe = -1;
  bool rhythm_fused(int g_files) const { return g_files >= rhythm_fused_min; }
  struct afx_batch* live = nullptr;   // the one batch whose data occupies the device buffers (afx_batch_upload .. afx_batch_free)
};

struct KernelTime { const char* name; cudaEvent_t a, b; };

struct afx_batch {
  afx_ctx* ctx = nullptr;
  int n_files = 0;
  std::vector<afx_file> in;
  std::vector<AfxFile> files;
  std::vector<AfxState> state_host;
  // plan
  std::vector<int> src_chunk_file, src_chunk_start, dst_chunk_file, dst_chunk_start, rs_chunk_file, rs_chunk_start;
  std::vector<RsBlock> rs_blocks; std::vector<int> rs_blk_file; std::vector<double> rs_chk;
  std::vector<std::shared_ptr<RsShape>> rs_shapes;   // the batch's distinct resampler shapes (kept alive whatever the process-wide cache evicts)
  struct Tail { long long off; long long count; }; std::vector<Tail> rs_tails;   // mono samples libresample never writes
  struct Group { int file0, nfiles, slot0, nslots, rslot0, nrslots; };
  std::vector<Group> groups;
  std::vector<int> file_order;        // per launch group: its file indices, longest first (per-file kernels start the long ones early)
  int max_gslots = 0, max_grslots = 0, max_fr = 0, rs_smem = 0;
  struct CopyRun { const unsigned char* host; size_t dev_off; size_t bytes; bool to_mono; };
  std::vector<AfxInject> inject;      // conditioning reductions made elsewhere (files conditioned in parts)
  std::vector<CopyRun> runs;
  size_t pcm_bytes = 0; long long mono_samples = 0, mono_src_samples = 0;
  int TF = 0, TFr = 0;
  PinBuf h_plan;                      // pinned staging of file table + chunk tables
  PinBuf h_results;                   // pinned results
  // host result layout (offsets in doubles inside h_results)
  size_t o_header = 0, o_state = 0, o_fs = 0, o_fsr = 0, o_fv = 0, o_stats = 0, total_doubles = 0;
  size_t o_hl = 0, o_hl_pitch = 0, o_hl_sig = 0, o_hl_feat = 0, o_hl_status = 0;
  AfxHighLevelDev hl;
  size_t o_ext_mfcc = 0, o_ext_chroma = 0, o_ext_idx = 0;
  AfxExtDev ext;
  // packed sink rows (AFX_FEAT_PACK): a separate pinned block; file i's region starts at pack_file_off[i]
  std::vector<unsigned long long> pack_file_off; size_t pack_bytes = 0;
  PinBuf h_pack;                      // [pack_bytes] blobs, then [n][AFX_N_BLOBS + 1] uint32 offsets
  AfxPackDev pack;
  AfxBatchDev dev;
  AfxCondPlan cond;
  cudaEvent_t ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
  bool uploaded = false, computed = false, downloaded = false, arrays_downloaded = false;
  long long launches = 0, h2d_bytes = 0, d2h_bytes = 0;
  std::vector<KernelTime> ktimes;
};


// shared helpers (afx_api.cu)
int afx_fail(afx_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess);
std::shared_ptr<RsShape> afx_rs_shape(int sr, int in_len, int src_rate, int out_len);   // cached libresample replay (process-wide)
int afx_rs_smem_need(int sr, int src_rate, int span);                                          // dynamic shared memory of k_resample
int afx_reference_round(double v);                                                            // TMath::d2iRound
struct AfxCondInput {  // a file whose conditioning reductions were made elsewhere (afx_part.cu): the batch holds one such file
  const float* mono;                  // analysis-rate mono samples [mono_begin, mono_begin + mono_count) of the file
  long long mono_begin, mono_count;
  AfxInject inj;
};
int afx_batch_create_impl(afx_ctx* ctx, const afx_file* files, int32_t n_files, const AfxCondInput* cond, afx_batch** out);

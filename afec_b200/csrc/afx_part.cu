// Long files conditioned in parts (BASELINE config 5): one file's sample range is cut into
// parts, each part is downmixed / resampled / reduced on its own GPU, and the few per-file reductions the
// reference makes over the whole file are combined on the host between three phases:
//
//   phase A  afx_part_peak       downmix + libresample-exact resample of the part, max |x| and sum of squares
//                                (SampleAnalyser.cpp:535-631)                 -> combine: max, sum
//   phase B  afx_part_trim       first / last sample above the -48 dB floor, which depends on the global peak
//                                through Amplification (SA.cpp:636-669)       -> combine: min, max
//   phase C  afx_part_effective  effective-length scans at -48 / -24 / -12 dB over the audible region, which
//                                depends on the global trim (SA.cpp:1715-1756) -> combine: min, max
//
// The analysis then only needs the <= 882000 audible samples behind the global trim point (SA.cpp:37, 760-764):
// afx_part_read hands out the pieces each part owns and afx_analyze_conditioned runs the regular kernel schedule
// on that window with the combined reductions injected in place of the conditioning passes.
//
// The kernels are the ones of afx_condition.cu: a part is described to them as a file whose buffers are offset
// so that GLOBAL sample indices address the part's slice, with the passes limited to the part's range.  Parts
// of a resampled file are cut at libresample block boundaries (the block / time-stamp replay of the whole file
// is data independent), and a part's source slice carries the filter halo of its first and last block.
#include "afx_internal.h"

#include <algorithm>
#include <climits>
#include <cstring>

#define CHUNK 8192
#define CKP(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return afx_fail(ctx, AFX_ERR_CUDA, what, e_); } while (0)

struct afx_partjob {
  afx_ctx* ctx = nullptr;
  afx_part part;
  AfxFile file;
  AfxState st;
  PartBufs bufs;
  AfxBatchDev dev;
  AfxCondPlan plan;
  bool resampled = false;
  long long zero_from = -1, zero_count = 0;     // analysis-rate samples libresample never delivers (stay 0)
  int phase = 0, rs_smem = 0;
  cudaEvent_t ev_h2d = nullptr;                 // the slice and the tables are on the device
  double own_sumsq = 0.0;                       // this part's share: every phase's output merges back to the global sums
  void release()      // hand the buffers back to the context (caller holds ctx->mu)
  {
    if (ctx->part_pool.size() < 8) ctx->part_pool.push_back(bufs); else bufs.release();
    bufs = PartBufs();
  }
};

static int analysis_len(int sr, long long nframes, int src_rate)
{
  const double speed = (double)src_rate / (double)sr;                 // SA.cpp:563-573
  if (speed == 1.0) return (int)nframes;
  const int nn = afx_reference_round((double)(int)nframes / speed);
  return nn < 1 ? 1 : nn;
}

extern "C" int afx_part_plan(int32_t sample_rate, int64_t nframes, int32_t src_rate, int32_t n_parts, afx_part* out)
{
  if (!out || n_parts < 1 || nframes <= 0 || nframes > 0x7fffffffLL || src_rate <= 0 || sample_rate <= 0) return AFX_ERR_ARG;
  const long long n = analysis_len(sample_rate, nframes, src_rate);
  if (src_rate == sample_rate) {
    for (int p = 0; p < n_parts; ++p) {
      long long b = (n * p / n_parts) & ~(long long)(CHUNK - 1), e = (p + 1 == n_parts) ? n : ((n * (p + 1) / n_parts) & ~(long long)(CHUNK - 1));
      out[p].out_begin = b; out[p].out_end = e; out[p].src_begin = b; out[p].src_end = e;
    }
    return AFX_OK;
  }
  std::shared_ptr<RsShape> sh = afx_rs_shape(sample_rate, (int)nframes, src_rate, (int)n);
  const int nb = (int)sh->blocks.size();
  int b0 = 0;
  for (int p = 0; p < n_parts; ++p) {
    // first block of the next part: the first one that starts at or after the even split point
    int b1 = nb;
    if (p + 1 < n_parts) {
      const long long target = n * (p + 1) / n_parts;
      b1 = b0;
      while (b1 < nb && sh->blocks[b1].out0 < target) ++b1;
    }
    afx_part& r = out[p];
    r.out_begin = (b0 < nb) ? sh->blocks[b0].out0 : n;
    r.out_end = (b1 < nb) ? sh->blocks[b1].out0 : n;
    if (p == 0) r.out_begin = 0;
    if (b1 > b0) {
      long long lo = sh->blocks[b0].in0, hi = lo;
      for (int k = b0; k < b1; ++k) { lo = std::min(lo, sh->blocks[k].in0); hi = std::max(hi, sh->blocks[k].in0 + sh->span[k]); }
      lo = std::max(0LL, lo) & ~3LL; hi = std::min((long long)nframes, hi);
      r.src_begin = lo; r.src_end = std::max(lo, hi);
    } else { r.src_begin = 0; r.src_end = 0; }
    b0 = b1;
  }
  return AFX_OK;
}

extern "C" void afx_part_sums_init(afx_part_sums* s)
{
  if (!s) return;
  memset(s, 0, sizeof(*s));
  s->first = INT64_MAX; s->last = -1;
  for (int k = 0; k < 3; ++k) { s->eff_first[k] = INT64_MAX; s->eff_last[k] = -1; }
}

extern "C" void afx_part_sums_merge(afx_part_sums* a, const afx_part_sums* b)
{
  if (!a || !b) return;
  a->maxabs = std::max(a->maxabs, b->maxabs);
  a->sumsq += b->sumsq;
  a->first = std::min(a->first, b->first); a->last = std::max(a->last, b->last);
  for (int k = 0; k < 3; ++k) { a->eff_first[k] = std::min(a->eff_first[k], b->eff_first[k]); a->eff_last[k] = std::max(a->eff_last[k], b->eff_last[k]); }
}

static int clamp_first(int64_t v) { return v >= 0x7fffffffLL ? 0x7fffffff : (int)v; }

extern "C" int afx_part_open(afx_ctx* ctx, const afx_file* whole, const afx_part* part, const void* pcm_slice, afx_partjob** out)
{
  if (!ctx || !whole || !part || !out) return afx_fail(ctx, AFX_ERR_ARG, "afx_part_open: null argument");
  *out = nullptr;
  if (whole->channels < 1 || whole->channels > 8 || whole->nframes <= 0 || whole->nframes * whole->channels > 0x7fffffffLL || whole->src_rate <= 0 ||
      afx_pcm_bytes(whole->format) == 0)
    return afx_fail(ctx, AFX_ERR_ARG, "afx_part_open: unsupported file description");
  if (part->src_begin < 0 || part->src_end < part->src_begin || part->src_end > whole->nframes || part->out_begin < 0 || part->out_end < part->out_begin ||
      (part->src_begin & 3) || (part->src_end > part->src_begin && !pcm_slice))
    return afx_fail(ctx, AFX_ERR_ARG, "afx_part_open: bad part (use afx_part_plan)");
  const AfxParams& P = ctx->P;
  const int n = analysis_len(P.sr, whole->nframes, whole->src_rate);
  std::shared_ptr<RsShape> sh;
  if (whole->src_rate != P.sr) sh = afx_rs_shape(P.sr, (int)whole->nframes, whole->src_rate, n);
  if (part->out_end > n) return afx_fail(ctx, AFX_ERR_ARG, "afx_part_open: part exceeds the file");
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  afx_partjob* j = new afx_partjob();
  j->ctx = ctx; j->part = *part;
  if (!ctx->part_pool.empty()) { j->bufs = ctx->part_pool.back(); ctx->part_pool.pop_back(); }
  j->resampled = whole->src_rate != P.sr;
  const size_t bps = (size_t)afx_pcm_bytes(whole->format);
  const long long ns = part->src_end - part->src_begin, no = part->out_end - part->out_begin;
  const size_t pcm_bytes = (size_t)ns * whole->channels * bps;

  AfxFile& f = j->file;
  memset(&f, 0, sizeof(f));
  f.channels = whole->channels; f.src_rate = whole->src_rate; f.format = whole->format; f.bit_depth = whole->bit_depth; f.file_size = whole->file_size;
  f.nframes_src = (int)whole->nframes; f.n = n; f.status = AFX_FILE_OK; f.inject = -1;
  f.pcm_off = -(long long)((size_t)part->src_begin * whole->channels * bps);      // global frame index -> slice
  f.src_off = -part->src_begin; f.mono_off = j->resampled ? -part->out_begin : -part->src_begin;
  f.src_end = (int)part->src_end; f.dst_end = (int)part->out_end;
  { long long lmax = std::max<long long>((long long)n + P.N / 2, P.N); if (lmax > P.analysis_cap) lmax = P.analysis_cap;
    f.frame_cap = (int)((lmax - P.N) / P.H + 1); f.rframe_cap = (int)((lmax - AFX_RFFT) / AFX_RHOP + 1); }

  // plan tables: chunk lists over the part's ranges, the part's resampler blocks and their time checkpoints
  std::vector<int> scf, scs, dcf, dcs;
  std::vector<RsBlock> blocks; std::vector<int> blkf; std::vector<double> chk;
  for (long long s = part->src_begin; s < part->src_end; s += CHUNK) { scf.push_back(0); scs.push_back((int)s); }
  for (long long s = part->out_begin; s < part->out_end; s += CHUNK) { dcf.push_back(0); dcs.push_back((int)s); }
  if (j->resampled) {
    long long covered = part->out_begin;
    for (size_t k = 0; k < sh->blocks.size(); ++k) {
      RsBlock rb = sh->blocks[k];
      if (rb.out0 < part->out_begin || rb.out0 >= part->out_end) continue;
      const long long c0 = rb.chk_off, c1 = c0 + (rb.nout + 63) / 64 + 1;
      rb.chk_off = (long long)chk.size();
      for (long long c = c0; c < c1 && c < (long long)sh->chk.size(); ++c) chk.push_back(sh->chk[c]);
      blocks.push_back(rb); blkf.push_back(0);
      covered = rb.out0 + rb.nout;
      j->rs_smem = std::max(j->rs_smem, afx_rs_smem_need(P.sr, f.src_rate, rb.span));
    }
    if (covered < part->out_end) { j->zero_from = covered; j->zero_count = part->out_end - covered; }   // SA.cpp:579-580
  }
  size_t po = 0;
  auto place = [&](size_t bytes) { size_t o = po; po += (bytes + 255) & ~(size_t)255; return o; };
  const size_t p_file = place(sizeof(AfxFile)), p_scf = place(scf.size() * 4), p_scs = place(scs.size() * 4), p_dcf = place(dcf.size() * 4),
    p_dcs = place(dcs.size() * 4), p_rb = place(blocks.size() * sizeof(RsBlock)), p_rbf = place(blkf.size() * 4), p_chk = place(chk.size() * 8);
  PartBufs& Bf = j->bufs;
  cudaError_t e = Bf.htab.reserve(po + 256);
  if (e == cudaSuccess) e = Bf.tab.reserve(po + 256);
  if (e == cudaSuccess) e = Bf.state.reserve(sizeof(AfxState) * 2);
  if (e == cudaSuccess) e = Bf.pcm.reserve(pcm_bytes + 64);
  if (e == cudaSuccess) e = Bf.mono.reserve((size_t)((j->resampled ? no : ns) + 16) * 4);
  if (e == cudaSuccess && j->resampled) e = Bf.mono_src.reserve((size_t)(ns + 16) * 4);
  if (e == cudaSuccess) {
    unsigned char* tab = (unsigned char*)Bf.htab.p;
    memcpy(tab + p_file, &f, sizeof(f));
    if (!scf.empty()) { memcpy(tab + p_scf, scf.data(), scf.size() * 4); memcpy(tab + p_scs, scs.data(), scs.size() * 4); }
    if (!dcf.empty()) { memcpy(tab + p_dcf, dcf.data(), dcf.size() * 4); memcpy(tab + p_dcs, dcs.data(), dcs.size() * 4); }
    if (!blocks.empty()) { memcpy(tab + p_rb, blocks.data(), blocks.size() * sizeof(RsBlock)); memcpy(tab + p_rbf, blkf.data(), blkf.size() * 4); }
    if (!chk.empty()) memcpy(tab + p_chk, chk.data(), chk.size() * 8);
    e = cudaMemcpyAsync(Bf.tab.p, tab, po, cudaMemcpyHostToDevice, ctx->copy_stream);
  }
  // asynchronous, on the context's copy stream: afx_part_peak makes the kernels wait for it, so a part opened
  // early uploads while earlier parts compute (pcm_slice must stay valid until afx_part_peak returned)
  if (e == cudaSuccess && pcm_bytes) e = cudaMemcpyAsync(Bf.pcm.p, pcm_slice, pcm_bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&j->ev_h2d, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(j->ev_h2d, ctx->copy_stream);
  if (e != cudaSuccess) { if (j->ev_h2d) cudaEventDestroy(j->ev_h2d); j->release(); delete j; return afx_fail(ctx, AFX_ERR_CUDA, "afx_part_open", e); }

  unsigned char* dp = (unsigned char*)Bf.tab.p;
  AfxBatchDev& D = j->dev;
  memset(&D, 0, sizeof(D));
  D.n_files = 1; D.g_files = 1;
  D.pcm = (const unsigned char*)Bf.pcm.p; D.mono = (float*)Bf.mono.p; D.mono_src = (float*)Bf.mono_src.p;
  D.files = (const AfxFile*)(dp + p_file); D.state = (AfxState*)Bf.state.p;
  AfxCondPlan& C = j->plan;
  memset(&C, 0, sizeof(C));
  C.src_chunk_file = (const int*)(dp + p_scf); C.src_chunk_start = (const int*)(dp + p_scs); C.n_src_chunks = (int)scf.size();
  C.dst_chunk_file = (const int*)(dp + p_dcf); C.dst_chunk_start = (const int*)(dp + p_dcs); C.n_dst_chunks = (int)dcf.size();
  if (j->resampled) { C.rs_chunk_file = C.dst_chunk_file; C.rs_chunk_start = C.dst_chunk_start; C.n_rs_chunks = C.n_dst_chunks; }
  C.rs_blocks = (const RsBlock*)(dp + p_rb); C.rs_blk_file = (const int*)(dp + p_rbf); C.rs_times = (const double*)(dp + p_chk);
  C.n_rs_blocks = (int)blocks.size(); C.rs_smem_bytes = j->rs_smem;
  *out = j;
  return AFX_OK;
}

static int fetch_state(afx_partjob* j)
{
  afx_ctx* ctx = j->ctx;
  CKP(cudaMemcpyAsync(&j->st, j->dev.state, sizeof(AfxState), cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(state)");
  CKP(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  CKP(cudaGetLastError(), "kernel launch");
  return AFX_OK;
}
static int push_state(afx_partjob* j)
{
  afx_ctx* ctx = j->ctx;
  CKP(cudaMemcpyAsync(j->dev.state, &j->st, sizeof(AfxState), cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(state)");
  CKP(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  return AFX_OK;
}

extern "C" int afx_part_peak(afx_partjob* j, afx_part_sums* out)
{
  if (!j || !out) return AFX_ERR_ARG;
  afx_ctx* ctx = j->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  CKP(cudaStreamWaitEvent(ctx->stream, j->ev_h2d, 0), "cudaStreamWaitEvent");
  if (j->zero_count > 0)
    CKP(cudaMemsetAsync((float*)j->bufs.mono.p + (j->zero_from - j->part.out_begin), 0, (size_t)j->zero_count * 4, ctx->stream), "cudaMemsetAsync(mono tail)");
  afx_launch_part_reduce(ctx->P, j->dev, j->plan, ctx->stream);
  int rc = fetch_state(j);
  if (rc != AFX_OK) return rc;
  afx_part_sums_init(out);
  float m; memcpy(&m, &j->st.maxabs_bits, 4);
  out->maxabs = m; out->sumsq = j->st.sumsq; j->own_sumsq = j->st.sumsq;
  j->phase = 1;
  return AFX_OK;
}

extern "C" int afx_part_trim(afx_partjob* j, const afx_part_sums* global, afx_part_sums* out)
{
  if (!j || !global || !out) return AFX_ERR_ARG;
  afx_ctx* ctx = j->ctx;
  if (j->phase < 1) return afx_fail(ctx, AFX_ERR_STATE, "afx_part_trim: afx_part_peak first");
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  memcpy(&j->st.maxabs_bits, &global->maxabs, 4);
  j->st.sumsq = global->sumsq; j->st.first = 0x7fffffff; j->st.last = -1;
  int rc = push_state(j);
  if (rc != AFX_OK) return rc;
  afx_launch_part_trim(ctx->P, j->dev, j->plan, ctx->stream);
  if ((rc = fetch_state(j)) != AFX_OK) return rc;
  *out = *global;
  out->sumsq = j->own_sumsq;
  out->first = (j->st.first == 0x7fffffff) ? INT64_MAX : j->st.first;
  out->last = j->st.last;
  j->phase = 2;
  return AFX_OK;
}

extern "C" int afx_part_effective(afx_partjob* j, const afx_part_sums* global, afx_part_sums* out)
{
  if (!j || !global || !out) return AFX_ERR_ARG;
  afx_ctx* ctx = j->ctx;
  if (j->phase < 2) return afx_fail(ctx, AFX_ERR_STATE, "afx_part_effective: afx_part_trim first");
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  j->st.first = clamp_first(global->first); j->st.last = (int)global->last;
  int rc = push_state(j);
  if (rc != AFX_OK) return rc;
  afx_launch_part_eff(ctx->P, j->dev, j->plan, ctx->stream);
  if ((rc = fetch_state(j)) != AFX_OK) return rc;
  *out = *global;
  out->sumsq = j->own_sumsq;
  for (int k = 0; k < 3; ++k) {
    out->eff_first[k] = (j->st.eff_first[k] == 0x7fffffff) ? INT64_MAX : j->st.eff_first[k];
    out->eff_last[k] = j->st.eff_last[k];
  }
  j->phase = 3;
  return AFX_OK;
}

// SA.cpp:651-669, 760-764: the analysis reads the audible samples behind the trim point, at most the 20 s cap
extern "C" int afx_part_window(afx_ctx* ctx, const afx_file* whole, const afx_part_sums* global, int64_t* begin, int64_t* count)
{
  if (!ctx || !whole || !global || !begin || !count) return AFX_ERR_ARG;
  const long long n = analysis_len(ctx->P.sr, whole->nframes, whole->src_rate);
  const long long lead = std::min<long long>(global->first, n);
  long long trail = 0;
  if (lead < n) { const long long last = std::max<long long>(global->last, lead); trail = n - 1 - last; }
  const long long audible = n - lead - trail;
  *begin = lead;
  *count = std::min<long long>(audible, ctx->P.analysis_cap);
  return AFX_OK;
}

extern "C" int64_t afx_part_read(afx_partjob* j, int64_t begin, int64_t count, float* dst)
{
  if (!j || (!dst && count > 0) || count < 0) return AFX_ERR_ARG;
  afx_ctx* ctx = j->ctx;
  if (j->phase < 1) return afx_fail(ctx, AFX_ERR_STATE, "afx_part_read: afx_part_peak first");
  const long long ob = j->resampled ? j->part.out_begin : j->part.src_begin, oe = j->resampled ? j->part.out_end : j->part.src_end;
  const long long lo = std::max<long long>(begin, ob), hi = std::min<long long>(begin + count, oe);
  if (hi <= lo) return 0;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  CKP(cudaMemcpyAsync(dst + (lo - begin), (const float*)j->bufs.mono.p + (lo - ob), (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(mono)");
  CKP(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  return hi - lo;
}

extern "C" void afx_part_close(afx_partjob* j)
{
  if (!j) return;
  cudaSetDevice(j->ctx->device);
  if (j->ev_h2d) { cudaEventSynchronize(j->ev_h2d); cudaEventDestroy(j->ev_h2d); }
  cudaStreamSynchronize(j->ctx->stream);
  {
    std::lock_guard<std::mutex> lk(j->ctx->mu);
    j->release();
  }
  delete j;
}

extern "C" int afx_analyze_conditioned(afx_ctx* ctx, const afx_file* whole, const afx_part_sums* global, const float* mono,
                                       int64_t mono_begin, int64_t mono_count, afx_batch** out)
{
  if (!ctx || !whole || !global || !out) return afx_fail(ctx, AFX_ERR_ARG, "afx_analyze_conditioned: null argument");
  AfxCondInput in;
  memset(&in, 0, sizeof(in));
  in.mono = mono; in.mono_begin = mono_begin; in.mono_count = mono_count;
  memcpy(&in.inj.maxabs_bits, &global->maxabs, 4);
  in.inj.sumsq = global->sumsq;
  in.inj.first = clamp_first(global->first); in.inj.last = (int)global->last;
  for (int k = 0; k < 3; ++k) { in.inj.eff_first[k] = clamp_first(global->eff_first[k]); in.inj.eff_last[k] = (int)global->eff_last[k]; }
  int rc = afx_batch_create_impl(ctx, whole, 1, &in, out);
  if (rc != AFX_OK) return rc;
  afx_batch* b = *out;
  if ((rc = afx_batch_upload(b)) != AFX_OK || (rc = afx_batch_compute(b)) != AFX_OK ||
      (rc = afx_batch_download(b)) != AFX_OK || (rc = afx_batch_sync(b)) != AFX_OK) {
    afx_batch_free(b); *out = nullptr; return rc;
  }
  return AFX_OK;
}

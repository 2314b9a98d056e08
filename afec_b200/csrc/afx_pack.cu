// K9: the sink's BLOB images on the GPU.
//
// afec-ll.db stores every framed series as a msgpack array of float64 (SqliteSampleDescriptorPool.cpp:601-713, 890-954;
// msgpack-c pack.hpp: array headers 0x9X / 0xdc + u16 / 0xdd + u32, every double as 0xcb + 8 bytes big endian) -- 122
// BLOB columns per row, 0.2-0.9 MB per file.  Packing them is pure byte shuffling over arrays that already sit in HBM, so
// it is done here and the device -> host copy delivers rows the sink can bind without touching them:
//
//   packed[file]     the row's BLOBs back to back in column order: 22 framed scalars, the two onset series,
//                    then per framed vector its VVR blob followed by its 13 per-band statistic VR blobs
//   packed_off[file] AFX_N_BLOBS + 1 byte offsets of the blobs inside packed[file]
//
// A file's region is placed by the host from the frame-slot CAPACITIES (known before any sample is read); the blob
// offsets inside it follow the real frame counts and are written by the device.
#include "afx_common.cuh"
#include "../../include/afec_b200.h"

#define PK_T 256

__host__ __device__ __forceinline__ unsigned pk_hdr(unsigned n) { return n < 16u ? 1u : (n < 65536u ? 3u : 5u); }
__host__ __device__ __forceinline__ unsigned pk_vr(unsigned n) { return pk_hdr(n) + 9u * n; }
__host__ __device__ __forceinline__ unsigned pk_vvr(unsigned frames, unsigned nb) { return pk_hdr(frames) + frames * (pk_hdr(nb) + 9u * nb); }

// bytes of one file's packed region for F main and Fr rhythm frames (monotone in both: caps give an upper bound)
size_t afx_pack_region_bytes(int F, int Fr)
{
  static const unsigned nbv[AFX_N_FV] = { 14, 14, 14, 14, 14, 28, 14 };
  size_t n = (size_t)AFX_N_FS_MAIN * pk_vr((unsigned)F) + 2 * (size_t)pk_vr((unsigned)Fr);
  for (int v = 0; v < AFX_N_FV; ++v) n += pk_vvr((unsigned)F, nbv[v]) + (size_t)AFX_N_STATS * pk_vr(nbv[v]);
  return n;
}

__device__ __forceinline__ unsigned char* pk_put_hdr(unsigned char* p, unsigned n)
{
  if (n < 16u) { p[0] = (unsigned char)(0x90u | n); return p + 1; }
  if (n < 65536u) { p[0] = 0xdc; p[1] = (unsigned char)(n >> 8); p[2] = (unsigned char)n; return p + 3; }
  p[0] = 0xdd; p[1] = (unsigned char)(n >> 24); p[2] = (unsigned char)(n >> 16); p[3] = (unsigned char)(n >> 8); p[4] = (unsigned char)n;
  return p + 5;
}
__device__ __forceinline__ void pk_put_double(unsigned char* p, double v)
{
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  p[0] = 0xcb;
  p[1] = (unsigned char)(hi >> 24); p[2] = (unsigned char)(hi >> 16); p[3] = (unsigned char)(hi >> 8); p[4] = (unsigned char)hi;
  p[5] = (unsigned char)(lo >> 24); p[6] = (unsigned char)(lo >> 16); p[7] = (unsigned char)(lo >> 8); p[8] = (unsigned char)lo;
}

// grid (n_files, 32): y < 24 one framed scalar series, 24 <= y < 31 one framed vector (its VVR blob), y == 31 the 91
// statistic blobs of the vectors and the offsets table
__global__ void __launch_bounds__(PK_T) k_pack(AfxBatchDev B, AfxPackDev O)
{
  const int fi = blockIdx.x, y = blockIdx.y, tid = threadIdx.x;
  const AfxFile f = B.files[fi];
  unsigned* __restrict__ offs = O.blob_off + (size_t)fi * (AFX_N_BLOBS + 1);
  if (f.status != 0) { if (y == 31) for (int i = tid; i <= AFX_N_BLOBS; i += PK_T) offs[i] = 0u; return; }
  const unsigned F = (unsigned)B.state[fi].F, Fr = (unsigned)B.state[fi].Fr;
  const size_t TF = (size_t)B.TF;
  const int nbv[AFX_N_FV] = { 14, 14, 14, 14, 14, 28, 14 };
  const int fvo[AFX_N_FV] = { FV_RMS, FV_FLATNESS, FV_FLUX, FV_COMPLEXITY, FV_CONTRAST, FV_BANDS28, FV_CEPSTRUM };
  unsigned char* __restrict__ base = O.packed + O.file_off[fi];
  // start of blob `y` (and, for y == 31, of the first statistic blob of every vector): a short serial walk over the layout
  unsigned off = 0;
  if (y < AFX_N_FS) off = (unsigned)min(y, AFX_N_FS_MAIN) * pk_vr(F) + (unsigned)max(y - AFX_N_FS_MAIN, 0) * pk_vr(Fr);
  else {
    off = AFX_N_FS_MAIN * pk_vr(F) + 2u * pk_vr(Fr);
    for (int v = 0; v < y - AFX_N_FS && v < AFX_N_FV; ++v) off += pk_vvr(F, (unsigned)nbv[v]) + AFX_N_STATS * pk_vr((unsigned)nbv[v]);
  }
  if (y < AFX_N_FS) {                                         // VR[n]
    const unsigned n = (y < AFX_N_FS_MAIN) ? F : Fr;
    const double* __restrict__ x = (y < AFX_N_FS_MAIN) ? B.fs + (size_t)y * TF + f.frame_off : B.fsr + (size_t)(y - AFX_N_FS_MAIN) * B.TFr + f.rframe_off;
    unsigned char* p = base + off;
    if (tid == 0) pk_put_hdr(p, n);
    p += pk_hdr(n);
    for (unsigned i = tid; i < n; i += PK_T) pk_put_double(p + 9u * i, x[i]);
  } else if (y < AFX_N_FS + AFX_N_FV) {                       // VVR[F][nb]
    const int v = y - AFX_N_FS;
    const unsigned nb = (unsigned)nbv[v], hb = pk_hdr(nb), rec = hb + 9u * nb;
    const double* __restrict__ x = B.fv + (size_t)fvo[v] * TF + (size_t)f.frame_off * nb;
    unsigned char* p = base + off;
    if (tid == 0) pk_put_hdr(p, F);
    p += pk_hdr(F);
    const unsigned total = F * nb;
    for (unsigned i = tid; i < total; i += PK_T) {
      const unsigned fr = i / nb, b = i - fr * nb;
      unsigned char* q = p + fr * rec;
      if (b == 0) pk_put_hdr(q, nb);
      pk_put_double(q + hb + 9u * b, x[i]);
    }
  } else {                                                    // 7 x 13 statistic blobs VR[nb] + the offsets table
    const double* __restrict__ st = B.stats + (size_t)fi * AFX_N_SERIES * AFX_N_STATS;
    unsigned o = AFX_N_FS_MAIN * pk_vr(F) + 2u * pk_vr(Fr);
    int series = AFX_N_FS, blob = AFX_N_FS;
    if (tid == 0) for (int s = 0; s < AFX_N_FS; ++s) offs[s] = (unsigned)min(s, AFX_N_FS_MAIN) * pk_vr(F) + (unsigned)max(s - AFX_N_FS_MAIN, 0) * pk_vr(Fr);
    for (int v = 0; v < AFX_N_FV; ++v) {
      const unsigned nb = (unsigned)nbv[v];
      if (tid == 0) offs[blob] = o;
      o += pk_vvr(F, nb); ++blob;
      for (int k = 0; k < AFX_N_STATS; ++k) {
        unsigned char* p = base + o;
        if (tid == 0) { offs[blob] = o; pk_put_hdr(p, nb); }
        if ((unsigned)tid < nb) pk_put_double(p + pk_hdr(nb) + 9u * tid, st[(size_t)(series + tid) * AFX_N_STATS + k]);
        o += pk_vr(nb); ++blob;
      }
      series += (int)nb;
    }
    if (tid == 0) offs[AFX_N_BLOBS] = o;
  }
}

void afx_launch_pack(const AfxBatchDev& B, const AfxPackDev& O, cudaStream_t s, long long* launches)
{
  if (B.n_files <= 0) return;
  k_pack<<<dim3(B.n_files, 32), PK_T, 0, s>>>(B, O); ++*launches;
}

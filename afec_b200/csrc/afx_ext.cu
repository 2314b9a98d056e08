// K10 (extension, BASELINE configs[2]): MFCC-13 over 40 mel filters + chroma-12 per main frame.
//
// The reference has neither (its cepstrum is MFCC-14 over 14 filters, it has no chroma); north_star
// item (4) asks for them as an extension validated against a CPU restatement (tests/ext_reference.py): they are the one
// GEMM-shaped piece of the path -- per frame a [1 x 1024] x [1024 x 52] contraction of the magnitude spectrum with two
// fixed weight matrices, followed by a tiny epilogue:
//   mel     E_m  = sum_k Mag[k]   W_mel[m][k]      40 triangular equal-gain filters, 20 Hz .. 15.5 kHz on the mel scale
//                                                  1127 ln(1 + f / 700) (LibXtract's construction, init.c:237-378, laid
//                                                  over all 1024 bins)
//   mfcc    c_n  = sum_m log(max(E_m, 2e-42)) cos(pi n (m + 1/2) / 40),  n = 0..12   (unnormalised DCT-II, vector.c:372-391)
//   chroma  C_c  = sum_k Mag[k]^2 W_chr[c][k]      bins 65.4 Hz .. 8372 Hz, each shared linearly between its two nearest
//                                                  semitones (pitch = 69 + 12 log2(f / 440), class = pitch mod 12)
//           chroma[c] = C_c / max_c C_c,  chroma_index = argmax_c C_c (first maximum)
//
// Two implementations of the contraction, chosen per context (AFX_EXT_TENSOR=1 in the environment selects the second):
//   k_ext_fp32   FP32 FMA tile: thread per frame, 52 FP32 accumulators, the spectrum reaches the threads through a
//                transposed shared-memory tile, the weights are broadcast float4 loads
//   k_ext_tc     3xTF32 on the 5th-generation tensor cores: tcgen05.mma kind::tf32, M = 128 frames per CTA, accumulators
//                in TMEM.  Every FP32 operand is split x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and
//                A W ~= A_hi W_hi + A_lo W_hi + A_hi W_lo (the dropped lo x lo term is 2^-22 relative): three MMAs per
//                K step.  The tensor core adds into its FP32 accumulator with TRUNCATION, so the error grows with the number
//                of MMAs chained into one accumulator (1.8e-5 on an MFCC coefficient with a single one): the K steps are dealt
//                round robin to EIGHT accumulator sets (all 512 TMEM columns) that the epilogue adds in FP64, small terms
//                first (0.47 x tolerance).  Operands are staged by the CTA's threads in the canonical K-major
//                SWIZZLE_128B shared-memory layout (8-row x 128-byte atoms); one thread issues the MMAs and commits
//                them to an mbarrier; the epilogue reads the accumulators back with tcgen05.ld.
// Both feed the same FP64 epilogue.  tests/test_gpu_ext.py compares both with the FP64 restatement: north_star's rule is
// "tensor cores only if the split-precision scheme stays inside tolerance".
#include "afx_common.cuh"
#include "../../include/afec_b200.h"

#include <cstdint>

#define EX_NMEL 40
#define EX_NCHR 12
#define EX_NOUT 52          // 40 mel + 12 chroma
#define EX_NMFCC 13
#define EX_T 128            // frames per CTA (both kernels)

// FP64 epilogue of one frame: acc[0..39] mel energies, acc[40..51] chroma energies
__device__ __forceinline__ void ext_epilogue(const AfxBatchDev& B, const AfxExtDev& X, int slot, const float* acc)
{
  double lg[EX_NMEL];
#pragma unroll
  for (int m = 0; m < EX_NMEL; ++m) { const double e = (double)acc[m]; lg[m] = log(e < 2e-42 ? 2e-42 : e); }
  double* __restrict__ mf = X.mfcc + (size_t)slot * EX_NMFCC;
#pragma unroll 1
  for (int n = 0; n < EX_NMFCC; ++n) {
    double a = 0.0;
#pragma unroll
    for (int m = 0; m < EX_NMEL; ++m) a = fma(lg[m], __ldg(X.dct + n * EX_NMEL + m), a);
    mf[n] = a;
  }
  double mx = 0.0; int arg = 0;
#pragma unroll
  for (int c = 0; c < EX_NCHR; ++c) { const double v = (double)acc[EX_NMEL + c]; if (v > mx) { mx = v; arg = c; } }
  double* __restrict__ ch = X.chroma + (size_t)slot * EX_NCHR;
#pragma unroll
  for (int c = 0; c < EX_NCHR; ++c) ch[c] = (mx > 0.0) ? (double)acc[EX_NMEL + c] / mx : 0.0;
  X.chroma_index[slot] = (double)arg;
}

// ---------------------------------------------------------------------------------------------------------
// FP32 FMA tile
__global__ void __launch_bounds__(EX_T) k_ext_fp32(AfxBatchDev B, AfxExtDev X)
{
  __shared__ float As[32][EX_T + 1];          // [bin][frame]
  __shared__ float4 Ws[32][EX_NOUT / 4];      // [bin][output / 4]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int rel0 = blockIdx.x * EX_T;
  const int rel = rel0 + tid;
  const bool in_range = rel < B.g_slots;
  const int slot = B.slot0 + (in_range ? rel : rel0);
  const int fi = B.slot_file[slot];
  const bool live = in_range && B.files[fi].status == 0 && (slot - B.files[fi].frame_off) < B.state[fi].F;
  const int last_row = B.g_slots - 1;
  float acc[EX_NOUT];
#pragma unroll
  for (int q = 0; q < EX_NOUT; ++q) acc[q] = 0.0f;
  for (int kb = 0; kb < AFX_NBIN; kb += 32) {
    // spectrum tile: warp w loads rows w, w + 4, ... (256 contiguous bytes each), stored transposed as float
#pragma unroll 8
    for (int r = wid; r < EX_T; r += EX_T / 32) {
      const int rc = min(rel0 + r, last_row);
      As[lane][r] = (float)B.mag[(size_t)rc * AFX_NBIN + kb + lane];
    }
    for (int i = tid; i < 32 * (EX_NOUT / 4); i += EX_T)
      Ws[i / (EX_NOUT / 4)][i % (EX_NOUT / 4)] = __ldg(reinterpret_cast<const float4*>(X.w_kmajor + (size_t)kb * EX_NOUT) + i);
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      const float x = As[j][tid], x2 = x * x;
#pragma unroll
      for (int q = 0; q < EX_NMEL / 4; ++q) {
        const float4 w = Ws[j][q];
        acc[4 * q] = fmaf(x, w.x, acc[4 * q]); acc[4 * q + 1] = fmaf(x, w.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(x, w.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x, w.w, acc[4 * q + 3]);
      }
#pragma unroll
      for (int q = EX_NMEL / 4; q < EX_NOUT / 4; ++q) {
        const float4 w = Ws[j][q];
        acc[4 * q] = fmaf(x2, w.x, acc[4 * q]); acc[4 * q + 1] = fmaf(x2, w.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(x2, w.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x2, w.w, acc[4 * q + 3]);
      }
    }
    __syncthreads();
  }
  if (live) ext_epilogue(B, X, slot, acc);
}

// ---------------------------------------------------------------------------------------------------------
// 3xTF32 on tcgen05.  PTX wrappers (sm_100a); SASS: UTCHMMA / LDTM.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float to_tf32(float x)
{
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor: start address >> 4, leading byte
// offset 1 (unused for swizzled K-major), stride byte offset = 1024 B between 8-row groups, version 1, layout type 2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
    :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
// bounded wait: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
  const uint32_t a = smem_u32(bar);
  for (uint32_t spins = 0;; ++spins) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(a), "r"(phase) : "memory");
    if (done) return;
    if (spins > (1u << 26)) __trap();
  }
}

// Shared memory of k_ext_tc, per stage (K step of 32 bins = one 128-byte swizzle row):
//   A_mag_hi, A_mag_lo, A_pow_hi, A_pow_lo : 128 rows x 128 B = 16 KB each
//   W_mel_hi, W_mel_lo : 48 rows x 128 B = 6 KB each;  W_chr_hi, W_chr_lo : 16 rows x 128 B = 2 KB each
#define TC_A_BYTES (EX_T * 128)
#define TC_WM_ROWS 48
#define TC_WC_ROWS 16
#define TC_STAGE_BYTES (4 * TC_A_BYTES + 2 * TC_WM_ROWS * 128 + 2 * TC_WC_ROWS * 128)
#define TC_STAGES 2
#define TC_NACC 8           // TMEM accumulator sets (x 64 columns = the SM's 512)
#define TC_SMEM (TC_STAGES * TC_STAGE_BYTES + 1024)

// byte offset of element (row, col) of a K-major SWIZZLE_128B tile whose rows hold 32 x 4 bytes: 8-row atoms of 1024 B,
// the 16-byte chunk index XORed with the row inside the atom
__device__ __forceinline__ uint32_t sw128_off(int row, int col)
{
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 2) ^ (row & 7)) & 7) << 4) + ((col & 3) << 2));
}

__global__ void __launch_bounds__(EX_T) k_ext_tc(AfxBatchDev B, AfxExtDev X)
{
  extern __shared__ unsigned char tc_raw[];
  __shared__ uint64_t bar_mma[TC_STAGES];      // "the MMAs that read this stage are done"
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)tc_raw + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 1024-byte aligned
  const int rel0 = blockIdx.x * EX_T;
  const int rel = rel0 + tid;
  const bool in_range = rel < B.g_slots;
  const int slot = B.slot0 + (in_range ? rel : rel0);
  const int fi = B.slot_file[slot];
  const bool live = in_range && B.files[fi].status == 0 && (slot - B.files[fi].frame_off) < B.state[fi].F;
  const int last_row = B.g_slots - 1;

  if (tid == 0) { for (int s = 0; s < TC_STAGES; ++s) mbar_init(&bar_mma[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  // TMEM: TC_NACC accumulator sets of 64 columns (mel at column 0 (48 wide), chroma at column 48 (16 wide) of a set).  The
  // tensor core adds into its FP32 accumulators with truncation, so the error of ONE accumulator grows linearly with the
  // number of MMAs chained into it (measured: 3e-6 relative after the 384 of a whole row); K step `it` therefore goes to
  // set it % TC_NACC (48 MMAs per set) and the sets are added in FP64 by the epilogue.
  if (wid == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t idesc_mel = umma_idesc_tf32(EX_T, TC_WM_ROWS), idesc_chr = umma_idesc_tf32(EX_T, TC_WC_ROWS);

  for (int it = 0; it < AFX_NBIN / 32; ++it) {
    const int s = it % TC_STAGES, kb = it * 32;
    unsigned char* st = smem + (size_t)s * TC_STAGE_BYTES;
    if (it >= TC_STAGES) mbar_wait(&bar_mma[s], (uint32_t)((it / TC_STAGES - 1) & 1));     // the stage's previous MMAs have read it
    // ---- stage the operands: spectrum rows (warp w: rows w, w + 4, ...; lane = bin) and the weight rows -------------
#pragma unroll 4
    for (int r = wid; r < EX_T; r += EX_T / 32) {
      const int rc = min(rel0 + r, last_row);
      const float x = (float)B.mag[(size_t)rc * AFX_NBIN + kb + lane], p = x * x;
      const float xh = to_tf32(x), xl = to_tf32(x - xh), ph = to_tf32(p), pl = to_tf32(p - ph);
      const uint32_t o = sw128_off(r, lane);
      *reinterpret_cast<float*>(st + o) = xh;
      *reinterpret_cast<float*>(st + TC_A_BYTES + o) = xl;
      *reinterpret_cast<float*>(st + 2 * TC_A_BYTES + o) = ph;
      *reinterpret_cast<float*>(st + 3 * TC_A_BYTES + o) = pl;
    }
    {
      unsigned char* wm = st + 4 * TC_A_BYTES; unsigned char* wc = wm + 2 * TC_WM_ROWS * 128;
      for (int i = tid; i < (TC_WM_ROWS + TC_WC_ROWS) * 32; i += EX_T) {
        const int row = i >> 5, col = i & 31;                      // weight row = output index, K-major
        const bool mel = row < TC_WM_ROWS;
        const int out = mel ? row : EX_NMEL + (row - TC_WM_ROWS);
        const bool valid = mel ? (row < EX_NMEL) : (row - TC_WM_ROWS < EX_NCHR);
        const float w = valid ? __ldg(X.w_nmajor + (size_t)out * AFX_NBIN + kb + col) : 0.0f;
        const float wh = to_tf32(w), wl = to_tf32(w - wh);
        const uint32_t o = sw128_off(mel ? row : row - TC_WM_ROWS, col);
        unsigned char* base = mel ? wm : wc;
        const uint32_t half = (uint32_t)(mel ? TC_WM_ROWS : TC_WC_ROWS) * 128u;
        *reinterpret_cast<float*>(base + o) = wh;
        *reinterpret_cast<float*>(base + half + o) = wl;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();
    // ---- one thread issues the stage's MMAs: 4 K-steps of 8 x (hi.hi + lo.hi + hi.lo) x (mel, chroma) --------------
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_mh = smem_u32(st), a_ml = a_mh + TC_A_BYTES, a_ph = a_mh + 2 * TC_A_BYTES, a_pl = a_mh + 3 * TC_A_BYTES;
      const uint32_t w_mh = a_mh + 4 * TC_A_BYTES, w_ml = w_mh + TC_WM_ROWS * 128, w_ch = w_mh + 2 * TC_WM_ROWS * 128, w_cl = w_ch + TC_WC_ROWS * 128;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t ko = (uint32_t)k * 32u;                     // 8 tf32 = 32 bytes along K inside the swizzle row
        const uint32_t first = (it >= TC_NACC || k > 0) ? 1u : 0u;   // 0: overwrite (the set's first MMA), 1: accumulate
        const uint32_t d = tmem + 64u * (uint32_t)(it % TC_NACC);
        // small terms first: lo.hi and hi.lo are ~2^-11 of hi.hi
        umma_tf32(d, umma_desc_sw128(a_ml + ko), umma_desc_sw128(w_mh + ko), idesc_mel, first);
        umma_tf32(d, umma_desc_sw128(a_mh + ko), umma_desc_sw128(w_ml + ko), idesc_mel, 1u);
        umma_tf32(d, umma_desc_sw128(a_mh + ko), umma_desc_sw128(w_mh + ko), idesc_mel, 1u);
        umma_tf32(d + 48, umma_desc_sw128(a_pl + ko), umma_desc_sw128(w_ch + ko), idesc_chr, first);
        umma_tf32(d + 48, umma_desc_sw128(a_ph + ko), umma_desc_sw128(w_cl + ko), idesc_chr, 1u);
        umma_tf32(d + 48, umma_desc_sw128(a_ph + ko), umma_desc_sw128(w_ch + ko), idesc_chr, 1u);
      }
      // arrives on the barrier when every MMA issued so far has completed (implies tcgen05.fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar_mma[s])) : "memory");
    }
  }
  // ---- epilogue: wait for the last stage's MMAs, read the accumulators (thread = TMEM lane = frame row) ----------------
  {
    constexpr int last = AFX_NBIN / 32 - 1;
    mbar_wait(&bar_mma[last % TC_STAGES], (uint32_t)((last / TC_STAGES) & 1));
    if (TC_STAGES > 1) mbar_wait(&bar_mma[(last - 1) % TC_STAGES], (uint32_t)(((last - 1) / TC_STAGES) & 1));
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float acc[EX_NOUT];
  {
    double sum[EX_NOUT];
#pragma unroll
    for (int q = 0; q < EX_NOUT; ++q) sum[q] = 0.0;
#pragma unroll 1
    for (int a = 0; a < TC_NACC; ++a) {
      const uint32_t taddr = tmem + ((uint32_t)(wid * 32) << 16) + 64u * (uint32_t)a;   // lane field in bits 31:16: warp w owns TMEM lanes 32 w .. 32 w + 31
      uint32_t v[64];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[16 * c]), "=r"(v[16 * c + 1]), "=r"(v[16 * c + 2]), "=r"(v[16 * c + 3]), "=r"(v[16 * c + 4]), "=r"(v[16 * c + 5]),
                       "=r"(v[16 * c + 6]), "=r"(v[16 * c + 7]), "=r"(v[16 * c + 8]), "=r"(v[16 * c + 9]), "=r"(v[16 * c + 10]), "=r"(v[16 * c + 11]),
                       "=r"(v[16 * c + 12]), "=r"(v[16 * c + 13]), "=r"(v[16 * c + 14]), "=r"(v[16 * c + 15])
                     : "r"(taddr + 16u * c));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < EX_NMEL; ++q) sum[q] += (double)__uint_as_float(v[q]);
#pragma unroll
      for (int q = 0; q < EX_NCHR; ++q) sum[EX_NMEL + q] += (double)__uint_as_float(v[48 + q]);
    }
#pragma unroll
    for (int q = 0; q < EX_NOUT; ++q) acc[q] = (float)sum[q];
  }
  if (live) ext_epilogue(B, X, slot, acc);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

void afx_launch_ext(const AfxBatchDev& B, const AfxExtDev& X, bool tensor, cudaStream_t s, long long* launches)
{
  if (B.g_slots <= 0) return;
  const int grid = (B.g_slots + EX_T - 1) / EX_T;
  if (tensor) {
    cudaFuncSetAttribute(k_ext_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    k_ext_tc<<<grid, EX_T, TC_SMEM, s>>>(B, X);
  } else {
    k_ext_fp32<<<grid, EX_T, 0, s>>>(B, X);
  }
  ++*launches;
}

// Shared pieces of the tcgen05 kernel family (render_tc_fwd.cu, render_tc_bwd.cu): global workspace
// layout, operand-image geometry, abortable mbarrier waits, bulk global->shared copies.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma.cuh"

// ---- operand images --------------------------------------------------------------------------
// Every bf16 operand lives in the SWIZZLE_NONE canonical layout of 8x8 core matrices (128 B each).
// "Row-group-major" image of R rows x C cols:  off(r,c) = (r/8)*128 + (c/8)*(R/8)*128 + (r%8)*16 + (c%8)*2
//   * weights  W_l [k][n]      : R = K_l (padded), C = 128      (B operand; MN-major in fwd, K-major in dgrad)
//   * samples  X   [s][c]      : R = 128 samples,  C = 128|32   (A operand K-major in fwd/dgrad, A/B MN-major in wgrad)
// so a 128-row sample image has RS = 128 B, CS = 2048 B.
#define TC_IMG_RS 128u
#define TC_SIMG_CS 2048u                 // column-group stride of a 128-row sample image
#define TC_SIMG_BYTES 32768u             // 128 x 128 bf16
#define TC_FIMG_BYTES 8192u              // 128 x 32 bf16 (features, 21 real columns)

__host__ __device__ constexpr uint32_t tc_layer_K(int l) { return l == 0 ? 32u : (l == 3 ? 160u : 128u); }
__host__ __device__ constexpr uint32_t tc_plane_bytes(int l) { return tc_layer_K(l) * 128u * 2u; }   // one bf16 plane
__host__ __device__ constexpr uint32_t tc_stage_bytes(int l) { return 2u * tc_plane_bytes(l); }      // [hi | lo]
__host__ __device__ constexpr uint32_t tc_stage_off(int l) {                                         // in the weight block
  return l == 0 ? 0u : (l == 1 ? 16384u : (l == 2 ? 81920u : 147456u));
}
// Saved activations per frame (bh_tc_fwd with acts), PL = 1 or 2 bf16 planes (hi | hi+lo):
//   h images    [plane][l=0..3][tile][32 KB]     plane stride n_pad*1024, layer stride n_pad*256
//   feat images [plane][tile][8 KB] at PL*n_pad*1024 (hi plane: col TC_ONES_COL = 1)
// Backward scratch per frame: delta images [plane][l][tile][32 KB], then aux images [tile][4 KB]
// (col 0 = hi, col 1 = lo bf16 part of d loss / d o).
#define TC_ONES_COL 21
#define TC_AIMG_BYTES 4096u              // 128 x 16 bf16
//   relu masks  [tile][l=0..3][column half][row][2 words] at PL*n_pad*1088: one bit per activation (h > 0), 64 B per
//               sample; all the dgrad chain needs of the activations (tc_mask_bits / tc_mask_expand below)
__host__ __device__ inline size_t tc_acts_bytes_per_frame(int n_pad, int PL) { return (size_t)n_pad * ((size_t)PL * (1024u + 64u) + 64u); }
__host__ __device__ inline size_t tc_mask_off(int n_pad, int PL) { return (size_t)n_pad * (size_t)PL * (1024u + 64u); }
#define TC_MASK_TILE_BYTES 8192u         // 4 layers x 2 halves x 128 rows x 8 B
__host__ __device__ inline size_t tc_delta_bytes_per_frame(int n_pad, int PL) { return (size_t)n_pad * ((size_t)PL * 1024u + 32u); }
#define TC_W_BYTES 229376u               // 16K + 64K + 64K + 80K
#define TC_BIMG_BYTES 4096u              // forward only: bias image of layer 1 / 2 (16 x 128 fp16) at TC_WS_W + TC_W_BYTES
#define TC_STAGE_MAX 81920u

// ---- global workspace (bh_tc_ws_bytes) -------------------------------------------------------
//   [0,256)        int32 status words (0 = ok): [0..5) flags of the CURRENT step (reset by tc_prepare_weights_kernel;
//                  the guarded Adam kernels read them), [8..55) optional cycle counters, [55] magic, [56..61) the same
//                  flags STICKY: set by any step since the last bhnerf_workspace_status call, which clears them
//   [256,4096)     fp32 constants: b0,b1,b2,b3 (4x128) | W4 (128) | b4 (1)
//   [4096, +224K)  fp16 weight images of the forward, per layer [hi plane | lo plane]; then 2 x 4 KB bias images
//   [262144,+224K) bf16 weight images of the dgrad chain (cotangents need the fp32 exponent range), same layout
#define TC_WS_STATUS 0u
#define TC_STATUS_MAGIC_WORD 55
#define TC_STATUS_MAGIC 0x62683230
#define TC_STATUS_STICKY 56
#define TC_WS_CONST 256u
#define TC_WS_W 4096u
#define TC_WS_WB 262144u
#define TC_CONST_FLOATS 768              // 642 used
#define TC_C_B(l) ((l) * 128)
#define TC_C_W4 512
#define TC_C_B4 640
#define TC_C_RANGE 641                   // != 0: the weights allow |h| > 65504, the forward epilogue tracks the range
#define TC_C_DOUTMAX 700                 // uint32 bits of max |d loss / d o| of the current backward launch (tc_dout_kernel)
#define TC_C_SCHED 701                   // next tile pair of the fused backward's dynamic schedule (zeroed with DOUTMAX)

// Optional cycle accounting (-DBH_TC_TIMING): block 0 writes, per role, the cycles spent waiting vs working into
// the workspace status words [8..20) (two int32 per counter).  Compiled out by default.
#ifdef BH_TC_TIMING
#define BH_TIMING_DECL(v) long long v = 0; long long _t0_##v = 0; (void)_t0_##v;
#define BH_TIMING_BEGIN _bh_t0 = clock64();
#define BH_TIMING_END(v) v += clock64() - _bh_t0;
#define BH_TIMING_T0 long long _bh_t0 = 0;
#define BH_TIMING_STORE_B(st, i, v, blk) if (blockIdx.x == (blk)) { (st)[i] = (int)(v & 0xffffffffll); (st)[(i) + 1] = (int)(v >> 32); }
#define BH_TIMING_STORE(st, i, v) BH_TIMING_STORE_B(st, i, v, 0)
#else
#define BH_TIMING_T0
#define BH_TIMING_DECL(v)
#define BH_TIMING_BEGIN
#define BH_TIMING_END(v)
#define BH_TIMING_STORE(st, i, v)
#define BH_TIMING_STORE_B(st, i, v, blk)
#endif

namespace tc {
using namespace umma;

// raise health flag k of this step and its sticky copy
__device__ __forceinline__ void raise_flag(int* status, int k) {
  atomicExch(status + k, 1);
  atomicExch(status + TC_STATUS_STICKY + k, 1);
}

// ---- abortable waits: a wedged pipeline flags an error and drains instead of hanging the GPU ----
struct Abort { volatile int* flag; };
__device__ __forceinline__ bool wait(uint64_t* bar, uint32_t parity, const Abort& ab) {
  for (uint32_t i = 0; i < (1u << 22); ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((i & 255u) == 255u && *ab.flag) return false;
  }
  *ab.flag = 1;
  return false;
}

__device__ __forceinline__ bool wait_cluster(uint64_t* bar, uint32_t parity, const Abort& ab) {   // acquire.cluster
  for (uint32_t i = 0; i < (1u << 22); ++i) {
    if (mbar_try_wait_cluster(bar, parity)) return true;
    if ((i & 255u) == 255u && *ab.flag) return false;
  }
  *ab.flag = 1;
  return false;
}

__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// the same with an L2 eviction-priority hint (createpolicy): streamed-once data should not push the cotangent ring out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

// pull `bytes` (multiple of 16) of global memory into L2 ahead of their use; one instruction, no destination
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

// ---- cotangent scale of the one-plane backward (fp16 operands) ----
// The saved activations of the one-plane plan are the forward's fp16 hi plane, so the cotangents must be fp16 too
// (tcgen05.mma takes one 16-bit format per instruction).  Their range is set by d loss/d o, which spans many decades
// across samples: all of them are multiplied by ONE power of two s, chosen from max|dout| of the launch so that
// max|dout| * s lies in [2, 4) -- the chain is linear, so s rides through every layer and the wgrad flush multiplies by
// 1/s (exact).  Contributions below 2^-26 of the largest flush to zero, above 2^-16 of it they keep all 11 bits; three
// layers of worst-case growth (row sums of |W|) stay below 65504 for any weights with row sums under ~25 -- beyond that
// the non-finite gradient is flagged (status word 4), not hidden.
__device__ __forceinline__ float tc_grad_scale(uint32_t max_bits, float& inv) {
  const int e = (int)((max_bits >> 23) & 0xFFu) - 127;          // floor(log2 max|dout|)
  if (max_bits == 0u || e < -100 || e > 100) { inv = 1.f; return 1.f; }
  inv = __uint_as_float((uint32_t)(127 + e - 1) << 23);
  return __uint_as_float((uint32_t)(127 + 1 - e) << 23);
}

#define TC_RZ_UNBIAS 1.000352215f       // 1 + 2^-11 * 0.5 / ln 2: mean relative truncation of cvt.rz to fp16

// ---- ReLU bit masks ----
// One 32-bit word covers 32 consecutive columns = 16 packed bf16 pairs.  Pair j (columns 2j, 2j+1) keeps its two bits
// at positions p and p + 16 with p = 8*(j&1) + 7 - (j>>1): after a left shift by (j>>1) they are the sign bits of
// bytes (j&1) and (j&1)+2, and ONE prmt with sign replication expands them to the 0xffff / 0x0000 halves that mask
// a packed pair -- the same two instructions per pair as a compare against the saved activation, without loading it.
__device__ __forceinline__ constexpr uint32_t tc_mask_pair_bits(int j) {
  return (1u << (8 * (j & 1) + 7 - (j >> 1))) | (1u << (8 * (j & 1) + 7 - (j >> 1) + 16));
}
// bits of pair j from a packed bf16 pair of (relu'd) activations
__device__ __forceinline__ uint32_t tc_mask_bits(uint32_t h2, int j) {
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&h2), __float2bfloat162_rn(0.f)) & tc_mask_pair_bits(j);
}
// same from a packed fp16 pair (one-plane plan: the saved activation IS the forward's fp16 hi plane)
__device__ __forceinline__ uint32_t tc_mask_bits_f16(uint32_t h2, int j) {
  return __hgt2_mask(*reinterpret_cast<const __half2*>(&h2), __float2half2_rn(0.f)) & tc_mask_pair_bits(j);
}
// 0xffff per half of pair j where the activation was positive
__device__ __forceinline__ uint32_t tc_mask_expand(uint32_t word, int j) {
  uint32_t r;
  const uint32_t t = word << (j >> 1);
  if (j & 1) asm("prmt.b32 %0, %1, 0, 0xBB99;" : "=r"(r) : "r"(t));
  else asm("prmt.b32 %0, %1, 0, 0xAA88;" : "=r"(r) : "r"(t));
  return r;
}
// this thread's two mask words of layer l (row = sample row of the tile, half = 64-column half)
__device__ __forceinline__ size_t tc_mask_word_off(int tile, int l, int half, int row) {
  return (size_t)tile * TC_MASK_TILE_BYTES + (size_t)((l * 2 + half) * 128 + row) * 8u;
}

__device__ __forceinline__ uint32_t sample_img_off(int row, int colgroup) {   // 16-byte chunk of 8 columns
  return (uint32_t)(row >> 3) * TC_IMG_RS + (uint32_t)colgroup * TC_SIMG_CS + (uint32_t)(row & 7) * 16u;
}

// split 8 fp32 values into bf16 hi / lo chunks (16 B each)
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
    l[j] = pack_bf16x2(x[2 * j] - bf16_lo(h[j]), x[2 * j + 1] - bf16_hi(h[j]));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// same split into fp16 planes: 11+11 significand bits, |x - hi - lo| <= max(2^-22 |x|, 2^-25)
__device__ __forceinline__ void split8_f16(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = pack_f16x2(x[2 * j], x[2 * j + 1]);
    l[j] = pack_f16x2(x[2 * j] - f16_lo(h[j]), x[2 * j + 1] - f16_hi(h[j]));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace tc

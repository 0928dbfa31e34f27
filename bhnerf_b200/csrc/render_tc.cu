// tcgen05 forward render kernel (sm_100a): warp -> posenc -> 4x128 MLP on the 5th-gen tensor cores
// (fp16 x3 split operands, fp32 accumulate in TMEM) -> sigmoid(o-10) -> masked emission.
// Reference semantics: bhnerf/network.py:18-64 (MLP), :98-122 (posenc), :191-237 (NeRF_Predictor.__call__),
// bhnerf/emission.py:143-211 (velocity_warp_coords).  DESIGN.md s4 describes the pipeline.
//
// One persistent CTA per SM, 20 warps:
//   warps 0-7   epilogue of tile slot 0: (column half) x (TMEM lane quadrant = warp%4); thread = one sample row x 64 columns
//   warps 8-15  epilogue of tile slot 1
//   warp  16    MMA issuer (one elected lane issues every tcgen05.mma / tcgen05.commit)
//   warp  17    weight producer (cp.async.bulk of the next layer's [hi|lo] weight images into a 2-stage ring)
//   warps 18-19 idle (they complete the fifth warpgroup for setmaxnreg)
// Two 128-sample tiles are in flight per CTA and ping-pong: while the tensor core runs layer l of one tile,
// the CUDA cores run the ReLU + hi/lo-split epilogue of the other.  Activations never leave the SM:
// D (fp32) is read from TMEM with tcgen05.ld and the next layer's A operand is written back to TMEM with
// tcgen05.st (TS-form MMA); only the 21 posenc features go through shared memory (SS-form, K-major).
// The features of a slot's NEXT tile are computed while its current tile waits for the tensor core (double-buffered
// feature images), so the ~3 k cycles of warp + posenc are off the tile's serial chain.
#include <cuda_fp16.h>
#include <type_traits>
#include <stdlib.h>
#include "tc_common.cuh"

using namespace tc;

namespace {

// 16 epilogue warps + MMA issuer + weight producer + 2 idle warps that complete the fifth warpgroup (setmaxnreg is a
// warpgroup-wide instruction).  Registers: every SM sub-partition hosts 4 epilogue warps + 1 other warp; the kernel launches
// at 96 per thread (5 x 96 = 480 per lane of a sub-partition, the pool setmaxnreg redistributes) and rebalances to
// 4 x 104 + 64: the epilogue holds 64 accumulator values + both operand planes and spilled at 96.
constexpr int kThreads = 640;
constexpr int kEpiRegs = 104, kAuxRegs = 64;
constexpr int kMmaWarp = 16;
// forward weight images in the workspace (fp16, [hi plane | lo plane] each), tc_prepare_weights_kernel:
//   L0 (K = 32: 21 features, 2 bias rows) | L1 | L2 | L3 rows 0..127 (hidden part) | L3 rows 128..159 (skip part, K = 32)
// The skip part is its own image so that it can stay resident in shared memory: the ring stages then need 64 KB (+ a 4 KB
// bias image) instead of 80 KB, which pays for the second set of feature images.
__host__ __device__ constexpr uint32_t fw_K(int j) { return (j == 0 || j == 4) ? 32u : 128u; }
__host__ __device__ constexpr uint32_t fw_plane(int j) { return fw_K(j) * 128u * 2u; }
__host__ __device__ constexpr uint32_t fw_off(int j) {
  return j == 0 ? 0u : (j == 1 ? 16384u : (j == 2 ? 81920u : (j == 3 ? 147456u : 212992u)));
}
static_assert(fw_off(4) + 2u * fw_plane(4) == TC_W_BYTES, "forward weight block");
constexpr uint32_t FW_STAGE = 65536u + TC_BIMG_BYTES;              // one ring stage: [hi | lo] (<= 64 KB) + bias image
constexpr uint32_t SM_WSTAGE = 0;                                  // 2 ring stages
constexpr uint32_t SM_WSKIP = 2 * FW_STAGE;                        // resident [hi 8K | lo 8K] skip rows of layer 3
constexpr uint32_t SM_FEAT = SM_WSKIP + 2 * fw_plane(4);           // [slot][buffer][hi 8K | lo 8K]
constexpr uint32_t SM_CONST = SM_FEAT + 2 * 2 * 2 * TC_FIMG_BYTES; // 768 floats
constexpr uint32_t SM_BARS = SM_CONST + TC_CONST_FLOATS * 4;       // mbarriers + tmem base + abort flags
constexpr uint32_t SM_PART = SM_BARS + 256;                        // [slot][row] partial of the last layer
constexpr uint32_t SM_TOTAL2 = SM_PART + 2 * 128 * 4;
static_assert(SM_TOTAL2 <= 232448, "forward shared memory");

enum { BAR_WFULL = 0, BAR_WEMPTY = 2, BAR_AREADY = 4, BAR_DREADY = 6, BAR_FREADY = 8, BAR_DFREE = 10, BAR_SKIP = 12, BAR_COUNT = 13 };

// ------------------------------------------------------------------------------------------------
// weight images: fp32 params -> bf16 hi/lo planes in the canonical UMMA layout (tc_common.cuh)
// ------------------------------------------------------------------------------------------------
__global__ void tc_prepare_weights_kernel(const float* __restrict__ params, uint8_t* __restrict__ ws) {
  int l = blockIdx.y;
  const int K = (int)tc_layer_K(l);
  int idx = blockIdx.x * blockDim.x + threadIdx.x;        // k*128 + n
  if (blockIdx.y == 4) {                                  // fp32 constants + status reset
    float* c = (float*)(ws + TC_WS_CONST);
    if (idx < 128) {
      c[TC_C_B(0) + idx] = params[OFF_B0 + idx]; c[TC_C_B(1) + idx] = params[OFF_B1 + idx];
      c[TC_C_B(2) + idx] = params[OFF_B2 + idx]; c[TC_C_B(3) + idx] = params[OFF_B3 + idx];
      c[TC_C_W4 + idx] = params[OFF_W4 + idx];
    }
    if (idx == 0) c[TC_C_B4] = params[OFF_B4];
    // per-step flags and cycle counters are reset; the sticky flags [56..64) survive until bhnerf_workspace_status reads
    // them (a workspace seen for the first time -- no magic word -- has them zeroed once)
    {
      int* stw = (int*)(ws + TC_WS_STATUS);
      if (blockIdx.x == 0) {
        if (idx < TC_STATUS_MAGIC_WORD) stw[idx] = 0;
        if (idx == 0 && stw[TC_STATUS_MAGIC_WORD] != TC_STATUS_MAGIC) {
          for (int k = 0; k < 8; ++k) stw[TC_STATUS_STICKY + k] = 0;
          stw[TC_STATUS_MAGIC_WORD] = TC_STATUS_MAGIC;
        }
      }
    }
    if (blockIdx.x == 0) {
      // Can a hidden activation leave the fp16 operand range?  With |coords/scale| <= 4 and |sin| <= 1:
      //   B_0 = max_n sum_k f_k |W0[k][n]| + |b0[n]| (f_k = 4 for k < 3, else 1),
      //   B_l = max_n B_{l-1} sum_k |W_l[k][n]| + |b_l[n]|   (l = 1, 2)
      // bound h0, h1, h2 -- the operands the forward rounds to fp16.  If a bound exceeds 6e4 (or is NaN) the epilogue
      // tracks max|h| per sample and flags status word 3; otherwise that check costs nothing.
      __shared__ float red[128];
      float B = 1.f, need = 0.f;
      const int woffs[3] = {OFF_W0, OFF_W1, OFF_W2}, boffs[3] = {OFF_B0, OFF_B1, OFF_B2}, kr[3] = {21, 128, 128};
      for (int l = 0; l < 3; ++l) {
        if (threadIdx.x < 128) {
          float sum = 0.f;
          for (int k = 0; k < kr[l]; ++k) sum += fabsf(params[woffs[l] + k * 128 + threadIdx.x]) * ((l == 0 && k < 3) ? 4.f : 1.f);
          red[threadIdx.x] = sum * B + fabsf(params[boffs[l] + threadIdx.x]);
        }
        __syncthreads();
        for (int o = 64; o > 0; o >>= 1) {
          if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]) + 0.f * (red[threadIdx.x] + red[threadIdx.x + o]);
          __syncthreads();
        }
        B = red[0];
        if (!(B <= 6.0e4f)) need = 1.f;
        __syncthreads();
      }
      if (threadIdx.x == 0) c[TC_C_RANGE] = need;
    }
    return;
  }
  if (idx >= K * 128) return;
  int k = idx >> 7, n = idx & 127;
  const int woff[4] = {OFF_W0, OFF_W1, OFF_W2, OFF_W3};
  const int boff[4] = {OFF_B0, OFF_B1, OFF_B2, OFF_B3};
  const int krows[4] = {21, 128, 128, 149};
  float w = (k < krows[l]) ? params[woff[l] + k * 128 + n] : 0.f;
  // The bias rides on the MMA: feature columns TC_ONES_COL and TC_ONES_COL+1 are the constant 1, and the weight
  // rows they meet hold fp16(b) and b - fp16(b) (both exact in the hi plane).  Layers 0 and 3 contract over the
  // feature image anyway (rows 21,22 / 149,150); layers 1 and 2 get a 16-row bias image of their own (below).
  const float bias = params[boff[l] + n];
  const float bias_hi = __half2float(__float2half_rn(bias));
  if (l == 0 || l == 3) {
    if (k == krows[l]) w = bias_hi;
    if (k == krows[l] + 1) w = bias - bias_hi;
  } else if (k < 16) {       // bias image [16 rows = feature cols 16..31][128], hi plane only (lo plane stays zero)
    const int kf = 16 + k;
    const float wb = kf == TC_ONES_COL ? bias_hi : (kf == TC_ONES_COL + 1 ? bias - bias_hi : 0.f);
    uint8_t* bimg = ws + TC_WS_W + TC_W_BYTES + (uint32_t)(l - 1) * TC_BIMG_BYTES;
    *reinterpret_cast<__half*>(bimg + img_off(k, n, TC_IMG_RS, 2u * 128u)) = __float2half_rn(wb);
  }
  uint32_t off = img_off(k, n, TC_IMG_RS, (uint32_t)(K / 8) * 128u);
  {   // forward: fp16 hi/lo planes; layer 3 as two images (rows 0..127 | rows 128..159, see fw_off)
    __half hi = __float2half_rn(w);
    __half lo = __float2half_rn(w - __half2float(hi));
    const int j = (l == 3 && k >= 128) ? 4 : l;
    const int kk = (j == 4) ? k - 128 : k;
    const uint32_t foff = img_off(kk, n, TC_IMG_RS, (fw_K(j) / 8u) * 128u);
    uint8_t* base = ws + TC_WS_W + fw_off(j);
    *reinterpret_cast<__half*>(base + foff) = hi;
    *reinterpret_cast<__half*>(base + fw_plane(j) + foff) = lo;
  }
  {   // dgrad chain: bf16 hi/lo planes
    __nv_bfloat16 hi = __float2bfloat16_rn(w);
    __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    uint8_t* base = ws + TC_WS_WB + tc_stage_off(l);
    *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + tc_plane_bytes(l) + off) = lo;
  }
}

// ------------------------------------------------------------------------------------------------
// MMA issue for one (layer, slot): D[128x128] = A[128xK] * W_l[Kx128], NPASS fp16 products
//   pass 0: a_hi*w_hi   pass 1: a_lo*w_hi   pass 2: a_hi*w_lo
// Called by ONE elected lane; fully unrolled with the descriptor halves precomputed, so each tcgen05.mma costs a
// 32-bit add instead of a descriptor rebuild (measured: 74 instead of 135 cycles per MMA, scripts/run_umma_bench.py).
// ------------------------------------------------------------------------------------------------
template <int NPASS, int L>
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t tmem_a, uint32_t feat_smem, uint32_t w_smem,
                                            uint32_t skip_smem, uint32_t idesc) {
  constexpr uint32_t K = fw_K(L), plane = fw_plane(L), w_cs = (K / 8) * 128u;
  constexpr int nks = (int)(K / 16);
  // B: image [k][n] read MN-major: K groups advance by RS (LBO), N groups by CS (SBO)
  const uint32_t b_hi = desc_hi(w_cs);
  const uint32_t b_lo[2] = {desc_lo(w_smem, TC_IMG_RS), desc_lo(w_smem + plane, TC_IMG_RS)};
  // A (features): image [s][k] read K-major: K groups advance by CS (LBO), M groups by RS (SBO)
  const uint32_t a_hi = desc_hi(TC_IMG_RS);
  const uint32_t a_lo[2] = {desc_lo(feat_smem, TC_SIMG_CS), desc_lo(feat_smem + TC_FIMG_BYTES, TC_SIMG_CS)};
#pragma unroll
  for (int pass = 0; pass < NPASS; ++pass) {
    const int ap = (pass == 1) ? 1 : 0, wp = (pass == 2) ? 1 : 0;
#pragma unroll
    for (int ks = 0; ks < nks; ++ks) {
      const uint32_t bl = b_lo[wp] + (uint32_t)ks * ((2u * TC_IMG_RS) >> 4);
      const uint32_t acc = (pass | ks) ? 1u : 0u;
      if (L == 0) mma_ss_raw(tmem_d, a_lo[ap] + (uint32_t)ks * ((2u * TC_SIMG_CS) >> 4), a_hi, bl, b_hi, idesc, acc);
      else mma_ts_raw(tmem_d, tmem_a + (uint32_t)ap * 64u + (uint32_t)ks * 8u, bl, b_hi, idesc, acc);
    }
    if (L == 3) {     // skip connection: [features] x W3 rows 128..159 (the resident K = 32 image; its rows 21, 22 carry b3)
      constexpr uint32_t s_plane = fw_plane(4), s_cs = (fw_K(4) / 8) * 128u;
      const uint32_t sb_hi = desc_hi(s_cs);
      const uint32_t sb_lo = desc_lo(skip_smem + (uint32_t)wp * s_plane, TC_IMG_RS);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        mma_ss_raw(tmem_d, a_lo[ap] + (uint32_t)ks * ((2u * TC_SIMG_CS) >> 4), a_hi, sb_lo + (uint32_t)ks * ((2u * TC_IMG_RS) >> 4),
                   sb_hi, idesc, 1u);
    }
  }
  if (L == 1 || L == 2)      // + [feature cols 16..31] x bias image: adds b_hi + b_lo through the two constant-1 columns
    mma_ss_raw(tmem_d, a_lo[0] + ((2u * TC_SIMG_CS) >> 4), a_hi, desc_lo(w_smem + 2u * plane, TC_IMG_RS), desc_hi(2u * 128u),
               idesc, 1u);
}

// ---- branch-free sine of the reference's reduced argument r in [0, 2*100pi) (safe_sin, network.py:16) ----
// Cody-Waite reduction by pi/2 (3 constants, j*C1 exact for j < 2^16) + the two minimax polynomials on
// [-pi/4, pi/4]; abs error ~1e-7 like sinf(), but no slow-path branch, so the nine sines of a thread interleave.
__device__ __forceinline__ float sin_reduced(float r) {
  const float j = rintf(r * 0.636619772367581343f);
  const int q = (int)j;
  float x = fmaf(-j, 1.5703125f, r);
  x = fmaf(-j, 4.837512969970703125e-4f, x);
  x = fmaf(-j, 7.54978995489188216e-8f, x);
  const float x2 = x * x;
  const float sn = fmaf(fmaf(fmaf(-1.9515295891e-4f, x2, 8.3321608736e-3f), x2, -1.6666654611e-1f), x2 * x, x);
  const float cs = fmaf(fmaf(fmaf(2.443315711809948e-5f, x2, -1.388731625493765e-3f), x2, 4.166664568298827e-2f),
                        x2 * x2, fmaf(-0.5f, x2, 1.0f));
  const float val = (q & 1) ? cs : sn;
  return (q & 2) ? -val : val;
}
// Default: two-constant Cody-Waite reduction by 2*pi (j <= 100, j*6.28125 exact) + MUFU.SIN, abs error 2^-21.4 on
// [-pi, pi] -- 7 instructions instead of ~20 for the polynomial pair above (-DBH_EXP_POLYSIN selects that one); the
// epilogue warps are issue-bound, so this is 7 % of the forward (2.91 vs 3.13 ms, cfg2 x 25 frames) at unchanged
// image / gradient error against the oracle (1.9e-6 / 1.2e-4).
__device__ __forceinline__ float sin_reduced_mufu(float r) {
  const float j = rintf(r * 0.15915494309189535f);
  float x = fmaf(-j, 6.28125f, r);
  x = fmaf(-j, 1.9353071795864769e-3f, x);
  return __sinf(x);
}
__device__ __forceinline__ float safe_sin_fast(float a) {      // same float32 ARGUMENT arithmetic as bh_safe_sin
  float r = (fabsf(a) < BH_100PI_F) ? a : fmodf(a, BH_100PI_F);
  if (r < 0.0f) r = __fadd_rn(r, BH_100PI_F);
#ifdef BH_EXP_POLYSIN
  return sin_reduced(r);
#else
  return sin_reduced_mufu(r);
#endif
}

// warped + scaled coordinates of one sample (bh_features without the encodings)
__device__ __forceinline__ bool warp_coords(float x, float y, float z, float om, float tg, float tfc,
                                            const FrameConsts& fc, float* u) {
  float tM = __fsub_rn(__fadd_rn(tfc, tg), fc.t_injection);
  bool valid = !(tM < 0.0f);
  float sn, cs;
  sincosf(__fmul_rn(tM, om), &sn, &cs);
  u[0] = __fdiv_rn(x * cs + y * sn, fc.scale);
  u[1] = __fdiv_rn(y * cs - x * sn, fc.scale);
  u[2] = __fdiv_rn(z, fc.scale);
  if (!valid) { u[0] = 0.f; u[1] = 0.f; u[2] = 0.f; }
  return valid;
}

// hi = x rounded toward zero with the ReLU folded in (so lo = x - hi >= 0 wherever x > 0 and the second
// relu-conversion is exact for x <= 0): 8 instructions per pair instead of 10
__device__ __forceinline__ uint32_t pack_f16x2_rz_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2_rn_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2_rn_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// residual of the fp16 hi plane, x - float(hi), for both halves of a packed pair: ONE full-rate FMA-pipe instruction per
// element (fma.rn.f32.f16 = FHFMA: hi * -1 + x, exact -- the residual has <= 13 significant bits).  The earlier form, a
// mantissa mask (LOP3) + FADD per element, ran on the ALU pipe, which binds this epilogue: F2FP, LOP3, PRMT and HSET2 all
// issue there at one warp-instruction per 2 cycles (profiles/r2_sm_egress_ingress_probe.log).
__device__ __forceinline__ void f16x2_residual(uint32_t hi, float x0, float x1, float& r0, float& r1) {
  asm("{\n\t.reg .b16 l, u, m;\n\tmov.b32 {l, u}, %2;\n\tmov.b16 m, 0xBC00;\n\t"
      "fma.rn.f32.f16 %0, l, m, %3;\n\tfma.rn.f32.f16 %1, u, m, %4;\n\t}"
      : "=f"(r0), "=f"(r1) : "r"(hi), "f"(x0), "f"(x1));
}

// SAVE = bf16 planes of every activation kept for the backward (0, 1, 2).  RANGE = track max|h| per sample: compiled
// as a second body of the same kernel and entered only when the weight bound of tc_prepare_weights_kernel cannot
// rule out an fp16 operand overflow (the tracking costs ~10 % of the forward, so the usual path does not carry it).
template <int NPASS, int SAVE, bool RANGE>
__device__ __forceinline__ void
tc_fwd_body(uint8_t* smem, const PackedView& v, const FrameConsts& fc, const uint8_t* __restrict__ ws,
            const float* __restrict__ t_frames, int Bt, float* __restrict__ e_out, uint8_t* __restrict__ acts,
            int* __restrict__ status) {
  uint8_t* wst = smem + SM_WSTAGE;
  uint8_t* wskip = smem + SM_WSKIP;
  uint8_t* featimg = smem + SM_FEAT;                  // [slot][buffer][plane] x 8 KB
  float* cst = (float*)(smem + SM_CONST);
  uint64_t* bars = (uint64_t*)(smem + SM_BARS);
  uint32_t* tmem_base_s = (uint32_t*)(bars + BAR_COUNT);
  int* abort_s = (int*)(tmem_base_s + 1);
  float* part = (float*)(smem + SM_PART);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_frame = v.n_pad / 128;
  const int NT = Bt * tiles_per_frame;
  Abort ab{abort_s};

  if (tid == 0) {
    mbar_init(&bars[BAR_WFULL + 0], 1); mbar_init(&bars[BAR_WFULL + 1], 1);
    mbar_init(&bars[BAR_WEMPTY + 0], 1); mbar_init(&bars[BAR_WEMPTY + 1], 1);
    mbar_init(&bars[BAR_AREADY + 0], 8); mbar_init(&bars[BAR_AREADY + 1], 8);
    mbar_init(&bars[BAR_DREADY + 0], 1); mbar_init(&bars[BAR_DREADY + 1], 1);
    mbar_init(&bars[BAR_FREADY + 0], 8); mbar_init(&bars[BAR_FREADY + 1], 8);
    mbar_init(&bars[BAR_DFREE + 0], 8); mbar_init(&bars[BAR_DFREE + 1], 8);
    mbar_init(&bars[BAR_SKIP], 1);
    abort_s[0] = 0; abort_s[1] = 0;
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_base_s, 512);
  for (int i = tid; i < TC_CONST_FLOATS; i += kThreads) cst[i] = ((const float*)(ws + TC_WS_CONST))[i];
  for (int i = tid; i < (int)(2 * 2 * 2 * TC_FIMG_BYTES / 16); i += kThreads)   // feature columns 24..31 stay zero
    reinterpret_cast<uint4*>(featimg)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_base_s;

  if (warp > kMmaWarp + 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kAuxRegs));      // idle warps of the last warpgroup
  } else if (warp == kMmaWarp + 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kAuxRegs));
    // ===================== weight producer =====================
    if (lane == 0) {
      mbar_expect_tx(&bars[BAR_SKIP], 2u * fw_plane(4));                      // resident skip rows of layer 3, once
      bulk_g2s(wskip, ws + TC_WS_W + fw_off(4), 2u * fw_plane(4), &bars[BAR_SKIP]);
      uint32_t wcnt = 0;
      for (int r = 0;; ++r) {
        int T0 = (r * (int)gridDim.x + (int)blockIdx.x) * 2;
        if (T0 >= NT) break;
        bool ok = true;
        for (int l = 0; l < 4 && ok; ++l, ++wcnt) {
          uint32_t st = wcnt & 1u;
          ok = wait(&bars[BAR_WEMPTY + st], ((wcnt >> 1) & 1u) ^ 1u, ab);
          if (!ok) break;
          const bool bimg = (l == 1 || l == 2);
          const uint32_t bytes = 2u * fw_plane(l);
          mbar_expect_tx(&bars[BAR_WFULL + st], bytes + (bimg ? TC_BIMG_BYTES : 0u));
          bulk_g2s(wst + st * FW_STAGE, ws + TC_WS_W + fw_off(l), bytes, &bars[BAR_WFULL + st]);
          if (bimg)
            bulk_g2s(wst + st * FW_STAGE + bytes, ws + TC_WS_W + TC_W_BYTES + (uint32_t)(l - 1) * TC_BIMG_BYTES,
                     TC_BIMG_BYTES, &bars[BAR_WFULL + st]);
        }
        if (!ok) break;
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kAuxRegs));
    // ===================== MMA issuer: the whole warp runs the control flow, one elected lane issues =====================
    {
      const uint32_t idesc = make_idesc_f16(128, 128, 0, 1);
      uint32_t wcnt = 0, a_phase[2] = {0u, 0u}, f_phase[2] = {0u, 0u}, df_phase[2] = {0u, 0u};
      BH_TIMING_T0 BH_TIMING_DECL(t_ww) BH_TIMING_DECL(t_wa) BH_TIMING_DECL(t_is)
      bool ok = wait(&bars[BAR_SKIP], 0, ab);
      auto layer_step = [&](auto Ltag, int T0, int r) {
        constexpr int L = decltype(Ltag)::value;
        uint32_t st = wcnt & 1u;
        BH_TIMING_BEGIN
        ok = ok && wait(&bars[BAR_WFULL + st], (wcnt >> 1) & 1u, ab);
        BH_TIMING_END(t_ww)
        for (int s = 0; s < 2 && ok; ++s) {
          if (T0 + s >= NT) continue;
          BH_TIMING_BEGIN
          if (L == 0) {
            // layer 0 of a tile needs its feature image (written one tile ahead) and, from the slot's second tile on, the
            // accumulator: the previous tile's layer-3 result must be in the epilogue warps' registers
            ok = wait(&bars[BAR_FREADY + s], f_phase[s], ab);
            f_phase[s] ^= 1u;
            if (ok && r > 0) { ok = wait(&bars[BAR_DFREE + s], df_phase[s], ab); df_phase[s] ^= 1u; }
          } else {
            ok = wait(&bars[BAR_AREADY + s], a_phase[s], ab);
            a_phase[s] ^= 1u;
          }
          BH_TIMING_END(t_wa)
          if (!ok) break;
          tc_fence_after_sync();
          BH_TIMING_BEGIN
          if (elect_one()) {
            issue_layer<NPASS, L>(tbase + (uint32_t)s * 256u, tbase + (uint32_t)s * 256u + 128u,
                                  smem_u32(featimg + (s * 2 + (r & 1)) * 2 * TC_FIMG_BYTES), smem_u32(wst + st * FW_STAGE),
                                  smem_u32(wskip), idesc);
            mma_commit_raw(&bars[BAR_DREADY + s]);
          }
          __syncwarp();
          BH_TIMING_END(t_is)
        }
        if (ok) {
          if (elect_one()) mma_commit_raw(&bars[BAR_WEMPTY + st]);
          __syncwarp();
        }
        ++wcnt;
      };
      for (int r = 0; ok; ++r) {
        int T0 = (r * (int)gridDim.x + (int)blockIdx.x) * 2;
        if (T0 >= NT) break;
        layer_step(std::integral_constant<int, 0>{}, T0, r);
        layer_step(std::integral_constant<int, 1>{}, T0, r);
        layer_step(std::integral_constant<int, 2>{}, T0, r);
        layer_step(std::integral_constant<int, 3>{}, T0, r);
      }
      if (lane == 0) { BH_TIMING_STORE(status, 8, t_ww) BH_TIMING_STORE(status, 10, t_wa) BH_TIMING_STORE(status, 12, t_is) }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: (slot, column half, lane quadrant) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    const int slot = warp >> 3, half = (warp >> 2) & 1, q = warp & 3, row = q * 32 + lane;
    const uint32_t t_lane = tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * 256u;
    const uint32_t pair_bar = 1u + (uint32_t)(slot * 4 + q);          // named barrier of the two warps of a row
    uint32_t d_phase = 0;
    bool ok = true;
    auto tile_of = [&](int r) { return (r * (int)gridDim.x + (int)blockIdx.x) * 2 + slot; };
    // ---- features of the slot's tile of round r: warp + posenc in registers -> feature image buffer (r & 1).  Half 0 of a
    // row writes u and the sines (cols 0..11), half 1 the cosines and the two constant-1 columns (12..22).  Runs one tile
    // AHEAD of the MLP, while the current tile waits for the tensor core.  Returns valid / ray of the sample.
    float u[3] = {0.f, 0.f, 0.f};                       // warped, scaled coordinates between the two parts
    auto features_a = [&](int r, bool& valid, int& ray) {
      const int T = tile_of(r), b = T / tiles_per_frame, tile = T - b * tiles_per_frame, i = tile * 128 + row;
      ray = v.ray[i];
      valid = warp_coords(v.x[i], v.y[i], v.z[i], v.omega[i], v.tgeo[i], bh_frame_time(t_frames[b], fc), fc, u);
      // the weight bound of tc_prepare_weights_kernel assumes |coords/scale| <= 4 (the reference uses scale = rmax, i.e.
      // <= 1): outside it the fp16 operand range cannot be vouched for without tracking
      if (!RANGE && fmaxf(fmaxf(fabsf(u[0]), fabsf(u[1])), fabsf(u[2])) > 4.f) abort_s[1] = 1;
    };
    auto features_b = [&](int r) {
      const int T = tile_of(r), b = T / tiles_per_frame, tile = T - b * tiles_per_frame;
      uint8_t* my_feat = featimg + (slot * 2 + (r & 1)) * 2 * TC_FIMG_BYTES;
      float f[16];
      if (half == 0) {
        f[0] = u[0]; f[1] = u[1]; f[2] = u[2];
#pragma unroll
        for (int ii = 0; ii < 3; ++ii)
#pragma unroll
          for (int c = 0; c < 3; ++c) f[3 + 3 * ii + c] = safe_sin_fast(u[c] * (float)(1 << ii));
        f[12] = f[13] = f[14] = f[15] = 0.f;
      } else {
#pragma unroll
        for (int ii = 0; ii < 3; ++ii)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            f[4 + 3 * ii + c] = safe_sin_fast(__fadd_rn(u[c] * (float)(1 << ii), BH_HALFPI_F));     // cols 12..20
        f[0] = f[1] = f[2] = f[3] = 0.f;
        f[13] = f[14] = 1.f;             // cols 21, 22 = 1: the weight rows they meet carry the bias (hi, lo parts)
        f[15] = 0.f;
      }
      // half 0: chunk 0 (cols 0..7) + first 8 bytes of chunk 1 (cols 8..11)
      // half 1: last 8 bytes of chunk 1 (cols 12..15) + chunk 2 (cols 16..23) [+ zero chunk 3 of the saved copy]
      uint4 hA, lA, hB, lB;
      split8_f16(f, hA, lA);
      split8_f16(f + 8, hB, lB);
      const uint32_t o0 = sample_img_off(row, 0), o1 = sample_img_off(row, 1), o2 = sample_img_off(row, 2);
      if (half == 0) {
        *reinterpret_cast<uint4*>(my_feat + o0) = hA;
        *reinterpret_cast<uint2*>(my_feat + o1) = make_uint2(hB.x, hB.y);
        if (NPASS > 1) {
          *reinterpret_cast<uint4*>(my_feat + TC_FIMG_BYTES + o0) = lA;
          *reinterpret_cast<uint2*>(my_feat + TC_FIMG_BYTES + o1) = make_uint2(lB.x, lB.y);
        }
      } else {
        *reinterpret_cast<uint2*>(my_feat + o1 + 8) = make_uint2(hA.z, hA.w);
        *reinterpret_cast<uint4*>(my_feat + o2) = hB;
        if (NPASS > 1) {
          *reinterpret_cast<uint2*>(my_feat + TC_FIMG_BYTES + o1 + 8) = make_uint2(lA.z, lA.w);
          *reinterpret_cast<uint4*>(my_feat + TC_FIMG_BYTES + o2) = lB;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_FREADY + slot]);
      if (SAVE) {
        // the backward's copy keeps ONE constant-1 column (21): wgrad reads d bias off it.  One-plane plan: the fp16 hi
        // plane just computed (column 22 cleared); two-plane plan: bf16 hi + lo planes.
        uint8_t* feat_save = acts + (size_t)b * tc_acts_bytes_per_frame(v.n_pad, SAVE) + (size_t)v.n_pad * 1024u * SAVE +
                             (size_t)tile * TC_FIMG_BYTES;
        if (SAVE == 2) {
          if (half == 1) f[14] = 0.f;
          split8(f, hA, lA);
          split8(f + 8, hB, lB);
        } else if (half == 1) {
          hB.w = 0u;                                    // cols 22, 23
        }
        const size_t lo_off = (size_t)v.n_pad * 64u;
        if (half == 0) {
          *reinterpret_cast<uint4*>(feat_save + o0) = hA;
          *reinterpret_cast<uint2*>(feat_save + o1) = make_uint2(hB.x, hB.y);
          if (SAVE == 2) {
            *reinterpret_cast<uint4*>(feat_save + lo_off + o0) = lA;
            *reinterpret_cast<uint2*>(feat_save + lo_off + o1) = make_uint2(lB.x, lB.y);
          }
        } else {
          // (columns 24..31 of the saved copy are never written: the wgrad discards the accumulator columns they feed)
          *reinterpret_cast<uint2*>(feat_save + o1 + 8) = make_uint2(hA.z, hA.w);
          *reinterpret_cast<uint4*>(feat_save + o2) = hB;
          if (SAVE == 2) {
            *reinterpret_cast<uint2*>(feat_save + lo_off + o1 + 8) = make_uint2(lA.z, lA.w);
            *reinterpret_cast<uint4*>(feat_save + lo_off + o2) = lB;
          }
        }
      }
    };
    BH_TIMING_T0 BH_TIMING_DECL(t_ft) BH_TIMING_DECL(t_wd) BH_TIMING_DECL(t_ep)
    bool n_valid = false; int n_ray = -1;               // of the tile whose features were computed last
    if (tile_of(0) < NT) {                              // first tile of this slot: nothing to hide under yet
      BH_TIMING_BEGIN
      features_a(0, n_valid, n_ray);
      features_b(0);
      BH_TIMING_END(t_ft)
    }
    for (int r = 0; ok; ++r) {
      int T0 = (r * (int)gridDim.x + (int)blockIdx.x) * 2;
      if (T0 >= NT) break;
      int T = T0 + slot;
      if (T >= NT) continue;
      const int b = T / tiles_per_frame, tile = T - b * tiles_per_frame;
      const int i = tile * 128 + row;
      const bool valid = n_valid; const int ray = n_ray;              // of THIS tile (computed one tile ago)
      const bool has_next = tile_of(r + 1) < NT;
      uint8_t* act_tile = SAVE ? acts + (size_t)b * tc_acts_bytes_per_frame(v.n_pad, SAVE) + (size_t)tile * TC_SIMG_BYTES : nullptr;
      float o = 0.f, xmax = 0.f;                        // xmax: largest activation written as an fp16 operand
      for (int l = 0; l < 4; ++l) {
        BH_TIMING_BEGIN
        ok = wait(&bars[BAR_DREADY + slot], d_phase, ab);
        BH_TIMING_END(t_wd)
        if (!ok) break;
        d_phase ^= 1u;
        tc_fence_after_sync();
        BH_TIMING_BEGIN
        uint8_t* act_img = SAVE ? act_tile + (size_t)l * ((size_t)v.n_pad * 256u) : nullptr;
        // this warp's 64 columns: both TMEM loads in flight before the first use
        uint32_t raw[2][32];
        uint32_t mword[2] = {0u, 0u};                   // relu bit masks of this thread's 64 columns (backward)
        tmem_ld32(t_lane + (uint32_t)(half * 64), raw[0]);
        tmem_ld32(t_lane + (uint32_t)(half * 64 + 32), raw[1]);
        tmem_wait_ld();
        if (l == 3) {                                   // the accumulator is in registers: free for the next tile's layer 0
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BAR_DFREE + slot]);
        }
        // 16 columns at a time (operand planes -> TMEM, saved copy -> global): short live ranges, the 64 accumulator values
        // are the only large register block of the epilogue
#pragma unroll
        for (int sc = 0; sc < 4; ++sc) {
          const int cc = sc >> 1, c0 = half * 64 + sc * 16, j0 = (sc & 1) * 8;      // j0: first pair of the mask word
          const uint32_t* rw = &raw[cc][(sc & 1) * 16];
          float x[16];                                  // pre-activation (the bias came through the MMA)
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(rw[j]);
          uint32_t hi[8], lo[8];
          if (l < 3) {                                  // next layer's A operand: fp16 hi/lo planes in TMEM
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              hi[j] = pack_f16x2_rz_relu(x[2 * j], x[2 * j + 1]);
              // lo = x - hi (hi = x rounded toward zero, ReLU folded in: lo >= 0 wherever x > 0, and for x <= 0 the
              // second relu-conversion gives 0)
              if (NPASS > 1) {
#ifdef BH_EXP_MASKLO     // round-1 form: mantissa mask + FADD (ALU pipe)
                lo[j] = pack_f16x2_rn_relu(x[2 * j] - __uint_as_float(__float_as_uint(x[2 * j]) & 0xFFFFE000u),
                                           x[2 * j + 1] - __uint_as_float(__float_as_uint(x[2 * j + 1]) & 0xFFFFE000u));
#else
                float r0, r1;
                f16x2_residual(hi[j], x[2 * j], x[2 * j + 1], r0, r1);
                lo[j] = pack_f16x2_rn_relu(r0, r1);
#endif
              }
            }
            if (RANGE) {
#pragma unroll
              for (int j = 0; j < 8; ++j) xmax = fmaxf(fmaxf(x[2 * j], x[2 * j + 1]), xmax);
            }
            tmem_st8(t_lane + 128u + (uint32_t)(c0 >> 1), hi);
            if (NPASS > 1) tmem_st8(t_lane + 192u + (uint32_t)(c0 >> 1), lo);
          } else {
            const float* w4 = cst + TC_C_W4 + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) o = fmaf(fmaxf(x[j], 0.f), w4[j], o);
          }
          if (SAVE) {
            // saved for the backward.  One-plane plan: the fp16 hi plane itself (x rounded toward zero, ReLU folded into
            // the cvt -- no second conversion); two-plane plan: bf16 hi + lo planes.
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (SAVE == 1) {
                if (l == 3) hi[j] = pack_f16x2_rz_relu(x[2 * j], x[2 * j + 1]);
#ifndef BH_EXP_NOMASK      // timing experiment: no relu bit masks (wrong gradients)
                mword[cc] |= tc_mask_bits_f16(hi[j], j0 + j);
#endif
              } else {
                hi[j] = pack_bf16x2_rn_relu(x[2 * j], x[2 * j + 1]);
                mword[cc] |= tc_mask_bits(hi[j], j0 + j);
                lo[j] = pack_bf16x2(fmaxf(x[2 * j], 0.f) - bf16_lo(hi[j]), fmaxf(x[2 * j + 1], 0.f) - bf16_hi(hi[j]));
              }
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
#ifdef BH_EXP_NOSAVESTG    // timing experiment: the activation images are not written (wrong gradients)
              if (hi[4 * g] == 0x12345678u)
#endif
#if defined(BH_EXP_STCS)
              __stcs(reinterpret_cast<uint4*>(act_img + sample_img_off(row, (c0 >> 3) + g)),
                     make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]));
#elif defined(BH_EXP_STWT)
              __stwt(reinterpret_cast<uint4*>(act_img + sample_img_off(row, (c0 >> 3) + g)),
                     make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]));
#else
              *reinterpret_cast<uint4*>(act_img + sample_img_off(row, (c0 >> 3) + g)) =
                  make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
#endif
              if (SAVE == 2)
                *reinterpret_cast<uint4*>(act_img + (size_t)v.n_pad * 1024u + sample_img_off(row, (c0 >> 3) + g)) =
                    make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
            }
          }
        }
        if (l < 3) {
          tmem_wait_st();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BAR_AREADY + slot]);
        }
        if (SAVE)
          *reinterpret_cast<uint2*>(acts + (size_t)b * tc_acts_bytes_per_frame(v.n_pad, SAVE) + tc_mask_off(v.n_pad, SAVE) +
                                    tc_mask_word_off(tile, l, half, row)) = make_uint2(mword[0], mword[1]);
        BH_TIMING_END(t_ep)
        if (l < 2 && has_next) {                        // the next tile's features, under the MMAs of layers 1 and 2
          BH_TIMING_BEGIN
          if (l == 0) features_a(r + 1, n_valid, n_ray);
          else features_b(r + 1);
          BH_TIMING_END(t_ft)
        }
      }
      if (!ok) break;
      if (RANGE && xmax > 65504.f) abort_s[1] = 1;      // fp16 operand range exceeded (hi saturates): flag, do not hide
      // last layer: the two halves of a row combine their partial dot products through shared memory
      if (half == 1) {
        part[slot * 128 + row] = o;
        asm volatile("bar.arrive %0, 64;" ::"r"(pair_bar) : "memory");
      } else {
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        o += part[slot * 128 + row] + cst[TC_C_B4];
        float e = bh_sigmoid_m10(o);
        if (!(fabsf(o) <= 3.0e38f)) abort_s[1] = 1;    // fp16 operand overflow (|activation| > 65504): flag, do not hide
        e_out[(size_t)b * v.n_pad + i] = (valid && ray >= 0) ? e : 0.f;       // network.py:232
      }
    }
    if (tid == 0) { BH_TIMING_STORE(status, 14, t_ft) BH_TIMING_STORE(status, 16, t_wd) BH_TIMING_STORE(status, 18, t_ep) }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tbase, 512);
  if (tid == 0 && *abort_s) raise_flag(status, 0);
  if (tid == 0 && abort_s[1]) raise_flag(status, 3);
}

template <int NPASS, int SAVE>
__global__ void __launch_bounds__(kThreads, 1)
tc_fwd_kernel(PackedView v, FrameConsts fc, const uint8_t* __restrict__ ws, const float* __restrict__ t_frames,
              int Bt, float* __restrict__ e_out, uint8_t* __restrict__ acts, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
#ifdef BH_TC_TIMING
  const long long t_kernel0 = clock64();
#endif
#ifndef BH_EXP_NORANGE
  if (((const float*)(ws + TC_WS_CONST))[TC_C_RANGE] != 0.f)
    tc_fwd_body<NPASS, SAVE, true>(smem, v, fc, ws, t_frames, Bt, e_out, acts, status);
  else
#endif
    tc_fwd_body<NPASS, SAVE, false>(smem, v, fc, ws, t_frames, Bt, e_out, acts, status);
#ifdef BH_TC_TIMING
  if (threadIdx.x == 0) {     // per-CTA record for scripts/tc_timing.py: cycles (in units of 1024) and the SM it ran on
    uint32_t smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
    ((uint32_t*)((uint8_t*)status + TC_WS_CONST))[710 + blockIdx.x] = ((uint32_t)((clock64() - t_kernel0) >> 10) << 10) | smid;
  }
#endif
}


int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int NPASS, int SAVE>
int launch_fwd(const PackedView& v, const FrameConsts& fc, const void* ws, const float* t_frames, int Bt,
               float* e_out, void* acts, cudaStream_t st) {
  auto kern = tc_fwd_kernel<NPASS, SAVE>;
  BH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL2));
  int NT = Bt * (v.n_pad / 128);
  int grid = (NT + 1) / 2; if (grid > num_sms()) grid = num_sms();
  kern<<<grid, kThreads, SM_TOTAL2, st>>>(v, fc, (const uint8_t*)ws, t_frames, Bt, e_out, (uint8_t*)acts,
                                        (int*)((uint8_t*)ws + TC_WS_STATUS));
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

size_t bh_tc_acts_bytes_per_frame(int n_pad, int planes) { return tc_acts_bytes_per_frame(n_pad, planes); }

// Precision plan of the backward.  One-plane plan: the saved activations are the forward's fp16 hi plane and
// the cotangents one fp16 plane (launch-wide power-of-two scale); the rounding noise of the wgrad reduction averages out as
// 1/sqrt(#sample-frames).  Measured against the float64 oracle at 1.1e4 sample-frames (profiles/r2_planes_fp16.log): 4.9e-4
// of the gradient for the Q/U lightcurve loss, 2.9e-4 for closure phases (both are small differences of large per-ray terms),
// 3e-5 .. 5e-5 for the other losses; at 4.5e6 sample-frames 9.6e-5 (that floor is the forward's own 3e-6 image error).  The
// plan is used from 2^16 sample-frames per step on (worst case 2e-4 there, tolerance 1e-3); below that both bf16 planes are
// kept (x3 products, two kernels, ~1e-5 at any size).  The per-pixel 'full' loss has no cancellation between rays (2.8e-5 at
// 1.1e4) and switches at 2^15.  BHNERF_TC_PLANES=1|2 overrides.
int bh_tc_planes_full_loss(int n_active, int Bt_total) {
  int base = bh_tc_planes(n_active, Bt_total);
  if (base == 1) return 1;
  const char* e = getenv("BHNERF_TC_PLANES");
  if (e && (e[0] == '1' || e[0] == '2')) return base;
  return ((long long)n_active * (long long)Bt_total < (1ll << 15)) ? 2 : 1;
}
int bh_tc_planes(int n_active, int Bt_total) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("BHNERF_TC_PLANES");
    forced = (e && (e[0] == '1' || e[0] == '2')) ? (e[0] - '0') : 0;
  }
  if (forced) return forced;
  return ((long long)n_active * (long long)Bt_total < (1ll << 16)) ? 2 : 1;
}
size_t bh_tc_ws_bytes() { return 1u << 20; }

int bh_tc_prepare_weights(const float* params, void* ws, cudaStream_t st) {
  BhProfScope ps(BH_CAT_MISC, 1, st);
  tc_prepare_weights_kernel<<<dim3(160 * 128 / 256, 5), 256, 0, st>>>(params, (uint8_t*)ws);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// BHNERF_PRECISION=fast: the STATED fast mode (north_star "tf32/bf16 MLP stated"; SURVEY.md s0.9: report its error
// separately).  One fp16 product per layer in the forward (a_hi * w_hi) and in the dgrad chain -- a third of the forward's
// tensor work.  It does NOT meet the 1e-4 / 1e-3 parity tolerances (measured errors: profiles/r2_fast_mode.log) and is never
// the default, the headline number or what the parity tests run.
bool bh_tc_fast() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("BHNERF_PRECISION"); on = (e && e[0] == 'f') ? 1 : 0; }
  return on == 1;
}

int bh_tc_fwd(const PackedView& v, const FrameConsts& fc, const void* ws, const float* params,
              const float* t_frames, int Bt, float* e_out, void* acts, int planes, cudaStream_t st) {
  (void)params;
  BhProfScope ps(BH_CAT_FWD, 1, st);
  if (bh_tc_fast() && planes != 2) {
    if (!acts) return launch_fwd<1, 0>(v, fc, ws, t_frames, Bt, e_out, nullptr, st);
    return launch_fwd<1, 1>(v, fc, ws, t_frames, Bt, e_out, acts, st);
  }
  if (!acts) return launch_fwd<3, 0>(v, fc, ws, t_frames, Bt, e_out, nullptr, st);
  return planes == 2 ? launch_fwd<3, 2>(v, fc, ws, t_frames, Bt, e_out, acts, st)
                     : launch_fwd<3, 1>(v, fc, ws, t_frames, Bt, e_out, acts, st);
}


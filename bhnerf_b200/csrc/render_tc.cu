// tcgen05 forward render kernel (sm_100a): warp -> posenc -> 4x128 MLP on the 5th-gen tensor cores
// (bf16x3 split operands, fp32 accumulate in TMEM) -> sigmoid(o-10) -> masked emission.
// Reference semantics: bhnerf/network.py:18-64 (MLP), :98-122 (posenc), :191-237 (NeRF_Predictor.__call__),
// bhnerf/emission.py:143-211 (velocity_warp_coords).  DESIGN.md s4 describes the pipeline.
//
// One persistent CTA per SM, 10 warps:
//   warps 0-3  epilogue of tile slot 0 (TMEM lane quadrant = warp%4; thread = one sample row)
//   warps 4-7  epilogue of tile slot 1
//   warp  8    MMA issuer (one elected lane issues every tcgen05.mma / tcgen05.commit)
//   warp  9    weight producer (cp.async.bulk of the next layer's [hi|lo] weight images into a 2-stage ring)
// Two 128-sample tiles are in flight per CTA and ping-pong: while the tensor core runs layer l of one tile,
// the CUDA cores run the bias+ReLU+hi/lo-split epilogue of the other.  Activations never leave the SM:
// D (fp32) is read from TMEM with tcgen05.ld and the next layer's A operand is written back to TMEM with
// tcgen05.st (TS-form MMA); only the 21 posenc features go through shared memory (SS-form, K-major).
#include <cuda_fp16.h>
#include <stdlib.h>
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int kThreads = 320;
constexpr uint32_t SM_WSTAGE = 0;                                  // 2 x 80 KB weight ring
constexpr uint32_t SM_FEAT = 2 * TC_STAGE_MAX;                     // 2 slots x [hi 8K | lo 8K]
constexpr uint32_t SM_CONST = SM_FEAT + 2 * 2 * TC_FIMG_BYTES;     // 768 floats
constexpr uint32_t SM_BARS = SM_CONST + TC_CONST_FLOATS * 4;       // 8 mbarriers + tmem base + abort flag
constexpr uint32_t SM_TOTAL = SM_BARS + 128;

enum { BAR_WFULL = 0, BAR_WEMPTY = 2, BAR_AREADY = 4, BAR_DREADY = 6 };

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
    if (idx < 64) ((int*)(ws + TC_WS_STATUS))[idx] = 0;
    return;
  }
  if (idx >= K * 128) return;
  int k = idx >> 7, n = idx & 127;
  const int woff[4] = {OFF_W0, OFF_W1, OFF_W2, OFF_W3};
  const int krows[4] = {21, 128, 128, 149};
  float w = (k < krows[l]) ? params[woff[l] + k * 128 + n] : 0.f;
  uint32_t off = img_off(k, n, TC_IMG_RS, (uint32_t)(K / 8) * 128u);
  {   // forward: fp16 hi/lo planes
    __half hi = __float2half_rn(w);
    __half lo = __float2half_rn(w - __half2float(hi));
    uint8_t* base = ws + TC_WS_W + tc_stage_off(l);
    *reinterpret_cast<__half*>(base + off) = hi;
    *reinterpret_cast<__half*>(base + tc_plane_bytes(l) + off) = lo;
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
// MMA issue for one (layer, slot): D[128x128] = A[128xK] * W_l[Kx128], NPASS bf16 products
//   pass 0: a_hi*w_hi   pass 1: a_lo*w_hi   pass 2: a_hi*w_lo
// ------------------------------------------------------------------------------------------------
template <int NPASS>
__device__ __forceinline__ void issue_layer(int l, uint32_t tmem_d, uint32_t tmem_a, uint32_t feat_smem,
                                            uint32_t w_smem, uint32_t idesc) {
  const uint32_t K = tc_layer_K(l), plane = tc_plane_bytes(l), w_cs = (K / 8) * 128u;
  const int nks = (int)(K / 16);
  uint32_t acc = 0;
#pragma unroll 1
  for (int pass = 0; pass < NPASS; ++pass) {
    const uint32_t a_lo = (pass == 1) ? 1u : 0u, w_lo = (pass == 2) ? 1u : 0u;
#pragma unroll 1
    for (int ks = 0; ks < nks; ++ks) {
      // B: image [k][n] read MN-major: K groups advance by RS (LBO), N groups by CS (SBO)
      uint64_t bd = make_desc(w_smem + w_lo * plane + (uint32_t)ks * 2u * TC_IMG_RS, TC_IMG_RS, w_cs);
      const bool a_from_smem = (l == 0) || (l == 3 && ks >= 8);
      if (a_from_smem) {
        // A: feature image [s][k] read K-major: K groups advance by CS (LBO), M groups by RS (SBO)
        uint32_t kk = (l == 0) ? (uint32_t)ks : (uint32_t)(ks - 8);
        uint64_t ad = make_desc(feat_smem + a_lo * TC_FIMG_BYTES + kk * 2u * TC_SIMG_CS, TC_SIMG_CS, TC_IMG_RS);
        mma_ss(tmem_d, ad, bd, idesc, acc);
      } else {
        mma_ts(tmem_d, tmem_a + a_lo * 64u + (uint32_t)ks * 8u, bd, idesc, acc);
      }
      acc = 1;
    }
  }
}

template <int NPASS, int SAVE>      // SAVE = bf16 planes of every activation kept for the backward (0, 1, 2)
__global__ void __launch_bounds__(kThreads, 1)
tc_fwd_kernel(PackedView v, FrameConsts fc, const uint8_t* __restrict__ ws, const float* __restrict__ t_frames,
              int Bt, float* __restrict__ e_out, uint8_t* __restrict__ acts, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* wst = smem + SM_WSTAGE;
  uint8_t* featimg = smem + SM_FEAT;
  float* cst = (float*)(smem + SM_CONST);
  uint64_t* bars = (uint64_t*)(smem + SM_BARS);
  uint32_t* tmem_base_s = (uint32_t*)(bars + 8);
  int* abort_s = (int*)(tmem_base_s + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_frame = v.n_pad / 128;
  const int NT = Bt * tiles_per_frame;
  Abort ab{abort_s};

  if (tid == 0) {
    mbar_init(&bars[BAR_WFULL + 0], 1); mbar_init(&bars[BAR_WFULL + 1], 1);
    mbar_init(&bars[BAR_WEMPTY + 0], 1); mbar_init(&bars[BAR_WEMPTY + 1], 1);
    mbar_init(&bars[BAR_AREADY + 0], 4); mbar_init(&bars[BAR_AREADY + 1], 4);
    mbar_init(&bars[BAR_DREADY + 0], 1); mbar_init(&bars[BAR_DREADY + 1], 1);
    abort_s[0] = 0; abort_s[1] = 0;
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(tmem_base_s, 512);
  for (int i = tid; i < TC_CONST_FLOATS; i += kThreads) cst[i] = ((const float*)(ws + TC_WS_CONST))[i];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_base_s;

  if (warp == 9) {
    // ===================== weight producer =====================
    if (lane == 0) {
      uint32_t wcnt = 0;
      for (int r = 0;; ++r) {
        int T0 = (r * (int)gridDim.x + (int)blockIdx.x) * 2;
        if (T0 >= NT) break;
        bool ok = true;
        for (int l = 0; l < 4 && ok; ++l, ++wcnt) {
          uint32_t st = wcnt & 1u;
          ok = wait(&bars[BAR_WEMPTY + st], ((wcnt >> 1) & 1u) ^ 1u, ab);
          if (!ok) break;
          mbar_expect_tx(&bars[BAR_WFULL + st], tc_stage_bytes(l));
          bulk_g2s(wst + st * TC_STAGE_MAX, ws + TC_WS_W + tc_stage_off(l), tc_stage_bytes(l), &bars[BAR_WFULL + st]);
        }
        if (!ok) break;
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 128, 0, 1);
      uint32_t wcnt = 0, a_phase[2] = {0u, 0u};
      for (int r = 0;; ++r) {
        int T0 = (r * (int)gridDim.x + (int)blockIdx.x) * 2;
        if (T0 >= NT) break;
        bool ok = true;
        for (int l = 0; l < 4 && ok; ++l, ++wcnt) {
          uint32_t st = wcnt & 1u;
          ok = wait(&bars[BAR_WFULL + st], (wcnt >> 1) & 1u, ab);
          if (!ok) break;
          for (int s = 0; s < 2; ++s) {
            if (T0 + s >= NT) continue;
            ok = wait(&bars[BAR_AREADY + s], a_phase[s], ab);
            if (!ok) break;
            a_phase[s] ^= 1u;
            tc_fence_after_sync();
            issue_layer<NPASS>(l, tbase + (uint32_t)s * 256u, tbase + (uint32_t)s * 256u + 128u,
                               smem_u32(featimg + s * 2 * TC_FIMG_BYTES), smem_u32(wst + st * TC_STAGE_MAX), idesc);
            mma_commit(&bars[BAR_DREADY + s]);
          }
          if (!ok) break;
          mma_commit(&bars[BAR_WEMPTY + st]);
        }
        if (!ok) break;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    const int slot = warp >> 2, q = warp & 3, row = q * 32 + lane;
    const uint32_t t_lane = tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * 256u;
    uint8_t* my_feat = featimg + slot * 2 * TC_FIMG_BYTES;
    uint32_t d_phase = 0;
    bool ok = true;
    for (int r = 0; ok; ++r) {
      int T0 = (r * (int)gridDim.x + (int)blockIdx.x) * 2;
      if (T0 >= NT) break;
      int T = T0 + slot;
      if (T >= NT) continue;
      const int b = T / tiles_per_frame, tile = T - b * tiles_per_frame;
      const int i = tile * 128 + row;
      // ---- warp + posenc in registers, split into bf16 hi/lo feature images ----
      float f[32];
      const float tfc = bh_frame_time(t_frames[b], fc);
      const bool valid = bh_features(v.x[i], v.y[i], v.z[i], v.omega[i], v.tgeo[i], tfc, fc, f);
#pragma unroll
      for (int k = BH_NF; k < 32; ++k) f[k] = 0.f;
      // column 21 = 1: the weight images have zero rows there, and the wgrad kernel reads bias gradients off it
      if (SAVE) f[TC_ONES_COL] = 1.f;
      uint8_t* feat_save = SAVE ? acts + (size_t)b * tc_acts_bytes_per_frame(v.n_pad, SAVE) +
                                      (size_t)v.n_pad * 1024u * SAVE + (size_t)tile * TC_FIMG_BYTES : nullptr;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 hi, lo;
        split8_f16(f + 8 * g, hi, lo);
        uint32_t off = sample_img_off(row, g);
        *reinterpret_cast<uint4*>(my_feat + off) = hi;
        if (NPASS > 1) *reinterpret_cast<uint4*>(my_feat + TC_FIMG_BYTES + off) = lo;
        if (SAVE) {                                     // the backward's operands are bf16 (range of the cotangents)
          split8(f + 8 * g, hi, lo);
          *reinterpret_cast<uint4*>(feat_save + off) = hi;
          if (SAVE == 2) *reinterpret_cast<uint4*>(feat_save + (size_t)v.n_pad * 64u + off) = lo;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_AREADY + slot]);
      uint8_t* act_tile = SAVE ? acts + (size_t)b * tc_acts_bytes_per_frame(v.n_pad, SAVE) + (size_t)tile * TC_SIMG_BYTES : nullptr;
      float o = cst[TC_C_B4];
      for (int l = 0; l < 4; ++l) {
        ok = wait(&bars[BAR_DREADY + slot], d_phase, ab);
        if (!ok) break;
        d_phase ^= 1u;
        tc_fence_after_sync();
        const float* bias = cst + TC_C_B(l);
        uint8_t* act_img = SAVE ? act_tile + (size_t)l * ((size_t)v.n_pad * 256u) : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(t_lane + (uint32_t)c0, raw);
          tmem_wait_ld();
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(__uint_as_float(raw[j]) + bias[c0 + j], 0.f);
          uint32_t hi[16], lo[16];
          if (l < 3) {                                  // next layer's A operand: fp16 hi/lo planes in TMEM
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              hi[j] = pack_f16x2(x[2 * j], x[2 * j + 1]);
              if (NPASS > 1) lo[j] = pack_f16x2(x[2 * j] - f16_lo(hi[j]), x[2 * j + 1] - f16_hi(hi[j]));
            }
            tmem_st16(t_lane + 128u + (uint32_t)(c0 >> 1), hi);
            if (NPASS > 1) tmem_st16(t_lane + 192u + (uint32_t)(c0 >> 1), lo);
          } else {
            const float* w4 = cst + TC_C_W4 + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) o = fmaf(x[j], w4[j], o);
          }
          if (SAVE) {                                   // saved for the backward as bf16 planes
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              hi[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
              if (SAVE == 2) lo[j] = pack_bf16x2(x[2 * j] - bf16_lo(hi[j]), x[2 * j + 1] - bf16_hi(hi[j]));
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              *reinterpret_cast<uint4*>(act_img + sample_img_off(row, (c0 >> 3) + g)) =
                  make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
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
      }
      if (!ok) break;
      float e = bh_sigmoid_m10(o);
      if (!(fabsf(o) <= 3.0e38f)) abort_s[1] = 1;      // fp16 operand overflow (|activation| > 65504): flag, do not hide
      e_out[(size_t)b * v.n_pad + i] = (valid && v.ray[i] >= 0) ? e : 0.f;     // network.py:232
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tbase, 512);
  if (tid == 0 && *abort_s) atomicExch(status, 1);
  if (tid == 0 && abort_s[1]) atomicExch(status + 3, 1);
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
  BH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
  int NT = Bt * (v.n_pad / 128);
  int grid = (NT + 1) / 2; if (grid > num_sms()) grid = num_sms();
  kern<<<grid, kThreads, SM_TOTAL, st>>>(v, fc, (const uint8_t*)ws, t_frames, Bt, e_out, (uint8_t*)acts,
                                        (int*)((uint8_t*)ws + TC_WS_STATUS));
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

size_t bh_tc_acts_bytes_per_frame(int n_pad, int planes) { return tc_acts_bytes_per_frame(n_pad, planes); }

// Precision plan of the backward: the saved activations / cotangents are bf16.  With the hi plane
// only, the rounding noise of the wgrad reduction averages out as 1/sqrt(#sample-frames) (measured 1e-4 of the
// gradient at 4.5e6 sample-frames); below 2^19 sample-frames per step both planes are kept (x3 products, error
// ~1e-5 at any size).  BHNERF_TC_PLANES=1|2 overrides.
int bh_tc_planes(int n_active, int Bt_total) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("BHNERF_TC_PLANES");
    forced = (e && (e[0] == '1' || e[0] == '2')) ? (e[0] - '0') : 0;
  }
  if (forced) return forced;
  return ((long long)n_active * (long long)Bt_total < (1ll << 19)) ? 2 : 1;
}
size_t bh_tc_ws_bytes() { return 1u << 20; }

int bh_tc_prepare_weights(const float* params, void* ws, cudaStream_t st) {
  BhProfScope ps(BH_CAT_MISC, 1, st);
  tc_prepare_weights_kernel<<<dim3(160 * 128 / 256, 5), 256, 0, st>>>(params, (uint8_t*)ws);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int bh_tc_fwd(const PackedView& v, const FrameConsts& fc, const void* ws, const float* params,
              const float* t_frames, int Bt, float* e_out, void* acts, int planes, cudaStream_t st) {
  (void)params;
  BhProfScope ps(BH_CAT_FWD, 1, st);
  if (!acts) return launch_fwd<3, 0>(v, fc, ws, t_frames, Bt, e_out, nullptr, st);
  return planes == 2 ? launch_fwd<3, 2>(v, fc, ws, t_frames, Bt, e_out, acts, st)
                     : launch_fwd<3, 1>(v, fc, ws, t_frames, Bt, e_out, acts, st);
}


// tcgen05 backward of the render (sm_100a): parameter gradient of image_plane_prediction
// (jax.value_and_grad pull-back, bhnerf/network.py:617,:677) from the activations the forward kernel saved.
//
//   tc_dgrad_kernel : d images -> d o -> delta_3 .. delta_0 (pre-activation cotangents), one 128-sample tile per
//                     slot, two slots ping-pong.  delta_{l-1} = (delta_l * W_l[:128]^T) .* (h_{l-1} > 0) runs on the
//                     tensor cores with delta_l as the TMEM A operand and weight images read K-major, two products per
//                     layer (W hi + W lo planes).  One-plane plan (PL = 1, the benchmarked one): cotangents are fp16 with
//                     ONE power-of-two scale per launch (tc_dout_kernel reduces max|d loss/d o|, tc_grad_scale) and the
//                     chain reads the FORWARD's fp16 weight images; two-plane plan (PL = 2, steps below 2^16 evaluated
//                     sample-frames): bf16 hi + lo planes of everything.  All three hidden weight layers stay resident
//                     in shared memory (192 KB, loaded once per CTA).
//   tc_wgrad_kernel : dW_l^T[n][k] = sum_s delta_l[s][n] * in_l[s][k] as MN-major x MN-major UMMAs over the sample
//                     axis, accumulated in TMEM (496 of 512 columns) across all tiles of a persistent CTA; bias
//                     gradients ride on the constant-one column of the feature image, dW4 on the CUDA cores.
//   tc_bwd_fused_kernel : both roles in ONE launch as a cluster of two CTAs (rank 0 = dgrad, rank 1 = wgrad) per SM
//                     pair.  The delta images go from the dgrad CTA to its partner through a small per-pair ring in
//                     global memory that stays L2-resident (74 pairs x 4 tile sets x 132 KB = 39 MB) instead of a
//                     per-frame HBM buffer; the hand-shake is mbarriers in the PARTNER's shared memory (remote arrive,
//                     acquire/release at cluster scope).  DSMEM is not the data path: it shares the SM's ~30 B/clk way
//                     out with the L2 stores (scripts/run_dsmem_probe.py).  The single-SM fusion does not fit: the wgrad
//                     accumulators take 496 of 512 TMEM columns.
//   tc_bwd_fused7_kernel (render_tc_bwd7.cuh, BHNERF_TC_BWD_V7=1): second generation of the pair, opt-in -- measured
//                     slower in every variant (profiles/r2_bwd_experiments.md).
#include <stdlib.h>
#include "tc_common.cuh"

using namespace tc;

namespace {

// Per-hop trace of the hand-off chain (-DBH_TC_TRACE, scripts/tc_trace_bwd.py): CTA 0 logs the low 32 bits of the clock at
// the hops of rounds 10 and 11 into the spare constant words 710.. of the workspace.  Compiled out by default.
#ifdef BH_TC_TRACE
#define BH_TRACE(idx) do { if (blockIdx.x == 0 && (r == 10 || r == 11)) { uint32_t _c; asm volatile("mov.u32 %0, %%clock;" : "=r"(_c)); \
    ((uint32_t*)((uint8_t*)status + TC_WS_CONST))[710 + (idx)] = _c; } } while (0)
#else
#define BH_TRACE(idx) do { } while (0)
#endif

// =====================================================================================================
// dgrad chain
// =====================================================================================================
#ifndef BH_DGRAD_PASSES
#define BH_DGRAD_PASSES 2     // products per layer of the one-plane chain: d*W_hi (+ d*W_lo)
#endif
constexpr int kDThreads = 576;           // 16 epilogue warps + MMA issuer + weight loader
constexpr int kDMmaWarp = 16;
constexpr uint32_t DW_BYTES = TC_W_BYTES - 16384u;                 // stages of layers 1..3, contiguous in ws
constexpr uint32_t D_SM_W = 0;
constexpr uint32_t D_SM_W4 = DW_BYTES;                             // 128 floats
constexpr uint32_t D_SM_BARS = D_SM_W4 + 512;
constexpr uint32_t D_SM_TOTAL = D_SM_BARS + 128;
#ifndef BH_RING_DEPTH
#define BH_RING_DEPTH 4
#endif
constexpr int kRingDepth = BH_RING_DEPTH;                          // tile sets per CTA pair in the global delta ring
constexpr uint32_t TSET_BYTES = 4u * TC_SIMG_BYTES + TC_AIMG_BYTES; // delta_0..3 images + aux image of one tile
enum { DB_WFULL = 0, DB_AREADY = 1, DB_DREADY = 3, DB_GFREE = 5 /* x kRingDepth, arrived by the wgrad CTA */ };
// fused mode: where the partner CTA of the pair is
struct PairLink {
  uint8_t* ring;          // this pair's kRingDepth tile sets in global memory
  uint32_t peer_bars;     // shared::cluster address of the partner's link barriers (wgrad: GFULL[d][l]; dgrad: GFREE[d])
};

#ifdef BH_EXP_NOSTG        // timing experiment: the delta images are not written (wrong results)
#define BH_DSTORE(p, v) do { if (((size_t)(p) & 1u)) *reinterpret_cast<uint4*>(p) = (v); } while (0)
#else
#define BH_DSTORE(p, v) (*reinterpret_cast<uint4*>(p) = (v))
#endif
// pack two fp32 cotangents and apply the relu mask m (0xffff per surviving half).  One-plane plan: ONE fp16 word (the
// cotangents carry the launch-wide power-of-two scale, tc_grad_scale); two-plane plan: bf16 hi / lo words.
template <int PL>
__device__ __forceinline__ void delta_pack(uint32_t m, float a, float b, uint32_t& hi, uint32_t& lo) {
  if (PL == 1) {
    hi = pack_f16x2(a, b) & m;
  } else {
    uint32_t p = pack_bf16x2(a, b);
    hi = p & m;
    lo = pack_bf16x2(a - bf16_lo(p), b - bf16_hi(p)) & m;
  }
}

// d loss / d o of every evaluated sample, once per backward, so that the dependent gathers (ray -> dI[ray]) are
// off the dgrad chain:  dout = e(1-e) * sum_c dI[b,c,ray] * w[c,i]   (sigmoid', network.py:230; kgeo.py:621).
// In place over e when the caller owns that buffer (dout == e is allowed: one thread reads and writes an element).
// Also reduces max|dout| of the launch into *dout_max (uint32 bits of a non-negative float order like the float; zeroed
// by the launcher): the scale of the fp16 cotangents (tc_grad_scale).
__global__ void __launch_bounds__(256)
tc_dout_kernel(PackedView v, const float* __restrict__ d_images, const float* e, int Bt, float* dout,
               uint32_t* __restrict__ dout_max) {
  const size_t n = (size_t)Bt * v.n_pad;
  float mx = 0.f;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / v.n_pad), i = (int)(idx - (size_t)b * v.n_pad);
    const int ray = v.ray[i];
    float g = 0.f;
    if (ray >= 0)
      for (int c = 0; c < v.S; ++c) g += d_images[((size_t)b * v.S + c) * v.P + ray] * v.w[(size_t)c * v.n_pad + i];
    const float ev = e[idx];
    const float d = g * ev * (1.f - ev);
    dout[idx] = d;
    mx = fmaxf(mx, fabsf(d));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(dout_max, __float_as_uint(mx));
}

template <int PL, bool FUSED, bool WIDE>
__device__ __forceinline__ void
dgrad_role(uint8_t* smem, const int cta, const int ncta, const PairLink link, const PackedView& v,
           const uint8_t* __restrict__ ws, const float* __restrict__ dout_all, int Bt, const int passes,
           const uint8_t* __restrict__ acts, uint8_t* __restrict__ deltas,
           float* __restrict__ d_params, int* __restrict__ status) {
  uint8_t* wsm = smem + D_SM_W;
  float* w4s = (float*)(smem + D_SM_W4);
  uint64_t* bars = (uint64_t*)(smem + D_SM_BARS);
  uint32_t* tmem_base_s = (uint32_t*)(bars + 10);
  int* abort_s = (int*)(tmem_base_s + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_frame = v.n_pad / 128;
  const int NT = Bt * tiles_per_frame;
  Abort ab{abort_s};

  if (tid == 0) {
    mbar_init(&bars[DB_WFULL], 1);
    mbar_init(&bars[DB_AREADY + 0], WIDE ? 16 : 8); mbar_init(&bars[DB_AREADY + 1], WIDE ? 16 : 8);
    mbar_init(&bars[DB_DREADY + 0], 1); mbar_init(&bars[DB_DREADY + 1], 1);
    for (int d = 0; d < kRingDepth; ++d) mbar_init(&bars[DB_GFREE + d], 1);
    *abort_s = 0;
    mbar_fence_init();
  }
  if (warp == kDMmaWarp) tmem_alloc(tmem_base_s, 512);
  if (tid < 128) w4s[tid] = ((const float*)(ws + TC_WS_CONST))[TC_C_W4 + tid];
  // one-plane plan: launch-wide power-of-two scale of the fp16 cotangents
  float ginv = 1.f;
  const float gscale = PL == 1 ? tc_grad_scale(((const uint32_t*)(ws + TC_WS_CONST))[TC_C_DOUTMAX], ginv) : 1.f;
  (void)ginv;
  tc_fence_before_sync();
  __syncthreads();
  if (FUSED) cluster_sync_all();          // the partner's barriers exist before anyone arrives on them
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_base_s;

  if (warp == kDMmaWarp + 1) {
    if (lane == 0) {      // resident weights: layers 1..3 [hi|lo] images, one shot
      if (PL == 1) {      // the forward's fp16 images (layer 3 = its 128-row hidden part): 3 x 64 KB, contiguous
        mbar_expect_tx(&bars[DB_WFULL], 196608u);
        const uint8_t* src = ws + TC_WS_W + 16384u;
        bulk_g2s(wsm, src, 65536u, &bars[DB_WFULL]);
        bulk_g2s(wsm + 65536u, src + 65536u, 65536u, &bars[DB_WFULL]);
        bulk_g2s(wsm + 131072u, src + 131072u, 65536u, &bars[DB_WFULL]);
      } else {            // bf16 images of the two-plane plan
        mbar_expect_tx(&bars[DB_WFULL], DW_BYTES);
        const uint8_t* src = ws + TC_WS_WB + 16384u;
        bulk_g2s(wsm, src, 65536u, &bars[DB_WFULL]);
        bulk_g2s(wsm + 65536u, src + 65536u, 65536u, &bars[DB_WFULL]);
        bulk_g2s(wsm + 131072u, src + 131072u, 81920u, &bars[DB_WFULL]);
      }
    }
    __syncwarp();
  } else if (warp == kDMmaWarp) {
    {   // whole warp runs the loop; elect.sync inside the issue wrappers picks the issuing lane
      const uint32_t idesc = PL == 1 ? make_idesc_f16(128, 128, 0, 0) : make_idesc(128, 128, 0, 0);   // A from TMEM, B K-major
      uint32_t a_phase[2] = {0u, 0u};
      BH_TIMING_T0 BH_TIMING_DECL(t_wa) BH_TIMING_DECL(t_is)
      bool ok = wait(&bars[DB_WFULL], 0, ab);
      for (int r = 0; ok; ++r) {
        int T0 = (r * ncta + cta) * 2;
        if (T0 >= NT) break;
        for (int l = 3; l >= 1 && ok; --l) {
          // one-plane plan: fp16 images of 128 rows each; two-plane plan: bf16 images, layer 3 with its 160 rows
          const uint32_t wl = smem_u32(wsm) + (PL == 1 ? (uint32_t)(l - 1) * 65536u : tc_stage_off(l) - 16384u);
          const uint32_t plane = PL == 1 ? 32768u : tc_plane_bytes(l), cs = PL == 1 ? 2048u : (tc_layer_K(l) / 8u) * 128u;
          for (int s = 0; s < 2; ++s) {
            if (T0 + s >= NT) continue;
            BH_TIMING_BEGIN
            ok = wait(&bars[DB_AREADY + s], a_phase[s], ab);
            BH_TIMING_END(t_wa)
            if (!ok) break;
            a_phase[s] ^= 1u;
            tc_fence_after_sync();
            const uint32_t td = tbase + (uint32_t)s * 256u, ta = td + 128u;
            if (lane == 0) BH_TRACE(((r - 10) * 6 + (3 - l) * 2 + s) * 2);
            BH_TIMING_BEGIN
            if (elect_one()) {
              // B' [K'=n][N'=k] = W_l[k][n]: the forward-layout image read K-major (K' groups = column groups);
              // passes: d_hi*W_hi, d_hi*W_lo, and with both planes d_lo*W_hi.  Unrolled, descriptor halves precomputed.
              const uint32_t b_hi = desc_hi(TC_IMG_RS), b_lo0 = desc_lo(wl, cs), b_lo1 = desc_lo(wl + plane, cs);
              const uint32_t kstep = (2u * cs) >> 4;
#pragma unroll
              for (int pp = 0; pp < (PL == 2 ? 3 : passes); ++pp)
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                  mma_ts_raw(td, ta + (pp == 2 ? 64u : 0u) + (uint32_t)ks * 8u, (pp == 1 ? b_lo1 : b_lo0) + (uint32_t)ks * kstep,
                             b_hi, idesc, (pp | ks) ? 1u : 0u);
              mma_commit_raw(&bars[DB_DREADY + s]);
            }
            __syncwarp();
            if (lane == 0) BH_TRACE(((r - 10) * 6 + (3 - l) * 2 + s) * 2 + 1);
            BH_TIMING_END(t_is)
          }
        }
      }
      if (lane == 0) { BH_TIMING_STORE(status, 20, t_wa) BH_TIMING_STORE(status, 22, t_is) }
    }
    __syncwarp();
  } else if (WIDE) {
    // Epilogue warps: (32-column group, TMEM lane quadrant); thread = one sample row x 32 columns.  ALL 16 warps work on
    // one tile at a time and alternate between the two slots layer by layer: the epilogue of one tile (half as long as
    // with 8 warps x 64 columns -- the chain is latency-bound) runs under the MMAs of the other.
    const int cgrp = warp >> 2, q = warp & 3, row = q * 32 + lane;
    const int half = cgrp >> 1, cc = cgrp & 1;                  // where these 32 columns sit in the ReLU mask words
    const uint32_t t_row = tbase + ((uint32_t)(q * 32) << 16);
    uint32_t d_phase[2] = {0u, 0u};
    float db4 = 0.f;
    bool ok = true;
    const size_t act_fs = tc_acts_bytes_per_frame(v.n_pad, PL), del_fs = tc_delta_bytes_per_frame(v.n_pad, PL);
    const size_t lstride = (size_t)v.n_pad * 256u, pstride = (size_t)v.n_pad * 1024u;
    uint32_t pub_addr[2] = {0u, 0u};
    // hand the finished tiles to the wgrad CTA: ONE proxy fence per warp covers the stores of both tiles of the round,
    // then one release per tile
    auto publish_both = [&]() {
      if (FUSED && (pub_addr[0] | pub_addr[1])) {
        fence_proxy_async_global();
        __syncwarp();
        if (lane == 0) {
          if (pub_addr[0]) mbar_arrive_remote(pub_addr[0]);
          if (pub_addr[1]) mbar_arrive_remote(pub_addr[1]);
        }
        pub_addr[0] = 0u; pub_addr[1] = 0u;
      }
    };
    // inputs of the next round's two tiles, loaded one round ahead: d loss/d o of the row and its four mask words
    float dout_next[2] = {0.f, 0.f};
    uint32_t mk_next[2][4];
    auto load_inputs = [&](int r) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int T = (r * ncta + cta) * 2 + s;
        if (T >= NT) continue;
        const int b = T / tiles_per_frame, tile = T - b * tiles_per_frame;
        dout_next[s] = dout_all[(size_t)b * v.n_pad + tile * 128 + row];
        const uint8_t* mbase = acts + (size_t)b * act_fs + tc_mask_off(v.n_pad, PL);
#pragma unroll
        for (int l = 0; l < 4; ++l)
          mk_next[s][l] = *reinterpret_cast<const uint32_t*>(mbase + tc_mask_word_off(tile, l, half, row) + cc * 4);
      }
    };
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int l = 0; l < 4; ++l) mk_next[s][l] = 0u;
    load_inputs(0);
    BH_TIMING_T0 BH_TIMING_DECL(t_top) BH_TIMING_DECL(t_gf) BH_TIMING_DECL(t_pub) BH_TIMING_DECL(t_wd) BH_TIMING_DECL(t_ep)
#ifdef BH_TC_TIMING
    const long long t_loop0 = clock64();
#endif
    // ---- top of a tile: delta_3[j] = dout * W4[j] * (h3[j] > 0) -> ring set + TMEM operand of the slot, hand to the MMA warp.
    // Compile-time variant -DBH_DGRAD_PIPELINE_TOP=1 (measured SLOWER, profiles/r2_bwd_experiments.md): the top of the NEXT
    // round's tile of a slot runs right after the slot's last accumulator (layer 1) has been pulled into registers, before
    // those values are packed and stored, so the tensor core already works on the next tile while this CTA finishes delta_0
    // and hands the round over.  It does not pay because the round is bound by the epilogue warps' own serial work, not by
    // the tensor pipe's idle time, and the earlier claim on a ring set couples the pair more tightly.
    uint8_t* del_cur[2] = {nullptr, nullptr};
    uint32_t rd_cur[2] = {0u, 0u};
    // mode 0: everything; 1: TMEM operand + hand to the MMA warp only; 2: ring stores (delta_3, aux) + db4 only
    auto do_top = [&](int r, int s, float dout_s, uint32_t mw, int mode) -> bool {
      const int T = (r * ncta + cta) * 2 + s, b = T / tiles_per_frame, tile = T - b * tiles_per_frame;
      const size_t dls = FUSED ? (size_t)TC_SIMG_BYTES : lstride;
      rd_cur[s] = (uint32_t)(2 * r + s) % kRingDepth;
      del_cur[s] = FUSED ? link.ring + (size_t)rd_cur[s] * TSET_BYTES : deltas + (size_t)b * del_fs + (size_t)tile * TC_SIMG_BYTES;
      uint8_t* aux = FUSED ? del_cur[s] + 4u * TC_SIMG_BYTES
                           : deltas + (size_t)b * del_fs + pstride * PL + (size_t)tile * TC_AIMG_BYTES;
      if (tid == 0 && s == 0) BH_TRACE(84 + (r - 10) * 4 + 2);
      if (FUSED && mode != 1) {       // ring set free: the partner has pulled tile rk - kRingDepth out of it
        BH_TIMING_BEGIN
        const bool good = wait_cluster(&bars[DB_GFREE + rd_cur[s]], (((uint32_t)(2 * r + s) / kRingDepth) & 1u) ^ 1u, ab);
        BH_TIMING_END(t_gf)
        if (!good) return false;
      }
      if (tid == 0) BH_TRACE(24 + (r - 10) * 30 + s * 2);
      BH_TIMING_BEGIN
      if (cgrp == 0 && mode != 1) {
        db4 += dout_s;
        const float dh = __bfloat162float(__float2bfloat16_rn(dout_s));      // aux image: col 0 = hi, col 1 = lo part of dout
        *reinterpret_cast<uint4*>(aux + sample_img_off(row, 0)) = make_uint4(pack_bf16x2(dh, dout_s - dh), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(aux + sample_img_off(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
      }
      uint32_t d[16], dl[16];
      const float douts = dout_s * gscale;
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) {
        const uint32_t off = sample_img_off(row, cgrp * 4 + gq);
        const float* w4 = w4s + (cgrp * 4 + gq) * 8;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          delta_pack<PL>(tc_mask_expand(mw, 4 * gq + jj), douts * w4[2 * jj], douts * w4[2 * jj + 1], d[4 * gq + jj], dl[4 * gq + jj]);
        if (mode != 1) {
#ifdef BH_EXP_NOD3STORE     // timing experiment: delta_3 is not written to the ring (wrong results)
          if (d[4 * gq] == 0x12345678u)
#endif
          BH_DSTORE(del_cur[s] + 3 * dls + off, make_uint4(d[4 * gq], d[4 * gq + 1], d[4 * gq + 2], d[4 * gq + 3]));
          if (PL == 2)
            *reinterpret_cast<uint4*>(del_cur[s] + pstride + 3 * dls + off) =
                make_uint4(dl[4 * gq], dl[4 * gq + 1], dl[4 * gq + 2], dl[4 * gq + 3]);
        }
      }
      if (mode != 2) {
        const uint32_t t_slot = t_row + (uint32_t)s * 256u;
        tmem_st16(t_slot + 128u + (uint32_t)(cgrp * 16), d);
        if (PL == 2) tmem_st16(t_slot + 192u + (uint32_t)(cgrp * 16), dl);
        tmem_wait_st();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[DB_AREADY + s]);
      }
      if (tid == 0) BH_TRACE(24 + (r - 10) * 30 + s * 2 + 1);
      BH_TIMING_END(t_top)
      return true;
    };
#ifndef BH_DGRAD_PIPELINE_TOP
#define BH_DGRAD_PIPELINE_TOP 0
#endif
    if (BH_DGRAD_PIPELINE_TOP && NT > cta * 2) {          // prologue: the tops of the first round
#pragma unroll
      for (int s = 0; s < 2; ++s)
        if (ok && cta * 2 + s < NT) ok = do_top(0, s, dout_next[s], mk_next[s][3], 1);
    }
    for (int r = 0; ok; ++r) {
      const int T0 = (r * ncta + cta) * 2;
      if (T0 >= NT) break;
      const bool has[2] = {true, T0 + 1 < NT};
      const int T0n = ((r + 1) * ncta + cta) * 2;
      const bool has_next[2] = {T0n < NT, T0n + 1 < NT};
      float dout[2];
      uint32_t mk[2][4];
      uint8_t* del_tile[2];
      uint32_t rd[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        dout[s] = dout_next[s];
#pragma unroll
        for (int l = 0; l < 4; ++l) mk[s][l] = mk_next[s][l];
      }
      if (tid == 0) BH_TRACE(84 + (r - 10) * 4);
      load_inputs(r + 1);
      if (tid == 0) BH_TRACE(84 + (r - 10) * 4 + 1);
#ifdef BH_EXP_LATE_PUBLISH   // experiment: the previous round is handed over here, after its stores have had ~0.9 k cycles to drain
      publish_both();
#endif
      const size_t dls = FUSED ? (size_t)TC_SIMG_BYTES : lstride;          // layer stride of the delta images
      {   // pipelined: the TMEM operand of these tiles went in during the previous round; their ring stores happen here
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (has[s] && ok) ok = do_top(r, s, dout[s], mk[s][3], BH_DGRAD_PIPELINE_TOP ? 2 : 0);
        if (!ok) break;
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) { del_tile[s] = del_cur[s]; rd[s] = rd_cur[s]; }
      // ---- the chain, alternating between the slots: D = delta_l * W_l^T  ->  delta_{l-1}
      for (int l = 3; l >= 1 && ok; --l) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (!has[s] || !ok) continue;
          uint8_t* d_img = del_tile[s] + (size_t)(l - 1) * dls;
          const uint32_t mw = mk[s][l - 1];
          const uint32_t t_slot = t_row + (uint32_t)s * 256u;
          BH_TIMING_BEGIN
          ok = wait(&bars[DB_DREADY + s], d_phase[s], ab);
          BH_TIMING_END(t_wd)
          if (!ok) break;
          d_phase[s] ^= 1u;
          tc_fence_after_sync();
          if (tid == 0) BH_TRACE(24 + (r - 10) * 30 + 4 + ((3 - l) * 2 + s) * 4);
          uint32_t raw[32];
          tmem_ld32(t_slot + (uint32_t)(cgrp * 32), raw);
          tmem_wait_ld();
          if (tid == 0) BH_TRACE(24 + (r - 10) * 30 + 4 + ((3 - l) * 2 + s) * 4 + 1);
          if (BH_DGRAD_PIPELINE_TOP && l == 1 && has_next[s]) {
            // the slot's accumulator is in registers and its last product is complete: the next tile's operand may go in
            tc_fence_before_sync();
            ok = do_top(r + 1, s, dout_next[s], mk_next[s][3], 1);
            if (!ok) break;
          }
          BH_TIMING_BEGIN
          uint32_t d[16], dl[16];
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const uint32_t off = sample_img_off(row, cgrp * 4 + gq);
            const uint32_t* rr = raw + 8 * gq;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              delta_pack<PL>(tc_mask_expand(mw, 4 * gq + jj), __uint_as_float(rr[2 * jj]), __uint_as_float(rr[2 * jj + 1]),
                             d[4 * gq + jj], dl[4 * gq + jj]);
            BH_DSTORE(d_img + off, make_uint4(d[4 * gq], d[4 * gq + 1], d[4 * gq + 2], d[4 * gq + 3]));
            if (PL == 2)
              *reinterpret_cast<uint4*>(d_img + pstride + off) = make_uint4(dl[4 * gq], dl[4 * gq + 1], dl[4 * gq + 2], dl[4 * gq + 3]);
          }
          if (l > 1) {
            tmem_st16(t_slot + 128u + (uint32_t)(cgrp * 16), d);
            if (PL == 2) tmem_st16(t_slot + 192u + (uint32_t)(cgrp * 16), dl);
            tmem_wait_st();
            tc_fence_before_sync();
            if (tid == 0) BH_TRACE(24 + (r - 10) * 30 + 4 + ((3 - l) * 2 + s) * 4 + 2);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[DB_AREADY + s]);
          }
          if (tid == 0) BH_TRACE(24 + (r - 10) * 30 + 4 + ((3 - l) * 2 + s) * 4 + 3);
          BH_TIMING_END(t_ep)
        }
      }
      if (!ok) break;
      if (FUSED) {
        // Hand both tiles over at the end of the round (the proxy fence + release costs ~0.9 k cycles per tile wherever it
        // is placed).  Measured alternatives, cycles per round: both deferred to the next round's delta_3 as the 8-warp
        // variant does: 17.2 k with the tensor-core dW4 (the pair serialises: the wgrad CTA idles until the next delta_3
        // while this CTA waits for ring sets the wgrad CTA has not started on), 13.5 k with the CUDA-core dW4; slot 0
        // right after its last epilogue: 13.6 k (the release waits for the fresh stores); this placement: 12.9 k.
        pub_addr[0] = link.peer_bars + rd[0] * 8u;
        if (has[1]) pub_addr[1] = link.peer_bars + rd[1] * 8u;
#ifndef BH_EXP_LATE_PUBLISH
        if (tid == 0) BH_TRACE(24 + (r - 10) * 30 + 28);
        BH_TIMING_BEGIN
        publish_both();
        BH_TIMING_END(t_pub)
        if (tid == 0) BH_TRACE(24 + (r - 10) * 30 + 29);
#endif
      }
    }
#ifdef BH_EXP_LATE_PUBLISH
    publish_both();                        // the last round
#endif
#ifdef BH_TC_TIMING
    if (tid == 0) {
      long long t_loop = clock64() - t_loop0;
      BH_TIMING_STORE(status, 24, t_top) BH_TIMING_STORE(status, 26, t_gf) BH_TIMING_STORE(status, 28, t_pub)
      BH_TIMING_STORE(status, 30, t_wd) BH_TIMING_STORE(status, 32, t_ep) BH_TIMING_STORE(status, 34, t_loop)
      if (FUSED) {   // per-pair record (scripts/tc_timing_bwd.py): loop cycles / 64 and the SM this dgrad CTA ran on
        uint32_t smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
        uint32_t* rec = (uint32_t*)((uint8_t*)status + TC_WS_CONST) + 710 + (blockIdx.x / 2) * 2;
        rec[0] = (uint32_t)(t_loop >> 6);
        rec[1] = smid | (smid + 1u) << 10 | ((uint32_t)((NT / 2 + ncta - 1 - cta) / ncta) << 20);
        atomicMax((unsigned long long*)(status + 50), (unsigned long long)t_loop);
      }
    }
#endif
    // d b4 = sum dout (network.py:64 bias of the last Dense)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) db4 += __shfl_xor_sync(0xffffffffu, db4, o);
    if (lane == 0 && db4 != 0.f) atomicAdd(d_params + OFF_B4, db4);
  } else {
    // epilogue warps: (slot, column half, TMEM lane quadrant); thread = one sample row x 64 columns
    const int slot = warp >> 3, half = (warp >> 2) & 1, q = warp & 3, row = q * 32 + lane;
    const uint32_t t_lane = tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * 256u;
    const int cg0 = half * 8;                      // first 8-column group of this warp
    uint32_t d_phase = 0;
    float db4 = 0.f;
    bool ok = true;
    const size_t act_fs = tc_acts_bytes_per_frame(v.n_pad, PL), del_fs = tc_delta_bytes_per_frame(v.n_pad, PL);
    const size_t lstride = (size_t)v.n_pad * 256u, pstride = (size_t)v.n_pad * 1024u;
    // fused: a finished tile is handed to the wgrad CTA with ONE release (MEMBAR.GPU: every store of the warp must
    // have reached L2), issued while the warp would wait for the next tile's first MMA anyway
    uint32_t pub_addr = 0u;
    auto publish_pending = [&]() {
      if (FUSED && pub_addr) {
#ifndef BH_EXP_NOFENCE
        fence_proxy_async_global();         // generic-proxy stores -> the partner's cp.async.bulk reads (async proxy)
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(pub_addr);
        pub_addr = 0u;
      }
    };
    // d loss / d o of this thread's sample (tc_dout_kernel), loaded one tile ahead
    auto load_dout = [&](int r) -> float {
      const int T = (r * ncta + cta) * 2 + slot;
      if (T >= NT) return 0.f;
      const int b = T / tiles_per_frame, tile = T - b * tiles_per_frame;
      return dout_all[(size_t)b * v.n_pad + tile * 128 + row];
    };
    float dout_next = load_dout(0);
    uint2 mk_next[4] = {make_uint2(0u, 0u), make_uint2(0u, 0u), make_uint2(0u, 0u), make_uint2(0u, 0u)};
    auto load_masks = [&](int r) {
      const int T = (r * ncta + cta) * 2 + slot;
      if (T >= NT) return;
      const int b = T / tiles_per_frame, tile = T - b * tiles_per_frame;
      const uint8_t* mbase = acts + (size_t)b * act_fs + tc_mask_off(v.n_pad, PL);
#pragma unroll
      for (int l = 0; l < 4; ++l) mk_next[l] = *reinterpret_cast<const uint2*>(mbase + tc_mask_word_off(tile, l, half, row));
    };
    load_masks(0);
    BH_TIMING_T0 BH_TIMING_DECL(t_top) BH_TIMING_DECL(t_gf) BH_TIMING_DECL(t_pub) BH_TIMING_DECL(t_wd) BH_TIMING_DECL(t_ep)
#ifdef BH_TC_TIMING
    const long long t_loop0 = clock64();
#endif
    for (int r = 0; ok; ++r) {
      int T0 = (r * ncta + cta) * 2;
      if (T0 >= NT) break;
      int T = T0 + slot;
      if (T >= NT) continue;
      const int b = T / tiles_per_frame, tile = T - b * tiles_per_frame;
      // delta images of this tile: the per-frame scratch, or (fused) tile set `rd` of the pair's ring
      const uint32_t rk = (uint32_t)(2 * r + slot), rd = rk % kRingDepth;
      uint8_t* del_tile = FUSED ? link.ring + (size_t)rd * TSET_BYTES : deltas + (size_t)b * del_fs + (size_t)tile * TC_SIMG_BYTES;
      const size_t dls = FUSED ? (size_t)TC_SIMG_BYTES : lstride;          // layer stride of the delta images
      uint8_t* aux = FUSED ? del_tile + 4u * TC_SIMG_BYTES
                           : deltas + (size_t)b * del_fs + pstride * PL + (size_t)tile * TC_AIMG_BYTES;
      // ReLU bit masks of this row's 64 columns, all four layers (loaded one tile ahead, like dout)
      uint2 mk[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) mk[l] = mk_next[l];
      const float dout = dout_next;
      dout_next = load_dout(r + 1);
      load_masks(r + 1);
      if (FUSED) {       // ring slot free: the partner has pulled tile rk - kRingDepth out of it
        BH_TIMING_BEGIN
        ok = wait_cluster(&bars[DB_GFREE + rd], ((rk / kRingDepth) & 1u) ^ 1u, ab);
        BH_TIMING_END(t_gf)
        if (!ok) break;
      }
      BH_TIMING_BEGIN
      if (half == 0) {
        db4 += dout;
        // aux image [128][16]: col 0 = bf16 hi part of dout, col 1 = lo part
        const float dh = __bfloat162float(__float2bfloat16_rn(dout));
        *reinterpret_cast<uint4*>(aux + sample_img_off(row, 0)) = make_uint4(pack_bf16x2(dh, dout - dh), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(aux + sample_img_off(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
      }
      // delta_3[j] = dout * W4[j] * (h3[j] > 0)
      const float douts = dout * gscale;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t d[16], dl[16];
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const uint32_t off = sample_img_off(row, cg0 + 4 * cc + gq);
          const uint32_t mw = cc ? mk[3].y : mk[3].x;
          const float* w4 = w4s + (cg0 + 4 * cc + gq) * 8;
          delta_pack<PL>(tc_mask_expand(mw, 4 * gq + 0), douts * w4[0], douts * w4[1], d[4 * gq + 0], dl[4 * gq + 0]);
          delta_pack<PL>(tc_mask_expand(mw, 4 * gq + 1), douts * w4[2], douts * w4[3], d[4 * gq + 1], dl[4 * gq + 1]);
          delta_pack<PL>(tc_mask_expand(mw, 4 * gq + 2), douts * w4[4], douts * w4[5], d[4 * gq + 2], dl[4 * gq + 2]);
          delta_pack<PL>(tc_mask_expand(mw, 4 * gq + 3), douts * w4[6], douts * w4[7], d[4 * gq + 3], dl[4 * gq + 3]);
          BH_DSTORE(del_tile + 3 * dls + off, make_uint4(d[4 * gq], d[4 * gq + 1], d[4 * gq + 2], d[4 * gq + 3]));
          if (PL == 2)
            *reinterpret_cast<uint4*>(del_tile + pstride + 3 * dls + off) =
                make_uint4(dl[4 * gq], dl[4 * gq + 1], dl[4 * gq + 2], dl[4 * gq + 3]);
        }
        tmem_st16(t_lane + 128u + (uint32_t)(half * 32 + cc * 16), d);
        if (PL == 2) tmem_st16(t_lane + 192u + (uint32_t)(half * 32 + cc * 16), dl);
      }
      tmem_wait_st();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[DB_AREADY + slot]);
      BH_TIMING_END(t_top)
      BH_TIMING_BEGIN
      publish_pending();                   // the previous tile of this slot (overlaps the MMA of delta_3)
      BH_TIMING_END(t_pub)
      for (int l = 3; l >= 1; --l) {       // D = delta_l * W_l^T  ->  delta_{l-1}
        uint8_t* d_img = del_tile + (size_t)(l - 1) * dls;
        const uint2 mkl = l == 3 ? mk[2] : (l == 2 ? mk[1] : mk[0]);
        BH_TIMING_BEGIN
        ok = wait(&bars[DB_DREADY + slot], d_phase, ab);
        BH_TIMING_END(t_wd)
        if (!ok) break;
        d_phase ^= 1u;
        tc_fence_after_sync();
        BH_TIMING_BEGIN
        uint32_t raw[2][32];
        tmem_ld32(t_lane + (uint32_t)(half * 64), raw[0]);
        tmem_ld32(t_lane + (uint32_t)(half * 64 + 32), raw[1]);
        tmem_wait_ld();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t d[16], dl[16];
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const uint32_t off = sample_img_off(row, cg0 + 4 * cc + gq);
            const uint32_t mw = cc ? mkl.y : mkl.x;
            const uint32_t* rr = raw[cc] + 8 * gq;
            delta_pack<PL>(tc_mask_expand(mw, 4 * gq + 0), __uint_as_float(rr[0]), __uint_as_float(rr[1]), d[4 * gq + 0], dl[4 * gq + 0]);
            delta_pack<PL>(tc_mask_expand(mw, 4 * gq + 1), __uint_as_float(rr[2]), __uint_as_float(rr[3]), d[4 * gq + 1], dl[4 * gq + 1]);
            delta_pack<PL>(tc_mask_expand(mw, 4 * gq + 2), __uint_as_float(rr[4]), __uint_as_float(rr[5]), d[4 * gq + 2], dl[4 * gq + 2]);
            delta_pack<PL>(tc_mask_expand(mw, 4 * gq + 3), __uint_as_float(rr[6]), __uint_as_float(rr[7]), d[4 * gq + 3], dl[4 * gq + 3]);
            BH_DSTORE(d_img + off, make_uint4(d[4 * gq], d[4 * gq + 1], d[4 * gq + 2], d[4 * gq + 3]));
            if (PL == 2)
              *reinterpret_cast<uint4*>(d_img + pstride + off) =
                  make_uint4(dl[4 * gq], dl[4 * gq + 1], dl[4 * gq + 2], dl[4 * gq + 3]);
          }
          if (l > 1) {
            tmem_st16(t_lane + 128u + (uint32_t)(half * 32 + cc * 16), d);
            if (PL == 2) tmem_st16(t_lane + 192u + (uint32_t)(half * 32 + cc * 16), dl);
          }
        }
        if (l > 1) {
          tmem_wait_st();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[DB_AREADY + slot]);
        }
        BH_TIMING_END(t_ep)
      }
      if (FUSED) pub_addr = link.peer_bars + rd * 8u;      // handed over during the next tile (or after the loop)
    }
    publish_pending();
#ifdef BH_TC_TIMING
    if (tid == 0) {
      long long t_loop = clock64() - t_loop0;
      BH_TIMING_STORE(status, 24, t_top) BH_TIMING_STORE(status, 26, t_gf) BH_TIMING_STORE(status, 28, t_pub)
      BH_TIMING_STORE(status, 30, t_wd) BH_TIMING_STORE(status, 32, t_ep) BH_TIMING_STORE(status, 34, t_loop)
    }
#endif
    // d b4 = sum dout (network.py:64 bias of the last Dense)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) db4 += __shfl_xor_sync(0xffffffffu, db4, o);
    if (lane == 0 && db4 != 0.f) atomicAdd(d_params + OFF_B4, db4);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (FUSED) cluster_sync_all();          // no CTA of the pair leaves while the other may still arrive on its barriers
  if (warp == kDMmaWarp) tmem_dealloc(tbase, 512);
  if (tid == 0 && *abort_s) raise_flag(status, 1);
}

template <int PL, bool WIDE>
__global__ void __launch_bounds__(kDThreads, 1)
tc_dgrad_kernel(PackedView v, const uint8_t* __restrict__ ws, const float* __restrict__ dout_all, int Bt, int passes,
                const uint8_t* __restrict__ acts, uint8_t* __restrict__ deltas,
                float* __restrict__ d_params, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  dgrad_role<PL, false, WIDE>(smem, (int)blockIdx.x, (int)gridDim.x, PairLink{nullptr, 0u}, v, ws, dout_all, Bt, passes, acts,
                        deltas, d_params, status);
}

// =====================================================================================================
// wgrad
// =====================================================================================================
constexpr int kWThreads = 192;             // warps 0-3 final epilogue, warp 4 MMA issuer, warp 5 producer
// One stage of the operand ring = the A image, the B image and the slice of the feature image that rides on the same
// MMA as extra N columns: [A 32K | B 32K | F 8K].  Small-N MMAs cost as much as N=128 ones in SS form (the A read
// from shared memory dominates: measured ~105 cycles per MMA for any N <= 128), so the skip/bias products are
// appended to the big ones (N = 160 / 144) instead of being issued on their own.
constexpr uint32_t W_STAGE_BYTES = 2u * TC_SIMG_BYTES + TC_FIMG_BYTES;
template <int PL> struct WCfg {
  static constexpr int kWStages = 3;
  static constexpr uint32_t SM_BARS = kWStages * W_STAGE_BYTES;
  static constexpr uint32_t SM_TOTAL = SM_BARS + 256;
};
// WB_GFULL[d]: the four delta images and the aux image of the tile in ring set d are written; 8 remote arrivals
// (the dgrad CTA's epilogue warps of that slot) per phase
enum { WB_FULL = 0, WB_EMPTY = 3, WB_DONE = 10, WB_GFULL = 11, WB_NBARS = 11 + 2 * kRingDepth };   // GFULL[partner][set]
static_assert(WB_NBARS * 8 + 16 <= 256, "wgrad barrier area");
// TMEM accumulator columns (lane = n, or j for dW4): [dW3 128 | dW3 skip rows + b3 32] [dW2 128 | b2 16]
// [dW1 128 | b1 16] [dW0 + b0 32] [dW4 16]
constexpr uint32_t ACC_W3 = 0, ACC_W3F = 128, ACC_W2 = 160, ACC_B2 = 288, ACC_W1 = 304, ACC_B1 = 432, ACC_W0F = 448,
                   ACC_W4 = 480;

// Per tile the CTA runs a list of stage fills ("sub-jobs").  job j picks the operands
//   0: delta_3 x [h2 | feat]   1: delta_2 x [h1 | feat cols 16..31]   2: delta_1 x [h0 | feat cols 16..31]
//   3: delta_0 x feat          4: h3 x aux
// (feature column 21 is the constant one: bias gradients) and with two planes every job runs its bf16 plane
// combinations (A plane pa, B plane pb):
//   jobs 0-3: (hi,hi) (lo,hi) (hi,lo)      job 4: (hi,-) (lo,-)   [aux carries both parts of dout]
template <int PL> __device__ __forceinline__ int wg_num_subjobs() { return PL == 2 ? 14 : 5; }
template <int PL> __device__ __forceinline__ void wg_subjob(int idx, int& j, int& pa, int& pb) {
  if (PL == 1) { j = idx; pa = 0; pb = 0; return; }
  if (idx < 12) { j = idx / 3; int c = idx - 3 * j; pa = (c == 1); pb = (c == 2); }
  else { j = 4; pa = idx - 12; pb = 0; }
}

// Tile order of one wgrad CTA.  Stand-alone: T = cta, cta + ncta, ...   Fused: the tile PAIRS of its ND dgrad partners,
// interleaved partner by partner.  Partner p is dgrad CTA cta*ND + p of ncta*ND; its it_p-th tile is
// (r * ncta*ND + cta*ND + p) * 2 + s with r = it_p / 2, s = it_p & 1 -- ring sequence number it_p, the numbering that dgrad
// CTA uses.  Returns 0 = all partners done, 1 = tile valid, 2 = skip (odd tail, or this partner is done already).
struct WgSeq { int T, p; uint32_t rd, ru; };
template <bool FUSED, int ND>
__device__ __forceinline__ int wg_tile(int it, int cta, int ncta, int NT, WgSeq& q) {
  if (!FUSED) { q.T = cta + it * ncta; q.p = 0; q.rd = 0u; q.ru = 0u; return q.T < NT ? 1 : 0; }
  const int p = it % ND, itp = it / ND;
  q.p = p; q.rd = (uint32_t)itp % kRingDepth; q.ru = ((uint32_t)itp / kRingDepth) & 1u;
  const int T0 = ((itp >> 1) * ncta * ND + cta * ND + p) * 2;
  if (T0 >= NT) return p == 0 ? 0 : 2;          // partner 0 has the smallest tile numbers: when it is done, all are
  q.T = T0 + (itp & 1);
  return q.T < NT ? 1 : 2;
}

template <int PL, bool FUSED, bool WIDE = false, int ND = 1>
__device__ __forceinline__ void
wgrad_role(uint8_t* smem, const int cta, const int ncta, const PairLink* links, int n_pad, int Bt,
           const uint8_t* __restrict__ acts, const uint8_t* __restrict__ deltas, float* __restrict__ d_params,
           int* __restrict__ status) {
  // status = first bytes of the tcgen05 workspace: the fp32 constants (and the cotangent scale) sit TC_WS_CONST behind it
  float ginv = 1.f;
  if (PL == 1) tc_grad_scale(((const uint32_t*)((const uint8_t*)status + TC_WS_CONST))[TC_C_DOUTMAX], ginv);
  static_assert(!(FUSED && PL == 2), "the fused pair runs the one-plane plan");
  constexpr int kWStages = WCfg<PL>::kWStages;
  constexpr uint32_t W_SM_BARS = WCfg<PL>::SM_BARS;
  uint64_t* bars = (uint64_t*)(smem + W_SM_BARS);
  uint32_t* tmem_base_s = (uint32_t*)(bars + WB_NBARS);
  int* abort_s = (int*)(tmem_base_s + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_frame = n_pad / 128;
  const int NT = Bt * tiles_per_frame;
  Abort ab{abort_s};
  const size_t act_fs = tc_acts_bytes_per_frame(n_pad, PL), del_fs = tc_delta_bytes_per_frame(n_pad, PL);
  const size_t lstride = (size_t)n_pad * 256u, pstride = (size_t)n_pad * 1024u;

  if (tid == 0) {
    for (int s = 0; s < kWStages; ++s) { mbar_init(&bars[WB_FULL + s], 1); mbar_init(&bars[WB_EMPTY + s], 1); }
    mbar_init(&bars[WB_DONE], 1);
    for (int s = 0; s < ND * kRingDepth; ++s) mbar_init(&bars[WB_GFULL + s], WIDE ? 16 : 8);
    abort_s[0] = 0; abort_s[1] = 0;
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(tmem_base_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  if (FUSED) cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_base_s;
  WgSeq q_first;
  const bool has_work = wg_tile<FUSED, ND>(0, cta, ncta, NT, q_first) == 1;

  if (warp == 5) {
    // ===================== producer: bulk copies of the saved images =====================
    if (lane == 0) {
      uint32_t cnt = 0;
      bool ok = true;
#ifdef BH_EXP_EVICT
      const uint64_t pol_stream = l2_policy_evict_first();
#endif
      BH_TIMING_T0 BH_TIMING_DECL(t_gfl) BH_TIMING_DECL(t_em)
      for (int it = 0; ok; ++it) {
        WgSeq sq;
        const int tv = wg_tile<FUSED, ND>(it, cta, ncta, NT, sq);
        if (tv == 0) break;
        if (tv == 2) continue;
        const int T = sq.T;
        const int b = T / tiles_per_frame, tile = T - b * tiles_per_frame;
        const uint8_t* act_tile = acts + (size_t)b * act_fs + (size_t)tile * TC_SIMG_BYTES;
        const uint8_t* feat_tile = acts + (size_t)b * act_fs + pstride * PL + (size_t)tile * TC_FIMG_BYTES;
        // fused: ring set rd of partner p; use number ru of that set selects the barrier phase
        const uint32_t rd = sq.rd, ru = sq.ru;
        const uint8_t* del_tile = FUSED ? links[sq.p].ring + (size_t)rd * TSET_BYTES
                                        : deltas + (size_t)b * del_fs + (size_t)tile * TC_SIMG_BYTES;
        const size_t dls = FUSED ? (size_t)TC_SIMG_BYTES : lstride;
        const uint8_t* aux_src = FUSED ? del_tile + 4u * TC_SIMG_BYTES
                                       : deltas + (size_t)b * del_fs + pstride * PL + (size_t)tile * TC_AIMG_BYTES;
        if (FUSED) {      // the dgrad CTA has written this tile's delta / aux images into ring set rd
          BH_TIMING_BEGIN
          ok = wait_cluster(&bars[WB_GFULL + sq.p * kRingDepth + rd], ru, ab);
          BH_TIMING_END(t_gfl)
          if (!ok) break;
          fence_proxy_async_global();
        }
        for (int idx = 0; idx < wg_num_subjobs<PL>() && ok; ++idx, ++cnt) {
          int j, pa, pb;
          wg_subjob<PL>(idx, j, pa, pb);
          uint32_t st = cnt % kWStages, ph = (cnt / kWStages) & 1u;
          BH_TIMING_BEGIN
          ok = wait(&bars[WB_EMPTY + st], ph ^ 1u, ab);
          BH_TIMING_END(t_em)
          if (!ok) break;
          uint8_t* dst = smem + st * W_STAGE_BYTES;
          uint64_t* full = &bars[WB_FULL + st];
          const uint8_t* a_src = (j < 4) ? del_tile + pa * pstride + (size_t)(3 - j) * dls
                                         : act_tile + pa * pstride + 3 * lstride;
          const uint8_t* f_src = feat_tile + (size_t)pb * ((size_t)n_pad * 64u);
          const uint32_t b_bytes = (j == 0) ? TC_SIMG_BYTES + TC_FIMG_BYTES
                                 : (j < 3) ? TC_SIMG_BYTES + TC_FIMG_BYTES / 2u
                                 : (j == 3) ? TC_FIMG_BYTES : TC_AIMG_BYTES;
          mbar_expect_tx(full, TC_SIMG_BYTES + b_bytes);
#ifdef BH_EXP_EVICT
          // saved activations are read once: evict-first, so that they do not push the cotangent ring out of L2
          if (FUSED && j == 4) bulk_g2s_hint(dst, a_src, TC_SIMG_BYTES, full, pol_stream); else
#endif
          bulk_g2s(dst, a_src, TC_SIMG_BYTES, full);
          uint8_t* bdst = dst + TC_SIMG_BYTES;
#ifdef BH_EXP_EVICT
          if (FUSED && j < 3) {
            bulk_g2s_hint(bdst, act_tile + pb * pstride + (size_t)(2 - j) * lstride, TC_SIMG_BYTES, full, pol_stream);
            if (j == 0) bulk_g2s(bdst + TC_SIMG_BYTES, f_src, TC_FIMG_BYTES, full);
            else bulk_g2s(bdst + TC_SIMG_BYTES, f_src + TC_FIMG_BYTES / 2u, TC_FIMG_BYTES / 2u, full);
          } else
#endif
          if (j < 3) {
            bulk_g2s(bdst, act_tile + pb * pstride + (size_t)(2 - j) * lstride, TC_SIMG_BYTES, full);
            if (j == 0) bulk_g2s(bdst + TC_SIMG_BYTES, f_src, TC_FIMG_BYTES, full);
            else bulk_g2s(bdst + TC_SIMG_BYTES, f_src + TC_FIMG_BYTES / 2u, TC_FIMG_BYTES / 2u, full);    // cols 16..31
          } else if (j == 3) {
            bulk_g2s(bdst, f_src, TC_FIMG_BYTES, full);
          } else {
            bulk_g2s(bdst, aux_src, TC_AIMG_BYTES, full);
          }
        }
      }
      BH_TIMING_STORE_B(status, 38, t_gfl, FUSED ? 1 : 0) BH_TIMING_STORE_B(status, 40, t_em, FUSED ? 1 : 0)
    }
    __syncwarp();
  } else if (warp == 4) {
    // ===================== MMA issuer (whole warp, elect.sync inside the issue wrappers) =====================
    {
      const uint32_t id160 = PL == 1 ? make_idesc_f16(128, 160, 1, 1) : make_idesc(128, 160, 1, 1),
                     id144 = PL == 1 ? make_idesc_f16(128, 144, 1, 1) : make_idesc(128, 144, 1, 1),
                     id32 = PL == 1 ? make_idesc_f16(128, 32, 1, 1) : make_idesc(128, 32, 1, 1),
                     id16 = PL == 1 ? make_idesc_f16(128, 16, 1, 1) : make_idesc(128, 16, 1, 1);
      uint32_t cnt = 0;
      bool ok = true;
      uint32_t later_tile = 0;               // 0 for the CTA's first tile: accumulators start from zero
      BH_TIMING_T0 BH_TIMING_DECL(t_fu) BH_TIMING_DECL(t_wi)
#ifdef BH_TC_TIMING
      const long long t_w0 = clock64();
#endif
      // all operands are [s][c] images read MN-major: K (= sample) groups advance by RS, M/N groups by CS
      for (int it = 0; ok; ++it) {
        WgSeq sq;
        const int tv = wg_tile<FUSED, ND>(it, cta, ncta, NT, sq);
        if (tv == 0) break;
        if (tv == 2) continue;
        for (int idx = 0; idx < wg_num_subjobs<PL>() && ok; ++idx, ++cnt) {
          int j, pa, pb;
          wg_subjob<PL>(idx, j, pa, pb);
          uint32_t st = cnt % kWStages, ph = (cnt / kWStages) & 1u;
          BH_TIMING_BEGIN
          ok = wait(&bars[WB_FULL + st], ph, ab);    // every consumer observes every phase (parity waits cannot skip one)
          BH_TIMING_END(t_fu)
          if (!ok) break;
          if (j == 4) continue;              // dW4 = h3^T dout runs on the CUDA cores of warps 0-3, which release the stage
          tc_fence_after_sync();
          BH_TIMING_BEGIN
          const uint32_t A = smem_u32(smem + st * W_STAGE_BYTES), B = A + TC_SIMG_BYTES;
          const uint32_t started = later_tile | ((pa | pb) ? 1u : 0u);     // (hi,hi) is each accumulator's first product
          const uint32_t acc_col = j == 0 ? ACC_W3 : j == 1 ? ACC_W2 : j == 2 ? ACC_W1 : j == 3 ? ACC_W0F : ACC_W4;
          const uint32_t idesc = j == 0 ? id160 : j < 3 ? id144 : j == 3 ? id32 : id16;
          if (elect_one()) {
            const uint32_t hi = desc_hi(TC_SIMG_CS), kstep = (2u * TC_IMG_RS) >> 4;
            const uint32_t a_lo = desc_lo(A, TC_IMG_RS), b_lo = desc_lo(B, TC_IMG_RS);
#pragma unroll
            for (uint32_t ks = 0; ks < 8; ++ks)
              mma_ss_raw(tbase + acc_col, a_lo + ks * kstep, hi, b_lo + ks * kstep, hi, idesc, started | (ks > 0 ? 1u : 0u));
            mma_commit_raw(&bars[WB_EMPTY + st]);
          }
          __syncwarp();
          BH_TIMING_END(t_wi)
        }
        later_tile = 1;
      }
      if (elect_one()) mma_commit_raw(&bars[WB_DONE]);
      __syncwarp();
#ifdef BH_TC_TIMING
      if (lane == 0) {
        long long t_wl = clock64() - t_w0;
        BH_TIMING_STORE_B(status, 44, t_fu, FUSED ? 1 : 0)
        BH_TIMING_STORE_B(status, 46, t_wi, FUSED ? 1 : 0) BH_TIMING_STORE_B(status, 48, t_wl, FUSED ? 1 : 0)
      }
#endif
    }
    __syncwarp();
  } else if (warp < 4 && has_work) {
    // ===================== dW4 on the CUDA cores =====================
    // dW4[j] = sum_s h3[s][j] * dout[s] is a 128 x 128 matrix-vector product per tile.  As an MMA (N = 16) it cost a full
    // SS-form instruction slot per K step -- 8 of the 40 MMAs of a tile, ~20 % of this CTA's issue time, which bounds the
    // fused pair -- for 16 K MACs.  These four warps idle until the final flush, so they take it: the stage holds the h3
    // image and the aux image (dout as bf16 hi + lo); thread (column group cg, row phase rp) reads one 16-byte chunk
    // (8 columns of one row) per row group -- a quarter-warp reads 128 contiguous bytes, no bank conflicts -- and keeps 8
    // running sums over all tiles of the CTA.
    bool ok = true;
    float w4acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) w4acc[c] = 0.f;
    const int rp = lane & 7, cg = warp * 4 + (lane >> 3);       // row phase; column group (8 columns) of this thread
    {
      uint32_t cnt = 0;
      for (int it = 0; ok; ++it) {
        WgSeq sq;
        const int tv = wg_tile<FUSED, ND>(it, cta, ncta, NT, sq);
        if (tv == 0) break;
        if (tv == 2) continue;
        for (int idx = 0; idx < wg_num_subjobs<PL>() && ok; ++idx, ++cnt) {
          int j, pa, pb;
          wg_subjob<PL>(idx, j, pa, pb);
          const uint32_t st = cnt % kWStages, ph = (cnt / kWStages) & 1u;
          ok = wait(&bars[WB_FULL + st], ph, ab);    // observe every phase, consume only job 4
          if (!ok) break;
          if (j != 4) continue;
          // fused: the aux image is the last thing pulled out of the ring set -> hand the set back to the dgrad CTA
          if (FUSED && tid == 0) mbar_arrive_remote(links[sq.p].peer_bars + sq.rd * 8u);
          const uint8_t* img = smem + st * W_STAGE_BYTES + cg * TC_SIMG_CS + rp * 16;
          const uint8_t* aux = smem + st * W_STAGE_BYTES + TC_SIMG_BYTES + rp * 16;
#pragma unroll 4
          for (int rg = 0; rg < 16; ++rg) {
            const uint32_t dw = *reinterpret_cast<const uint32_t*>(aux + rg * TC_IMG_RS);     // row rg*8+rp: (hi, lo) of dout
            const float dout = bf16_lo(dw) + bf16_hi(dw);
            const uint4 hv = *reinterpret_cast<const uint4*>(img + rg * TC_IMG_RS);
            if (PL == 1) {        // h3 saved as fp16
              w4acc[0] = fmaf(f16_lo(hv.x), dout, w4acc[0]); w4acc[1] = fmaf(f16_hi(hv.x), dout, w4acc[1]);
              w4acc[2] = fmaf(f16_lo(hv.y), dout, w4acc[2]); w4acc[3] = fmaf(f16_hi(hv.y), dout, w4acc[3]);
              w4acc[4] = fmaf(f16_lo(hv.z), dout, w4acc[4]); w4acc[5] = fmaf(f16_hi(hv.z), dout, w4acc[5]);
              w4acc[6] = fmaf(f16_lo(hv.w), dout, w4acc[6]); w4acc[7] = fmaf(f16_hi(hv.w), dout, w4acc[7]);
            } else {
              w4acc[0] = fmaf(bf16_lo(hv.x), dout, w4acc[0]); w4acc[1] = fmaf(bf16_hi(hv.x), dout, w4acc[1]);
              w4acc[2] = fmaf(bf16_lo(hv.y), dout, w4acc[2]); w4acc[3] = fmaf(bf16_hi(hv.y), dout, w4acc[3]);
              w4acc[4] = fmaf(bf16_lo(hv.z), dout, w4acc[4]); w4acc[5] = fmaf(bf16_hi(hv.z), dout, w4acc[5]);
              w4acc[6] = fmaf(bf16_lo(hv.w), dout, w4acc[6]); w4acc[7] = fmaf(bf16_hi(hv.w), dout, w4acc[7]);
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");      // all four warps are done with the stage
          if (tid == 0) mbar_arrive(&bars[WB_EMPTY + st]);
        }
      }
    }
    if (ok) {                                                  // dW4: sum the 8 row phases, one atomic per column
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float vsum = w4acc[c];
        vsum += __shfl_xor_sync(0xffffffffu, vsum, 1);
        vsum += __shfl_xor_sync(0xffffffffu, vsum, 2);
        vsum += __shfl_xor_sync(0xffffffffu, vsum, 4);
        if (PL == 1) vsum *= TC_RZ_UNBIAS;                           // h3 is rz-truncated too (see the flush below)
        if (rp == 0 && vsum != 0.f) atomicAdd(d_params + OFF_W4 + cg * 8 + c, vsum);
        if (rp == 0 && !(fabsf(vsum) <= 3.0e38f)) abort_s[1] = 1;
      }
    }
    // ===================== final epilogue: TMEM accumulators -> d_params (atomic accumulate) =====================
    ok = ok && wait(&bars[WB_DONE], 0, ab);
    tc_fence_after_sync();
    if (ok) {
      const int n = warp * 32 + lane;
      const uint32_t t_lane = tbase + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < ACC_W4; c0 += 16) {
        uint32_t raw[16];
        tmem_ld16(t_lane + c0, raw);
        tmem_wait_ld();
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const uint32_t c = c0 + (uint32_t)jj;
          // undo the cotangent scale (power of two: exact).  One-plane plan: the saved h0..h2 are the forward's fp16 hi
          // plane, rounded TOWARD ZERO -- every product with them is low by the mean relative truncation,
          // 2^-11 * (1/2)/ln 2 = 3.52e-4 for mantissas spread log-uniformly over a binade (measured on cfg1 / cfg2:
          // 3.5e-4 .. 3.6e-4 of dW1..dW3, profiles/r2_grad_ab_fp16.log); the flush removes that bias
          const bool from_h = PL == 1 && (c < ACC_W3F || (c >= ACC_W2 && c < ACC_B2) || (c >= ACC_W1 && c < ACC_B1));
          const float val = __uint_as_float(raw[jj]) * (from_h ? ginv * TC_RZ_UNBIAS : ginv);
          int dst = -1;
          if (c < ACC_W3F) dst = OFF_W3 + (int)c * 128 + n;
          else if (c < ACC_W2) { int kf = (int)(c - ACC_W3F); dst = kf < BH_NF ? OFF_W3 + (128 + kf) * 128 + n : (kf == TC_ONES_COL ? OFF_B3 + n : -1); }
          else if (c < ACC_B2) dst = OFF_W2 + (int)(c - ACC_W2) * 128 + n;
          else if (c < ACC_W1) dst = (c - ACC_B2 == TC_ONES_COL - 16) ? OFF_B2 + n : -1;
          else if (c < ACC_B1) dst = OFF_W1 + (int)(c - ACC_W1) * 128 + n;
          else if (c < ACC_W0F) dst = (c - ACC_B1 == TC_ONES_COL - 16) ? OFF_B1 + n : -1;
          else { int kf = (int)(c - ACC_W0F); dst = kf < BH_NF ? OFF_W0 + kf * 128 + n : (kf == TC_ONES_COL ? OFF_B0 + n : -1); }
          if (dst >= 0 && val != 0.f) atomicAdd(d_params + dst, val);
          if (dst >= 0 && !(fabsf(val) <= 3.0e38f)) abort_s[1] = 1;     // non-finite gradient (cotangent overflow): flag it
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (FUSED) cluster_sync_all();
  if (warp == 4) tmem_dealloc(tbase, 512);
  if (tid == 0 && *abort_s) raise_flag(status, 2);
  if (tid == 0 && abort_s[1]) raise_flag(status, 4);
}

template <int PL>
__global__ void __launch_bounds__(kWThreads, 1)
tc_wgrad_kernel(int n_pad, int Bt, const uint8_t* __restrict__ acts, const uint8_t* __restrict__ deltas,
                float* __restrict__ d_params, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const PairLink none{nullptr, 0u};
  wgrad_role<PL, false, false, 1>(smem, (int)blockIdx.x, (int)gridDim.x, &none, n_pad, Bt, acts, deltas, d_params, status);
}

// =====================================================================================================
// fused backward: cluster of two CTAs per SM pair, rank 0 = dgrad chain, rank 1 = wgrad (one-plane plan)
// =====================================================================================================
constexpr uint32_t F_SM_TOTAL = D_SM_TOTAL > WCfg<1>::SM_TOTAL ? D_SM_TOTAL : WCfg<1>::SM_TOTAL;

// ND dgrad CTAs (cluster ranks 0..ND-1) feed ONE wgrad CTA (rank ND).  ND = 1 is the production shape (74 pairs, all
// resident); ND = 2 is kept as a measured alternative (see bwd_dgrad_per_cluster).
template <bool WIDE, int ND>
__global__ void __cluster_dims__(ND + 1, 1, 1) __launch_bounds__(kDThreads, 1)
tc_bwd_fused_kernel(PackedView v, const uint8_t* __restrict__ ws, const float* __restrict__ dout_all, int Bt, int passes,
                    const uint8_t* __restrict__ acts, uint8_t* __restrict__ ring,
                    float* __restrict__ d_params, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t rank = cluster_ctarank();
  const int cl = (int)(blockIdx.x / (ND + 1)), ncl = (int)(gridDim.x / (ND + 1));
  if (rank < (uint32_t)ND) {
    PairLink link;
    link.ring = ring + (size_t)(cl * ND + (int)rank) * kRingDepth * TSET_BYTES;
    link.peer_bars = mapa_u32(smem_u32(smem + WCfg<1>::SM_BARS + (WB_GFULL + rank * kRingDepth) * 8), (uint32_t)ND);
    dgrad_role<1, true, WIDE>(smem, cl * ND + (int)rank, ncl * ND, link, v, ws, dout_all, Bt, passes, acts, nullptr, d_params, status);
  } else {
    PairLink links[ND];
#pragma unroll
    for (int p = 0; p < ND; ++p) {
      links[p].ring = ring + (size_t)(cl * ND + p) * kRingDepth * TSET_BYTES;
      links[p].peer_bars = mapa_u32(smem_u32(smem + D_SM_BARS + DB_GFREE * 8), (uint32_t)p);
    }
    wgrad_role<1, true, WIDE, ND>(smem, cl, ncl, links, v.n_pad, Bt, acts, nullptr, d_params, status);
  }
}

#include "render_tc_bwd7.cuh"

int g_num_sms_b = 0;
int num_sms_b() {
  if (g_num_sms_b == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms_b, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms_b <= 0) g_num_sms_b = 148;
  }
  return g_num_sms_b;
}

// BHNERF_TC_DGRAD_WIDE=0 selects the 8-warps-per-tile dgrad epilogue (A/B measurements)
bool dgrad_wide_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("BHNERF_TC_DGRAD_WIDE"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

// dgrad CTAs per wgrad CTA in the fused backward (BHNERF_TC_BWD_ND=1|2, default 1).  Measured on B200 (cfg2 x 25 frames):
// clusters of 3 are parity-green but SLOWER, 4.35 ms against 2.80 ms -- only 45 of them are resident at once (135 of 148
// SMs) and one wgrad CTA cannot feed on two rings: its stage ring stalls behind the CUDA-core dW4 and the dgrad CTAs wait
// 10 k cycles per round for free ring sets.
int bwd_dgrad_per_cluster() {
  static int nd = 0;
  if (nd == 0) { const char* e = getenv("BHNERF_TC_BWD_ND"); nd = (e && e[0] == '2') ? 2 : 1; }
  return nd;
}

bool bwd_v7_enabled() {                    // BHNERF_TC_BWD_V7=1 selects the second-generation fused pair (slower as measured: opt-in)
  static int on = -1;
  if (on < 0) { const char* e = getenv("BHNERF_TC_BWD_V7"); on = (e && e[0] == '1') ? 1 : 0; }
  return on == 1;
}

bool bwd_fused_enabled() {                 // BHNERF_TC_FUSED=0 selects the two-kernel backward (A/B measurements)
  static int on = -1;
  if (on < 0) { const char* e = getenv("BHNERF_TC_FUSED"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

template <int PL>
int launch_bwd(const PackedView& v, const void* ws, const float* d_images, int Bt, const float* e_saved, const void* acts,
               void* delta_ws, float* dout, float* d_params, cudaStream_t st) {
  int* status = (int*)((uint8_t*)ws + TC_WS_STATUS);
  const int NT = Bt * (v.n_pad / 128);
  // products per layer of the one-plane chain: d*W_hi + d*W_lo; the stated fast mode (BHNERF_PRECISION=fast) drops W_lo
  const int dgrad_passes = (PL == 1 && bh_tc_fast()) ? 1 : BH_DGRAD_PASSES;
  {
    BhProfScope ps(BH_CAT_HEADS, 1, st);
    size_t n = (size_t)Bt * v.n_pad;
    int grid = (int)((n + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    uint32_t* dout_max = (uint32_t*)((uint8_t*)ws + TC_WS_CONST) + TC_C_DOUTMAX;
    static_assert(TC_C_SCHED == TC_C_DOUTMAX + 1, "one memset clears both words");
    BH_CHECK_CUDA(cudaMemsetAsync(dout_max, 0, 2 * sizeof(uint32_t), st));
    tc_dout_kernel<<<grid, 256, 0, st>>>(v, d_images, e_saved, Bt, dout, dout_max);
    BH_CHECK_CUDA(cudaGetLastError());
  }
  if (PL == 1 && bwd_fused_enabled() && bwd_v7_enabled() && !bh_tc_fast()) {
    BhProfScope ps(BH_CAT_BWD, 1, st);
    auto kern = tc_bwd_fused7_kernel;
    BH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F7_SM_TOTAL));
    static int max_cl7 = 0;
    if (max_cl7 == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * (num_sms_b() / 2)); cfg.blockDim = dim3(kDThreads); cfg.dynamicSmemBytes = F7_SM_TOTAL;
      cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, (const void*)kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms_b() / 2; }
      max_cl7 = n;
    }
    int ncl = (NT + 1) / 2; if (ncl > num_sms_b() / 2) ncl = num_sms_b() / 2; if (ncl > max_cl7) ncl = max_cl7;
    kern<<<2 * ncl, kDThreads, F7_SM_TOTAL, st>>>(v, (const uint8_t*)ws, dout, Bt, (const uint8_t*)acts, (uint8_t*)delta_ws,
                                                  d_params, status);
    BH_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (PL == 1 && bwd_fused_enabled()) {
    BhProfScope ps(BH_CAT_BWD, 1, st);
    const int nd = bwd_dgrad_per_cluster();
    auto kern = !dgrad_wide_enabled() ? tc_bwd_fused_kernel<false, 1>
                                      : (nd == 2 ? tc_bwd_fused_kernel<true, 2> : tc_bwd_fused_kernel<true, 1>);
    BH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SM_TOTAL));
    const int ndc = dgrad_wide_enabled() ? nd : 1;
    int ncl = ((NT + 1) / 2 + ndc - 1) / ndc; if (ncl > num_sms_b() / (ndc + 1)) ncl = num_sms_b() / (ndc + 1);
    // a cluster lives inside one GPC: launch no more clusters than can be resident at once (persistent kernel, a second
    // wave would double its time)
    static int max_cl[3] = {0, 0, 0};
    if (max_cl[ndc] == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((ndc + 1) * (num_sms_b() / (ndc + 1))); cfg.blockDim = dim3(kDThreads); cfg.dynamicSmemBytes = F_SM_TOTAL;
      cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = ndc + 1; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, (const void*)kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms_b() / (ndc + 1); }
      max_cl[ndc] = n;
      if (getenv("BHNERF_DEBUG")) fprintf(stderr, "bhnerf_b200: fused backward, clusters of %d: %d resident at once\n", ndc + 1, n);
    }
    if (ncl > max_cl[ndc]) ncl = max_cl[ndc];
    kern<<<(ndc + 1) * ncl, kDThreads, F_SM_TOTAL, st>>>(v, (const uint8_t*)ws, dout, Bt, dgrad_passes, (const uint8_t*)acts,
                                                         (uint8_t*)delta_ws, d_params, status);
    BH_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  {
    BhProfScope ps(BH_CAT_BWD, 1, st);
    auto kern = dgrad_wide_enabled() ? tc_dgrad_kernel<PL, true> : tc_dgrad_kernel<PL, false>;
    BH_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D_SM_TOTAL));
    int grid = (NT + 1) / 2; if (grid > num_sms_b()) grid = num_sms_b();
    kern<<<grid, kDThreads, D_SM_TOTAL, st>>>(v, (const uint8_t*)ws, dout, Bt, dgrad_passes, (const uint8_t*)acts, (uint8_t*)delta_ws,
                                              d_params, status);
    BH_CHECK_CUDA(cudaGetLastError());
  }
  {
    BhProfScope ps(BH_CAT_WGRAD, 1, st);
    BH_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WCfg<PL>::SM_TOTAL));
    int grid = NT < num_sms_b() ? NT : num_sms_b();
    tc_wgrad_kernel<PL><<<grid, kWThreads, WCfg<PL>::SM_TOTAL, st>>>(v.n_pad, Bt, (const uint8_t*)acts, (const uint8_t*)delta_ws,
                                                             d_params, status);
    BH_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace

// Backward scratch beyond the saved activations.  One-plane plan: the fixed delta ring of the fused CTA pairs
// (frame-count independent, L2-resident); two-plane plan or BHNERF_TC_FUSED=0: per-frame delta images.
size_t bh_tc_delta_bytes_per_frame(int n_pad, int planes) {
  return (planes == 1 && bwd_fused_enabled()) ? 0 : tc_delta_bytes_per_frame(n_pad, planes);
}
size_t bh_tc_delta_fixed_bytes(int planes) {
  return (planes == 1 && bwd_fused_enabled()) ? (size_t)(2 * (num_sms_b() / 3) + 2) * kRingDepth * TSET_BYTES : 0;   // >= dgrad CTAs of either cluster shape
}

int bh_tc_bwd(const PackedView& v, const void* ws, const float* params, const float* d_images, int Bt,
              const float* e_saved, const void* acts, void* delta_ws, float* dout_ws, int planes, float* d_params,
              cudaStream_t st) {
  (void)params;
  if (!e_saved || !acts || !delta_ws || !dout_ws) { bh_set_error("bh_tc_bwd: saved e / activations / delta + dout scratch required"); return 1; }
  return planes == 2 ? launch_bwd<2>(v, ws, d_images, Bt, e_saved, acts, delta_ws, dout_ws, d_params, st)
                     : launch_bwd<1>(v, ws, d_images, Bt, e_saved, acts, delta_ws, dout_ws, d_params, st);
}

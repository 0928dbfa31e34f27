// Fused backward, second generation ("v7"): the same CTA pair as tc_bwd_fused_kernel (rank 0 = dgrad chain, rank 1 = wgrad,
// cotangent images handed over through the per-pair ring in L2), re-plumbed around what the cycle accounting of the first
// generation showed (profiles/r2_bwd_experiments.md): both CTAs were bound by latencies, not by the tensor pipe.
//
//   dgrad CTA   * one fp16 product per layer (W hi plane only; measured gradient error 2.4e-4, tolerance 1e-3): the
//                 resident weights shrink from 192 KB to 96 KB, which pays for
//               * shared-memory STAGING of the cotangent images: the epilogue warps write an image once with st.shared, a
//                 dedicated store thread ships it to the ring with cp.async.bulk (shared -> global) and hands the tile to
//                 the partner after cp.async.bulk.wait_group -- no epilogue warp ever waits for a global store to land
//                 (the per-warp fence.proxy.async.global of the first generation cost ~1.8 k of 12.9 k cycles per round,
//                 and the 16-byte st.global bursts more).
//   wgrad CTA   * TWO operand rings instead of one: X = the saved activations (from HBM: the long latency), four slots
//                 deep and filled ahead of the dgrad CTA's progress; Y = the cotangent images (from the L2 ring), two
//                 slots, filled as tiles are handed over.  With one 3-stage ring of [A|B|F] stages the MMA warp waited
//                 ~6.8 k of 12.9 k cycles per round for loads.
// Included inside the anonymous namespace of render_tc_bwd.cu (shares its constants and helpers).

// ---------------- the ring of this generation ----------------
// A tile set holds delta_0..2 only (delta_3 is rebuilt by the partner), and the ring is deeper than the first generation's:
// the store thread hands a tile over two images late (so that it never waits for a write to land), which costs slack.
#ifndef BH_V7_RING
#define BH_V7_RING 7
#endif
#ifndef BH_V7_LAG
#define BH_V7_LAG 2
#endif
constexpr int kRing7 = BH_V7_RING;
constexpr uint32_t TSET7 = 3u * TC_SIMG_BYTES;
static_assert((size_t)kRing7 * TSET7 <= (size_t)(100 * kRingDepth) * TSET_BYTES / 74, "v7 ring fits the first generation's allocation");

// ---------------- shared-memory layouts ----------------
constexpr uint32_t D7_W_BYTES = 3u * 32768u;                      // W1, W2, W3[:128] hi planes
constexpr int kStg = 3;                                           // staging buffers of one cotangent image each
constexpr uint32_t D7_SM_W = 0;
constexpr uint32_t D7_SM_STG = D7_W_BYTES;                         // kStg x 32 KB
constexpr uint32_t D7_SM_W4 = D7_SM_STG + kStg * TC_SIMG_BYTES;    // 128 floats
constexpr uint32_t D7_SM_BARS = D7_SM_W4 + 512u;
constexpr uint32_t D7_SM_TOTAL = D7_SM_BARS + 320u;
enum { D7_WFULL = 0, D7_AREADY = 1, D7_DREADY = 3, D7_GFREE = 5, D7_SFULL = 5 + kRing7, D7_SFREE = D7_SFULL + kStg,
       D7_NBARS = D7_SFREE + kStg };
static_assert(D7_NBARS * 8 + 16 <= 256, "dgrad v7 barrier area (the schedule slots sit at +256)");

// X slots: [act 32 KB | feature slice]: slot 0 (h2 + all 32 feature columns), slots 1, 2 (h1 / h0 + feature columns 16..31);
// "slot 3" is the buffer the generator warps rebuild delta_3 in.  Y slots: slot 0 (delta_1), slot 1 (delta_2 / delta_0 + the
// whole feature image for the delta_0 x feat job).  An item always lands in the same slot.
constexpr uint32_t W7_X0 = 0, W7_X1 = W7_X0 + 32768u + 8192u, W7_X2 = W7_X1 + 32768u + 4096u, W7_X3 = W7_X2 + 32768u + 4096u;
constexpr uint32_t W7_Y0 = W7_X3 + 32768u, W7_Y1 = W7_Y0 + 32768u;
constexpr uint32_t W7_W4 = W7_Y1 + 32768u + 8192u;                // 128 floats
constexpr uint32_t W7_SM_BARS = W7_W4 + 512u;
constexpr uint32_t W7_SM_TOTAL = W7_SM_BARS + 320u;
// Y barriers are per JOB (j = 0..3: delta_3, delta_2, delta_1, delta_0), one phase per tile: jobs j and j+2 share a slot but are
// filled by different agents (delta_3 is rebuilt by the generator warps, the rest is loaded), and an agent that watched only
// every other phase of a shared barrier could not tell "not yet" from "two phases on".
enum { W7_XFULL = 0, W7_XEMPTY = 4, W7_YFULL = 8, W7_YEMPTY = 12, W7_DONE = 16, W7_GFULL = 17, W7_NBARS = 17 + kRing7 };
constexpr int kGenWarp0 = 8, kGenWarps = 8;                       // wgrad CTA: the warps that rebuild delta_3 locally
static_assert(W7_NBARS * 8 + 16 <= 256, "wgrad v7 barrier area (the schedule slots sit at +256)");
static_assert(W7_SM_TOTAL <= 232448 && D7_SM_TOTAL <= 232448, "v7 shared memory");
constexpr uint32_t F7_SM_TOTAL = D7_SM_TOTAL > W7_SM_TOTAL ? D7_SM_TOTAL : W7_SM_TOTAL;
__device__ __forceinline__ uint32_t w7_xslot(int i) { return i == 0 ? W7_X0 : (i == 1 ? W7_X1 : (i == 2 ? W7_X2 : W7_X3)); }

// ---------------- dynamic schedule ----------------
// The CTA pairs do not run at the same speed (a static round-robin finished its slowest pair at 17.0 k cycles per round
// against 12.8 k for pair 0: distance to the L2 slices that hold the ring and the activations differs by GPC), so tile pairs
// are handed out by an atomic counter.  The dgrad CTA's store thread draws round k+2 when it starts round k and publishes
// {k+1, first tile} as ONE 8-byte word into slot k % 8 of both CTAs' shared memory; every role reads the tile of its round
// from there (all roles of a pair are within 5 rounds of each other, see the ring depth).
constexpr uint32_t SCHED7_OFF = 256u;                             // inside the 320-byte barrier area of either role
constexpr int kSched = 8;
__device__ __forceinline__ int sched_get(const uint8_t* sched, int k, const Abort& ab) {
  const uint32_t a = smem_u32(sched) + (uint32_t)(k & (kSched - 1)) * 8u;
  for (uint32_t i = 0; i < (1u << 24); ++i) {
    uint32_t seq, t0;
    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(seq), "=r"(t0) : "r"(a) : "memory");
    if (seq == (uint32_t)k + 1u) return (int)t0;
    if ((i & 255u) == 255u && *ab.flag) return -1;
  }
  *ab.flag = 1;
  return -1;
}
__device__ __forceinline__ void sched_put(uint8_t* sched, uint32_t peer_sched, int k, int T0) {
  const uint32_t off = (uint32_t)(k & (kSched - 1)) * 8u;
  asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(smem_u32(sched) + off), "r"((uint32_t)k + 1u), "r"((uint32_t)T0) : "memory");
  asm volatile("st.volatile.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(peer_sched + off), "r"((uint32_t)k + 1u), "r"((uint32_t)T0) : "memory");
}
// tile `it` of this pair (two per round): 0 = no more work, 1 = valid, 2 = the odd tail of the last round
__device__ __forceinline__ int wg_tile7(int it, const uint8_t* sched, int NT, WgSeq& q, const Abort& ab) {
  q.p = 0; q.rd = (uint32_t)it % kRing7; q.ru = ((uint32_t)it / kRing7) & 1u;
  const int T0 = sched_get(sched, it >> 1, ab);
  if (T0 < 0) return 0;
  q.T = T0 + (it & 1);
  return q.T < NT ? 1 : 2;
}

// wait on a LOCAL barrier the partner CTA arrives on: poll at CTA scope, acquire at cluster scope once the phase is over
__device__ __forceinline__ bool wait_partner(uint64_t* bar, uint32_t parity, const Abort& ab) {
#ifdef BH_EXP_POLLCLUSTER
  return wait_cluster(bar, parity, ab);
#else
  const bool ok = wait(bar, parity, ab);
  asm volatile("fence.acq_rel.cluster;" ::: "memory");
  return ok;
#endif
}

__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// =====================================================================================================
// dgrad role, v7
// =====================================================================================================
__device__ __forceinline__ void
dgrad_role_v7(uint8_t* smem, const uint32_t peer_sched, const PairLink link, const PackedView& v,
              const uint8_t* __restrict__ ws, const float* __restrict__ dout_all, int Bt,
              const uint8_t* __restrict__ acts, float* __restrict__ d_params, int* __restrict__ status) {
  uint8_t* wsm = smem + D7_SM_W;
  uint8_t* stg = smem + D7_SM_STG;
  float* w4s = (float*)(smem + D7_SM_W4);
  uint64_t* bars = (uint64_t*)(smem + D7_SM_BARS);
  uint8_t* sched = smem + D7_SM_BARS + SCHED7_OFF;
  uint32_t* tmem_base_s = (uint32_t*)(bars + D7_NBARS);
  int* abort_s = (int*)(tmem_base_s + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_frame = v.n_pad / 128;
  const int NT = Bt * tiles_per_frame;
  Abort ab{abort_s};

  if (tid == 0) {
    mbar_init(&bars[D7_WFULL], 1);
    mbar_init(&bars[D7_AREADY + 0], 16); mbar_init(&bars[D7_AREADY + 1], 16);
    mbar_init(&bars[D7_DREADY + 0], 1); mbar_init(&bars[D7_DREADY + 1], 1);
    for (int d = 0; d < kRing7; ++d) mbar_init(&bars[D7_GFREE + d], 1);         // MMA warp of the partner
    for (int b = 0; b < kStg; ++b) { mbar_init(&bars[D7_SFULL + b], 16); mbar_init(&bars[D7_SFREE + b], 1); }
    for (int k = 0; k < 2 * kSched; ++k) reinterpret_cast<volatile uint32_t*>(sched)[k] = 0u;
    *abort_s = 0;
    mbar_fence_init();
  }
  if (warp == kDMmaWarp) tmem_alloc(tmem_base_s, 512);
  if (tid < 128) w4s[tid] = ((const float*)(ws + TC_WS_CONST))[TC_C_W4 + tid];
  float ginv = 1.f;
  const float gscale = tc_grad_scale(((const uint32_t*)(ws + TC_WS_CONST))[TC_C_DOUTMAX], ginv);
  (void)ginv;
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                     // the partner's barriers exist before anyone arrives on them
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_base_s;

  if (warp == kDMmaWarp + 1) {
    // ===================== loader + STORE thread =====================
    if (lane == 0) {
      mbar_expect_tx(&bars[D7_WFULL], D7_W_BYTES);        // resident weights: the hi planes of the forward's fp16 images
      const uint8_t* src = ws + TC_WS_W + 16384u;
      bulk_g2s(wsm, src, 32768u, &bars[D7_WFULL]);
      bulk_g2s(wsm + 32768u, src + 65536u, 32768u, &bars[D7_WFULL]);
      bulk_g2s(wsm + 65536u, src + 131072u, 32768u, &bars[D7_WFULL]);
      // images in the order the epilogue warps stage them: delta_2 of both tiles, then delta_1, delta_0.  (delta_3 never
      // leaves the SM: the partner rebuilds it from d loss/d o, W4 and the layer-3 masks -- an SM moves only ~30 B/clk out,
      // to L2 or to its partner's shared memory alike (scripts/run_epi_probe.py), and that is what bounds this CTA.)
      uint32_t q = 0, rel = 0;                            // image sequence number (-> staging buffer q % kStg); images released
      uint32_t pend_tk[2] = {0u, 0u}, pend_q[2] = {0u, 0u};     // tiles whose last image (sequence number pend_q) is in flight
      int npend = 0;
      bool ok = true;
      int* sched_ctr = (int*)(ws + TC_WS_CONST) + TC_C_SCHED;
      auto put = [&](int k, int drawn) {                  // round k of this pair <- tile pair `drawn`
        const int T0 = drawn * 2;
        sched_put(sched, peer_sched, k, T0 < NT ? T0 : -1);
      };
      put(0, atomicAdd(sched_ctr, 1)); put(1, atomicAdd(sched_ctr, 1));
      int drawn_next = atomicAdd(sched_ctr, 1);           // drawn one round before it is published: nobody waits for the atomic
      auto hand_over = [&](uint32_t tk) {
        __threadfence();
        mbar_arrive_remote(link.peer_bars + (tk % kRing7) * 8u);
      };
      for (int r = 0; ok; ++r) {
        const int T0 = sched_get(sched, r, ab);
        if (T0 < 0) break;
        put(r + 2, drawn_next);
        drawn_next = atomicAdd(sched_ctr, 1);
        const bool has1 = T0 + 1 < NT;
        for (int step = 1; step < 4 && ok; ++step) {      // step 1 = delta_2, ..., 3 = delta_0
          for (int s = 0; s < 2 && ok; ++s) {
            if (s == 1 && !has1) continue;
            const uint32_t tk = (uint32_t)(2 * r + s), rd = tk % kRing7;
            uint8_t* set = link.ring + (size_t)rd * TSET7;
            if (step == 1) {                              // first store into the set: the partner must have emptied it
              ok = wait_partner(&bars[D7_GFREE + rd], ((tk / kRing7) & 1u) ^ 1u, ab);
              if (!ok) break;
            }
            const uint32_t b = q % kStg;
            ok = wait(&bars[D7_SFULL + b], (q / kStg) & 1u, ab);
            if (!ok) break;
            bulk_s2g(set + (size_t)(3 - step) * TC_SIMG_BYTES, stg + b * TC_SIMG_BYTES, TC_SIMG_BYTES);
            bulk_commit();
            if (npend > 0 && q >= pend_q[0] + (uint32_t)BH_V7_LAG) {
              // a tile's last image is BH_V7_LAG groups old: wait for its WRITE (not only the read of the staging buffer): it
              // has landed long ago, then hand the tile over
              asm volatile("cp.async.bulk.wait_group %0;" ::"n"(BH_V7_LAG) : "memory");
              while (npend > 0 && q >= pend_q[0] + (uint32_t)BH_V7_LAG) {
                hand_over(pend_tk[0]);
                pend_tk[0] = pend_tk[1]; pend_q[0] = pend_q[1]; --npend;
              }
            } else {
              // at most the two newest groups may still be reading their staging buffers: images <= q-2 are free again
              bulk_wait_read<2>();
            }
            while (rel + 2u <= q) { mbar_arrive(&bars[D7_SFREE + rel % kStg]); ++rel; }
            if (step == 3) { pend_tk[npend] = tk; pend_q[npend] = q; ++npend; }
            ++q;
          }
        }
      }
      bulk_wait_all();
      if (ok) for (int i = 0; i < npend; ++i) hand_over(pend_tk[i]);
      while (rel < q) { mbar_arrive(&bars[D7_SFREE + rel % kStg]); ++rel; }
      bulk_wait_all();
    }
    __syncwarp();
  } else if (warp == kDMmaWarp) {
    {
      const uint32_t idesc = make_idesc_f16(128, 128, 0, 0);      // A from TMEM, B K-major
      uint32_t a_phase[2] = {0u, 0u};
      BH_TIMING_T0 BH_TIMING_DECL(t_wa) BH_TIMING_DECL(t_is)
      bool ok = wait(&bars[D7_WFULL], 0, ab);
      for (int r = 0; ok; ++r) {
        const int T0 = sched_get(sched, r, ab);
        if (T0 < 0) break;
        for (int l = 3; l >= 1 && ok; --l) {
          const uint32_t wl = smem_u32(wsm) + (uint32_t)(l - 1) * 32768u;
          for (int s = 0; s < 2; ++s) {
            if (T0 + s >= NT) continue;
            BH_TIMING_BEGIN
            ok = wait(&bars[D7_AREADY + s], a_phase[s], ab);
            BH_TIMING_END(t_wa)
            if (!ok) break;
            a_phase[s] ^= 1u;
            tc_fence_after_sync();
            const uint32_t td = tbase + (uint32_t)s * 256u, ta = td + 128u;
            BH_TIMING_BEGIN
            if (elect_one()) {
              const uint32_t b_hi = desc_hi(TC_IMG_RS), b_lo0 = desc_lo(wl, 2048u);
              const uint32_t kstep = (2u * 2048u) >> 4;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                mma_ts_raw(td, ta + (uint32_t)ks * 8u, b_lo0 + (uint32_t)ks * kstep, b_hi, idesc, ks ? 1u : 0u);
              mma_commit_raw(&bars[D7_DREADY + s]);
            }
            __syncwarp();
            BH_TIMING_END(t_is)
          }
        }
      }
      if (lane == 0) { BH_TIMING_STORE(status, 20, t_wa) BH_TIMING_STORE(status, 22, t_is) }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: (32-column group, TMEM lane quadrant), all 16 on one tile at a time =====================
    const int cgrp = warp >> 2, qd = warp & 3, row = qd * 32 + lane;
    const int half = cgrp >> 1, cc = cgrp & 1;
    const uint32_t t_row = tbase + ((uint32_t)(qd * 32) << 16);
    uint32_t d_phase[2] = {0u, 0u};
    float db4 = 0.f;
    bool ok = true;
    const size_t act_fs = tc_acts_bytes_per_frame(v.n_pad, 1);
    float dout_next[2] = {0.f, 0.f};
    uint32_t mk_next[2][4];
    auto load_inputs = [&](int r) {
      const int T0n = sched_get(sched, r, ab);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int T = T0n + s;
        if (T0n < 0 || T >= NT) continue;
        const int b = T / tiles_per_frame, tile = T - b * tiles_per_frame;
        dout_next[s] = dout_all[(size_t)b * v.n_pad + tile * 128 + row];
        const uint8_t* mbase = acts + (size_t)b * act_fs + tc_mask_off(v.n_pad, 1);
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
    uint32_t q = 0;                                      // image sequence number (same order as the store thread's)
    BH_TIMING_T0 BH_TIMING_DECL(t_top) BH_TIMING_DECL(t_gf) BH_TIMING_DECL(t_wd) BH_TIMING_DECL(t_ep)
#ifdef BH_TC_TIMING
    const long long t_loop0 = clock64();
#endif
    // one cotangent image: packed words d[16] of this thread's 32 columns -> staging buffer of image q (+ TMEM A operand)
    auto emit = [&](const uint32_t (&d)[16], int s, bool to_tmem, bool to_stage) -> bool {
      const uint32_t b = q % kStg;
      if (to_stage) {
        BH_TIMING_BEGIN
        bool good = wait(&bars[D7_SFREE + b], ((q / kStg) & 1u) ^ 1u, ab);
        BH_TIMING_END(t_gf)
        if (!good) return false;
        uint8_t* img = stg + b * TC_SIMG_BYTES;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq)
          *reinterpret_cast<uint4*>(img + sample_img_off(row, cgrp * 4 + gq)) = make_uint4(d[4 * gq], d[4 * gq + 1], d[4 * gq + 2], d[4 * gq + 3]);
      }
      if (to_tmem) {
        tmem_st16(t_row + (uint32_t)s * 256u + 128u + (uint32_t)(cgrp * 16), d);
        tmem_wait_st();
        tc_fence_before_sync();
      }
      if (to_stage) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (to_stage) mbar_arrive(&bars[D7_SFULL + b]);
        if (to_tmem) mbar_arrive(&bars[D7_AREADY + s]);
      }
      if (to_stage) ++q;
      return true;
    };
    int r_done = 0;
    for (int r = 0; ok; ++r) {
      const int T0 = sched_get(sched, r, ab);
      if (T0 < 0) break;
      r_done = r + 1;
      const bool has[2] = {true, T0 + 1 < NT};
      float dout[2];
      uint32_t mk[2][4];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        dout[s] = dout_next[s];
#pragma unroll
        for (int l = 0; l < 4; ++l) mk[s][l] = mk_next[s][l];
      }
      load_inputs(r + 1);
      // ---- top of both tiles: delta_3[j] = dout * W4[j] * (h3[j] > 0), straight into the TMEM operand of the first product
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (!has[s] || !ok) continue;
        BH_TIMING_BEGIN
        if (cgrp == 0) db4 += dout[s];
        uint32_t d[16], dl[16];
        const uint32_t mw = mk[s][3];
        const float douts = dout[s] * gscale;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const float* w4 = w4s + (cgrp * 4 + gq) * 8;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            delta_pack<1>(tc_mask_expand(mw, 4 * gq + jj), douts * w4[2 * jj], douts * w4[2 * jj + 1], d[4 * gq + jj], dl[4 * gq + jj]);
        }
        ok = emit(d, s, true, false);
        BH_TIMING_END(t_top)
      }
      if (!ok) break;
      // ---- the chain, alternating between the slots: D = delta_l * W_l^T  ->  delta_{l-1}
      for (int l = 3; l >= 1 && ok; --l) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (!has[s] || !ok) continue;
          const uint32_t mw = mk[s][l - 1];
          BH_TIMING_BEGIN
          ok = wait(&bars[D7_DREADY + s], d_phase[s], ab);
          BH_TIMING_END(t_wd)
          if (!ok) break;
          d_phase[s] ^= 1u;
          tc_fence_after_sync();
          BH_TIMING_BEGIN
          uint32_t raw[32];
          tmem_ld32(t_row + (uint32_t)s * 256u + (uint32_t)(cgrp * 32), raw);
          tmem_wait_ld();
          uint32_t d[16], dl[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj)
            delta_pack<1>(tc_mask_expand(mw, jj), __uint_as_float(raw[2 * jj]), __uint_as_float(raw[2 * jj + 1]), d[jj], dl[jj]);
          ok = emit(d, s, l > 1, true);
          BH_TIMING_END(t_ep)
        }
      }
    }
#ifdef BH_TC_TIMING
    if (tid == 0) {
      long long t_loop = clock64() - t_loop0;
      BH_TIMING_STORE(status, 24, t_top) BH_TIMING_STORE(status, 26, t_gf) BH_TIMING_STORE(status, 30, t_wd)
      BH_TIMING_STORE(status, 32, t_ep) BH_TIMING_STORE(status, 34, t_loop)
      atomicMax((unsigned long long*)(status + 50), (unsigned long long)t_loop);
      {   // per-pair record: loop cycles / 64, rounds done, SM ids of the two CTAs (the partner fills its own)
        uint32_t smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
        uint32_t* rec = (uint32_t*)((uint8_t*)status + TC_WS_CONST) + 710 + (blockIdx.x / 2) * 2;
        rec[0] = (uint32_t)(t_loop >> 6);
        atomicOr(rec + 1, smid | ((uint32_t)r_done << 20));
      }
    }
#endif
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) db4 += __shfl_xor_sync(0xffffffffu, db4, o);
    if (lane == 0 && db4 != 0.f) atomicAdd(d_params + OFF_B4, db4);
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();          // no CTA of the pair leaves while the other may still arrive on its barriers
  if (warp == kDMmaWarp) tmem_dealloc(tbase, 512);
  if (tid == 0 && *abort_s) raise_flag(status, 1);
}

// =====================================================================================================
// wgrad role, v7 (one plane, one dgrad partner)
// =====================================================================================================
__device__ __forceinline__ void
wgrad_role_v7(uint8_t* smem, const PairLink link, int n_pad, int Bt,
              const uint8_t* __restrict__ ws, const float* __restrict__ dout_all,
              const uint8_t* __restrict__ acts, float* __restrict__ d_params, int* __restrict__ status) {
  float ginv = 1.f;
  const float gscale = tc_grad_scale(((const uint32_t*)(ws + TC_WS_CONST))[TC_C_DOUTMAX], ginv);
  float* w4s = (float*)(smem + W7_W4);
  uint64_t* bars = (uint64_t*)(smem + W7_SM_BARS);
  const uint8_t* sched = smem + W7_SM_BARS + SCHED7_OFF;
  uint32_t* tmem_base_s = (uint32_t*)(bars + W7_NBARS);
  int* abort_s = (int*)(tmem_base_s + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_frame = n_pad / 128;
  const int NT = Bt * tiles_per_frame;
  Abort ab{abort_s};
  const size_t act_fs = tc_acts_bytes_per_frame(n_pad, 1);
  const size_t lstride = (size_t)n_pad * 256u, pstride = (size_t)n_pad * 1024u;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&bars[W7_XFULL + s], 1); mbar_init(&bars[W7_XEMPTY + s], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&bars[W7_YFULL + s], 1); mbar_init(&bars[W7_YEMPTY + s], 1); }
    mbar_init(&bars[W7_DONE], 1);
    for (int s = 0; s < kRing7; ++s) mbar_init(&bars[W7_GFULL + s], 1);            // the partner's store thread
    for (int k = 0; k < 2 * kSched; ++k) reinterpret_cast<volatile uint32_t*>(smem + W7_SM_BARS + SCHED7_OFF)[k] = 0u;
    abort_s[0] = 0; abort_s[1] = 0;
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(tmem_base_s, 512);
  if (tid < 128) w4s[tid] = ((const float*)(ws + TC_WS_CONST))[TC_C_W4 + tid];
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_base_s;
  WgSeq q_first;
  const bool has_work = wg_tile7(0, sched, NT, q_first, ab) == 1;

  if (warp == 5) {
    // ===================== X producer: saved activations (HBM), independent of the partner's progress =====================
    if (lane == 0) {
      bool ok = true;
      uint32_t tcount = 0;                               // tiles processed: every X slot is used once per tile
      for (int it = 0; ok; ++it) {
        WgSeq sq;
        const int tv = wg_tile7(it, sched, NT, sq, ab);
        if (tv == 0) break;
        if (tv == 2) continue;
        const int b = sq.T / tiles_per_frame, tile = sq.T - b * tiles_per_frame;
        const uint8_t* act_tile = acts + (size_t)b * act_fs + (size_t)tile * TC_SIMG_BYTES;
        const uint8_t* feat_tile = acts + (size_t)b * act_fs + pstride + (size_t)tile * TC_FIMG_BYTES;
        for (int i = 0; i < 3 && ok; ++i) {              // item i: h2 + feat | h1 + feat[16..] | h0 + feat[16..]
          ok = wait(&bars[W7_XEMPTY + i], (tcount & 1u) ^ 1u, ab);
          if (!ok) break;
          uint8_t* dst = smem + w7_xslot(i);
          uint64_t* full = &bars[W7_XFULL + i];
          const uint32_t fbytes = i == 0 ? TC_FIMG_BYTES : TC_FIMG_BYTES / 2u;
          mbar_expect_tx(full, TC_SIMG_BYTES + fbytes);
          bulk_g2s(dst, act_tile + (size_t)(2 - i) * lstride, TC_SIMG_BYTES, full);
          if (i == 0) bulk_g2s(dst + TC_SIMG_BYTES, feat_tile, TC_FIMG_BYTES, full);
          else bulk_g2s(dst + TC_SIMG_BYTES, feat_tile + TC_FIMG_BYTES / 2u, TC_FIMG_BYTES / 2u, full);
        }
        ++tcount;
      }
    }
    __syncwarp();
  } else if (warp == 6) {
    // ===================== Y producer: cotangent images + aux of a tile, once the partner has handed it over =====================
    if (lane == 0) {
      bool ok = true;
      uint32_t tcount = 0;
      BH_TIMING_T0 BH_TIMING_DECL(t_gfl) BH_TIMING_DECL(t_em)
      for (int it = 0; ok; ++it) {
        WgSeq sq;
        const int tv = wg_tile7(it, sched, NT, sq, ab);
        if (tv == 0) break;
        if (tv == 2) continue;
        const int b = sq.T / tiles_per_frame, tile = sq.T - b * tiles_per_frame;
        const uint8_t* feat_tile = acts + (size_t)b * act_fs + pstride + (size_t)tile * TC_FIMG_BYTES;
        const uint8_t* set = link.ring + (size_t)sq.rd * TSET7;
        BH_TIMING_BEGIN
        ok = wait_partner(&bars[W7_GFULL + sq.rd], sq.ru, ab);
        BH_TIMING_END(t_gfl)
        if (!ok) break;
        fence_proxy_async_global();
        for (int j = 1; j < 4 && ok; ++j) {              // delta_2, delta_1, delta_0 (+ the feature image); delta_3 is rebuilt here
          const uint32_t ys = (uint32_t)j & 1u;
          // the slot's previous occupant: slot 1 alternates delta_2 / delta_0, slot 0 holds delta_1 of every tile
          BH_TIMING_BEGIN
          ok = j == 3 ? wait(&bars[W7_YEMPTY + 1], tcount & 1u, ab)
                      : wait(&bars[W7_YEMPTY + (j == 1 ? 3 : 2)], (tcount & 1u) ^ 1u, ab);
          BH_TIMING_END(t_em)
          if (!ok) break;
          uint8_t* dst = smem + (ys ? W7_Y1 : W7_Y0);
          uint64_t* full = &bars[W7_YFULL + j];
          mbar_expect_tx(full, TC_SIMG_BYTES + (j == 3 ? TC_FIMG_BYTES : 0u));
          bulk_g2s(dst, set + (size_t)(3 - j) * TC_SIMG_BYTES, TC_SIMG_BYTES, full);
          if (j == 3) bulk_g2s(dst + TC_SIMG_BYTES, feat_tile, TC_FIMG_BYTES, full);
        }
        if (!ok) break;
        ++tcount;
      }
      BH_TIMING_STORE_B(status, 38, t_gfl, 1) BH_TIMING_STORE_B(status, 40, t_em, 1)
    }
    __syncwarp();
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    {
      const uint32_t id160 = make_idesc_f16(128, 160, 1, 1), id144 = make_idesc_f16(128, 144, 1, 1), id32 = make_idesc_f16(128, 32, 1, 1);
      bool ok = true;
      uint32_t tcount = 0;
      BH_TIMING_T0 BH_TIMING_DECL(t_fu) BH_TIMING_DECL(t_wi) BH_TIMING_DECL(t_f0) BH_TIMING_DECL(t_fx)
#ifdef BH_TC_TIMING
      const long long t_w0 = clock64();
#endif
      for (int it = 0; ok; ++it) {
        WgSeq sq;
        const int tv = wg_tile7(it, sched, NT, sq, ab);
        if (tv == 0) break;
        if (tv == 2) continue;
        for (int j = 0; j < 4 && ok; ++j) {              // job j: delta_{3-j} x [in | feat]
          const uint32_t ys = (uint32_t)j & 1u;
          BH_TIMING_BEGIN
          ok = wait(&bars[W7_YFULL + j], tcount & 1u, ab);
          if (j == 0) { BH_TIMING_END(t_f0) } else { BH_TIMING_END(t_fu) }
          BH_TIMING_BEGIN
          if (ok && j < 3) ok = wait(&bars[W7_XFULL + j], tcount & 1u, ab);
          BH_TIMING_END(t_fx)
          if (!ok) break;
          tc_fence_after_sync();
          BH_TIMING_BEGIN
          const uint32_t A = smem_u32(smem + (j == 0 ? W7_X3 : (ys ? W7_Y1 : W7_Y0)));
          const uint32_t B = j < 3 ? smem_u32(smem + w7_xslot(j)) : A + TC_SIMG_BYTES;     // job 3: the feature image behind delta_0
          const uint32_t acc_col = j == 0 ? ACC_W3 : j == 1 ? ACC_W2 : j == 2 ? ACC_W1 : ACC_W0F;
          const uint32_t idesc = j == 0 ? id160 : j < 3 ? id144 : id32;
          if (elect_one()) {
            const uint32_t hi = desc_hi(TC_SIMG_CS), kstep = (2u * TC_IMG_RS) >> 4;
            const uint32_t a_lo = desc_lo(A, TC_IMG_RS), b_lo = desc_lo(B, TC_IMG_RS);
#pragma unroll
            for (uint32_t ks = 0; ks < 8; ++ks)
              mma_ss_raw(tbase + acc_col, a_lo + ks * kstep, hi, b_lo + ks * kstep, hi, idesc, (tcount | ks) ? 1u : 0u);
            mma_commit_raw(&bars[W7_YEMPTY + j]);
            if (j < 3) mma_commit_raw(&bars[W7_XEMPTY + j]);
          }
          __syncwarp();
          BH_TIMING_END(t_wi)
        }
        if (!ok) break;
        // every cotangent image of the set has been pulled out of the ring (each Y full phase was observed)
        if (lane == 0) mbar_arrive_remote(link.peer_bars + sq.rd * 8u);
        __syncwarp();
        ++tcount;
      }
      if (elect_one()) mma_commit_raw(&bars[W7_DONE]);
      __syncwarp();
#ifdef BH_TC_TIMING
      if (lane == 0) {
        long long t_wl = clock64() - t_w0;
        BH_TIMING_STORE_B(status, 44, t_fu, 1) BH_TIMING_STORE_B(status, 46, t_wi, 1) BH_TIMING_STORE_B(status, 48, t_wl, 1)
        BH_TIMING_STORE_B(status, 42, t_f0, 1) BH_TIMING_STORE_B(status, 36, t_fx, 1)
      }
#endif
    }
    __syncwarp();
  } else if (warp >= kGenWarp0 && warp < kGenWarp0 + kGenWarps) {
    // ===================== delta_3 of every tile, rebuilt locally: dout * W4[j] * (h3[j] > 0) -> its own buffer =====================
    // (bit-identical to what the partner's epilogue feeds its first product: same inputs, same delta_pack), and dW4 = h3 x dout
    // on the CUDA cores with h3 read straight from global memory while the image is being built.
    const int gt = tid - kGenWarp0 * 32, row = gt & 127, hf = gt >> 7;          // image: one sample row, 64 of its 128 columns
    const int gw = warp - kGenWarp0, rp = lane & 7, cg = (gw & 3) * 4 + (lane >> 3), rhalf = gw >> 2;   // dW4: 8 columns, 8 of 64 rows
    bool ok = true;
    uint32_t tcount = 0;
    float w4acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) w4acc[c] = 0.f;
    // the image's inputs (this row's d loss/d o and layer-3 mask words) are fetched one tile ahead
    float dout_nx = 0.f;
    uint2 mw_nx = make_uint2(0u, 0u);
    auto fetch = [&](int it) {
      for (;; ++it) {
        WgSeq sq;
        const int tv = wg_tile7(it, sched, NT, sq, ab);
        if (tv == 0) return;
        if (tv == 2) continue;
        const int b = sq.T / tiles_per_frame, tile = sq.T - b * tiles_per_frame;
        dout_nx = dout_all[(size_t)b * n_pad + tile * 128 + row];
        mw_nx = *reinterpret_cast<const uint2*>(acts + (size_t)b * act_fs + tc_mask_off(n_pad, 1) + tc_mask_word_off(tile, 3, hf, row));
        return;
      }
    };
    fetch(0);
    BH_TIMING_T0 BH_TIMING_DECL(t_gw)
    for (int it = 0; ok; ++it) {
      WgSeq sq;
      const int tv = wg_tile7(it, sched, NT, sq, ab);
      if (tv == 0) break;
      if (tv == 2) continue;
      const int b = sq.T / tiles_per_frame, tile = sq.T - b * tiles_per_frame;
      const float douts = dout_nx * gscale;
      const uint2 mw2 = mw_nx;
      fetch(it + 1);
      const float* dtile = dout_all + (size_t)b * n_pad + tile * 128;
      const uint8_t* h3 = acts + (size_t)b * act_fs + (size_t)tile * TC_SIMG_BYTES + 3u * lstride + cg * TC_SIMG_CS + rhalf * 8u * TC_IMG_RS + rp * 16;
      uint4 hv[8];
      float dr[8];
#pragma unroll
      for (int rg = 0; rg < 8; ++rg) {
#ifdef BH_EXP_NODW4
        hv[rg] = make_uint4(0u, 0u, 0u, 0u); dr[rg] = 0.f; (void)h3;
#else
        hv[rg] = *reinterpret_cast<const uint4*>(h3 + rg * TC_IMG_RS);
        dr[rg] = dtile[rhalf * 64 + rg * 8 + rp];
#endif
      }
      BH_TIMING_BEGIN
      ok = wait(&bars[W7_YEMPTY + 0], (tcount & 1u) ^ 1u, ab);           // the previous tile's delta_3 has been consumed
      BH_TIMING_END(t_gw)
      if (!ok) break;
      uint8_t* img = smem + W7_X3;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const uint32_t mw = cc ? mw2.y : mw2.x;
        const int cgrp = hf * 2 + cc;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const float* w4 = w4s + (cgrp * 4 + gq) * 8;
          uint32_t d[4], dl[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            delta_pack<1>(tc_mask_expand(mw, 4 * gq + jj), douts * w4[2 * jj], douts * w4[2 * jj + 1], d[jj], dl[jj]);
          *reinterpret_cast<uint4*>(img + sample_img_off(row, cgrp * 4 + gq)) = make_uint4(d[0], d[1], d[2], d[3]);
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, %0;" ::"n"(kGenWarps * 32) : "memory");
      if (gt == 0) mbar_arrive(&bars[W7_YFULL + 0]);
#pragma unroll
      for (int rg = 0; rg < 8; ++rg) {
        const float dout = dr[rg];
        w4acc[0] = fmaf(f16_lo(hv[rg].x), dout, w4acc[0]); w4acc[1] = fmaf(f16_hi(hv[rg].x), dout, w4acc[1]);
        w4acc[2] = fmaf(f16_lo(hv[rg].y), dout, w4acc[2]); w4acc[3] = fmaf(f16_hi(hv[rg].y), dout, w4acc[3]);
        w4acc[4] = fmaf(f16_lo(hv[rg].z), dout, w4acc[4]); w4acc[5] = fmaf(f16_hi(hv[rg].z), dout, w4acc[5]);
        w4acc[6] = fmaf(f16_lo(hv[rg].w), dout, w4acc[6]); w4acc[7] = fmaf(f16_hi(hv[rg].w), dout, w4acc[7]);
      }
      ++tcount;
    }
    if (gt == 0) { BH_TIMING_STORE_B(status, 28, t_gw, 1) }
    if (ok) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float vsum = w4acc[c];
        vsum += __shfl_xor_sync(0xffffffffu, vsum, 1);
        vsum += __shfl_xor_sync(0xffffffffu, vsum, 2);
        vsum += __shfl_xor_sync(0xffffffffu, vsum, 4);
        vsum *= TC_RZ_UNBIAS;
        if (rp == 0 && vsum != 0.f) atomicAdd(d_params + OFF_W4 + cg * 8 + c, vsum);
        if (rp == 0 && !(fabsf(vsum) <= 3.0e38f)) abort_s[1] = 1;
      }
    }
  } else if (warp < 4 && has_work) {
    // ===================== final flush =====================
    bool ok = true;
    // ---- final flush: TMEM accumulators -> d_params ----
    ok = ok && wait(&bars[W7_DONE], 0, ab);
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
          const bool from_h = (c < ACC_W3F || (c >= ACC_W2 && c < ACC_B2) || (c >= ACC_W1 && c < ACC_B1));
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
          if (dst >= 0 && !(fabsf(val) <= 3.0e38f)) abort_s[1] = 1;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 4) tmem_dealloc(tbase, 512);
  if (tid == 0 && *abort_s) raise_flag(status, 2);
  if (tid == 0 && abort_s[1]) raise_flag(status, 4);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDThreads, 1)
tc_bwd_fused7_kernel(PackedView v, const uint8_t* __restrict__ ws, const float* __restrict__ dout_all, int Bt,
                     const uint8_t* __restrict__ acts, uint8_t* __restrict__ ring, float* __restrict__ d_params,
                     int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t rank = cluster_ctarank();
  const int cl = (int)(blockIdx.x / 2), ncl = (int)(gridDim.x / 2);
#ifdef BH_TC_TIMING
  const long long t_kernel0 = clock64();
#endif
  PairLink link;
  link.ring = ring + (size_t)cl * kRing7 * TSET7;
  if (rank == 0) {
    link.peer_bars = mapa_u32(smem_u32(smem + W7_SM_BARS + W7_GFULL * 8), 1u);
    dgrad_role_v7(smem, mapa_u32(smem_u32(smem + W7_SM_BARS + SCHED7_OFF), 1u), link, v, ws, dout_all, Bt, acts, d_params, status);
  } else {
    link.peer_bars = mapa_u32(smem_u32(smem + D7_SM_BARS + D7_GFREE * 8), 0u);
    wgrad_role_v7(smem, link, v.n_pad, Bt, ws, dout_all, acts, d_params, status);
  }
#ifdef BH_TC_TIMING
  if (threadIdx.x == 0 && rank == 1) {
    uint32_t smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
    atomicOr((uint32_t*)((uint8_t*)status + TC_WS_CONST) + 710 + cl * 2 + 1, smid << 10);
  }
  if (threadIdx.x == 0) atomicMax((unsigned long long*)(status + 52), (unsigned long long)(clock64() - t_kernel0));
#endif
}

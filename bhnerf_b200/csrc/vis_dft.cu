// Separable visibility head on the tensor cores (opt-in; BASELINE.json north_star: "pixel-to-visibility DFT as a
// tensor-core complex GEMM").  loss_fn_eht multiplies every frame's image with an explicit complex matrix A (nvis x npix,
// bhnerf/network.py:542-544) that ehtim builds as  A[k,(i,j)] = pulse_k * exp(-2 pi i (u_k x_i + v_k y_j))  on the regular
// pixel grid.  An opaque A makes the head a batched GEMV bound by the HBM stream of A (heads.cu: bhnerf_vis_head, 8 V P bytes
// per frame per pass).  When the caller hands over (u, v) instead, the kernel factorises:
//   forward   T[k,i] = sum_j E_v[k,j] I[i,j]      -- GEMM  (M = k, N = i, K = j)   E_v = exp(-2 pi i v_k y_j) = C - iS
//             vis_k  = pulse_k sum_i E_u[k,i] T[k,i]                                 (row-wise, in the epilogue)
//   backward  G[k,i] = conj(pulse_k E_u[k,i]) d_vis_k;   d_I[i,j] = sum_k Re(G[k,i] conj(E_v[k,j]))   -- GEMM  (M = i, N = j, K = k)
// The DFT factors are generated in shared memory from (u, v) -- nothing of size V x P is ever read: 1.5 KB of (u, v) per frame
// instead of 25 MB of A.  tcgen05.mma kind::f16, fp32 accumulate in TMEM; the images / cotangents go in as fp16 hi + lo
// planes (22 bits) and so do the unit-modulus factors (with one fp16 plane their 2^-12 error showed up at 1.5e-4 of the
// visibilities of a 16 x 16 image with a few dominant pixels); three products per accumulator, hi*hi + hi*lo + lo*hi.
#include <cuda_fp16.h>
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int kDftThreads = 128;         // 4 warps: one TMEM lane quadrant each; thread = one GEMM row
// shared memory: operand images of 128 rows x 128 columns fp16 (32 KB each, the canonical layout of tc_common.cuh)
constexpr uint32_t DF_IMG = TC_SIMG_BYTES;
constexpr uint32_t DF_SM_A0 = 0, DF_SM_A1 = DF_IMG, DF_SM_B0 = 2 * DF_IMG, DF_SM_B1 = 3 * DF_IMG, DF_SM_A2 = 4 * DF_IMG,
                   DF_SM_A3 = 5 * DF_IMG;
constexpr uint32_t DF_SM_BARS = 6 * DF_IMG;
constexpr uint32_t DF_SM_TOTAL = DF_SM_BARS + 64;

__device__ __forceinline__ void store_h(uint8_t* img, int row, int col, float v) {
  *reinterpret_cast<__half*>(img + sample_img_off(row, col >> 3) + (uint32_t)(col & 7) * 2u) = __float2half_rn(v);
}
// fp16 hi / lo planes of v (22 significant bits) into two images
__device__ __forceinline__ void store_h2(uint8_t* img_hi, uint8_t* img_lo, int row, int col, float v) {
  const __half h = __float2half_rn(v);
  const uint32_t off = sample_img_off(row, col >> 3) + (uint32_t)(col & 7) * 2u;
  *reinterpret_cast<__half*>(img_hi + off) = h;
  *reinterpret_cast<__half*>(img_lo + off) = __float2half_rn(v - __half2float(h));
}

struct DftGrid { float x0, dx, y0, dy; int NA, NB; };

// ---------------- forward: one CTA per (128 visibilities, frame) ----------------
// A operands (K-major, rows = k): C = cos(2 pi v_k y_j), S = sin(2 pi v_k y_j);  B operands (K-major, rows = i): I hi, I lo.
// D_re[k,i] = sum_j C I;  D_ms[k,i] = sum_j S I  (T = D_re - i D_ms).
__global__ void __launch_bounds__(kDftThreads, 1)
vis_dft_fwd_kernel(const float2* __restrict__ uv, const float2* __restrict__ pulse, const float* __restrict__ images, int V,
                   DftGrid g, float2* __restrict__ vis, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = (uint64_t*)(smem + DF_SM_BARS);
  uint32_t* tmem_base_s = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y, k0 = blockIdx.x * 128, k = k0 + tid;
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(tmem_base_s, 256);
  const float2 uvk = k < V ? uv[(size_t)b * V + k] : make_float2(0.f, 0.f);
  const float* img = images + (size_t)b * g.NA * g.NB;
  // this thread's row of C and S (row = its visibility), and row i = tid of the image planes
  for (int j = 0; j < 128; ++j) {
    float sn = 0.f, cs = 0.f;
    if (j < g.NB) sincospif(2.f * uvk.y * (g.y0 + j * g.dy), &sn, &cs);
    store_h2(smem + DF_SM_A0, smem + DF_SM_A2, tid, j, j < g.NB ? cs : 0.f);       // C hi | C lo
    store_h2(smem + DF_SM_A1, smem + DF_SM_A3, tid, j, j < g.NB ? sn : 0.f);       // S hi | S lo
  }
  for (int idx = tid; idx < 128 * 128; idx += kDftThreads) {       // coalesced over j
    const int i = idx >> 7, j = idx & 127;
    const float x = (i < g.NA && j < g.NB) ? img[(size_t)i * g.NB + j] : 0.f;
    const __half h = __float2half_rn(x);
    *reinterpret_cast<__half*>(smem + DF_SM_B0 + sample_img_off(i, j >> 3) + (uint32_t)(j & 7) * 2u) = h;
    *reinterpret_cast<__half*>(smem + DF_SM_B1 + sample_img_off(i, j >> 3) + (uint32_t)(j & 7) * 2u) = __float2half_rn(x - __half2float(h));
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_base_s;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, 128, 0, 0);
      const uint32_t hi = desc_hi(TC_IMG_RS), kstep = (2u * TC_SIMG_CS) >> 4;
      const uint32_t a0 = desc_lo(smem_u32(smem + DF_SM_A0), TC_SIMG_CS), a1 = desc_lo(smem_u32(smem + DF_SM_A1), TC_SIMG_CS);
      const uint32_t a2 = desc_lo(smem_u32(smem + DF_SM_A2), TC_SIMG_CS), a3 = desc_lo(smem_u32(smem + DF_SM_A3), TC_SIMG_CS);
      const uint32_t b0 = desc_lo(smem_u32(smem + DF_SM_B0), TC_SIMG_CS), b1 = desc_lo(smem_u32(smem + DF_SM_B1), TC_SIMG_CS);
#pragma unroll
      for (uint32_t ks = 0; ks < 8; ++ks) {      // hi*hi + hi*lo + lo*hi per accumulator
        mma_ss_raw(tbase, a0 + ks * kstep, hi, b0 + ks * kstep, hi, idesc, ks ? 1u : 0u);
        mma_ss_raw(tbase, a0 + ks * kstep, hi, b1 + ks * kstep, hi, idesc, 1u);
        mma_ss_raw(tbase, a2 + ks * kstep, hi, b0 + ks * kstep, hi, idesc, 1u);
        mma_ss_raw(tbase + 128u, a1 + ks * kstep, hi, b0 + ks * kstep, hi, idesc, ks ? 1u : 0u);
        mma_ss_raw(tbase + 128u, a1 + ks * kstep, hi, b1 + ks * kstep, hi, idesc, 1u);
        mma_ss_raw(tbase + 128u, a3 + ks * kstep, hi, b0 + ks * kstep, hi, idesc, 1u);
      }
      mma_commit_raw(bar);
    }
    __syncwarp();
  }
  if (!mbar_wait(bar, 0)) { if (tid == 0) atomicExch(err, 1); }
  tc_fence_after_sync();
  // epilogue: vis_k = pulse_k sum_i (Cu - i Su)(D_re - i D_ms),  E_u[k,i] = exp(-2 pi i u_k x_i) by rotation from x_0
  const uint32_t t_lane = tbase + ((uint32_t)(warp * 32) << 16);
  float re = 0.f, im = 0.f;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t dr[32], ds[32];
    tmem_ld32(t_lane + (uint32_t)c0, dr);
    tmem_ld32(t_lane + 128u + (uint32_t)c0, ds);
    tmem_wait_ld();
#pragma unroll 8
    for (int jj = 0; jj < 32; ++jj) {
      const int i = c0 + jj;
      float su, cu;
      sincospif(2.f * uvk.x * (g.x0 + i * g.dx), &su, &cu);
      const float tr = __uint_as_float(dr[jj]), ts = __uint_as_float(ds[jj]);      // T = tr - i ts
      // (cu - i su)(tr - i ts) = (cu tr - su ts) - i (cu ts + su tr)
      re += cu * tr - su * ts;
      im -= cu * ts + su * tr;
    }
  }
  if (k < V) {
    const float2 p = pulse ? pulse[(size_t)b * V + k] : make_float2(1.f, 0.f);
    vis[(size_t)b * V + k] = make_float2(p.x * re - p.y * im, p.x * im + p.y * re);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

// ---------------- backward: one CTA per (128 image rows i, frame); loops over the visibilities in tiles of 128 ----------------
// w_k = conj(pulse_k) d_vis_k * scale;  G[k,i] = w_k exp(+2 pi i u_k x_i) = Gr + i Gi;  d_I[i,j] = sum_k Gr C[k,j] - Gi S[k,j]
// A operands (K-major, rows = i, cols = k): Gr / -Gi as fp16 hi + lo;  B operands (MN-major, rows = k, cols = j): C, S.
// The images are refilled between the two chains (Gr C, then -Gi S) of every visibility tile.
__global__ void __launch_bounds__(kDftThreads, 1)
vis_dft_bwd_kernel(const float2* __restrict__ uv, const float2* __restrict__ pulse, const float2* __restrict__ dvis, int V,
                   DftGrid g, const uint32_t* __restrict__ wmax_bits, float* __restrict__ d_images, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = (uint64_t*)(smem + DF_SM_BARS);
  uint32_t* tmem_base_s = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y, i0 = blockIdx.x * 128, i = i0 + tid;
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(tmem_base_s, 128);
  float winv = 1.f;
  const float wscale = tc_grad_scale(*wmax_bits, winv) * 64.f;        // max |w| -> [128, 256): fp16-safe for any sigma
  winv *= (1.f / 64.f);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_base_s;
  const float xi = g.x0 + i * g.dx;
  uint32_t phase = 0, started = 0;
  bool ok = true;
  for (int k0 = 0; k0 < V && ok; k0 += 128) {
    for (int chain = 0; chain < 2 && ok; ++chain) {                   // 0: Gr x C     1: (-Gi) x S
      // A: row = this thread's image row i, columns = the 128 visibilities of the tile
      for (int kk = 0; kk < 128; ++kk) {
        const int k = k0 + kk;
        float val = 0.f;
        if (k < V && i < g.NA) {
          const float2 u = uv[(size_t)b * V + k], dv = dvis[(size_t)b * V + k];
          const float2 p = pulse ? pulse[(size_t)b * V + k] : make_float2(1.f, 0.f);
          const float wr = (p.x * dv.x + p.y * dv.y) * wscale, wi = (p.x * dv.y - p.y * dv.x) * wscale;   // conj(p) dv
          float su, cu;
          sincospif(2.f * u.x * xi, &su, &cu);                       // exp(+2 pi i u x) = cu + i su
          val = chain == 0 ? (wr * cu - wi * su) : -(wr * su + wi * cu);
        }
        const __half h = __float2half_rn(val);
        *reinterpret_cast<__half*>(smem + DF_SM_A0 + sample_img_off(tid, kk >> 3) + (uint32_t)(kk & 7) * 2u) = h;
        *reinterpret_cast<__half*>(smem + DF_SM_A1 + sample_img_off(tid, kk >> 3) + (uint32_t)(kk & 7) * 2u) = __float2half_rn(val - __half2float(h));
      }
      // B: row = visibility k0 + tid, columns = j
      {
        const int k = k0 + tid;
        const float vk = k < V ? uv[(size_t)b * V + k].y : 0.f;
        for (int j = 0; j < 128; ++j) {
          float sn = 0.f, cs = 0.f;
          if (k < V && j < g.NB) sincospif(2.f * vk * (g.y0 + j * g.dy), &sn, &cs);
          store_h2(smem + DF_SM_B0, smem + DF_SM_B1, tid, j, chain == 0 ? cs : sn);
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (warp == 0) {
        if (elect_one()) {
          const uint32_t idesc = make_idesc_f16(128, 128, 0, 1);       // A K-major, B MN-major
          const uint32_t a_hi = desc_hi(TC_IMG_RS), akstep = (2u * TC_SIMG_CS) >> 4;
          const uint32_t b_hi = desc_hi(TC_SIMG_CS), bkstep = (2u * TC_IMG_RS) >> 4;
          const uint32_t a0 = desc_lo(smem_u32(smem + DF_SM_A0), TC_SIMG_CS), a1 = desc_lo(smem_u32(smem + DF_SM_A1), TC_SIMG_CS);
          const uint32_t b0 = desc_lo(smem_u32(smem + DF_SM_B0), TC_IMG_RS), b1 = desc_lo(smem_u32(smem + DF_SM_B1), TC_IMG_RS);
#pragma unroll
          for (uint32_t ks = 0; ks < 8; ++ks) {
            mma_ss_raw(tbase, a0 + ks * akstep, a_hi, b0 + ks * bkstep, b_hi, idesc, (started | ks) ? 1u : 0u);
            mma_ss_raw(tbase, a1 + ks * akstep, a_hi, b0 + ks * bkstep, b_hi, idesc, 1u);
            mma_ss_raw(tbase, a0 + ks * akstep, a_hi, b1 + ks * bkstep, b_hi, idesc, 1u);
          }
          mma_commit_raw(bar);
        }
        __syncwarp();
      }
      started = 1;
      ok = mbar_wait(bar, phase);                                    // the operand images are free again
      phase ^= 1u;
      tc_fence_after_sync();
    }
  }
  if (!ok && tid == 0) atomicExch(err, 1);
  const uint32_t t_lane = tbase + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t d[32];
    tmem_ld32(t_lane + (uint32_t)c0, d);
    tmem_wait_ld();
    if (i < g.NA)
#pragma unroll 8
      for (int jj = 0; jj < 32; ++jj)
        if (c0 + jj < g.NB) d_images[((size_t)b * g.NA + i) * g.NB + c0 + jj] = __uint_as_float(d[jj]) * winv;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 128);
}

// max |conj(pulse) d_vis| of the launch (uint32 bits of a non-negative float order like the float)
__global__ void vis_dft_wmax_kernel(const float2* __restrict__ pulse, const float2* __restrict__ dvis, int n, uint32_t* __restrict__ out) {
  float mx = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float2 dv = dvis[i];
    const float2 p = pulse ? pulse[i] : make_float2(1.f, 0.f);
    mx = fmaxf(mx, fmaxf(fabsf(p.x * dv.x + p.y * dv.y), fabsf(p.x * dv.y - p.y * dv.x)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(out, __float_as_uint(mx));
}

}  // namespace

static int check_grid(int NA, int NB) {
  BH_REQUIRE(NA >= 1 && NA <= 128 * 65535 && NB >= 1 && NB <= 128, "vis_dft: the separable head handles up to 128 pixels along beta (NB), got %d", NB);
  return 0;
}

extern "C" int bhnerf_vis_dft_fwd(const float* uv, const float* pulse, const float* images, int32_t Bt, int32_t V, int32_t NA,
                                  int32_t NB, float x0, float dx, float y0, float dy, float* vis, int32_t* status_dev, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(uv && images && vis && status_dev && Bt > 0 && V > 0 && Bt <= 65535, "vis_dft_fwd: bad argument");
  if (int r = check_grid(NA, NB)) return r;
  BH_REQUIRE(NA <= 128, "vis_dft_fwd: up to 128 pixels along alpha (NA), got %d", NA);
  BH_CHECK_CUDA(cudaFuncSetAttribute(vis_dft_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SM_TOTAL));
  BhProfScope ps(BH_CAT_VIS, 1, st);
  DftGrid g{x0, dx, y0, dy, NA, NB};
  vis_dft_fwd_kernel<<<dim3((V + 127) / 128, Bt), kDftThreads, DF_SM_TOTAL, st>>>((const float2*)uv, (const float2*)pulse, images, V, g,
                                                                             (float2*)vis, status_dev);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bhnerf_vis_dft_bwd(const float* uv, const float* pulse, const float* d_vis, int32_t Bt, int32_t V, int32_t NA,
                                  int32_t NB, float x0, float dx, float y0, float dy, float* d_images, int32_t* status_dev,
                                  void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(uv && d_vis && d_images && status_dev && Bt > 0 && V > 0 && Bt <= 65535, "vis_dft_bwd: bad argument");
  if (int r = check_grid(NA, NB)) return r;
  BH_CHECK_CUDA(cudaFuncSetAttribute(vis_dft_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SM_TOTAL));
  uint32_t* wmax = (uint32_t*)(status_dev + 1);
  BH_CHECK_CUDA(cudaMemsetAsync(wmax, 0, sizeof(uint32_t), st));
  BhProfScope ps(BH_CAT_VIS, 2, st);
  int n = Bt * V, blocks = (n + 255) / 256; if (blocks > 296) blocks = 296;
  vis_dft_wmax_kernel<<<blocks, 256, 0, st>>>((const float2*)pulse, (const float2*)d_vis, n, wmax);
  DftGrid g{x0, dx, y0, dy, NA, NB};
  vis_dft_bwd_kernel<<<dim3((NA + 127) / 128, Bt), kDftThreads, DF_SM_TOTAL, st>>>((const float2*)uv, (const float2*)pulse,
                                                                              (const float2*)d_vis, V, g, wmax, d_images, status_dev);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

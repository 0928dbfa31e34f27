// Hardware probes behind the fused backward's design, run by scripts/run_dsmem_probe.py:
//   dsmem_bw_kernel   : bytes/clk one CTA can push into its cluster partner's shared memory
//                       mode 0: st.shared::cluster.v4 from `nthreads` threads (what the dgrad epilogue would do)
//                       mode 1: cp.async.bulk.shared::cluster.shared::cta in `chunk`-byte pieces (copy engine)
//   mixed_fmt_kernel  : tcgen05.mma kind::f16 with A = bf16 and B = fp16 in ONE instruction (is it legal / exact?)
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include "umma.cuh"

using namespace umma;

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { return mapa_u32(addr, rank); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(1024, 1)
dsmem_bw_kernel(int mode, int reps, int chunk, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];      // [0,64K) source (mode 1) | [64K,128K) destination window
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = tid; i < 128 * 1024 / 16; i += nthr) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, 1u, 2u, 3u);
  __syncthreads();
  cluster_sync_all();
  const uint32_t dst_local = smem_u32(smem + 64 * 1024);
  const uint32_t dst_remote = mapa(dst_local, rank ^ 1u);
  const uint32_t bar_remote = mapa(smem_u32(&bar), rank ^ 1u);
  long long t0 = clock64();
  if (rank == 0) {
    if (mode == 0) {
      for (int r = 0; r < reps; ++r) {
        // 32 KB per rep: every thread stores 16-byte pieces at consecutive addresses (a warp covers 512 contiguous bytes)
        for (int i = tid; i < 32 * 1024 / 16; i += nthr) {
          const uint32_t a = dst_remote + (uint32_t)(((r & 1) * 32 * 1024) + i * 16);
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r), "r"(i), "r"(tid), "r"(7) : "memory");
        }
      }
      asm volatile("fence.acq_rel.cluster;" ::: "memory");
    } else if (tid == 0) {
      // the destination CTA's barrier counts the bytes; this CTA only issues
      const int per_rep = 32 * 1024 / chunk;
      for (int r = 0; r < reps; ++r)
        for (int c = 0; c < per_rep; ++c) {
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           dst_remote + (uint32_t)((r & 1) * 32 * 1024 + c * chunk)),
                       "r"(smem_u32(smem) + (uint32_t)(c * chunk)), "r"(chunk), "r"(bar_remote)
                       : "memory");
        }
    }
  } else if (mode == 1 && tid == 0) {
    // receiver: one phase counts all the bytes (tx-count is 20 bits: the host keeps reps*32 KB <= 512 KB)
    mbar_expect_tx(&bar, (uint32_t)(reps * 32 * 1024));
    if (!mbar_wait(&bar, 0, 1u << 24)) out[2] = -1;
  }
  cluster_sync_all();
  long long t1 = clock64();
  if (tid == 0 && rank == 0) { out[0] = t1 - t0; out[1] = (long long)reps * 32 * 1024; }
  if (tid == 0 && rank == 1) { out[3] = ((volatile uint32_t*)(smem + 64 * 1024))[0]; }
}

extern "C" int dsmem_bw_run(int mode, int nthreads, int reps, int chunk, long long* out_dev, void* stream) {
  size_t smem = 128 * 1024;
  if (cudaFuncSetAttribute(dsmem_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
  dsmem_bw_kernel<<<2, nthreads, smem, (cudaStream_t)stream>>>(mode, reps, chunk, out_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

// Are the two ways out of an SM independent?  CTA 0 of every pair pushes `reps` x 32 KB into its partner's shared memory
// (flags & 1: cp.async.bulk smem -> dsmem) and/or the same amount into an L2-resident global region (flags & 2:
// cp.async.bulk smem -> global), from two different threads.  out[4*pair + 0/1] = cycles until the DSMEM bytes have
// landed / until the global copies have completed.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
egress2_kernel(int flags, int reps, uint8_t* __restrict__ gbuf, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];      // [0,64K) source | [64K,128K) destination window
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = tid; i < 128 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, 1u, 2u, 3u);
  fence_proxy_async_smem();
  __syncthreads();
  cluster_sync_all();
  const uint32_t dst_remote = mapa(smem_u32(smem + 64 * 1024), rank ^ 1u);
  const uint32_t bar_remote = mapa(smem_u32(&bar), rank ^ 1u);
  const long long t0 = clock64();
  if (rank == 0) {
    if ((flags & 1) && tid == 0) {
      for (int r = 0; r < reps; ++r)
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         dst_remote + (uint32_t)((r & 1) * 32 * 1024)), "r"(smem_u32(smem) + (uint32_t)((r & 1) * 32768)), "r"(32768), "r"(bar_remote) : "memory");
    }
    if ((flags & 2) && tid == 32) {
      uint8_t* mine = gbuf + (size_t)pair * 262144u;
      for (int r = 0; r < reps; ++r) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine + (size_t)(r & 7) * 32768), "r"(smem_u32(smem) + (uint32_t)((r & 1) * 32768)), "r"(32768u) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      out[4 * pair + 1] = clock64() - t0;
    }
  } else if ((flags & 1) && tid == 0) {
    mbar_expect_tx(&bar, (uint32_t)(reps * 32 * 1024));
    if (!mbar_wait(&bar, 0, 1u << 24)) out[4 * pair + 2] = -1;
    out[4 * pair + 0] = clock64() - t0;
  }
  cluster_sync_all();
}

extern "C" int egress2_run(int flags, int reps, int npairs, uint8_t* gbuf, long long* out_dev, void* stream) {
  size_t smem = 128 * 1024;
  if (cudaFuncSetAttribute(egress2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
  egress2_kernel<<<2 * npairs, 128, smem, (cudaStream_t)stream>>>(flags, reps, gbuf, out_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

// D[128 x 64] = A[128 x 64] (bf16, smem K-major) * B[64 x 64] (fp16, smem MN-major), one CTA
__global__ void __launch_bounds__(128, 1)
mixed_fmt_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int a_fmt, int b_fmt,
                 int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = 64, N = 64;
  uint8_t* a_img = smem; uint8_t* b_img = smem + 32 * 1024;
  for (int idx = tid; idx < 128 * K; idx += 128) {
    int m = idx / K, k = idx % K;
    uint8_t* p = a_img + img_off(m, k, 128, 16 * 128);
    if (a_fmt) *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(A[m * K + k]);
    else *reinterpret_cast<__half*>(p) = __float2half_rn(A[m * K + k]);
  }
  for (int idx = tid; idx < K * N; idx += 128) {
    int k = idx / N, n = idx % N;
    uint8_t* p = b_img + img_off(k, n, 128, (K / 8) * 128);
    if (b_fmt) *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(B[k * N + n]);
    else *reinterpret_cast<__half*>(p) = __float2half_rn(B[k * N + n]);
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  if (warp == 0) {
    uint32_t idesc = (1u << 4) | ((uint32_t)(a_fmt & 1) << 7) | ((uint32_t)(b_fmt & 1) << 10) | (1u << 16) |
                     ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int ks = 0; ks < K / 16; ++ks)
      mma_ss(tbase, make_desc(smem_u32(a_img) + ks * 2 * 2048, 2048, 128), make_desc(smem_u32(b_img) + ks * 2 * 128, 128, (K / 8) * 128),
             idesc, ks > 0);
    mma_commit(&bar);
  }
  bool ok = mbar_wait(&bar, 0, 1u << 22);
  tc_fence_after_sync();
  if (!ok) {
    if (lane == 0) atomicExch(status, 1);
  } else {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 64);
}

extern "C" int mixed_fmt_run(const float* A, const float* B, float* D, int a_fmt, int b_fmt, int* status_dev, void* stream) {
  size_t smem = 64 * 1024;
  if (cudaFuncSetAttribute(mixed_fmt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
  mixed_fmt_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, a_fmt, b_fmt, status_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

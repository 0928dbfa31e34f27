// Hardware probe for the UMMA operand forms the tcgen05 kernels rely on (run by scripts/run_umma_probe.py
// and tests/test_gpu_umma.py).  One CTA computes D[128 x N] = A[128 x K] * B[K x N] with bf16 operands:
//   a_mode 0: A image in smem, read K-major         a_mode 1: A in TMEM (lane = row, 2 bf16 per column)
//   b_mode 0: B image [k][n] read MN-major (N contiguous)   b_mode 1: B image [n][k] read K-major
//   a_mode 2: A image [k][m] in smem read MN-major (wgrad form: A^T stored with M contiguous)
// swap_a / swap_b exchange the LBO/SBO fields (hypothesis testing).  All waits are bounded.
#include <cuda_runtime.h>
#include <stdio.h>
#include "umma.cuh"

using namespace umma;

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, int N,
                  int a_mode, int b_mode, int swap_a, int swap_b, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* a_img = smem;                                   // up to 160*128*2 = 40 KB
  uint8_t* b_img = smem + 48 * 1024;                       // up to 160*160*2 = 50 KB
  const int KG = K / 8, NG = N / 8;
  // ---- A image ----
  uint32_t a_RS = 128, a_CS;
  if (a_mode == 0) {           // rows = m (128), cols = k
    a_CS = 16 * 128;
    for (int idx = tid; idx < 128 * K; idx += 128) {
      int m = idx / K, k = idx % K;
      *reinterpret_cast<__nv_bfloat16*>(a_img + img_off(m, k, a_RS, a_CS)) = __float2bfloat16_rn(A[m * K + k]);
    }
  } else if (a_mode == 2) {    // rows = k, cols = m (M contiguous): A^T image, read MN-major
    a_CS = (uint32_t)KG * 128;
    for (int idx = tid; idx < 128 * K; idx += 128) {
      int m = idx / K, k = idx % K;
      *reinterpret_cast<__nv_bfloat16*>(a_img + img_off(k, m, a_RS, a_CS)) = __float2bfloat16_rn(A[m * K + k]);
    }
  } else {
    a_CS = 0;
  }
  // ---- B image ----
  uint32_t b_RS = 128, b_CS;
  if (b_mode == 0) {           // rows = k, cols = n
    b_CS = (uint32_t)KG * 128;
    for (int idx = tid; idx < K * N; idx += 128) {
      int k = idx / N, n = idx % N;
      *reinterpret_cast<__nv_bfloat16*>(b_img + img_off(k, n, b_RS, b_CS)) = __float2bfloat16_rn(B[k * N + n]);
    }
  } else {                     // rows = n, cols = k
    b_CS = (uint32_t)NG * 128;
    for (int idx = tid; idx < K * N; idx += 128) {
      int k = idx / N, n = idx % N;
      *reinterpret_cast<__nv_bfloat16*>(b_img + img_off(n, k, b_RS, b_CS)) = __float2bfloat16_rn(B[k * N + n]);
    }
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t d_col = 0, a_col = 256;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (a_mode == 1) {           // thread = row m; column c holds (k=2c low half, k=2c+1 high half)
    int m = tid;
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = pack_bf16x2(A[m * K + 2 * (c0 + j)], A[m * K + 2 * (c0 + j) + 1]);
      tmem_st8(tbase + lane_base + a_col + c0, r);
    }
    tmem_wait_st();
  }
  fence_proxy_async_smem();    // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc(128, N, a_mode == 2 ? 1 : 0, b_mode == 0 ? 1 : 0);
    for (int ks = 0; ks < K / 16; ++ks) {
      // B: K-group stride = RS (MN-major image [k][n]) or CS (K-major image [n][k])
      uint64_t bd;
      if (b_mode == 0) {
        uint32_t lbo = b_RS, sbo = b_CS;
        if (swap_b) { uint32_t t = lbo; lbo = sbo; sbo = t; }
        bd = make_desc(smem_u32(b_img) + ks * 2 * b_RS, lbo, sbo);
      } else {
        uint32_t lbo = b_CS, sbo = b_RS;
        if (swap_b) { uint32_t t = lbo; lbo = sbo; sbo = t; }
        bd = make_desc(smem_u32(b_img) + ks * 2 * b_CS, lbo, sbo);
      }
      if (a_mode == 1) {
        mma_ts(tbase + d_col, tbase + a_col + ks * 8, bd, idesc, ks > 0);
      } else if (a_mode == 0) {
        uint32_t lbo = a_CS, sbo = a_RS;
        if (swap_a) { uint32_t t = lbo; lbo = sbo; sbo = t; }
        mma_ss(tbase + d_col, make_desc(smem_u32(a_img) + ks * 2 * a_CS, lbo, sbo), bd, idesc, ks > 0);
      } else {
        uint32_t lbo = a_RS, sbo = a_CS;
        if (swap_a) { uint32_t t = lbo; lbo = sbo; sbo = t; }
        mma_ss(tbase + d_col, make_desc(smem_u32(a_img) + ks * 2 * a_RS, lbo, sbo), bd, idesc, ks > 0);
      }
    }
    mma_commit(&bar);
  }
  bool ok = mbar_wait(&bar, 0, 1u << 22);
  tc_fence_after_sync();
  if (!ok) {
    if (lane == 0) atomicExch(status, 1);
  } else {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tbase + lane_base + d_col + c0, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

extern "C" int umma_probe_run(const float* A, const float* B, float* D, int K, int N, int a_mode, int b_mode,
                              int swap_a, int swap_b, int* status_dev, void* stream) {
  if (K % 16 || K > 160 || N % 16 || N > 160 || N < 16) return 1;
  size_t smem = 100 * 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return 2;
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, K, N, a_mode, b_mode, swap_a, swap_b, status_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

// ---------------------------------------------------------------------------------------------------------
// Throughput micro-benchmark of the UMMA operand forms/layouts (scripts/run_umma_bench.py): every CTA issues
// reps x nks MMAs (M=128, N, K=16 each) into one accumulator and reports cycles per MMA.  Operand contents are
// irrelevant (whatever is in shared memory); strides select the SWIZZLE_NONE core-matrix arrangement:
//   a_kstride/a_mstride, b_kstride/b_nstride = byte distance between adjacent K groups / MN groups (8 elements).
// a_mode: 0 smem K-major, 1 TMEM, 2 smem MN-major.  b_mn_major: 1 = MN-major, 0 = K-major.
__global__ void __launch_bounds__(128, 1)
umma_bench_kernel(int N, int nks, int reps, int a_mode, int b_mn_major, uint32_t a_kstride, uint32_t a_mstride,
                  uint32_t b_kstride, uint32_t b_nstride, int f16, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  if (warp == 0) {
    uint32_t idesc = (f16 & 1) ? make_idesc_f16(128, N, a_mode == 2 ? 1 : 0, b_mn_major) : make_idesc(128, N, a_mode == 2 ? 1 : 0, b_mn_major);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 80 * 1024);
    long long t0 = clock64();
    if (f16 & 8) {            // unrolled, descriptors precomputed, one elected lane for the whole batch
      const uint32_t b_lo0 = desc_lo(b_base, b_kstride), b_hi = desc_hi(b_nstride);
      const uint32_t a_lo0 = desc_lo(a_base, a_kstride), a_hi = desc_hi(a_mstride);
      const uint32_t bstep = (2u * b_kstride) >> 4, astep = (2u * a_kstride) >> 4;
      if (elect_one()) {
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            if (a_mode == 1) mma_ts_raw(tbase, tbase + 256u + (uint32_t)ks * 8u, b_lo0 + (uint32_t)ks * bstep, b_hi, idesc, 1u);
            else mma_ss_raw(tbase, a_lo0 + (uint32_t)ks * astep, a_hi, b_lo0 + (uint32_t)ks * bstep, b_hi, idesc, 1u);
          }
        }
        mma_commit_raw(&bar);
      }
      __syncwarp();
    } else
    for (int r = 0; r < reps; ++r)
      for (int ks = 0; ks < nks; ++ks) {
        uint64_t bd = make_desc(b_base + (uint32_t)ks * 2u * b_kstride, b_kstride, b_nstride);
        // f16 bit 1: alternate between two accumulators (independent chains); bit 2: no accumulate
        const uint32_t dcol = (f16 & 2) ? (uint32_t)(ks & 1) * 128u : 0u;
        const uint32_t accf = (f16 & 4) ? 0u : 1u;
        if (a_mode == 1) mma_ts(tbase + dcol, tbase + 256u + (uint32_t)ks * 8u, bd, idesc, accf);
        else mma_ss(tbase + dcol, make_desc(a_base + (uint32_t)ks * 2u * a_kstride, a_kstride, a_mstride), bd, idesc, accf);
      }
    if (!(f16 & 8)) mma_commit(&bar);
    mbar_wait(&bar, 0, 1u << 26);
    long long t1 = clock64();
    if ((tid & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

extern "C" int umma_bench_run(int grid, int N, int nks, int reps, int a_mode, int b_mn_major, uint32_t a_kstride,
                              uint32_t a_mstride, uint32_t b_kstride, uint32_t b_nstride, int f16, long long* out_dev,
                              void* stream) {
  size_t smem = 160 * 1024;
  if (cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
  umma_bench_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(N, nks, reps, a_mode, b_mn_major, a_kstride, a_mstride,
                                                             b_kstride, b_nstride, f16, out_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

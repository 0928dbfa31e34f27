// Issue-throughput probe behind the epilogue design of the tcgen05 kernels (scripts/run_epi_probe.py): which SM
// pipe each instruction of the activation epilogue runs on (F2FP conversions, LOP3/PRMT, FADD, HSET2, HADD2.F32 ...)
// and at what rate, alone and mixed.  Every op is a loop-carried chain (8 independent chains per op kind and thread),
// so nothing can be hoisted or eliminated; 16 warps per SM = 4 per SMSP, one CTA per SM, like the epilogue warps.
// Not part of the product library.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace {

enum Op { F2FP_F16 = 0, F2FP_RZRELU, F2FP_BF16, LOP3, FADD, HSET2, HADD2_F32, PRMT, FFMA, FMNMX, IMAD, SHF, F2FP_E4M3, NOPS };

template <int OP>
__device__ __forceinline__ void step(uint32_t& r, uint32_t c) {
  if (OP == F2FP_F16) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(r)), "f"(__uint_as_float(c)));
  else if (OP == F2FP_RZRELU) asm volatile("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(r)), "f"(__uint_as_float(c)));
  else if (OP == F2FP_BF16) asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(r)), "f"(__uint_as_float(c)));
  else if (OP == LOP3) asm volatile("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(r) : "r"(r), "r"(c), "r"(0x0f0f1234u));
  else if (OP == FADD) { float f; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(r)), "f"(__uint_as_float(c))); r = __float_as_uint(f); }
  else if (OP == HSET2) asm volatile("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(c));
  else if (OP == HADD2_F32) { float f; asm volatile("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, l;\n\t}" : "=f"(f) : "r"(r)); r = __float_as_uint(f); }
  else if (OP == PRMT) asm volatile("prmt.b32 %0, %1, %2, 0xBB99;" : "=r"(r) : "r"(r), "r"(c));
  else if (OP == FFMA) { float f; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(f) : "f"(__uint_as_float(r)), "f"(__uint_as_float(c)), "f"(1.0e-3f)); r = __float_as_uint(f); }
  else if (OP == FMNMX) { float f; asm volatile("max.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(r)), "f"(__uint_as_float(c))); r = __float_as_uint(f); }
  else if (OP == IMAD) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(r), "r"(c), "r"(12345u));
  else if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %1, %2, 7;" : "=r"(r) : "r"(r), "r"(c));
  else if (OP == F2FP_E4M3) { uint16_t h; asm volatile("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(h) : "f"(__uint_as_float(r)), "f"(__uint_as_float(c))); r = (r & 0xffff0000u) | h; }
}

// counts[k] ops of kind k per loop iteration and chain group (compile-time mix); 8 chains per kind
template <int A, int NA, int B, int NB, int C, int NC, int D, int ND>
__global__ void __launch_bounds__(512, 1) mix_kernel(const uint32_t* __restrict__ seed, uint32_t* __restrict__ out, int iters,
                                                     long long* __restrict__ cycles) {
  uint32_t ra[8], rb[8], rc[8], rd[8];
  const uint32_t c0 = seed[threadIdx.x & 31] | 0x3f000000u;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ra[j] = seed[(threadIdx.x + j) & 255]; rb[j] = seed[(threadIdx.x + 8 + j) & 255];
    rc[j] = seed[(threadIdx.x + 16 + j) & 255]; rd[j] = seed[(threadIdx.x + 24 + j) & 255];
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int n = 0; n < NA; ++n) step<A>(ra[j], c0);
#pragma unroll
        for (int n = 0; n < NB; ++n) step<B>(rb[j], c0);
#pragma unroll
        for (int n = 0; n < NC; ++n) step<C>(rc[j], c0);
#pragma unroll
        for (int n = 0; n < ND; ++n) step<D>(rd[j], c0);
      }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc ^= ra[j] ^ rb[j] ^ rc[j] ^ rd[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int A, int NA, int B, int NB, int C, int NC, int D, int ND>
int run_mix(const uint32_t* seed, uint32_t* out, int iters, long long* cycles, cudaStream_t st) {
  auto k = mix_kernel<A, NA, B, NB, C, NC, D, ND>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<148, 512, 200 * 1024, st>>>(seed, out, iters, cycles);
  return (int)cudaGetLastError();
}

}  // namespace

// id -> mix.  Returns the number of warp-instructions of the mix per loop iteration per warp (0 = unknown id).
extern "C" int epi_probe_run(int id, const uint32_t* seed, uint32_t* out, int iters, long long* cycles, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
#define SINGLE(ID, OP) case ID: run_mix<OP, 1, NOPS, 0, NOPS, 0, NOPS, 0>(seed, out, iters, cycles, st); return 32;
#define PAIR(ID, OP1, OP2) case ID: run_mix<OP1, 1, OP2, 1, NOPS, 0, NOPS, 0>(seed, out, iters, cycles, st); return 64;
  switch (id) {
    SINGLE(0, F2FP_F16) SINGLE(1, F2FP_RZRELU) SINGLE(2, F2FP_BF16) SINGLE(3, LOP3) SINGLE(4, FADD) SINGLE(5, HSET2)
    SINGLE(6, HADD2_F32) SINGLE(7, PRMT) SINGLE(8, FFMA) SINGLE(9, FMNMX) SINGLE(10, IMAD) SINGLE(11, SHF) SINGLE(12, F2FP_E4M3)
    PAIR(20, F2FP_F16, LOP3) PAIR(21, F2FP_F16, FADD) PAIR(22, F2FP_F16, HSET2) PAIR(23, F2FP_F16, HADD2_F32)
    PAIR(24, LOP3, FADD) PAIR(25, FADD, HSET2) PAIR(26, FADD, HADD2_F32) PAIR(27, LOP3, PRMT) PAIR(28, F2FP_F16, PRMT)
    PAIR(29, F2FP_F16, FFMA) PAIR(30, LOP3, HSET2) PAIR(31, LOP3, HADD2_F32) PAIR(32, HSET2, HADD2_F32) PAIR(33, F2FP_F16, F2FP_BF16)
    PAIR(34, LOP3, IMAD) PAIR(35, FADD, IMAD) PAIR(36, F2FP_F16, IMAD)
    // the forward epilogue's mix per activation pair today: 3 F2FP + 4 LOP3 + 2 FADD + 1 HSET2
    case 50: run_mix<F2FP_F16, 3, LOP3, 4, FADD, 2, HSET2, 1>(seed, out, iters, cycles, st); return 320;
    // fp16 saves (no bf16 copy): 2 F2FP + 3 LOP3 + 2 FADD + 1 HSET2
    case 51: run_mix<F2FP_F16, 2, LOP3, 3, FADD, 2, HSET2, 1>(seed, out, iters, cycles, st); return 256;
    // same with one residual through HADD2.F32 instead of a mantissa mask: 2 F2FP + 2 LOP3 + 2 FADD + (HSET2 + HADD2.F32)
    case 52: run_mix<F2FP_F16, 2, LOP3, 2, FADD, 2, HADD2_F32, 2>(seed, out, iters, cycles, st); return 256;
    // dgrad epilogue per pair: 1 F2FP + 1 LOP3 + 1 PRMT
    case 53: run_mix<F2FP_F16, 1, LOP3, 1, PRMT, 1, NOPS, 0>(seed, out, iters, cycles, st); return 96;
    default: return 0;
  }
}

// =====================================================================================================
// Stand-alone replica of the forward's layer epilogue (render_tc.cu): 16 warps, each thread one TMEM lane x 64 columns:
// tcgen05.ld -> fp16 hi (rz, relu) / lo planes -> tcgen05.st -> relu bit masks -> saved hi plane to global memory.
// No tensor core, no barriers: what the CUDA-core side of one SM can sustain, and which part of it costs what
// (flags switch parts off).  cycles[blk] = cycles for `iters` tile-layers per slot (2 slots in parallel).
// =====================================================================================================
#include "umma.cuh"
using namespace umma;
enum { EF_LDTM = 1, EF_SPLIT = 2, EF_STTM = 4, EF_MASK = 8, EF_STG = 16, EF_LO = 32 };

template <int FLAGS>
__global__ void __launch_bounds__(512, 1) epi_full_kernel(uint8_t* __restrict__ gbuf, size_t gbytes, int iters,
                                                          long long* __restrict__ cycles, uint32_t* __restrict__ sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const int slot = warp >> 3, half = (warp >> 2) & 1, q = warp & 3, row = q * 32 + lane;
  const uint32_t t_lane = tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * 256u;
  // seed the accumulator columns with something that is not constant
  {
    uint32_t init[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) init[j] = __float_as_uint(0.001f * (float)((tid * 7 + j * 13) % 1000) - 0.3f);
    for (int c = 0; c < 128; c += 16) tmem_st16(t_lane + (uint32_t)c, init);
    tmem_wait_st();
  }
  __syncthreads();
  uint32_t acc = 0, mw0 = 0, mw1 = 0;
  const size_t per_iter = 2u * 32768u;                       // both slots' tile-layer images
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t raw[2][32];
    if (FLAGS & EF_LDTM) {
      tmem_ld32(t_lane + (uint32_t)(half * 64), raw[0]);
      tmem_ld32(t_lane + (uint32_t)(half * 64 + 32), raw[1]);
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) { raw[0][j] = acc + j * 0x01010101u; raw[1][j] = acc ^ (j * 0x00100301u); }
    }
    uint8_t* act_img = gbuf + (((size_t)blockIdx.x * iters + it) * per_iter + (size_t)slot * 32768u) % gbytes;
#pragma unroll
    for (int sc = 0; sc < 4; ++sc) {
      const int cc = sc >> 1, c0 = half * 64 + sc * 16, j0 = (sc & 1) * 8;
      const uint32_t* rw = &raw[cc][(sc & 1) * 16];
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x0 = __uint_as_float(rw[2 * j]), x1 = __uint_as_float(rw[2 * j + 1]);
        if (FLAGS & EF_SPLIT) {
          asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi[j]) : "f"(x1), "f"(x0));
          if (FLAGS & EF_LO) {
            float r0, r1;
            asm("{\n\t.reg .b16 l, u, m;\n\tmov.b32 {l, u}, %2;\n\tmov.b16 m, 0xBC00;\n\t"
                "fma.rn.f32.f16 %0, l, m, %3;\n\tfma.rn.f32.f16 %1, u, m, %4;\n\t}" : "=f"(r0), "=f"(r1) : "r"(hi[j]), "f"(x0), "f"(x1));
            asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo[j]) : "f"(r1), "f"(r0));
          } else lo[j] = rw[2 * j];
        } else { hi[j] = rw[2 * j]; lo[j] = rw[2 * j + 1]; }
      }
      if (FLAGS & EF_STTM) {
        tmem_st8(t_lane + 128u + (uint32_t)(c0 >> 1), hi);
        tmem_st8(t_lane + 192u + (uint32_t)(c0 >> 1), lo);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc ^= lo[j];
      }
      if (FLAGS & EF_MASK) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t m = __hgt2_mask(*reinterpret_cast<const __half2*>(&hi[j]), __float2half2_rn(0.f)) &
                             ((1u << (8 * ((j0 + j) & 1) + 7 - ((j0 + j) >> 1))) | (1u << (8 * ((j0 + j) & 1) + 7 - ((j0 + j) >> 1) + 16)));
          if (cc) mw1 |= m; else mw0 |= m;
        }
      }
      if (FLAGS & EF_STG) {
#pragma unroll
        for (int g = 0; g < 2; ++g)
          *reinterpret_cast<uint4*>(act_img + (uint32_t)(row >> 3) * 128u + (uint32_t)((c0 >> 3) + g) * 2048u + (uint32_t)(row & 7) * 16u) =
              make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc ^= hi[j];
      }
    }
    if (FLAGS & EF_STTM) { tmem_wait_st(); }
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + tid] = acc ^ mw0 ^ mw1;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

extern "C" int epi_full_run(int flags, uint8_t* gbuf, size_t gbytes, int iters, long long* cycles, uint32_t* sink, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
#define EFCASE(F) case F: epi_full_kernel<F><<<148, 512, 0, st>>>(gbuf, gbytes, iters, cycles, sink); break;
  switch (flags) {
    EFCASE(63) EFCASE(63 - 16) EFCASE(63 - 8) EFCASE(63 - 8 - 16) EFCASE(63 - 4) EFCASE(63 - 1) EFCASE(1) EFCASE(1 + 4) EFCASE(2 + 32)
    EFCASE(63 - 32) EFCASE(1 + 16) EFCASE(16) EFCASE(2 + 32 + 8)
    default: return -1;
  }
  return (int)cudaGetLastError();
}

// =====================================================================================================
// SM egress probe: how many bytes per clock can ONE SM push to an L2-resident region (the cotangent ring of the fused
// backward: 264 KB per round per dgrad CTA), with 16-byte st.global from all warps (mode 0), with cp.async.bulk
// shared -> global issued by one thread (mode 1: one 32 KB copy in flight at a time, mode 2: four in flight), and both at
// once (mode 3)?  Every CTA owns a private 256 KB region (148 x 256 KB = 38 MB: L2 resident).
// =====================================================================================================
__global__ void __launch_bounds__(512, 1) egress_kernel(uint8_t* __restrict__ gbuf, int iters, int mode, long long* __restrict__ cycles,
                                                        long long* __restrict__ bytes) {
  extern __shared__ __align__(1024) uint8_t esm[];
  const int tid = threadIdx.x;
  uint8_t* mine = gbuf + (size_t)blockIdx.x * 262144u;
  for (int i = tid; i < 131072 / 16; i += 512) reinterpret_cast<uint4*>(esm)[i] = make_uint4(i, tid, 3u, 4u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  long long nb = 0;
  for (int it = 0; it < iters; ++it) {
    if (mode == 0 || mode == 3) {          // 512 threads x 16 B x 16 = 128 KB per iteration
#pragma unroll 4
      for (int k = 0; k < 16; ++k)
        *reinterpret_cast<uint4*>(mine + ((size_t)(k * 512 + tid) * 16u)) = make_uint4(it, k, tid, 7u);
      nb += 131072;
    }
    if ((mode == 1 || mode == 3) && tid == 0) {
      for (int k = 0; k < 4; ++k) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine + 131072 + k * 32768), "r"(smem_u32(esm + k * 32768)), "r"(32768u) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      nb += 131072;
    }
    if (mode == 2 && tid == 0) {
      for (int k = 0; k < 4; ++k) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine + 131072 + k * 32768), "r"(smem_u32(esm + k * 32768)), "r"(32768u) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      nb += 131072;
    }
    if ((mode == 5 || mode == 6) && tid == 0) {     // 16 groups of two 4 KB copies; mode 5 waits for each group's read, mode 6 lets one group run ahead
      for (int k = 0; k < 16; ++k) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine + 131072 + k * 8192), "r"(smem_u32(esm + (k & 3) * 8192)), "r"(4096u) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine + 131072 + k * 8192 + 4096), "r"(smem_u32(esm + (k & 3) * 8192 + 4096)), "r"(4096u) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (mode == 5) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
      nb += 131072;
    }
    if (mode == 4 && tid < 32) {            // 8 x 16 KB copies issued by 8 lanes
      if (tid < 8) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine + 131072 + tid * 16384), "r"(smem_u32(esm + tid * 16384)), "r"(16384u) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      nb += 131072;
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __threadfence();
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) { cycles[blockIdx.x] = t1 - t0; bytes[blockIdx.x] = nb; }
}

// SM ingress: cp.async.bulk global -> shared, four 32 KB copies in flight, from an L2-resident private region (mode 0)
// or streaming through a large buffer (mode 1: every CTA walks its own slice of `gbytes`, HBM).
__global__ void __launch_bounds__(128, 1) ingress_kernel(const uint8_t* __restrict__ gbuf, size_t gbytes, int iters, int mode,
                                                         long long* __restrict__ cycles, long long* __restrict__ bytes) {
  extern __shared__ __align__(1024) uint8_t esm[];
  __shared__ uint64_t bar[4];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int i = 0; i < 4; ++i) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i]))); } asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const size_t slice = gbytes / gridDim.x / 32768 * 32768;
  const uint8_t* mine = mode == 0 ? gbuf + (size_t)blockIdx.x * 262144u : gbuf + (size_t)blockIdx.x * slice;
  const size_t span = mode == 0 ? 262144u : slice;
  const long long t0 = clock64();
  if (tid == 0) {
    size_t off = 0;
    for (int it = 0; it < iters + 4; ++it) {
      const int b = it & 3;
      if (it >= 4) {
        uint32_t done = 0, ph = ((it >> 2) - 1) & 1;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar[b])), "r"(ph) : "memory");
      }
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(32768u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(esm + b * 32768)), "l"(mine + off), "r"(32768u), "r"(smem_u32(&bar[b])) : "memory");
        off += 32768; if (off + 32768 > span) off = 0;
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) { cycles[blockIdx.x] = t1 - t0; bytes[blockIdx.x] = (long long)iters * 32768; }
}

extern "C" int ingress_run(const uint8_t* gbuf, size_t gbytes, int iters, int mode, int nblocks, long long* cycles, long long* bytes, void* stream) {
  cudaFuncSetAttribute(ingress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  ingress_kernel<<<nblocks, 128, 131072 + 1024, (cudaStream_t)stream>>>(gbuf, gbytes, iters, mode, cycles, bytes);
  return (int)cudaGetLastError();
}

// the same two store forms, but STREAMING: every CTA walks through its own slice of a buffer far larger than L2 (what the
// forward's saved activations are): mode 0 = st.global.v4 from 16 warps, 1 = cp.async.bulk 2 x 4 KB per group, one group ahead,
// 2 = cp.async.bulk 32 KB per group, one group ahead
__global__ void __launch_bounds__(512, 1) egress_stream_kernel(uint8_t* __restrict__ gbuf, size_t gbytes, int iters, int mode,
                                                               long long* __restrict__ cycles, long long* __restrict__ bytes) {
  extern __shared__ __align__(1024) uint8_t esm[];
  const int tid = threadIdx.x;
  const size_t slice = gbytes / gridDim.x / 131072 * 131072;
  uint8_t* mine = gbuf + (size_t)blockIdx.x * slice;
  for (int i = tid; i < 131072 / 16; i += 512) reinterpret_cast<uint4*>(esm)[i] = make_uint4(i, tid, 3u, 4u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  size_t off = 0;
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll 4
      for (int k = 0; k < 16; ++k)
        *reinterpret_cast<uint4*>(mine + off + ((size_t)(k * 512 + tid) * 16u)) = make_uint4(it, k, tid, 7u);
    } else if (tid == 0) {
      if (mode == 1) {
        for (int k = 0; k < 16; ++k) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine + off + k * 8192), "r"(smem_u32(esm + (k & 3) * 8192)), "r"(4096u) : "memory");
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine + off + k * 8192 + 4096), "r"(smem_u32(esm + (k & 3) * 8192 + 4096)), "r"(4096u) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
      } else {
        for (int k = 0; k < 4; ++k) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine + off + k * 32768), "r"(smem_u32(esm + k * 32768)), "r"(32768u) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
      }
    }
    off += 131072; if (off + 131072 > slice) off = 0;
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __threadfence();
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) { cycles[blockIdx.x] = t1 - t0; bytes[blockIdx.x] = (long long)iters * 131072; }
}

extern "C" int egress_stream_run(uint8_t* gbuf, size_t gbytes, int iters, int mode, int nblocks, long long* cycles, long long* bytes, void* stream) {
  cudaFuncSetAttribute(egress_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  egress_stream_kernel<<<nblocks, 512, 131072 + 1024, (cudaStream_t)stream>>>(gbuf, gbytes, iters, mode, cycles, bytes);
  return (int)cudaGetLastError();
}

extern "C" int egress_run(uint8_t* gbuf, int iters, int mode, int nblocks, long long* cycles, long long* bytes, void* stream) {
  cudaFuncSetAttribute(egress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  egress_kernel<<<nblocks, 512, 131072 + 1024, (cudaStream_t)stream>>>(gbuf, iters, mode, cycles, bytes);
  return (int)cudaGetLastError();
}

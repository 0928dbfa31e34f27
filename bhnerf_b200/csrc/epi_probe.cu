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

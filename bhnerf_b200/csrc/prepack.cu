// Scene prepack: frame-independent domain mask, weight folding and ray-major compaction.
// Replaces the per-call masking of emission.fill_unsupervised_emission (bhnerf/emission.py:343-374,
// applied at network.py:231 with UN-warped coords) and folds w_s = g^2*dtau*Sigma*J_s
// (kgeo.radiative_trasfer, bhnerf/kgeo.py:618-621; J broadcast network.py:415-418).
// Dead samples contribute exactly 0 to images and gradients in the reference, so dropping them is
// result-preserving.
#include "common.cuh"

__device__ __forceinline__ bool bh_sample_active(const float* coords, const float* g, const float* dtau,
                                                 const float* Sigma, size_t PG, size_t idx, float rmin2,
                                                 float rmax2, float zw) {
  float x = coords[idx], y = coords[PG + idx], z = coords[2 * PG + idx];
  float r2 = x * x + y * y + z * z;
  // emission.py:370-373: zero where r^2 < rmin^2, r^2 > rmax^2, |z| > z_width (strict)
  bool dead = (r2 < rmin2) || (r2 > rmax2) || (fabsf(z) > zw);
  float w = g[idx] * g[idx] * dtau[idx] * Sigma[idx];
  return !dead && (w != 0.0f);
}

// one warp per ray: count active samples
__global__ void prepack_count_kernel(const float* __restrict__ coords, const float* __restrict__ g,
                                     const float* __restrict__ dtau, const float* __restrict__ Sigma,
                                     int P, int G, float rmin2, float rmax2, float zw,
                                     int32_t* __restrict__ counts) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P) return;
  size_t PG = (size_t)P * G;
  int cnt = 0;
  for (int k0 = 0; k0 < G; k0 += 32) {
    int k = k0 + lane;
    bool a = (k < G) && bh_sample_active(coords, g, dtau, Sigma, PG, (size_t)warp * G + k, rmin2, rmax2, zw);
    cnt += __popc(__ballot_sync(0xffffffffu, a));
  }
  if (lane == 0) counts[warp] = cnt;
}

// single-block exclusive scan in place: counts[0..P) -> row_ptr[0..P], total in row_ptr[P]
__global__ void prepack_scan_kernel(int32_t* __restrict__ rp, int P) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < P; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = (i < P) ? rp[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int ws = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
      int wi = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_sums[lane] = wi - ws;   // exclusive
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + warp_sums[wid] + incl - v;
    if (i < P) rp[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) rp[P] = carry_s;
}

// one warp per ray: scatter active samples to their compacted slots (ray-major order kept)
__global__ void prepack_fill_kernel(const float* __restrict__ coords, const float* __restrict__ Omega,
                                    const float* __restrict__ g, const float* __restrict__ dtau,
                                    const float* __restrict__ Sigma, const float* __restrict__ t_geos,
                                    const float* __restrict__ J, int P, int G, int S, float rmin2,
                                    float rmax2, float zw, const int32_t* __restrict__ row_ptr,
                                    float* __restrict__ fbase, int n_pad) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P) return;
  size_t PG = (size_t)P * G, np = (size_t)n_pad;
  float* ox = fbase; float* oy = fbase + np; float* oz = fbase + 2 * np;
  float* oom = fbase + 3 * np; float* otg = fbase + 4 * np; float* ow = fbase + 5 * np;
  int32_t* oray = (int32_t*)(fbase + (5 + (size_t)S) * np);
  int32_t* okidx = oray + np;
  int pos = row_ptr[warp];
  for (int k0 = 0; k0 < G; k0 += 32) {
    int k = k0 + lane;
    size_t idx = (size_t)warp * G + k;
    bool a = (k < G) && bh_sample_active(coords, g, dtau, Sigma, PG, idx, rmin2, rmax2, zw);
    unsigned m = __ballot_sync(0xffffffffu, a);
    if (a) {
      int o = pos + __popc(m & ((1u << lane) - 1u));
      ox[o] = coords[idx]; oy[o] = coords[PG + idx]; oz[o] = coords[2 * PG + idx];
      oom[o] = Omega[idx]; otg[o] = t_geos[idx];
      float w = g[idx] * g[idx] * dtau[idx] * Sigma[idx];
      for (int s = 0; s < S; ++s) ow[(size_t)s * np + o] = J ? w * J[(size_t)s * PG + idx] : w;
      oray[o] = warp;
      okidx[o] = k;
    }
    pos += __popc(m);
  }
}

__global__ void prepack_pad_kernel(float* __restrict__ fbase, int S, int n_active, int n_pad) {
  int i = n_active + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  size_t np = (size_t)n_pad;
  for (int a = 0; a < 5 + S; ++a) fbase[(size_t)a * np + i] = 0.0f;
  ((int32_t*)(fbase + (5 + (size_t)S) * np))[i] = -1;
  ((int32_t*)(fbase + (6 + (size_t)S) * np))[i] = 0;
}

extern "C" size_t bhnerf_packed_bytes(int32_t P, int32_t G, int32_t S) {
  size_t nmax = (size_t)bh_round_up(P * G, 128);
  return bh_rp_pad(P) * 4 + (size_t)(7 + S) * nmax * 4;
}

extern "C" int bhnerf_prepack(const float* coords, const float* Omega, const float* g, const float* dtau,
                              const float* Sigma, const float* t_geos, const float* J, int32_t P, int32_t G,
                              int32_t S, float rmin, float rmax, float z_width, void* packed,
                              size_t packed_bytes, bhnerf_scene_t* scene, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(P > 0 && G > 0 && S >= 1 && S <= 4, "prepack: bad shape P=%d G=%d S=%d", P, G, S);
  BH_REQUIRE((long long)P * G < (1ll << 31), "prepack: P*G too large");
  BH_REQUIRE(packed_bytes >= bhnerf_packed_bytes(P, G, S), "prepack: packed buffer too small");
  BH_REQUIRE(scene != nullptr, "prepack: scene is NULL");
  int32_t* rp = (int32_t*)packed;
  float rmin2 = rmin * rmin, rmax2 = rmax * rmax;
  int threads = 256, blocks = (P * 32 + threads - 1) / threads;
  bh_prof_begin(BH_CAT_MISC, 4, st); bh_prof_end(BH_CAT_MISC, st);
  prepack_count_kernel<<<blocks, threads, 0, st>>>(coords, g, dtau, Sigma, P, G, rmin2, rmax2, z_width, rp);
  prepack_scan_kernel<<<1, 1024, 0, st>>>(rp, P);
  int32_t n_active = 0;
  BH_CHECK_CUDA(cudaMemcpyAsync(&n_active, rp + P, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  BH_CHECK_CUDA(cudaStreamSynchronize(st));
  int n_pad = bh_round_up(n_active > 0 ? n_active : 1, 128);
  float* fbase = (float*)((char*)packed + bh_rp_pad(P) * 4);
  prepack_fill_kernel<<<blocks, threads, 0, st>>>(coords, Omega, g, dtau, Sigma, t_geos, J, P, G, S, rmin2,
                                                 rmax2, z_width, rp, fbase, n_pad);
  int npadfill = n_pad - n_active;
  if (npadfill > 0)
    prepack_pad_kernel<<<(npadfill + 127) / 128, 128, 0, st>>>(fbase, S, n_active, n_pad);
  BH_CHECK_CUDA(cudaGetLastError());
  scene->packed = packed; scene->n_active = n_active; scene->n_pad = n_pad;
  scene->P = P; scene->G = G; scene->S = S;
  return 0;
}

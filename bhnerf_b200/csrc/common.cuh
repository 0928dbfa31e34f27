// Shared definitions for the bhnerf_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/bhnerf_b200.h"

#define BH_TILE_SIMT 128           // samples per CTA in the fp32 SIMT kernels
#define BH_W 128                   // hidden width (network.py:148 net_width)
#define BH_NF 21                   // posenc features (network.py:98-122, deg 3)

// flat parameter offsets: layer i kernel (in,out) row-major then bias (out)
#define OFF_W0 0
#define OFF_B0 (OFF_W0 + 21 * 128)
#define OFF_W1 (OFF_B0 + 128)
#define OFF_B1 (OFF_W1 + 128 * 128)
#define OFF_W2 (OFF_B1 + 128)
#define OFF_B2 (OFF_W2 + 128 * 128)
#define OFF_W3 (OFF_B2 + 128)
#define OFF_B3 (OFF_W3 + 149 * 128)
#define OFF_W4 (OFF_B3 + 128)
#define OFF_B4 (OFF_W4 + 128)
static_assert(OFF_B4 + 1 == BHNERF_N_PARAMS, "param layout");

// ---- error plumbing (thread-local message, never throws across the ABI) ----
void bh_set_error(const char* fmt, ...);
#define BH_CHECK_CUDA(expr)                                                         \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      bh_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,            \
                   cudaGetErrorString(_e));                                         \
      return 2;                                                                     \
    }                                                                               \
  } while (0)
#define BH_REQUIRE(cond, ...)                                                       \
  do {                                                                              \
    if (!(cond)) { bh_set_error(__VA_ARGS__); return 1; }                           \
  } while (0)

// ---- launch accounting + optional per-category CUDA-event timing (bench.py's live roofline) ----
enum { BH_CAT_FWD = 0, BH_CAT_BWD = 1, BH_CAT_WGRAD = 2, BH_CAT_HEADS = 3, BH_CAT_MISC = 4, BH_CAT_COMM = 5, BH_CAT_VIS = 6,
       BH_NCAT = BHNERF_N_CATEGORIES };
static_assert(BH_CAT_VIS + 1 == BH_NCAT, "profile categories");
void bh_prof_begin(int cat, int n_launches, cudaStream_t st);
void bh_prof_end(int cat, cudaStream_t st);
struct BhProfScope {
  int cat; cudaStream_t st;
  BhProfScope(int c, int n, cudaStream_t s) : cat(c), st(s) { bh_prof_begin(c, n, s); }
  ~BhProfScope() { bh_prof_end(cat, st); }
};

// ---- packed scene view ----
// buffer layout: row_ptr[int32 rp_pad] | x | y | z | omega | tgeo | w[S] | ray[int32] | kidx[int32]  (each n_pad)
struct PackedView {
  const int32_t* row_ptr;  // [P+1] CSR offsets of each ray's active samples (ray-major order)
  const float* x; const float* y; const float* z; const float* omega; const float* tgeo;
  const float* w;          // [S][n_pad]  g^2*dtau*Sigma*J_s
  const int32_t* ray;      // [n_pad] ray index, -1 for padding
  const int32_t* kidx;     // [n_pad] sample index along the ray (dense index = ray*G + kidx)
  int n_active, n_pad, P, S;
};
__host__ __device__ inline int bh_round_up(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ inline size_t bh_rp_pad(int P) { return (size_t)bh_round_up(P + 1, 128); }
inline PackedView bh_view(const bhnerf_scene_t* sc) {
  PackedView v;
  const char* base = (const char*)sc->packed;
  size_t np = (size_t)sc->n_pad;
  v.row_ptr = (const int32_t*)base;
  const float* f = (const float*)(base + bh_rp_pad(sc->P) * 4);
  v.x = f; v.y = f + np; v.z = f + 2 * np; v.omega = f + 3 * np; v.tgeo = f + 4 * np;
  v.w = f + 5 * np;
  v.ray = (const int32_t*)(f + (5 + (size_t)sc->S) * np);
  v.kidx = v.ray + np;
  v.n_active = sc->n_active; v.n_pad = sc->n_pad; v.P = sc->P; v.S = sc->S;
  return v;
}

// per-frame scalars of the warp
struct FrameConsts { float t_start_obs, GM_c3, t_injection, scale; };

// ---- warp + positional encoding of one sample (bit-faithful time arithmetic) ----
// t_M = ((t_frame - t_start_obs)/GM_c3 + t_geos) - t_injection  in fp32, this op order, no
// contraction (emission.py:200-201); valid = !(t_M < 0) (emission.py:205, network.py:226);
// theta = t_M*Omega; rotation about z by -theta (emission.py:207-210, utils.py:126-132);
// u = valid ? warped/scale : 0 (network.py:227,229);
// feat = [u | sin(2^i u) i=0..2 (scale-major, xyz inner) | sin(2^i u + pi/2)] (network.py:98-122).
// The sin ARGUMENTS reproduce the reference's float32 arithmetic bit-for-bit: xb = u*2^i (exact),
// xb + fl32(pi/2), then safe_sin's python-sign modulo by fl32(100*pi) (network.py:16): every negative
// argument gets 314.15927f ADDED and rounded to the float32 grid at 314 (ulp 3.05e-5).  That rounding is
// part of the reference's result (it moves gradients by up to 1.4e-2), so it is kept;
// sinf() is the full-precision (non-MUFU) path.
#define BH_100PI_F 314.15927f      /* (float)(100*pi) */
#define BH_HALFPI_F 1.5707964f     /* (float)(pi/2)   */
__device__ __forceinline__ float bh_frame_time(float t_frame, const FrameConsts& fc) {
  return __fdiv_rn(__fsub_rn(t_frame, fc.t_start_obs), fc.GM_c3);
}
__device__ __forceinline__ float bh_safe_sin(float a) {
  float r = (fabsf(a) < BH_100PI_F) ? a : fmodf(a, BH_100PI_F);
  if (r < 0.0f) r = __fadd_rn(r, BH_100PI_F);
  return sinf(r);
}
__device__ __forceinline__ bool bh_features(float x, float y, float z, float om, float tg,
                                            float tfc, const FrameConsts& fc, float* f /*[21]*/) {
  float tM = __fsub_rn(__fadd_rn(tfc, tg), fc.t_injection);
  bool valid = !(tM < 0.0f);
  float th = __fmul_rn(tM, om);
  float sn, cs;
  sincosf(th, &sn, &cs);
  float u[3];
  u[0] = __fdiv_rn(x * cs + y * sn, fc.scale);
  u[1] = __fdiv_rn(y * cs - x * sn, fc.scale);
  u[2] = __fdiv_rn(z, fc.scale);
  if (!valid) { u[0] = 0.f; u[1] = 0.f; u[2] = 0.f; }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    f[c] = u[c];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float xb = u[c] * (float)(1 << i);
      f[3 + 3 * i + c] = bh_safe_sin(xb);
      f[12 + 3 * i + c] = bh_safe_sin(__fadd_rn(xb, BH_HALFPI_F));
    }
  }
  return valid;
}

__device__ __forceinline__ float bh_sigmoid_m10(float o) {   // sigmoid(o - 10), network.py:230
  return 1.0f / (1.0f + expf(10.0f - o));
}

// ---- kernel launchers implemented across the .cu files ----
int bh_launch_ray_integrate(const PackedView& v, const float* e, int Bt, float* images, cudaStream_t st);

// SIMT (fp32) family
// saved activations per frame (fp32): [h0,h1,h2,h3][128][n_pad] then feat [21][n_pad]
__host__ __device__ inline size_t bh_simt_acts_floats_per_frame(int n_pad) { return (size_t)(4 * 128 + BH_NF) * n_pad; }
__host__ __device__ inline size_t bh_simt_delta_floats_per_frame(int n_pad) { return (size_t)(4 * 128) * n_pad; }
int bh_simt_fwd(const PackedView& v, const FrameConsts& fc, const float* params, const float* t_frames,
                int Bt, float* e_out, float* acts /*or null*/, cudaStream_t st);
int bh_simt_bwd(const PackedView& v, const float* params, const float* d_images /*[Bt,S,P] slice*/,
                int Bt, const float* e_saved, const float* acts, float* delta_ws, float* wt_ws,
                float* d_params /*accumulated into*/, cudaStream_t st);
#define BH_SIMT_WT_FLOATS (3 * 128 * 128)

// TC (tcgen05) family.  `planes` = bf16 planes kept of every saved activation / cotangent (1: hi, 2: hi+lo).
bool bh_tc_fast();                                          // BHNERF_PRECISION=fast (stated single-product mode)
int bh_tc_planes(int n_active, int Bt_total);               // precision plan of a step
int bh_tc_planes_full_loss(int n_active, int Bt_total);     // ... of the fused step with the per-pixel image loss
size_t bh_tc_acts_bytes_per_frame(int n_pad, int planes);
size_t bh_tc_delta_bytes_per_frame(int n_pad, int planes); // backward scratch (delta images), 0 with the fused backward
size_t bh_tc_delta_fixed_bytes(int planes);                  // delta ring of the fused backward (frame-count independent)
size_t bh_tc_ws_bytes();                                     // weight images + status words
int bh_tc_prepare_weights(const float* params, void* ws, cudaStream_t st);
int bh_tc_fwd(const PackedView& v, const FrameConsts& fc, const void* ws, const float* params,
              const float* t_frames, int Bt, float* e_out, void* acts /*or null*/, int planes, cudaStream_t st);
int bh_tc_bwd(const PackedView& v, const void* ws, const float* params, const float* d_images, int Bt,
              const float* e_saved, const void* acts, void* delta_ws, float* dout_ws /*[Bt,n_pad], may alias e_saved*/,
              int planes, float* d_params /*accumulated into*/, cudaStream_t st);

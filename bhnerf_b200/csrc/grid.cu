// Grid renderers: the velocity warp + ray integral of the render path with a trilinear voxel lookup in place of the
// MLP.  Two reference functions share it:
//   mode 0  emission.image_plane_dynamics (bhnerf/emission.py:234-303): velocity_warp_coords -> interpolate_coords
//           (emission.py:213-232 = utils.world_to_image_coords, utils.py:160-166, + scipy.ndimage.map_coordinates
//           order=1, mode='constant', cval=0: a point outside [0, n-1] on any axis gives exactly 0) -> J broadcast ->
//           kgeo.radiative_trasfer (kgeo.py:595-622)
//   mode 1  network.GRID_Predictor.__call__ (bhnerf/network.py:306-352): the same warp, net_input =
//           (coords + scale) / (2 scale) * (res - 1), jax.scipy.ndimage.map_coordinates(order=1, cval=0) -- which
//           blends with cval corner by corner instead of cutting at the edge -- then sigmoid(v - 10), the domain
//           fill (done by the prepack) and the injection-time mask; plus its pull-back to the grid.
// Bound: L2/HBM gather (8 corner loads per evaluated sample; a 64^3 grid is 1 MB and stays in L2), no tensor work.
// One warp per (frame, ray): lanes stride over the ray's compacted samples (contiguous in the packed scene, so the
// geodesic stream is coalesced) and a shuffle reduction gives the pixel -- deterministic, no atomics in the forward.
#include "common.cuh"

namespace {

struct GridDesc {
  const float* grid;
  int nx, ny, nz;
  float fx, fy, fz;        // extent of the grid along each axis (world units); voxel centres span [-f/2, f/2]
};

__device__ __forceinline__ float sigmoid_m10_from(float v) { return 1.0f / (1.0f + expf(10.0f - v)); }

// image coordinates of the warped point (no division by the predictor scale here: world units)
__device__ __forceinline__ bool grid_coords(float x, float y, float z, float om, float tg, float tfc, const FrameConsts& fc,
                                            const GridDesc& gd, float* ic) {
  const float tM = __fsub_rn(__fadd_rn(tfc, tg), fc.t_injection);           // emission.py:200-201
  const bool valid = !(tM < 0.0f);
  float sn, cs;
  sincosf(__fmul_rn(tM, om), &sn, &cs);
  const float xw = x * cs + y * sn, yw = y * cs - x * sn;                    // rotation about z by -theta
  ic[0] = (xw + 0.5f * gd.fx) / gd.fx * (float)(gd.nx - 1);                  // utils.py:164
  ic[1] = (yw + 0.5f * gd.fy) / gd.fy * (float)(gd.ny - 1);
  ic[2] = (z + 0.5f * gd.fz) / gd.fz * (float)(gd.nz - 1);
  return valid;
}

// corner indices, weights and per-corner validity of the order-1 lookup.  Returns false if the point contributes 0.
template <int MODE>
__device__ __forceinline__ bool corners(const float* ic, const GridDesc& gd, int* i0, float* fr, bool* lo_ok, bool* hi_ok) {
  const int n[3] = {gd.nx, gd.ny, gd.nz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float c = ic[a];
    if (MODE == 0) {           // scipy 'constant': no interpolation beyond the edges
      if (!(c >= 0.0f && c <= (float)(n[a] - 1))) return false;
    } else {                   // jax: corner by corner; nothing within reach beyond one cell outside
      if (!(c > -1.0f && c < (float)n[a])) return false;
    }
    const float f = floorf(c);
    i0[a] = (int)f;
    fr[a] = c - f;
    lo_ok[a] = i0[a] >= 0 && i0[a] < n[a];
    hi_ok[a] = i0[a] + 1 >= 0 && i0[a] + 1 < n[a];
  }
  return true;
}

template <int MODE>
__device__ __forceinline__ float trilinear(const float* ic, const GridDesc& gd) {
  int i0[3]; float fr[3]; bool lo[3], hi[3];
  if (!corners<MODE>(ic, gd, i0, fr, lo, hi)) return 0.0f;
  float acc = 0.0f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int dx = c >> 2, dy = (c >> 1) & 1, dz = c & 1;
    const bool ok = (dx ? hi[0] : lo[0]) && (dy ? hi[1] : lo[1]) && (dz ? hi[2] : lo[2]);
    const float w = (dx ? fr[0] : 1.0f - fr[0]) * (dy ? fr[1] : 1.0f - fr[1]) * (dz ? fr[2] : 1.0f - fr[2]);
    if (ok) acc = fmaf(w, __ldg(gd.grid + ((size_t)(i0[0] + dx) * gd.ny + (i0[1] + dy)) * gd.nz + (i0[2] + dz)), acc);
  }
  return acc;
}

template <int MODE>
__global__ void __launch_bounds__(256)
grid_render_fwd_kernel(PackedView v, FrameConsts fc, GridDesc gd, const float* __restrict__ t_frames, int Bt,
                       float* __restrict__ images, float* __restrict__ e_out) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wid < (long long)Bt * v.P; wid += nwarps) {
    const int b = (int)(wid / v.P), p = (int)(wid - (long long)b * v.P);
    const int i0 = v.row_ptr[p], i1 = v.row_ptr[p + 1];
    const float tfc = bh_frame_time(t_frames[b], fc);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = i0 + lane; i < i1; i += 32) {
      float ic[3];
      const bool valid = grid_coords(v.x[i], v.y[i], v.z[i], v.omega[i], v.tgeo[i], tfc, fc, gd, ic);
      float e = 0.0f;
      if (valid) {             // before injection: NaN coordinates in the reference; mode 1 masks them (network.py:348)
        e = trilinear<MODE>(ic, gd);
        if (MODE == 1) e = sigmoid_m10_from(e);
      }
      if (e_out) e_out[(size_t)b * v.n_pad + i] = e;
#pragma unroll
      for (int s = 0; s < 4; ++s)
        if (s < v.S) acc[s] = fmaf(e, v.w[(size_t)s * v.n_pad + i], acc[s]);
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s >= v.S) break;
      float a = acc[s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) images[((size_t)b * v.S + s) * v.P + p] = a;
    }
  }
}

// d grid: every evaluated sample scatters  (sum_s dI[b,s,ray] w[s,i]) * sigmoid'(v) * corner weight  (mode 1 only)
__global__ void __launch_bounds__(256)
grid_render_bwd_kernel(PackedView v, FrameConsts fc, GridDesc gd, const float* __restrict__ t_frames, int Bt,
                       const float* __restrict__ d_images, float* __restrict__ d_grid) {
  const size_t n = (size_t)Bt * v.n_pad;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / v.n_pad), i = (int)(idx - (size_t)b * v.n_pad);
    const int ray = v.ray[i];
    if (ray < 0) continue;
    float gsum = 0.f;
    for (int s = 0; s < v.S; ++s) gsum += d_images[((size_t)b * v.S + s) * v.P + ray] * v.w[(size_t)s * v.n_pad + i];
    if (gsum == 0.f) continue;
    float ic[3];
    if (!grid_coords(v.x[i], v.y[i], v.z[i], v.omega[i], v.tgeo[i], bh_frame_time(t_frames[b], fc), fc, gd, ic)) continue;
    int i0[3]; float fr[3]; bool lo[3], hi[3];
    if (!corners<1>(ic, gd, i0, fr, lo, hi)) continue;
    const float e = sigmoid_m10_from(trilinear<1>(ic, gd));
    const float dv = gsum * e * (1.0f - e);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int dx = c >> 2, dy = (c >> 1) & 1, dz = c & 1;
      const bool ok = (dx ? hi[0] : lo[0]) && (dy ? hi[1] : lo[1]) && (dz ? hi[2] : lo[2]);
      const float w = (dx ? fr[0] : 1.0f - fr[0]) * (dy ? fr[1] : 1.0f - fr[1]) * (dz ? fr[2] : 1.0f - fr[2]);
      if (ok && w != 0.f) atomicAdd(d_grid + ((size_t)(i0[0] + dx) * gd.ny + (i0[1] + dy)) * gd.nz + (i0[2] + dz), dv * w);
    }
  }
}

// emission.interpolate_coords (bhnerf/emission.py:213-232) as a stand-alone stage: coords [N,3] (world units, last axis
// x,y,z as velocity_warp_coords returns them) -> trilinear lookup; NaN coordinates give 0 (mode 0: scipy 'constant').
template <int MODE>
__global__ void __launch_bounds__(256)
interpolate_coords_kernel(GridDesc gd, const float* __restrict__ coords, long long N, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    float ic[3];
    ic[0] = (coords[3 * i + 0] + 0.5f * gd.fx) / gd.fx * (float)(gd.nx - 1);
    ic[1] = (coords[3 * i + 1] + 0.5f * gd.fy) / gd.fy * (float)(gd.ny - 1);
    ic[2] = (coords[3 * i + 2] + 0.5f * gd.fz) / gd.fz * (float)(gd.nz - 1);
    out[i] = trilinear<MODE>(ic, gd);
  }
}

int check_grid_args(const bhnerf_scene_t* sc, const float* grid, int nx, int ny, int nz, float fx, float fy, float fz,
                    const float* t_frames, int Bt) {
  BH_REQUIRE(sc && sc->packed, "grid_render: scene is NULL / not prepacked");
  BH_REQUIRE(sc->n_pad > 0 && sc->n_pad % 128 == 0 && sc->S >= 1 && sc->S <= 4, "grid_render: bad n_pad/S");
  BH_REQUIRE(grid && t_frames && Bt > 0, "grid_render: NULL argument or Bt <= 0");
  BH_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2, "grid_render: the grid needs >= 2 voxels per axis");
  BH_REQUIRE(fx > 0.f && fy > 0.f && fz > 0.f && sc->GM_c3 > 0.f, "grid_render: fov and GM_c3 must be > 0");
  return 0;
}

}  // namespace

extern "C" int bhnerf_grid_render_fwd(const bhnerf_scene_t* sc, const float* grid, int32_t nx, int32_t ny, int32_t nz,
                                      float fov_x, float fov_y, float fov_z, int32_t mode, const float* t_frames,
                                      int32_t Bt, float* images, float* e_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int r = check_grid_args(sc, grid, nx, ny, nz, fov_x, fov_y, fov_z, t_frames, Bt)) return r;
  BH_REQUIRE(images && (mode == 0 || mode == 1), "grid_render_fwd: images is NULL or unknown mode %d", mode);
  PackedView v = bh_view(sc);
  FrameConsts fc; fc.t_start_obs = sc->t_start_obs; fc.GM_c3 = sc->GM_c3; fc.t_injection = sc->t_injection; fc.scale = sc->scale;
  GridDesc gd{grid, nx, ny, nz, fov_x, fov_y, fov_z};
  BhProfScope ps(BH_CAT_FWD, 1, st);
  long long warps = (long long)Bt * v.P;
  int blocks = (int)((warps + 7) / 8); if (blocks > 148 * 64) blocks = 148 * 64;
  if (mode == 0) grid_render_fwd_kernel<0><<<blocks, 256, 0, st>>>(v, fc, gd, t_frames, Bt, images, e_out);
  else grid_render_fwd_kernel<1><<<blocks, 256, 0, st>>>(v, fc, gd, t_frames, Bt, images, e_out);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bhnerf_interpolate_coords(const float* grid, int32_t nx, int32_t ny, int32_t nz, float fov_x, float fov_y,
                                         float fov_z, int32_t mode, const float* coords, int64_t N, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(grid && coords && out && N > 0, "interpolate_coords: NULL argument or N <= 0");
  BH_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2 && fov_x > 0.f && fov_y > 0.f && fov_z > 0.f, "interpolate_coords: bad grid");
  BH_REQUIRE(mode == 0 || mode == 1, "interpolate_coords: unknown mode %d", mode);
  GridDesc gd{grid, nx, ny, nz, fov_x, fov_y, fov_z};
  BhProfScope ps(BH_CAT_MISC, 1, st);
  int blocks = (int)((N + 255) / 256); if (blocks > 148 * 32) blocks = 148 * 32;
  if (mode == 0) interpolate_coords_kernel<0><<<blocks, 256, 0, st>>>(gd, coords, (long long)N, out);
  else interpolate_coords_kernel<1><<<blocks, 256, 0, st>>>(gd, coords, (long long)N, out);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bhnerf_grid_render_bwd(const bhnerf_scene_t* sc, const float* grid, int32_t nx, int32_t ny, int32_t nz,
                                      float fov_x, float fov_y, float fov_z, const float* t_frames, int32_t Bt,
                                      const float* d_images, float* d_grid, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int r = check_grid_args(sc, grid, nx, ny, nz, fov_x, fov_y, fov_z, t_frames, Bt)) return r;
  BH_REQUIRE(d_images && d_grid, "grid_render_bwd: NULL argument");
  PackedView v = bh_view(sc);
  FrameConsts fc; fc.t_start_obs = sc->t_start_obs; fc.GM_c3 = sc->GM_c3; fc.t_injection = sc->t_injection; fc.scale = sc->scale;
  GridDesc gd{grid, nx, ny, nz, fov_x, fov_y, fov_z};
  BH_CHECK_CUDA(cudaMemsetAsync(d_grid, 0, (size_t)nx * ny * nz * sizeof(float), st));
  BhProfScope ps(BH_CAT_BWD, 1, st);
  size_t n = (size_t)Bt * v.n_pad;
  int blocks = (int)((n + 255) / 256); if (blocks > 148 * 32) blocks = 148 * 32;
  grid_render_bwd_kernel<<<blocks, 256, 0, st>>>(v, fc, gd, t_frames, Bt, d_images, d_grid);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// =====================================================================================================
// Geodesic post-processing: the per-sample algebra between kgeo's tracer output and network.raytracing_args.
//   Geodesics.get_dataset (kgeo/kgeo/kerr_raytracing_utils.py:220-279): x,y,z from Boyer-Lindquist (r,theta,phi),
//     Sigma = r^2 + a^2 cos^2 theta, dtau = concat(0, diff(mino)) along the ray
//   Keplerian angular velocity (bhnerf Tutorial3 cell 2 / alma.py:49): sign * sqrt(M) / (r^1.5 + a sqrt(M))
//   kgeo.azimuthal_velocity_vector + doppler_factor (bhnerf/kgeo.py:199-248): u^t from the Kerr metric, only k_t = -E and
//     k_phi = E lam survive for an azimuthal 4-velocity (kgeo.py:111-114)  =>  g = 1 / (u^t (1 - lam Omega)); NaN -> fill
// float64 arithmetic (the reference runs it in numpy float64), float32 outputs in the layout bhnerf_prepack takes.
// One thread per sample; HBM-bound, 40 B read + 36 B written per sample, runs once per (spin, inclination).
// =====================================================================================================
__global__ void __launch_bounds__(256)
geodesic_inputs_kernel(const double* __restrict__ r, const double* __restrict__ th, const double* __restrict__ ph,
                       const double* __restrict__ t, const double* __restrict__ mino, const double* __restrict__ lam,
                       const double* __restrict__ omega_in, long long P, int G, double a, double M, double omega_sign,
                       double fillna, float* __restrict__ coords, float* __restrict__ Omega, float* __restrict__ g,
                       float* __restrict__ dtau, float* __restrict__ Sigma, float* __restrict__ t_geos) {
  const long long n = P * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / G;
    const int k = (int)(i - p * G);
    const double rr = r[i], sth = sin(th[i]), cth = cos(th[i]);
    coords[i] = (float)(rr * cos(ph[i]) * sth);
    coords[n + i] = (float)(rr * sin(ph[i]) * sth);
    coords[2 * n + i] = (float)(rr * cth);
    const double Sg = rr * rr + a * a * cth * cth;
    const double Delta = rr * rr + a * a - 2.0 * M * rr;
    const double Xi = (rr * rr + a * a) * (rr * rr + a * a) - a * a * Delta * sth * sth;
    Sigma[i] = (float)Sg;
    dtau[i] = k == 0 ? 0.f : (float)(mino[i] - mino[i - 1]);
    t_geos[i] = (float)t[i];
    const double Om = omega_in ? omega_in[i] : omega_sign * sqrt(M) / (rr * sqrt(rr) + a * sqrt(M));
    Omega[i] = (float)Om;
    const double g_tt = -(1.0 - 2.0 * M * rr / Sg), g_phph = Xi * sth * sth / Sg, g_tph = -2.0 * M * a * rr * sth * sth / Sg;
    const double ut = 1.0 / sqrt(-(g_tt + 2.0 * Om * g_tph + g_phph * Om * Om));
    double gg = 1.0 / -(-ut + lam[p] * ut * Om);                    // E cancels (kgeo.py:245)
    if (gg != gg) gg = fillna;
    g[i] = (float)gg;
  }
}

extern "C" int bhnerf_geodesic_inputs(const double* r, const double* theta, const double* phi, const double* t,
                                      const double* mino, const double* lam, const double* Omega_in, int64_t P, int32_t G,
                                      double spin, double M, double omega_sign, double fillna, float* coords,
                                      float* Omega, float* g, float* dtau, float* Sigma, float* t_geos, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(r && theta && phi && t && mino && lam && coords && Omega && g && dtau && Sigma && t_geos,
             "geodesic_inputs: NULL argument");
  BH_REQUIRE(P > 0 && G > 0 && M > 0.0, "geodesic_inputs: P, G and M must be > 0");
  BhProfScope ps(BH_CAT_MISC, 1, st);
  long long n = (long long)P * G;
  int blocks = (int)((n + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
  geodesic_inputs_kernel<<<blocks, 256, 0, st>>>(r, theta, phi, t, mino, lam, Omega_in, (long long)P, G, spin, M, omega_sign,
                                                 fillna, coords, Omega, g, dtau, Sigma, t_geos);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

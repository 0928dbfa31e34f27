/* bhnerf_b200 -- C ABI of the B200-native bhnerf render/train hot path.
 *
 * The reference (aviadlevis/bhnerf) has no native FFI for this path: the seam is its Python API
 * (SURVEY.md s8b).  Each entry point below names the reference function it replaces
 * (paths under the reference repo).  A JAX binding registers these through jax.ffi
 * (see INTEGRATION.md); the ctypes host in bhnerf_b200/_lib.py binds them directly.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; arrays are row-major fp32
 *     unless noted; complex64 is interleaved (re,im) fp32
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*) and returns;
 *     the only exception is bhnerf_prepack (one stream sync to return n_active)
 *   - no allocation on the call path: scratch comes from the caller-provided workspace
 *   - return value: 0 = ok, non-zero = error (message via bhnerf_last_error, thread-local)
 *   - no CPU fallback exists: on a machine without an sm_100 GPU the calls fail with an error
 */
#ifndef BHNERF_B200_H
#define BHNERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BHNERF_ABI_VERSION 2
#define BHNERF_N_PARAMS 55169      /* 21x128+128, 128x128+128 (x2), 149x128+128, 128x1+1 */
#define BHNERF_N_FEAT 21
#define BHNERF_WIDTH 128

/* which kernel family runs the MLP */
#define BHNERF_IMPL_SIMT 0         /* fp32 FFMA reference kernels (CUDA cores)              */
#define BHNERF_IMPL_TC   1         /* tcgen05 tensor cores, fp16 x3 split operands, fp32 acc */

/* loss kinds: bhnerf/network.py:476-484 ('full','lc') and :542-564 ('vis','amp','cphase') */
#define BHNERF_LOSS_FULL 0
#define BHNERF_LOSS_LC   1
#define BHNERF_LOSS_VIS  2
#define BHNERF_LOSS_AMP  3
#define BHNERF_LOSS_CPHASE 4

/* A prepacked scene = the frame-independent part of network.raytracing_args
 * (bhnerf/network.py:850-894) + the NeRF_Predictor constants (bhnerf/network.py:147-157).
 * Filled by bhnerf_prepack; plain data, may be copied freely. */
typedef struct bhnerf_scene {
  const void* packed;     /* device buffer written by bhnerf_prepack                       */
  int32_t n_active;       /* samples inside the recovery domain with non-zero weight       */
  int32_t n_pad;          /* n_active rounded up to a multiple of 128 (padding has w = 0)  */
  int32_t P;              /* rays = num_alpha*num_beta                                     */
  int32_t G;              /* samples per ray (ngeo)                                        */
  int32_t S;              /* Stokes channels: 1 when J is the scalar 1.0, else J.shape[0]  */
  float t_start_obs;      /* [t_units]  raytracing_args['t_start_obs']                     */
  float GM_c3;            /* GM/c^3 in t_units (emission.py:183-185)                       */
  float t_injection;      /* [M]                                                           */
  float scale;            /* NeRF_Predictor.scale (network.py:229)                         */
} bhnerf_scene_t;

const char* bhnerf_last_error(void);
int bhnerf_version(void);
/* number of SMs / compute capability of the current device; fails if it is not sm_100 */
int bhnerf_device_check(int* sm_count_host, int* cc_major_host, int* cc_minor_host);

/* ---- setup: replaces the per-call masking of emission.fill_unsupervised_emission
 * (bhnerf/emission.py:343-374, applied at network.py:231) and folds
 * w_s = g^2 * dtau * Sigma * J_s (kgeo.radiative_trasfer, bhnerf/kgeo.py:618-621; J broadcast
 * network.py:415-418).  The domain mask uses UN-warped coords so it is frame independent.
 * coords [3,P,G]; Omega,g,dtau,Sigma,t_geos [P,G]; J [S,P,G] or NULL (scalar 1.0).          */
size_t bhnerf_packed_bytes(int32_t P, int32_t G, int32_t S);   /* upper bound (all active)   */
int bhnerf_prepack(const float* coords, const float* Omega, const float* g, const float* dtau,
                   const float* Sigma, const float* t_geos, const float* J,
                   int32_t P, int32_t G, int32_t S, float rmin, float rmax, float z_width,
                   void* packed, size_t packed_bytes, bhnerf_scene_t* scene_host, void* stream);

/* ---- forward render: replaces network.image_plane_prediction (bhnerf/network.py:373-420) =
 * NeRF_Predictor.__call__ (:191-237: velocity_warp_coords emission.py:143-211, posenc :98-122,
 * MLP :18-64, sigmoid(o-10) :230, masks :231-232) + kgeo.radiative_trasfer (kgeo.py:595-622).
 * images [Bt,S,P] (overwritten).  e_out [Bt,n_pad] (required): per-sample masked emission, the
 * residual the backward needs.  acts_out or NULL: saved activations for the backward
 * (bhnerf_acts_bytes(scene,Bt,impl) bytes; fp16 (one-plane plan) or bf16 hi+lo tiles for TC,
 * fp32 rows for SIMT).                                                                       */
size_t bhnerf_acts_bytes(const bhnerf_scene_t* scene, int32_t Bt, int32_t impl);
size_t bhnerf_fwd_workspace_bytes(int32_t impl);   /* fp16 / bf16 weight images + status words of the TC family; 0 for SIMT */
int bhnerf_render_fwd(const bhnerf_scene_t* scene, const float* params, const float* t_frames,
                      int32_t Bt, float* images, float* e_out, void* acts_out, void* workspace,
                      size_t workspace_bytes, int32_t impl, void* stream);

/* ---- backward render: replaces the jax.value_and_grad pull-back through
 * image_plane_prediction (bhnerf/network.py:617, :677), gradient w.r.t. params only.
 * d_params [55169] is OVERWRITTEN.  e_saved/acts_saved from the matching forward, or NULL:
 * the backward then recomputes them frame-chunk by frame-chunk inside `workspace`
 * (>= bhnerf_bwd_workspace_bytes).                                                           */
size_t bhnerf_bwd_workspace_bytes(const bhnerf_scene_t* scene, int32_t Bt, int32_t impl);
/* the frame-count independent part of the two workspace sizes above/below (weight images + the
 * L2-resident delta ring of the fused tcgen05 backward); the rest scales with the frames per chunk */
size_t bhnerf_bwd_fixed_workspace_bytes(const bhnerf_scene_t* scene, int32_t Bt, int32_t impl);
int bhnerf_render_bwd(const bhnerf_scene_t* scene, const float* params, const float* t_frames,
                      int32_t Bt, const float* d_images, const float* e_saved,
                      const void* acts_saved, float* d_params, void* workspace,
                      size_t workspace_bytes, int32_t impl, void* stream);

/* ---- loss heads (value + d_images in one pass)
 * image: loss_fn_image (bhnerf/network.py:422-484). target/sigma/offset are [Bt,S,P] ('full')
 * or [Bt,S] ('lc').  loss[1] overwritten; d_images [Bt,S,P] overwritten.                     */
int bhnerf_loss_image(const float* images, const float* target, const float* sigma,
                      const float* offset, float loss_scale, int32_t kind, int32_t Bt, int32_t S,
                      int32_t P, float* loss, float* d_images, void* stream);
/* eht: loss_fn_eht (bhnerf/network.py:486-564).  A [Bt,V,P] complex64 (one DFT matrix per
 * frame), images [Bt,P] (S must be 1), target [Bt,V] complex64 ('vis') or fp32 ('amp'),
 * sigma [Bt,V].  vis [Bt,V] complex64 overwritten.
 * 'cphase' (network.py:555-559): A is [Bt,3,V,P] = call vis_fwd/vis_bwd with 3V rows; loss_vis then takes
 * vis/d_vis [Bt,3,V], target (radians) and sigma [Bt,V], and V = number of closure phases.      */
int bhnerf_vis_fwd(const float* A, const float* images, int32_t Bt, int32_t V, int32_t P,
                   float* vis, void* stream);
int bhnerf_loss_vis(const float* vis, const float* target, const float* sigma, float loss_scale,
                    int32_t kind, int32_t Bt, int32_t V, float* loss, float* d_vis, void* stream);
int bhnerf_vis_bwd(const float* A, const float* d_vis, int32_t Bt, int32_t V, int32_t P,
                   float* d_images, void* stream);

/* The whole eht head of a step in one call (loss_fn_eht after image_plane_prediction, network.py:542-564, and its
 * pull-back to the images): per GROUP of frames  vis = A I -> chi^2 (accumulated into loss[1]) -> d_images = A^H d_vis.
 * group_bytes = bytes of A per group: 0 = the whole batch (both passes stream A at the HBM roofline, the measured
 * optimum); a value <= ~48 MB lets the L2 serve the backward pass at the price of 3 launches per group.
 * rows = rows of A per frame: nvis ('vis','amp') or 3*ncphase ('cphase'); a polarization axis folds into Bt.
 * vis / d_vis [Bt,rows] complex64 scratch+output; d_images [Bt,P] overwritten (NULL: forward + loss only).          */
int bhnerf_vis_head(const float* A, const float* images, const float* target, const float* sigma,
                    float loss_scale, int32_t kind, int32_t Bt, int32_t rows, int32_t P, float* loss,
                    float* vis, float* d_vis, float* d_images, size_t group_bytes, void* stream);

/* ---- separable visibility head on the tensor cores (opt-in).  ehtim fills the per-frame matrix of loss_fn_eht
 * (bhnerf/network.py:542-544) with the Fourier kernel of the regular pixel grid,
 *   A[b,k,(i,j)] = pulse[b,k] * exp(-2 pi i (u[b,k] x_i + v[b,k] y_j)),   x_i = x0 + i dx (alpha axis), y_j = y0 + j dy.
 * Given (u, v) instead of A, vis = A vec(I) is a GEMM with the DFT factor of one axis (tcgen05, factors generated in shared
 * memory) and a row-wise contraction with the factor of the other: nothing of size nvis x npix is read.  uv [Bt,V,2] in
 * cycles per unit of x / y; pulse [Bt,V] complex64 or NULL (1); images / d_images [Bt,NA,NB] (NA, NB <= 128 forward;
 * NB <= 128 backward); vis / d_vis [Bt,V] complex64.  status_dev: 2 int32 of device scratch (word 0 != 0 after the
 * kernel = a bounded wait expired).  Same results as bhnerf_vis_fwd / bhnerf_vis_bwd with that A, to 1e-4 / 1e-3.      */
int bhnerf_vis_dft_fwd(const float* uv, const float* pulse, const float* images, int32_t Bt, int32_t V, int32_t NA,
                       int32_t NB, float x0, float dx, float y0, float dy, float* vis, int32_t* status_dev, void* stream);
int bhnerf_vis_dft_bwd(const float* uv, const float* pulse, const float* d_vis, int32_t Bt, int32_t V, int32_t NA,
                       int32_t NB, float x0, float dx, float y0, float dy, float* d_images, int32_t* status_dev,
                       void* stream);

/* ---- fused train step for the separable image losses: per frame chunk
 * fwd -> ray integral -> loss -> bwd with activations kept in `workspace`.  Replaces
 * gradient_step_image up to (not including) pmean/apply_gradients (network.py:617-619).
 * images [Bt,S,P], loss[1], d_params[55169] overwritten.                                     */
size_t bhnerf_train_workspace_bytes(const bhnerf_scene_t* scene, int32_t Bt, int32_t impl);
int bhnerf_train_step_image(const bhnerf_scene_t* scene, const float* params,
                            const float* t_frames, int32_t Bt, const float* target,
                            const float* sigma, const float* offset, float loss_scale,
                            int32_t kind, float* loss, float* images, float* d_params,
                            void* workspace, size_t workspace_bytes, int32_t impl, void* stream);

/* ---- stand-alone dense stages with the reference's public names (fused inside render_*).
 * velocity_warp_coords: emission.velocity_warp_coords (bhnerf/emission.py:143-211), rot_axis = z.
 *   coords [3,N], Omega/t_geos [N], out [Bt,N,3]; NaN where t_M < 0 (emission.py:205).
 * fill_unsupervised_emission: bhnerf/emission.py:343-374, in place on emission [R,N], coords [3,N].
 * radiative_transfer: kgeo.radiative_trasfer (bhnerf/kgeo.py:595-622), emission [R,P,G] -> out [R,P]. */
int bhnerf_velocity_warp_coords(const float* coords, const float* Omega, const float* t_geos,
                                int64_t N, const float* t_frames, int32_t Bt, float t_start_obs,
                                float GM_c3, float t_injection, float* out, void* stream);
int bhnerf_fill_unsupervised_emission(float* emission, const float* coords, int32_t R, int64_t N,
                                      float rmin, float rmax, float z_width, float fill_value,
                                      void* stream);
int bhnerf_radiative_transfer(const float* emission, const float* g, const float* dtau,
                              const float* Sigma, int32_t R, int32_t P, int32_t G, float* out,
                              void* stream);

/* ---- grid renderers: the same warp + ray integral with a trilinear voxel lookup instead of the MLP.
 * mode 0 = emission.image_plane_dynamics (bhnerf/emission.py:234-303: velocity_warp_coords -> interpolate_coords
 *   :213-232 [utils.world_to_image_coords utils.py:160-166 + scipy map_coordinates order=1, cval=0: exactly 0
 *   outside the grid extent] -> J broadcast -> kgeo.radiative_trasfer); samples before injection (NaN coordinates
 *   in the reference, never reached by its callers) contribute 0.
 * mode 1 = network.GRID_Predictor.__call__ (bhnerf/network.py:306-352) + image_plane_prediction: lookup with
 *   fov = 2*scale per axis and jax.scipy.ndimage.map_coordinates' corner-wise cval blending, sigmoid(v - 10), domain
 *   fill (from the prepack), injection mask.
 * `scene` = bhnerf_prepack of the geodesics (for mode 0 prepack with rmin = 0 and rmax / z_width large enough to
 * cover the grid: culled samples contribute exactly 0 either way).  grid [nx,ny,nz] fp32, voxel centres spanning
 * [-fov/2, fov/2] per axis.  images [Bt,S,P] overwritten; e_out [Bt,n_pad] or NULL (per-sample emission).
 * grid_render_bwd (mode 1): d_grid [nx,ny,nz] OVERWRITTEN with the pull-back of d_images [Bt,S,P] to the grid
 * (jax.value_and_grad w.r.t. params['grid'], network.py:617).                                                  */
int bhnerf_grid_render_fwd(const bhnerf_scene_t* scene, const float* grid, int32_t nx, int32_t ny,
                           int32_t nz, float fov_x, float fov_y, float fov_z, int32_t mode,
                           const float* t_frames, int32_t Bt, float* images, float* e_out, void* stream);
/* stand-alone lookup: emission.interpolate_coords (bhnerf/emission.py:213-232; mode 0 = scipy rule, 1 = jax rule).
 * coords [N,3] world units (x,y,z innermost, the layout velocity_warp_coords returns), out [N]; NaN coordinates -> 0. */
int bhnerf_interpolate_coords(const float* grid, int32_t nx, int32_t ny, int32_t nz, float fov_x,
                              float fov_y, float fov_z, int32_t mode, const float* coords, int64_t N,
                              float* out, void* stream);
int bhnerf_grid_render_bwd(const bhnerf_scene_t* scene, const float* grid, int32_t nx, int32_t ny,
                           int32_t nz, float fov_x, float fov_y, float fov_z, const float* t_frames,
                           int32_t Bt, const float* d_images, float* d_grid, void* stream);

/* ---- geodesic post-processing (setup, once per (spin, inclination)): tracer output -> the arrays of
 * network.raytracing_args.  Fuses Geodesics.get_dataset's algebra (kgeo/kgeo/kerr_raytracing_utils.py:220-279:
 * x,y,z, Sigma, dtau = concat(0, diff(mino))), the Keplerian angular velocity (Tutorial3 cell 2, alma.py:49) when
 * Omega_in is NULL (omega_sign = +-1 sets the rotation sense), and kgeo.azimuthal_velocity_vector + doppler_factor
 * (bhnerf/kgeo.py:199-248; NaN -> fillna).  Inputs float64 [P,G] (lam [P], per ray), outputs float32:
 * coords [3,P,G], Omega, g, dtau, Sigma, t_geos [P,G].                                                            */
int bhnerf_geodesic_inputs(const double* r, const double* theta, const double* phi, const double* t,
                           const double* mino, const double* lam, const double* Omega_in, int64_t P,
                           int32_t G, double spin, double M, double omega_sign, double fillna,
                           float* coords, float* Omega, float* g, float* dtau, float* Sigma,
                           float* t_geos, void* stream);

/* ---- polarization factors (setup, once per (spin, inclination)): J = (I, Q, U) per geodesic sample, the chain of
 * alma.image_plane_model (bhnerf/alma.py:47-60): kgeo.azimuthal_velocity_vector (bhnerf/kgeo.py:199-223) -> doppler_factor
 * (:225-248) -> magnetic_field_fluid_frame(arad, avert, ator) (:274-313) normalised by its mean strength inside the recovery
 * domain (alma.py:55-57) -> parallel_transport(Q_frac, V_frac = 0, spectral_index) (:438-519) -> nan_to_num.  float64
 * inputs: r, theta, affine [P,G] (tracer output; affine = Geodesics.sig_s), per-ray lam, eta, alpha, beta [P]
 * (kerr_raytracing_utils.py:264-266), Omega_in [P,G] or NULL (Keplerian, omega_sign = +-1).  J [3,P,G] float32, the `J`
 * entry of network.raytracing_args (before emission.rotate_evpa).  workspace >= bhnerf_polarization_workspace_bytes. */
size_t bhnerf_polarization_workspace_bytes(int64_t P, int32_t G);
int bhnerf_polarization_factors(const double* r, const double* theta, const double* affine, const double* lam,
                                const double* eta, const double* alpha, const double* beta, const double* Omega_in,
                                int64_t P, int32_t G, double spin, double inclination, double omega_sign, double arad,
                                double avert, double ator, double Q_frac, double rmin, double rmax, double z_width,
                                int32_t spectral_index, float* J, void* workspace, size_t workspace_bytes, void* stream);

/* ---- optimiser: optax.adam + polynomial_schedule(power=1) applied by
 * TrainState.apply_gradients (bhnerf/network.py:171-182, :621).  grad_scale multiplies the
 * gradient first (1/ndev turns an all-reduce SUM into jax.lax.pmean, network.py:620).
 * count = number of updates already applied.  guard (or NULL): device pointer to the health flags of the step
 * that produced `grads` (= the tcgen05 workspace, see bhnerf_workspace_status); if one is set the update is
 * skipped, so a gradient from an overflowed or aborted step is never applied.                */
/* dst[n] += src[n] on the device (per-chunk gradients / losses of the chunked eht step) */
int bhnerf_add_inplace(float* dst, const float* src, int32_t n, void* stream);
int bhnerf_adam_step(float* params, const float* grads, float* mu, float* nu, int32_t n,
                     int32_t count, float lr_init, float lr_final, int32_t transition_steps,
                     float b1, float b2, float eps, float grad_scale, const int32_t* guard, void* stream);

/* The same update with the step counter in device memory (*count_dev is read, then advanced by one): nothing in the
 * call depends on host state, so a whole train step can be captured in a CUDA graph and replayed.                */
int bhnerf_adam_step_dev(float* params, const float* grads, float* mu, float* nu, int32_t n,
                         int32_t* count_dev, float lr_init, float lr_final, int32_t transition_steps,
                         float b1, float b2, float eps, float grad_scale, const int32_t* guard, void* stream);

/* ---- gradient exchange across ranks (one process per GPU): replaces jax.lax.pmean(grads, 'batch')
 * (bhnerf/network.py:620, :680).  One NCCL all-reduce on `stream` (CUDA-graph capturable); NCCL is resolved with dlopen
 * at the first call.  Rank 0 calls bhnerf_comm_unique_id, the host side broadcasts the BHNERF_COMM_ID_BYTES bytes
 * (any out-of-band channel), every rank calls bhnerf_comm_init with its rank -- collectively, like ncclCommInitRank.
 * allreduce_mean: buf[n] <- mean over ranks (in place).  allreduce_sum: buf[n] <- sum over ranks: the partial
 * lightcurves / visibilities of ray sharding (each rank renders P/world rays of every frame).    */
#define BHNERF_COMM_ID_BYTES 128
int bhnerf_comm_unique_id(void* id_host);
int bhnerf_comm_init(int32_t rank, int32_t world, const void* id_host, void** comm_out_host);
int bhnerf_comm_destroy(void* comm);
int bhnerf_allreduce_mean(float* buf, int64_t n, void* comm, void* stream);
int bhnerf_allreduce_sum(float* buf, int64_t n, void* comm, void* stream);

/* ---- lightcurves of rendered images, and the 'lc' loss from them: the two halves of bhnerf_loss_image(kind = LC)
 * (bhnerf/network.py:478-480), split so that ray-sharded ranks can all-reduce their partial lightcurves in between.
 * lightcurve: lc[Bt,S] = sum_p images[Bt,S,P] (overwritten).  loss_lightcurve: loss[1] = scale*sum|(lc-t-off)/sigma|^2
 * and d_images[Bt,S,P] = 2*scale*(lc-t-off)/sigma^2 for every ray (both overwritten).                             */
int bhnerf_lightcurve(const float* images, int32_t Bt, int32_t S, int32_t P, float* lc, void* stream);
int bhnerf_loss_lightcurve(const float* lc, const float* target, const float* sigma, const float* offset,
                           float loss_scale, int32_t Bt, int32_t S, int32_t P, float* loss, float* d_images,
                           void* stream);

/* ---- accounting for benchmarks: kernels launched by this library, and (between begin/end) CUDA-event
 * time per category {0 render fwd, 1 render bwd, 2 wgrad (two-kernel backward), 3 heads (ray integral, image losses,
 * d loss/d o), 4 misc, 5 collectives (bhnerf_allreduce_*), 6 visibility head}.
 * profile_end synchronises the device.  ms_host/scopes_host/launches_host: host arrays of BHNERF_N_CATEGORIES.   */
#define BHNERF_N_CATEGORIES 7
int64_t bhnerf_launch_count(void);
int bhnerf_profile_begin(void);
int bhnerf_profile_end(double* ms_host, int64_t* scopes_host, int64_t* launches_host);

/* ---- health flags of the tcgen05 steps that used `workspace` (kept in its first 256 bytes).  The flags are STICKY: one
 * raised by any step since the previous call is reported, then cleared by this call.  (The current step's own copy, words
 * [0..5) of the workspace, is what the guarded Adam update looks at.)  Synchronises `stream`.  flags_host[8]:
 *   [0] forward pipeline aborted (a bounded mbarrier wait expired)   [1] dgrad chain aborted   [2] wgrad aborted
 *   [3] forward: |activation| exceeded the fp16 operand range (65504) -> images invalid
 *   [4] backward: non-finite parameter gradient (cotangent overflow)  [5..7] reserved (0)
 * The reference has no counterpart (XLA fp32 cannot overflow here); callers poll this off the hot path.        */
int bhnerf_workspace_status(void* workspace, int32_t* flags_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BHNERF_B200_H */

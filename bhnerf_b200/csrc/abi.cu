// extern "C" entry points (include/bhnerf_b200.h): argument checking, workspace carving, chunking.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[1024] = "";
void bh_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* bhnerf_last_error(void) { return g_err; }
extern "C" int bhnerf_version(void) { return BHNERF_ABI_VERSION; }

extern "C" int bhnerf_device_check(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  BH_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  BH_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  BH_REQUIRE(p.major == 10, "bhnerf_b200 is built for sm_100a only; device is sm_%d%d", p.major, p.minor);
  return 0;
}

// ---- launch accounting / event timing ----
#include <mutex>
#include <vector>
static std::mutex g_prof_mu;
static unsigned long long g_launches[BH_NCAT] = {0};
static bool g_prof_on = false;
struct ProfRec { int cat; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
void bh_prof_begin(int cat, int n_launches, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_launches[cat] += (unsigned long long)n_launches;
  if (!g_prof_on) return;
  ProfRec r; r.cat = cat; r.a = prof_event(); r.b = prof_event();
  cudaEventRecord(r.a, st);
  g_prof_recs.push_back(r);
}
void bh_prof_end(int cat, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_on || g_prof_recs.empty()) return;
  for (size_t i = g_prof_recs.size(); i-- > 0;)
    if (g_prof_recs[i].cat == cat) { cudaEventRecord(g_prof_recs[i].b, st); break; }
}
extern "C" int bhnerf_profile_begin(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
  g_prof_recs.clear();
  for (int c = 0; c < BH_NCAT; ++c) g_launches[c] = 0;
  g_prof_on = true;
  return 0;
}
// ms, scopes (timed launch groups), launches (kernels launched): BHNERF_N_CATEGORIES entries each, since profile_begin
extern "C" int bhnerf_profile_end(double* ms_host, int64_t* scopes_host, int64_t* launches_host) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  BH_CHECK_CUDA(cudaDeviceSynchronize());
  for (int c = 0; c < BH_NCAT; ++c) { if (ms_host) ms_host[c] = 0.0; if (scopes_host) scopes_host[c] = 0; }
  for (auto& r : g_prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      if (ms_host) ms_host[r.cat] += ms;
      if (scopes_host) scopes_host[r.cat] += 1;
    }
    g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b);
  }
  g_prof_recs.clear();
  if (launches_host) for (int c = 0; c < BH_NCAT; ++c) launches_host[c] = (int64_t)g_launches[c];
  g_prof_on = false;
  return 0;
}
extern "C" int64_t bhnerf_launch_count(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  unsigned long long t = 0;
  for (int c = 0; c < BH_NCAT; ++c) t += g_launches[c];
  return (int64_t)t;
}

extern "C" int bhnerf_workspace_status(void* workspace, int32_t* flags_host, void* stream) {
  BH_REQUIRE(workspace && flags_host, "workspace_status: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  // words [55] magic | [56..64) sticky flags (tc_common.cuh): read, then cleared, so a flag raised by ANY step since the
  // previous call is reported exactly once
  int32_t w[9];
  BH_CHECK_CUDA(cudaMemcpyAsync(w, (const char*)workspace + 55 * sizeof(int32_t), sizeof(w), cudaMemcpyDeviceToHost, st));
  BH_CHECK_CUDA(cudaStreamSynchronize(st));
  const bool used = w[0] == 0x62683230;                 // a workspace no tcgen05 step has touched yet reports all-clear
  for (int i = 0; i < 8; ++i) flags_host[i] = (used && i < 5) ? w[1 + i] : 0;
  if (used) BH_CHECK_CUDA(cudaMemsetAsync((char*)workspace + 56 * sizeof(int32_t), 0, 8 * sizeof(int32_t), st));
  return 0;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static FrameConsts frame_consts(const bhnerf_scene_t* sc) {
  FrameConsts fc; fc.t_start_obs = sc->t_start_obs; fc.GM_c3 = sc->GM_c3; fc.t_injection = sc->t_injection;
  fc.scale = sc->scale; return fc;
}
static int check_scene(const bhnerf_scene_t* sc, int Bt, int impl) {
  BH_REQUIRE(sc && sc->packed, "scene is NULL / not prepacked");
  BH_REQUIRE(sc->n_pad > 0 && sc->n_pad % 128 == 0 && sc->S >= 1 && sc->S <= 4, "scene: bad n_pad/S");
  BH_REQUIRE(Bt > 0, "Bt must be > 0");
  BH_REQUIRE(impl == BHNERF_IMPL_SIMT || impl == BHNERF_IMPL_TC, "unknown impl %d", impl);
  BH_REQUIRE(sc->GM_c3 > 0.f && sc->scale > 0.f, "scene: GM_c3 and scale must be > 0");
  return 0;
}

// `pl` = precision plan of the TC family for this step (bh_tc_planes of the WHOLE step's frame count)
static size_t acts_bytes_per_frame(const bhnerf_scene_t* sc, int impl, int pl) {
  return impl == BHNERF_IMPL_SIMT ? bh_simt_acts_floats_per_frame(sc->n_pad) * 4
                                  : bh_tc_acts_bytes_per_frame(sc->n_pad, pl);
}
extern "C" size_t bhnerf_acts_bytes(const bhnerf_scene_t* sc, int32_t Bt, int32_t impl) {
  return acts_bytes_per_frame(sc, impl, bh_tc_planes(sc->n_active, Bt)) * (size_t)Bt;
}

// fixed (frame-count independent) part of the backward workspace: weight images (+ the delta ring of the fused
// tcgen05 backward, one-plane plan)
static size_t bwd_fixed_bytes(int impl, int pl) {
  return impl == BHNERF_IMPL_SIMT ? align_up(BH_SIMT_WT_FLOATS * 4, 256)
                                  : align_up(bh_tc_ws_bytes(), 256) + align_up(bh_tc_delta_fixed_bytes(pl), 256);
}
// where the backward's delta scratch lives: the fixed ring if the plan has one, else `per_chunk`
static void* delta_ptr(char* ws, int impl, int pl, void* per_chunk) {
  return (impl == BHNERF_IMPL_TC && bh_tc_delta_fixed_bytes(pl) > 0) ? (void*)(ws + align_up(bh_tc_ws_bytes(), 256)) : per_chunk;
}
// per-frame scratch of the backward beyond saved activations
static size_t bwd_frame_bytes(const bhnerf_scene_t* sc, int impl, int pl) {
  return impl == BHNERF_IMPL_SIMT ? bh_simt_delta_floats_per_frame(sc->n_pad) * 4
                                  : bh_tc_delta_bytes_per_frame(sc->n_pad, pl) + (size_t)sc->n_pad * 4 /* d loss/d o */;
}

// TC variant needs scratch for the bf16 weight images; the SIMT variant ignores it.
extern "C" size_t bhnerf_fwd_workspace_bytes(int32_t impl) {
  return impl == BHNERF_IMPL_TC ? align_up(bh_tc_ws_bytes(), 256) : 0;
}
extern "C" int bhnerf_render_fwd(const bhnerf_scene_t* sc, const float* params, const float* t_frames,
                                    int32_t Bt, float* images, float* e_out, void* acts_out, void* workspace,
                                    size_t workspace_bytes, int32_t impl, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int r = check_scene(sc, Bt, impl)) return r;
  BH_REQUIRE(params && t_frames && images && e_out, "render_fwd: NULL argument (e_out is required)");
  PackedView v = bh_view(sc);
  FrameConsts fc = frame_consts(sc);
  if (impl == BHNERF_IMPL_SIMT) {
    if (int r = bh_simt_fwd(v, fc, params, t_frames, Bt, e_out, (float*)acts_out, st)) return r;
  } else {
    BH_REQUIRE(workspace && workspace_bytes >= bhnerf_fwd_workspace_bytes(impl), "render_fwd: workspace too small");
    if (int r = bh_tc_prepare_weights(params, workspace, st)) return r;
    if (int r = bh_tc_fwd(v, fc, workspace, params, t_frames, Bt, e_out, acts_out, bh_tc_planes(sc->n_active, Bt), st)) return r;
  }
  return bh_launch_ray_integrate(v, e_out, Bt, images, st);
}

extern "C" size_t bhnerf_bwd_workspace_bytes(const bhnerf_scene_t* sc, int32_t Bt, int32_t impl) {
  // enough for ONE frame of recompute; more lets the backward chunk more frames per launch
  const int pl = bh_tc_planes(sc->n_active, Bt);
  return bwd_fixed_bytes(impl, pl) + acts_bytes_per_frame(sc, impl, pl) + bwd_frame_bytes(sc, impl, pl) +
         (size_t)sc->n_pad * 4 + 1024;
}
// the frame-count independent part of bhnerf_bwd_workspace_bytes / bhnerf_train_workspace_bytes
extern "C" size_t bhnerf_bwd_fixed_workspace_bytes(const bhnerf_scene_t* sc, int32_t Bt, int32_t impl) {
  return bwd_fixed_bytes(impl, bh_tc_planes(sc->n_active, Bt));
}

extern "C" int bhnerf_render_bwd(const bhnerf_scene_t* sc, const float* params, const float* t_frames, int32_t Bt,
                                 const float* d_images, const float* e_saved, const void* acts_saved,
                                 float* d_params, void* workspace, size_t workspace_bytes, int32_t impl,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int r = check_scene(sc, Bt, impl)) return r;
  BH_REQUIRE(params && t_frames && d_images && d_params && workspace, "render_bwd: NULL argument");
  PackedView v = bh_view(sc);
  FrameConsts fc = frame_consts(sc);
  char* ws = (char*)workspace;
  const int pl = bh_tc_planes(sc->n_active, Bt);
  size_t fixed = bwd_fixed_bytes(impl, pl);
  BH_REQUIRE(workspace_bytes >= fixed, "render_bwd: workspace too small");
  char* p = ws + fixed;
  size_t avail = workspace_bytes - fixed;
  // the forward is recomputed unless BOTH residuals were passed in; the recompute carves activations AND e out of the
  // workspace, so both are counted whenever it runs (a mixed call used to overrun a capped workspace)
  const bool recompute = (acts_saved == nullptr) || (e_saved == nullptr);
  size_t per_frame = bwd_frame_bytes(sc, impl, pl) +
                     (recompute ? acts_bytes_per_frame(sc, impl, pl) + (size_t)sc->n_pad * 4 : 0);
  int Bc = per_frame ? (int)(avail / per_frame) : Bt;
  if (Bc > Bt) Bc = Bt;
  if (Bc >= 1) { const int nchunks = (Bt + Bc - 1) / Bc; Bc = (Bt + nchunks - 1) / nchunks; }   // equal chunks
  BH_REQUIRE(Bc >= 1, "render_bwd: workspace (%zu B) cannot hold one frame (%zu B + %zu B fixed)",
             workspace_bytes, per_frame, fixed);
  BH_CHECK_CUDA(cudaMemsetAsync(d_params, 0, BHNERF_N_PARAMS * sizeof(float), st));
  if (impl == BHNERF_IMPL_TC) { if (int r = bh_tc_prepare_weights(params, ws, st)) return r; }
  for (int b0 = 0; b0 < Bt; b0 += Bc) {
    int nb = (Bt - b0 < Bc) ? Bt - b0 : Bc;
    char* q = p;
    float* delta = (float*)delta_ptr(ws, impl, pl, q); q += bwd_frame_bytes(sc, impl, pl) * nb;
    float* dout = (float*)(q - (size_t)sc->n_pad * 4 * nb);       // tail of the per-frame scratch (TC family)
    const void* acts = acts_saved ? (const char*)acts_saved + acts_bytes_per_frame(sc, impl, pl) * b0 : nullptr;
    const float* e = e_saved ? e_saved + (size_t)b0 * sc->n_pad : nullptr;
    if (recompute) {
      void* acts_ws = q; q += acts_bytes_per_frame(sc, impl, pl) * nb;
      float* e_ws = (float*)q; q += (size_t)sc->n_pad * 4 * nb;
      BH_REQUIRE((size_t)(q - ws) <= workspace_bytes, "render_bwd: internal workspace accounting error");
      if (impl == BHNERF_IMPL_SIMT) {
        if (int r = bh_simt_fwd(v, fc, params, t_frames + b0, nb, e_ws, (float*)acts_ws, st)) return r;
      } else {
        if (int r = bh_tc_fwd(v, fc, ws, params, t_frames + b0, nb, e_ws, acts_ws, pl, st)) return r;
      }
      acts = acts_ws; e = e_ws;
    }
    const float* dI = d_images + (size_t)b0 * sc->S * sc->P;
    if (impl == BHNERF_IMPL_SIMT) {
      if (int r = bh_simt_bwd(v, params, dI, nb, e, (const float*)acts, delta, (float*)ws, d_params, st)) return r;
    } else {
      if (int r = bh_tc_bwd(v, ws, params, dI, nb, e, acts, delta, dout, pl, d_params, st)) return r;
    }
  }
  return 0;
}

// ---- fused train step for separable image losses ----
static size_t train_frame_bytes(const bhnerf_scene_t* sc, int impl, int pl) {
  return acts_bytes_per_frame(sc, impl, pl) + bwd_frame_bytes(sc, impl, pl) + (size_t)sc->n_pad * 4 +
         (size_t)sc->S * sc->P * 4;
}
extern "C" size_t bhnerf_train_workspace_bytes(const bhnerf_scene_t* sc, int32_t Bt, int32_t impl) {
  // all Bt frames in one chunk; smaller workspaces are accepted down to one frame
  // large enough for either precision plan the step may pick (it depends on the loss kind, bhnerf_train_step_image)
  size_t need = 0;
  const int plans[2] = {bh_tc_planes(sc->n_active, Bt), bh_tc_planes_full_loss(sc->n_active, Bt)};
  for (int pl : plans) {
    size_t b = bwd_fixed_bytes(impl, pl) + train_frame_bytes(sc, impl, pl) * (size_t)Bt + 1024;
    if (b > need) need = b;
  }
  return need;
}

int bh_loss_image_accum(const float* images, const float* target, const float* sigma, const float* offset,
                        float loss_scale, int kind, int Bt, int S, int P, float* loss, float* d_images,
                        cudaStream_t st);

extern "C" int bhnerf_train_step_image(const bhnerf_scene_t* sc, const float* params, const float* t_frames,
                                       int32_t Bt, const float* target, const float* sigma, const float* offset,
                                       float loss_scale, int32_t kind, float* loss, float* images, float* d_params,
                                       void* workspace, size_t workspace_bytes, int32_t impl, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int r = check_scene(sc, Bt, impl)) return r;
  BH_REQUIRE(params && t_frames && target && sigma && offset && loss && images && d_params && workspace,
             "train_step_image: NULL argument");
  BH_REQUIRE(kind == BHNERF_LOSS_FULL || kind == BHNERF_LOSS_LC, "image dtype (%d) not supported", kind);
  PackedView v = bh_view(sc);
  FrameConsts fc = frame_consts(sc);
  char* ws = (char*)workspace;
  // the fused step owns its workspace layout, so its precision plan may depend on the loss (render_tc.cu)
  const int pl = kind == BHNERF_LOSS_FULL ? bh_tc_planes_full_loss(sc->n_active, Bt) : bh_tc_planes(sc->n_active, Bt);
  size_t fixed = bwd_fixed_bytes(impl, pl);
  size_t per_frame = train_frame_bytes(sc, impl, pl);
  BH_REQUIRE(workspace_bytes >= fixed + per_frame, "train_step_image: workspace (%zu B) cannot hold one frame (%zu B)",
             workspace_bytes, fixed + per_frame);
  int Bc = (int)((workspace_bytes - fixed) / per_frame);
  if (Bc > Bt) Bc = Bt;
  { const int nchunks = (Bt + Bc - 1) / Bc; Bc = (Bt + nchunks - 1) / nchunks; }      // equal chunks instead of a short tail
  BH_CHECK_CUDA(cudaMemsetAsync(d_params, 0, BHNERF_N_PARAMS * sizeof(float), st));
  BH_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  if (impl == BHNERF_IMPL_TC) { if (int r = bh_tc_prepare_weights(params, ws, st)) return r; }
  size_t tstride = (kind == BHNERF_LOSS_FULL) ? (size_t)sc->S * sc->P : (size_t)sc->S;
  for (int b0 = 0; b0 < Bt; b0 += Bc) {
    int nb = (Bt - b0 < Bc) ? Bt - b0 : Bc;
    char* q = ws + fixed;
    void* acts = q; q += acts_bytes_per_frame(sc, impl, pl) * nb;
    float* delta = (float*)delta_ptr(ws, impl, pl, q); q += bwd_frame_bytes(sc, impl, pl) * nb;
    float* dout = (float*)(q - (size_t)sc->n_pad * 4 * nb);
    float* e = (float*)q; q += (size_t)sc->n_pad * 4 * nb;
    float* dI = (float*)q;
    float* img = images + (size_t)b0 * sc->S * sc->P;
    if (impl == BHNERF_IMPL_SIMT) {
      if (int r = bh_simt_fwd(v, fc, params, t_frames + b0, nb, e, (float*)acts, st)) return r;
    } else {
      if (int r = bh_tc_fwd(v, fc, ws, params, t_frames + b0, nb, e, acts, pl, st)) return r;
    }
    if (int r = bh_launch_ray_integrate(v, e, nb, img, st)) return r;
    if (int r = bh_loss_image_accum(img, target + b0 * tstride, sigma + b0 * tstride, offset + b0 * tstride,
                                    loss_scale, kind, nb, sc->S, sc->P, loss, dI, st)) return r;
    if (impl == BHNERF_IMPL_SIMT) {
      if (int r = bh_simt_bwd(v, params, dI, nb, e, (const float*)acts, delta, (float*)ws, d_params, st)) return r;
    } else {
      if (int r = bh_tc_bwd(v, ws, params, dI, nb, e, acts, delta, dout, pl, d_params, st)) return r;
    }
  }
  return 0;
}

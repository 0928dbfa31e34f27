// XLA-FFI shim over the C ABI (include/bhnerf_b200.h): the jax.ffi custom-call targets a bhnerf
// maintainer registers so that bhnerf/network.py keeps its JAX API while the render / train step
// runs in libbhnerf_b200.so.  NOT compiled in this repository's image (jaxlib, and therefore
// xla/ffi/api/ffi.h, is not installed -- SURVEY.md s0.4); build it next to a JAX install with
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -Iinclude integration/xla_ffi_shim.cc -Lbhnerf_b200/lib -lbhnerf_b200 -o libbhnerf_xla.so
// All logic stays below the C ABI; a handler only unpacks buffers, picks the stream and forwards.
#include "xla/ffi/api/ffi.h"

#include "bhnerf_b200.h"

namespace ffi = xla::ffi;

static ffi::Error to_error(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, bhnerf_last_error());
}

// The prepacked scene (frame-independent part of network.raytracing_args, bhnerf/network.py:850-894) travels through XLA
// as an opaque uint8 buffer (the packed arrays) plus TYPED scalar attributes, one per field of bhnerf_scene_t
// (bhnerf_b200/jax_scene.py: ATTRS); bhnerf_prepack is called once per raytracing_args from Python.
struct SceneAttrs {
  int32_t n_active, n_pad, P, G, S;
  float t_start_obs, GM_c3, t_injection, scale;
};
static bhnerf_scene_t scene_of(ffi::AnyBuffer packed, const SceneAttrs& a) {
  bhnerf_scene_t sc;
  sc.packed = packed.untyped_data();
  sc.n_active = a.n_active; sc.n_pad = a.n_pad; sc.P = a.P; sc.G = a.G; sc.S = a.S;
  sc.t_start_obs = a.t_start_obs; sc.GM_c3 = a.GM_c3; sc.t_injection = a.t_injection; sc.scale = a.scale;
  return sc;
}
#define BHNERF_SCENE_ATTR_PARAMS                                                                                   \
  int32_t n_active, int32_t n_pad, int32_t P, int32_t G, int32_t S, float t_start_obs, float GM_c3, float t_injection, \
      float scale
#define BHNERF_SCENE_ATTR_VALUES SceneAttrs{n_active, n_pad, P, G, S, t_start_obs, GM_c3, t_injection, scale}
#define BHNERF_SCENE_ATTR_BINDINGS                                                                                  \
  .Attr<int32_t>("n_active").Attr<int32_t>("n_pad").Attr<int32_t>("P").Attr<int32_t>("G").Attr<int32_t>("S")         \
      .Attr<float>("t_start_obs").Attr<float>("GM_c3").Attr<float>("t_injection").Attr<float>("scale")

// images[Bt,S,P], e[Bt,n_pad] = render(packed scene, params[55169], t_frames[Bt])
// replaces network.image_plane_prediction (bhnerf/network.py:373-420)
static ffi::Error RenderFwdImpl(cudaStream_t stream, ffi::AnyBuffer packed, ffi::Buffer<ffi::F32> params,
                                ffi::Buffer<ffi::F32> t_frames, BHNERF_SCENE_ATTR_PARAMS,
                                ffi::ResultBuffer<ffi::F32> images, ffi::ResultBuffer<ffi::F32> e,
                                ffi::ResultBuffer<ffi::U8> acts, ffi::ResultBuffer<ffi::U8> workspace) {
  bhnerf_scene_t sc = scene_of(packed, BHNERF_SCENE_ATTR_VALUES);
  const int32_t Bt = static_cast<int32_t>(t_frames.element_count());
  return to_error(bhnerf_render_fwd(&sc, params.typed_data(), t_frames.typed_data(), Bt, images->typed_data(),
                                    e->typed_data(), acts->element_count() ? acts->untyped_data() : nullptr,
                                    workspace->untyped_data(), workspace->element_count(), BHNERF_IMPL_TC, stream));
}

// d_params[55169] = pull-back of d_images through the render (jax.value_and_grad, bhnerf/network.py:617,:677)
static ffi::Error RenderBwdImpl(cudaStream_t stream, ffi::AnyBuffer packed, ffi::Buffer<ffi::F32> params,
                                ffi::Buffer<ffi::F32> t_frames, ffi::Buffer<ffi::F32> d_images,
                                ffi::Buffer<ffi::F32> e, ffi::Buffer<ffi::U8> acts, BHNERF_SCENE_ATTR_PARAMS,
                                ffi::ResultBuffer<ffi::F32> d_params, ffi::ResultBuffer<ffi::U8> workspace) {
  bhnerf_scene_t sc = scene_of(packed, BHNERF_SCENE_ATTR_VALUES);
  const int32_t Bt = static_cast<int32_t>(t_frames.element_count());
  return to_error(bhnerf_render_bwd(&sc, params.typed_data(), t_frames.typed_data(), Bt, d_images.typed_data(),
                                    e.typed_data(), acts.untyped_data(), d_params->typed_data(),
                                    workspace->untyped_data(), workspace->element_count(), BHNERF_IMPL_TC, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(BhnerfRenderFwd, RenderFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  BHNERF_SCENE_ATTR_BINDINGS
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(BhnerfRenderBwd, RenderBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  BHNERF_SCENE_ATTR_BINDINGS
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

// Ray integral, loss heads (image / lightcurve / visibility) and the Adam update.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// images[b,s,p] = sum_{i in ray p} e[b,i] * w[s,i]        (kgeo.radiative_trasfer, kgeo.py:618-621)
// one thread per (b,p); the CSR ranges of neighbouring rays are adjacent in memory. Deterministic.
// ---------------------------------------------------------------------------------------------
__global__ void ray_integrate_kernel(PackedView v, const float* __restrict__ e, int Bt,
                                     float* __restrict__ images) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  int b = blockIdx.y;
  if (p >= v.P) return;
  int i0 = v.row_ptr[p], i1 = v.row_ptr[p + 1];
  const float* eb = e + (size_t)b * v.n_pad;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = i0; i < i1; ++i) {
    float ev = eb[i];
    for (int s = 0; s < v.S; ++s) acc[s] += ev * v.w[(size_t)s * v.n_pad + i];
  }
  for (int s = 0; s < v.S; ++s) images[((size_t)b * v.S + s) * v.P + p] = acc[s];
}

int bh_launch_ray_integrate(const PackedView& v, const float* e, int Bt, float* images, cudaStream_t st) {
  BhProfScope ps(BH_CAT_HEADS, 1, st);
  dim3 grid((v.P + 127) / 128, Bt);
  ray_integrate_kernel<<<grid, 128, 0, st>>>(v, e, Bt, images);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// image losses (network.py:476-484)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce_sum(float v, float* sh) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (wid == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  __syncthreads();
  if (threadIdx.x == 0) sh[0] = v;
  __syncthreads();
  return sh[0];
}

// 'full': loss = scale * sum |(I - t - off)/sigma|^2 ; dI = 2*scale*(I - t - off)/sigma^2
__global__ void loss_full_kernel(const float* __restrict__ I, const float* __restrict__ t,
                                 const float* __restrict__ sg, const float* __restrict__ off, float scale,
                                 size_t n, float* __restrict__ loss, float* __restrict__ dI) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float r = (I[i] - t[i] - off[i]) / sg[i];
    acc += r * r;
    dI[i] = 2.f * scale * r / sg[i];
  }
  float tot = block_reduce_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, scale * tot);
}

// 'lc': lightcurve[b,s] = sum_p I[b,s,p]; loss = scale*sum |(lc - t - off)/sigma|^2; one block per (b,s)
__global__ void loss_lc_kernel(const float* __restrict__ I, const float* __restrict__ t,
                               const float* __restrict__ sg, const float* __restrict__ off, float scale,
                               int P, float* __restrict__ loss, float* __restrict__ dI) {
  __shared__ float sh[32];
  int bs = blockIdx.x;
  const float* row = I + (size_t)bs * P;
  float acc = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) acc += row[p];
  float lc = block_reduce_sum(acc, sh);
  float r = (lc - t[bs] - off[bs]) / sg[bs];
  float d = 2.f * scale * r / sg[bs];
  for (int p = threadIdx.x; p < P; p += blockDim.x) dI[(size_t)bs * P + p] = d;
  if (threadIdx.x == 0) atomicAdd(loss, scale * r * r);
}

// the two halves of the 'lc' head, for ray-sharded ranks (partial lightcurves are all-reduced in between)
__global__ void lightcurve_kernel(const float* __restrict__ I, int P, float* __restrict__ lc) {
  __shared__ float sh[32];
  const float* row = I + (size_t)blockIdx.x * P;
  float acc = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) acc += row[p];
  float tot = block_reduce_sum(acc, sh);
  if (threadIdx.x == 0) lc[blockIdx.x] = tot;
}
__global__ void loss_from_lightcurve_kernel(const float* __restrict__ lc, const float* __restrict__ t,
                                            const float* __restrict__ sg, const float* __restrict__ off, float scale,
                                            int P, float* __restrict__ loss, float* __restrict__ dI) {
  int bs = blockIdx.x;
  float r = (lc[bs] - t[bs] - off[bs]) / sg[bs];
  float d = 2.f * scale * r / sg[bs];
  for (int p = threadIdx.x; p < P; p += blockDim.x) dI[(size_t)bs * P + p] = d;
  if (threadIdx.x == 0) atomicAdd(loss, scale * r * r);
}
extern "C" int bhnerf_lightcurve(const float* images, int32_t Bt, int32_t S, int32_t P, float* lc, void* stream) {
  BH_REQUIRE(images && lc && Bt > 0 && S > 0 && P > 0, "lightcurve: bad argument");
  BhProfScope ps(BH_CAT_HEADS, 1, (cudaStream_t)stream);
  lightcurve_kernel<<<Bt * S, 256, 0, (cudaStream_t)stream>>>(images, P, lc);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int bhnerf_loss_lightcurve(const float* lc, const float* target, const float* sigma, const float* offset,
                                      float loss_scale, int32_t Bt, int32_t S, int32_t P, float* loss, float* d_images,
                                      void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(lc && target && sigma && offset && loss && d_images && Bt > 0 && S > 0 && P > 0, "loss_lightcurve: bad argument");
  BH_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  BhProfScope ps(BH_CAT_HEADS, 1, st);
  loss_from_lightcurve_kernel<<<Bt * S, 256, 0, st>>>(lc, target, sigma, offset, loss_scale, P, loss, d_images);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// accumulates into *loss (caller zeroes it)
int bh_loss_image_accum(const float* images, const float* target, const float* sigma, const float* offset,
                        float loss_scale, int kind, int Bt, int S, int P, float* loss, float* d_images,
                        cudaStream_t st) {
  BhProfScope ps(BH_CAT_HEADS, 1, st);
  if (kind == BHNERF_LOSS_FULL) {
    size_t n = (size_t)Bt * S * P;
    int blocks = (int)((n + 255) / 256); if (blocks > 1184) blocks = 1184;
    loss_full_kernel<<<blocks, 256, 0, st>>>(images, target, sigma, offset, loss_scale, n, loss, d_images);
  } else {
    loss_lc_kernel<<<Bt * S, 256, 0, st>>>(images, target, sigma, offset, loss_scale, P, loss, d_images);
  }
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bhnerf_loss_image(const float* images, const float* target, const float* sigma,
                                 const float* offset, float loss_scale, int32_t kind, int32_t Bt, int32_t S,
                                 int32_t P, float* loss, float* d_images, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(kind == BHNERF_LOSS_FULL || kind == BHNERF_LOSS_LC, "loss_image: image dtype (%d) not supported", kind);
  BH_REQUIRE(images && target && sigma && offset && loss && d_images, "loss_image: NULL argument");
  BH_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  return bh_loss_image_accum(images, target, sigma, offset, loss_scale, kind, Bt, S, P, loss, d_images, st);
}

// ---------------------------------------------------------------------------------------------
// visibility head (network.py:542-564): vis[b,v] = sum_p A[b,v,p] * I[b,p], complex64 A -- a DIFFERENT DFT matrix per
// frame (uv coverage rotates with the Earth), so this is a batched GEMV with no operand reuse: the bound is the HBM
// stream of A, 8*V*P bytes per frame per pass.  Forward: one block per (b,v) row, 16-byte loads, 4 independent loads in
// flight per thread, deterministic block reduction.  Backward: d_images[b,p] = sum_v Re(conj(A[b,v,p]) d_vis[b,v]); the
// rows are split into chunks so that a frame gives 8x more blocks than P/512 (fp32 atomics into the zeroed d_images).
// bhnerf_vis_head runs forward -> chi^2 -> backward per GROUP of frames (default: the whole batch; optionally groups small
// enough for the group's A to stay in the 126 MB L2 for the backward pass).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vis_fwd_kernel(const float2* __restrict__ A, const float* __restrict__ I, int V, int P,
                                                      float2* __restrict__ vis) {
  __shared__ float sh[32];
  const int v = blockIdx.x, b = blockIdx.y;
  const float2* row = A + ((size_t)b * V + v) * P;
  const float* img = I + (size_t)b * P;
  float re = 0.f, im = 0.f;
  if ((P & 1) == 0) {
    const float4* row4 = (const float4*)row;
    const float2* img2 = (const float2*)img;
    const int n = P / 2;
    int q = threadIdx.x;
    for (; q + 3 * 256 < n; q += 4 * 256) {
      float4 a0 = __ldcs(row4 + q), a1 = __ldcs(row4 + q + 256), a2 = __ldcs(row4 + q + 512), a3 = __ldcs(row4 + q + 768);
      float2 x0 = img2[q], x1 = img2[q + 256], x2 = img2[q + 512], x3 = img2[q + 768];
      re += a0.x * x0.x + a0.z * x0.y + a1.x * x1.x + a1.z * x1.y + a2.x * x2.x + a2.z * x2.y + a3.x * x3.x + a3.z * x3.y;
      im += a0.y * x0.x + a0.w * x0.y + a1.y * x1.x + a1.w * x1.y + a2.y * x2.x + a2.w * x2.y + a3.y * x3.x + a3.w * x3.y;
    }
    for (; q < n; q += 256) {
      float4 a = __ldcs(row4 + q);
      float2 x = img2[q];
      re += a.x * x.x + a.z * x.y;
      im += a.y * x.x + a.w * x.y;
    }
  } else {
    for (int p = threadIdx.x; p < P; p += 256) { float2 a = row[p]; float x = img[p]; re += a.x * x; im += a.y * x; }
  }
  re = block_reduce_sum(re, sh);
  im = block_reduce_sum(im, sh);
  if (threadIdx.x == 0) vis[(size_t)b * V + v] = make_float2(re, im);
}

#define VIS_BWD_ROWS 24       // rows of A per block of the backward
__global__ void __launch_bounds__(256) vis_bwd_kernel(const float2* __restrict__ A, const float2* __restrict__ dvis, int V, int P,
                                                      float* __restrict__ dI) {
  __shared__ float2 dv_s[VIS_BWD_ROWS];
  const int b = blockIdx.z, v0 = blockIdx.y * VIS_BWD_ROWS, nv = min(VIS_BWD_ROWS, V - v0);
  if (threadIdx.x < nv) dv_s[threadIdx.x] = dvis[(size_t)b * V + v0 + threadIdx.x];
  __syncthreads();
  const int p = (blockIdx.x * 256 + threadIdx.x) * 2;          // two pixels per thread: 16-byte loads
  if (p >= P) return;
  const float2* col = A + ((size_t)b * V + v0) * P + p;
  float acc0 = 0.f, acc1 = 0.f;
  if (p + 1 < P && (P & 1) == 0) {
#pragma unroll 4
    for (int v = 0; v < nv; ++v) {
      const float4 a = __ldg((const float4*)(col + (size_t)v * P));       // second pass over A: expected in L2
      acc0 += a.x * dv_s[v].x + a.y * dv_s[v].y;
      acc1 += a.z * dv_s[v].x + a.w * dv_s[v].y;
    }
    atomicAdd(dI + (size_t)b * P + p, acc0);
    atomicAdd(dI + (size_t)b * P + p + 1, acc1);
  } else {
    for (int v = 0; v < nv; ++v) {
      const float2 a = col[(size_t)v * P];
      acc0 += a.x * dv_s[v].x + a.y * dv_s[v].y;
      if (p + 1 < P) { const float2 a1 = col[(size_t)v * P + 1]; acc1 += a1.x * dv_s[v].x + a1.y * dv_s[v].y; }
    }
    atomicAdd(dI + (size_t)b * P + p, acc0);
    if (p + 1 < P) atomicAdd(dI + (size_t)b * P + p + 1, acc1);
  }
}

// 'vis': chisq = sum (|vis - t|/sigma)^2 ; 'amp': chisq = sum |(|vis| - t)/sigma|^2
__global__ void loss_vis_kernel(const float2* __restrict__ vis, const float* __restrict__ target,
                                const float* __restrict__ sg, float scale, int kind, int n,
                                float* __restrict__ loss, float2* __restrict__ dvis) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float2 v = vis[i];
    float s2 = sg[i] * sg[i];
    if (kind == BHNERF_LOSS_VIS) {
      float2 t = ((const float2*)target)[i];
      float dr = v.x - t.x, di = v.y - t.y;
      acc += (dr * dr + di * di) / s2;
      dvis[i] = make_float2(2.f * scale * dr / s2, 2.f * scale * di / s2);
    } else {
      float amp = sqrtf(v.x * v.x + v.y * v.y);
      float r = amp - target[i];
      acc += r * r / s2;
      float k = (amp > 0.f) ? 2.f * scale * r / (s2 * amp) : 0.f;
      dvis[i] = make_float2(k * v.x, k * v.y);
    }
  }
  float tot = block_reduce_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, scale * tot);
}

// 'cphase' (network.py:555-559): closure phase = angle(v1*v2*v3) over the baseline-triangle axis of vis [Bt,3,V];
// chisq = sum (1 - cos(target - cphase)) / sigma^2.  d cphase / d v_k = (-im_k, re_k) / |v_k|^2.
__global__ void loss_cphase_kernel(const float2* __restrict__ vis, const float* __restrict__ target,
                                   const float* __restrict__ sg, float scale, int V, int n,
                                   float* __restrict__ loss, float2* __restrict__ dvis) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int b = i / V, v = i - b * V;
    const size_t base = (size_t)b * 3 * V + v;
    const float2 v1 = vis[base], v2 = vis[base + V], v3 = vis[base + 2 * (size_t)V];
    const float pr = v1.x * v2.x - v1.y * v2.y, pi = v1.x * v2.y + v1.y * v2.x;
    const float br = pr * v3.x - pi * v3.y, bi = pr * v3.y + pi * v3.x;
    const float phi = atan2f(bi, br);
    const float s2 = sg[i] * sg[i];
    float sn, cs;
    sincosf(target[i] - phi, &sn, &cs);
    acc += (1.f - cs) / s2;
    const float gphi = -scale * sn / s2;                   // d loss / d cphase
    const float2 vv[3] = {v1, v2, v3};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float m2 = vv[k].x * vv[k].x + vv[k].y * vv[k].y;
      const float f = (m2 > 0.f) ? gphi / m2 : 0.f;
      dvis[base + (size_t)k * V] = make_float2(-f * vv[k].y, f * vv[k].x);
    }
  }
  float tot = block_reduce_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, scale * tot);
}

// dst += src (chunked steps of the host mirror accumulate per-chunk gradients / losses on the device)
__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
extern "C" int bhnerf_add_inplace(float* dst, const float* src, int32_t n, void* stream) {
  BH_REQUIRE(dst && src && n > 0, "add_inplace: bad argument");
  BhProfScope ps(BH_CAT_MISC, 1, (cudaStream_t)stream);
  add_inplace_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dst, src, n);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int launch_vis_fwd(const float* A, const float* images, int Bt, int V, int P, float* vis, cudaStream_t st) {
  BhProfScope ps(BH_CAT_VIS, 1, st);
  vis_fwd_kernel<<<dim3(V, Bt), 256, 0, st>>>((const float2*)A, images, V, P, (float2*)vis);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int bhnerf_vis_fwd(const float* A, const float* images, int32_t Bt, int32_t V, int32_t P, float* vis,
                              void* stream) {
  BH_REQUIRE(A && images && vis && Bt > 0 && V > 0 && P > 0 && Bt <= 65535, "vis_fwd: bad argument");
  return launch_vis_fwd(A, images, Bt, V, P, vis, (cudaStream_t)stream);
}

static int launch_loss_vis(const float* vis, const float* target, const float* sigma, float loss_scale, int kind, int Bt,
                           int V, float* loss, float* d_vis, cudaStream_t st);
extern "C" int bhnerf_loss_vis(const float* vis, const float* target, const float* sigma, float loss_scale,
                               int32_t kind, int32_t Bt, int32_t V, float* loss, float* d_vis, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(kind == BHNERF_LOSS_VIS || kind == BHNERF_LOSS_AMP || kind == BHNERF_LOSS_CPHASE,
             "loss_vis: eht dtype (%d) not supported", kind);
  BH_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  return launch_loss_vis(vis, target, sigma, loss_scale, kind, Bt, V, loss, d_vis, st);
}

static int launch_vis_bwd(const float* A, const float* d_vis, int Bt, int V, int P, float* d_images, cudaStream_t st) {
  BH_CHECK_CUDA(cudaMemsetAsync(d_images, 0, (size_t)Bt * P * sizeof(float), st));
  BhProfScope ps(BH_CAT_VIS, 1, st);
  dim3 grid((P / 2 + 255) / 256 + ((P & 1) ? 1 : 0), (V + VIS_BWD_ROWS - 1) / VIS_BWD_ROWS, Bt);
  vis_bwd_kernel<<<grid, 256, 0, st>>>((const float2*)A, (const float2*)d_vis, V, P, d_images);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int bhnerf_vis_bwd(const float* A, const float* d_vis, int32_t Bt, int32_t V, int32_t P,
                              float* d_images, void* stream) {
  BH_REQUIRE(A && d_vis && d_images && Bt > 0 && V > 0 && P > 0 && Bt <= 65535, "vis_bwd: bad argument");
  return launch_vis_bwd(A, d_vis, Bt, V, P, d_images, (cudaStream_t)stream);
}

static int launch_loss_vis(const float* vis, const float* target, const float* sigma, float loss_scale, int kind, int Bt,
                           int V, float* loss, float* d_vis, cudaStream_t st) {
  int n = Bt * V;
  int blocks = (n + 255) / 256; if (blocks > 592) blocks = 592;
  BhProfScope ps(BH_CAT_VIS, 1, st);
  if (kind == BHNERF_LOSS_CPHASE)
    loss_cphase_kernel<<<blocks, 256, 0, st>>>((const float2*)vis, target, sigma, loss_scale, V, n, loss, (float2*)d_vis);
  else
    loss_vis_kernel<<<blocks, 256, 0, st>>>((const float2*)vis, target, sigma, loss_scale, kind, n, loss, (float2*)d_vis);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// The whole eht head of a step in one call: per group of frames  vis = A I  ->  chi^2 (accumulated)  ->  d_images = A^H d_vis.
// rows = rows of A per frame (V for 'vis'/'amp', 3*ncphase for 'cphase'); target stride per frame is rows ('vis': complex)
// or rows/3 ('cphase').  group_bytes: how much of A one group may cover (0 = the whole batch in one group).  Measured on
// B200 at the cfg3 shape (scripts/vis_head_bench.py, profiles/r2_vis_head_bench.log): both passes stream A at the HBM
// roofline when the batch is one group (0.49 ms for 2 x 1.6 GB = 6.5 TB/s); groups small enough for the L2 to serve the
// backward pass (<= 48 MB) lose more to their 3 launches each than the saved HBM pass gains (1.08 ms), so one group is the
// default and the L2-sized grouping stays an option.
extern "C" int bhnerf_vis_head(const float* A, const float* images, const float* target, const float* sigma,
                               float loss_scale, int32_t kind, int32_t Bt, int32_t rows, int32_t P, float* loss,
                               float* vis, float* d_vis, float* d_images, size_t group_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(A && images && target && sigma && loss && vis && d_vis && Bt > 0 && rows > 0 && P > 0, "vis_head: bad argument");
  BH_REQUIRE(kind == BHNERF_LOSS_VIS || kind == BHNERF_LOSS_AMP || kind == BHNERF_LOSS_CPHASE,
             "vis_head: eht dtype (%d) not supported", kind);
  BH_REQUIRE(kind != BHNERF_LOSS_CPHASE || rows % 3 == 0, "vis_head: cphase needs rows = 3 * ncphase");
  const size_t frame_bytes = (size_t)rows * P * 8;
  int Bg = group_bytes == 0 ? Bt : (int)(group_bytes / frame_bytes); if (Bg < 1) Bg = 1; if (Bg > Bt) Bg = Bt;
  const int V = kind == BHNERF_LOSS_CPHASE ? rows / 3 : rows;              // targets per frame
  const size_t tstride = (size_t)V * (kind == BHNERF_LOSS_VIS ? 2 : 1);
  BH_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  for (int b0 = 0; b0 < Bt; b0 += Bg) {
    const int nb = Bt - b0 < Bg ? Bt - b0 : Bg;
    const float* Ag = A + (size_t)b0 * rows * P * 2;
    float* visg = vis + (size_t)b0 * rows * 2;
    float* dvisg = d_vis + (size_t)b0 * rows * 2;
    if (int r = launch_vis_fwd(Ag, images + (size_t)b0 * P, nb, rows, P, visg, st)) return r;
    if (int r = launch_loss_vis(visg, target + (size_t)b0 * tstride, sigma + (size_t)b0 * V, loss_scale, kind, nb, V, loss, dvisg, st)) return r;
    if (d_images) { if (int r = launch_vis_bwd(Ag, dvisg, nb, rows, P, d_images + (size_t)b0 * P, st)) return r; }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// optax.adam + polynomial_schedule(power=1) (network.py:173-174, :621)
// ---------------------------------------------------------------------------------------------
// `guard` (or NULL): the five health flags of the step that produced `g` (first words of the tcgen05 workspace).  If any
// is set the update is skipped -- a gradient from an overflowed / aborted step is never applied; the sticky copy of the
// flags makes the caller's next bhnerf_workspace_status poll raise.
__device__ __forceinline__ bool adam_guard_tripped(const int* guard) {
  return guard && (guard[0] | guard[1] | guard[2] | guard[3] | guard[4]);
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu,
                            float* __restrict__ nu, int n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2, float gscale, const int* __restrict__ guard) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || adam_guard_tripped(guard)) return;
  float gi = g[i] * gscale;
  float m = b1 * mu[i] + (1.f - b1) * gi;
  float v = b2 * nu[i] + (1.f - b2) * gi * gi;
  mu[i] = m; nu[i] = v;
  float mh = m / bc1, vh = v / bc2;
  p[i] = p[i] - lr * mh / (sqrtf(vh) + eps);
}

extern "C" int bhnerf_adam_step(float* params, const float* grads, float* mu, float* nu, int32_t n,
                                int32_t count, float lr_init, float lr_final, int32_t transition_steps,
                                float b1, float b2, float eps, float grad_scale, const int32_t* guard, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(transition_steps > 0, "adam: transition_steps must be > 0");
  int c = count < 0 ? 0 : (count > transition_steps ? transition_steps : count);
  double frac = 1.0 - (double)c / (double)transition_steps;
  float lr = (float)((double)(lr_init - lr_final) * frac + (double)lr_final);
  double t = (double)count + 1.0;
  float bc1 = (float)(1.0 - pow((double)b1, t)), bc2 = (float)(1.0 - pow((double)b2, t));
  BhProfScope ps(BH_CAT_MISC, 1, st);
  adam_kernel<<<(n + 255) / 256, 256, 0, st>>>(params, grads, mu, nu, n, lr, b1, b2, eps, bc1, bc2, grad_scale, guard);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Same update with the step counter in DEVICE memory, so that a whole train step (render -> loss -> gradient ->
// Adam) can be captured once in a CUDA graph and replayed: the schedule and the bias corrections are derived from
// *count_dev inside the kernel (double precision, one thread per block), and a second one-thread kernel advances it.
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu,
                                float* __restrict__ nu, int n, const int* __restrict__ count_dev, float lr_init,
                                float lr_final, int transition_steps, float b1, float b2, float eps, float gscale,
                                const int* __restrict__ guard) {
  __shared__ float sh[3];
  if (threadIdx.x == 0) {
    const int count = *count_dev;
    const int c = count < 0 ? 0 : (count > transition_steps ? transition_steps : count);
    const double frac = 1.0 - (double)c / (double)transition_steps;
    sh[0] = (float)((double)(lr_init - lr_final) * frac + (double)lr_final);
    const double t = (double)count + 1.0;
    sh[1] = (float)(1.0 - pow((double)b1, t));
    sh[2] = (float)(1.0 - pow((double)b2, t));
  }
  __syncthreads();
  const float lr = sh[0], bc1 = sh[1], bc2 = sh[2];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || adam_guard_tripped(guard)) return;
  float gi = g[i] * gscale;
  float m = b1 * mu[i] + (1.f - b1) * gi;
  float v = b2 * nu[i] + (1.f - b2) * gi * gi;
  mu[i] = m; nu[i] = v;
  float mh = m / bc1, vh = v / bc2;
  p[i] = p[i] - lr * mh / (sqrtf(vh) + eps);
}
__global__ void counter_inc_kernel(int* c) { *c += 1; }

extern "C" int bhnerf_adam_step_dev(float* params, const float* grads, float* mu, float* nu, int32_t n,
                                    int32_t* count_dev, float lr_init, float lr_final, int32_t transition_steps,
                                    float b1, float b2, float eps, float grad_scale, const int32_t* guard,
                                    void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(transition_steps > 0 && count_dev, "adam_dev: transition_steps must be > 0 and count_dev non-NULL");
  BhProfScope ps(BH_CAT_MISC, 2, st);
  adam_dev_kernel<<<(n + 255) / 256, 256, 0, st>>>(params, grads, mu, nu, n, count_dev, lr_init, lr_final, transition_steps,
                                                   b1, b2, eps, grad_scale, guard);
  counter_inc_kernel<<<1, 1, 0, st>>>(count_dev);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// stand-alone dense stages (the reference exposes them as public functions; inside the render
// kernels they are fused).  Elementwise / short reductions, HBM-bound.
// ---------------------------------------------------------------------------------------------
// emission.velocity_warp_coords (emission.py:143-211): out[b,i,:] = R_z(-theta) coords[:,i], NaN where t_M<0
__global__ void warp_coords_kernel(const float* __restrict__ coords, const float* __restrict__ Omega,
                                   const float* __restrict__ t_geos, size_t N, const float* __restrict__ t_frames,
                                   FrameConsts fc, float* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int b = blockIdx.y;
  if (i >= N) return;
  float tfc = bh_frame_time(t_frames[b], fc);
  float tM = __fsub_rn(__fadd_rn(tfc, t_geos[i]), fc.t_injection);
  float x = coords[i], y = coords[N + i], z = coords[2 * N + i];
  float* o = out + ((size_t)b * N + i) * 3;
  if (tM < 0.0f) { float nan = __int_as_float(0x7fc00000); o[0] = nan; o[1] = nan; o[2] = nan; return; }
  float sn, cs;
  sincosf(__fmul_rn(tM, Omega[i]), &sn, &cs);
  o[0] = x * cs + y * sn; o[1] = y * cs - x * sn; o[2] = z;
}
extern "C" int bhnerf_velocity_warp_coords(const float* coords, const float* Omega, const float* t_geos, int64_t N,
                                           const float* t_frames, int32_t Bt, float t_start_obs, float GM_c3,
                                           float t_injection, float* out, void* stream) {
  BH_REQUIRE(coords && Omega && t_geos && t_frames && out && N > 0 && Bt > 0 && GM_c3 > 0.f, "velocity_warp_coords: bad argument");
  FrameConsts fc; fc.t_start_obs = t_start_obs; fc.GM_c3 = GM_c3; fc.t_injection = t_injection; fc.scale = 1.f;
  dim3 grid((unsigned)((N + 255) / 256), Bt);
  BhProfScope ps(BH_CAT_MISC, 1, (cudaStream_t)stream);
  warp_coords_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(coords, Omega, t_geos, (size_t)N, t_frames, fc, out);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// emission.fill_unsupervised_emission (emission.py:343-374), in place on e[R,N] with coords[3,N]
__global__ void fill_unsupervised_kernel(float* __restrict__ e, const float* __restrict__ coords, size_t N,
                                         float rmin2, float rmax2, float zw, float fill) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float x = coords[i], y = coords[N + i], z = coords[2 * N + i];
  float r2 = x * x + y * y + z * z;
  if ((r2 < rmin2) || (r2 > rmax2) || (fabsf(z) > zw)) e[(size_t)blockIdx.y * N + i] = fill;
}
extern "C" int bhnerf_fill_unsupervised_emission(float* emission, const float* coords, int32_t R, int64_t N,
                                                 float rmin, float rmax, float z_width, float fill_value,
                                                 void* stream) {
  BH_REQUIRE(emission && coords && R > 0 && N > 0, "fill_unsupervised_emission: bad argument");
  dim3 grid((unsigned)((N + 255) / 256), R);
  BhProfScope ps(BH_CAT_MISC, 1, (cudaStream_t)stream);
  fill_unsupervised_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(emission, coords, (size_t)N, rmin * rmin,
                                                                   rmax * rmax, z_width, fill_value);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// kgeo.radiative_trasfer (kgeo.py:595-622): out[r,p] = sum_k g^2 * e[r,p,k] * dtau * Sigma; one warp per (r,p)
__global__ void radiative_transfer_kernel(const float* __restrict__ e, const float* __restrict__ g,
                                          const float* __restrict__ dtau, const float* __restrict__ Sigma,
                                          int P, int G, float* __restrict__ out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int r = blockIdx.y;
  if (warp >= P) return;
  const float* er = e + ((size_t)r * P + warp) * G;
  size_t base = (size_t)warp * G;
  float acc = 0.f;
  for (int k = lane; k < G; k += 32) { float gg = g[base + k]; acc += gg * gg * er[k] * dtau[base + k] * Sigma[base + k]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[(size_t)r * P + warp] = acc;
}
extern "C" int bhnerf_radiative_transfer(const float* emission, const float* g, const float* dtau, const float* Sigma,
                                         int32_t R, int32_t P, int32_t G, float* out, void* stream) {
  BH_REQUIRE(emission && g && dtau && Sigma && out && R > 0 && P > 0 && G > 0, "radiative_transfer: bad argument");
  dim3 grid((P * 32 + 255) / 256, R);
  BhProfScope ps(BH_CAT_MISC, 1, (cudaStream_t)stream);
  radiative_transfer_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(emission, g, dtau, Sigma, P, G, out);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// fp32 CUDA-core (FFMA) kernels for the render hot path: the on-device correctness reference for the
// tcgen05 family and the path for odd shapes.  One thread per (frame, active sample).
//   fwd   : warp -> posenc -> 4x128 relu MLP (+skip) -> sigmoid(o-10) -> masked emission e
//   chain : d e -> delta_3..delta_0 (dgrad), bias grads, dW4
//   wgrad : dW_l = in_l^T delta_l as a tiled SGEMM over saved activations
// Reference semantics: bhnerf/network.py:18-64 (MLP), :98-122 (posenc), :191-237 (predictor).
#include "common.cuh"

#define PITCH 129                                   // smem row pitch (floats): conflict-free both ways
#define SIMT_SMEM_FWD ((2 * 128 * PITCH + BH_NF * 128) * sizeof(float))


// out[j][s] = relu(bias[j] + sum_k in[k][s] * W[k][j]) for the thread's sample s; W rows are read
// with block-uniform float4 loads (L1 broadcast).
template <int K, bool RELU>
__device__ __forceinline__ void simt_layer(const float* __restrict__ W, const float* __restrict__ bias,
                                           const float* in, int in_pitch, const float* in2, int K2,
                                           int in2_pitch, float* out, int s) {
  for (int j0 = 0; j0 < 128; j0 += 16) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = __ldg(bias + j0 + i);
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float hk = in[k * in_pitch + s];
      const float4* w4 = (const float4*)(W + (size_t)k * 128 + j0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 w = __ldg(w4 + q);
        acc[4 * q + 0] = fmaf(hk, w.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(hk, w.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(hk, w.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(hk, w.w, acc[4 * q + 3]);
      }
    }
    for (int k = 0; k < K2; ++k) {                  // skip-connection rows (posenc), network.py:61
      float hk = in2[k * in2_pitch + s];
      const float4* w4 = (const float4*)(W + (size_t)(K + k) * 128 + j0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 w = __ldg(w4 + q);
        acc[4 * q + 0] = fmaf(hk, w.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(hk, w.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(hk, w.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(hk, w.w, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) out[(j0 + i) * PITCH + s] = RELU ? fmaxf(acc[i], 0.f) : acc[i];
  }
}

template <bool SAVE>
__global__ void __launch_bounds__(128, 1)
simt_fwd_kernel(PackedView v, FrameConsts fc, const float* __restrict__ params,
                const float* __restrict__ t_frames, float* __restrict__ e_out, float* __restrict__ acts) {
  extern __shared__ float smem[];
  float* bufA = smem; float* bufB = smem + 128 * PITCH; float* feat = smem + 2 * 128 * PITCH;
  int s = threadIdx.x, i = blockIdx.x * 128 + s, b = blockIdx.y;
  float tfc = bh_frame_time(t_frames[b], fc);
  float f[BH_NF];
  bool valid = bh_features(v.x[i], v.y[i], v.z[i], v.omega[i], v.tgeo[i], tfc, fc, f);
#pragma unroll
  for (int k = 0; k < BH_NF; ++k) feat[k * 128 + s] = f[k];
  size_t np = (size_t)v.n_pad;
  float* ab = SAVE ? acts + (size_t)b * bh_simt_acts_floats_per_frame(v.n_pad) : nullptr;
  if (SAVE) {
#pragma unroll
    for (int k = 0; k < BH_NF; ++k) ab[(size_t)(4 * 128 + k) * np + i] = f[k];
  }
  // every thread only touches its own column s of the smem buffers: no block barriers needed
  simt_layer<BH_NF, true>(params + OFF_W0, params + OFF_B0, feat, 128, nullptr, 0, 0, bufA, s);
  if (SAVE) for (int j = 0; j < 128; ++j) ab[(size_t)(0 * 128 + j) * np + i] = bufA[j * PITCH + s];
  simt_layer<128, true>(params + OFF_W1, params + OFF_B1, bufA, PITCH, nullptr, 0, 0, bufB, s);
  if (SAVE) for (int j = 0; j < 128; ++j) ab[(size_t)(1 * 128 + j) * np + i] = bufB[j * PITCH + s];
  simt_layer<128, true>(params + OFF_W2, params + OFF_B2, bufB, PITCH, nullptr, 0, 0, bufA, s);
  if (SAVE) for (int j = 0; j < 128; ++j) ab[(size_t)(2 * 128 + j) * np + i] = bufA[j * PITCH + s];
  simt_layer<128, true>(params + OFF_W3, params + OFF_B3, bufA, PITCH, feat, BH_NF, 128, bufB, s);
  if (SAVE) for (int j = 0; j < 128; ++j) ab[(size_t)(3 * 128 + j) * np + i] = bufB[j * PITCH + s];
  float o = __ldg(params + OFF_B4);
#pragma unroll 8
  for (int j = 0; j < 128; ++j) o = fmaf(bufB[j * PITCH + s], __ldg(params + OFF_W4 + j), o);
  float e = bh_sigmoid_m10(o);
  e_out[(size_t)b * np + i] = (valid && v.ray[i] >= 0) ? e : 0.f;   // network.py:232
}

int bh_simt_fwd(const PackedView& v, const FrameConsts& fc, const float* params, const float* t_frames, int Bt,
                float* e_out, float* acts, cudaStream_t st) {
  BhProfScope ps(BH_CAT_FWD, 1, st);
  dim3 grid(v.n_pad / 128, Bt);
  if (acts) {
    BH_CHECK_CUDA(cudaFuncSetAttribute(simt_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)SIMT_SMEM_FWD));
    simt_fwd_kernel<true><<<grid, 128, SIMT_SMEM_FWD, st>>>(v, fc, params, t_frames, e_out, acts);
  } else {
    BH_CHECK_CUDA(cudaFuncSetAttribute(simt_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)SIMT_SMEM_FWD));
    simt_fwd_kernel<false><<<grid, 128, SIMT_SMEM_FWD, st>>>(v, fc, params, t_frames, e_out, nullptr);
  }
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// backward chain.  delta_l = d loss / d (pre-activation of layer l), per sample.
// ---------------------------------------------------------------------------------------------
__global__ void simt_transpose_w_kernel(const float* __restrict__ params, float* __restrict__ wt) {
  // wt[l-1][n][k] = W_l[k][n] for l = 1,2,3 (hidden rows k<128 only)
  int l = blockIdx.y, idx = blockIdx.x * blockDim.x + threadIdx.x;   // idx = n*128 + k
  int n = idx >> 7, k = idx & 127;
  const int offs[3] = {OFF_W1, OFF_W2, OFF_W3};
  wt[(size_t)l * 16384 + idx] = params[offs[l] + k * 128 + n];
}

// sum over the 128 samples of row j of a [128][PITCH] smem tile; thread j owns row j
__device__ __forceinline__ float row_sum(const float* buf, int j) {
  float a = 0.f;
#pragma unroll 8
  for (int s = 0; s < 128; ++s) a += buf[j * PITCH + s];
  return a;
}

__global__ void __launch_bounds__(128, 1)
simt_chain_kernel(PackedView v, const float* __restrict__ params, const float* __restrict__ wt,
                  const float* __restrict__ d_images, const float* __restrict__ e_saved,
                  const float* __restrict__ acts, float* __restrict__ delta, float* __restrict__ d_params) {
  extern __shared__ float smem[];
  float* dA = smem; float* dB = smem + 128 * PITCH; float* dout_s = smem + 2 * 128 * PITCH;
  int s = threadIdx.x, i = blockIdx.x * 128 + s, b = blockIdx.y;
  size_t np = (size_t)v.n_pad;
  const float* ab = acts + (size_t)b * bh_simt_acts_floats_per_frame(v.n_pad);
  float* db = delta + (size_t)b * bh_simt_delta_floats_per_frame(v.n_pad);
  // d loss / d o = e(1-e) * sum_s dI[b,s,ray] * w[s,i]     (sigmoid', network.py:230; kgeo.py:621)
  int ray = v.ray[i];
  float e = e_saved[(size_t)b * np + i];
  float g = 0.f;
  if (ray >= 0)
    for (int c = 0; c < v.S; ++c) g += d_images[((size_t)b * v.S + c) * v.P + ray] * v.w[(size_t)c * np + i];
  float dout = g * e * (1.f - e);
  dout_s[s] = dout;
  // delta_3[j] = dout * W4[j] * (h3[j] > 0); stage h3 tile in dB for the dW4 reduction
  for (int j = 0; j < 128; ++j) {
    float h3 = ab[(size_t)(3 * 128 + j) * np + i];
    dB[j * PITCH + s] = h3;
    float d = (h3 > 0.f) ? dout * __ldg(params + OFF_W4 + j) : 0.f;
    dA[j * PITCH + s] = d;
    db[(size_t)(3 * 128 + j) * np + i] = d;
  }
  __syncthreads();
  {   // dW4[j] = sum_s h3[j][s]*dout[s]; db4 = sum_s dout[s]; db3[j] = sum_s delta_3[j][s]
    int j = s;
    float a = 0.f;
#pragma unroll 8
    for (int t = 0; t < 128; ++t) a = fmaf(dB[j * PITCH + t], dout_s[t], a);
    atomicAdd(d_params + OFF_W4 + j, a);
    atomicAdd(d_params + OFF_B3 + j, row_sum(dA, j));
    if (j == 0) { float t4 = 0.f; for (int t = 0; t < 128; ++t) t4 += dout_s[t]; atomicAdd(d_params + OFF_B4, t4); }
  }
  __syncthreads();
  float* cur = dA; float* nxt = dB;
  const int boffs[3] = {OFF_B0, OFF_B1, OFF_B2};
  for (int l = 3; l >= 1; --l) {
    // delta_{l-1}[k] = (sum_n delta_l[n] * W_l[k][n]) * (h_{l-1}[k] > 0)
    const float* wtl = wt + (size_t)(l - 1) * 16384;     // [n][k]
    for (int k0 = 0; k0 < 128; k0 += 16) {
      float acc[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] = 0.f;
#pragma unroll 4
      for (int n = 0; n < 128; ++n) {
        float dn = cur[n * PITCH + s];
        const float4* w4 = (const float4*)(wtl + (size_t)n * 128 + k0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 w = __ldg(w4 + q);
          acc[4 * q + 0] = fmaf(dn, w.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(dn, w.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(dn, w.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(dn, w.w, acc[4 * q + 3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        float h = ab[(size_t)((l - 1) * 128 + k0 + q) * np + i];
        float d = (h > 0.f) ? acc[q] : 0.f;
        nxt[(k0 + q) * PITCH + s] = d;
        db[(size_t)((l - 1) * 128 + k0 + q) * np + i] = d;
      }
    }
    __syncthreads();
    atomicAdd(d_params + boffs[l - 1] + s, row_sum(nxt, s));
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
}

// dW[k][n] += sum_i A[k][i] * B[n][i] over this CTA's sample chunk.  A: [KA][n_pad] (k-th input row),
// B: [128][n_pad] (delta rows).  256 threads; thread (ty,tx) owns k = ty*8..+7, n = tx + 16*j.
#define WG_CHUNK 4096
__global__ void __launch_bounds__(256, 2)
simt_wgrad_kernel(const float* __restrict__ A, int KA, const float* __restrict__ B, int n_pad,
                  size_t a_frame_stride, size_t b_frame_stride, float* __restrict__ dW) {
  __shared__ float As[128][33];
  __shared__ float Bs[128][33];
  int b = blockIdx.y;
  const float* Ab = A + (size_t)b * a_frame_stride;
  const float* Bb = B + (size_t)b * b_frame_stride;
  int i_begin = blockIdx.x * WG_CHUNK, i_end = min(i_begin + WG_CHUNK, n_pad);
  int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[a][c] = 0.f;
  for (int i0 = i_begin; i0 < i_end; i0 += 32) {
    // load 128 rows x 32 samples of each operand, coalesced along samples
    for (int r = threadIdx.x >> 5; r < 128; r += 8) {
      int c = threadIdx.x & 31;
      As[r][c] = (r < KA) ? Ab[(size_t)r * n_pad + i0 + c] : 0.f;
      Bs[r][c] = Bb[(size_t)r * n_pad + i0 + c];
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < 32; ++c) {
      float a[8], bb[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { a[q] = As[ty * 8 + q][c]; bb[q] = Bs[tx + 16 * q][c]; }
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[q][r] = fmaf(a[q], bb[r], acc[q][r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    int k = ty * 8 + q;
    if (k < KA) {
#pragma unroll
      for (int r = 0; r < 8; ++r) atomicAdd(dW + (size_t)k * 128 + tx + 16 * r, acc[q][r]);
    }
  }
}

int bh_simt_bwd(const PackedView& v, const float* params, const float* d_images, int Bt, const float* e_saved,
                const float* acts, float* delta_ws, float* wt_ws, float* d_params, cudaStream_t st) {
  { BhProfScope ps(BH_CAT_MISC, 1, st); simt_transpose_w_kernel<<<dim3(64, 3), 256, 0, st>>>(params, wt_ws); }
  size_t smem = (2 * 128 * PITCH + 128) * sizeof(float);
  BH_CHECK_CUDA(cudaFuncSetAttribute(simt_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(v.n_pad / 128, Bt);
  { BhProfScope ps(BH_CAT_BWD, 1, st);
    simt_chain_kernel<<<grid, 128, smem, st>>>(v, params, wt_ws, d_images, e_saved, acts, delta_ws, d_params); }
  size_t np = (size_t)v.n_pad;
  size_t afs = bh_simt_acts_floats_per_frame(v.n_pad), dfs = bh_simt_delta_floats_per_frame(v.n_pad);
  dim3 wg((v.n_pad + WG_CHUNK - 1) / WG_CHUNK, Bt);
  const float* feat = acts + 4 * 128 * np;
  BhProfScope ps(BH_CAT_WGRAD, 5, st);
  // layer 0: in = feat (21 rows);  layers 1,2: in = h_{l-1};  layer 3: in = [h2 | feat] (network.py:61)
  simt_wgrad_kernel<<<wg, 256, 0, st>>>(feat, BH_NF, delta_ws + 0 * 128 * np, v.n_pad, afs, dfs, d_params + OFF_W0);
  simt_wgrad_kernel<<<wg, 256, 0, st>>>(acts + 0 * 128 * np, 128, delta_ws + 1 * 128 * np, v.n_pad, afs, dfs, d_params + OFF_W1);
  simt_wgrad_kernel<<<wg, 256, 0, st>>>(acts + 1 * 128 * np, 128, delta_ws + 2 * 128 * np, v.n_pad, afs, dfs, d_params + OFF_W2);
  simt_wgrad_kernel<<<wg, 256, 0, st>>>(acts + 2 * 128 * np, 128, delta_ws + 3 * 128 * np, v.n_pad, afs, dfs, d_params + OFF_W3);
  simt_wgrad_kernel<<<wg, 256, 0, st>>>(feat, BH_NF, delta_ws + 3 * 128 * np, v.n_pad, afs, dfs, d_params + OFF_W3 + 128 * 128);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}

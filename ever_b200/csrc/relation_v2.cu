// FSRelationV2 extras (ever/module/fs_relation.py:76-163): the scene encoder's GroupNorm(32) + ReLU on the N x C scene
// vector (:86-96), the channel concatenation [r * p, p] (:156) as strided row copies, and Dropout2d (:101,:156) as a
// per-(image, channel) scale.  All HBM-/latency-bound helper kernels; the heavy parts of the module (1x1 convolutions, BN)
// run on the tensor-core / BatchNorm kernels.
#include "common.cuh"

namespace evb {

// x, y: [N, C] fp32 (the scene embedding after a 1x1 conv, spatial size 1): GroupNorm statistics run over the C/G channels
// of a group.  One thread per (n, g).  stat[n][g] = {mean, rstd}.  y = relu(gamma * xhat + beta).
__global__ void gn_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ stat, int N,
                                   int C, int G, float eps, int round_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * G) return;
  const int n = idx / G, g = idx % G, cg = C / G;
  const float* xr = x + (long long)n * C + g * cg;
  float s = 0.f;
  for (int j = 0; j < cg; ++j) s += xr[j];
  const float mean = s / cg;
  float q = 0.f;
  for (int j = 0; j < cg; ++j) { const float d = xr[j] - mean; q += d * d; }
  const float rstd = rsqrtf(q / cg + eps);
  stat[idx * 2] = mean;
  stat[idx * 2 + 1] = rstd;
  for (int j = 0; j < cg; ++j) {
    const int c = g * cg + j;
    const float v = fmaxf((xr[j] - mean) * rstd * gamma[c] + beta[c], 0.f);
    y[(long long)n * C + c] = round_out ? bf16_round(v) : v;   // a bf16 conv consumes it (autocast casts its input)
  }
}
// dx[n][c] from dy (masked by y > 0): dxhat = g * gamma; dx = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat * xhat))
__global__ void gn_relu_bwd_x_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                                     const float* __restrict__ gamma, const float* __restrict__ stat, float* __restrict__ dx,
                                     int N, int C, int G) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * G) return;
  const int n = idx / G, g = idx % G, cg = C / G;
  const float mean = stat[idx * 2], rstd = stat[idx * 2 + 1];
  const long long base = (long long)n * C + g * cg;
  float s1 = 0.f, s2 = 0.f;
  for (int j = 0; j < cg; ++j) {
    const float gg = y[base + j] > 0.f ? dy[base + j] : 0.f;
    const float dxh = gg * gamma[g * cg + j];
    const float xh = (x[base + j] - mean) * rstd;
    s1 += dxh;
    s2 += dxh * xh;
  }
  s1 /= cg;
  s2 /= cg;
  for (int j = 0; j < cg; ++j) {
    const float gg = y[base + j] > 0.f ? dy[base + j] : 0.f;
    const float dxh = gg * gamma[g * cg + j];
    const float xh = (x[base + j] - mean) * rstd;
    dx[base + j] = rstd * (dxh - s1 - xh * s2);
  }
}
// dgamma[c] (+)= sum_n g * xhat, dbeta[c] (+)= sum_n g  (fixed order over n)
__global__ void gn_relu_bwd_p_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                                     const float* __restrict__ stat, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     int N, int C, int G, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int g = c / (C / G);
  float sg = 0.f, sb = 0.f;
  for (int n = 0; n < N; ++n) {
    const long long i = (long long)n * C + c;
    const float gg = y[i] > 0.f ? dy[i] : 0.f;
    const float mean = stat[(n * G + g) * 2], rstd = stat[(n * G + g) * 2 + 1];
    sg += gg * (x[i] - mean) * rstd;
    sb += gg;
  }
  dgamma[c] = accumulate ? dgamma[c] + sg : sg;
  dbeta[c] = accumulate ? dbeta[c] + sb : sb;
}

// dst[r][0:cols] (+)= src[r][0:cols], bf16 rows with independent strides (cols % 8 == 0, 16-byte aligned rows)
__global__ void __launch_bounds__(256)
copy2d_bf16_kernel(const __nv_bfloat16* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst, int ldd, long long rows,
                   int cols8, int accumulate) {
  const long long total = rows * cols8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols8;
    const int c = (int)(i % cols8) * 8;
    bf16x8 v = *reinterpret_cast<const bf16x8*>(src + r * lds + c);
    if (accumulate) {
      float a[8], b[8];
      unpack8(v, a);
      unpack8(*reinterpret_cast<const bf16x8*>(dst + r * ldd + c), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
      v = pack8(a);
    }
    *reinterpret_cast<bf16x8*>(dst + r * ldd + c) = v;
  }
}

// y[n][hw][c] = bf16(x[n][hw][c] * m[n][c])   (Dropout2d: m in {0, bf16(1 / (1 - p))}; backward is the same map on dy)
__global__ void __launch_bounds__(256)
channel_scale_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ m, __nv_bfloat16* __restrict__ y, int N,
                     long long HW, int C) {
  const int c8 = C / 8;
  const long long total = (long long)N * HW * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    const long long n = i / (HW * c8);
    float v[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + i * 8), v);
    const float4 m0 = *reinterpret_cast<const float4*>(m + n * C + c), m1 = *reinterpret_cast<const float4*>(m + n * C + c + 4);
    v[0] *= m0.x; v[1] *= m0.y; v[2] *= m0.z; v[3] *= m0.w;
    v[4] *= m1.x; v[5] *= m1.y; v[6] *= m1.z; v[7] *= m1.w;
    *reinterpret_cast<bf16x8*>(y + i * 8) = pack8(v);
  }
}

// y = keep ? bf16(x * scale) : 0   (nn.Dropout: keep = the element survived, scale = 1 / (1 - p); backward: same map on dy)
__global__ void __launch_bounds__(256)
dropout_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ keep, float scale,
                     __nv_bfloat16* __restrict__ y, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float v[8], k[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + i * 8), v);
    unpack8(*reinterpret_cast<const bf16x8*>(keep + i * 8), k);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = k[j] != 0.f ? v[j] * scale : 0.f;
    *reinterpret_cast<bf16x8*>(y + i * 8) = pack8(v);
  }
}

}  // namespace evb

using namespace evb;
#define ST ((cudaStream_t)stream)
#define LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA)

extern "C" int evb_groupnorm_relu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stat, int N,
                                      int C, int G, float eps, int round_out, void* stream) {
  if (N < 1 || G < 1 || C % G) return EVB_ERR_ARG;
  gn_relu_fwd_kernel<<<(N * G + 127) / 128, 128, 0, ST>>>(x, gamma, beta, y, stat, N, C, G, eps, round_out);
  return LAUNCH_OK();
}
extern "C" int evb_groupnorm_relu_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* stat,
                                      float* dx, float* dgamma, float* dbeta, int N, int C, int G, int accumulate,
                                      void* stream) {
  if (N < 1 || G < 1 || C % G) return EVB_ERR_ARG;
  gn_relu_bwd_x_kernel<<<(N * G + 127) / 128, 128, 0, ST>>>(dy, x, y, gamma, stat, dx, N, C, G);
  if (dgamma && dbeta) gn_relu_bwd_p_kernel<<<(C + 127) / 128, 128, 0, ST>>>(dy, x, y, stat, dgamma, dbeta, N, C, G, accumulate);
  return LAUNCH_OK();
}
extern "C" int evb_copy2d_bf16(const void* src, int lds, void* dst, int ldd, long long rows, int cols, int accumulate,
                               void* stream) {
  if (cols % 8 || lds % 8 || ldd % 8 || rows < 0) return EVB_ERR_ARG;
  if (rows == 0 || cols == 0) return EVB_OK;
  long long b = (rows * (cols / 8) + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  copy2d_bf16_kernel<<<(int)b, 256, 0, ST>>>((const __nv_bfloat16*)src, lds, (__nv_bfloat16*)dst, ldd, rows, cols / 8,
                                             accumulate);
  return LAUNCH_OK();
}
extern "C" int evb_channel_scale(const void* x, const float* m, void* y, int N, long long HW, int C, void* stream) {
  if (C % 8 || N < 1 || HW < 1) return EVB_ERR_ARG;
  long long b = ((long long)N * HW * (C / 8) + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  channel_scale_kernel<<<(int)b, 256, 0, ST>>>((const __nv_bfloat16*)x, m, (__nv_bfloat16*)y, N, HW, C);
  return LAUNCH_OK();
}
extern "C" int evb_dropout_apply(const void* x, const void* keep, float scale, void* y, long long numel, void* stream) {
  if (numel % 8 || numel < 0) return EVB_ERR_ARG;
  if (numel == 0) return EVB_OK;
  long long b = (numel / 8 + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  dropout_apply_kernel<<<(int)b, 256, 0, ST>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)keep, scale,
                                               (__nv_bfloat16*)y, numel / 8);
  return LAUNCH_OK();
}

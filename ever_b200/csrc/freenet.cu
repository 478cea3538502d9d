// Small kernels of the FreeNet path (Z-Zheng/FreeNet: 3x3 conv + GroupNorm + ReLU blocks with squeeze-excitation, nearest
// top-down fusion).  The heavy parts run on the shared kernels: tcgen05 convolutions, evb_bn_stats (per-channel
// sums), evb_bn_apply (per-channel affine + ReLU), evb_norm_bwd_reduce / evb_norm_bwd_apply.  Here: the per-GROUP statistics
// that turn per-channel sums into GroupNorm's affine map and its backward constants, and the squeeze-excitation gate
// (ever/module/se_block.py:9-24).
#include "common.cuh"

namespace evb {

// One thread per group.  Inputs: per-channel mean_c / rstd_c of ONE sample as evb_bn_stats produced them with eps_bn
// (var_c = 1 / rstd_c^2 - eps_bn).  Group statistics over its Creal / G channels (each with the same pixel count):
// mu_g = mean_c(mean_c), E[x^2]_g = mean_c(var_c + mean_c^2).  Outputs per channel: scale = gamma * rstd_g,
// shift = beta - mu_g * scale (so y = x * scale + shift is GroupNorm), gmean = mu_g, grstd = rstd_g (for the backward).
// Channels >= Creal (zero padding to the tensor-core tile) get the identity map.
__global__ void gn_fold_kernel(const float* __restrict__ mean_c, const float* __restrict__ rstd_c, float eps_bn,
                               const float* __restrict__ gamma, const float* __restrict__ beta, int C, int Creal, int G,
                               float eps, float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ gmean,
                               float* __restrict__ grstd) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = Creal / G;
  if (g < G) {
    double m1 = 0.0, m2 = 0.0;
    for (int j = 0; j < cg; ++j) {
      const int c = g * cg + j;
      const double mu = mean_c[c], rs = rstd_c[c];
      double var = 1.0 / (rs * rs) - (double)eps_bn;
      if (var < 0) var = 0;
      m1 += mu;
      m2 += var + mu * mu;
    }
    m1 /= cg;
    m2 /= cg;
    double var_g = m2 - m1 * m1;
    if (var_g < 0) var_g = 0;
    const float rs_g = (float)(1.0 / sqrt(var_g + (double)eps));
    for (int j = 0; j < cg; ++j) {
      const int c = g * cg + j;
      const float a = gamma[c] * rs_g;
      scale[c] = a;
      shift[c] = beta[c] - (float)m1 * a;
      gmean[c] = (float)m1;
      grstd[c] = rs_g;
    }
  }
  if (g == 0)
    for (int c = Creal; c < C; ++c) { scale[c] = 1.f; shift[c] = 0.f; gmean[c] = 0.f; grstd[c] = 1.f; }
}

// GroupNorm backward constants from the per-channel sums dbeta_c = sum g, dgamma_c = sum g * xhat (xhat with the GROUP
// statistics): A_g = sum_c gamma_c dbeta_c, B_g = sum_c gamma_c dgamma_c; dx = a g + k0 - c2 x with a = gamma_c rstd_g,
// c2 = rstd_g^2 B_g / m, k0 = c2 mu_g - rstd_g A_g / m (m = pixels * channels per group).
__global__ void gn_bwd_consts_kernel(const float* __restrict__ dgamma_c, const float* __restrict__ dbeta_c,
                                     const float* __restrict__ gamma, const float* __restrict__ gmean,
                                     const float* __restrict__ grstd, int C, int Creal, int G, float inv_m,
                                     float* __restrict__ c2, float* __restrict__ k0) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = Creal / G;
  if (g < G) {
    double A = 0.0, B = 0.0;
    for (int j = 0; j < cg; ++j) {
      const int c = g * cg + j;
      A += (double)gamma[c] * dbeta_c[c];
      B += (double)gamma[c] * dgamma_c[c];
    }
    const float rs = grstd[g * cg], mu = gmean[g * cg];
    const float c2v = rs * rs * (float)B * inv_m;
    const float k0v = c2v * mu - rs * (float)A * inv_m;
    for (int j = 0; j < cg; ++j) { c2[g * cg + j] = c2v; k0[g * cg + j] = k0v; }
  }
  if (g == 0)
    for (int c = Creal; c < C; ++c) { c2[c] = 0.f; k0[c] = 0.f; }
}

// squeeze-excitation gate: sig = bf16(sigmoid(s)) (nn.Sigmoid on the bf16 output of the second nn.Linear under autocast)
__global__ void sigmoid_fwd_kernel(const float* __restrict__ s, float* __restrict__ sig, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sig[i] = bf16_round(1.f / (1.f + __expf(-s[i])));
}
// ds = bf16(dsig * sig * (1 - sig))
__global__ void sigmoid_bwd_kernel(const float* __restrict__ dsig, const float* __restrict__ sig, float* __restrict__ ds,
                                   int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ds[i] = bf16_round(dsig[i] * sig[i] * (1.f - sig[i]));
}

}  // namespace evb

using namespace evb;
#define ST ((cudaStream_t)stream)
#define LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA)

extern "C" int evb_gn_fold(const float* mean_c, const float* rstd_c, float eps_bn, const float* gamma, const float* beta,
                           int C, int Creal, int G, float eps, float* scale, float* shift, float* gmean, float* grstd,
                           void* stream) {
  if (G < 1 || Creal < G || Creal % G || Creal > C) return EVB_ERR_ARG;
  gn_fold_kernel<<<(G + 63) / 64, 64, 0, ST>>>(mean_c, rstd_c, eps_bn, gamma, beta, C, Creal, G, eps, scale, shift, gmean,
                                               grstd);
  return LAUNCH_OK();
}
extern "C" int evb_gn_bwd_consts(const float* dgamma_c, const float* dbeta_c, const float* gamma, const float* gmean,
                                 const float* grstd, int C, int Creal, int G, float inv_m, float* c2, float* k0,
                                 void* stream) {
  if (G < 1 || Creal < G || Creal % G || Creal > C) return EVB_ERR_ARG;
  gn_bwd_consts_kernel<<<(G + 63) / 64, 64, 0, ST>>>(dgamma_c, dbeta_c, gamma, gmean, grstd, C, Creal, G, inv_m, c2, k0);
  return LAUNCH_OK();
}
extern "C" int evb_sigmoid_fwd(const float* s, float* sig, int n, void* stream) {
  if (n < 1) return EVB_ERR_ARG;
  sigmoid_fwd_kernel<<<(n + 255) / 256, 256, 0, ST>>>(s, sig, n);
  return LAUNCH_OK();
}
extern "C" int evb_sigmoid_bwd(const float* dsig, const float* sig, float* ds, int n, void* stream) {
  if (n < 1) return EVB_ERR_ARG;
  sigmoid_bwd_kernel<<<(n + 255) / 256, 256, 0, ST>>>(dsig, sig, ds, n);
  return LAUNCH_OK();
}

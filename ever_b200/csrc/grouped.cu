// Grouped 3x3 convolution of the ResNeXt bottleneck (ever/module/_resnets.py:21-24,80-84: conv3x3(width, width, stride,
// groups); resnext50_32x4d / resnext101_32x4d / resnext101_32x8d, :291-324) on the dense tensor-core kernels: the
// [Co][Ci/G][k][k] fp32 master weight is expanded into a block-diagonal dense [Co][Ci][k][k] matrix (zeros outside the
// group's block, written once at allocation) that the weight pack reads, and the dense weight gradient the wgrad kernel
// produces is read back on the diagonal blocks only.  Both are latency-bound helper kernels (<= 9.4 M floats).
#include "common.cuh"

namespace evb {

// dense[o][g(o) * cpg * kk + j] = w[o][j],  j in [0, cpg * kk)  (OIHW rows: a group's inputs are contiguous in a row)
__global__ void __launch_bounds__(256)
group_expand_kernel(const float* __restrict__ w, float* __restrict__ dense, int Co, int cpg_kk, int opg, long long ld) {
  const long long total = (long long)Co * cpg_kk;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(t / cpg_kk), j = (int)(t % cpg_kk);
    dense[o * ld + (long long)(o / opg) * cpg_kk + j] = w[t];
  }
}
// gw[o][j] (+)= dense[o][g(o) * cpg * kk + j]
__global__ void __launch_bounds__(256)
group_extract_kernel(const float* __restrict__ dense, float* __restrict__ gw, int Co, int cpg_kk, int opg, long long ld,
                     int accumulate) {
  const long long total = (long long)Co * cpg_kk;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(t / cpg_kk), j = (int)(t % cpg_kk);
    const float v = dense[o * ld + (long long)(o / opg) * cpg_kk + j];
    gw[t] = accumulate ? gw[t] + v : v;
  }
}

}  // namespace evb

using namespace evb;

static inline int grp_blocks(long long total) {
  long long b = (total + 255) / 256;
  return (int)(b > 148 * 8 ? 148 * 8 : b);
}
static inline bool grp_args_ok(int Co, int Ci, int kk, int groups) {
  return Co > 0 && Ci > 0 && kk > 0 && groups > 0 && Co % groups == 0 && Ci % groups == 0;
}

extern "C" int evb_group_expand(const float* w, float* dense, int Co, int Ci, int kk, int groups, void* stream) {
  if (!grp_args_ok(Co, Ci, kk, groups)) return EVB_ERR_ARG;
  const int cpg_kk = Ci / groups * kk;
  group_expand_kernel<<<grp_blocks((long long)Co * cpg_kk), 256, 0, (cudaStream_t)stream>>>(w, dense, Co, cpg_kk, Co / groups,
                                                                                          (long long)Ci * kk);
  return cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA;
}
extern "C" int evb_group_extract(const float* dense, float* gw, int Co, int Ci, int kk, int groups, int accumulate,
                                 void* stream) {
  if (!grp_args_ok(Co, Ci, kk, groups)) return EVB_ERR_ARG;
  const int cpg_kk = Ci / groups * kk;
  group_extract_kernel<<<grp_blocks((long long)Co * cpg_kk), 256, 0, (cudaStream_t)stream>>>(dense, gw, Co, cpg_kk, Co / groups,
                                                                                           (long long)Ci * kk, accumulate);
  return cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA;
}

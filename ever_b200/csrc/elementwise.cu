// HBM-bound kernels around the convolutions: BatchNorm (batch statistics, apply(+residual)(+ReLU),
// backward), max-pool, bilinear (align_corners) up-sampling fwd/bwd, 2x2 sum-pool (nearest-up backward),
// decoder merge, global average pool, stem im2col, weight packing.  NHWC bf16, 16-byte vector accesses,
// fp32 math, rounding points chosen to mirror the reference's bf16-autocast flow.
//
// Reference ops replaced: nn.BatchNorm2d (ever/module/_resnets.py:46-49,83-87; fpn.py:166; fs_relation.py:43,50),
// nn.ReLU, residual add (_resnets.py:66-67,109-110), nn.MaxPool2d(3,2,1) (_resnets.py:153),
// nn.UpsamplingBilinear2d via Bf16compatible (ever/module/ops.py:152-166, fpn.py:168,180), F.interpolate nearest
// backward (fpn.py:100), sum()/len() merge (fpn.py:189), F.adaptive_avg_pool2d (fs_relation.py:177).
#include "common.cuh"

namespace evb {

constexpr int kEwThreads = 256;
// kNbPad (common.cuh): stride of the per-channel partial rows; colsum grids are capped at 296 blocks (2 per SM)

static inline int ew_blocks(long long work, int per_block, int cap = 148 * 16) {
  long long b = (work + per_block - 1) / per_block;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------------------------------------
// Per-channel sums over M rows of an [M, C] bf16 matrix.  MODE 0: sum x, sum x^2.
// MODE 1 (BN backward): g = dy * mask, sums g and g * xhat, xhat = (x - mean) * rstd.
//   mask_mode 0: none; 1: (ymask > 0) from a stored post-activation tensor; 2: (x*scale+shift > 0) recomputed.
// Output: partial[which][C][kNbPad] fp32, one column per block (reduced in fixed order by the finalize kernels ->
// deterministic; a warp reads a channel's partials coalesced).
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kEwThreads)
colsum_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
              const __nv_bfloat16* __restrict__ ymask, const float* __restrict__ mean, const float* __restrict__ rstd,
              const float* __restrict__ scale, const float* __restrict__ shift, int mask_mode, long long M, int C,
              float* __restrict__ partial) {
  extern __shared__ float sm[];  // [rows_par][cg][16]
  const int cg = C / 8;
  const int rows_par = kEwThreads / cg > 0 ? kEwThreads / cg : 1;
  const int t = threadIdx.x;
  const int g = t % cg, rsub = t / cg;
  const long long rows_per_block = (M + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
  if (rsub < rows_par) {
    for (int gg = g; gg < cg; gg += kEwThreads) {  // cg <= 256 in practice: single trip
      float mu[8], rs[8], sc[8], sh[8];
      if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          mu[j] = mean[gg * 8 + j];
          rs[j] = rstd[gg * 8 + j];
          if (mask_mode == 2) { sc[j] = scale[gg * 8 + j]; sh[j] = shift[gg * 8 + j]; }
        }
      }
      constexpr int UN = MODE == 0 ? 4 : 2;
      for (long long rb = r0 + rsub; rb < r1; rb += (long long)rows_par * UN) {
        bf16x8 xq[UN], gq[UN], yq[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const long long r = rb + (long long)u * rows_par;
          if (r < r1) {
            xq[u] = *reinterpret_cast<const bf16x8*>(x + r * C + gg * 8);
            if (MODE == 1) {
              gq[u] = *reinterpret_cast<const bf16x8*>(dy + r * C + gg * 8);
              if (mask_mode == 1) yq[u] = *reinterpret_cast<const bf16x8*>(ymask + r * C + gg * 8);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const long long r = rb + (long long)u * rows_par;
          if (r >= r1) continue;
          float xv[8];
          unpack8(xq[u], xv);
          if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { s0[j] += xv[j]; s1[j] += xv[j] * xv[j]; }
          } else {
            float gv[8];
            unpack8(gq[u], gv);
            if (mask_mode == 1) {
              float yv[8];
              unpack8(yq[u], yv);
#pragma unroll
              for (int j = 0; j < 8; ++j) gv[j] = yv[j] > 0.f ? gv[j] : 0.f;
            } else if (mask_mode == 2) {
#pragma unroll
              for (int j = 0; j < 8; ++j) gv[j] = bf16_round(xv[j] * sc[j] + sh[j]) > 0.f ? gv[j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { s0[j] += gv[j]; s1[j] += gv[j] * (xv[j] - mu[j]) * rs[j]; }
          }
        }
      }
    }
  }
  float* my = sm + (size_t)t * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) { my[j] = s0[j]; my[8 + j] = s1[j]; }
  __syncthreads();
  // outputs: for channel c = gg*8+j : which in {0,1}
  for (int o = t; o < 2 * C; o += kEwThreads) {
    const int which = o / C, c = o % C;
    const int gg = c / 8, j = c % 8;
    float acc = 0.f;
    for (int rs_ = 0; rs_ < rows_par; ++rs_) acc += sm[((size_t)rs_ * cg + gg) * 16 + which * 8 + j];
    partial[((size_t)which * C + c) * kNbPad + blockIdx.x] = acc;
  }
}

// BN training statistics finalize: mean, biased var -> rstd, folded scale/shift, running-stat update
// (momentum, unbiased var), reference semantics SURVEY.md Appendix D.
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// one warp per channel: lanes stride over the partial blocks, fixed-order tree -> deterministic
template <int LDP = kNbPad>
__device__ __forceinline__ void reduce_partials(const float* __restrict__ partial, int nblk, int C, int c, double& s,
                                                double& ss) {
  const int lane = threadIdx.x & 31;
  double a = 0.0, b = 0.0;
  const float* pa = partial + (size_t)c * LDP;
  const float* pb = partial + ((size_t)C + c) * LDP;
#pragma unroll
  for (int j = 0; j < LDP / 32; ++j) {
    const int k = lane + 32 * j;
    if (k < nblk) { a += pa[k]; b += pb[k]; }
  }
  s = warp_sum_d(a);
  ss = warp_sum_d(b);
}

__global__ void bn_finalize_kernel(const float* __restrict__ partial, int nblk, long long M, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ scale, float* __restrict__ shift) {
  pdl_wait();   // may be launched programmatically behind the conv whose epilogue wrote `partial`
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;
  double s, ss;
  reduce_partials(partial, nblk, C, c, s, ss);
  if ((threadIdx.x & 31) != 0) return;
  const double mu = s / (double)M;
  double var = ss / (double)M - mu * mu;
  if (var < 0) var = 0;
  const float rs = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)mu;
  rstd[c] = rs;
  const float a = gamma[c] * rs;
  scale[c] = a;
  shift[c] = beta[c] - (float)mu * a;
  if (running_mean) {
    const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

// SyncBatchNorm support: the per-CTA partial columns of a rank are summed (fixed order) into sums[2][C]; the caller
// all-reduces sums over the ranks and finalizes with the global row count.
__global__ void bn_partial_sums_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ sums) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;
  double s, ss;
  reduce_partials(partial, nblk, C, c, s, ss);
  if ((threadIdx.x & 31) != 0) return;
  sums[c] = (float)s;
  sums[C + c] = (float)ss;
}
__global__ void bn_finalize_sums_kernel(const float* __restrict__ sums, double M, int C, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float* __restrict__ running_mean,
                                        float* __restrict__ running_var, float momentum, float eps, float* __restrict__ mean,
                                        float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = (double)sums[c] / M;
  double var = (double)sums[C + c] / M - mu * mu;
  if (var < 0) var = 0;
  const float rs = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)mu;
  rstd[c] = rs;
  const float a = gamma[c] * rs;
  scale[c] = a;
  shift[c] = beta[c] - (float)mu * a;
  if (running_mean) {
    const double unb = M > 1 ? var * M / (M - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}
// backward constants of dx = a g + k0 - c2 x from (all-reduced) sums: c2 = a rstd dgamma / M, k0 = c2 mean - a dbeta / M
__global__ void bn_bwd_consts_kernel(const float* __restrict__ scale, const float* __restrict__ rstd,
                                     const float* __restrict__ mean, const float* __restrict__ dgamma,
                                     const float* __restrict__ dbeta, float inv_m, int C, float* __restrict__ c2,
                                     float* __restrict__ k0) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float a = scale[c];
  const float v = a * rstd[c] * dgamma[c] * inv_m;
  c2[c] = v;
  k0[c] = v * mean[c] - a * dbeta[c] * inv_m;
}

// eval / frozen BN: fold running stats
__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, int C,
                               float* scale, float* shift, float* mean, float* rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float rs = rsqrtf(rv[c] + eps);
  const float a = gamma[c] * rs;
  scale[c] = a;
  shift[c] = beta[c] - rm[c] * a;
  if (mean) { mean[c] = rm[c]; rstd[c] = rs; }
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int accumulate, float* __restrict__ fresh) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;
  double s, ss;
  reduce_partials(partial, nblk, C, c, s, ss);
  if ((threadIdx.x & 31) != 0) return;
  if (fresh) { fresh[c] = (float)s; fresh[C + c] = (float)ss; }  // this launch's own sums: [dbeta | dgamma]
  if (dgamma) { if (accumulate) dgamma[c] += (float)ss; else dgamma[c] = (float)ss; }
  if (dbeta) { if (accumulate) dbeta[c] += (float)s; else dbeta[c] = (float)s; }
}


// dx = scale * (g - (dbeta + xhat * dgamma) / M) = a*g + k0 - c2*x,  g = dy * mask ;  optional dres (+)= g
//   c2 = a * rstd * dgamma / M,  k0 = c2 * mean - a * dbeta / M   (hoisted per channel); kernels below
// ------------------------------------------------------------------------------------------------
// Narrow-vector variants of the BatchNorm streaming kernels.  A thread owns V (4 or 8) consecutive channels, so its
// per-channel constants live in 3-4 * V registers; with V = 4 the kernels need ~half the registers of the V = 8 versions,
// twice as many CTAs fit per SM and twice the bytes are in flight per SM (measured on B200: the V = 8 backward kernels
// ran at 46 % of the copy bandwidth, in-flight-limited).  MASK is a template parameter so unused constants are pruned.
// ------------------------------------------------------------------------------------------------
// BN backward reduction: per-channel sums of g = dy * mask and g * (x - mean) over M rows (rstd is applied by the finalize
// kernel).  MASK 0: none; 1: ymask > 0; 2: recomputed bf16(x*scale+shift) > 0; 3: ymask is a uint32 per 8-channel vector, bit j = channel j survived the ReLU (written by bn_apply_ca<1>: a quarter of the bytes of re-reading the bf16 output).  partial[which][C][kNbPadBwd], column = block.
constexpr int kNbPadBwd = 640;   // up to 4 reduce blocks per SM

// finalize of bn_bwd_reduce_kernel: dbeta = sum g, dgamma = rstd * sum g (x - mean); fixed-order fp64 reduction
__global__ void bn_bwd_finalize2_kernel(const float* __restrict__ partial, int nblk, int C, const float* __restrict__ rstd,
                                        float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate,
                                        float* __restrict__ fresh) {
  pdl_wait();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;
  double s, ss;
  reduce_partials<kNbPadBwd>(partial, nblk, C, c, s, ss);
  if ((threadIdx.x & 31) != 0) return;
  const float dg = (float)(ss * (double)rstd[c]);
  fresh[c] = (float)s;
  fresh[C + c] = dg;
  if (dgamma) { if (accumulate) dgamma[c] += dg; else dgamma[c] = dg; }
  if (dbeta) { if (accumulate) dbeta[c] += (float)s; else dbeta[c] = (float)s; }
}



// ------------------------------------------------------------------------------------------------
// cp.async-staged BatchNorm backward.  The register-staged kernels above keep at most UN 16-byte loads per tensor in
// flight per thread and stall between "issue" and "consume" phases (measured 2.7-3.7 TB/s of the 6.5 TB/s copy peak on
// the 67 MB layers).  Here every thread streams its own vectors through a private ring of S shared-memory slots filled
// by cp.async (LDGSTS, L1-bypassing): S-1 loads per tensor stay in flight per thread without holding registers, no block
// barrier is needed (a thread only reads slots it filled itself), and 3 CTAs x 256 threads fit per SM.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kCaStages = 8;

// per-channel sums of g = dy*mask and g*(x-mean); same partial layout / finalize as bn_bwd_reduce_kernel
template <int MASK>
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_reduce_ca_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                        const __nv_bfloat16* __restrict__ ymask, const float* __restrict__ mean,
                        const float* __restrict__ scale, const float* __restrict__ shift, long long M, int C,
                        float* __restrict__ partial) {
  constexpr int NT = (MASK == 1 || MASK == 3) ? 3 : 2;   // tensors streamed
  constexpr int S = kCaStages;
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16x8* ring = reinterpret_cast<bf16x8*>(smraw);   // [S][NT][blockDim]
  const int nthr = blockDim.x;
  const int cg = C / 8;
  const int rows_par = nthr / cg;
  const int t = threadIdx.x;
  const int g = t % cg, rsub = t / cg;
  const long long rows_per_block = (M + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  float s0[8], s1[8], mu[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s0[j] = s1[j] = 0.f;
    mu[j] = mean[g * 8 + j];
    if (MASK == 2) { sc[j] = scale[g * 8 + j]; sh[j] = shift[g * 8 + j]; }
  }
  const size_t col = (size_t)g * 8;
  const long long first = r0 + rsub;
  const int niter = first < r1 ? (int)((r1 - first + rows_par - 1) / rows_par) : 0;
  auto issue = [&](int it) {
    if (it < niter) {
      const size_t off = (size_t)(first + (long long)it * rows_par) * C + col;
      bf16x8* slot = ring + (size_t)(it % S) * NT * nthr + t;
      cp_async16(slot, x + off);
      cp_async16(slot + nthr, dy + off);
      if (MASK == 1) cp_async16(slot + 2 * nthr, ymask + off);
      if (MASK == 3) cp_async4(slot + 2 * nthr, reinterpret_cast<const unsigned*>(ymask) + off / 8);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int it = 0; it < S - 1; ++it) issue(it);
  for (int it = 0; it < niter; ++it) {
    issue(it + S - 1);
    cp_async_wait<S - 1>();
    const bf16x8* slot = ring + (size_t)(it % S) * NT * nthr + t;
    float xv[8], gv[8];
    unpack8(slot[0], xv);
    unpack8(slot[nthr], gv);
    if (MASK == 1) {
      float yv[8];
      unpack8(slot[2 * nthr], yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) gv[j] = yv[j] > 0.f ? gv[j] : 0.f;
    } else if (MASK == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) gv[j] = bf16_round(xv[j] * sc[j] + sh[j]) > 0.f ? gv[j] : 0.f;
    } else if (MASK == 3) {
      const unsigned mb = *reinterpret_cast<const unsigned*>(slot + 2 * nthr);
#pragma unroll
      for (int j = 0; j < 8; ++j) gv[j] = (mb >> j & 1u) ? gv[j] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0[j] += gv[j]; s1[j] += gv[j] * (xv[j] - mu[j]); }
  }
  cp_async_wait<0>();
  __syncthreads();   // everyone is done with the ring: reuse it for the cross-row reduction
  float* sm = reinterpret_cast<float*>(smraw);
  float* my = sm + (size_t)t * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) { my[j] = s0[j]; my[8 + j] = s1[j]; }
  __syncthreads();
  for (int o = t; o < 2 * C; o += nthr) {
    const int which = o / C, c = o % C;
    const int gg = c / 8, j = c % 8;
    float acc = 0.f;
    for (int rs_ = 0; rs_ < rows_par; ++rs_) acc += sm[((size_t)rs_ * cg + gg) * 16 + which * 8 + j];
    partial[((size_t)which * C + c) * kNbPadBwd + blockIdx.x] = acc;
  }
}

// dx = a*g + k0 - c2*x ; optional dres (+)= g   (the dres accumulate input is a plain load: it is rare and L2-resident)
template <int MASK>
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_apply_ca_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                       const __nv_bfloat16* __restrict__ ymask, const float* __restrict__ mean,
                       const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                       const float* __restrict__ dgamma, const float* __restrict__ dbeta, int frozen,
                       __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres, int dres_acc, unsigned nvec, int C,
                       float inv_m) {
  constexpr int NT = (MASK == 1 || MASK == 3) ? 3 : 2;
  constexpr int S = kCaStages;
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16x8* ring = reinterpret_cast<bf16x8*>(smraw);   // [S][NT][blockDim]
  const int nthr = blockDim.x;
  const int t = threadIdx.x;
  const int cg = C / 8;
  const unsigned tid = blockIdx.x * nthr + t;
  const unsigned stride = gridDim.x * nthr;
  const int c0 = (int)(tid % (unsigned)cg) * 8;
  pdl_wait();
  float a[8], k0[8], c2[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    a[j] = scale[c];
    if (MASK == 2) sh[j] = shift[c];
    if (frozen == 2) { c2[j] = dgamma[c]; k0[j] = dbeta[c]; }   // explicit constants (evb_norm_bwd_apply: GroupNorm, plain ReLU)
    else if (frozen) { c2[j] = 0.f; k0[j] = 0.f; }
    else {
      c2[j] = a[j] * rstd[c] * dgamma[c] * inv_m;
      k0[j] = c2[j] * mean[c] - a[j] * dbeta[c] * inv_m;
    }
  }
  const int niter = tid < nvec ? (int)((nvec - tid + stride - 1) / stride) : 0;
  const bf16x8* gx = reinterpret_cast<const bf16x8*>(x);
  const bf16x8* gg = reinterpret_cast<const bf16x8*>(dy);
  const bf16x8* gy = reinterpret_cast<const bf16x8*>(ymask);
  auto issue = [&](int it) {
    if (it < niter) {
      const size_t i = (size_t)tid + (size_t)it * stride;
      bf16x8* slot = ring + (size_t)(it % S) * NT * nthr + t;
      cp_async16(slot, gx + i);
      cp_async16(slot + nthr, gg + i);
      if (MASK == 1) cp_async16(slot + 2 * nthr, gy + i);
      if (MASK == 3) cp_async4(slot + 2 * nthr, reinterpret_cast<const unsigned*>(ymask) + i);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int it = 0; it < S - 1; ++it) issue(it);
  for (int it = 0; it < niter; ++it) {
    issue(it + S - 1);
    cp_async_wait<S - 1>();
    const size_t i = (size_t)tid + (size_t)it * stride;
    const bf16x8* slot = ring + (size_t)(it % S) * NT * nthr + t;
    float xf[8], g[8];
    unpack8(slot[0], xf);
    unpack8(slot[nthr], g);
    if (MASK == 1) {
      float yf[8];
      unpack8(slot[2 * nthr], yf);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = yf[j] > 0.f ? g[j] : 0.f;
    } else if (MASK == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = bf16_round(xf[j] * a[j] + sh[j]) > 0.f ? g[j] : 0.f;
    } else if (MASK == 3) {
      const unsigned mb = *reinterpret_cast<const unsigned*>(slot + 2 * nthr);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = (mb >> j & 1u) ? g[j] : 0.f;
    }
    if (dres) {
      float d[8];
      if (dres_acc) {
        unpack8(reinterpret_cast<const bf16x8*>(dres)[i], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += g[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = g[j];
      }
      reinterpret_cast<bf16x8*>(dres)[i] = pack8(d);
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a[j] * g[j] + k0[j] - c2[j] * xf[j];
    reinterpret_cast<bf16x8*>(dx)[i] = pack8(o);
  }
  cp_async_wait<0>();
}

// y = act(bf16(x*scale+shift) [+ res]) with the cp.async rings (RES: the residual tensor is streamed too)
template <int RES>
__global__ void __launch_bounds__(kEwThreads)
bn_apply_ca_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                   const __nv_bfloat16* __restrict__ res, __nv_bfloat16* __restrict__ y, unsigned nvec, int C, int relu,
                   unsigned* __restrict__ mbits) {
  constexpr int NT = RES ? 2 : 1;
  constexpr int S = kCaStages;
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16x8* ring = reinterpret_cast<bf16x8*>(smraw);   // [S][NT][blockDim]
  const int nthr = blockDim.x, t = threadIdx.x;
  const int cg = C / 8;
  const unsigned tid = blockIdx.x * nthr + t;
  const unsigned stride = gridDim.x * nthr;
  const int c0 = (int)(tid % (unsigned)cg) * 8;
  pdl_wait();
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = scale[c0 + j]; b[j] = shift[c0 + j]; }
  const int niter = tid < nvec ? (int)((nvec - tid + stride - 1) / stride) : 0;
  const bf16x8* gx = reinterpret_cast<const bf16x8*>(x);
  const bf16x8* gr = reinterpret_cast<const bf16x8*>(res);
  auto issue = [&](int it) {
    if (it < niter) {
      const size_t i = (size_t)tid + (size_t)it * stride;
      bf16x8* slot = ring + (size_t)(it % S) * NT * nthr + t;
      cp_async16(slot, gx + i);
      if (RES) cp_async16(slot + nthr, gr + i);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int it = 0; it < S - 1; ++it) issue(it);
  for (int it = 0; it < niter; ++it) {
    issue(it + S - 1);
    cp_async_wait<S - 1>();
    const size_t i = (size_t)tid + (size_t)it * stride;
    const bf16x8* slot = ring + (size_t)(it % S) * NT * nthr + t;
    float v[8];
    unpack8(slot[0], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = v[j] * a[j] + b[j];
    if (RES) {
      float r[8];
      unpack8(slot[nthr], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = bf16_round(v[j]) + r[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    const bf16x8 out = pack8(v);
    reinterpret_cast<bf16x8*>(y)[i] = out;
    if (RES && mbits) {   // survivors of the ReLU as the backward pass sees them: the ROUNDED output > 0
      float r8[8];
      unpack8(out, r8);
      unsigned mb = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) mb |= (r8[j] > 0.f ? 1u : 0u) << j;
      mbits[i] = mb;
    }
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------ max pool 3x3 s2 p1
__global__ void __launch_bounds__(kEwThreads)
maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx, int N,
                   int H, int W, int C) {
  const int cg = C / 8, Ho = H / 2, Wo = W / 2;
  const unsigned total = (unsigned) (long long)N * Ho * Wo * cg;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    unsigned p = i / cg;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * ho - 1 + r;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = 2 * wo - 1 + s;
        if (w < 0 || w >= W) continue;
        float v[8];
        unpack8(*reinterpret_cast<const bf16x8*>(x + (((long long)n * H + h) * W + w) * C + g * 8), v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] > best[j] || (best[j] == -INFINITY)) { best[j] = v[j]; bi[j] = r * 3 + s; }  // first max wins
      }
    }
    reinterpret_cast<bf16x8*>(y)[i] = pack8(best);
    uint2 packed;
    packed.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    packed.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    reinterpret_cast<uint2*>(idx)[i] = packed;
  }
}

__global__ void __launch_bounds__(kEwThreads)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx, __nv_bfloat16* __restrict__ dx,
                   int N, int H, int W, int C) {
  const int cg = C / 8, Ho = H / 2, Wo = W / 2;
  const unsigned total = (unsigned) (long long)N * H * W * cg;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    unsigned p = i / cg;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int n = (int)(p / H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // windows (ho, wo) containing (h, w): 2ho-1 <= h <= 2ho+1
    for (int ho = (h) / 2; ho <= (h + 1) / 2; ++ho) {
      if (ho < 0 || ho >= Ho) continue;
      const int r = h - (2 * ho - 1);
      for (int wo = (w) / 2; wo <= (w + 1) / 2; ++wo) {
        if (wo >= Wo) continue;
        const int s = w - (2 * wo - 1);
        const int code = r * 3 + s;
        const long long o = (((long long)n * Ho + ho) * Wo + wo) * cg + g;
        const uint2 pk = reinterpret_cast<const uint2*>(idx)[o];
        float gv[8];
        unpack8(reinterpret_cast<const bf16x8*>(dy)[o], gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int b = (j < 4 ? (pk.x >> (8 * j)) : (pk.y >> (8 * (j - 4)))) & 0xff;
          if (b == code) acc[j] += gv[j];
        }
      }
    }
    reinterpret_cast<bf16x8*>(dx)[i] = pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------------ bilinear (align_corners)
// y[N, f*h, f*w, C] = bf16( bilinear( act(x) ) ),  act(x) = bf16(relu(x*scale+shift)) when scale != null.
// Input rows have stride ldx (>= C) elements per pixel, output ldy.
// P2: C/8, Wo and Ho are powers of two (every FarSeg shape): the index decomposition is shifts and masks instead of
// five integer divisions per output vector (the kernel is instruction-bound: ~40 % of its instructions were div/mod).
template <bool P2>
__global__ void __launch_bounds__(kEwThreads)
bilinear_up_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                   __nv_bfloat16* __restrict__ y, int N, int h, int w, int C, int ldx, int ldy, int f, int lg_cg, int lg_wo,
                   int lg_ho) {
  // blockDim is a multiple of C/8: a thread's channel group is fixed, the BN fold is hoisted, UN pixels in flight
  constexpr int UN = 2;
  const int cg = C / 8, Ho = h * f, Wo = w * f;
  const float sy = Ho > 1 ? (float)(h - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const unsigned total = (unsigned)N * Ho * Wo * cg;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned stride = gridDim.x * blockDim.x;
  const int g = P2 ? (int)(tid & (unsigned)(cg - 1)) : (int)(tid % cg);
  float a[8], b[8];
  if (scale) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = scale[g * 8 + j]; b[j] = shift[g * 8 + j]; }
  }
  for (unsigned i0 = tid; i0 < total; i0 += stride * UN) {
    bf16x8 q[UN][4];
    float ly[UN], lx[UN];
    long long oidx[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const unsigned i = i0 + u * stride;
      if (i >= total) continue;
      int ox, oy, n;
      if (P2) {
        unsigned p = i >> lg_cg;
        ox = (int)(p & (unsigned)(Wo - 1)); p >>= lg_wo;
        oy = (int)(p & (unsigned)(Ho - 1));
        n = (int)(p >> lg_ho);
      } else {
        unsigned p = i / cg;
        ox = (int)(p % Wo); p /= Wo;
        oy = (int)(p % Ho);
        n = (int)(p / Ho);
      }
      const float fy = sy * oy, fx = sx * ox;
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
      ly[u] = fy - y0; lx[u] = fx - x0;
      const __nv_bfloat16* base = x + (long long)n * h * w * ldx + g * 8;
      q[u][0] = *reinterpret_cast<const bf16x8*>(base + ((long long)y0 * w + x0) * ldx);
      q[u][1] = *reinterpret_cast<const bf16x8*>(base + ((long long)y0 * w + x1) * ldx);
      q[u][2] = *reinterpret_cast<const bf16x8*>(base + ((long long)y1 * w + x0) * ldx);
      q[u][3] = *reinterpret_cast<const bf16x8*>(base + ((long long)y1 * w + x1) * ldx);
      oidx[u] = (((long long)n * Ho + oy) * Wo + ox) * ldy + g * 8;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const unsigned i = i0 + u * stride;
      if (i >= total) continue;
      float v00[8], v01[8], v10[8], v11[8], o[8];
      unpack8(q[u][0], v00); unpack8(q[u][1], v01); unpack8(q[u][2], v10); unpack8(q[u][3], v11);
      const float hy = 1.f - ly[u], hx = 1.f - lx[u];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (scale) {
          v00[j] = fmaxf(bf16_round(v00[j] * a[j] + b[j]), 0.f);
          v01[j] = fmaxf(bf16_round(v01[j] * a[j] + b[j]), 0.f);
          v10[j] = fmaxf(bf16_round(v10[j] * a[j] + b[j]), 0.f);
          v11[j] = fmaxf(bf16_round(v11[j] * a[j] + b[j]), 0.f);
        }
        // same association as ATen upsample_bilinear2d: hy*(hx*v00 + lx*v01) + ly*(hx*v10 + lx*v11)
        o[j] = hy * (hx * v00[j] + lx[u] * v01[j]) + ly[u] * (hx * v10[j] + lx[u] * v11[j]);
      }
      *reinterpret_cast<bf16x8*>(y + oidx[u]) = pack8(o);
    }
  }
}

// Gather form of the transpose: dx[n,iy,ix,:] = sum over outputs that read (iy,ix) of weight * dy.
__global__ void __launch_bounds__(kEwThreads)
bilinear_up_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int h, int w, int C,
                       int lddy, int lddx, int f) {
  const int cg = C / 8, Ho = h * f, Wo = w * f;
  const float sy = Ho > 1 ? (float)(h - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const unsigned total = (unsigned) (long long)N * h * w * cg;
  const float inv_sy = sy > 0.f ? 1.f / sy : 0.f, inv_sx = sx > 0.f ? 1.f / sx : 0.f;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    unsigned p = i / cg;
    const int ix = (int)(p % w); p /= w;
    const int iy = (int)(p % h);
    const int n = (int)(p / h);
    // outputs whose source coordinate falls in (iy-1, iy+1): o in ((iy-1)/s, (iy+1)/s), widened by one for rounding;
    // the exact membership test below uses the same float formula as the forward kernel.
    int ylo = 0, yhi = Ho - 1, xlo = 0, xhi = Wo - 1;
    if (sy > 0.f) { ylo = max(0, (int)floorf((iy - 1) * inv_sy) - 1); yhi = min(Ho - 1, (int)ceilf((iy + 1) * inv_sy) + 1); }
    if (sx > 0.f) { xlo = max(0, (int)floorf((ix - 1) * inv_sx) - 1); xhi = min(Wo - 1, (int)ceilf((ix + 1) * inv_sx) + 1); }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int oy = ylo; oy <= yhi; ++oy) {
      const float fy = sy * oy;
      const int y0 = (int)fy;
      const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
      const float ly = fy - y0;
      float wy = 0.f;
      if (y0 == iy) wy += 1.f - ly;
      if (y1 == iy) wy += ly;
      if (wy == 0.f) continue;
      for (int ox = xlo; ox <= xhi; ++ox) {
        const float fx = sx * ox;
        const int x0 = (int)fx;
        const int x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float lx = fx - x0;
        float wx = 0.f;
        if (x0 == ix) wx += 1.f - lx;
        if (x1 == ix) wx += lx;
        if (wx == 0.f) continue;
        const float wgt = wy * wx;
        float gv[8];
        unpack8(*reinterpret_cast<const bf16x8*>(dy + (((long long)n * Ho + oy) * Wo + ox) * lddy + g * 8), gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += wgt * gv[j];
      }
    }
    *reinterpret_cast<bf16x8*>(dx + (((long long)n * h + iy) * w + ix) * lddx + g * 8) = pack8(acc);
  }
}

// Separable form of the same transpose (fewer taps, coalesced): pass X reduces along the output row into an fp32
// intermediate tmp[N, Ho, w, C]; pass Y reduces tmp along the output column into dx (bf16).
__global__ void __launch_bounds__(kEwThreads)
bilinear_bwd_x_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ tmp, int N, int h, int w, int C, int lddy,
                      int f) {
  const int cg = C / 8, Ho = h * f, Wo = w * f;
  const float sx = Wo > 1 ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const float inv_sx = sx > 0.f ? 1.f / sx : 0.f;
  const unsigned total = (unsigned)N * Ho * w * cg;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    unsigned p = i / cg;
    const int ix = (int)(p % w); p /= w;   // p = n * Ho + oy
    int xlo = 0, xhi = Wo - 1;
    if (sx > 0.f) { xlo = max(0, (int)floorf((ix - 1) * inv_sx) - 1); xhi = min(Wo - 1, (int)ceilf((ix + 1) * inv_sx) + 1); }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const __nv_bfloat16* row = dy + (long long)p * Wo * lddy + g * 8;
    for (int ox = xlo; ox <= xhi; ++ox) {
      const float fx = sx * ox;
      const int x0 = (int)fx;
      const int x1 = x0 + (x0 < w - 1 ? 1 : 0);
      const float lx = fx - x0;
      float wx = 0.f;
      if (x0 == ix) wx += 1.f - lx;
      if (x1 == ix) wx += lx;
      if (wx == 0.f) continue;
      float gv[8];
      unpack8(*reinterpret_cast<const bf16x8*>(row + (long long)ox * lddy), gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += wx * gv[j];
    }
    float4* dst = reinterpret_cast<float4*>(tmp + (long long)i * 8);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}
// Row-staged pass X: one block per output row (n, oy) and chunk of <= 8 channel groups.  The row of dy is read ONCE,
// coalesced, into shared memory; every (ix, group) output then gathers its taps from there in the same order as
// bilinear_bwd_x_kernel (bit-identical sums).  The gather version re-read every dy vector ~2x through L1 with ~11 dependent
// candidate taps per output (67 us for the x4 logit gradient of the C2 step, against ~15 us of HBM time).
__global__ void __launch_bounds__(kEwThreads)
bilinear_bwd_x_row_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ tmp, int w, int C, int lddy, int f,
                          int ncg) {
  extern __shared__ __align__(16) unsigned char bwdx_sm[];
  bf16x8* sh = reinterpret_cast<bf16x8*>(bwdx_sm);   // [Wo][ncg]
  const int cg = C / 8, Wo = w * f;
  const long long p = blockIdx.x;                    // n * Ho + oy
  const int g0 = blockIdx.y * ncg;
  const int ng = min(ncg, cg - g0);
  const __nv_bfloat16* row = dy + p * Wo * lddy + g0 * 8;
  for (int v = threadIdx.x; v < Wo * ng; v += blockDim.x) {
    const int ox = v / ng, g = v - ox * ng;
    sh[ox * ncg + g] = *reinterpret_cast<const bf16x8*>(row + (long long)ox * lddy + g * 8);
  }
  __syncthreads();
  const float sx = Wo > 1 ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const float inv_sx = sx > 0.f ? 1.f / sx : 0.f;
  for (int v = threadIdx.x; v < w * ng; v += blockDim.x) {
    const int ix = v / ng, g = v - ix * ng;
    int xlo = 0, xhi = Wo - 1;
    if (sx > 0.f) { xlo = max(0, (int)floorf((ix - 1) * inv_sx) - 1); xhi = min(Wo - 1, (int)ceilf((ix + 1) * inv_sx) + 1); }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int ox = xlo; ox <= xhi; ++ox) {
      const float fx = sx * ox;
      const int x0 = (int)fx;
      const int x1 = x0 + (x0 < w - 1 ? 1 : 0);
      const float lx = fx - x0;
      float wx = 0.f;
      if (x0 == ix) wx += 1.f - lx;
      if (x1 == ix) wx += lx;
      if (wx == 0.f) continue;
      float gv[8];
      unpack8(sh[ox * ncg + g], gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += wx * gv[j];
    }
    float4* dst = reinterpret_cast<float4*>(tmp + ((p * w + ix) * cg + g0 + g) * 8);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}
__global__ void __launch_bounds__(kEwThreads)
bilinear_bwd_y_kernel(const float* __restrict__ tmp, __nv_bfloat16* __restrict__ dx, int N, int h, int w, int C, int lddx,
                      int f) {
  const int cg = C / 8, Ho = h * f;
  const float sy = Ho > 1 ? (float)(h - 1) / (float)(Ho - 1) : 0.f;
  const float inv_sy = sy > 0.f ? 1.f / sy : 0.f;
  const unsigned total = (unsigned)N * h * w * cg;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    unsigned p = i / cg;
    const int ix = (int)(p % w); p /= w;
    const int iy = (int)(p % h);
    const int n = (int)(p / h);
    int ylo = 0, yhi = Ho - 1;
    if (sy > 0.f) { ylo = max(0, (int)floorf((iy - 1) * inv_sy) - 1); yhi = min(Ho - 1, (int)ceilf((iy + 1) * inv_sy) + 1); }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int oy = ylo; oy <= yhi; ++oy) {
      const float fy = sy * oy;
      const int y0 = (int)fy;
      const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
      const float ly = fy - y0;
      float wy = 0.f;
      if (y0 == iy) wy += 1.f - ly;
      if (y1 == iy) wy += ly;
      if (wy == 0.f) continue;
      const float4* src = reinterpret_cast<const float4*>(tmp + ((((long long)n * Ho + oy) * w + ix) * cg + g) * 8);
      const float4 a = src[0], b = src[1];
      acc[0] += wy * a.x; acc[1] += wy * a.y; acc[2] += wy * a.z; acc[3] += wy * a.w;
      acc[4] += wy * b.x; acc[5] += wy * b.y; acc[6] += wy * b.z; acc[7] += wy * b.w;
    }
    *reinterpret_cast<bf16x8*>(dx + (((long long)n * h + iy) * w + ix) * lddx + g * 8) = pack8(acc);
  }
}

// dcoarse[n,h,w,:] (+)= sum of the 2x2 block of dfine  (backward of nearest x2)
__global__ void __launch_bounds__(kEwThreads)
sumpool2_kernel(const __nv_bfloat16* __restrict__ dfine, __nv_bfloat16* __restrict__ dcoarse, int N, int h, int w, int C,
                int accumulate) {
  const int cg = C / 8;
  const unsigned total = (unsigned) (long long)N * h * w * cg;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    unsigned p = i / cg;
    const int x = (int)(p % w); p /= w;
    const int y = (int)(p % h);
    const int n = (int)(p / h);
    float acc[8];
    if (accumulate) unpack8(reinterpret_cast<const bf16x8*>(dcoarse)[i], acc);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx_ = 0; dx_ < 2; ++dx_) {
        float v[8];
        unpack8(*reinterpret_cast<const bf16x8*>(dfine + (((long long)n * 2 * h + 2 * y + dy) * 2 * w + 2 * x + dx_) * C + g * 8), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    reinterpret_cast<bf16x8*>(dcoarse)[i] = pack8(acc);
  }
}

// out = (((0+a)+b)+c)+d) / 4 with bf16 rounding after every add (python sum() of bf16 tensors, fpn.py:189)
__global__ void __launch_bounds__(kEwThreads)
merge4_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, const __nv_bfloat16* c, const __nv_bfloat16* d,
              __nv_bfloat16* out, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float va[8], vb[8], vc[8], vd[8];
    unpack8(reinterpret_cast<const bf16x8*>(a)[i], va);
    unpack8(reinterpret_cast<const bf16x8*>(b)[i], vb);
    unpack8(reinterpret_cast<const bf16x8*>(c)[i], vc);
    unpack8(reinterpret_cast<const bf16x8*>(d)[i], vd);
#pragma unroll
    for (int j = 0; j < 8; ++j) va[j] = bf16_round(bf16_round(bf16_round(va[j] + vb[j]) + vc[j]) + vd[j]) * 0.25f;
    reinterpret_cast<bf16x8*>(out)[i] = pack8(va);
  }
}

// y = bf16(x * alpha)   (and optional add of a second tensor: y = bf16(x*alpha + z))
__global__ void __launch_bounds__(kEwThreads)
scale_add_kernel(const __nv_bfloat16* x, float alpha, const __nv_bfloat16* z, __nv_bfloat16* y, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    unpack8(reinterpret_cast<const bf16x8*>(x)[i], v);
    if (z) {
      float u[8];
      unpack8(reinterpret_cast<const bf16x8*>(z)[i], u);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] * alpha + u[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= alpha;
    }
    reinterpret_cast<bf16x8*>(y)[i] = pack8(v);
  }
}

// ------------------------------------------------------------------------------------------------ global average pool
// out[n][c] = bf16(mean over hw)   (fp32 storage holding bf16-rounded values)
__global__ void gap_fwd_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int HW, int C) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  const __nv_bfloat16* p = x + (long long)n * HW * C + c;
  for (int i = 0; i < HW; ++i) s += __bfloat162float(p[(long long)i * C]);
  out[n * C + c] = bf16_round(s / (float)HW);
}
// dx[n,hw,c] += bf16(dscene[n][c] / HW)
__global__ void gap_bwd_kernel(const float* __restrict__ dscene, __nv_bfloat16* __restrict__ dx, int HW, int C,
                               long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int n = (int)(i / ((long long)HW * C));
    const float g = bf16_round(dscene[n * C + c] / (float)HW);
    dx[i] = __float2bfloat16_rn(__bfloat162float(dx[i]) + g);
  }
}

// ------------------------------------------------------------------------------------------------ stem im2col
// x: NCHW fp32 [N,Cin,H,W] -> A: [N*Ho*Wo][KP] bf16, k = c*ks*ks + r*ks + s (ks x ks window, given stride / pad), zero
// padded to KP.  (7,2,3) = the 7x7 stem; (3,2,1) = the first conv of the deep (v1c) stem.
__global__ void __launch_bounds__(kEwThreads)
stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ a, int N, int Cin, int H, int W, int KP, int ks,
                   int stride, int pad) {
  const int Ho = H / stride, Wo = W / stride;
  const int kg = KP / 8;
  const unsigned total = (unsigned) (long long)N * Ho * Wo * kg;
  const int kk = ks * ks;
  const int K = Cin * kk;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = (int)(i % kg);
    unsigned p = i / kg;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float val = 0.f;
      if (k < K) {
        const int c = k / kk, rs = k % kk, r = rs / ks, s = rs % ks;
        const int h = stride * ho - pad + r, w = stride * wo - pad + s;
        if (h >= 0 && h < H && w >= 0 && w < W) val = x[(((long long)n * Cin + c) * H + h) * W + w];
      }
      v[j] = val;
    }
    reinterpret_cast<bf16x8*>(a)[i] = pack8(v);
  }
}

// Same lowering from a raw uint8 HWC tile [N,H,W,Cin] with the reference's normalisation fused in:
// value = (u8 - mean[c]) / std[c]  (th_mean_std_normalize, ever/preprocess/function.py:9-32), then bf16.
__global__ void __launch_bounds__(kEwThreads)
stem_im2col_u8_kernel(const uint8_t* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ stdv,
                      __nv_bfloat16* __restrict__ a, int N, int Cin, int H, int W, int KP, int ks, int stride, int pad) {
  const int Ho = H / stride, Wo = W / stride;
  const int kg = KP / 8;
  const unsigned total = (unsigned)N * Ho * Wo * kg;
  const int kk = ks * ks;
  const int K = Cin * kk;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = (int)(i % kg);
    unsigned p = i / kg;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float val = 0.f;
      if (k < K) {
        const int c = k / kk, rs = k % kk, r = rs / ks, s_ = rs % ks;
        const int h = stride * ho - pad + r, w = stride * wo - pad + s_;
        if (h >= 0 && h < H && w >= 0 && w < W)
          val = ((float)x[(((long long)n * H + h) * W + w) * Cin + c] - mean[c]) / stdv[c];
      }
      v[j] = val;
    }
    reinterpret_cast<bf16x8*>(a)[i] = pack8(v);
  }
}

// ------------------------------------------------------------------------------------------------ confusion matrix
// cm[t * K + p] += #pixels with label t and prediction p (labels outside [0,K), e.g. ignore_index, are skipped).
// ever/metric/confusion_matrix.py:11-25 (scipy COO accumulation on the host in the reference).  Integer atomics.
__global__ void __launch_bounds__(256)
confusion_matrix_kernel(const uint8_t* __restrict__ pred, const long long* __restrict__ labels, long long P, int K,
                        unsigned long long* __restrict__ cm) {
  extern __shared__ unsigned int hist[];
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const long long t = labels[p];
    const int q = pred[p];
    if (t >= 0 && t < K && q < K) atomicAdd(&hist[(int)t * K + q], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K; i += blockDim.x)
    if (hist[i]) atomicAdd(&cm[i], (unsigned long long)hist[i]);
}

// ------------------------------------------------------------------------------------------------ weight packing
// w: OIHW fp32 [Co][Ci][k][k] -> wf: bf16 [k*k][CoP][CiP]  and  wb: bf16 [k*k][CiPb][CoPb]  (zero padded)
__global__ void pack_weight_kernel(const float* __restrict__ w, int Co, int Ci, int kk, __nv_bfloat16* __restrict__ wf,
                                   int CoP, int CiP, __nv_bfloat16* __restrict__ wb, int CiPb, int CoPb) {
  const long long nf = (long long)kk * CoP * CiP;
  const long long nb = wb ? (long long)kk * CiPb * CoPb : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nf + nb; i += (long long)gridDim.x * blockDim.x) {
    if (i < nf) {
      const int ci = (int)(i % CiP);
      const int co = (int)((i / CiP) % CoP);
      const int t = (int)(i / ((long long)CiP * CoP));
      wf[i] = __float2bfloat16_rn((co < Co && ci < Ci) ? w[((long long)co * Ci + ci) * kk + t] : 0.f);
    } else {
      const long long k = i - nf;
      const int co = (int)(k % CoPb);
      const int ci = (int)((k / CoPb) % CiPb);
      const int t = (int)(k / ((long long)CoPb * CiPb));
      wb[k] = __float2bfloat16_rn((co < Co && ci < Ci) ? w[((long long)co * Ci + ci) * kk + t] : 0.f);
    }
  }
}


// 64 x 64 (co, ci) tiles, bf16 staging, 16-byte stores: the 32 x 32 kernel above issues one 2-byte store per element and is
// instruction-bound (190 us for 135 MB of packs; 0.48 ms in situ right after the optimizer's writes).  Here a warp reads
// one output-channel row of the tile (T*kk contiguous floats) per iteration, the tile is staged as bf16 in the input order
// [co][ci][tap] (row stride padded to an odd number of 32-bit words), and every thread writes whole 8-element vectors of
// both packs: wf[tap][co][ci0 + 8g ..] and wb[tap][ci][co0 + 8g ..] -- 8 lanes cover one 128-byte line.
template <int T>
__global__ void __launch_bounds__(256)
pack_weights_tiled_kernel(const long long* __restrict__ desc, const int* __restrict__ block_map, int block0) {
  extern __shared__ __align__(16) unsigned char pk_smem[];
  __nv_bfloat16* tl = reinterpret_cast<__nv_bfloat16*>(pk_smem);   // [T co][T*kk + 2]
  const int bid = blockIdx.x + block0;
  const long long* d = desc + (long long)block_map[bid] * 12;
  const float* w = reinterpret_cast<const float*>(d[0]);
  __nv_bfloat16* wf = reinterpret_cast<__nv_bfloat16*>(d[1]);
  __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(d[2]);
  const int Co = (int)d[3], Ci = (int)d[4], kk = (int)d[5], CoP = (int)d[6], CiP = (int)d[7], CiPb = (int)d[8],
            CoPb = (int)d[9];
  const int t = bid - (int)d[10];
  const long long w_ld = d[11] ? d[11] : (long long)Ci * kk;
  const int tiles_ci = (Ci + T - 1) / T;
  const int co0 = (t / tiles_ci) * T, ci0 = (t % tiles_ci) * T;
  const int nci = min(T, Ci - ci0), nco = min(T, Co - co0);
  const int rs = T * kk + 2;            // row stride in elements: (T*kk + 2) / 2 words is odd for every kk
  const int run = nci * kk;             // contiguous floats per output channel in OIHW
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < nco; c += 8) {
    const float* src = w + (long long)(co0 + c) * w_ld + (long long)ci0 * kk;
    __nv_bfloat16* dst = tl + c * rs;
    for (int e = lane; e < run; e += 32) dst[e] = __float2bfloat16_rn(src[e]);
  }
  __syncthreads();
  constexpr int G = T / 8;              // 8-element groups per tile edge
  const int nvec = kk * T * G;
  // wf[tap][co][ci]: vector = 8 consecutive ci of one (tap, co); lanes 0..G-1 cover one row segment of the tile
  for (int idx = threadIdx.x; idx < nvec; idx += 256) {
    const int g = idx % G, c = (idx / G) % T, tap = idx / (G * T);
    if (c >= nco || g * 8 >= nci) continue;
    bf16x8 v;
    __nv_bfloat16* ve = reinterpret_cast<__nv_bfloat16*>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ci = g * 8 + j;
      ve[j] = ci < nci ? tl[c * rs + ci * kk + tap] : __float2bfloat16_rn(0.f);
    }
    *reinterpret_cast<bf16x8*>(wf + ((long long)tap * CoP + co0 + c) * CiP + ci0 + g * 8) = v;
  }
  if (wb) {
    // wb[tap][ci][co]: vector = 8 consecutive co of one (tap, ci)
    for (int idx = threadIdx.x; idx < nvec; idx += 256) {
      const int g = idx % G, ci = (idx / G) % T, tap = idx / (G * T);
      if (ci >= nci || g * 8 >= nco) continue;
      bf16x8 v;
      __nv_bfloat16* ve = reinterpret_cast<__nv_bfloat16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        ve[j] = c < nco ? tl[c * rs + ci * kk + tap] : __float2bfloat16_rn(0.f);
      }
      *reinterpret_cast<bf16x8*>(wb + ((long long)tap * CiPb + ci0 + ci) * CoPb + co0 + g * 8) = v;
    }
  }
}

// generic strided fp32 2-D copy: dst[r*ldd + c] (+)= src[r*lds + c]
__global__ void copy2d_kernel(const float* src, int lds, float* dst, int ldd, int rows, int cols, int accumulate) {
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float v = src[(long long)r * lds + c];
    float* d = dst + (long long)r * ldd + c;
    *d = accumulate ? *d + v : v;
  }
}

}  // namespace evb

using namespace evb;
#define ST ((cudaStream_t)stream)
#define LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA)

static int colsum_blocks(long long M, int C) {
  const int cg = C / 8;
  const int rows_par = kEwThreads / cg > 0 ? kEwThreads / cg : 1;
  long long b = (M + (long long)rows_par * 8 - 1) / ((long long)rows_par * 8);
  if (b > 296) b = 296;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" long long evb_bn_workspace(long long M, int C) { (void)M; return ((long long)kNbPadBwd + 1) * 2 * C * sizeof(float); }

// Training BN statistics of x[M,C] + folded scale/shift + running-stat update.
extern "C" int evb_bn_stats(const void* x, long long M, int C, const float* gamma, const float* beta, float* running_mean,
                            float* running_var, float momentum, float eps, float* mean, float* rstd, float* scale,
                            float* shift, void* ws, void* stream) {
  if (C % 8 || C > 2048) return EVB_ERR_ARG;
  const int nb = colsum_blocks(M, C);
  colsum_kernel<0><<<nb, kEwThreads, kEwThreads * 16 * sizeof(float), ST>>>(
      (const __nv_bfloat16*)x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, M, C, (float*)ws);
  bn_finalize_kernel<<<(C + 7) / 8, 256, 0, ST>>>((const float*)ws, nb, M, C, gamma, beta, running_mean, running_var,
                                                      momentum, eps, mean, rstd, scale, shift);
  return LAUNCH_OK();
}

// BN training statistics from per-CTA partial sums produced by the convolution epilogue (evb_conv2d_fwd_stats):
// partial is [2][C][kNbPad] (sum, sum of squares of the bf16-rounded conv output), nblk columns are valid.
extern "C" int evb_bn_finalize(const float* partial, int nblk, long long M, int C, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps, float* mean, float* rstd,
                               float* scale, float* shift, void* stream) {
  if (nblk < 1 || nblk > kNbPad) return EVB_ERR_ARG;
  if (evb_launch_pdl_small(bn_finalize_kernel, dim3((C + 7) / 8), dim3(256), 0, ST, partial, nblk, M, C, gamma, beta,
                           running_mean, running_var, momentum, eps, mean, rstd, scale, shift) != cudaSuccess)
    return EVB_ERR_CUDA;
  return LAUNCH_OK();
}

// nn.SyncBatchNorm (train.sync_bn, ever/trainer/th_ddp_trainer.py:21-22): statistics over the batch of ALL ranks.
// evb_bn_partial_sums: this rank's sums[2][C] from the conv epilogue's partial columns; the caller all-reduces them;
// evb_bn_finalize_sums: mean / rstd / folded scale / shift and the running statistics from global sums over M_total rows;
// evb_bn_bwd_consts: the constants of evb_norm_bwd_apply from the all-reduced backward sums.
extern "C" int evb_bn_partial_sums(const float* partial, int nblk, int C, float* sums, void* stream) {
  if (nblk < 1 || nblk > kNbPad || C < 1) return EVB_ERR_ARG;
  bn_partial_sums_kernel<<<(C + 7) / 8, 256, 0, ST>>>(partial, nblk, C, sums);
  return LAUNCH_OK();
}
extern "C" int evb_bn_finalize_sums(const float* sums, double M_total, int C, const float* gamma, const float* beta,
                                    float* running_mean, float* running_var, float momentum, float eps, float* mean,
                                    float* rstd, float* scale, float* shift, void* stream) {
  if (M_total < 1 || C < 1) return EVB_ERR_ARG;
  bn_finalize_sums_kernel<<<(C + 127) / 128, 128, 0, ST>>>(sums, M_total, C, gamma, beta, running_mean, running_var, momentum,
                                                           eps, mean, rstd, scale, shift);
  return LAUNCH_OK();
}
extern "C" int evb_bn_bwd_consts(const float* scale, const float* rstd, const float* mean, const float* dgamma,
                                 const float* dbeta, float inv_m, int C, float* c2, float* k0, void* stream) {
  if (C < 1) return EVB_ERR_ARG;
  bn_bwd_consts_kernel<<<(C + 127) / 128, 128, 0, ST>>>(scale, rstd, mean, dgamma, dbeta, inv_m, C, c2, k0);
  return LAUNCH_OK();
}

extern "C" int evb_bn_fold(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, int C,
                           float* scale, float* shift, float* mean, float* rstd, void* stream) {
  bn_fold_kernel<<<(C + 127) / 128, 128, 0, ST>>>(gamma, beta, rm, rv, eps, C, scale, shift, mean, rstd);
  return LAUNCH_OK();
}

// y = act(bf16(x*scale+shift) [+res])
static int g_bn_blocks_per_sm = 4;   // reduce-kernel grid cap (blocks per SM), <= kNbPadBwd / 148
extern "C" int evb_set_bn_reduce_blocks(int per_sm) {
  if (per_sm < 1 || per_sm * 148 > kNbPadBwd) return EVB_ERR_ARG;
  g_bn_blocks_per_sm = per_sm;
  return EVB_OK;
}
static int bn_apply_impl(const void* x, const float* scale, const float* shift, const void* res, void* y, void* mask32,
                         long long M, int C, int relu, void* stream);
extern "C" int evb_bn_apply(const void* x, const float* scale, const float* shift, const void* res, void* y, long long M,
                            int C, int relu, void* stream) {
  return bn_apply_impl(x, scale, shift, res, y, nullptr, M, C, relu, stream);
}
// same with a residual operand, also writing the ReLU survivors as one uint32 per 8-channel vector (bit j = channel j of
// the rounded output is > 0) for evb_bn_bwd(mask_mode = 3)
extern "C" int evb_bn_apply_mask(const void* x, const float* scale, const float* shift, const void* res, void* y,
                                 void* mask32, long long M, int C, void* stream) {
  if (!res || !mask32) return EVB_ERR_ARG;
  return bn_apply_impl(x, scale, shift, res, y, mask32, M, C, 1, stream);
}
static int bn_apply_impl(const void* x, const float* scale, const float* shift, const void* res, void* y, void* mask32,
                         long long M, int C, int relu, void* stream) {
  if (C % 8) return EVB_ERR_ARG;
  if (C / 8 > kEwThreads) return EVB_ERR_ARG;
  if (M * C / 8 >= (1LL << 31) - (1LL << 24)) return EVB_ERR_ARG;   // vectors are indexed with 32 bits
  {   // cp.async rings, 16-byte vectors
    const long long nvec = M * C / 8;
    const int cg_ = C / 8;
    const int bt = (kEwThreads / cg_) * cg_;
    static bool attr_set = false;
    if (!attr_set) {
      if (cudaFuncSetAttribute(bn_apply_ca_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 256 * 16) != cudaSuccess ||
          cudaFuncSetAttribute(bn_apply_ca_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 256 * 16) != cudaSuccess)
        return EVB_ERR_CUDA;
      attr_set = true;
    }
    cudaError_t e;
    if (res)
      e = evb_launch_pdl_small(bn_apply_ca_kernel<1>, dim3(ew_blocks(nvec, bt * 8, 148 * 3)), dim3(bt),
                               (size_t)kCaStages * 2 * bt * 16, ST, (const __nv_bfloat16*)x, scale, shift,
                               (const __nv_bfloat16*)res, (__nv_bfloat16*)y, (unsigned)nvec, C, relu, (unsigned*)mask32);
    else
      e = evb_launch_pdl_small(bn_apply_ca_kernel<0>, dim3(ew_blocks(nvec, bt * 8, 148 * 4)), dim3(bt),
                               (size_t)kCaStages * bt * 16, ST, (const __nv_bfloat16*)x, scale, shift,
                               (const __nv_bfloat16*)nullptr, (__nv_bfloat16*)y, (unsigned)nvec, C, relu, (unsigned*)nullptr);
    if (e != cudaSuccess) return EVB_ERR_CUDA;
    return LAUNCH_OK();
  }
  return EVB_ERR_ARG;
}

template <int MASK>
static int launch_bn_bwd_ca(const void* dy, const void* x, const void* ymask, const float* mean, const float* rstd,
                            const float* scale, const float* shift, int frozen, void* dx, void* dres, int dres_acc,
                            float* dgamma, float* dbeta, int param_acc, long long M, int C, void* ws, cudaStream_t st,
                            int phase = 0) {
  // phase 0: reduce + finalize + apply (BatchNorm).  1: reduce + finalize only (per-channel sums -> dgamma / dbeta).
  // 2: apply only, with explicit per-channel constants c2 = dgamma[], k0 = dbeta[] (frozen must be 2).
  constexpr int NT = (MASK == 1 || MASK == 3) ? 3 : 2;
  const int cg = C / 8;
  const int bt = (kEwThreads / cg) * cg;
  const int rows_par = bt / cg;
  const size_t smem = (size_t)kCaStages * NT * bt * 16;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(bn_bwd_reduce_ca_kernel<MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 256 * 16) !=
            cudaSuccess ||
        cudaFuncSetAttribute(bn_bwd_apply_ca_kernel<MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 256 * 16) !=
            cudaSuccess)
      return EVB_ERR_CUDA;
    attr_set = true;
  }
  long long nb = (M + (long long)rows_par * 16 - 1) / ((long long)rows_par * 16);   // >= 16 iterations per thread
  int per_sm = g_bn_blocks_per_sm;
  if (per_sm > (NT == 3 ? 2 : 3)) per_sm = NT == 3 ? 2 : 3;   // resident CTAs per SM with the 64 / 96 KB rings
  const int cap = per_sm * 148;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  float* fresh = ws ? (float*)ws + (size_t)kNbPadBwd * 2 * C : nullptr;
  if (phase != 2) {
    bn_bwd_reduce_ca_kernel<MASK><<<(int)nb, bt, smem, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,
                                                           (const __nv_bfloat16*)ymask, mean, scale, shift, M, C, (float*)ws);
    if (evb_launch_pdl_small(bn_bwd_finalize2_kernel, dim3((C + 7) / 8), dim3(256), 0, st, (const float*)ws, (int)nb, C, rstd,
                             dgamma, dbeta, param_acc, fresh) != cudaSuccess)
      return EVB_ERR_CUDA;
    if (phase == 1) return EVB_OK;
  }
  const long long nvec = M * C / 8;
  const float* c_dgamma = phase == 2 ? dgamma : (const float*)(fresh + C);
  const float* c_dbeta = phase == 2 ? dbeta : (const float*)fresh;
  if (evb_launch_pdl_small(bn_bwd_apply_ca_kernel<MASK>, dim3(ew_blocks(nvec, bt * 8, 148 * (NT == 3 ? 2 : 3))), dim3(bt), smem,
                           st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)ymask, mean, rstd,
                           scale, shift, c_dgamma, c_dbeta, frozen, (__nv_bfloat16*)dx,
                           (__nv_bfloat16*)dres, dres_acc, (unsigned)nvec, C, 1.0f / (float)M) != cudaSuccess)
    return EVB_ERR_CUDA;
  return EVB_OK;
}

extern "C" int evb_bn_bwd(const void* dy, const void* x, const void* ymask, const float* mean, const float* rstd,
                          const float* scale, const float* shift, int mask_mode, int frozen, void* dx, void* dres,
                          int dres_acc, float* dgamma, float* dbeta, int param_acc, long long M, int C, void* ws,
                          void* stream) {
  if (C % 8 || C > 2048 || mask_mode < 0 || mask_mode > 3) return EVB_ERR_ARG;
  if (M * C / 8 >= (1LL << 31) - (1LL << 24)) return EVB_ERR_ARG;
  int rc;
  if (mask_mode == 3) rc = launch_bn_bwd_ca<3>(dy, x, ymask, mean, rstd, scale, shift, frozen, dx, dres, dres_acc, dgamma, dbeta, param_acc, M, C, ws, ST);
  else if (mask_mode == 0) rc = launch_bn_bwd_ca<0>(dy, x, ymask, mean, rstd, scale, shift, frozen, dx, dres, dres_acc, dgamma, dbeta, param_acc, M, C, ws, ST);
  else if (mask_mode == 1) rc = launch_bn_bwd_ca<1>(dy, x, ymask, mean, rstd, scale, shift, frozen, dx, dres, dres_acc, dgamma, dbeta, param_acc, M, C, ws, ST);
  else rc = launch_bn_bwd_ca<2>(dy, x, ymask, mean, rstd, scale, shift, frozen, dx, dres, dres_acc, dgamma, dbeta, param_acc, M, C, ws, ST);
  if (rc) return rc;
  return LAUNCH_OK();
}

// The two halves of evb_bn_bwd as separate calls, for normalisations whose statistics are not per-channel batch statistics
// (GroupNorm: ever_b200 FreeNet path) and for element-wise gates (plain ReLU, squeeze-excitation):
// evb_norm_bwd_reduce: dbeta[c] (+)= sum_rows g, dgamma[c] (+)= rstd[c] * sum_rows g * (x - mean[c]), g = dy * mask.
extern "C" int evb_norm_bwd_reduce(const void* dy, const void* x, const void* ymask, const float* mean, const float* rstd,
                                   const float* scale, const float* shift, int mask_mode, float* dgamma, float* dbeta,
                                   int param_acc, long long M, int C, void* ws, void* stream) {
  if (C % 8 || C > 2048 || mask_mode < 0 || mask_mode > 2 || !dgamma || !dbeta) return EVB_ERR_ARG;
  if (M * C / 8 >= (1LL << 31) - (1LL << 24)) return EVB_ERR_ARG;
  int rc;
  if (mask_mode == 0) rc = launch_bn_bwd_ca<0>(dy, x, ymask, mean, rstd, scale, shift, 0, nullptr, nullptr, 0, dgamma, dbeta, param_acc, M, C, ws, ST, 1);
  else if (mask_mode == 1) rc = launch_bn_bwd_ca<1>(dy, x, ymask, mean, rstd, scale, shift, 0, nullptr, nullptr, 0, dgamma, dbeta, param_acc, M, C, ws, ST, 1);
  else rc = launch_bn_bwd_ca<2>(dy, x, ymask, mean, rstd, scale, shift, 0, nullptr, nullptr, 0, dgamma, dbeta, param_acc, M, C, ws, ST, 1);
  if (rc) return rc;
  return LAUNCH_OK();
}
// evb_norm_bwd_apply: dx = a[c] * (dy * mask) + k0[c] - c2[c] * x with explicit per-channel constants (a = the forward scale);
// dres (+)= dy * mask as in evb_bn_bwd.  A plain ReLU backward is a = 1, k0 = c2 = 0, mask_mode 1.
extern "C" int evb_norm_bwd_apply(const void* dy, const void* x, const void* ymask, const float* a, const float* shift,
                                  const float* c2, const float* k0, int mask_mode, void* dx, void* dres, int dres_acc,
                                  long long M, int C, void* stream) {
  if (C % 8 || C > 2048 || mask_mode < 0 || mask_mode > 2 || !a || !c2 || !k0) return EVB_ERR_ARG;
  if (M * C / 8 >= (1LL << 31) - (1LL << 24)) return EVB_ERR_ARG;
  int rc;
  float* c2_ = const_cast<float*>(c2);
  float* k0_ = const_cast<float*>(k0);
  if (mask_mode == 0) rc = launch_bn_bwd_ca<0>(dy, x, ymask, a, a, a, shift, 2, dx, dres, dres_acc, c2_, k0_, 0, M, C, nullptr, ST, 2);
  else if (mask_mode == 1) rc = launch_bn_bwd_ca<1>(dy, x, ymask, a, a, a, shift, 2, dx, dres, dres_acc, c2_, k0_, 0, M, C, nullptr, ST, 2);
  else rc = launch_bn_bwd_ca<2>(dy, x, ymask, a, a, a, shift, 2, dx, dres, dres_acc, c2_, k0_, 0, M, C, nullptr, ST, 2);
  if (rc) return rc;
  return LAUNCH_OK();
}

// per-channel sum of dy[M,C] (conv bias gradient): db[c] (+)= sum_rows dy
extern "C" int evb_bias_grad(const void* dy, long long M, int C, float* db, float* scratch_sq, int accumulate, void* ws,
                             void* stream) {
  if (C % 8 || C > 2048) return EVB_ERR_ARG;
  const int nb = colsum_blocks(M, C);
  colsum_kernel<0><<<nb, kEwThreads, kEwThreads * 16 * sizeof(float), ST>>>(
      (const __nv_bfloat16*)dy, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, M, C, (float*)ws);
  (void)scratch_sq;
  bn_bwd_finalize_kernel<<<(C + 7) / 8, 256, 0, ST>>>((const float*)ws, nb, C, nullptr, db, accumulate, nullptr);
  return LAUNCH_OK();
}

extern "C" int evb_maxpool3x3s2_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, void* stream) {
  if (C % 8 || (H & 1) || (W & 1)) return EVB_ERR_ARG;
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  maxpool_fwd_kernel<<<ew_blocks(total, kEwThreads * 2), kEwThreads, 0, ST>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y,
                                                                           (uint8_t*)idx, N, H, W, C);
  return LAUNCH_OK();
}
extern "C" int evb_maxpool3x3s2_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, void* stream) {
  const long long total = (long long)N * H * W * (C / 8);
  maxpool_bwd_kernel<<<ew_blocks(total, kEwThreads * 2), kEwThreads, 0, ST>>>((const __nv_bfloat16*)dy,
                                                                           (const uint8_t*)idx, (__nv_bfloat16*)dx, N, H, W, C);
  return LAUNCH_OK();
}

extern "C" int evb_bilinear_up(const void* x, const float* scale, const float* shift, void* y, int N, int h, int w, int C,
                               int ldx, int ldy, int f, void* stream) {
  if (C % 8 || ldx % 8 || ldy % 8 || f < 1 || f > 4) return EVB_ERR_ARG;
  const long long total = (long long)N * h * f * w * f * (C / 8);
  if (C / 8 > kEwThreads) return EVB_ERR_ARG;
  const int bt = (kEwThreads / (C / 8)) * (C / 8);
  const int cg = C / 8, Ho = h * f, Wo = w * f;
  auto lg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
  const int lc = lg(cg), lw = lg(Wo), lh = lg(Ho);
  if (lc >= 0 && lw >= 0 && lh >= 0)
    bilinear_up_kernel<true><<<ew_blocks(total, bt * 2), bt, 0, ST>>>((const __nv_bfloat16*)x, scale, shift, (__nv_bfloat16*)y, N,
                                                                    h, w, C, ldx, ldy, f, lc, lw, lh);
  else
    bilinear_up_kernel<false><<<ew_blocks(total, bt * 2), bt, 0, ST>>>((const __nv_bfloat16*)x, scale, shift, (__nv_bfloat16*)y,
                                                                     N, h, w, C, ldx, ldy, f, 0, 0, 0);
  return LAUNCH_OK();
}
extern "C" int evb_bilinear_up_bwd(const void* dy, void* dx, int N, int h, int w, int C, int lddy, int lddx, int f,
                                   void* stream) {
  if (C % 8 || lddx % 8 || lddy % 8 || f < 1 || f > 4) return EVB_ERR_ARG;
  const long long total = (long long)N * h * w * (C / 8);
  bilinear_up_bwd_kernel<<<ew_blocks(total, kEwThreads), kEwThreads, 0, ST>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx,
                                                                           N, h, w, C, lddy, lddx, f);
  return LAUNCH_OK();
}

extern "C" long long evb_bilinear_up_bwd_workspace(int N, int h, int w, int C, int f) {
  return (long long)N * h * f * w * C * sizeof(float);
}
// separable two-pass version of evb_bilinear_up_bwd (same result up to fp32 summation order); ws: fp32 [N, f*h, w, C]
extern "C" int evb_bilinear_up_bwd_sep(const void* dy, void* dx, int N, int h, int w, int C, int lddy, int lddx, int f,
                                       void* ws, long long ws_bytes, void* stream) {
  if (C % 8 || lddx % 8 || lddy % 8 || f < 1 || f > 4) return EVB_ERR_ARG;
  if (ws_bytes < evb_bilinear_up_bwd_workspace(N, h, w, C, f)) return EVB_ERR_ARG;
  const long long t1 = (long long)N * h * f * w * (C / 8), t2 = (long long)N * h * w * (C / 8);
  const int Wo = w * f, cg = C / 8;
  // narrow rows (the logit gradient: <= 64 channels) go through the row-staged kernel (x4 at 8x512^2: 81 -> 47 us for both
  // passes); for the 256-channel decoder maps the gather kernel is as fast (its taps hit L1) and stays
  if (cg <= 8 && Wo <= 3072 && (long long)N * h * f < (1LL << 31)) {   // row of dy staged in <= 48 KB of shared memory
    int ncg = 3072 / Wo;
    if (ncg > 8) ncg = 8;
    if (ncg > cg) ncg = cg;
    bilinear_bwd_x_row_kernel<<<dim3((unsigned)(N * h * f), (unsigned)((cg + ncg - 1) / ncg)), kEwThreads,
                                (size_t)Wo * ncg * 16, ST>>>((const __nv_bfloat16*)dy, (float*)ws, w, C, lddy, f, ncg);
  } else {
    bilinear_bwd_x_kernel<<<ew_blocks(t1, kEwThreads * 2), kEwThreads, 0, ST>>>((const __nv_bfloat16*)dy, (float*)ws, N, h, w,
                                                                              C, lddy, f);
  }
  bilinear_bwd_y_kernel<<<ew_blocks(t2, kEwThreads), kEwThreads, 0, ST>>>((const float*)ws, (__nv_bfloat16*)dx, N, h, w, C, lddx,
                                                                        f);
  return LAUNCH_OK();
}

extern "C" int evb_sumpool2(const void* dfine, void* dcoarse, int N, int h, int w, int C, int accumulate, void* stream) {
  const long long total = (long long)N * h * w * (C / 8);
  sumpool2_kernel<<<ew_blocks(total, kEwThreads * 2), kEwThreads, 0, ST>>>((const __nv_bfloat16*)dfine,
                                                                        (__nv_bfloat16*)dcoarse, N, h, w, C, accumulate);
  return LAUNCH_OK();
}

extern "C" int evb_merge4(const void* a, const void* b, const void* c, const void* d, void* out, long long numel,
                          void* stream) {
  const long long nvec = numel / 8;
  merge4_kernel<<<ew_blocks(nvec, kEwThreads * 4), kEwThreads, 0, ST>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                                     (const __nv_bfloat16*)c, (const __nv_bfloat16*)d,
                                                                     (__nv_bfloat16*)out, nvec);
  return LAUNCH_OK();
}

extern "C" int evb_scale_add(const void* x, float alpha, const void* z, void* y, long long numel, void* stream) {
  const long long nvec = numel / 8;
  scale_add_kernel<<<ew_blocks(nvec, kEwThreads * 4), kEwThreads, 0, ST>>>((const __nv_bfloat16*)x, alpha,
                                                                        (const __nv_bfloat16*)z, (__nv_bfloat16*)y, nvec);
  return LAUNCH_OK();
}

extern "C" int evb_gap_fwd(const void* x, float* out, int N, int HW, int C, void* stream) {
  dim3 grid((C + 127) / 128, N);
  gap_fwd_kernel<<<grid, 128, 0, ST>>>((const __nv_bfloat16*)x, out, HW, C);
  return LAUNCH_OK();
}
extern "C" int evb_gap_bwd(const float* dscene, void* dx, int N, int HW, int C, void* stream) {
  const long long total = (long long)N * HW * C;
  gap_bwd_kernel<<<ew_blocks(total, kEwThreads * 4), kEwThreads, 0, ST>>>(dscene, (__nv_bfloat16*)dx, HW, C, total);
  return LAUNCH_OK();
}

namespace evb {
// Fast im2col for small windows (KP / 8 <= 256): blockDim is a multiple of KP/8, so a thread owns one 8-element slot of the
// K axis for every pixel it visits: the (channel, row, column) of its 8 taps, their offsets from the window origin and
// their validity tests are computed once; interior pixels (whole window inside the image) take a path without bounds
// tests.  The generic kernel spent ~25 integer instructions per element on div/mod and was instruction-bound (214 us for
// the 201 MB stem matrix; the write alone takes ~35 us).
template <bool U8>
__global__ void __launch_bounds__(kEwThreads)
stem_im2col_fast_kernel(const void* __restrict__ xin, const float* __restrict__ mean, const float* __restrict__ stdv,
                        __nv_bfloat16* __restrict__ a, int N, int Cin, int H, int W, int KP, int ks, int stride, int pad) {
  const int Ho = H / stride, Wo = W / stride;
  const int kg = KP / 8;
  const int kk = ks * ks, K = Cin * kk;
  const int g = threadIdx.x % kg;
  const int psub = threadIdx.x / kg, ppb = blockDim.x / kg;   // pixel lane within the block, pixels per block iteration
  int off[8], rr[8], ss[8];
  float mu[8], is[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = g * 8 + j;
    if (k < K) {
      const int c = k / kk, rs = k % kk;
      rr[j] = rs / ks - pad;
      ss[j] = rs % ks - pad;
      off[j] = U8 ? (rr[j] * W + ss[j]) * Cin + c : c * H * W + rr[j] * W + ss[j];
      if (U8) { mu[j] = mean[c]; is[j] = stdv[c]; }
    } else {
      rr[j] = 1 << 20;   // never valid
      ss[j] = 0;
      off[j] = 0;
      if (U8) { mu[j] = 0.f; is[j] = 1.f; }
    }
  }
  const int lo = pad, hi_h = H - (ks - 1 - pad), hi_w = W - (ks - 1 - pad);   // window fully inside: lo <= h0 < hi
  const long long npix = (long long)N * Ho * Wo;
  const bool tail = g * 8 + 8 > K;   // slot touches the zero padding of the K axis
  for (long long p = (long long)blockIdx.x * ppb + psub; p < npix; p += (long long)gridDim.x * ppb) {
    const int wo = (int)(p % Wo);
    const long long t = p / Wo;
    const int ho = (int)(t % Ho), n = (int)(t / Ho);
    const int h0 = ho * stride, w0 = wo * stride;
    const long long base = U8 ? (((long long)n * H + h0) * W + w0) * Cin : ((long long)n * Cin * H + h0) * W + w0;
    float v[8];
    if (!tail && h0 >= lo && h0 < hi_h && w0 >= lo && w0 < hi_w) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (U8) v[j] = ((float)__ldg(reinterpret_cast<const uint8_t*>(xin) + base + off[j]) - mu[j]) / is[j];
        else v[j] = __ldg(reinterpret_cast<const float*>(xin) + base + off[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int h = h0 + rr[j], w = w0 + ss[j];
        float val = 0.f;
        if (h >= 0 && h < H && w >= 0 && w < W) {
          if (U8) val = ((float)__ldg(reinterpret_cast<const uint8_t*>(xin) + base + off[j]) - mu[j]) / is[j];
          else val = __ldg(reinterpret_cast<const float*>(xin) + base + off[j]);
        }
        v[j] = val;
      }
    }
    reinterpret_cast<bf16x8*>(a)[p * kg + g] = pack8(v);
  }
}
}  // namespace evb

static int im2col_check(int N, int Cin, int H, int W, int KP, int ks, int stride, int pad) {
  if (KP % 8 || ks < 1 || ks > 7 || stride < 1 || stride > 2 || pad < 0 || pad > 3 || KP < Cin * ks * ks) return EVB_ERR_ARG;
  if (H % stride || W % stride) return EVB_ERR_ARG;
  if ((long long)N * (H / stride) * (W / stride) * (KP / 8) >= (1LL << 32)) return EVB_ERR_ARG;
  return EVB_OK;
}
// general window: A[N*(H/stride)*(W/stride)][KP] from NCHW fp32 (ks x ks, stride 1|2, pad)
extern "C" int evb_im2col_nchw(const float* x, void* a, int N, int Cin, int H, int W, int KP, int ks, int stride, int pad,
                               void* stream) {
  if (im2col_check(N, Cin, H, W, KP, ks, stride, pad)) return EVB_ERR_ARG;
  const long long total = (long long)N * (H / stride) * (W / stride) * (KP / 8);
  const int kg = KP / 8;
  if (kg <= kEwThreads && (long long)Cin * H * W < (1LL << 30)) {
    const int bt = (kEwThreads / kg) * kg;
    const long long npix = total / kg;
    stem_im2col_fast_kernel<false><<<ew_blocks(npix, (bt / kg) * 4), bt, 0, ST>>>(x, nullptr, nullptr, (__nv_bfloat16*)a, N, Cin,
                                                                               H, W, KP, ks, stride, pad);
    return LAUNCH_OK();
  }
  stem_im2col_kernel<<<ew_blocks(total, kEwThreads), kEwThreads, 0, ST>>>(x, (__nv_bfloat16*)a, N, Cin, H, W, KP, ks, stride,
                                                                       pad);
  return LAUNCH_OK();
}
extern "C" int evb_im2col_u8(const void* x, const float* mean, const float* stdv, void* a, int N, int Cin, int H, int W,
                             int KP, int ks, int stride, int pad, void* stream) {
  if (im2col_check(N, Cin, H, W, KP, ks, stride, pad)) return EVB_ERR_ARG;
  const long long total = (long long)N * (H / stride) * (W / stride) * (KP / 8);
  const int kg = KP / 8;
  if (kg <= kEwThreads && (long long)Cin * H * W < (1LL << 30)) {
    const int bt = (kEwThreads / kg) * kg;
    const long long npix = total / kg;
    stem_im2col_fast_kernel<true><<<ew_blocks(npix, (bt / kg) * 4), bt, 0, ST>>>(x, mean, stdv, (__nv_bfloat16*)a, N, Cin, H, W,
                                                                              KP, ks, stride, pad);
    return LAUNCH_OK();
  }
  stem_im2col_u8_kernel<<<ew_blocks(total, kEwThreads), kEwThreads, 0, ST>>>((const uint8_t*)x, mean, stdv, (__nv_bfloat16*)a,
                                                                          N, Cin, H, W, KP, ks, stride, pad);
  return LAUNCH_OK();
}
// the 7x7 stride-2 pad-3 stem (ever/module/_resnets.py:149-150)
extern "C" int evb_stem_im2col(const float* x, void* a, int N, int Cin, int H, int W, int KP, void* stream) {
  return evb_im2col_nchw(x, a, N, Cin, H, W, KP, 7, 2, 3, stream);
}
extern "C" int evb_stem_im2col_u8(const void* x, const float* mean, const float* stdv, void* a, int N, int Cin, int H, int W,
                                  int KP, void* stream) {
  return evb_im2col_u8(x, mean, stdv, a, N, Cin, H, W, KP, 7, 2, 3, stream);
}
extern "C" int evb_confusion_matrix(const void* pred, const void* labels, long long P, int K, void* cm, void* stream) {
  if (K < 1 || K > 64) return EVB_ERR_ARG;
  long long b = (P + 256 * 16 - 1) / (256 * 16);
  if (b > 148 * 4) b = 148 * 4;
  if (b < 1) b = 1;
  confusion_matrix_kernel<<<(int)b, 256, K * K * sizeof(unsigned int), ST>>>((const uint8_t*)pred, (const long long*)labels, P, K,
                                                                             (unsigned long long*)cm);
  return LAUNCH_OK();
}

extern "C" int evb_pack_weight(const float* w, int Co, int Ci, int kk, void* wf, int CoP, int CiP, void* wb, int CiPb,
                               int CoPb, void* stream) {
  const long long total = (long long)kk * CoP * CiP + (wb ? (long long)kk * CiPb * CoPb : 0);
  pack_weight_kernel<<<ew_blocks(total, kEwThreads * 4), kEwThreads, 0, ST>>>(w, Co, Ci, kk, (__nv_bfloat16*)wf, CoP, CiP,
                                                                           (__nv_bfloat16*)wb, CiPb, CoPb);
  return LAUNCH_OK();
}

extern "C" int evb_pack_weights_tiled(const void* desc, const void* block_map, int block0, int nblocks, void* stream) {
  if (nblocks <= 0) return EVB_OK;
  constexpr int T = 64;
  const size_t smem = (size_t)T * (T * 9 + 2) * sizeof(__nv_bfloat16);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(pack_weights_tiled_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return EVB_ERR_CUDA;
    attr_set = true;
  }
  pack_weights_tiled_kernel<T><<<nblocks, 256, smem, ST>>>((const long long*)desc, (const int*)block_map, block0);
  return LAUNCH_OK();
}
// blocks [block0, block0 + nblocks) of the same table: lets the caller pack the first layers' weights on the main stream
// and the rest on a second stream that overlaps the start of the forward pass
// zero-fill as a memset node (no kernel): gradient buffers whose first writer accumulates
extern "C" int evb_zero_bytes(void* dst, long long nbytes, void* stream) {
  if (nbytes < 0) return EVB_ERR_ARG;
  if (nbytes == 0) return EVB_OK;
  return cudaMemsetAsync(dst, 0, (size_t)nbytes, ST) == cudaSuccess ? EVB_OK : EVB_ERR_CUDA;
}
extern "C" int evb_copy2d_f32(const float* src, int lds, float* dst, int ldd, int rows, int cols, int accumulate,
                              void* stream) {
  copy2d_kernel<<<ew_blocks((long long)rows * cols, kEwThreads), kEwThreads, 0, ST>>>(src, lds, dst, ldd, rows, cols,
                                                                                   accumulate);
  return LAUNCH_OK();
}

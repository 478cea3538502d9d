// Convolution weight gradient on tcgen05 tensor cores (sm_100a).
//
//   dW[tap][ci, co] = sum over pixels  X[pix + tap offset, ci] * dY[pix, co]
//
// Both operands are "MN-major" for the MMA: the reduction (K) dimension is the pixel index and the
// contiguous dimension in HBM/smem is the channel, so the NHWC tiles TMA delivers ([pixels][64 ch], 128 B
// rows, 128B swizzle) are consumed directly, no transpose.  M = 128 input channels, N = NT output channels,
// K = 64 pixels per pipeline stage (4 MMAs of K=16).  One CTA = (pixel split, tap, ci tile, co tile); fp32
// partials per split go to a workspace and a second kernel reduces the splits in a fixed order
// (deterministic) into the OIHW fp32 gradient.
//
// Replaces (reference): the weight-gradient half of every nn.Conv2d backward on the FarSeg path
// (aten::convolution_backward -> cuDNN wgrad), ever/module/_resnets.py:21-29, fpn.py:165,179 etc.
#include <cstdlib>
#include "common.cuh"

namespace evb {

struct WTap {
  int dw, dh, coff, phase, slab;
};

struct WgradParams {
  int ntaps;
  WTap taps[9];
  int bw, bh, bn;                 // pixel chunk box (product 64)
  int tiles_w, tiles_h, tiles_n;  // chunk grid
  int nchunks, chunks_per_split, nsplit;
  int tiles_ci, tiles_co;
  float* ws;   // deterministic mode: [split][tap][CinP][CoutP] partials, reduced by a second kernel
  int CinP, CoutP, Cin, Cout;
};

// MT = number of 128-row input-channel tiles per CTA.  MT = 2 keeps two accumulators in TMEM that share every dY (B)
// tile: one third less L2->SM traffic per FLOP (the wgrad main loop is L2-fed, see profiles/).
template <int NT, int MT>
struct WgradCfg {
  static constexpr int A_BYTES = MT * 2 * 64 * 128;
  static constexpr int B_BYTES = (NT / 64) * 64 * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = (STAGE >= 65536) ? 3 : (STAGE >= 49152) ? 4 : (STAGE >= 32768) ? 6 : 8;
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
  static constexpr int TMEM_COLS = NT * MT;
};

template <int NT, int MT>
__global__ void __launch_bounds__(192, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
             const __grid_constant__ WgradParams p) {
  using Cfg = WgradCfg<NT, MT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* tfull = empty + Cfg::STAGES;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int tco = t % p.tiles_co; t /= p.tiles_co;
  const int tci = t % p.tiles_ci; t /= p.tiles_ci;
  const int tap = t % p.ntaps;
  const int split = t / p.ntaps;
  const int ci0 = tci * 128 * MT, co0 = tco * NT;
  const int c_begin = split * p.chunks_per_split;
  const int c_end = min(p.nchunks, c_begin + p.chunks_per_split);
  const int niter = max(0, c_end - c_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tslot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = *tslot;
  pdl_wait();   // prologue above overlaps the predecessor's tail

  if (warp == 0) {
    if (lane == 0) {
      const WTap tp = p.taps[tap];
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        int cc = c;
        const int tw = cc % p.tiles_w; cc /= p.tiles_w;
        const int th = cc % p.tiles_h;
        const int tn = cc / p.tiles_h;
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        mbar_wait(&empty[stage], phase ^ 1, 0x400 + stage);
        uint8_t* a_dst = smem + stage * Cfg::STAGE;
        mbar_arrive_expect_tx(&full[stage], Cfg::STAGE);
#pragma unroll
        for (int s = 0; s < 2 * MT; ++s)
          tma_load_5d(a_dst + s * 8192, &tmX, &full[stage], tp.coff + ci0 + s * 64, w0 + tp.dw, h0 + tp.dh, n0, tp.phase);
#pragma unroll
        for (int s = 0; s < NT / 64; ++s)
          tma_load_5d(a_dst + Cfg::A_BYTES + s * 8192, &tmDY, &full[stage], co0 + s * 64, w0, h0, n0, 0);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, NT, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < niter; ++it) {
        mbar_wait(&full[stage], phase, 0x500 + stage);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE);
        const uint32_t b_base = a_base + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // 16 pixels (K) per MMA = 2 swizzle atoms of 8 rows x 128 B; LBO = 64-channel slab stride
          const uint64_t bdesc = make_smem_desc(b_base + k * 2048, 8192, 1024);
#pragma unroll
          for (int m = 0; m < MT; ++m)
            umma_bf16(taddr + m * NT, make_smem_desc(a_base + m * 16384 + k * 2048, 8192, 1024), bdesc, idesc,
                      (it | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;  // input channel within the tile
    mbar_wait(tfull, 0, 0x600);
    tc_fence_after();
#pragma unroll 1
    for (int m = 0; m < MT; ++m) {
      const int ci = ci0 + m * 128 + r;
      const uint32_t tm = taddr + (uint32_t(q * 32) << 16) + m * NT;
      {
        float* orow = p.ws + (((long long)split * p.ntaps + tap) * p.CinP + ci) * p.CoutP + co0;
#pragma unroll 1
        for (int c = 0; c < NT / 32; ++c) {
          uint32_t v[32];
          if (niter > 0) {
            tmem_ld32(tm + c * 32, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0;
          }
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(orow + c * 32 + g * 4) = make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(taddr, Cfg::TMEM_COLS);
  }
}

// dW[co][ci][tap] (OIHW fp32) = (accumulate ? dW : 0) + sum_split ws[split][tap][ci][co]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int nsplit, int ntaps, int Cin,
                                    int Cout, int CinP, int CoutP, int accumulate) {
  __shared__ float tile[32][33];
  const int tap = blockIdx.z;
  const int co_b = blockIdx.x * 32, ci_b = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int ci = ci_b + i, co = co_b + tx;
    float s = 0.f;
    if (ci < Cin && co < Cout) {
      const float* src = ws + ((long long)tap * CinP + ci) * CoutP + co;
      const long long ss = (long long)ntaps * CinP * CoutP;
      float s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int k = 0;
      for (; k + 3 < nsplit; k += 4) {
        s += src[k * ss]; s1 += src[(k + 1) * ss]; s2 += src[(k + 2) * ss]; s3 += src[(k + 3) * ss];
      }
      for (; k < nsplit; ++k) s += src[k * ss];
      s += s1 + s2 + s3;
    }
    tile[i][tx] = s;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int co = co_b + i, ci = ci_b + tx;
    if (ci < Cin && co < Cout) {
      float* dst = dw + ((long long)co * Cin + ci) * ntaps + tap;
      const float v = tile[tx][i];
      *dst = accumulate ? (*dst + v) : v;
    }
  }
}

// Coalesced deterministic reduce + transpose to OIHW.  256 threads: tile of 32 co x CIT ci x all taps.
//   read : ws[split][tap][ci][co] as float4 along co, splits summed in fixed order with 8 loads in flight
//   write: dw[co][ci][tap]        -- for each co, CIT ci x NTAPS contiguous floats
template <int NTAPS, int CIT>
__global__ void __launch_bounds__(256)
wgrad_reduce_oihw_kernel(const float* __restrict__ ws, float* __restrict__ dw, int nsplit, int Cin, int Cout, int CinP,
                         int CoutP, int accumulate) {
  __shared__ float tile[NTAPS][CIT][33];
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * CIT;
  const long long ss = (long long)NTAPS * CinP * CoutP;
  constexpr int ITEMS = NTAPS * CIT * 8;
  for (int it = threadIdx.x; it < ITEMS; it += 256) {
    const int quad = it & 7, cil = (it >> 3) % CIT, tap = it / (8 * CIT);
    const int ci = ci0 + cil, co = co0 + quad * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ci < Cin && co < Cout) {   // CoutP is a multiple of 64: the float4 never leaves the padded row
      const float* src = ws + ((long long)tap * CinP + ci) * CoutP + co;
      int k = 0;
      for (; k + 7 < nsplit; k += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(src + (k + u) * ss);
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      }
      for (; k < nsplit; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(src + k * ss);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    tile[tap][cil][quad * 4 + 0] = acc.x;
    tile[tap][cil][quad * 4 + 1] = acc.y;
    tile[tap][cil][quad * 4 + 2] = acc.z;
    tile[tap][cil][quad * 4 + 3] = acc.w;
  }
  __syncthreads();
  constexpr int RUN = CIT * NTAPS;  // contiguous output floats per co
  for (int idx = threadIdx.x; idx < 32 * RUN; idx += 256) {
    const int c = idx / RUN, e = idx % RUN;
    const int cil = e / NTAPS, tap = e % NTAPS;
    const int oco = co0 + c, oci = ci0 + cil;
    if (oco < Cout && oci < Cin) {
      float* dst = dw + ((long long)oco * Cin + oci) * NTAPS + tap;
      const float v = tile[tap][cil][c];
      *dst = accumulate ? (*dst + v) : v;
    }
  }
}

// split-parallel variant for many splits / few outputs: block (32 co, 8 split lanes), one ci per blockIdx.y
__global__ void wgrad_reduce_splitpar_kernel(const float* __restrict__ ws, float* __restrict__ dw, int nsplit, int ntaps,
                                             int Cin, int Cout, int CinP, int CoutP, int accumulate) {
  __shared__ float red[8][33];
  const int tap = blockIdx.z, ci = blockIdx.y;
  const int co = blockIdx.x * 32 + threadIdx.x;
  const long long ss = (long long)ntaps * CinP * CoutP;
  float s = 0.f;
  if (co < Cout) {
    const float* src = ws + ((long long)tap * CinP + ci) * CoutP + co;
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = threadIdx.y;
    for (; k + 24 < nsplit; k += 32) {  // 4 independent loads in flight
      s += src[k * ss]; s1 += src[(k + 8) * ss]; s2 += src[(k + 16) * ss]; s3 += src[(k + 24) * ss];
    }
    for (; k < nsplit; k += 8) s += src[k * ss];
    s += s1 + s2 + s3;
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && co < Cout) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    float* dst = dw + ((long long)co * Cin + ci) * ntaps + tap;
    *dst = accumulate ? (*dst + t) : t;
  }
}

static int pow2_le(int x, int cap) {
  int r = 1;
  while (r * 2 <= x && r * 2 <= cap) r *= 2;
  return r;
}

static int make_act_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int stride, int bw, int bh, int bn) {
  uint64_t dims[5], strides[4];
  uint32_t box[5] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn, 1};
  if (stride == 1) {
    dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = N; dims[4] = 1;
    strides[0] = (uint64_t)C * 2; strides[1] = (uint64_t)W * C * 2; strides[2] = (uint64_t)H * W * C * 2;
    strides[3] = (uint64_t)N * H * W * C * 2;
  } else {
    if ((H & 1) || (W & 1)) return EVB_ERR_ARG;
    dims[0] = 2 * (uint64_t)C; dims[1] = W / 2; dims[2] = H / 2; dims[3] = N; dims[4] = 2;
    strides[0] = (uint64_t)2 * C * 2; strides[1] = (uint64_t)2 * W * C * 2;
    strides[2] = (uint64_t)H * W * C * 2; strides[3] = (uint64_t)W * C * 2;
  }
  return evb_make_tmap_bf16(m, ptr, 5, dims, strides, box);
}

template <int NT, int MT>
static int launch_wgrad(const CUtensorMap& a, const CUtensorMap& b, const WgradParams& p, cudaStream_t st) {
  using Cfg = WgradCfg<NT, MT>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_kernel<NT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess)
      return EVB_ERR_CUDA;
    attr_set = true;
  }
  const int grid = p.nsplit * p.ntaps * p.tiles_ci * p.tiles_co;
  if (evb_launch_pdl(wgrad_kernel<NT, MT>, dim3(grid), dim3(192), Cfg::SMEM, st, a, b, p) != cudaSuccess) return EVB_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA;
}

struct WgradPlan {
  int nt, mt, nsplit, cps, CinP, CoutP, tiles_ci, tiles_co, nchunks, bw, bh, bn, tw, th, tn;
  size_t ws_bytes;
};

// CTAs a weight-gradient launch aims for.  A full wave (148) is best for the kernel alone, but the weight gradients run on a
// side stream NEXT TO the dgrad / BatchNorm chain: half a wave means half as many split-K partials to write and re-read (the
// kernel is L2-fed) and leaves the other SMs to the critical path.  Measured in situ on the C2 step (bench.py, 200 steps,
// same box): 148 -> 9.08 / 9.13 ms, 111 -> 8.89, 96 -> 8.87 / 9.00, 84 -> 8.97, 74 -> 8.83 / 8.87, 64 -> 9.04, 56 -> 9.01.
static int wgrad_wave() {
  static int wave = 0;
  if (!wave) {
    const char* e = getenv("EVB_WGRAD_WAVE");
    wave = e ? atoi(e) : 74;
    if (wave < 16 || wave > 592) wave = 74;
  }
  return wave;
}

static WgradPlan plan_wgrad(int N, int Ho, int Wo, int Cin, int Cout, int ntaps, int force_nt, int force_split,
                            int allow_mt2 = 1) {
  WgradPlan pl{};
  const int wave = wgrad_wave();
  pl.nt = force_nt ? force_nt : (Cout >= 256 ? 256 : (Cout >= 128 ? 128 : 64));
  pl.tiles_co = (Cout + pl.nt - 1) / pl.nt;
  pl.CoutP = pl.tiles_co * pl.nt;
  pl.bw = pow2_le(Wo, 64);
  pl.bh = pow2_le(Ho, 64 / pl.bw);
  pl.bn = 64 / (pl.bw * pl.bh);
  pl.tw = (Wo + pl.bw - 1) / pl.bw;
  pl.th = (Ho + pl.bh - 1) / pl.bh;
  pl.tn = (N + pl.bn - 1) / pl.bn;
  pl.nchunks = pl.tw * pl.th * pl.tn;
  // two M tiles per CTA when the channel count allows it and a full wave of CTAs can still be formed
  pl.mt = 1;
  if (allow_mt2 && pl.nt >= 128 && Cin >= 256) {
    const int base2 = ntaps * ((Cin + 255) / 256) * pl.tiles_co;
    int sp = wave / base2 > 0 ? wave / base2 : 1;
    if (sp > (pl.nchunks + 7) / 8) sp = (pl.nchunks + 7) / 8;
    if (base2 * sp >= wave * 3 / 4) pl.mt = 2;
  }
  pl.tiles_ci = (Cin + 128 * pl.mt - 1) / (128 * pl.mt);
  pl.CinP = pl.tiles_ci * 128 * pl.mt;
  const int base = ntaps * pl.tiles_ci * pl.tiles_co;
  int split = force_split ? force_split : wave / base;  // round down: at most one wave of CTAs
  if (split > pl.nchunks) split = pl.nchunks;
  if (!force_split && split > 74) split = 74;   // bounds the reduction depth (and workspace) of small-channel convs
  if (!force_split) {
    const int max_by_work = (pl.nchunks + 7) / 8;  // at least ~8 chunks (512 pixels) per CTA (16 / 32 measured no better)
    if (split > max_by_work) split = max_by_work;
  }
  if (split < 1) split = 1;
  pl.cps = (pl.nchunks + split - 1) / split;
  pl.nsplit = (pl.nchunks + pl.cps - 1) / pl.cps;
  pl.ws_bytes = (size_t)pl.nsplit * ntaps * pl.CinP * pl.CoutP * sizeof(float);
  return pl;
}

}  // namespace evb

using namespace evb;

// Workspace bytes needed by evb_conv2d_wgrad for this problem (fp32 split-K partials).
extern "C" long long evb_conv2d_wgrad_workspace(int N, int Ho, int Wo, int Cin, int Cout, int ksize, int force_nt,
                                                int force_split) {
  return (long long)plan_wgrad(N, Ho, Wo, Cin, Cout, ksize * ksize, force_nt, force_split).ws_bytes;
}

// dw[Cout][Cin][k][k] fp32 (+)= sum_pixels x (*) dy.  x: [N,H,W,Cin] bf16 (the conv input), dy: [N,Ho,Wo,Cout] bf16.
static int wgrad_impl(const void* x, int N, int H, int W, int Cin, const void* dy, int Cout, int ksize, int stride,
                      float* dw, int accumulate, void* ws, long long ws_bytes, int force_nt, int force_split, void* stream) {
  if ((ksize != 1 && ksize != 3) || (stride != 1 && stride != 2)) return EVB_ERR_ARG;
  if (Cin % 64 || Cout % 64) return EVB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = H / stride, Wo = W / stride, ntaps = ksize * ksize;
  const WgradPlan pl = plan_wgrad(N, Ho, Wo, Cin, Cout, ntaps, force_nt, force_split, 1);
  if ((size_t)ws_bytes < pl.ws_bytes) return EVB_ERR_ARG;
  WgradParams p{};
  p.ntaps = ntaps;
  for (int r = 0; r < ksize; ++r)
    for (int s = 0; s < ksize; ++s) {
      WTap& t = p.taps[r * ksize + s];
      const int oh = r - ksize / 2, ow = s - ksize / 2;
      t.slab = r * ksize + s;
      if (stride == 1) {
        t.dh = oh; t.dw = ow; t.coff = 0; t.phase = 0;
      } else {
        const int ph = oh & 1, pw = ow & 1;
        t.dh = (oh - ph) / 2; t.dw = (ow - pw) / 2; t.phase = ph; t.coff = pw * Cin;
      }
    }
  p.bw = pl.bw; p.bh = pl.bh; p.bn = pl.bn;
  p.tiles_w = pl.tw; p.tiles_h = pl.th; p.tiles_n = pl.tn;
  p.nchunks = pl.nchunks; p.chunks_per_split = pl.cps; p.nsplit = pl.nsplit;
  p.tiles_ci = pl.tiles_ci; p.tiles_co = pl.tiles_co;
  p.ws = (float*)ws; p.CinP = pl.CinP; p.CoutP = pl.CoutP;
  p.Cin = Cin; p.Cout = Cout;
  CUtensorMap tmX, tmDY;
  int rc = make_act_map(&tmX, x, N, H, W, Cin, stride, pl.bw, pl.bh, pl.bn);
  if (rc) return rc;
  rc = make_act_map(&tmDY, dy, N, Ho, Wo, Cout, 1, pl.bw, pl.bh, pl.bn);
  if (rc) return rc;
  switch (pl.nt * 10 + pl.mt) {
    case 2562: rc = launch_wgrad<256, 2>(tmX, tmDY, p, st); break;
    case 2561: rc = launch_wgrad<256, 1>(tmX, tmDY, p, st); break;
    case 1282: rc = launch_wgrad<128, 2>(tmX, tmDY, p, st); break;
    case 1281: rc = launch_wgrad<128, 1>(tmX, tmDY, p, st); break;
    case 641: rc = launch_wgrad<64, 1>(tmX, tmDY, p, st); break;
    default: rc = EVB_ERR_ARG;
  }
  if (rc) return rc;
  if (pl.nsplit > 16 && (long long)Cin * Cout * ntaps <= 65536) {   // few outputs, many splits
    dim3 grid((Cout + 31) / 32, Cin, ntaps), block(32, 8);
    wgrad_reduce_splitpar_kernel<<<grid, block, 0, st>>>((const float*)ws, dw, pl.nsplit, ntaps, Cin, Cout, pl.CinP,
                                                         pl.CoutP, accumulate);
  } else {
    if (ntaps == 9) {
      dim3 grid((Cout + 31) / 32, (Cin + 7) / 8);
      wgrad_reduce_oihw_kernel<9, 8><<<grid, 256, 0, st>>>((const float*)ws, dw, pl.nsplit, Cin, Cout, pl.CinP, pl.CoutP,
                                                           accumulate);
    } else {
      dim3 grid((Cout + 31) / 32, (Cin + 31) / 32);
      wgrad_reduce_oihw_kernel<1, 32><<<grid, 256, 0, st>>>((const float*)ws, dw, pl.nsplit, Cin, Cout, pl.CinP, pl.CoutP,
                                                            accumulate);
    }
  }
  return cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA;
}

// dw[Cout][Cin][k][k] fp32 (+)= sum_pixels x (*) dy, deterministic: split-K partials in `ws`, fixed-order reduction.
extern "C" int evb_conv2d_wgrad(const void* x, int N, int H, int W, int Cin, const void* dy, int Cout, int ksize,
                                int stride, float* dw, int accumulate, void* ws, long long ws_bytes, int force_nt,
                                int force_split, void* stream) {
  return wgrad_impl(x, N, H, W, Cin, dy, Cout, ksize, stride, dw, accumulate, ws, ws_bytes, force_nt, force_split, stream);
}

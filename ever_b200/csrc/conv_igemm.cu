// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC bf16 activations.
//
//   D[128 output pixels, NT output channels] = sum over taps, channel blocks of
//        A[128 pixels, 64 ch]  (TMA 5-D box of the activation tensor, shifted by the tap, OOB -> 0 = padding)
//      x B[NT out-ch, 64 ch]   (TMA 3-D box of the packed weights [tap][Cout][Cin])
//
// One CTA per (pixel tile, out-channel tile).  Warp 0 = TMA producer, warp 1 = tcgen05.mma issuer
// (single elected thread) + TMEM allocator, warps 2-5 = epilogue (tcgen05.ld -> bias / residual add ->
// bf16 -> global).  fp32 accumulators live in TMEM.  Forward, stride-1 dgrad (flipped taps on the
// transposed weight pack), stride-2 forward (5-D "phase" view of the input) and stride-2 dgrad (one launch
// per output phase) are all the same kernel with a different tap table.
//
// Replaces (reference): every nn.Conv2d forward / input-gradient on the FarSeg path, i.e.
// ever/module/_resnets.py:21-29,139-150, ever/module/ops.py:53-55, ever/module/fpn.py:165,179,
// ever/module/fs_relation.py:25-27,42,49 (cuDNN through ATen in the reference).
#include "common.cuh"

namespace evb {

struct ConvTap {
  int dw, dh;   // pixel offset of the A box for this tap
  int coff;     // channel offset inside dim 0 of the A map (phase view: q * C)
  int phase;    // coordinate on dim 4 of the A map
  int slab;     // which [Cout][Cin] weight slab
};
constexpr int kMaxTaps = 9;

struct IgemmParams {
  int ntaps, kblocks;
  ConvTap taps[kMaxTaps];
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n, tiles_c;
  __nv_bfloat16* out;
  long long o_sn, o_sh, o_sw;
  int No, Ho, Wo, Cout;
  const float* bias;
  const __nv_bfloat16* add;
  long long a_sn, a_sh, a_sw;
  int add_shift, add_mode;
  float* stats;   // STATS kernels: BN partial sums [2][Cout][kNbPad], column = blockIdx / tiles_c
};

// ------------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM loops over output tiles.  The fp32 accumulator is double-buffered in TMEM
// (2 x NT columns) so the epilogue of tile i (tcgen05.ld -> bias/add -> bf16 -> 128B-swizzled smem -> TMA store)
// overlaps the TMA/MMA main loop of tile i+1; the smem ring keeps streaming across tile boundaries.
// ------------------------------------------------------------------------------------------------------------
template <int NT>
struct Igemm2Cfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = NT * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = (NT == 256) ? 4 : (NT == 128) ? 4 : 5;
  static constexpr int SLAB = 128 * 128;  // 128 rows x 64 channels bf16, one TMA store box
  static constexpr int NSLAB_G = (NT == 256) ? 1 : 2;  // staging buffers per epilogue group
  static constexpr int SMEM = STAGES * STAGE + 2 * NSLAB_G * SLAB + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * NT;
  static constexpr int THREADS = 64 + 256;  // producer warp, MMA warp, 2 epilogue groups of 4 warps
};

// EPI: the epilogue applies a per-channel bias and/or adds a second tensor (compile-time so the plain path is branch-free)
// STATS: the epilogue also accumulates per-channel sum / sum of squares of the bf16-rounded outputs (BatchNorm batch
//        statistics, fused: no separate read pass over the conv output); requires gridDim.x % tiles_c == 0.
template <int NT, bool EPI, bool STATS>
__global__ void __launch_bounds__(320, 1)
igemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ CUtensorMap tmC, const __grid_constant__ IgemmParams p) {
  using Cfg = Igemm2Cfg<NT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* slab_base = smem + Cfg::STAGES * Cfg::STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(slab_base + 2 * Cfg::NSLAB_G * Cfg::SLAB);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* tfull = empty + Cfg::STAGES;   // [2]
  uint64_t* tempty = tfull + 2;            // [2]
  uint32_t* tslot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint32_t* rowmask = tslot + 4;           // [tile parity][2 groups][4 quadrants]: valid-row bits of a tile (STATS)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = p.tiles_c * p.tiles_w * p.tiles_h * p.tiles_n;
  const int niter = p.ntaps * p.kblocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tslot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = *tslot;
  pdl_wait();   // everything above overlapped the predecessor's tail; global memory is touched only below

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int t = tile;
        const int tc = t % p.tiles_c; t /= p.tiles_c;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h;
        const int tn = t / p.tiles_h;
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn, cout0 = tc * NT;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const ConvTap tp = p.taps[tap];
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1, 0x700 + stage);
            uint8_t* a_dst = smem + stage * Cfg::STAGE;
            mbar_arrive_expect_tx(&full[stage], Cfg::STAGE);
            tma_load_5d(a_dst, &tmA, &full[stage], tp.coff + kb * 64, w0 + tp.dw, h0 + tp.dh, n0, tp.phase);
            tma_load_3d(a_dst + Cfg::A_BYTES, &tmB, &full[stage], kb * 64, cout0, tp.slab);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, NT, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++local) {
        const int buf = local & 1;
        const uint32_t use = (uint32_t)(local >> 1);
        mbar_wait(&tempty[buf], (use & 1) ^ 1, 0x800 + buf);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = taddr + buf * NT;
        for (int it = 0; it < niter; ++it) {
          mbar_wait(&full[stage], phase, 0x900 + stage);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE);
          const uint32_t b_base = a_base + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(tacc, make_smem_desc(a_base + k * 32, 0, 1024), make_smem_desc(b_base + k * 32, 0, 1024), idesc,
                      (it | k) != 0);
          }
          umma_commit(&empty[stage]);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else {
    // ---- two epilogue groups of 4 warps (128 threads) alternate over the 64-channel slabs of a tile;
    //      warp w touches TMEM lanes [32*(w%4), +32)
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // row of the tile = TMEM lane
    const int et = threadIdx.x - 64 - grp * 128; // 0..127 within the group
    const int bar_id = 1 + grp;
    uint8_t* my_slabs = slab_base + grp * Cfg::NSLAB_G * Cfg::SLAB;
    const int wl = r % p.bw, hl = (r / p.bw) % p.bh, nl = r / (p.bw * p.bh);
    const uint32_t sw_row = (uint32_t)(r * 128);
    const uint32_t sw_x = (uint32_t)(r & 7);
    int local = 0;
    int nstore = 0;                              // slabs stored so far by this group
    float ssum[2] = {0.f, 0.f}, ssq[2] = {0.f, 0.f};   // STATS: channel (et & 63) of this group's slabs, row half (et >> 6)
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++local) {
      int t = tile;
      const int tc = t % p.tiles_c; t /= p.tiles_c;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int tn = t / p.tiles_h;
      const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn, cout0 = tc * NT;
      const __nv_bfloat16* arow = nullptr;
      if (EPI) {
        const int n = n0 + nl, h = h0 + hl, w = w0 + wl;
        const bool valid = (n < p.No) && (h < p.Ho) && (w < p.Wo);
        if (p.add_mode && valid) arow = p.add + n * p.a_sn + (h >> p.add_shift) * p.a_sh + (w >> p.add_shift) * p.a_sw;
      }
      if (STATS) {
        // rows outside the image are clipped by the TMA store but a 3x3 tap can make them non-zero: mask them out of the
        // statistics.  One 32-bit mask per TMEM quadrant; visible to the group after the post-staging barrier below
        // (the previous tile's column pass is separated from this write by that tile's own barriers + the tfull wait).
        const bool valid = (n0 + nl < p.No) && (h0 + hl < p.Ho) && (w0 + wl < p.Wo);
        const uint32_t m = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) rowmask[(local & 1) * 8 + grp * 4 + q] = m;   // parity double-buffer: see hazard note in DESIGN.md
      }
      const int buf = local & 1;
      const uint32_t use = (uint32_t)(local >> 1);
      mbar_wait(&tfull[buf], use & 1, 0xA00 + buf);
      tc_fence_after();
      const uint32_t tacc = taddr + buf * NT + (uint32_t(q * 32) << 16);
#pragma unroll 1
      for (int sl = grp; sl < NT / 64; sl += 2) {
        const int ch0 = cout0 + sl * 64;
        if (ch0 >= p.Cout) break;                // uniform across the group
        uint32_t v[64];
        {
          uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
          uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
          tmem_ld32(tacc + sl * 64, lo);
          tmem_ld32(tacc + sl * 64 + 32, hi);
        }
        uint8_t* slab = my_slabs + (nstore % Cfg::NSLAB_G) * Cfg::SLAB;
        // the TMA store that last read this staging buffer must have finished reading it
        if (nstore >= Cfg::NSLAB_G) {
          if (et == 0) tma_store_wait_read<Cfg::NSLAB_G - 1>();
          named_bar_sync(bar_id, 128);
        }
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
          if (EPI) {
            const int ch = ch0 + g * 8;
            if (p.bias && ch < p.Cout) {
              const float4 b0 = *reinterpret_cast<const float4*>(p.bias + ch);
              const float4 b1 = *reinterpret_cast<const float4*>(p.bias + ch + 4);
              // torch's bf16 convolution adds the bias as a separate bf16 op (cudnn_convolution, then output.add_(bias) with
              // the autocast-cast bf16 bias): y = bf16(bf16(acc) + bf16(b)).  Mirrored bit for bit (99.997 % identical
              // outputs on B200, tools/diag_tf.py); a single rounding of acc + b differs on 37 % of the elements.
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = bf16_round(f[j]) + bf16_round(bb[j]);
            }
            if (arow && ch < p.Cout) {
              float a[8];
              unpack8(*reinterpret_cast<const bf16x8*>(arow + ch), a);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = bf16_round(f[j]) + a[j];
            }
          }
          // 128B-swizzled staging: 16-byte chunk g of row r lives at chunk (g ^ (r & 7))
          *reinterpret_cast<bf16x8*>(slab + sw_row + ((uint32_t(g) ^ sw_x) << 4)) = pack8(f);
        }
        fence_proxy_async();                     // generic-proxy smem writes -> visible to the TMA (async proxy)
        named_bar_sync(bar_id, 128);
        if (et == 0) {
          tma_store_5d(&tmC, slab, ch0, w0, h0, n0, 0);
          tma_store_commit();
        }
        if (STATS) {
          // column pass over the staged (bf16-rounded) slab: thread = (channel c, row half); rows outside the image
          // are exact zeros (zero-filled input, no bias) and add nothing
          const int c = et & 63, r0 = (et >> 6) * 64;
          const uint32_t cb = (uint32_t)(c >> 3), co2 = (uint32_t)(c & 7) * 2;
          float s_ = 0.f, q_ = 0.f;
          const uint32_t* rm_ = rowmask + (local & 1) * 8 + grp * 4 + (r0 >> 5);
          const uint32_t m0 = rm_[0], m1 = rm_[1];
#pragma unroll 8
          for (int i = 0; i < 64; ++i) {
            const uint32_t rr = (uint32_t)(r0 + i);
            float v_ = __bfloat162float(
                *reinterpret_cast<const __nv_bfloat16*>(slab + rr * 128 + ((cb ^ (rr & 7)) << 4) + co2));
            v_ = (((i < 32 ? m0 : m1) >> (i & 31)) & 1u) ? v_ : 0.f;
            s_ += v_;
            q_ += v_ * v_;
          }
          ssum[(sl - grp) >> 1] += s_;
          ssq[(sl - grp) >> 1] += q_;
        }
        ++nstore;
      }
      // all TMEM reads of this accumulator by this warp are complete: hand it back to the MMA warp
      tc_fence_before();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
    if (et == 0) tma_store_wait_all();
    if (STATS) {
      // combine the two row halves through this group's (now idle) staging buffer and publish the CTA's partial sums
      named_bar_sync(bar_id, 128);
      float* scr = reinterpret_cast<float*>(my_slabs);   // [j][half][64][2]
      const int c = et & 63, hh = et >> 6;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        scr[((j * 2 + hh) * 64 + c) * 2 + 0] = ssum[j];
        scr[((j * 2 + hh) * 64 + c) * 2 + 1] = ssq[j];
      }
      named_bar_sync(bar_id, 128);
      if (hh == 0) {
        const int cout0 = (blockIdx.x % p.tiles_c) * NT;
        const int slot = blockIdx.x / p.tiles_c;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int sl = grp + 2 * j;
          const int ch = cout0 + sl * 64 + c;
          if (sl < NT / 64 && ch < p.Cout) {
            p.stats[(size_t)ch * kNbPad + slot] = scr[((j * 2 + 0) * 64 + c) * 2 + 0] + scr[((j * 2 + 1) * 64 + c) * 2 + 0];
            p.stats[((size_t)p.Cout + ch) * kNbPad + slot] =
                scr[((j * 2 + 0) * 64 + c) * 2 + 1] + scr[((j * 2 + 1) * 64 + c) * 2 + 1];
          }
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

static int g_num_sms = 0;

static int pow2_le(int x, int cap) {
  int r = 1;
  while (r * 2 <= x && r * 2 <= cap) r *= 2;
  return r;
}

struct ADesc {  // activation tensor [N,H,W,C] bf16 viewed for a conv of the given stride
  const void* ptr;
  int N, H, W, C;
};

static int make_a_map(CUtensorMap* m, const ADesc& a, int stride, int bw, int bh, int bn) {
  uint64_t dims[5], strides[4];
  uint32_t box[5] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn, 1};
  if (stride == 1) {
    dims[0] = a.C; dims[1] = a.W; dims[2] = a.H; dims[3] = a.N; dims[4] = 1;
    strides[0] = (uint64_t)a.C * 2; strides[1] = (uint64_t)a.W * a.C * 2; strides[2] = (uint64_t)a.H * a.W * a.C * 2;
    strides[3] = (uint64_t)a.N * a.H * a.W * a.C * 2;
  } else {  // stride 2: (q*C+c, w', h', n, p) with w = 2w'+q, h = 2h'+p
    if ((a.H & 1) || (a.W & 1)) return EVB_ERR_ARG;
    dims[0] = 2 * (uint64_t)a.C; dims[1] = a.W / 2; dims[2] = a.H / 2; dims[3] = a.N; dims[4] = 2;
    strides[0] = (uint64_t)2 * a.C * 2; strides[1] = (uint64_t)2 * a.W * a.C * 2;
    strides[2] = (uint64_t)a.H * a.W * a.C * 2; strides[3] = (uint64_t)a.W * a.C * 2;
  }
  return evb_make_tmap_bf16(m, a.ptr, 5, dims, strides, box);
}

template <int NT, bool EPI, bool STATS>
static int launch_igemm2_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const IgemmParams& p,
                           cudaStream_t st, int* nblk_out) {
  using Cfg = Igemm2Cfg<NT>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(igemm2_kernel<NT, EPI, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) !=
        cudaSuccess)
      return EVB_ERR_CUDA;
    attr_set = true;
  }
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  const int ntiles = p.tiles_c * p.tiles_w * p.tiles_h * p.tiles_n;
  int grid = ntiles < g_num_sms ? ntiles : g_num_sms;
  if (STATS) {  // a CTA must stay on one output-channel tile: grid is a multiple of tiles_c (ntiles always is)
    if (grid < ntiles) grid = (grid / p.tiles_c) * p.tiles_c;
    if (grid < p.tiles_c) return EVB_ERR_ARG;
    if (nblk_out) *nblk_out = grid / p.tiles_c;
  }
  if (evb_launch_pdl(igemm2_kernel<NT, EPI, STATS>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM, st, tmA, tmB, tmC, p) !=
      cudaSuccess)
    return EVB_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA;
}
template <int NT>
static int launch_igemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const IgemmParams& p,
                         cudaStream_t st, int* nblk_out) {
  if (p.stats) {
    if (p.add_mode) return EVB_ERR_ARG;
    // with a bias the statistics are those of bf16(conv + bias): rows outside the image are masked out of the sums
    if (p.bias) return launch_igemm2_t<NT, true, true>(tmA, tmB, tmC, p, st, nblk_out);
    return launch_igemm2_t<NT, false, true>(tmA, tmB, tmC, p, st, nblk_out);
  }
  if (p.bias || p.add_mode) return launch_igemm2_t<NT, true, false>(tmA, tmB, tmC, p, st, nullptr);
  return launch_igemm2_t<NT, false, false>(tmA, tmB, tmC, p, st, nullptr);
}

// Core host entry: `a` is the tensor the A boxes are cut from, (To_n, To_h, To_w) the pixel grid the tiles
// enumerate (output grid for fwd / stride-1 dgrad, per-phase grid for stride-2 dgrad).
static int run_igemm(const ADesc& a, int a_stride, const void* wpk, int w_rows, int w_cin, int w_slabs,
                     IgemmParams p, int force_nt, cudaStream_t st, int* nblk_out = nullptr) {
  if (a.C % 64 || w_cin % 64 || p.Cout % 8) return EVB_ERR_ARG;
  p.bw = pow2_le(p.Wo, 128);
  p.bh = pow2_le(p.Ho, 128 / p.bw);
  p.bn = 128 / (p.bw * p.bh);
  p.tiles_w = (p.Wo + p.bw - 1) / p.bw;
  p.tiles_h = (p.Ho + p.bh - 1) / p.bh;
  p.tiles_n = (p.No + p.bn - 1) / p.bn;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  int nt = 64;
  if (force_nt) {
    nt = force_nt;
  } else {
    const int cands[3] = {256, 128, 64};
    bool found = false;
    for (int i = 0; i < 3 && !found; ++i) {
      const int c = cands[i];
      if (c > 64 && c > p.Cout) continue;
      if ((long long)tiles_m * ((p.Cout + c - 1) / c) >= 118) { nt = c; found = true; }
    }
    if (!found) nt = 64;
  }
  p.tiles_c = (p.Cout + nt - 1) / nt;
  CUtensorMap tmA, tmB;
  int rc = make_a_map(&tmA, a, a_stride, p.bw, p.bh, p.bn);
  if (rc) return rc;
  uint64_t wd[3] = {(uint64_t)w_cin, (uint64_t)w_rows, (uint64_t)w_slabs};
  uint64_t ws[2] = {(uint64_t)w_cin * 2, (uint64_t)w_cin * w_rows * 2};
  uint32_t wb[3] = {64, (uint32_t)nt, 1};
  rc = evb_make_tmap_bf16(&tmB, wpk, 3, wd, ws, wb);
  if (rc) return rc;
  {
    // output tensor map for the TMA-store epilogue: (channel, w, h, n, 1) with the caller's strides (covers the strided
    // per-phase outputs of the stride-2 dgrad); channels >= Cout and pixels outside the image are clipped by the TMA.
    CUtensorMap tmC;
    uint64_t od[5] = {(uint64_t)p.Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.No, 1};
    uint64_t os[4] = {(uint64_t)p.o_sw * 2, (uint64_t)p.o_sh * 2, (uint64_t)p.o_sn * 2, (uint64_t)p.o_sn * p.No * 2};
    uint32_t ob[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn, 1};
    rc = evb_make_tmap_bf16(&tmC, p.out, 5, od, os, ob);
    if (rc) return rc;
    switch (nt) {
      case 256: return launch_igemm2<256>(tmA, tmB, tmC, p, st, nblk_out);
      case 128: return launch_igemm2<128>(tmA, tmB, tmC, p, st, nblk_out);
      case 64: return launch_igemm2<64>(tmA, tmB, tmC, p, st, nblk_out);
    }
    return EVB_ERR_ARG;
  }
  return EVB_ERR_ARG;
}

}  // namespace evb

using namespace evb;

// y[N,Ho,Wo,Cout] = conv(x[N,H,W,Cin], w) (+bias) (+add).  ksize in {1,3}, pad = ksize/2, stride in {1,2}.
// wpk: bf16 [ksize*ksize][w_rows][Cin], w_rows >= Cout.  add_mode: 0 none, 1 same-shape bf16 tensor,
// 2 half-resolution tensor [N,Ho/2,Wo/2,Cout] sampled nearest (FPN top-down, ever/module/fpn.py:96-105).
static int conv2d_fwd_impl(const void* x, int N, int H, int W, int Cin, const void* wpk, int w_rows, int ksize, int stride,
                           void* y, int Cout, const float* bias, const void* add, int add_mode, int force_nt, float* stats,
                           int* nblk_out, void* stream) {
  if ((ksize != 1 && ksize != 3) || (stride != 1 && stride != 2)) return EVB_ERR_ARG;
  IgemmParams p{};
  p.stats = stats;
  const int Ho = H / stride, Wo = W / stride;
  p.kblocks = Cin / 64;
  p.ntaps = ksize * ksize;
  for (int r = 0; r < ksize; ++r)
    for (int s = 0; s < ksize; ++s) {
      ConvTap& t = p.taps[r * ksize + s];
      const int oh = r - ksize / 2, ow = s - ksize / 2;  // input offset relative to stride*h
      t.slab = r * ksize + s;
      if (stride == 1) {
        t.dh = oh; t.dw = ow; t.coff = 0; t.phase = 0;
      } else {  // 2h+oh = 2(h+dh)+ph
        const int ph = oh & 1, pw = ow & 1;
        t.dh = (oh - ph) / 2; t.dw = (ow - pw) / 2; t.phase = ph; t.coff = pw * Cin;
      }
    }
  p.out = (__nv_bfloat16*)y;
  p.o_sw = Cout; p.o_sh = (long long)Wo * Cout; p.o_sn = (long long)Ho * Wo * Cout;
  p.No = N; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
  p.bias = bias;
  p.add_mode = add_mode ? 1 : 0;
  if (add_mode) {
    p.add = (const __nv_bfloat16*)add;
    p.add_shift = add_mode == 2 ? 1 : 0;
    const int Ha = Ho >> p.add_shift, Wa = Wo >> p.add_shift;
    p.a_sw = Cout; p.a_sh = (long long)Wa * Cout; p.a_sn = (long long)Ha * Wa * Cout;
  }
  ADesc a{x, N, H, W, Cin};
  return run_igemm(a, stride, wpk, w_rows, Cin, ksize * ksize, p, force_nt, (cudaStream_t)stream, nblk_out);
}

extern "C" int evb_conv2d_fwd(const void* x, int N, int H, int W, int Cin, const void* wpk, int w_rows, int ksize,
                              int stride, void* y, int Cout, const float* bias, const void* add, int add_mode,
                              int force_nt, void* stream) {
  return conv2d_fwd_impl(x, N, H, W, Cin, wpk, w_rows, ksize, stride, y, Cout, bias, add, add_mode, force_nt, nullptr,
                         nullptr, stream);
}

// Convolution (no bias / add) whose epilogue also emits the BatchNorm batch-statistic partial sums of its bf16 output:
// partial = fp32 [2][Cout][320] (sum | sum of squares per channel, one column per CTA), *nblk_out columns are valid.
// Feed them to evb_bn_finalize: replaces the separate evb_bn_stats read pass over the conv output.
extern "C" int evb_conv2d_fwd_stats(const void* x, int N, int H, int W, int Cin, const void* wpk, int w_rows, int ksize,
                                    int stride, void* y, int Cout, float* partial, int* nblk_out, void* stream) {
  if (!partial || !nblk_out) return EVB_ERR_ARG;
  return conv2d_fwd_impl(x, N, H, W, Cin, wpk, w_rows, ksize, stride, y, Cout, nullptr, nullptr, 0, 0, partial, nblk_out,
                         stream);
}
// same with a per-channel fp32 bias added before the bf16 rounding (conv + bias -> BN of the FS-Relation encoders,
// ever/module/fs_relation.py:41-52): the statistics are those of the biased, rounded output
extern "C" int evb_conv2d_fwd_bias_stats(const void* x, int N, int H, int W, int Cin, const void* wpk, int w_rows, int ksize,
                                         int stride, void* y, int Cout, const float* bias, float* partial, int* nblk_out,
                                         void* stream) {
  if (!partial || !nblk_out || !bias) return EVB_ERR_ARG;
  return conv2d_fwd_impl(x, N, H, W, Cin, wpk, w_rows, ksize, stride, y, Cout, bias, nullptr, 0, 0, partial, nblk_out,
                         stream);
}

// dx[N,H,W,Cin] (+)= conv_transpose(dy[N,Ho,Wo,Cout], w).  wpk_t: bf16 [ksize*ksize][w_rows>=Cin][Cout]
// (the transposed pack, same tap order r*ksize+s as the forward pack).  accumulate: add into the bf16 dx.
extern "C" int evb_conv2d_dgrad(const void* dy, int N, int Ho, int Wo, int Cout, const void* wpk_t, int w_rows,
                                int ksize, int stride, void* dx, int H, int W, int Cin, int accumulate, int force_nt,
                                void* stream) {
  if ((ksize != 1 && ksize != 3) || (stride != 1 && stride != 2)) return EVB_ERR_ARG;
  if (H != Ho * stride || W != Wo * stride) return EVB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  ADesc a{dy, N, Ho, Wo, Cout};
  IgemmParams base{};
  base.kblocks = Cout / 64;
  base.out = (__nv_bfloat16*)dx;
  base.No = N; base.Cout = Cin;
  base.bias = nullptr;
  if (stride == 1) {
    IgemmParams p = base;
    p.ntaps = ksize * ksize;
    for (int r = 0; r < ksize; ++r)
      for (int s = 0; s < ksize; ++s) {
        ConvTap& t = p.taps[r * ksize + s];
        t.dh = -(r - ksize / 2); t.dw = -(s - ksize / 2); t.coff = 0; t.phase = 0; t.slab = r * ksize + s;
      }
    p.o_sw = Cin; p.o_sh = (long long)W * Cin; p.o_sn = (long long)H * W * Cin;
    p.Ho = H; p.Wo = W;
    if (accumulate) { p.add_mode = 1; p.add = (const __nv_bfloat16*)dx; p.a_sn = p.o_sn; p.a_sh = p.o_sh; p.a_sw = p.o_sw; }
    return run_igemm(a, 1, wpk_t, w_rows, Cout, ksize * ksize, p, force_nt, st);
  }
  // stride 2: output phase (ph,pw): dx[2i+ph, 2j+pw] = sum over taps with (ph - (r-pad)) even of dy[i + dh, j + dw] w[r,s]
  const int pad = ksize / 2;
  if (ksize == 1 && !accumulate &&
      cudaMemsetAsync(dx, 0, (size_t)N * H * W * Cin * 2, st) != cudaSuccess)
    return EVB_ERR_CUDA;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      IgemmParams p = base;
      p.ntaps = 0;
      for (int r = 0; r < ksize; ++r)
        for (int s = 0; s < ksize; ++s) {
          const int nh = ph - (r - pad), nw = pw - (s - pad);  // 2*h_out = 2i + nh
          if ((nh & 1) || (nw & 1)) continue;
          ConvTap& t = p.taps[p.ntaps++];
          t.dh = nh / 2; t.dw = nw / 2; t.coff = 0; t.phase = 0; t.slab = r * ksize + s;
        }
      p.out = (__nv_bfloat16*)dx + ((long long)ph * W + pw) * Cin;
      p.o_sw = 2LL * Cin; p.o_sh = 2LL * W * Cin; p.o_sn = (long long)H * W * Cin;
      p.Ho = H / 2; p.Wo = W / 2;
      if (accumulate) { p.add_mode = 1; p.add = p.out; p.a_sn = p.o_sn; p.a_sh = p.o_sh; p.a_sw = p.o_sw; }
      if (p.ntaps == 0) continue;  // phase receives no gradient (1x1 stride 2): dx was zeroed above
      int rc = run_igemm(a, 1, wpk_t, w_rows, Cout, ksize * ksize, p, force_nt, st);
      if (rc) return rc;
    }
  return EVB_OK;
}

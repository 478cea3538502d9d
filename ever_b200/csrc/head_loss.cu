// FS-Relation head kernels, scene-embedding MLP (tiny GEMV-class linears), and the fused
// softmax cross-entropy + Dice loss (forward statistics and logit gradient).
//
// Reference: FSRelation.forward ever/module/fs_relation.py:57-73; scene encoder :22-28; FarSegHead GAP :177;
// F.cross_entropy(ignore_index=255) in the user model; dice_loss_with_logits ever/module/loss.py:54-75
// (select :26-37, dice_coeff :40-51, all_reduce_sum :20-23).  Arithmetic restated in SURVEY.md Appendix D.
#include "common.cuh"

namespace evb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ FS-Relation
// u1, u2: [M, C] bf16 = 1x1 conv outputs (+bias) of the content encoder / feature re-encoder, pre-BN.
// cf = relu(bn1(u1)), pf = relu(bn2(u2)); r = sigmoid(sum_c bf16(sf_c * cf_c)); z = bf16(r * pf).
// One warp per pixel, C = 256*NG: lane owns channel groups [lane*8 + 256*a, +8); the per-channel BN folds are hoisted
// into registers and PX pixels are in flight per warp.
// PR: the product sf_c * cf_c is a bf16 tensor in the reference (FSRelation: bf16 scene vector, :57-73); FSRelationV2's scene
// vector leaves its GroupNorm in fp32 (autocast runs group_norm in fp32), so there the product stays fp32 (PR = false).
// ldz: elements per pixel of the output rows (FSRelationV2 writes z into the first half of its concat buffer).
template <int NG, int PX, bool PR>
__global__ void __launch_bounds__(256)
relation_fwd_kernel(const __nv_bfloat16* __restrict__ u1, const __nv_bfloat16* __restrict__ u2,
                    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ scale2,
                    const float* __restrict__ shift2, const float* __restrict__ sf, __nv_bfloat16* __restrict__ z,
                    float* __restrict__ rel, long long M, int HW, int ldz) {
  constexpr int C = 256 * NG;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float s1[NG][8], b1[NG][8], s2[NG][8], b2[NG][8], sfv[NG][8];
#pragma unroll
  for (int a = 0; a < NG; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane * 8 + 256 * a + j;
      s1[a][j] = scale[c]; b1[a][j] = shift[c]; s2[a][j] = scale2[c]; b2[a][j] = shift2[c];
    }
  int cur_n = -1;
  for (long long m0 = warp0 * PX; m0 < M; m0 += nwarps * PX) {
    bf16x8 q1[PX][NG], q2[PX][NG];
#pragma unroll
    for (int p = 0; p < PX; ++p)
      if (m0 + p < M) {
#pragma unroll
        for (int a = 0; a < NG; ++a) {
          q1[p][a] = *reinterpret_cast<const bf16x8*>(u1 + (m0 + p) * C + lane * 8 + 256 * a);
          q2[p][a] = *reinterpret_cast<const bf16x8*>(u2 + (m0 + p) * C + lane * 8 + 256 * a);
        }
      }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      const long long m = m0 + p;
      if (m >= M) break;
      const int n = (int)(m / HW);
      if (n != cur_n) {
        cur_n = n;
#pragma unroll
        for (int a = 0; a < NG; ++a)
#pragma unroll
          for (int j = 0; j < 8; ++j) sfv[a][j] = sf[n * C + lane * 8 + 256 * a + j];
      }
      float dot = 0.f;
#pragma unroll
      for (int a = 0; a < NG; ++a) {
        float v[8];
        unpack8(q1[p][a], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float pr_ = sfv[a][j] * fmaxf(bf16_round(v[j] * s1[a][j] + b1[a][j]), 0.f);
          dot += PR ? bf16_round(pr_) : pr_;
        }
      }
      dot = warp_sum(dot);
      const float r = 1.f / (1.f + __expf(-dot));
      if (lane == 0) rel[m] = r;
#pragma unroll
      for (int a = 0; a < NG; ++a) {
        float v[8];
        unpack8(q2[p][a], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = r * fmaxf(bf16_round(v[j] * s2[a][j] + b2[a][j]), 0.f);
        *reinterpret_cast<bf16x8*>(z + m * ldz + lane * 8 + 256 * a) = pack8(v);
      }
    }
  }
}

// g1/g2 [M,C] = gradient w.r.t. the two BN outputs (ReLU masks applied); dsf_part[warp][c] = sum over the warp's pixels
// of dlogit * cf (px_per_warp divides HW, so a warp never straddles two images; reduced deterministically afterwards).
template <int NG, int PX, bool PR>
__global__ void __launch_bounds__(256, NG == 1 ? 2 : 1)
relation_bwd_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ u1,
                    const __nv_bfloat16* __restrict__ u2, const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ scale2, const float* __restrict__ shift2, const float* __restrict__ sf,
                    const float* __restrict__ rel, __nv_bfloat16* __restrict__ g1, __nv_bfloat16* __restrict__ g2,
                    float* __restrict__ dsf_part, long long M, int HW, int px_per_warp, int lddz) {
  constexpr int C = 256 * NG;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long mbeg = warp * px_per_warp;
  if (mbeg >= M) return;
  const long long mend = mbeg + px_per_warp < M ? mbeg + px_per_warp : M;
  float s1[NG][8], b1[NG][8], s2[NG][8], b2[NG][8], sfv[NG][8], acc[NG][8];
#pragma unroll
  for (int a = 0; a < NG; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane * 8 + 256 * a + j;
      s1[a][j] = scale[c]; b1[a][j] = shift[c]; s2[a][j] = scale2[c]; b2[a][j] = shift2[c];
      acc[a][j] = 0.f;
    }
  int cur_n = -1;
  for (long long m0 = mbeg; m0 < mend; m0 += PX) {
    bf16x8 q1[PX][NG], q2[PX][NG], qd[PX][NG];
    float rr[PX];
#pragma unroll
    for (int p = 0; p < PX; ++p)
      if (m0 + p < mend) {
        rr[p] = rel[m0 + p];
#pragma unroll
        for (int a = 0; a < NG; ++a) {
          const long long off = (m0 + p) * C + lane * 8 + 256 * a;
          q1[p][a] = *reinterpret_cast<const bf16x8*>(u1 + off);
          q2[p][a] = *reinterpret_cast<const bf16x8*>(u2 + off);
          qd[p][a] = *reinterpret_cast<const bf16x8*>(dz + (m0 + p) * lddz + lane * 8 + 256 * a);
        }
      }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      const long long m = m0 + p;
      if (m >= mend) break;
      const int n = (int)(m / HW);
      if (n != cur_n) {
        cur_n = n;
#pragma unroll
        for (int a = 0; a < NG; ++a)
#pragma unroll
          for (int j = 0; j < 8; ++j) sfv[a][j] = sf[n * C + lane * 8 + 256 * a + j];
      }
      const float r = rr[p];
      float dr = 0.f;
#pragma unroll
      for (int a = 0; a < NG; ++a) {
        float v[8], d[8], o[8];
        unpack8(q2[p][a], v);
        unpack8(qd[p][a], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float pf = fmaxf(bf16_round(v[j] * s2[a][j] + b2[a][j]), 0.f);
          dr += d[j] * pf;
          o[j] = pf > 0.f ? d[j] * r : 0.f;
        }
        *reinterpret_cast<bf16x8*>(g2 + m * C + lane * 8 + 256 * a) = pack8(o);
      }
      dr = warp_sum(dr);
      // autocast flow of the reference: sum(dim=1) runs in fp32 on the bf16 product, so its backward hands a bf16-ROUNDED
      // d(logit) to the bf16 multiply; d(cf) = bf16(dl * sf) and d(sf) sums bf16(dl * cf) (SURVEY.md 8a dtype flow)
      // (PR = false, FSRelationV2: the product and the scene vector are fp32, nothing is rounded before d(cf) is stored)
      const float dl_ = dr * r * (1.f - r);
      const float dlogit = PR ? bf16_round(dl_) : dl_;
#pragma unroll
      for (int a = 0; a < NG; ++a) {
        float v[8], o[8];
        unpack8(q1[p][a], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float cf = fmaxf(bf16_round(v[j] * s1[a][j] + b1[a][j]), 0.f);
          acc[a][j] += PR ? bf16_round(dlogit * cf) : dlogit * cf;
          o[j] = cf > 0.f ? dlogit * sfv[a][j] : 0.f;
        }
        *reinterpret_cast<bf16x8*>(g1 + m * C + lane * 8 + 256 * a) = pack8(o);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < NG; ++a) {
    float* dst = dsf_part + warp * C + lane * 8 + 256 * a;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
  }
}

// dsf[n][c] = sum_{j < wpi} part[n * wpi + j][c]  (fixed order).  block (32 c, 8 row lanes), grid (C/32, N)
__global__ void relation_dsf_reduce_kernel(const float* __restrict__ part, float* __restrict__ dsf, int wpi, int C) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x, n = blockIdx.y;
  float s = 0.f;
  for (int j = threadIdx.y; j < wpi; j += 8) s += part[((long long)n * wpi + j) * C + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    dsf[n * C + c] = t;
  }
}

// ------------------------------------------------------------------------------------------------ tiny linears
// y[n][o] = round_bf16( sum_i bf16(W[o][i]) * x[n][i] + b[o] ), optional ReLU.  One warp per (n, o).
__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                                  float* __restrict__ y, int N, int I, int O, int relu) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= N * O) return;
  const int n = wid / O, o = wid % O;
  float s = 0.f;
  for (int i = lane; i < I; i += 32) s += bf16_round(W[(long long)o * I + i]) * x[(long long)n * I + i];
  s = warp_sum(s);
  if (lane == 0) {
    s = bf16_round(s + (b ? b[o] : 0.f));
    y[wid] = relu ? fmaxf(s, 0.f) : s;
  }
}
// dW[o][i] (+)= sum_n g[n][o] x[n][i];  db[o] (+)= sum_n g[n][o];  g = dy * (y>0 if relu)
__global__ void linear_bwd_w_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                                    float* __restrict__ dW, float* __restrict__ db, int N, int I, int O, int relu,
                                    int accumulate) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)O * I) return;
  const int o = (int)(idx / I), i = (int)(idx % I);
  float s = 0.f, sb = 0.f;
  for (int n = 0; n < N; ++n) {
    float g = dy[n * O + o];
    if (relu && !(y[n * O + o] > 0.f)) g = 0.f;
    g = bf16_round(g);
    s += g * x[(long long)n * I + i];
    sb += g;
  }
  dW[idx] = accumulate ? dW[idx] + s : s;
  if (i == 0 && db) db[o] = accumulate ? db[o] + sb : sb;
}
// dx[n][i] (+)= sum_o g[n][o] bf16(W[o][i]).  block (128 i, 8 o-lanes), grid (I/128, N)
__global__ void __launch_bounds__(1024)
linear_bwd_x_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ W,
                    float* __restrict__ dx, int N, int I, int O, int relu, int accumulate) {
  __shared__ float red[8][129];
  const int i = blockIdx.x * 128 + threadIdx.x, n = blockIdx.y;
  float s = 0.f;
  if (i < I) {
    for (int o = threadIdx.y; o < O; o += 8) {
      float g = dy[n * O + o];
      if (relu && !(y[n * O + o] > 0.f)) g = 0.f;
      s += bf16_round(g) * bf16_round(W[(long long)o * I + i]);
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < I) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    t = bf16_round(t);
    const long long idx = (long long)n * I + i;
    dx[idx] = accumulate ? dx[idx] + t : t;
  }
}

// ------------------------------------------------------------------------------------------------ CE + Dice
// logits: [P, LD] bf16 (NHWC, first K channels are classes), labels: int64 [P], 255 (ignore_index) = ignored.
// stats layout (fp32): [0]=sum of -log p_t over valid, [1]=n_valid, [2..2+K)=I_c, [2+K..2+2K)=sum p_c, [2+2K..2+3K)=sum y_c
constexpr int kMaxK = 64;   // classes per loss group (logit rows of LD = 16 / 32 / 64 bf16)

template <int PASS, int KT, int KMAX = 16>
__global__ void __launch_bounds__(256)
loss_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels, long long P, int Krt, int LD,
            int ignore_index, float* __restrict__ partial, const float* __restrict__ coef,
            __nv_bfloat16* __restrict__ dlogits) {
  // PASS 0: statistics -> partial[block][2+3K];  PASS 1: dlogits from coef = {inv_nvalid, A_c[K], B_c[K]}
  // KT > 0: K is a compile-time constant (arrays stay in registers); KT == 0: generic K <= KMAX (16 / 32 / 64)
  const int K = KT > 0 ? KT : Krt;
  constexpr int KA = KT > 0 ? KT : KMAX;
  // BINS (statistics pass, K <= 16): the two per-class sums only the pixel's OWN class contributes to (I_c = sum p_t, and
  // the class histogram) go to a private shared-memory column per thread, indexed by the label -- one read-modify-write
  // each instead of K compare + select + add chains (the pass was instruction-bound: ~700 SASS instructions per pixel);
  // bank = thread id, so no conflicts; fixed-order reduction below, so the sums stay deterministic.
  constexpr bool BINS = PASS == 0 && KA <= 16;
  __shared__ float red[8][2 + 3 * KA];
  __shared__ float bins[BINS ? 2 * KA : 1][BINS ? 256 : 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float acc[2 + 3 * KA];
  if (PASS == 0) {
#pragma unroll
    for (int i = 0; i < 2 + 3 * KA; ++i) acc[i] = 0.f;
    if (BINS) {
#pragma unroll
      for (int i = 0; i < 2 * KA; ++i) bins[i][threadIdx.x] = 0.f;
    }
  }
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const long long t = labels[p];
    const bool valid = t != ignore_index;
    float z[KA];
    const __nv_bfloat16* row = logits + p * LD;
#pragma unroll
    for (int c0 = 0; c0 < K; c0 += 8) {
      float v[8];
      unpack8(*reinterpret_cast<const bf16x8*>(row + c0), v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < K) z[c0 + j] = v[j];
    }
    float mx = z[0];
#pragma unroll
    for (int c = 1; c < K; ++c) mx = fmaxf(mx, z[c]);
    // one exponential per class: e_c = exp(z_c - max) (arguments <= 0: the ex2.approx intrinsic is within ~1e-6 relative),
    // softmax p_c = e_c / sum e; the kernel is instruction-bound, a second exp(z_c - lse) per class cost 40 % of it
    float e[KA];
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < K; ++c) { e[c] = __expf(z[c] - mx); se += e[c]; }
    const float lse = mx + logf(se);
    const float inv_se = 1.f / se;
    if (PASS == 0 && BINS) {
      if (valid) {
        const int ti = (int)t;
        const bool in_range = ti >= 0 && ti < K;   // a label outside [0, K) matches no class (the reference asserts)
        const float zt = in_range ? __bfloat162float(row[in_range ? ti : 0]) : 0.f;   // same bf16 value as z[t] (L1 hit)
        acc[0] += lse - zt;
        acc[1] += 1.f;
#pragma unroll
        for (int c = 0; c < K; ++c) acc[2 + K + c] += e[c] * inv_se;
        if (in_range) {
          bins[ti][threadIdx.x] += __expf(zt - mx) * inv_se;   // == e[t] * inv_se bit for bit
          bins[KA + ti][threadIdx.x] += 1.f;
        }
      }
    } else if (PASS == 0) {
      if (valid) {
        float zt = 0.f;
#pragma unroll
        for (int c = 0; c < K; ++c) zt = (c == (int)t) ? z[c] : zt;
        acc[0] += lse - zt;
        acc[1] += 1.f;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          // select arithmetic instead of `if (c == t) acc[..] += ..`: the compiler turned that chain into acc[2 + t] with a
          // dynamic index, which moved acc[] and e[] to local memory (248-byte stack frame, kernel 2x slower)
          const float pc = e[c] * inv_se;
          const float hit = (c == (int)t) ? 1.f : 0.f;
          acc[2 + K + c] += pc;
          acc[2 + c] += hit * pc;
          acc[2 + 2 * K + c] += hit;
        }
      }
    } else {
      float d[KA];
      if (valid) {
        const float inv_n = coef[0];
        float pc[KA], dot = 0.f;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          pc[c] = e[c] * inv_se;
          const float gc = coef[1 + K + c] - (c == (int)t ? coef[1 + c] : 0.f);  // B_c - y_c * A_c
          dot += pc[c] * gc;
        }
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const float gc = coef[1 + K + c] - (c == (int)t ? coef[1 + c] : 0.f);
          const float dce = (pc[c] - (c == (int)t ? 1.f : 0.f)) * inv_n;
          const float ddice = pc[c] * (gc - dot);
          // the reference sums two bf16 gradient tensors (one per loss term)
          d[c] = bf16_round(dce) + bf16_round(ddice);
        }
      } else {
#pragma unroll
        for (int c = 0; c < K; ++c) d[c] = 0.f;
      }
#pragma unroll
      for (int c0 = 0; c0 < (KA + 7) / 8 * 8 || c0 < 16; c0 += 8) {
        if (c0 >= LD) break;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (c0 + j < K) ? d[c0 + j] : 0.f;
        *reinterpret_cast<bf16x8*>(dlogits + p * LD + c0) = pack8(v);
      }
      for (int c0 = ((KA + 7) / 8 * 8 > 16 ? (KA + 7) / 8 * 8 : 16); c0 < LD; c0 += 8)   // padding channels up to LD
        *reinterpret_cast<uint4*>(dlogits + p * LD + c0) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (PASS == 0) {
    const int n = 2 + 3 * K;
    if (BINS) {
      __syncthreads();
      // class c (I_c) and KA + c (histogram): 256 thread columns summed in a fixed order by one warp each
      for (int b = wid; b < 2 * KA; b += 8) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) v += bins[b][lane + 32 * j];
        v = warp_sum(v);
        const int c = b < KA ? b : b - KA;
        if (lane == 0 && c < K) partial[(long long)blockIdx.x * n + (b < KA ? 2 + c : 2 + 2 * K + c)] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < 2 + 3 * KA; ++i) {   // no early break: keeps the indices static so acc[] stays in registers
      if (i < n) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) red[wid][i] = v;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (BINS && i >= 2 && (i < 2 + K || i >= 2 + 2 * K)) continue;   // written from the bins above
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += red[w][i];
      partial[(long long)blockIdx.x * n + i] = s;
    }
  }
}

// Binary head (K == 1): masked BCE-with-logits (ever/module/loss.py:229-235, _masked_ignore :10-17) + sigmoid Dice
// (dice_loss_with_logits K==1 branch, loss.py:66-68).  logits: channel 0 of [P, LD]; labels in {0, 1, ignore}.
// stats = {sum softplus(z) - z*y, n_valid, I = sum p*y, sum p, sum y}; coef = {bce_w / n_valid, A, B}.
template <int PASS>
__global__ void __launch_bounds__(256)
loss_binary_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels, long long P, int LD,
                   int ignore_index, float* __restrict__ partial, const float* __restrict__ coef,
                   __nv_bfloat16* __restrict__ dlogits) {
  __shared__ float red[8][5];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const long long t = labels[p];
    const bool valid = t != ignore_index;
    const float z = __bfloat162float(logits[p * LD]);
    const float y = (float)t;
    const float pr = 1.f / (1.f + expf(-z));
    if (PASS == 0) {
      if (valid) {
        acc[0] += fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));
        acc[1] += 1.f;
        acc[2] += pr * y;
        acc[3] += pr;
        acc[4] += y;
      }
    } else {
      float d = 0.f;
      if (valid) {
        const float g = coef[2] - y * coef[1];  // B - y * A
        d = bf16_round((pr - y) * coef[0]) + bf16_round(pr * (1.f - pr) * g);
      }
      float v[8] = {d, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      *reinterpret_cast<bf16x8*>(dlogits + p * LD) = pack8(v);
      const float zz[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int c0 = 8; c0 < LD; c0 += 8) *reinterpret_cast<bf16x8*>(dlogits + p * LD + c0) = pack8(zz);
    }
  }
  if (PASS == 0) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float v = warp_sum(acc[i]);
      if (lane == 0) red[wid][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
      float s_ = 0.f;
      for (int w = 0; w < 8; ++w) s_ += red[w][threadIdx.x];
      partial[(long long)blockIdx.x * 5 + threadIdx.x] = s_;
    }
  }
}

// stats[i] = sum_blocks partial (double accumulation, fixed order)
// one warp per statistic: lanes stride over the block partials, fp64, fixed shuffle tree -> deterministic
__global__ void loss_reduce_kernel(const float* __restrict__ partial, int nblk, int n, float* __restrict__ stats) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int b = lane; b < nblk; b += 32) s += partial[(long long)b * n + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) stats[i] = (float)s;
}

// From (possibly all-reduced) statistics: losses[0]=ce, losses[1]=dice; coef = {ce_scale/n_valid, A_c, B_c}.
// dice_stats may point to globally summed Dice statistics (I, sum p, sum y: 3K floats); dice_grad_scale
// multiplies the Dice gradient (world size, see DESIGN.md: DDP averages what autograd's all-reduce summed).
__global__ void loss_finalize_kernel(const float* __restrict__ stats, const float* __restrict__ dice_stats, int K,
                                     float smooth, float ce_weight, float dice_weight, float dice_grad_scale,
                                     float* __restrict__ losses, float* __restrict__ coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float nvalid = stats[1];
  // no valid pixel: 0 / 0 = NaN like F.cross_entropy's mean over an empty selection (its gradient is zero, not NaN)
  losses[0] = stats[0] / nvalid;
  coef[0] = nvalid > 0.f ? ce_weight / nvalid : 0.f;
  float coeff_sum = 0.f;
  for (int c = 0; c < K; ++c) {
    const float I = dice_stats[c], Z = dice_stats[K + c] + dice_stats[2 * K + c] + smooth;
    coeff_sum += (2.f * I + smooth) / Z;
    coef[1 + c] = dice_weight * dice_grad_scale * 2.f / ((float)K * Z);
    coef[1 + K + c] = dice_weight * dice_grad_scale * (2.f * I + smooth) / ((float)K * Z * Z);
  }
  losses[1] = 1.f - coeff_sum / (float)K;
}

// ------------------------------------------------------------------------------------------------ fused SGD
// L2 norm partials of a flat fp32 gradient arena
__global__ void __launch_bounds__(256) sqsum_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    s += v * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
// 256 threads: strided fp64 partial sums, fixed shared-memory tree -> deterministic
__global__ void norm_finalize_kernel(const float* partial, int nblk, float* out /*[0]=norm,[1]=clip coef*/, float max_norm) {
  __shared__ double red[256];
  double s = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 256) s += partial[b];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x) return;
  const float norm = (float)sqrt(red[0]);
  out[0] = norm;
  float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
  out[1] = coef < 1.f ? coef : 1.f;
}
// torch.optim.SGD (momentum, weight decay, dampening 0, no nesterov) + clip coefficient, in place
__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ mom, long long n, const float* __restrict__ lr_ptr,
           float momentum, float wd, const float* __restrict__ clip, int first_step, int zero_grad,
           const unsigned char* __restrict__ trainable) {
  const float c = clip ? clip[1] : 1.f;
  const float lr = *lr_ptr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // torch.optim.SGD skips parameters without a gradient (requires_grad=False): no decay, no momentum, no update
    if (trainable && !trainable[i]) continue;
    float d = g[i] * c + wd * w[i];
    float b = first_step ? d : momentum * mom[i] + d;
    mom[i] = b;
    w[i] -= lr * b;
    if (zero_grad) g[i] = 0.f;
  }
}

}  // namespace evb

using namespace evb;
#define ST ((cudaStream_t)stream)
#define LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA)

template <bool PR>
static int relation_fwd_launch(const void* u1, const void* u2, const float* scale1, const float* shift1, const float* scale2,
                               const float* shift2, const float* sf, void* z, float* rel, long long M, int HW, int C, int ldz,
                               void* stream) {
  if ((C != 256 && C != 512 && C != 1024) || ldz < C || ldz % 8) return EVB_ERR_ARG;
  long long blocks = (M + 31) / 32;  // 8 warps x 4 pixels
  if (blocks > 148 * 8) blocks = 148 * 8;
  const __nv_bfloat16 *a = (const __nv_bfloat16*)u1, *b = (const __nv_bfloat16*)u2;
  __nv_bfloat16* zz = (__nv_bfloat16*)z;
  if (C == 256) relation_fwd_kernel<1, 4, PR><<<(int)blocks, 256, 0, ST>>>(a, b, scale1, shift1, scale2, shift2, sf, zz, rel, M, HW, ldz);
  else if (C == 512) relation_fwd_kernel<2, 2, PR><<<(int)blocks, 256, 0, ST>>>(a, b, scale1, shift1, scale2, shift2, sf, zz, rel, M, HW, ldz);
  else relation_fwd_kernel<4, 1, PR><<<(int)blocks, 256, 0, ST>>>(a, b, scale1, shift1, scale2, shift2, sf, zz, rel, M, HW, ldz);
  return LAUNCH_OK();
}
extern "C" int evb_relation_fwd(const void* u1, const void* u2, const float* scale1, const float* shift1,
                                const float* scale2, const float* shift2, const float* sf, void* z, float* rel, long long M,
                                int HW, int C, void* stream) {
  return relation_fwd_launch<true>(u1, u2, scale1, shift1, scale2, shift2, sf, z, rel, M, HW, C, C, stream);
}
// FSRelationV2 flavour (ever/module/fs_relation.py:142-163): fp32 scene vector / fp32 product, z rows of ldz elements
extern "C" int evb_relation_fwd_v2(const void* u1, const void* u2, const float* scale1, const float* shift1,
                                   const float* scale2, const float* shift2, const float* sf, void* z, int ldz, float* rel,
                                   long long M, int HW, int C, void* stream) {
  return relation_fwd_launch<false>(u1, u2, scale1, shift1, scale2, shift2, sf, z, rel, M, HW, C, ldz, stream);
}
static int relation_px_per_warp(int HW) {
  int p = 16;
  while (p > 1 && HW % p) p >>= 1;
  return p;
}
extern "C" long long evb_relation_bwd_workspace(long long M, int HW, int C) {
  return (M / relation_px_per_warp(HW)) * (long long)C * sizeof(float);
}
// dsf[N][C] is overwritten: per-warp partial sums go to `ws` and are reduced in a fixed order (deterministic).
template <bool PR>
static int relation_bwd_launch(const void* dz, int lddz, const void* u1, const void* u2, const float* scale1,
                               const float* shift1, const float* scale2, const float* shift2, const float* sf,
                               const float* rel, void* g1, void* g2, float* dsf, long long M, int HW, int C, void* ws,
                               void* stream) {
  if ((C != 256 && C != 512 && C != 1024) || M % HW || lddz < C || lddz % 8) return EVB_ERR_ARG;
  const int px_per_warp = relation_px_per_warp(HW);
  const long long warps = M / px_per_warp;
  const long long blocks = (warps + 7) / 8;
  const __nv_bfloat16 *d = (const __nv_bfloat16*)dz, *a = (const __nv_bfloat16*)u1, *b = (const __nv_bfloat16*)u2;
  __nv_bfloat16 *o1 = (__nv_bfloat16*)g1, *o2 = (__nv_bfloat16*)g2;
  float* part = (float*)ws;
  if (C == 256)
    relation_bwd_kernel<1, 2, PR><<<(int)blocks, 256, 0, ST>>>(d, a, b, scale1, shift1, scale2, shift2, sf, rel, o1, o2, part, M, HW, px_per_warp, lddz);
  else if (C == 512)
    relation_bwd_kernel<2, 1, PR><<<(int)blocks, 256, 0, ST>>>(d, a, b, scale1, shift1, scale2, shift2, sf, rel, o1, o2, part, M, HW, px_per_warp, lddz);
  else
    relation_bwd_kernel<4, 1, PR><<<(int)blocks, 256, 0, ST>>>(d, a, b, scale1, shift1, scale2, shift2, sf, rel, o1, o2, part, M, HW, px_per_warp, lddz);
  relation_dsf_reduce_kernel<<<dim3(C / 32, (unsigned)(M / HW)), dim3(32, 8), 0, ST>>>(part, dsf, HW / px_per_warp, C);
  return LAUNCH_OK();
}
extern "C" int evb_relation_bwd(const void* dz, const void* u1, const void* u2, const float* scale1, const float* shift1,
                                const float* scale2, const float* shift2, const float* sf, const float* rel, void* g1,
                                void* g2, float* dsf, long long M, int HW, int C, void* ws, void* stream) {
  return relation_bwd_launch<true>(dz, C, u1, u2, scale1, shift1, scale2, shift2, sf, rel, g1, g2, dsf, M, HW, C, ws, stream);
}
extern "C" int evb_relation_bwd_v2(const void* dz, int lddz, const void* u1, const void* u2, const float* scale1,
                                   const float* shift1, const float* scale2, const float* shift2, const float* sf,
                                   const float* rel, void* g1, void* g2, float* dsf, long long M, int HW, int C, void* ws,
                                   void* stream) {
  return relation_bwd_launch<false>(dz, lddz, u1, u2, scale1, shift1, scale2, shift2, sf, rel, g1, g2, dsf, M, HW, C, ws,
                                    stream);
}

extern "C" int evb_linear_fwd(const float* x, const float* W, const float* b, float* y, int N, int I, int O, int relu,
                              void* stream) {
  const int warps = N * O;
  linear_fwd_kernel<<<(warps + 7) / 8, 256, 0, ST>>>(x, W, b, y, N, I, O, relu);
  return LAUNCH_OK();
}
extern "C" int evb_linear_bwd(const float* dy, const float* y, const float* x, const float* W, float* dW, float* db,
                              float* dx, int N, int I, int O, int relu, int acc_w, int acc_x, void* stream) {
  linear_bwd_w_kernel<<<(int)(((long long)O * I + 255) / 256), 256, 0, ST>>>(dy, y, x, dW, db, N, I, O, relu, acc_w);
  if (dx) linear_bwd_x_kernel<<<dim3((I + 127) / 128, N), dim3(128, 8), 0, ST>>>(dy, y, W, dx, N, I, O, relu, acc_x);
  return LAUNCH_OK();
}

static int loss_blocks(long long P) {
  long long b = (P + 256 * 4 - 1) / (256 * 4);
  if (b > 148 * 4) b = 148 * 4;
  return b < 1 ? 1 : (int)b;
}
extern "C" long long evb_loss_workspace(long long P, int K) { return (long long)loss_blocks(P) * (2 + 3 * K) * sizeof(float); }

// Pass A: statistics of softmax-CE + Dice over valid pixels -> stats[2+3K] (device, fp32).
extern "C" int evb_loss_stats(const void* logits, const void* labels, long long P, int K, int LD, int ignore_index,
                              float* stats, void* ws, void* stream) {
  if (K < 1 || K > kMaxK || (LD != 16 && LD != 32 && LD != 64) || K > LD) return EVB_ERR_ARG;
  const int nb = loss_blocks(P), n = 2 + 3 * K;
  if (K == 1) {
    loss_binary_kernel<0><<<nb, 256, 0, ST>>>((const __nv_bfloat16*)logits, (const long long*)labels, P, LD, ignore_index,
                                              (float*)ws, nullptr, nullptr);
    loss_reduce_kernel<<<(n * 32 + 127) / 128, 128, 0, ST>>>((const float*)ws, nb, n, stats);
    return LAUNCH_OK();
  }
#define EVB_LOSS_LAUNCH(PASS_, GRID_, ...)                                                        \
  switch (K) {                                                                                   \
    case 15: loss_kernel<PASS_, 15><<<GRID_, 256, 0, ST>>>(__VA_ARGS__); break;                  \
    case 7: loss_kernel<PASS_, 7><<<GRID_, 256, 0, ST>>>(__VA_ARGS__); break;                    \
    case 5: loss_kernel<PASS_, 5><<<GRID_, 256, 0, ST>>>(__VA_ARGS__); break;                    \
    case 2: loss_kernel<PASS_, 2><<<GRID_, 256, 0, ST>>>(__VA_ARGS__); break;                    \
    default:                                                                                     \
      if (K <= 16) loss_kernel<PASS_, 0, 16><<<GRID_, 256, 0, ST>>>(__VA_ARGS__);                \
      else if (K <= 32) loss_kernel<PASS_, 0, 32><<<GRID_, 256, 0, ST>>>(__VA_ARGS__);           \
      else loss_kernel<PASS_, 0, 64><<<GRID_, 256, 0, ST>>>(__VA_ARGS__);                        \
  }
  EVB_LOSS_LAUNCH(0, nb, (const __nv_bfloat16*)logits, (const long long*)labels, P, K, LD, ignore_index, (float*)ws,
                  nullptr, nullptr)
  loss_reduce_kernel<<<(n * 32 + 127) / 128, 128, 0, ST>>>((const float*)ws, nb, n, stats);
  return LAUNCH_OK();
}
// Between A and B the caller may all-reduce stats[2 : 2+3K] (Dice statistics) across ranks into dice_stats.
extern "C" int evb_loss_finalize(const float* stats, const float* dice_stats, int K, float smooth, float ce_weight,
                                 float dice_weight, float dice_grad_scale, float* losses, float* coef, void* stream) {
  loss_finalize_kernel<<<1, 32, 0, ST>>>(stats, dice_stats ? dice_stats : stats + 2, K, smooth, ce_weight, dice_weight,
                                         dice_grad_scale, losses, coef);
  return LAUNCH_OK();
}
// Pass B: dlogits[P, LD] bf16 (padding channels zeroed).
extern "C" int evb_loss_grad(const void* logits, const void* labels, long long P, int K, int LD, int ignore_index,
                             const float* coef, void* dlogits, void* stream) {
  if (K < 1 || K > kMaxK || (LD != 16 && LD != 32 && LD != 64) || K > LD) return EVB_ERR_ARG;
  if (K == 1) {
    loss_binary_kernel<1><<<loss_blocks(P), 256, 0, ST>>>((const __nv_bfloat16*)logits, (const long long*)labels, P, LD,
                                                          ignore_index, nullptr, coef, (__nv_bfloat16*)dlogits);
    return LAUNCH_OK();
  }
  EVB_LOSS_LAUNCH(1, loss_blocks(P), (const __nv_bfloat16*)logits, (const long long*)labels, P, K, LD, ignore_index,
                  nullptr, coef, (__nv_bfloat16*)dlogits)
  return LAUNCH_OK();
}

extern "C" long long evb_sgd_workspace(long long n) {
  long long b = (n + 256 * 8 - 1) / (256 * 8);
  if (b > 148 * 8) b = 148 * 8;
  return (b + 2) * sizeof(float);
}
// Global L2 norm of the gradient arena -> norm_out[0], clip coefficient -> norm_out[1] (ERModule.clip_grad,
// ever/interface/module.py:96-108: clip_grad_norm_(max_norm, norm_type=2)).
extern "C" int evb_grad_norm(const float* g, long long n, float max_norm, float* norm_out, void* ws, void* stream) {
  long long b = (n + 256 * 8 - 1) / (256 * 8);
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  sqsum_kernel<<<(int)b, 256, 0, ST>>>(g, n, (float*)ws);
  norm_finalize_kernel<<<1, 256, 0, ST>>>((const float*)ws, (int)b, norm_out, max_norm);
  return LAUNCH_OK();
}
// torch.optim.SGD step over a flat arena (ever/opt/optimizer.py:7-9), lr read from device memory.
extern "C" int evb_sgd_step(float* w, float* g, float* mom, long long n, const float* lr, float momentum, float wd,
                            const float* clip, int first_step, int zero_grad, void* stream) {
  long long b = (n + 256 * 4 - 1) / (256 * 4);
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  sgd_kernel<<<(int)b, 256, 0, ST>>>(w, g, mom, n, lr, momentum, wd, clip, first_step, zero_grad, nullptr);
  return LAUNCH_OK();
}
// Same with a per-slot trainable mask (uint8, 1 = update): frozen parameters (freeze_at, batchnorm_trainable=False) keep
// their weights and momentum untouched, as torch.optim.SGD does for parameters whose grad is None.
extern "C" int evb_sgd_step_masked(float* w, float* g, float* mom, long long n, const float* lr, float momentum, float wd,
                                   const float* clip, int first_step, int zero_grad, const unsigned char* trainable,
                                   void* stream) {
  long long b = (n + 256 * 4 - 1) / (256 * 4);
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  sgd_kernel<<<(int)b, 256, 0, ST>>>(w, g, mom, n, lr, momentum, wd, clip, first_step, zero_grad, trainable);
  return LAUNCH_OK();
}

// prob[N,K,H,W] fp32 = softmax over the K class channels of logits[P, LD] bf16 (eval path: logit.softmax(dim=1)),
// mask[P] uint8 = argmax (lowest index wins ties, as torch.argmax on CUDA).
namespace evb {
template <int KMAX>
__global__ void softmax_nchw_kernel(const __nv_bfloat16* __restrict__ logits, float* __restrict__ prob,
                                    uint8_t* __restrict__ mask, long long P, int HW, int K, int LD) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const __nv_bfloat16* row = logits + p * LD;
    if (K == 1) {  // binary head: sigmoid probability, mask = p > 0.5
      const float pr = 1.f / (1.f + expf(-__bfloat162float(row[0])));
      if (prob) prob[p] = pr;
      if (mask) mask[p] = pr > 0.5f ? 1 : 0;
      continue;
    }
    float z[KMAX];
    for (int c = 0; c < K; ++c) z[c] = __bfloat162float(row[c]);
    float mx = z[0];
    int am = 0;
    for (int c = 1; c < K; ++c)
      if (z[c] > mx) { mx = z[c]; am = c; }
    float se = 0.f;
    for (int c = 0; c < K; ++c) { z[c] = expf(z[c] - mx); se += z[c]; }
    const long long n = p / HW, hw = p % HW;
    if (prob)
      for (int c = 0; c < K; ++c) prob[(n * K + c) * HW + hw] = z[c] / se;
    if (mask) mask[p] = (uint8_t)am;
  }
}
}  // namespace evb
extern "C" int evb_softmax_nchw(const void* logits, float* prob, void* mask, long long P, int HW, int K, int LD,
                                void* stream) {
  if (K < 1 || K > kMaxK) return EVB_ERR_ARG;
  long long b = (P + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  const __nv_bfloat16* lg = (const __nv_bfloat16*)logits;
  if (K <= 16) softmax_nchw_kernel<16><<<(int)b, 256, 0, ST>>>(lg, prob, (uint8_t*)mask, P, HW, K, LD);
  else if (K <= 32) softmax_nchw_kernel<32><<<(int)b, 256, 0, ST>>>(lg, prob, (uint8_t*)mask, P, HW, K, LD);
  else softmax_nchw_kernel<64><<<(int)b, 256, 0, ST>>>(lg, prob, (uint8_t*)mask, P, HW, K, LD);
  return LAUNCH_OK();
}

// Index-map kernels either side of the hot path: dihedral (flip / rot90 / transpose)
// + crop + constant-pad gathers over pixel-interleaved images, and the sliding-window / test-time-augmentation
// probability canvas.  Pure byte / fp32 movement: HBM-bound, bit-exact against the torch ops the reference calls.
//
// Reference ops replaced: torch.rot90 / torch.flip / slicing / F.pad in ever/preprocess/thsegm.py:7-147
// (THRandomRotate90k, THRandomHorizontalFlip, THRandomVerticalFlip, THRandomCrop), th_divisible_pad / th_pad_to_size
// (ever/preprocess/function.py:35-83), the TTA transforms of ever/magic/transform/segm.py:9-72 with
// `sum(outs) / len(outs)` (ever/magic/transform/tta.py:11-23), and the window accumulation a caller of
// ever/magic/bigimage/sliding_window.py:8-33 performs on the host.
#include "common.cuh"

namespace evb {

// one row of the per-sample map table (int32 x 12):
//   dst(i, j) = src[s][a00*i + a01*j + b0][a10*i + a11*j + b1]  if that source pixel lies in [y0,y1) x [x0,x1), else fill
struct PixMap {
  int a00, a01, b0, a10, a11, b1, s, y0, y1, x0, x1, cb;   // cb: canvas index (canvas kernels only)
};
static_assert(sizeof(PixMap) == 48, "map row is 12 int32");

template <typename V>
__global__ void __launch_bounds__(256)
pixel_gather_kernel(const V* __restrict__ src, int Hs, int Ws, int vec_per_pix, const PixMap* __restrict__ maps,
                    V* __restrict__ dst, int Ho, int Wo, long long total, V fill) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t % vec_per_pix);
    long long p = t / vec_per_pix;
    const int j = (int)(p % Wo); p /= Wo;
    const int i = (int)(p % Ho);
    const int n = (int)(p / Ho);
    const PixMap m = maps[n];
    const int si = m.a00 * i + m.a01 * j + m.b0;
    const int sj = m.a10 * i + m.a11 * j + m.b1;
    V val = fill;
    if (si >= m.y0 && si < m.y1 && sj >= m.x0 && sj < m.x1)
      val = src[(((long long)m.s * Hs + si) * Ws + sj) * vec_per_pix + v];
    dst[t] = val;
  }
}

// canvas[cb][k][y][x] += prob[s][k][i][j] for every map row (in table order -> fixed summation order, no atomics) of
// canvas cb whose footprint covers (y, x); (i, j) = A (y, x) + b is the canvas->tile map.  count[cb][y][x] += coverage.
__global__ void __launch_bounds__(256)
canvas_accumulate_kernel(const float* __restrict__ prob, int N, int K, int h, int w, const PixMap* __restrict__ maps,
                         float* __restrict__ canvas, float* __restrict__ count, int B, int Hc, int Wc, int ylo, int yhi,
                         int xlo, int xhi) {
  const int bw = xhi - xlo, bh = yhi - ylo;
  const long long total = (long long)B * bh * bw;
  const long long plane = (long long)Hc * Wc;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = xlo + (int)(t % bw);
    const int y = ylo + (int)((t / bw) % bh);
    const int cb = (int)(t / ((long long)bw * bh));
    float cnt = 0.f;
    float* dst = canvas + (long long)cb * K * plane + (long long)y * Wc + x;
    for (int n = 0; n < N; ++n) {
      const PixMap m = maps[n];
      if (m.cb != cb || y < m.y0 || y >= m.y1 || x < m.x0 || x >= m.x1) continue;
      const int i = m.a00 * y + m.a01 * x + m.b0;
      const int j = m.a10 * y + m.a11 * x + m.b1;
      if (i < 0 || i >= h || j < 0 || j >= w) continue;
      const float* src = prob + ((long long)m.s * K * h + i) * w + j;
      for (int k = 0; k < K; ++k) dst[k * plane] += src[(long long)k * h * w];
      cnt += 1.f;
    }
    if (count && cnt > 0.f) count[(long long)cb * plane + (long long)y * Wc + x] += cnt;
  }
}

// prob[b][k][p] = canvas[b][k][p] / count[b][p] (in place when prob == canvas); mask[b][p] = argmax_k (lowest index wins
// ties, like torch.argmax) or (prob > 0.5) for K == 1.  Pixels never covered (count == 0) get 0 and mask 0.
__global__ void __launch_bounds__(256)
canvas_finalize_kernel(const float* canvas, const float* __restrict__ count, float uniform_count, int K, long long P,
                       long long BP, float* prob, uint8_t* __restrict__ mask) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < BP; t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / P, p = t % P;
    const float c = count ? count[t] : uniform_count;
    const float* cv = canvas + b * K * P + p;
    float* pr = prob ? prob + b * K * P + p : nullptr;
    float best = -1.f;
    int arg = 0;
    for (int k = 0; k < K; ++k) {
      const float v = c > 0.f ? cv[(long long)k * P] / c : 0.f;
      if (pr) pr[(long long)k * P] = v;
      if (v > best) { best = v; arg = k; }
    }
    if (mask) mask[t] = K == 1 ? (uint8_t)(best > 0.5f) : (uint8_t)arg;
  }
}

static inline int sp_blocks(long long work) {
  long long b = (work + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return b < 1 ? 1 : (int)b;
}

// dst[p][y][x] (+)= bilinear(src[p]) with align_corners = true, fp32 planes (the TTA `Scale` transform and its inverse,
// ever/magic/transform/segm.py:71-88 = F.interpolate(mode='bilinear', align_corners=True)).  Same fp32 operation order as
// torch's upsample_bilinear2d: source coordinate = (in - 1) / (out - 1) * index, weights (1 - l, l), rows combined last.
__global__ void __launch_bounds__(256)
resize_bilinear_ac_kernel(const float* __restrict__ src, int h, int w, float* __restrict__ dst, int ho, int wo,
                          long long total, float rh, float rw, int accumulate) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % wo);
    long long q = t / wo;
    const int y = (int)(q % ho);
    const long long p = q / ho;
    const float fy = rh * y, fx = rw * x;
    const int y0 = (int)fy, x0 = (int)fx;
    const int yp = y0 < h - 1 ? 1 : 0, xp = x0 < w - 1 ? 1 : 0;
    const float ly1 = fy - y0, ly0 = 1.f - ly1, lx1 = fx - x0, lx0 = 1.f - lx1;
    const float* s0 = src + (p * h + y0) * w + x0;
    const float* s1 = s0 + (long long)yp * w;
    const float top = __fadd_rn(__fmul_rn(lx0, s0[0]), __fmul_rn(lx1, s0[xp]));
    const float bot = __fadd_rn(__fmul_rn(lx0, s1[0]), __fmul_rn(lx1, s1[xp]));
    const float v = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
    dst[t] = accumulate ? dst[t] + v : v;
  }
}

}  // namespace evb

using namespace evb;

// dst[N,Ho,Wo,pix_bytes] <- gather of src[Ns,Hs,Ws,pix_bytes] through the per-sample int32[12] map rows
// (device memory).  elem_bytes in {1,2,4,8} is the element size used for `fill` (e.g. 8 for int64 labels padded
// with 255, ever/preprocess/thcomm.py:67-88); pix_bytes must be a multiple of elem_bytes.
extern "C" int evb_pixel_gather(const void* src, int Ns, int Hs, int Ws, int pix_bytes, int elem_bytes, long long fill,
                                const void* maps, void* dst, int N, int Ho, int Wo, void* stream) {
  (void)Ns;
  if (pix_bytes <= 0 || (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8) ||
      pix_bytes % elem_bytes)
    return EVB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const PixMap* mp = (const PixMap*)maps;
  const int vpp = pix_bytes / elem_bytes;
  const long long total = (long long)N * Ho * Wo * vpp;
  if (total == 0) return EVB_OK;
  const int grid = sp_blocks(total);
  switch (elem_bytes) {
    case 1: pixel_gather_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)src, Hs, Ws, vpp, mp, (uint8_t*)dst, Ho, Wo, total, (uint8_t)fill); break;
    case 2: pixel_gather_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)src, Hs, Ws, vpp, mp, (uint16_t*)dst, Ho, Wo, total, (uint16_t)fill); break;
    case 4: pixel_gather_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)src, Hs, Ws, vpp, mp, (uint32_t*)dst, Ho, Wo, total, (uint32_t)fill); break;
    default: pixel_gather_kernel<unsigned long long><<<grid, 256, 0, st>>>((const unsigned long long*)src, Hs, Ws, vpp, mp, (unsigned long long*)dst, Ho, Wo, total, (unsigned long long)fill); break;
  }
  return cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA;
}

// canvas[B,K,Hc,Wc] fp32 += the probability tiles prob[*,K,h,w] fp32 placed by the N canvas->tile map rows;
// count[B,Hc,Wc] (may be NULL) += coverage.  Only canvas pixels inside [ylo,yhi) x [xlo,xhi) (the batch's bounding box)
// are visited.
extern "C" int evb_canvas_accumulate(const float* prob, int N, int K, int h, int w, const void* maps, float* canvas,
                                     float* count, int B, int Hc, int Wc, int ylo, int yhi, int xlo, int xhi,
                                     void* stream) {
  if (ylo < 0 || xlo < 0 || yhi > Hc || xhi > Wc || N < 0 || K < 1 || B < 1) return EVB_ERR_ARG;
  if (yhi <= ylo || xhi <= xlo || N == 0) return EVB_OK;
  const long long total = (long long)B * (yhi - ylo) * (xhi - xlo);
  canvas_accumulate_kernel<<<sp_blocks(total), 256, 0, (cudaStream_t)stream>>>(prob, N, K, h, w, (const PixMap*)maps, canvas,
                                                                            count, B, Hc, Wc, ylo, yhi, xlo, xhi);
  return cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA;
}

// prob[B,K,P] = canvas / count[B,P] (count NULL: divide by uniform_count, the TTA case `sum(outs) / len(outs)`),
// mask[B,P] = argmax
extern "C" int evb_canvas_finalize(const float* canvas, const float* count, float uniform_count, int B, int K, long long P,
                                   float* prob, void* mask, void* stream) {
  if (K < 1 || P < 0 || B < 1) return EVB_ERR_ARG;
  if (P == 0) return EVB_OK;
  canvas_finalize_kernel<<<sp_blocks(B * P), 256, 0, (cudaStream_t)stream>>>(canvas, count, uniform_count, K, P, B * P, prob,
                                                                           (uint8_t*)mask);
  return cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA;
}

// dst[P,ho,wo] (+)= bilinear resize (align_corners = true) of src[P,h,w], fp32 planes
extern "C" int evb_resize_bilinear_ac(const float* src, long long P, int h, int w, float* dst, int ho, int wo, int accumulate,
                                      void* stream) {
  if (P < 0 || h < 1 || w < 1 || ho < 1 || wo < 1) return EVB_ERR_ARG;
  const long long total = P * ho * wo;
  if (total == 0) return EVB_OK;
  const float rh = ho > 1 ? (float)(h - 1) / (float)(ho - 1) : 0.f, rw = wo > 1 ? (float)(w - 1) / (float)(wo - 1) : 0.f;
  resize_bilinear_ac_kernel<<<sp_blocks(total), 256, 0, (cudaStream_t)stream>>>(src, h, w, dst, ho, wo, total, rh, rw, accumulate);
  return cudaGetLastError() == cudaSuccess ? EVB_OK : EVB_ERR_CUDA;
}

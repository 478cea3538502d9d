/* libevb200.so -- C ABI of the B200-native FarSeg/ChangeStar hot path.
 *
 * The reference (Z-Zheng/ever) is pure Python over PyTorch: it has no FFI of its own.  The boundary this
 * library replaces is therefore the ATen op each reference nn.Module call dispatches to; every entry point
 * below names the reference call site (file:line under the reference tree) it stands in for.  INTEGRATION.md
 * shows the ctypes stub a maintainer binds these with from ever's own Python.
 *
 * Conventions
 *   - return value: 0 = EVB_OK, 1 = bad argument, 2 = CUDA launch/runtime error, 3 = driver entry point missing.
 *     Nothing throws across the boundary; evb_last_cuda_error() gives the CUDA error string.
 *   - all buffers are caller-allocated DEVICE memory (raw pointers); no hidden allocation: kernels that need
 *     scratch take a workspace pointer sized by the matching *_workspace() query.
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it (no synchronisation).
 *   - activations: NHWC bf16 (channels contiguous), C % 8 == 0 (C % 64 == 0 for convolution operands);
 *     parameters / statistics / gradients of parameters: fp32; labels: int64, ignore_index = 255.
 *   - one host thread per GPU (one process per GPU); the library keeps no global state except cached
 *     kernel attributes.
 */
#ifndef EVB200_H_
#define EVB200_H_
#ifdef __cplusplus
extern "C" {
#endif

int evb_version(void);
const char* evb_last_cuda_error(void);
int evb_device_sync_check(void);
/* programmatic dependent launch of the tensor-core kernels (their prologue overlaps the predecessor's tail); default on,
 * EVB_PDL=0 in the environment or evb_set_pdl(0) turns it off */
int evb_set_pdl(int on);
/* same for the BatchNorm finalize / apply kernels behind them; default on, EVB_PDL_SMALL=0 or evb_set_pdl_small(0) = off */
int evb_set_pdl_small(int on);

/* ---- convolution (tcgen05 implicit GEMM).  Replaces nn.Conv2d forward/backward:
 * ever/module/_resnets.py:21-29,139-150 (ResNet 3x3/1x1/stem), ever/module/ops.py:53-55 (ConvBlock),
 * ever/module/fpn.py:165,179 (decoder, classifier), ever/module/fs_relation.py:25-27,42,49. */
/* y[N,H/s,W/s,Cout] = conv(x[N,H,W,Cin]) (+bias[Cout] fp32) (+add).  ksize 1|3, pad ksize/2, stride 1|2.
 * wpk: bf16 [ksize*ksize][w_rows>=Cout][Cin].  add_mode 0 none | 1 same-shape bf16 | 2 half-resolution bf16
 * sampled nearest (FPN top-down add, ever/module/fpn.py:96-105).  force_nt: 0 = auto tile, else 64|128|256. */
int evb_conv2d_fwd(const void* x, int N, int H, int W, int Cin, const void* wpk, int w_rows, int ksize, int stride,
                   void* y, int Cout, const float* bias, const void* add, int add_mode, int force_nt, void* stream);
/* Convolution (no bias / add) whose epilogue also emits BatchNorm batch-statistic partial sums of its bf16 output:
 * partial = fp32 [2][Cout][320] (sum | sum of squares per channel, one column per CTA), *nblk_out columns valid; pass them
 * to evb_bn_finalize (replaces the evb_bn_stats read pass; conv + BN of ever/module/_resnets.py:92-112, fpn.py:165-166). */
int evb_conv2d_fwd_stats(const void* x, int N, int H, int W, int Cin, const void* wpk, int w_rows, int ksize, int stride,
                         void* y, int Cout, float* partial, int* nblk_out, void* stream);
/* same, with a per-channel fp32 bias added before the bf16 rounding (conv+bias -> BN, ever/module/fs_relation.py:41-52) */
int evb_conv2d_fwd_bias_stats(const void* x, int N, int H, int W, int Cin, const void* wpk, int w_rows, int ksize, int stride,
                              void* y, int Cout, const float* bias, float* partial, int* nblk_out, void* stream);
/* dx[N,H,W,Cin] (+)= conv_transpose(dy[N,Ho,Wo,Cout]).  wpk_t: bf16 [ksize*ksize][w_rows>=Cin][Cout]. */
int evb_conv2d_dgrad(const void* dy, int N, int Ho, int Wo, int Cout, const void* wpk_t, int w_rows, int ksize,
                     int stride, void* dx, int H, int W, int Cin, int accumulate, int force_nt, void* stream);
/* dw[Cout][Cin][k][k] fp32 (+)= x (*) dy, split-K over pixels with deterministic reduction. */
long long evb_conv2d_wgrad_workspace(int N, int Ho, int Wo, int Cin, int Cout, int ksize, int force_nt, int force_split);
int evb_conv2d_wgrad(const void* x, int N, int H, int W, int Cin, const void* dy, int Cout, int ksize, int stride,
                     float* dw, int accumulate, void* ws, long long ws_bytes, int force_nt, int force_split, void* stream);
/* fp32 OIHW master weights -> bf16 packs [kk][CoP][CiP] (forward) and [kk][CiPb][CoPb] (dgrad), zero padded. */
int evb_pack_weight(const float* w, int Co, int Ci, int kk, void* wf, int CoP, int CiP, void* wb, int CiPb, int CoPb,
                    void* stream);
/* every convolution of a model in one launch: desc int64[n][12], block_map int32[nblocks] (one block per 64 x 64 (co, ci)
 * tile, all taps; see elementwise.cu); bf16 staging, 16-byte stores of both layouts; blocks [block0, block0 + nblocks) */
int evb_pack_weights_tiled(const void* desc, const void* block_map, int block0, int nblocks, void* stream);
/* 7x7 stride-2 pad-3 stem lowered to a GEMM: x NCHW fp32 -> A[N*H/2*W/2][KP] bf16, k = c*49 + r*7 + s
 * (ResNet.stem_forward, ever/module/_resnets.py:205-212). */
int evb_stem_im2col(const float* x, void* a, int N, int Cin, int H, int W, int KP, void* stream);
/* same from a raw uint8 HWC tile [N,H,W,Cin] with (u8 - mean[c]) / std[c] fused (th_mean_std_normalize,
 * ever/preprocess/function.py:9-32): the caller-side input pipeline, 4x fewer H2D bytes */
int evb_stem_im2col_u8(const void* x, const float* mean, const float* stdv, void* a, int N, int Cin, int H, int W, int KP,
                       void* stream);
/* general ks x ks window (stride 1|2, pad): the first 3x3 stride-2 conv of the deep "v1c" stem (ever/module/_resnets.py:137-147)
 * uses (3, 2, 1); evb_stem_im2col[_u8] = (7, 2, 3) */
int evb_im2col_nchw(const float* x, void* a, int N, int Cin, int H, int W, int KP, int ks, int stride, int pad, void* stream);
int evb_im2col_u8(const void* x, const float* mean, const float* stdv, void* a, int N, int Cin, int H, int W, int KP, int ks,
                  int stride, int pad, void* stream);
/* eval: cm[t*K + p] (int64) += pixels with label t predicted p; labels outside [0,K) skipped
 * (ConfusionMatrix.forward, ever/metric/confusion_matrix.py:11-25) */
int evb_confusion_matrix(const void* pred, const void* labels, long long P, int K, void* cm, void* stream);

/* ---- BatchNorm2d (+ReLU, + residual add).  Replaces nn.BatchNorm2d / nn.ReLU / `out += identity`:
 * ever/module/_resnets.py:46-49,58,66-67,83-87,97,101,109-110; fpn.py:166-167; fs_relation.py:43-44,50-51. */
long long evb_bn_workspace(long long M, int C);
/* training statistics of x[M,C] -> mean, rstd, folded scale/shift; running stats updated (momentum, unbiased var) */
int evb_bn_stats(const void* x, long long M, int C, const float* gamma, const float* beta, float* running_mean,
                 float* running_var, float momentum, float eps, float* mean, float* rstd, float* scale, float* shift,
                 void* ws, void* stream);
/* same outputs as evb_bn_stats, from the partial sums written by evb_conv2d_fwd_stats */
int evb_bn_finalize(const float* partial, int nblk, long long M, int C, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, float* mean, float* rstd,
                    float* scale, float* shift, void* stream);
/* nn.SyncBatchNorm (train.sync_bn, ever/trainer/th_ddp_trainer.py:21-22): batch statistics over ALL ranks.  partial_sums: this
 * rank's sums[2][C] (sum | sum of squares) from the conv epilogue's partial columns -- the caller all-reduces them --;
 * finalize_sums: mean / rstd / scale / shift + running statistics from global sums over M_total rows; bwd_consts: the constants
 * (c2, k0) of evb_norm_bwd_apply from the all-reduced backward sums (dgamma, dbeta of evb_norm_bwd_reduce), inv_m = 1/M_total */
int evb_bn_partial_sums(const float* partial, int nblk, int C, float* sums, void* stream);
int evb_bn_finalize_sums(const float* sums, double M_total, int C, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float momentum, float eps, float* mean, float* rstd, float* scale, float* shift,
                         void* stream);
int evb_bn_bwd_consts(const float* scale, const float* rstd, const float* mean, const float* dgamma, const float* dbeta,
                      float inv_m, int C, float* c2, float* k0, void* stream);
/* eval / frozen BN: fold running statistics (ever/module/resnet.py:155-160,227-234) */
int evb_bn_fold(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, int C, float* scale,
                float* shift, float* mean, float* rstd, void* stream);
/* y = act(bf16(x*scale+shift) [+ res]) */
int evb_bn_apply(const void* x, const float* scale, const float* shift, const void* res, void* y, long long M, int C,
                 int relu, void* stream);
/* y = relu(bf16(x*scale+shift) + res), and mask32[v] (one uint32 per 8-channel vector v of y) = bit j set when channel j of
 * the rounded output is > 0: the ReLU survivors for evb_bn_bwd(mask_mode = 3), a quarter of the bytes of re-reading y
 * (Bottleneck / BasicBlock tail: out = relu(bn(conv(x)) + identity), ever/module/_resnets.py:64-69,105-112) */
int evb_bn_apply_mask(const void* x, const float* scale, const float* shift, const void* res, void* y, void* mask32,
                      long long M, int C, void* stream);
/* backward of the above.  mask_mode 0 none | 1 (ymask>0) | 2 recomputed from x | 3 ymask = the uint32 bit masks of
 * evb_bn_apply_mask.  dres (+)= masked dy. */
/* grid cap of the BN backward reduction, blocks per SM (1..4, default 4) */
int evb_set_bn_reduce_blocks(int per_sm);
int evb_bn_bwd(const void* dy, const void* x, const void* ymask, const float* mean, const float* rstd, const float* scale,
               const float* shift, int mask_mode, int frozen, void* dx, void* dres, int dres_acc, float* dgamma,
               float* dbeta, int param_acc, long long M, int C, void* ws, void* stream);
/* the two halves of evb_bn_bwd on their own (GroupNorm, squeeze-excitation and plain-ReLU backward are built from them):
 * reduce: dbeta[c] (+)= sum_rows g, dgamma[c] (+)= rstd[c] * sum_rows g * (x - mean[c]), g = dy * mask;
 * apply:  dx = a[c] * g + k0[c] - c2[c] * x with explicit per-channel constants, dres (+)= g */
int evb_norm_bwd_reduce(const void* dy, const void* x, const void* ymask, const float* mean, const float* rstd,
                        const float* scale, const float* shift, int mask_mode, float* dgamma, float* dbeta, int param_acc,
                        long long M, int C, void* ws, void* stream);
int evb_norm_bwd_apply(const void* dy, const void* x, const void* ymask, const float* a, const float* shift, const float* c2,
                       const float* k0, int mask_mode, void* dx, void* dres, int dres_acc, long long M, int C, void* stream);
/* nn.GroupNorm(G, Creal) on an NHWC tensor of ONE sample (FreeNet's conv3x3_gn_relu blocks): evb_bn_stats gives the per-channel
 * mean / rstd (with eps_bn); evb_gn_fold combines the channels of each group into GroupNorm's statistics and emits the
 * per-channel affine map (scale, shift) for evb_bn_apply plus (gmean, grstd) for the backward; channels >= Creal (zero padding)
 * get the identity.  evb_gn_bwd_consts turns the per-channel sums of evb_norm_bwd_reduce (called with mean = gmean,
 * rstd = grstd) into the constants of evb_norm_bwd_apply; inv_m = 1 / (pixels * Creal / G). */
int evb_gn_fold(const float* mean_c, const float* rstd_c, float eps_bn, const float* gamma, const float* beta, int C, int Creal,
                int G, float eps, float* scale, float* shift, float* gmean, float* grstd, void* stream);
int evb_gn_bwd_consts(const float* dgamma_c, const float* dbeta_c, const float* gamma, const float* gmean, const float* grstd,
                      int C, int Creal, int G, float inv_m, float* c2, float* k0, void* stream);
/* squeeze-excitation gate (ever/module/se_block.py:9-24): sig = bf16(sigmoid(s)); ds = bf16(dsig * sig * (1 - sig)) */
int evb_sigmoid_fwd(const float* s, float* sig, int n, void* stream);
int evb_sigmoid_bwd(const float* dsig, const float* sig, float* ds, int n, void* stream);
/* db[C] (+)= column sums of dy[M,C]  (conv bias gradient) */
int evb_bias_grad(const void* dy, long long M, int C, float* db, float* unused, int accumulate, void* ws, void* stream);

/* ---- pooling / resampling.  nn.MaxPool2d(3,2,1) ever/module/_resnets.py:153; nn.UpsamplingBilinear2d via
 * Bf16compatible ever/module/ops.py:152-166, fpn.py:168,180; nearest x2 backward fpn.py:100;
 * sum(list)/len fpn.py:189; F.adaptive_avg_pool2d fs_relation.py:177. */
int evb_maxpool3x3s2_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, void* stream);
int evb_maxpool3x3s2_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, void* stream);
/* y[N,f*h,f*w,:C] = bilinear(align_corners)(act(x)); act = bf16(relu(x*scale+shift)) when scale != NULL.
 * ldx / ldy: elements per pixel of the input / output rows. */
int evb_bilinear_up(const void* x, const float* scale, const float* shift, void* y, int N, int h, int w, int C, int ldx,
                    int ldy, int f, void* stream);
int evb_bilinear_up_bwd(const void* dy, void* dx, int N, int h, int w, int C, int lddy, int lddx, int f, void* stream);
/* separable two-pass variant (fewer taps, coalesced); ws: fp32 [N, f*h, w, C] */
long long evb_bilinear_up_bwd_workspace(int N, int h, int w, int C, int f);
int evb_bilinear_up_bwd_sep(const void* dy, void* dx, int N, int h, int w, int C, int lddy, int lddx, int f, void* ws,
                            long long ws_bytes, void* stream);
int evb_sumpool2(const void* dfine, void* dcoarse, int N, int h, int w, int C, int accumulate, void* stream);
int evb_merge4(const void* a, const void* b, const void* c, const void* d, void* out, long long numel, void* stream);
int evb_scale_add(const void* x, float alpha, const void* z, void* y, long long numel, void* stream);
int evb_gap_fwd(const void* x, float* out, int N, int HW, int C, void* stream);
int evb_gap_bwd(const float* dscene, void* dx, int N, int HW, int C, void* stream);
int evb_copy2d_f32(const float* src, int lds, float* dst, int ldd, int rows, int cols, int accumulate, void* stream);
/* dst[0:nbytes] = 0 on the stream (cudaMemsetAsync: a memset node in a captured step, not a kernel) */
int evb_zero_bytes(void* dst, long long nbytes, void* stream);

/* ---- FS-Relation (FSRelation.forward, ever/module/fs_relation.py:57-73) and the scene-embedding MLP (:22-28) */
int evb_relation_fwd(const void* u1, const void* u2, const float* scale1, const float* shift1, const float* scale2,
                     const float* shift2, const float* sf, void* z, float* rel, long long M, int HW, int C, void* stream);
long long evb_relation_bwd_workspace(long long M, int HW, int C);
int evb_relation_bwd(const void* dz, const void* u1, const void* u2, const float* scale1, const float* shift1,
                     const float* scale2, const float* shift2, const float* sf, const float* rel, void* g1, void* g2,
                     float* dsf, long long M, int HW, int C, void* ws, void* stream);
/* FSRelationV2 (ever/module/fs_relation.py:76-163).  Relation with an fp32 scene vector / fp32 product (the scene encoder
 * ends in GroupNorm + ReLU, which autocast runs in fp32) and strided rows: z is written into the first C of ldz elements per
 * pixel (the [r * p, p] concatenation buffer, :156), dz is read with row stride lddz. */
int evb_relation_fwd_v2(const void* u1, const void* u2, const float* scale1, const float* shift1, const float* scale2,
                        const float* shift2, const float* sf, void* z, int ldz, float* rel, long long M, int HW, int C,
                        void* stream);
int evb_relation_bwd_v2(const void* dz, int lddz, const void* u1, const void* u2, const float* scale1, const float* shift1,
                        const float* scale2, const float* shift2, const float* sf, const float* rel, void* g1, void* g2,
                        float* dsf, long long M, int HW, int C, void* ws, void* stream);
/* y = relu(GroupNorm(G)(x)) on the N x C scene vector (nn.GroupNorm(32, C) + nn.ReLU, fs_relation.py:89-94); stat[N][G][2] =
 * {mean, rstd} saved for the backward; dgamma / dbeta (+)= sums over the batch in fixed order */
int evb_groupnorm_relu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stat, int N, int C, int G,
                           float eps, int round_out_bf16, void* stream);
int evb_groupnorm_relu_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* stat, float* dx,
                           float* dgamma, float* dbeta, int N, int C, int G, int accumulate, void* stream);
/* dst[r][0:cols] (+)= src[r][0:cols] for bf16 rows with strides lds / ldd (elements; all % 8): torch.cat(dim=1) of NHWC
 * tensors and its backward split (fs_relation.py:156) */
int evb_copy2d_bf16(const void* src, int lds, void* dst, int ldd, long long rows, int cols, int accumulate, void* stream);
/* y[n,hw,c] = bf16(x[n,hw,c] * m[n,c]): nn.Dropout2d (fs_relation.py:101,118) forward and backward; m = the {0, 1/(1-p)}
 * channel mask the caller draws with torch's own generator (same Philox stream as the reference's feature_dropout) */
int evb_channel_scale(const void* x, const float* m, void* y, int N, long long HW, int C, void* stream);
/* y = keep != 0 ? bf16(x * scale) : 0 on bf16 tensors of numel elements: nn.Dropout (AssymetricDecoder's classifier
 * dropout_rate, ever/module/fpn.py:175-176,190) forward and backward; keep = the survivors drawn with torch's generator */
int evb_dropout_apply(const void* x, const void* keep, float scale, void* y, long long numel, void* stream);
int evb_linear_fwd(const float* x, const float* W, const float* b, float* y, int N, int I, int O, int relu, void* stream);
int evb_linear_bwd(const float* dy, const float* y, const float* x, const float* W, float* dW, float* db, float* dx, int N,
                   int I, int O, int relu, int acc_w, int acc_x, void* stream);

/* ---- loss: F.cross_entropy(ignore_index=255) + dice_loss_with_logits (ever/module/loss.py:54-75, select :26-37,
 * dice_coeff :40-51).  stats = {sum -log p_t, n_valid, I_c[K], sum p_c[K], sum y_c[K]}; the caller may all-reduce
 * stats+2 (3K floats) across ranks between evb_loss_stats and evb_loss_finalize (all_reduce_sum, loss.py:20-23). */
long long evb_loss_workspace(long long P, int K);
int evb_loss_stats(const void* logits, const void* labels, long long P, int K, int LD, int ignore_index, float* stats,
                   void* ws, void* stream);
int evb_loss_finalize(const float* stats, const float* dice_stats, int K, float smooth, float ce_weight, float dice_weight,
                      float dice_grad_scale, float* losses, float* coef, void* stream);
int evb_loss_grad(const void* logits, const void* labels, long long P, int K, int LD, int ignore_index, const float* coef,
                  void* dlogits, void* stream);
/* eval: prob[N,K,H,W] fp32 = softmax, mask[P] uint8 = argmax (logit.softmax(dim=1)) */
int evb_softmax_nchw(const void* logits, float* prob, void* mask, long long P, int HW, int K, int LD, void* stream);

/* ---- optimizer step over flat fp32 arenas: clip_grad_norm_ + torch.optim.SGD + zero_grad
 * (ERModule.apply_gradients / clip_grad, ever/interface/module.py:83-108; ever/opt/optimizer.py:7-9) */
long long evb_sgd_workspace(long long n);
int evb_grad_norm(const float* g, long long n, float max_norm, float* norm_out, void* ws, void* stream);
int evb_sgd_step(float* w, float* g, float* mom, long long n, const float* lr, float momentum, float wd,
                 const float* clip, int first_step, int zero_grad, void* stream);
/* same with a uint8 per-slot mask (1 = trainable): slots of parameters with requires_grad=False (ResNetEncoder freeze_at /
 * frozen BN, ever/module/resnet.py:155-173) are skipped entirely, as torch.optim.SGD skips parameters whose grad is None */
int evb_sgd_step_masked(float* w, float* g, float* mom, long long n, const float* lr, float momentum, float wd,
                        const float* clip, int first_step, int zero_grad, const unsigned char* trainable, void* stream);

/* ---- index-map kernels either side of the path.
 * Map rows are int32[12] in device memory: {a00, a01, b0, a10, a11, b1, s, y0, y1, x0, x1, cb}.
 * evb_pixel_gather: dst[n,i,j,:] = src[s][a00*i+a01*j+b0][a10*i+a11*j+b1][:] when that source pixel lies in
 * [y0,y1) x [x0,x1), else `fill` (per element of elem_bytes).  One launch does any composition of torch.rot90 / torch.flip /
 * transpose / crop / constant pad on a batch of pixel-interleaved images or label maps: THRandomRotate90k,
 * THRandomHorizontalFlip, THRandomVerticalFlip, THRandomCrop (ever/preprocess/thsegm.py:7-147), th_divisible_pad /
 * th_pad_to_size (ever/preprocess/function.py:35-83), the TTA transforms (ever/magic/transform/segm.py:9-72). */
int evb_pixel_gather(const void* src, int Ns, int Hs, int Ws, int pix_bytes, int elem_bytes, long long fill,
                     const void* maps, void* dst, int N, int Ho, int Wo, void* stream);
/* canvas[B,K,Hc,Wc] fp32 += prob[*,K,h,w] tiles (a map row sends pixel (y,x) inside [y0,y1)x[x0,x1) of canvas `cb` (12th
 * int of the row) to tile s pixel (a00*y+a01*x+b0, a10*y+a11*x+b1)); rows are summed in table order (deterministic, no
 * atomics); count[B,Hc,Wc] += coverage.
 * The accumulation a caller of sliding_window() (ever/magic/bigimage/sliding_window.py:8-33) does on the host, and
 * `sum(outs)` of tta() (ever/magic/transform/tta.py:11-23) with the inverse transforms folded into the maps. */
int evb_canvas_accumulate(const float* prob, int N, int K, int h, int w, const void* maps, float* canvas, float* count,
                          int B, int Hc, int Wc, int ylo, int yhi, int xlo, int xhi, void* stream);
/* prob[B,K,P] = canvas / count[B,P] (count NULL: / uniform_count, tta.py:21), mask[B,P] uint8 = argmax (K == 1: > 0.5) */
int evb_canvas_finalize(const float* canvas, const float* count, float uniform_count, int B, int K, long long P,
                        float* prob, void* mask, void* stream);

/* dst[P,ho,wo] (+)= bilinear resize with align_corners = true of the fp32 planes src[P,h,w]: the TTA `Scale` transform and
 * its inverse (ever/magic/transform/segm.py:71-88, F.interpolate(mode='bilinear', align_corners=True)). */
int evb_resize_bilinear_ac(const float* src, long long P, int h, int w, float* dst, int ho, int wo, int accumulate,
                           void* stream);

/* ---- grouped convolution (ResNeXt bottleneck conv2, ever/module/_resnets.py:21-24,80-84,291-324) on the dense kernels.
 * evb_group_expand: dense[Co][Ci][kk] fp32 <- the block-diagonal image of w[Co][Ci/groups][kk] (only the diagonal blocks are
 * written: the caller zero-fills `dense` once).  evb_group_extract: gw[Co][Ci/groups][kk] (+)= the diagonal blocks of the
 * dense weight gradient evb_conv2d_wgrad produced. */
int evb_group_expand(const float* w, float* dense, int Co, int Ci, int kk, int groups, void* stream);
int evb_group_extract(const float* dense, float* gw, int Co, int Ci, int kk, int groups, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EVB200_H_ */

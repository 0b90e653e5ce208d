/* libeosvos_b200.so -- C ABI of the B200-native e-OSVOS hot path.
 *
 * The reference (dvl-tum/e-osvos) has no FFI/plugin layer of its own: its arithmetic lives in
 * PyTorch 1.2 / torchvision 0.4 / cuDNN.  Each entry point below
 * therefore cites the reference *call site* whose library call it replaces.  Conventions:
 *   - the caller owns every buffer; pointers are device pointers unless stated otherwise;
 *   - the CUDA stream is passed explicitly (eosvos_stream_t == cudaStream_t);
 *   - return 0 on success, a negative code on failure; eosvos_last_error() gives the message;
 *   - no hidden allocation, no hidden synchronisation, nothing throws across the boundary;
 *   - sm_100a only: there is no CPU or other-architecture fallback.
 * Activations / tensor-core operands are NHWC 16-bit floats: IEEE fp16 in the default build, bfloat16 with
 * -DEOSVOS_ACT_BF16 (eosvos_act_dtype() tells which); parameters, gradients and losses are fp32.  Entry
 * points that produce parameter gradients take `alpha`, applied to every accumulated value (the inverse
 * of the static loss scale the fp16 backward runs under).
 */
#ifndef EOSVOS_B200_H
#define EOSVOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EOSVOS_B200_VERSION 100

typedef void* eosvos_stream_t;

/* epilogue flags for the contraction entry points */
#define EOSVOS_FLAG_RELU 1      /* y = max(y, 0)                                             */
#define EOSVOS_FLAG_OUT_FP32 2  /* write fp32 instead of bf16                                */
#define EOSVOS_FLAG_RES_HALF 4  /* residual operand lives on the 2x coarser grid (FPN)       */

const char* eosvos_last_error(void);
int eosvos_version(void);
int eosvos_device_check(int device);
unsigned long long eosvos_launch_count(void);
int eosvos_act_dtype(void); /* 0 = bfloat16, 1 = float16 (default build) */

/* ---- K1: dense contractions on tcgen05 (reference: cuDNN conv / ATen addmm reached from
 *      src/networks/mask_rcnn.py:716 forward and src/meta_optim/meta_optim.py:202-204 backward) */

/* y[N,Ho,Wo,Cout] = act(conv(x[N,H,W,Cin], w[Cout,KH,KW,Cin]) + bias + res).
 * gn_sum (optional, fp32 [N][32][2], caller-zeroed): per-(image, group) sum / sum of squares of
 * the fp32 outputs, for the GroupNorm that follows (mask_rcnn.py:523-534). */
int eosvos_conv2d_fprop(const void* x, const void* w, const float* bias, const void* res, void* y, float* gn_sum,
                        int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int flags,
                        int bn_hint, eosvos_stream_t stream);
/* dx[N,H,W,Cin] = conv_transpose(dy[N,Ho,Wo,Cout], wt[Cin,KH,KW,Cout]) (+ acc[N,H,W,Cin], optional, stride 1 only:
 * the gradient that reached the same activation through another branch -- autograd's separate sum kernel fused
 * into the epilogue) */
int eosvos_conv2d_dgrad(const void* dy, const void* wt, void* dx, const void* acc, int N, int H, int W, int Cin,
                        int Cout, int KH, int KW, int stride, int pad, int flags, int bn_hint, eosvos_stream_t stream);
/* dw (fp32) += alpha * x (*) dy ; the caller zeroes dw.  dw_layout 0: memory order [Cout][Cin][KH][KW] (torch
 * contiguous); 1: [Cout][KH][KW][Cin] (torch channels_last strides of the same logical [Cout,Cin,KH,KW] tensor:
 * the input channel is innermost, so the epilogue issues 16-byte vector reductions) */
int eosvos_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int KH,
                        int KW, int stride, int pad, float alpha, int bn_hint, int split_hint, int dw_layout,
                        eosvos_stream_t stream);
/* dw[m * s_m + (n / n_inner) * s_n_outer + (n % n_inner) * s_n_inner] += sum_r dy[r][m] * x[r][n]  for n < n_valid
 * (Linear / stem-im2col weight gradients with an arbitrary destination layout; n_valid 0 = all n_cols columns,
 * smaller when trailing x columns are zero padding) */
int eosvos_gemm_wgrad(const void* x, const void* dy, float* dw, long long rows, int n_cols, int m_cols,
                      long long s_m, int n_inner, long long s_n_inner, long long s_n_outer, float alpha,
                      int bn_hint, int split_hint, int n_valid, eosvos_stream_t stream);
/* 2x2 / stride-2 transposed convolution of the mask head (tv MaskRCNNPredictor.conv5_mask) */
int eosvos_deconv2x2_fprop(const void* x, const void* wd, const float* bias4, void* y, int N, int h, int w, int Cin,
                           int Cout, int flags, int bn_hint, eosvos_stream_t stream);
int eosvos_deconv2x2_dgrad(const void* dy, const void* wdt, void* dx, int N, int h, int w, int Cin, int Cout,
                           int flags, int bn_hint, eosvos_stream_t stream);
int eosvos_deconv2x2_wgrad(const void* x, const void* dy, float* dw, int N, int h, int w, int Cin, int Cout,
                           float alpha, int bn_hint, int split_hint, eosvos_stream_t stream);

/* ---- K2/K3: GroupNorm(32) (+residual) (+ReLU)  (reference: mask_rcnn.py:523-534 -> ATen native_group_norm) */
int eosvos_gn_stats(const void* x, float* sums, int N, int HW, int C, eosvos_stream_t stream);
int eosvos_gn_apply(const void* x, const float* sums, const float* gamma, const float* beta, const void* res, void* y,
                    int N, int HW, int C, float eps, int relu, eosvos_stream_t stream);
int eosvos_gn_backward(const void* x, const float* sums, const float* gamma, const float* beta, const void* dy,
                       const void* yout, float* part, void* dx, void* dres, float* dgamma, float* dbeta, int N, int HW,
                       int C, float eps, int mask_mode, float alpha, eosvos_stream_t stream);

/* ---- K4: multi-scale RoIAlign (reference: mask_rcnn.py:113,147 -> torchvision::roi_align).  rois [R][5] =
 *      (image index, x1, y1, x2, y2); a NEGATIVE image index marks a padding row of a fixed-size list: zeros forward,
 *      no contribution backward (same convention in eosvos_mask_targets) */
int eosvos_roi_align_fwd(const void* const* feats, const int* Hs, const int* Ws, const float* scales, const float* rois,
                         void* out, int R, int P, int C, int sampling, eosvos_stream_t stream);
int eosvos_roi_align_bwd(float* const* dfeats, const int* Hs, const int* Ws, const float* scales, const float* rois,
                         const void* dout, int R, int P, int C, int sampling, eosvos_stream_t stream);
/* mask targets (reference: mask_rcnn.py:38,70 -> tv project_masks_on_boxes) */
int eosvos_mask_targets(const uint8_t* masks, const float* rois, float* out, int R, int M, int H, int W,
                        eosvos_stream_t stream);

/* ---- K7: fused mask losses, forward + gradient (reference: mask_rcnn.py:24-92, loss_lovasz.py:18-126) */
int eosvos_mask_loss_lovasz(const float* logits, const long long* labels, const float* targets, float* loss_out,
                            float* loss_per_roi, float* dlogits, int R, int Cc, int P, eosvos_stream_t stream);
int eosvos_mask_loss_bce(const float* logits, const long long* labels, const float* targets, float* loss_out,
                         float* dlogits, int R, int Cc, int P, eosvos_stream_t stream);

/* ---- K8: inference tail (reference: tv roi_heads.py:56-82,378-502; helper_func.py:113-121;
 *      mask_rcnn.py:626-632) */
int eosvos_mask_paste_threshold(const float* logits, const int* det_of_chan, const long long* det_label,
                                const float* det_box, float* probs, float* target, int* stats, int B, int K, int H,
                                int W, int M, int Cc, float thresh, eosvos_stream_t stream);
int eosvos_mask_to_bbox(const float* target, int* stats, int B, int K, int H, int W, eosvos_stream_t stream);

/* ---- DAVIS J / F counts of a sequence (stage after the hot path; reference: db_eval_sequence of the external davis
 *      package, called at src/util/helper_func.py:444-458).  pred, gt: [T,H,W] uint8 object ids; objects id0+1..id0+K
 *      (K <= 8 per call); bmap: [T,2,H,W] uint8 scratch; counts: [T,K,6] int32 = intersection, union, boundary pixels
 *      of pred / gt, pred boundary pixels matched within `radius` of a gt boundary pixel, and the converse. */
int eosvos_jf_counts(const unsigned char* pred, const unsigned char* gt, unsigned char* bmap, int* counts, int T,
                     int id0, int K, int H, int W, int radius, eosvos_stream_t stream);

/* ---- K5: segmented NMS, one segment per (image, FPN level) or per image (reference: mask_rcnn.py:249 ->
 *      tv rpn.py filter_proposals; mask_rcnn.py:392 -> torchvision::nms) */
long long eosvos_nms_scratch_bytes(int num_segments, int max_seg);
int eosvos_nms_segments(const float* boxes, const int* seg_off, int num_segments, int max_seg, float thresh,
                        void* scratch, unsigned char* keep, eosvos_stream_t stream);

/* ---- K5/K6: proposal / detection post-processing on padded, statically shaped buffers (reference:
 *      mask_rcnn.py:237-249 decode + tv rpn.py filter_proposals; :251-332 EXTEND / REPLACE augmentation; :347-420
 *      postprocess_detections; tv roi_heads.py:642-678 assign_targets_to_proposals + box_coder.encode).  See
 *      csrc/rpn.cu for the layouts. */
long long eosvos_rpn_scratch_bytes(int N, const int* hw, int num_levels, int A);
long long eosvos_rpn_scratch_zero_bytes(int N, int num_levels);
int eosvos_rpn_select(const void* const* heads, const int* hw, int num_levels, int A, int N, const float* anchors,
                      const float* image_hw, int pre_nms_top_n, float bbox_clip, float min_size, float score_thresh,
                      void* scratch, float* boxes, float* scores, unsigned char* valid, eosvos_stream_t stream);
int eosvos_rpn_postnms(const int* hw, int num_levels, int A, int N, int pre_nms_top_n, const float* boxes,
                       const float* scores, const unsigned char* valid, const unsigned char* keep, int post_n,
                       int out_stride, int out_offset, float* out_boxes, float* out_scores, int* out_count,
                       eosvos_stream_t stream);
int eosvos_extend_boxes(const int* stats, const int* fallback_stats, const float* rnd, int B, int G, int n_aug,
                        float ratio_w, float ratio_h, float img_w, float img_h, float share, int out_stride,
                        int out_offset, float* out_boxes, eosvos_stream_t stream);
int eosvos_det_top1(const float* head, const float* proposals, int B, int R, int num_classes,
                    const float* coder_weights4, float bbox_clip, float score_thresh, float min_size, float img_w,
                    float img_h, float back_w, float back_h, float* det_box, float* det_score, long long* det_label,
                    int* det_row, float* det_roi, int* chan, eosvos_stream_t stream);
int eosvos_roi_match(const float* proposals, const int* count, const float* gt_boxes, const long long* gt_labels,
                     const int* gt_off, int B, int P, int max_gt, float iou_thresh, float* all_boxes, long long* labels,
                     long long* matched, int* counts, eosvos_stream_t stream);
int eosvos_rpn_anchor_match(const float* anchors, int num_anchors, const float* gt_boxes, const int* gt_off, int N,
                            float fg_iou, float bg_iou, unsigned* gt_best, long long* labels, int* matched, int* counts,
                            eosvos_stream_t stream);
int eosvos_rpn_loss(const void* const* heads, void* const* dys, const int* hw, int num_levels, int A,
                    const long long* sampled, int num_sampled, const long long* labels, const int* matched,
                    const float* anchors, const float* gt_boxes, const int* gt_off, float beta, int mode, float* out,
                    const float* g_obj, const float* g_box, eosvos_stream_t stream);
/* sparse backward of the RPN head over the sampled anchors only (csrc/rpn.cu; the dense equivalent is the cuDNN dgrad +
 * wgrad of tv rpn.py RPNHead reached from meta_optim.py:202-204) */
int eosvos_rpn_sparse_head(const void* const* heads, const void* const* ts, const void* const* fs, const int* Hs,
                           const int* Ws, int num_levels, int A, int C, const long long* sampled, int M,
                           const long long* labels, const int* matched, const float* anchors, const float* gt_boxes,
                           const int* gt_off, float beta, const float* g_obj, const float* g_box, const float* w_cls,
                           const float* w_box, float grad_scale, void* dt, int* ev_pix, void* xg, float* dw_cls,
                           float* db_cls, float* dw_box, float* db_box, eosvos_stream_t stream);
int eosvos_rpn_sparse_scatter(void* const* dfs, const int* Hs, const int* Ws, int num_levels, int C, const int* ev_pix,
                              int M, const void* G, eosvos_stream_t stream);
long long eosvos_roi_sample_scratch_bytes(int B, int rows);
int eosvos_roi_sample(const long long* labels, const long long* table, int B, int rows, int S, int Pmax, void* scratch,
                      long long* inds, long long* pos_in, eosvos_stream_t stream);
int eosvos_roi_encode(const float* all_boxes, const long long* labels, const long long* matched, const float* gt_boxes,
                      const int* gt_off, const long long* inds, int B, int S, int rows_per_image,
                      const float* coder_weights4, float* rois5, long long* out_labels, long long* out_matched,
                      float* reg_targets, eosvos_stream_t stream);

/* ---- K9: MetaOptimizer update (reference: meta_optim.py:177-214, meta_model.py:78-80) */
/* table_dev: int64 [T][8] = (p, g, lr, out, numel, elements per lr row, g_taps, g_cin); g_taps > 1 means the gradient
 * of a [Cout][Cin][taps] filter is stored [Cout][taps][Cin] (channels_last, eosvos_conv2d_wgrad dw_layout 1);
 * chunks_dev: int32 [n][2] = (tensor, chunk index), chunk = eosvos_meta_update_chunk_elems() elements.  out may alias p.
 * nonfinite_flag (optional, device int): OR-ed with 1 when an updated value is Inf / NaN (overflow of the scaled
 * 16-bit backward); never reset by the library. */
int eosvos_meta_update_chunk_elems(void);
int eosvos_meta_update(const long long* table_dev, const int* chunks_dev, int num_chunks, int use_log,
                       int* nonfinite_flag, eosvos_stream_t stream);
/* d loss / d learning rate through one fused update (first-order BPTT of meta_run.py:124-214): see csrc/meta_update.cu
 * for the table ([T][10] int64) and work-list ([n][3] int32) layouts */
int eosvos_lr_grad(const long long* table_dev, const int* work_dev, int num_work, int use_log, eosvos_stream_t stream);
/* ---- K10: outer RAdam step of meta-training (reference: radam.py:28-94, train_meta.py:361-373) */
int eosvos_radam_step(float* p, const float* g, float* m, float* v, long long n, float gscale, float clip, float beta1,
                      float beta2, float one_minus_beta1, float one_minus_beta2, float eps, float lr, float wd,
                      float step_size, int rectified, float clamp_lo, float clamp_hi, int do_clamp,
                      eosvos_stream_t stream);

/* ---- helpers around the kernels (reference: tv transform.py:119-160 etc., see csrc/misc.cu) */
int eosvos_permute_cast(const void* src, void* dst, const long long* dims, const long long* sstride,
                        const long long* dstride, int src_dtype, int dst_dtype, eosvos_stream_t stream);
/* all fp32 -> 16-bit operand layouts of one iteration in one launch; table int64 [T][14] =
 * (src, dst, dims[4], src strides[4], dst strides[4]), chunks int32 [n][2] = (tensor, chunk index) */
int eosvos_permute_cast_multi_chunk_elems(void);
int eosvos_permute_cast_multi(const long long* table_dev, const int* chunks_dev, int num_chunks, eosvos_stream_t stream);
/* Tensor-core operand layouts of many fp32 parameter tensors in one launch (what cuDNN does internally for the
 * reference, src/networks/mask_rcnn.py:716): dst[x*dx + y*dy + z*dz] = (16-bit) src[(x*Y + y)*Z + z], dx == 1 or
 * dy == 1.  table: int64 [n][10] = (src, dst, X, Y, Z, dx, dy, dz, TX, TY) with TX*TY*Z <= tile_elems;
 * tiles: int32 [m][2] = (tensor, tile index), tile = (x / TX) * ceil(Y / TY) + y / TY. */
int eosvos_weight_prep_tile_elems(void);
int eosvos_weight_prep_multi(const long long* table_dev, const int* tiles_dev, int num_tiles, eosvos_stream_t stream);
/* device half of the first-frame augmentation (reference: src/data/custom_transforms.py:40-51,188-211) */
int eosvos_affine_warp_cubic(const float* src, const float* minv, const int* flip, float* dst, int B, int H, int W,
                             eosvos_stream_t stream);
/* Label half of the same augmentation: cv2.warpAffine(gt, M, flags=INTER_NEAREST) after the optional flip, bit for bit
 * (reference custom_transforms.py:57-89 warps the label with cv2 on the host).  src: [H,W] fp32 ids; minv: [B,6] DOUBLE,
 * the inverted matrix as OpenCV forms it; dst: [B,1,H,W]. */
int eosvos_label_warp_nearest(const float* src, const double* minv, const int* flip, float* dst, int B, int H, int W,
                              eosvos_stream_t stream);
int eosvos_transform(const float* img, void* out, int B, int h, int w, int oh, int ow, int Hp, int Wp, int Cs,
                     const float* mean3, const float* std3, eosvos_stream_t stream);
int eosvos_mask_resize_nearest(const uint8_t* src, uint8_t* dst, int G, int h, int w, int oh, int ow,
                               eosvos_stream_t stream);
int eosvos_im2col_stem(const void* x, void* col, int N, int H, int W, int Cs, int KH, int KW, int stride, int pad,
                       int Kp, eosvos_stream_t stream);
int eosvos_maxpool_fwd(const void* x, void* y, uint8_t* argmax, int N, int H, int W, int C, int ksz, int stride, int pad,
                       eosvos_stream_t stream);
int eosvos_maxpool_bwd(const uint8_t* argmax, const void* dy, void* dx, int N, int H, int W, int C, int ksz, int stride,
                       int pad, eosvos_stream_t stream);
int eosvos_subsample2(const void* x, void* y, int N, int H, int W, int C, int backward, eosvos_stream_t stream);
int eosvos_sum2x2(const void* dfine, void* dcoarse, int N, int Hc, int Wc, int C, eosvos_stream_t stream);
int eosvos_relu_bwd(const void* dy, const void* y, void* out, long long numel, eosvos_stream_t stream);
int eosvos_colsum(const void* dy, float* out, long long M, int C, float alpha, eosvos_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EOSVOS_B200_H */

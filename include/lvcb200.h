/*
 * lvcb200.h -- C ABI of liblvcb200.so, the B200 (sm_100a) implementation of the LVC pseudo-label
 * mining hot path (prannaykaul/lvc @ 3b5e5fa).
 *
 * The reference has no C ABI on this path: it is a Python import surface (detectron2.layers /
 * lvc.modeling) over torch / torchvision ops.  Each entry point below therefore cites
 * the reference *Python* interface it replaces; INTEGRATION.md shows the ctypes binding a maintainer adds
 * on the reference side.  Conventions (all entry points):
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless it says "host";
 *   - never allocates, never synchronises the device (except where stated); work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - returns 0 on success, a cudaError_t (>0) on a CUDA failure, or a negative LVCB200_E* code;
 *     lvcb200_last_error() returns a thread-local message;
 *   - inputs are never modified; outputs are fully written;
 *   - empty inputs (0 boxes / 0 rois / 0 queries) are valid and return 0 (roi_align.py / nms.py behaviour).
 */
#ifndef LVCB200_H_
#define LVCB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define LVCB200_ABI_VERSION 2
#define LVCB200_EINVAL (-1)      /* bad argument (shape, alignment, NULL) */
#define LVCB200_EWORKSPACE (-2)  /* workspace too small */
#define LVCB200_EUNSUPPORTED (-3)

int lvcb200_abi_version(void);
const char* lvcb200_last_error(void);
/* number of kernel launches issued through this library by the calling process (bench: gpu_launches) */
int64_t lvcb200_launch_count(void);

/* ---------------------------------------------------------------------------------------------------
 * RoIAlign, reference layout.  Replaces detectron2.layers.roi_align / ROIAlign.forward
 * (detectron2/layers/roi_align.py:14-15,63-108 -> torchvision.ops.roi_align; arithmetic spec
 * detectron2/layers/csrc/ROIAlign/ROIAlign_cuda.cu:12-139).
 * input [N,C,H,W] fp32 contiguous; rois [R,5] = (batch_idx, x1, y1, x2, y2) fp32; output [R,C,ph,pw] fp32.
 * ------------------------------------------------------------------------------------------------- */
int lvcb200_roi_align_nchw_f32(const float* input, int N, int C, int H, int W, const float* rois, int R,
                               int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, int aligned,
                               float* output, void* stream);

/* FPN level assignment.  Replaces assign_boxes_to_levels (detectron2/modeling/poolers.py:23-59).
 * boxes [R,4] fp32 -> levels [R] int64 (level - min_level). */
int lvcb200_assign_boxes_to_levels(const float* boxes, int64_t R, int min_level, int max_level,
                                   int canonical_box_size, int canonical_level, int64_t* levels, void* stream);

/* One feature-map level in channels-last storage: element (n, y, x, c) lives at
 * base[((n * img_stride) + y * row_stride + x) * C_stride + c]   (strides in pixels / elements),
 * which covers both dense NHWC and the zero-bordered planes the conv engine writes. */
typedef struct {
  const void* base;     /* bf16 or fp32, see `dtype` of the call */
  int H, W;             /* valid extent */
  int64_t img_stride;   /* pixels between images */
  int64_t row_stride;   /* pixels between rows */
  int64_t c_stride;     /* elements per pixel (>= C) */
  float spatial_scale;  /* 1/stride of this level */
} lvcb200_fmap;

#define LVCB200_F32 0
#define LVCB200_BF16 1
#define LVCB200_F16 3
#define LVCB200_OUT_NCHW 0 /* [R, C, ph, pw]  (reference order; fc1 weight index c*49+h*7+w) */
#define LVCB200_OUT_NHWC 1 /* [R, ph, pw, C]  (engine order; fc1 weight permuted at load time) */

/* Fused multi-level pooler.  Replaces ROIPooler.forward (detectron2/modeling/poolers.py:191-246):
 * convert_boxes_to_pooler_format + assign_boxes_to_levels + per-level ROIAlign(aligned=True) + scatter,
 * in ONE launch, no host sync.  rois [R,5]; levels_out (optional, may be NULL) [R] int64.
 * out: [R, C*ph*pw] elements of out_dtype in out_layout, row pitch out_pitch elements. */
int lvcb200_roi_pool_fpn(const lvcb200_fmap* levels /*host*/, int n_levels, int in_dtype, int C, const float* rois,
                         int64_t R, int pooled, int sampling_ratio, int canonical_box_size, int canonical_level,
                         int min_level, void* out, int out_dtype, int out_layout, int64_t out_pitch,
                         int64_t* levels_out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * NMS.  Replaces detectron2.layers.nms / batched_nms (detectron2/layers/nms.py:7,10-29 ->
 * torchvision.ops.nms / batched_nms).  boxes [n,4] fp32, scores [n] fp32, idxs [n] int64 or NULL (plain nms).
 * mode: 0 = coordinate trick (boxes + idx*(max+1) in fp32, the branch the reference takes on CUDA for
 *           n <= 25000), 1 = vanilla per-class (n > 25000, and nms.py:22-29 for n >= 40000),
 *       -1 = pick like the reference would on a CUDA device.
 * keep [n] int64: kept original indices in descending score order (ties: lower index first);
 * num_keep: device int64 scalar.  workspace: device scratch of >= lvcb200_batched_nms_workspace(n) bytes.
 * ------------------------------------------------------------------------------------------------- */
size_t lvcb200_batched_nms_workspace(int64_t n);
int lvcb200_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n, float iou_threshold,
                        int mode, int64_t* keep, int64_t* num_keep, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ---------------------------------------------------------------------------------------------------
 * RPN post-processing.  Replaces RPN.predict_proposals (detectron2/modeling/proposal_generator/rpn.py:455-508)
 * = DefaultAnchorGenerator (anchor_generator.py:157-208) + Box2BoxTransform.apply_deltas
 * (box_regression.py:73-110) + find_top_rpn_proposals (proposal_utils.py:13-118), inference branch.
 * Per level l: logits element (n,y,x,a) at logits[n*img_stride + (y*W + x)*pix_stride + a] when
 * row_stride == 0, else at logits[n*img_stride + y*row_stride + x*pix_stride + a]; deltas likewise with
 * (a*4 + coord).  Only the selected top-k anchors are decoded.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* logits;
  const float* deltas;
  int H, W, A;
  int stride;                 /* anchor stride in pixels (4,8,16,32,64) */
  int64_t img_stride_l, row_stride_l, pix_stride_l; /* logits strides (elements) */
  int64_t img_stride_d, row_stride_d, pix_stride_d; /* deltas strides (elements) */
  float cell_anchors[3 * 4];  /* A x (x1,y1,x2,y2), generate_cell_anchors values (A <= 3) */
} lvcb200_rpn_level;

typedef struct {
  int n_images, n_levels;
  int pre_nms_topk, post_nms_topk;
  float nms_thresh, min_box_size;
  float weights[4];           /* Box2BoxTransform weights */
  int nms_mode;               /* as lvcb200_batched_nms */
} lvcb200_rpn_params;

size_t lvcb200_rpn_proposals_workspace(const lvcb200_rpn_params* p /*host*/);
/* image_sizes [n_images,2] int32 (h, w) device.  Outputs: proposals [n_images, post_nms_topk, 4] fp32,
 * prop_logits [n_images, post_nms_topk] fp32, counts [n_images] int32 (rows beyond count are zero). */
int lvcb200_rpn_proposals(const lvcb200_rpn_level* levels /*host*/, const lvcb200_rpn_params* p /*host*/,
                          const int32_t* image_sizes, float* proposals, float* prop_logits, int32_t* counts,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Box-head post-processing.  Replaces FastRCNNOutputs.predict_boxes / predict_probs / inference
 * (lvc/modeling/roi_heads/fast_rcnn.py:440-493), fast_rcnn_inference_single_image (:95-137) and
 * detector_postprocess (detectron2/modeling/postprocessing.py:10-79).
 * cls_logits [R, K+1] (row pitch logit_pitch), box_deltas [R, 4K] (pitch delta_pitch) or [R,4]
 * (class_agnostic), proposals [R,4], roi_image [R] int32 image index of each row (rows grouped by image,
 * at most max_rois_per_image per image; -1 marks a padding row that is skipped).  row_scale (optional) multiplies the logits of row r (cosine head).
 * image_sizes [n_images,2] int32 (h,w) network input size; out_sizes [n_images,2] int32 requested output size.
 * Outputs per image, padded to topk: det_boxes [n_images, topk, 4], det_scores, det_classes (int64),
 * det_rows (int64, row index within the image), det_counts [n_images] int32.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int n_images, num_classes, max_rois_per_image, class_agnostic;
  float weights[4];
  float score_thresh, nms_thresh;
  int topk_per_image;
  int nms_mode;
} lvcb200_det_params;

size_t lvcb200_detections_workspace(const lvcb200_det_params* p /*host*/);
int lvcb200_detections(const float* cls_logits, int64_t logit_pitch, const float* row_scale, const float* box_deltas,
                       int64_t delta_pitch, const float* proposals, const int32_t* roi_image, int64_t R,
                       const lvcb200_det_params* p /*host*/, const int32_t* image_sizes, const int32_t* out_sizes,
                       float* det_boxes, float* det_scores, int64_t* det_classes, int64_t* det_rows, int32_t* det_counts,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Candidate filter right behind the detector (score mode of get_ret_anns, tools/create_coco_dataset_from_dets_all.py:129-193;
 * row a15 / f2).  Inputs are lvcb200_detections' outputs (boxes already in output-image coordinates) plus, per image, the
 * image area height*width as double (pycocotools' float(h)*float(w)), `novel` [num_classes] uint8 (1 = class to mine) and,
 * optionally, `excluded` [n_images, num_classes] uint8 (1 = this image holds the class's few-shot ground truth, :133-134).
 * flags [n_images, topk] int8: 1 = pseudo-label candidate (ignore_qe 0), 2 = ignore region (ignore_qe 1; only with full != 0),
 * 0 = dropped (also every padding slot >= det_counts).  n_keep (optional) [n_images] int32 = number of flags == 1.  topk <= 1024. */
int lvcb200_candidate_filter(const float* det_boxes, const float* det_scores, const int64_t* det_classes, const int32_t* det_counts,
                             const double* image_area, int n_images, int topk, int num_classes, const uint8_t* novel,
                             const uint8_t* excluded, double k_min, double k_max, double ar, int full, int8_t* flags, int32_t* n_keep,
                             void* stream);

/* Class-agnostic box update of the box corrector.  Replaces BoxOnlyLayersCascade.predict_boxes
 * (lvc/modeling/roi_heads/roi_heads_cascade.py:197-211 -> Box2BoxTransform.apply_deltas, box_regression.py:73-110) followed
 * by Boxes.clip (CascadeROIHeads._create_proposals_from_boxes, cascade_rcnn.py:348-369).  boxes [R,4], deltas [R,>=4] (row pitch
 * delta_pitch), roi_image [R] int32, image_sizes [n,2] int32 (h,w), weights4 = host float[4]; clip != 0 clips to the image. */
int lvcb200_apply_deltas_clip(const float* boxes, const float* deltas, int64_t delta_pitch, const int32_t* roi_image,
                              const int32_t* image_sizes, int64_t R, const float* weights4 /*host*/, int clip, float* out,
                              void* stream);

/* ---------------------------------------------------------------------------------------------------
 * kNN label verification.  Replaces run_nearest_neighbours + get_nn_class_confirmatory
 * (tools/run_nearest_neighbours.py:142-162, 214-227): centred cosine similarity against the support bank,
 * top-`topk` neighbours, class votes, mode over the first `knn` votes (smallest class on ties),
 * keep = (mode == detector class).
 * bank [S,D] fp32, bank_cls [S] int64, queries [Q,D] fp32, query_cls [Q] int64.
 * Outputs: top_idx [Q,topk] int64 (bank rows, best first), votes [Q,topk] int64, keep [Q] uint8;
 * top_sim [Q,topk] fp32 optional (NULL to skip).  S <= 4096, topk <= 16, D % 4 == 0.
 * lvcb200_knn_prepare centres/normalises the bank once (after the all-gather); bank_prepared holds
 * lvcb200_knn_prepared_bytes(S,D) bytes.
 * ------------------------------------------------------------------------------------------------- */
size_t lvcb200_knn_prepared_bytes(int S, int D);
int lvcb200_knn_prepare(const float* bank, int S, int D, void* bank_prepared, void* stream);
int lvcb200_knn_verify(const void* bank_prepared, const int64_t* bank_cls, int S, int D, const float* queries,
                       const int64_t* query_cls, int64_t Q, int topk, int knn, int64_t* top_idx, float* top_sim,
                       int64_t* votes, uint8_t* keep, void* stream);
/* QUERY_EXPAND.COSINE_SIM = False (run_nearest_neighbours.py:154-159): neighbours ranked by -cdist(bank, query) (plain Euclidean
 * distance, no centring).  Same buffers and outputs as above (top_sim = -distance); exact fp32 SIMT path. */
int lvcb200_knn_prepare_euclid(const float* bank, int S, int D, void* bank_prepared, void* stream);
int lvcb200_knn_verify_euclid(const void* bank_prepared, const int64_t* bank_cls, int S, int D, const float* queries,
                              const int64_t* query_cls, int64_t Q, int topk, int knn, int64_t* top_idx, float* top_sim,
                              int64_t* votes, uint8_t* keep, void* stream);
/* Same contract and the same (exact fp32) results, with the Q x S x D contraction on the tensor cores.  Default path: queries and
 * bank as bf16 hi/lo pairs, three-term product with fp32 accumulation in TMEM (scores good to 2e-5 |q|), the 12 best scores per
 * query selected in the GEMM's epilogue; queries whose order is not certain at that accuracy are re-scored exactly on their 11
 * candidates (or, if even the candidate set is uncertain, by the exact SIMT kernel).  LVCB200_KNN_TC=1 selects the round-1 path
 * (TF32 scores from the fp32 queries, rigorous candidate superset, exact re-scoring of ~13 rows per query).
 * Needs 64 <= S <= 4096, D % 8 == 0, and a device workspace of lvcb200_knn_tc_workspace(Q, S, D) bytes. */
size_t lvcb200_knn_tc_workspace(int64_t Q, int S, int D);
int lvcb200_knn_tc_select(int version /* 1 or 3; process-wide, for A/B measurements and tests */);
int lvcb200_knn_verify_tc(const void* bank_prepared, const int64_t* bank_cls, int S, int D, const float* queries,
                          const int64_t* query_cls, int64_t Q, int topk, int knn, int64_t* top_idx, float* top_sim,
                          int64_t* votes, uint8_t* keep, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Dense layers: tcgen05 / TMEM / TMA GEMM with row-shifted A operands ("shift-GEMM"), which is how the
 * engine runs every conv of the ResNet-FPN stack (detectron2/layers/wrappers.py:94-98 Conv2d.forward +
 * FrozenBatchNorm2d.forward batch_norm.py:45-65 + relu, resnet.py:195-211), the RPN head (rpn.py:120-139),
 * the box head FCs (lvc/modeling/roi_heads/box_head.py:82-91) and the predictors (fast_rcnn.py:583-598).
 *
 *   D[m, n] = act( sum_t sum_k A[m + shift[t], k] * W[n, t*K + k] + bias[n] + residual[m, n] ) * rowmask[m]
 *
 * A: bf16 [M_rows, K] row pitch lda (elements); rows outside [0, M_rows) read as zero.
 * W: bf16 [N, taps*K] K-major (OIHW conv weight permuted to O,(kh,kw),I with FrozenBN scale folded in).
 * A conv over a zero-bordered channels-last plane is this with shift[t] = (kh-1)*(W+2) + (kw-1).
 * border != NULL zeroes output rows that are border pixels of the plane geometry (plane_h, plane_w = padded
 * extents; rows are n*plane_h*plane_w + y*plane_w + x) so the result is again a valid zero-bordered plane.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int a_dtype;                                  /* LVCB200_BF16; or LVCB200_F32: A and W are fp32, multiplied as TF32 (kNN scores) */
  const void* A; int64_t lda; int64_t M_rows;   /* rows addressable in A (TMA bound) */
  const void* W; int64_t ldw;                   /* [N, taps*K] */
  const float* bias;                            /* [N] fp32 or NULL */
  const void* residual; int64_t ldr;            /* bf16 [M, N] or NULL (added before activation) */
  void* D; int64_t ldd; int d_dtype;            /* LVCB200_BF16, LVCB200_F16 or LVCB200_F32 */
  int64_t M; int N; int K;                      /* GEMM extents (K per tap), K % 64 == 0, N % 16 == 0 */
  int taps; int32_t shift[9];                   /* row shift per tap */
  int relu;                                     /* 0 none, 1 ReLU, 2 exact (erf) GELU (bf16 output, N > 128, no residual: the ViT MLP) */
  int plane_h, plane_w;                         /* 0 = no border masking */
  /* optional (appended; NULL = off): FPN top-down path fused into the lateral conv (fpn.py:128-134): D += nearest-2x-upsample(up),
   * up = the coarser level's zero-bordered bf16 plane [n, up_plane_h, up_plane_w, >= N] (row pitch ldu elements), this plane being
   * exactly twice its interior size.  bf16 output, N % 64 == 0, no residual; added in fp32 before the single bf16 rounding. */
  const void* upsample_add; int64_t ldu; int up_plane_h, up_plane_w;
  /* optional (appended; 0 = off): STRICT mode -- fp32-grade results on the bf16 tensor pipe (the reference is fp32 end to end,
   * wrappers.py:94-98).  Every real operand x is a bf16 pair x = hi + lo (hi = bf16(x), lo = bf16(x - hi)).  A, residual and a
   * bf16 D are pair matrices: hi half at row 0, lo half at row split_rows (a multiple of 128, >= M; M_rows covers both halves);
   * W is [N, taps * 2K] with [W_hi | W_lo] per tap.  D = A_hi W_hi + A_lo W_hi + A_hi W_lo (+ bias + R_hi + R_lo), fp32
   * accumulation; a bf16 D is written as a pair, an fp32 D as is.  No upsample_add, not chainable. */
  int64_t split_rows;
} lvcb200_gemm_desc;

int lvcb200_gemm_bf16(const lvcb200_gemm_desc* d /*host*/, void* stream);

/* Layer chain: ONE persistent launch runs n dependent bf16 -> bf16 layers (a ResNet stage, resnet.py:708-731: every
 * BottleneckBlock.forward :195-211 of the stage) with tile-granular dependencies instead of kernel boundaries.  Layer i may read
 * (as A or residual) the D buffer of any earlier layer of the chain -- matched by base pointer, same leading dimension --
 * or buffers written before the launch; it may not overwrite a buffer an earlier layer of the chain touches.  Each layer is
 * the lvcb200_gemm_bf16 contract restricted to bf16 operands / outputs with N % 64 == 0 and K % 64 == 0, and produces
 * bit-identical results.  lvcb200_gemm_chain_plan is synchronous (encodes the tensor maps, uploads the layer table into the
 * caller's device workspace of lvcb200_gemm_chain_workspace() bytes, 256-byte aligned); lvcb200_gemm_chain_run only enqueues
 * a counter memset and the launch on `stream` (CUDA-graph capturable) and can be repeated while the buffers stay in place. */
typedef struct {
  void* workspace;                              /* device: completion counters | layer table */
  int64_t counter_bytes; int64_t total_tiles;
  int n_layers; int grid;
} lvcb200_chain_plan;
size_t lvcb200_gemm_chain_workspace(const lvcb200_gemm_desc* descs /*host*/, int n);
int lvcb200_gemm_chain_plan(const lvcb200_gemm_desc* descs /*host*/, int n, void* workspace, size_t workspace_bytes,
                            lvcb200_chain_plan* plan /*host, out*/);
int lvcb200_gemm_chain_run(const lvcb200_chain_plan* plan /*host*/, void* stream);

/* Small memory-bound helpers of the conv engine. */
/* Preprocess + 4x4 space-to-depth.  Replaces GeneralizedRCNN.preprocess_image (lvc/modeling/meta_arch/rcnn.py:324-333:
 * (x - mean) / std, zero pad to /32) and prepares the 7x7/2 stem conv (resnet.py:588-590) as a 3x3 shift-GEMM:
 * images = device array of n pointers to planar [3,H_i,W_i] images (fp32 or uint8, BGR), image_sizes [n,2] int32 (h,w);
 * out = zero-bordered bf16 plane [n, Hpad/4 + 2, Wpad/4 + 2, 64] with channel (iy*4+ix)*3 + c (48 used). */
#define LVCB200_U8 2
int lvcb200_stem_s2d4(const void* const* images, int image_dtype, const int32_t* image_sizes, int n, int Hpad, int Wpad,
                      const float* mean, const float* inv_std, void* out, void* stream);
/* 3x3/2 max-pool (pad 1) (resnet.py:591) over the stem output stored space-to-depth:
 * in = plane [n, Ho+2, Wo+2, 4*C] with channel ((Y&1)*2 + (X&1))*C + c of stem pixel (Y,X); out = plane [n, Ho+2, Wo+2, C]. */
int lvcb200_maxpool_s2d(const void* in, int n, int Ho, int Wo, int C, void* out, void* stream);
/* stride-2 subsample of a zero-bordered plane (input side of the stride-2 1x1 convs, resnet.py:163-171, and
 * LastLevelMaxPool p6, fpn.py:165-177). */
int lvcb200_subsample2(const void* in, int n, int H, int W, int C, void* out, void* stream);
/* out += nearest-2x-upsample(top)  (FPN top-down path, fpn.py:131-133), zero-bordered bf16 planes. */
int lvcb200_upsample2_add(const void* top, int n, int Ht, int Wt, int C, void* inout, int H, int W, void* stream);

/* STRICT engine mode helpers (see lvcb200_gemm_desc.split_rows): planes / matrices held as bf16 hi/lo pairs, the lo half
 * `split_rows` rows after the hi half.  stem_s2d4_pair / maxpool_s2d_pair / upsample2_add_pair are the pair forms of the three
 * kernels above (arithmetic in fp32 on hi + lo, result re-split); subsample2 is a pure copy and is simply applied to each half.
 * pair_merge: out[r, c] = hi + lo as fp32 (dense [rows, cols]); pair_split: the inverse.  cols % 8 == 0. */
int lvcb200_stem_s2d4_pair(const void* const* images, int image_dtype, const int32_t* image_sizes, int n, int Hpad, int Wpad,
                           const float* mean, const float* inv_std, void* out, int64_t split_rows, void* stream);
int lvcb200_maxpool_s2d_pair(const void* in, int64_t in_split_rows, int n, int Ho, int Wo, int C, void* out, int64_t out_split_rows,
                             void* stream);
int lvcb200_upsample2_add_pair(const void* top, int64_t top_split_rows, int n, int Ht, int Wt, int C, void* inout,
                               int64_t io_split_rows, int H, int W, void* stream);
int lvcb200_pair_merge(const void* pair, int64_t split_rows, int64_t rows, int cols, float* out, void* stream);
int lvcb200_pair_split(const float* in, int64_t rows, int cols, void* pair, int64_t split_rows, void* stream);
/* out[r] = scale / (||x_r||_2 + eps), the per-row factor of CosineSimOutputLayers.forward (lvc/modeling/roi_heads/fast_rcnn.py:826-829).
 * x [R, C] row pitch ld: LVCB200_BF16 (lo_off != 0: a bf16 pair, lo element = hi element + lo_off) or LVCB200_F32. */
int lvcb200_row_inv_norm(const void* x, int dtype, int64_t lo_off, int64_t R, int C, int64_t ld, float scale, float eps, float* out,
                         void* stream);
/* convert_boxes_to_pooler_format (detectron2/modeling/poolers.py:69-96) on the RPN's padded output: proposals [n, P, 4] + counts [n]
 * -> rois [n*P, 5] = (image, x1, y1, x2, y2), roi_image [n*P] int32 (image index, -1 for padding rows). */
int lvcb200_make_rois(const float* proposals, const int32_t* counts, int n, int P, float* rois, int32_t* roi_image, void* stream);

/* "Next" row: crop front end of the kNN descriptors.  Replaces get_crops_qe (lvc/data/utils.py:485-519) +
 * preprocess_crops (tools/run_nearest_neighbours.py:102-105): image = planar [3,H,W] uint8 or fp32 on the device;
 * geom [n,8] int32 per box = (y0, x0, actual_h, actual_w, top_pad, left_pad, padded_h, padded_w) as produced by
 * lvc_b200.crops.crop_geometry (the python-slicing / get_padding rules of the reference); out [n,3,S,S] fp32,
 * nearest-resized (F.interpolate mode='nearest') and normalised with mean / inv_std when given. */
int lvcb200_crops_qe(const void* image, int image_dtype, int H, int W, const int32_t* geom, int n, int S, const float* mean,
                     const float* inv_std, float* out, void* stream);

/* "Next" row: the DINO ViT-S/8 forward between the crops and the kNN bank (tools/run_nearest_neighbours.py:108-128,
 * 292-295: `crop_features = model(crops)`, model = torch.hub facebookresearch/dino:main `dino_vits8`, a dependency outside the reference
 * tree; architecture restated from its published definition).  The linear layers run on lvcb200_gemm_bf16; these are the kernels between:
 *   vit_patchify : crops [B,3,S,S] fp32 -> patch rows [B*(S/patch)^2, 3*patch*patch] bf16, columns in (c, iy, ix) order (= conv weight order)
 *   vit_assemble : x[b, 0] = cls + pos[0]; x[b, 1 + i] = patch_tokens[b, i] + pos[1 + i]      (bf16 [B*(Np+1), D])
 *   layernorm    : rows of bf16 x (row pitch ldx) -> (x - mean) * rstd * gamma + beta, bf16 or fp32 out (row pitch ldo), eps as given
 *   gelu         : exact (erf) GELU in place over n bf16 values (n % 8 == 0)
 *   attention    : qkv [B*N, 3*H*64] bf16 (q | k | v, each head-major) -> softmax(q k^T * scale) v, out [B*N, H*64] bf16; head_dim 64 */
int lvcb200_vit_patchify(const float* crops, int B, int S, int patch, void* out, void* stream);
int lvcb200_vit_assemble(const void* patch_tokens, const float* cls_token, const float* pos_embed, int B, int Np, int D, void* x, void* stream);
int lvcb200_layernorm(const void* x, int64_t rows, int D, int64_t ldx, const float* gamma, const float* beta, float eps, void* out,
                      int out_dtype, int64_t ldo, void* stream);
int lvcb200_gelu(void* x, int64_t n, void* stream);
int lvcb200_attention(const void* qkv, int B, int N, int H, int head_dim, float scale, void* out, void* stream);
/* The same attention on tcgen05.mma / TMEM (attention_tc.cu): S = Q K^T and O = P V on the 5th-generation tensor cores (Q, K, V tiles by
 * TMA straight out of qkv; V consumed as an MN-major operand), softmax by two threads per query row straight out of TMEM.  qkv / out as
 * for lvcb200_attention; out must be 16-byte aligned. */
int lvcb200_attention_tc(const void* qkv, int B, int N, int H, int head_dim, float scale, void* out, void* stream);

/* "Next" row: training-side users of the path's operators.
 * roi_align_backward: autograd backward of detectron2.layers.roi_align (ROIAlign_cuda.cu:141-306 RoIAlignBackwardFeature):
 *   grad_output [R,C,ph,pw] fp32, rois [R,5] -> grad_input [N,C,H,W] fp32 (zeroed here, then accumulated with RED.ADD).
 * pairwise_iou: detectron2/structures/boxes.py:315-347; iou [G,P] fp32 row-major, bit-exact with the library's fp32 arithmetic.
 * match_boxes: Matcher.__call__ + set_low_quality_matches_ (detectron2/modeling/matcher.py:61-126) over `quality` [G,P] when given,
 *   else over the IoU of gt_boxes [G,4] x boxes [P,4] computed on the fly (the matrix is never materialised).  thresholds = the
 *   n interior thresholds (the reference adds -inf / +inf), labels = n + 1 values in {-1,0,1}.  matches [P] int64 (first maximum),
 *   match_labels [P] int8, matched_vals [P] fp32 optional.  workspace (lvcb200_match_boxes_workspace(G) bytes) is only needed with
 *   allow_low_quality_matches.  G == 0: matches 0, labels[0]. */
int lvcb200_roi_align_backward_nchw_f32(const float* grad_output, const float* rois, int R, int N, int C, int H, int W, int pooled_h,
                                        int pooled_w, float spatial_scale, int sampling_ratio, int aligned, float* grad_input,
                                        void* stream);
int lvcb200_pairwise_iou(const float* boxes1, int64_t G, const float* boxes2, int64_t P, float* iou, void* stream);
size_t lvcb200_match_boxes_workspace(int64_t G);
int lvcb200_match_boxes(const float* gt_boxes, int64_t G, const float* boxes, const float* quality, int64_t P,
                        const float* thresholds /*host*/, int n_thresholds, const int8_t* labels /*host*/,
                        int allow_low_quality_matches, int64_t* matches, int8_t* match_labels, float* matched_vals, void* workspace,
                        size_t workspace_bytes, void* stream);

/* The bank exchange of the kNN (tools/run_nearest_neighbours.py:303-309) over NVLink peer memory instead of an NCCL launch: peer_bufs[r]
 * (HOST array of W <= 16 device pointers, every rank's padded shard [1 + cap, row_floats] fp32 mapped into this process, e.g. torch symmetric
 * memory) -> out [sum rows, row_floats], rank-major; row 0 of every shard (the header) is skipped.  The caller orders the reads behind the
 * peers' writes (device-side barrier).  rows: HOST [W]. */
int lvcb200_gather_rows_p2p(const void* const* peer_bufs, int W, const int32_t* rows, int row_floats, void* out, void* stream);

/* subsample_labels (detectron2/modeling/sampling.py:9-54) for n_vectors label vectors [n_vectors, N] (int64, or int8 when
 * labels_are_int8) with the two torch.randperm draws replaced by caller-supplied random keys [n_vectors, N] uint32: the positives
 * (label != -1 && != bg_label) / negatives (== bg_label) with the smallest (key, index) pairs are taken, in that order
 * (== cls_idx[argsort(keys[cls_idx], stable)[:take]]); num_pos = min(n_pos, int(num_samples * positive_fraction)),
 * num_neg = min(n_neg, num_samples - num_pos).  pos_idx / neg_idx: [n_vectors, num_samples] int64, -1 beyond counts[v] = {num_pos,
 * num_neg} (int32 [n_vectors, 2]).  out_labels (optional, int8 [n_vectors, N]): RPN._subsample_labels (rpn.py:249-266): -1 everywhere,
 * 1 at the sampled positives, 0 at the sampled negatives.  No host synchronisation; num_samples <= 1024. */
int lvcb200_subsample_labels(const void* labels, int labels_are_int8, const uint32_t* keys, int n_vectors, int64_t N, int num_samples,
                             double positive_fraction, int64_t bg_label, int64_t* pos_idx, int64_t* neg_idx, int32_t* counts,
                             int8_t* out_labels, void* stream);

/* RPN.losses (detectron2/modeling/proposal_generator/rpn.py:328-400) before normalisation and loss weights: out2[0] = sum of
 * binary_cross_entropy_with_logits over anchors with gt_labels >= 0, out2[1] = sum of smooth_l1(pred_anchor_deltas -
 * Box2BoxTransform(weights).get_deltas(anchors, gt_boxes), beta) over gt_labels == 1 (beta < 1e-5: L1).  anchors [A,4] (all levels
 * concatenated), logits [N,A], deltas [N,A,4], gt_labels [N,A] int8, gt_boxes [N,A,4] (matched gt per anchor), out2: 2 device doubles. */
int lvcb200_rpn_losses(const float* anchors, const float* logits, const float* deltas, const int8_t* gt_labels, const float* gt_boxes,
                       int64_t N, int64_t A, const float* weights /*host [4]*/, float smooth_l1_beta, double* out2, void* stream);

/* FastRCNNOutputs.losses (lvc/modeling/roi_heads/fast_rcnn.py:267-279, 296-358, 424-438) before the division by R and the loss weights:
 * out2[0] = sum over rows of cross_entropy(cls_logits [R, K+1], gt_classes [R] int64), out2[1] = sum over foreground rows (0 <= gt < K) of
 * smooth_l1(box_deltas[r, 4*gt .. 4*gt+3] - Box2BoxTransform(weights).get_deltas(proposals, gt_boxes), beta); n_delta_cols = 4*K, or 4 for
 * class-agnostic regression.  out2: 2 device doubles. */
int lvcb200_fast_rcnn_losses(const float* cls_logits, const float* box_deltas, int n_delta_cols, const int64_t* gt_classes,
                             const float* proposals, const float* gt_boxes, int64_t R, int num_classes, const float* weights /*host [4]*/,
                             float smooth_l1_beta, double* out2, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* LVCB200_H_ */

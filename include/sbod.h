/*
 * sbod.h — C ABI of libsbod.so, the B200 (sm_100a) detection box pipeline.
 *
 * This is the drop-in boundary for the hot path of shuaiqi361/shape_based_object_detection
 * (SURVEY.md §8): prior<->GT IoU + best-match assignment, gcxgcy encode/decode, regression +
 * classification losses with hard-negative mining, and decode + threshold + NMS + top-k.
 * The reference has no FFI layer of its own (pure Python/PyTorch); every entry point below
 * cites the reference function it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - tensors are dense row-major fp32 / int64 / int32 / uint8 as stated;
 *   - every call is asynchronous on `stream` (a cudaStream_t), never allocates or frees
 *     caller memory, never throws; scratch comes from a caller-supplied workspace whose size
 *     is returned by the matching *_workspace_bytes query;
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = SBOD_ERR_* validation code.
 */
#ifndef SBOD_H_
#define SBOD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sbod_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define SBOD_API __attribute__((visibility("default")))
#else
#define SBOD_API
#endif

#define SBOD_ABI_VERSION 2

#define SBOD_OK 0
#define SBOD_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, unknown enum) */
#define SBOD_ERR_WORKSPACE (-2)   /* workspace too small or misaligned */
#define SBOD_ERR_UNSUPPORTED (-3) /* shape outside what the kernels support */
#define SBOD_ERR_ALIGNMENT (-4)   /* a streamed tensor is not 16-byte aligned */

SBOD_API int sbod_abi_version(void);
SBOD_API const char* sbod_error_string(int code);

/* Process-wide switches for A/B measurements (no reference counterpart). Defaults: all on. */
#define SBOD_OPT_PDL 0           /* programmatic dependent launch between the kernels of one call */
#define SBOD_OPT_PEER_EXCHANGE 1 /* one-shot NVLink exchange of the loss sums (sbod_comm_*) instead of the caller's all-reduce */
SBOD_API int sbod_set_option(int key, int value);

/* ------------------------------------------------------------------------------------------
 * Dense pairwise IoU.
 *   SBOD_IOU_METRICS  : metrics.py:208-252 find_jaccard_overlap (EPS=1e-5 in the denominator,
 *                       zero-size GT rows -> 0, zero-size anchors -> -1, anchor rule wins)
 *   SBOD_IOU_JACCARD  : operators/iou_utils.py:215-233 jaccard (no EPS, no masks, 0/0 -> NaN)
 * a: [A,4] xyxy, b: [B,4] xyxy, out: [A,B].
 * ---------------------------------------------------------------------------------------- */
#define SBOD_IOU_METRICS 0
#define SBOD_IOU_JACCARD 1
#define SBOD_IOU_INTERSECT 2 /* operators/iou_utils.py:192-212 intersect: clamp(min-max,0) area */
SBOD_API int sbod_iou_matrix(const float* a, int A, const float* b, int B, int mode, float* out,
                    sbod_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Box format converters and gcxgcy codec, elementwise over n boxes.
 *   dataset/transforms.py:26-83  xy_to_cxcy, cxcy_to_xy, cxcy_to_gcxgcy, gcxgcy_to_cxcy
 *   operators/iou_utils.py:167-177 point_form (== cxcy_to_xy), :324-368 encode / decode
 * For the codec ops `priors` is [n,4] cxcy; v0/v1 are the two variances: the transforms.py
 * functions hard-code 1/10 and 1/5 (pass v0=0.1f, v1=0.2f and SBOD_CODEC_TRANSFORMS so the
 * arithmetic order matches), iou_utils encode/decode take them as arguments.
 * ---------------------------------------------------------------------------------------- */
#define SBOD_BOX_XY_TO_CXCY 0
#define SBOD_BOX_CXCY_TO_XY 1
SBOD_API int sbod_box_convert(const float* in, float* out, int n, int op, sbod_stream_t stream);

#define SBOD_CODEC_TRANSFORMS 0 /* transforms.py arithmetic: /(pwh/10), log(..)*5 */
#define SBOD_CODEC_IOU_UTILS 1  /* iou_utils.py arithmetic: /(v0*pwh), log(..)/v1; decode -> xyxy */
SBOD_API int sbod_box_encode(const float* boxes, const float* priors_cxcy, float* out, int n, int flavour,
                    float v0, float v1, sbod_stream_t stream);
SBOD_API int sbod_box_decode(const float* locs, const float* priors_cxcy, float* out, int n, int flavour,
                    float v0, float v1, sbod_stream_t stream);

/* Backward of the six functions above with respect to their first argument (in the reference they are
 * differentiable torch expressions: IouLoss(pred_mode='Center') back-propagates through decode,
 * operators/Loss.py:176-178). op: 0 xy_to_cxcy, 1 cxcy_to_xy, 2 cxcy_to_gcxgcy, 3 encode,
 * 4 gcxgcy_to_cxcy, 5 decode; `in` = the forward input (unused for ops 0 and 1), grad_out / grad_in [n,4]. */
SBOD_API int sbod_box_op_bwd(int op, const float* in, const float* priors_cxcy, const float* grad_out,
                    float* grad_in, int n, float v0, float v1, sbod_stream_t stream);

/* Prior / anchor / location tables on the device (the step before the path): the python triple loops of
 * models/SSD300.py:389-443, SSD512.py:417-474, RetinaNet.py:261-300, RefineDet512.py:655-695 (level, row i,
 * column j, box shape -> (cx, cy, w, h), clamp_(0, 1)) and FCOS.compute_location (FCOSDet.py:235-251, centres
 * only: pass n_shapes[l] == 0 for every level, out is then [n_out, 2]). Same arithmetic as the reference
 * (float64, rounded once to fp32). All array arguments are HOST arrays: rows / cols / n_shapes [n_levels],
 * scale [n_levels][4] = (mul_x, div_x, mul_y, div_y) with cx = (j + 0.5) * mul_x / div_x, shapes
 * [sum n_shapes][2] = (w, h). out: device, [n_out, 4] (or [n_out, 2]), n_out = sum rows * cols * n_shapes. */
SBOD_API int sbod_prior_grid(int n_levels, const int32_t* rows, const int32_t* cols, const int32_t* n_shapes,
                    const double* scale, const double* shapes, int clamp01, float* out, long long n_out,
                    sbod_stream_t stream);

/* RefineDet512.offset2bbox, models/RefineDet512.py:643-653: two-stage decode ARM -> ODM -> xyxy.
 * arm/odm: [N,P,4], priors_cxcy: [P,4], out: [N,P,4]. */
SBOD_API int sbod_offset2bbox(const float* arm_locs, const float* odm_locs, const float* priors_cxcy,
                     float* out, int N, int P, sbod_stream_t stream);

/* Self-test of the device-side correctly rounded division the matching kernel uses instead of
 * div.rn (no reference counterpart; metrics.py:247 is a plain torch division): out[i] = a[i] / b[i]
 * by the fast sequence, ref[i] by div.rn, fast_ok[i] = 1 where the kernel would take the fast
 * sequence. All pointers are device pointers. */
SBOD_API int sbod_selftest_div(const float* a, const float* b, long long n, float* out, float* ref,
                               uint8_t* fast_ok, sbod_stream_t stream);

/* RefineDet ODM easy-negative mask: out[i] = softmax(arm_scores[i, :])[1] < theta, arm_scores
 * [n_rows,2] (models/RefineDet512.py:894-895). */
SBOD_API int sbod_arm_easy_negative(const float* arm_scores, long long n_rows, float theta, uint8_t* out,
                           sbod_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Paired (elementwise) IoU family, operators/iou_utils.py:6-164 bbox_overlaps_{iou,giou,diou,ciou}
 * b1,b2: [M,4] xyxy -> out [M]. Backward: grad wrt b1 and b2 given grad_out [M] (CIoU treats
 * alpha / v / arctan as constants exactly as the reference's no_grad block, :86-92).
 * ---------------------------------------------------------------------------------------- */
#define SBOD_PAIR_IOU 0
#define SBOD_PAIR_GIOU 1
#define SBOD_PAIR_DIOU 2
#define SBOD_PAIR_CIOU 3
SBOD_API int sbod_pair_iou_fwd(const float* b1, const float* b2, int M, int kind, float* out,
                      sbod_stream_t stream);
SBOD_API int sbod_pair_iou_bwd(const float* b1, const float* b2, const float* grad_out, int M, int kind,
                      float* grad_b1, float* grad_b2, sbod_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Row losses used stand-alone by operators/Loss.py.
 *   sbod_smooth_l1      : Loss.py:203-226 SmoothL1Loss elementwise term (pred,target [M,4] -> [M,4])
 *   sbod_softmax_focal  : Loss.py:9-38 focal_loss per-row term (logits [M,C], target [M] int64);
 *                         row_out[M] = loss of the target column (the only non-zero column),
 *                         grad_logits (optional, may be NULL) = d(sum row_out)/d logits.
 *   sbod_sigmoid_focal  : Loss.py:41-80 SigmoidFocalLoss (uses columns 1..C-1, class ids 1..C-1)
 *   sbod_bce_focal      : Loss.py:83-103 FocalLoss (one-hot over all C columns, clamped sigmoid,
 *                         BCE-with-logits); row_out[M] = row sums, grad_logits optional
 * ---------------------------------------------------------------------------------------- */
SBOD_API int sbod_smooth_l1(const float* pred, const float* target, int n_elem, float beta, float* out,
                   float* grad_pred /* nullable: d out / d pred */, sbod_stream_t stream);
SBOD_API int sbod_softmax_focal(const float* logits, const int64_t* target, int M, int C, float alpha_fg,
                       float alpha_bg, float gamma, float* row_out, float* grad_logits,
                       sbod_stream_t stream);
SBOD_API int sbod_sigmoid_focal(const float* logits, const int64_t* target, int M, int C, float alpha,
                       float gamma, float* row_out, float* grad_logits, sbod_stream_t stream);
SBOD_API int sbod_bce_focal(const float* logits, const int64_t* target, int M, int C, float alpha,
                   float gamma, float* row_out, float* grad_logits, sbod_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stand-alone greedy NMS == torchvision.ops.nms as called at models/utils.py:145,265 and
 * detect_scripts/detect_tools.py:62,182,202,304,324 (stable descending score order; suppress
 * iff inter/(area_i+area_j-inter) > thr), and operators/iou_utils.py:385-450 nms (top_k cap).
 * boxes [n,4] xyxy, scores [n]; keep_out [n] int64 (indices into the input, kept order),
 * count_out [1] int32. top_k <= 0 means no cap.
 * ---------------------------------------------------------------------------------------- */
SBOD_API size_t sbod_nms_workspace_bytes(int n);
SBOD_API int sbod_nms(const float* boxes, const float* scores, int n, float iou_thr, int top_k,
             int64_t* keep_out, int32_t* count_out, void* workspace, size_t workspace_bytes,
             sbod_stream_t stream);
/* operators/iou_utils.py:453-530 diounms as written (criterion IoU - (d/c)^beta1 with the centre
 * term of :507); same outputs and workspace as sbod_nms. */
SBOD_API int sbod_diou_nms(const float* boxes, const float* scores, int n, float thr, int top_k, float beta1,
                  int64_t* keep_out, int32_t* count_out, void* workspace, size_t workspace_bytes,
                  sbod_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Batched assignment (the loop inlined in every *Loss.forward: models/SSD300.py:501-542,
 * SSD512.py:532-572, RetinaNet.py:409-449, RefineDet512.py:746-785 and :847-886).
 * GT boxes/labels are packed CSR style: gt_boxes [T,4] xyxy, gt_labels [T] int64,
 * gt_offsets [N+1] int32 (image i owns rows gt_offsets[i]..gt_offsets[i+1]).
 * anchors_xy: [P,4] shared priors (per_image_anchors=0) or [N,P,4] (RefineDet ODM, =1).
 * Outputs (all [N,P]): ov = overlap_for_each_prior after the forced-match fill,
 * obj = object_for_each_prior (int32), cls = true_classes (int64, nullable),
 * neg = true_neg_classes (int64, nullable).
 * ---------------------------------------------------------------------------------------- */
SBOD_API size_t sbod_assign_workspace_bytes(int N, int gmax);
SBOD_API int sbod_assign(const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets, int N,
                int gmax, const float* anchors_xy, int per_image_anchors, int P, float thr_pos,
                float thr_neg, float* ov_out, int32_t* obj_out, int64_t* cls_out,
                int64_t* neg_out, void* workspace, size_t workspace_bytes, sbod_stream_t stream);

/* SSD-pytorch style match, operators/iou_utils.py:236-321 (match / match_ious): jaccard without
 * EPS/masks, best prior per GT forced with overlap 2 (no >0 filter, true GT ids),
 * conf = labels[idx]+1, conf[ov<thr]=0. One image. priors_cxcy [P,4]; truths [G,4] xyxy;
 * labels [G] int64. Writes loc_out [P,4] (encoded with variances v0,v1, or the matched xyxy
 * boxes when encode_loc==0 == match_ious) and conf_out [P] int64. */
SBOD_API size_t sbod_match_workspace_bytes(int G, int P);
SBOD_API int sbod_match(float threshold, const float* truths, int G, const float* priors_cxcy, int P,
               float v0, float v1, const int64_t* labels, int encode_loc, float* loc_out,
               int64_t* conf_out, void* workspace, size_t workspace_bytes, sbod_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused train path: assignment + encode/decode + loc loss + classification loss + hard-negative
 * mining, forward and backward. Replaces the bodies of MultiBoxLoss300.forward
 * (models/SSD300.py:477-594), MultiBoxLoss512.forward (SSD512.py:508-626),
 * RetinaFocalLoss.forward (RetinaNet.py:385-506), RefineDetLoss.compute_arm_loss /
 * compute_odm_loss (RefineDet512.py:730-939).
 * ---------------------------------------------------------------------------------------- */
#define SBOD_REG_L1_ELEM_MEAN 0 /* nn.L1Loss(): sum|d| / (4*n_pos)            SSD300.py:465 */
#define SBOD_REG_SMOOTH_L1 1    /* SmoothL1Loss(beta,'mean'): sum / n_pos rows Loss.py:203-226 */
#define SBOD_REG_IOU 2          /* IouLoss('Corner','mean',losstype): sum(1-x)/n_pos Loss.py:164 */
#define SBOD_REG_GIOU 3
#define SBOD_REG_DIOU 4
#define SBOD_REG_CIOU 5

#define SBOD_CLS_CE_MINE_NONPOS 0 /* CE; per image, positives zeroed, top 3*n_pos   SSD512.py:597-623 */
#define SBOD_CLS_CE_MINE_NEG 1    /* CE; per image, all but true_neg==-1 zeroed     RetinaNet.py:476-503 */
#define SBOD_CLS_CE_MINE_BATCH 2  /* CE; batch-global over true_neg==-1 rows        SSD300.py:567-591 */
#define SBOD_CLS_FOCAL_SUM 3      /* focal_loss(pos rows + neg rows), not normalised SSD512.py:587-593 */
#define SBOD_CLS_FOCAL_NORM 4     /* same / sum(n_pos)                              RetinaNet.py:464-472 */

typedef struct sbod_loss_desc {
  /* ---- inputs ---- */
  const float* locs;         /* [N,P,4] predicted gcxgcy offsets */
  const float* scores;       /* [N,P,C] logits, 16-byte aligned base */
  const float* priors_cxcy;  /* [P,4] */
  const float* priors_xy;    /* [P,4] == cxcy_to_xy(priors_cxcy), as the reference ctor computes */
  const float* anchors_xy;   /* NULL, or [N,P,4] xyxy per-image refined anchors (RefineDet ODM,
                                RefineDet512.py:850-851); their cxcy form is derived on the fly */
  const float* gt_boxes;     /* [T,4] xyxy */
  const int64_t* gt_labels;  /* [T] */
  const int32_t* gt_offsets; /* [N+1] */
  const uint8_t* exclude;    /* NULL, or [N,P]: ARM easy negatives (RefineDet512.py:894-899,924) */
  int32_t N, P, C, gmax;     /* gmax >= max GT per image (host knows it from the list lengths) */
  float thr_pos, thr_neg;    /* float32(threshold), float32(threshold - 0.1) */
  int32_t reg_kind, cls_kind;
  int32_t binarize_labels;   /* RefineDet ARM: labels -> (label > 0)   RefineDet512.py:781 */
  int32_t neg_pos_ratio;
  float reg_weight;          /* config.reg_weights (alpha) */
  float smooth_l1_beta;      /* 1/9 */
  float focal_alpha, focal_gamma;
  /* ---- per-prior state written by forward, read by backward; each [N,P] ---- */
  float* ov;     /* overlap_for_each_prior (after forced-match fill) */
  int32_t* obj;  /* object_for_each_prior */
  float* lse;    /* log-sum-exp of the logits row */
  float* ce;     /* cross entropy of the row against its true class */
  uint8_t* sel;  /* bit0: positive, bit1: hard negative / focal negative, bit2: hard-negative mining candidate */
  float* sel_thr; /* [N,2] per image: cross-entropy threshold of the mined negatives (a candidate row is mined iff
                     its CE is above it, or equal to it when the second value is 1); written by forward */
  /* ---- outputs ---- */
  double* partials; /* [N,4]  per image: sum loc, sum conf over positives, sum conf over mined negatives, n_pos */
  double* sums;     /* [4]    batch sums of the same (the only data that crosses GPUs) */
  float* loss;      /* [4]    total, conf, loc, n_pos */
  void* workspace;
  size_t workspace_bytes;
  /* optional [N,P,C]: if set, sbod_loss_forward also zero-fills this buffer (bulk TMA stores riding
   * under the logits stream). Passing the same pointer as grad_scores to sbod_loss_backward then
   * skips the zero-fill pass of the sparse backward. */
  float* grad_scores_prefill;
  /* optional: device descriptor of a communicator (sbod_comm_device_ptr). If set, the forward all-reduces
   * `sums` across the ranks of the communicator INSIDE its last kernel (one-shot exchange over NVLink peer
   * memory) and forms `loss` from the global sums: no separate collective, no sbod_loss_finalize.
   * Not allowed with SBOD_CLS_CE_MINE_BATCH (batch-global mining does not shard). */
  const void* comm;
} sbod_loss_desc;

SBOD_API size_t sbod_loss_workspace_bytes(const sbod_loss_desc* d);
/* Leading bytes of the loss workspace that must be zero before the first call with a given (N, gmax)
 * (counters, per-object keys, ticket queues); every call leaves them zero again. The rest is scratch. */
SBOD_API size_t sbod_loss_workspace_zero_bytes(const sbod_loss_desc* d);
/* The zero region of a workspace must be cleared once before its first use (every call leaves it clean
 * again); the same holds for the sbod_assign / sbod_detect workspaces. */
SBOD_API int sbod_workspace_init(void* workspace, size_t bytes, sbod_stream_t stream);
SBOD_API int sbod_loss_forward(const sbod_loss_desc* d, sbod_stream_t stream);
/* sbod_loss_forward = three kernels chained by programmatic dependent launches: the match + log-sum-exp kernel
 * (streams the logits once), classify_kernel (forced-match override, classes, candidate histogram, foreground
 * rows) and mine_kernel (top-k sum of the candidates without a sort, batch fold, cross-GPU exchange, loss).
 * Profiling / bench hook: launch one stage of the forward (0 = the match kernel; 1 = classify + mine). Stage 0
 * may be repeated; stage 1 must follow before the workspace is used by a full forward again. */
SBOD_API int sbod_loss_forward_stage(const sbod_loss_desc* d, int stage, sbod_stream_t stream);
/* Recompute d->loss from d->sums (after a cross-GPU all-reduce of d->sums). */
SBOD_API int sbod_loss_finalize(const sbod_loss_desc* d, sbod_stream_t stream);
/* grad_loss: device scalar (upstream gradient). grad_locs [N,P,4], grad_scores [N,P,C]. */
SBOD_API int sbod_loss_backward(const sbod_loss_desc* d, const float* grad_loss, float* grad_locs,
                       float* grad_scores, sbod_stream_t stream);
/* Expand the per-prior state into the reference's int64 tensors (tests / debugging):
 * true_classes, true_neg_classes [N,P]. */
SBOD_API int sbod_loss_targets(const sbod_loss_desc* d, int64_t* cls_out, int64_t* neg_out,
                      sbod_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Eval path: activation + decode + clamp + score threshold + per-class NMS + top-k, batched.
 * Replaces detect (models/utils.py:181-297), detect_tools.detect (detect_scripts/detect_tools.py:
 * 100-219) and detect_refine (:222-341).
 * ---------------------------------------------------------------------------------------- */
#define SBOD_ACT_SOFTMAX 0
#define SBOD_ACT_SIGMOID 1
#define SBOD_ACT_NONE 2    /* scores are probabilities already (FCOS.postprocess output) */
#define SBOD_BOX_OFFSET 0 /* gcxgcy wrt priors */
#define SBOD_BOX_CENTER 1 /* cxcy */
#define SBOD_BOX_CORNER 2 /* xyxy (clamped in place in the reference, models/utils.py:224) */

typedef struct sbod_detect_desc {
  float* locs;               /* [N,P,4]; written only when clamp_inplace != 0 */
  const float* scores;       /* [N,P,C] logits, 16-byte aligned base */
  const float* priors_cxcy;  /* [P,4] (SBOD_BOX_OFFSET only) */
  const uint8_t* prior_keep; /* NULL, or [N,P] prior_positives_idx */
  int32_t N, P, C;
  int32_t act_kind, box_kind, clamp_inplace;
  float min_score, max_overlap;
  int32_t top_k;
  float second_nms_thr;      /* < 0: off; detect_tools: 0.7 class-agnostic second NMS */
  int32_t pre_nms_topk;      /* <= 0: off; per-class candidate cap (BASELINE config 3) */
  int32_t class_agnostic;    /* != 0: detect_objects (models/utils.py:87-178, detect_tools.py:10-97, intended
                                semantics - the reference reaches exit()): every prior is one candidate scored
                                with its best foreground probability, ONE class-agnostic NMS, label = arg-max
                                class. PARITY UNPINNED. Excludes second_nms_thr >= 0 and pre_nms_topk > 0. */
  /* outputs; out_cap >= max(top_k, 1) rows per image */
  float* out_boxes;    /* [N,out_cap,4] */
  int64_t* out_labels; /* [N,out_cap] */
  float* out_scores;   /* [N,out_cap] */
  int32_t* out_prior;  /* [N,out_cap] prior index of each detection, -1 for the placeholder */
  int32_t* out_counts; /* [N] */
  int32_t out_cap;
  void* workspace;
  size_t workspace_bytes;
} sbod_detect_desc;

SBOD_API size_t sbod_detect_workspace_bytes(const sbod_detect_desc* d);
/* leading bytes of the workspace that must be zero before the first call (sbod_workspace_init) */
SBOD_API size_t sbod_detect_workspace_zero_bytes(const sbod_detect_desc* d);
SBOD_API int sbod_detect(const sbod_detect_desc* d, sbod_stream_t stream);
/* sbod_detect = three kernels: the bound pass (streams the logits once: per prior an upper bound of its best
 * foreground probability), the refine pass (exact probabilities of the rows whose bound can matter, candidate
 * keys) and the NMS kernel (lazy top-k NMS; it evaluates the remaining rows itself in the rare case that the
 * first band of candidates runs out - exact, no host round trip, no extra launch).
 * Stages, for profiling and for callers that want to enqueue the streaming pass early (e.g. on a side stream):
 *   0 = bound pass + refine, 1 = NMS, 2 = bound pass only, 3 = refine only, 4 = refine + NMS.
 * sbod_detect == stage 2 followed by stage 4 in stream order. */
SBOD_API int sbod_detect_stage(const sbod_detect_desc* d, int stage, sbod_stream_t stream);
/* The class probabilities exactly as sbod_detect evaluates them (softmax / sigmoid over the C columns, as
 * models/utils.py:201-204), scores [N,P,C] -> out [N,P,C]. The kept (class, prior) sequence of sbod_detect is
 * bit-exact GIVEN these score bits; the bits themselves agree with torch's CPU softmax to a few ulp (an
 * accurate exp and a true division, but another summation order), so two same-class candidates whose scores
 * differ by an ulp or two can swap places with respect to a CPU run. Tests use this entry point to check the
 * two statements separately. */
SBOD_API int sbod_detect_probabilities(const float* scores, int N, int P, int C, int act_kind, float* out,
                              sbod_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Anchor-free (FCOS) targets + loss + post-processing. The reference's FCOSLoss / FCOS.postprocess
 * (models/FCOSDet.py:253-270, 311-544) do not run (SURVEY.md §8 a-F); these entry points implement
 * what that file specifies (constants :333-335, centre sampling / min-area assignment :343-474,
 * centerness target :479-486, loss composition :527-544) with the canonical FCOS meaning where a
 * line cannot execute. PARITY UNPINNED with respect to the reference (pinned to oracle/ only).
 *   loss = SigmoidFocal(scores, labels)/(n_pos + N) + reg_weight * sum((1-DIoU)*ctr)/sum(ctr)
 *          + BCEWithLogits(centerness[pos], ctr)
 * locations [P,2] cell centres (normalised image coordinates); loc_aux [P,4] = (centre-sampling radius
 * in x, in y, size-of-interest lo, hi) per location - two radii because a rectangular input (800 x 1333,
 * BASELINE config 5) has different strides in the two normalised axes.
 * Sharding by image: sums[6] = [focal, sum((1-diou)*w), sum(w), bce, n_pos, n_images] is the only data that
 * crosses GPUs - all-reduce it, then sbod_fcos_finalize on every rank (backward reads the reduced sums).
 * ---------------------------------------------------------------------------------------- */
typedef struct sbod_fcos_desc {
  const float* locs;        /* [N,P,4] predicted l,t,r,b distances (normalised) */
  const float* scores;      /* [N,P,C] logits; column 0 is unused (Loss.py:51-58) */
  const float* centerness;  /* [N,P] logits */
  const float* locations;   /* [P,2] */
  const float* loc_aux;     /* [P,4], 16-byte aligned */
  const float* gt_boxes;    /* [T,4] xyxy */
  const int64_t* gt_labels; /* [T] */
  const int32_t* gt_offsets;/* [N+1] */
  int32_t N, P, C;
  int32_t center_sample;
  float reg_weight, focal_alpha, focal_gamma;
  int32_t* lab;   /* [N,P] label target, 0 = background (written by forward) */
  float* tgt;     /* [N,P,4] l,t,r,b target of the assigned object */
  double* sums;   /* [6] focal, sum((1-diou)*w), sum(w), bce, n_pos, n_images */
  float* loss;    /* [4] total, conf, loc, center */
  void* workspace;
  size_t workspace_bytes;
  const void* comm; /* optional communicator (see sbod_loss_desc.comm): sums all-reduced inside sbod_fcos_forward */
} sbod_fcos_desc;
SBOD_API size_t sbod_fcos_workspace_bytes(const sbod_fcos_desc* d);
SBOD_API int sbod_fcos_forward(const sbod_fcos_desc* d, sbod_stream_t stream);
/* Recompute d->loss from d->sums (after a cross-GPU all-reduce of d->sums). */
SBOD_API int sbod_fcos_finalize(const sbod_fcos_desc* d, sbod_stream_t stream);
SBOD_API int sbod_fcos_backward(const sbod_fcos_desc* d, const float* grad_loss, float* grad_locs,
                       float* grad_scores, float* grad_center, sbod_stream_t stream);
/* FCOS.postprocess: out_scores = sigmoid(cls) * sigmoid(center)[..., None]; out_locs = xyxy boxes */
SBOD_API int sbod_fcos_postprocess(const float* box_pred, const float* cls_pred, const float* center_pred,
                          const float* locations, int N, int P, int C, float* out_locs,
                          float* out_scores, sbod_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Communicator for the one exchange of the path: the loss sums of a batch that is
 * sharded by image over the GPUs of one node, one process per GPU. No reference counterpart (the
 * reference is single-GPU); it replaces the NCCL all-reduce a host would otherwise issue between the
 * forward and the backward. Mailboxes in each GPU's HBM, mapped into every process with CUDA IPC; the
 * kernels post their sums into all mailboxes with plain stores over NVLink / NVSwitch and add the
 * contributions in rank order (bit-identical on every rank).
 *   sbod_comm_create   on the current device; writes this rank's IPC handle (sbod_comm_handle_bytes() bytes,
 *                      host memory) - gather the handles of all ranks (e.g. torch.distributed.all_gather) -
 *   sbod_comm_connect  opens the peers' mailboxes; afterwards sbod_comm_device_ptr(comm) goes into
 *                      sbod_loss_desc.comm / sbod_fcos_desc.comm.
 *   sbod_comm_allreduce stand-alone sum of k <= 7 doubles (device pointer) through the same mailboxes (tests).
 * Every rank must issue the same sequence of exchanging calls.
 * ---------------------------------------------------------------------------------------- */
SBOD_API size_t sbod_comm_handle_bytes(void);
SBOD_API int sbod_comm_create(int rank, int world, void** comm_out, void* handle_out_host);
SBOD_API int sbod_comm_connect(void* comm, const void* all_handles_host);
SBOD_API const void* sbod_comm_device_ptr(void* comm);
SBOD_API int sbod_comm_allreduce(void* comm, double* vals_dev, int k, sbod_stream_t stream);
SBOD_API int sbod_comm_destroy(void* comm);

/* ------------------------------------------------------------------------------------------
 * metrics.calculate_mAP (metrics.py:8-145): VOC07 11-point interpolated average precision per class
 * of a set of detections against the ground truth, greedy matching in descending-score order.
 * Detections and objects are CSR by image (det_offsets / gt_offsets: [n_images + 1] int32); labels
 * int64 in 1..n_classes-1; true_difficulties uint8; gmax = max objects per image; threshold = IoU a
 * match must exceed (0.5); recall_thresholds11 = HOST array of the 11 recall levels as fp32
 * (torch.arange(0, 1.1, .1)). out_ap: [n_classes-1] device floats, class c at index c-1; the mean is
 * left to the caller. Score ties are ordered by detection index (the reference's unstable sort
 * leaves them unspecified). All other pointers are device pointers.
 * ---------------------------------------------------------------------------------------- */
SBOD_API size_t sbod_map_workspace_bytes(int n_detections);
SBOD_API int sbod_map(const float* det_boxes, const int64_t* det_labels, const float* det_scores,
             const int32_t* det_offsets, int n_detections, const float* true_boxes,
             const int64_t* true_labels, const uint8_t* true_difficulties, const int32_t* gt_offsets,
             int n_objects, int n_images, int gmax, int n_classes, double threshold,
             const float* recall_thresholds11, float* out_ap, void* workspace, size_t workspace_bytes,
             sbod_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * End-to-end helper with HOST buffers (pinned or pageable): copy the inputs to the device, run
 * sbod_loss_forward, copy the loss back - the call a C / cgo / JNI host would make.
 * The device scratch is the caller's (dev_arena, arena_bytes from the *_arena_bytes query).
 * ---------------------------------------------------------------------------------------- */
SBOD_API size_t sbod_loss_forward_host_arena_bytes(const sbod_loss_desc* d, int T);
SBOD_API int sbod_loss_forward_host(const sbod_loss_desc* d_host /* input pointers are HOST pointers */,
                           int T, float* loss_host /* [4] */, void* dev_arena, size_t arena_bytes,
                           sbod_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SBOD_H_ */

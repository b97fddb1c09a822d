/*
 * cvpp.h — C ABI of libcvpp.so, the B200 (sm_100a) detection post-processing library.
 *
 * The reference (calmiLovesAI/ComputerVision.pytorch) is pure Python and has no FFI; its
 * "boundary" for this path is a handful of Python callables.  Each entry
 * point below names the reference callable(s) whose arithmetic it replaces (file:line in the
 * reference tree); INTEGRATION.md shows the ctypes stub a maintainer would add on the
 * reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless stated otherwise; the
 *     library allocates nothing, keeps no state besides idempotent per-device caches (SM count,
 *     shared-memory opt-in) and never synchronises:
 *     all work is enqueued on `stream` (a CUstream / cudaStream_t handle; NULL = legacy
 *     default stream);
 *   - returns CVPP_OK (0) or a negative CVPP_ERR_* code; cvpp_last_error() returns a
 *     thread-local description of the last failure on the calling thread;
 *   - fp32 tensors must be contiguous in their innermost dimension and 16-byte aligned.
 *
 * Candidate key (uint64), the unit the filter kernels emit and the sort orders:
 *     [63:52] class id            (12 bits, nc <= 4096)
 *     [51:21] 0x7fffffff - bits(score)   (score is a non-negative fp32; smaller = better)
 *     [20: 0] anchor / prior index (21 bits)
 * Ascending key order == class ascending, score descending, lower anchor first on ties —
 * the order of torchvision's stable descending sort inside nms (SURVEY.md §8c tie rule).
 */
#ifndef CVPP_H
#define CVPP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cvpp_stream_t;
typedef void* cvpp_event_t; /* a cudaEvent_t / CUevent handle owned by the caller */

#if defined(__GNUC__)
#define CVPP_API __attribute__((visibility("default")))
#else
#define CVPP_API
#endif

enum {
  CVPP_OK = 0,
  CVPP_ERR_INVALID_ARG = -1, /* bad shape / threshold / NULL pointer */
  CVPP_ERR_ALIGNMENT = -2,   /* pointer or stride not 16-byte aligned */
  CVPP_ERR_WORKSPACE = -3,   /* workspace too small */
  CVPP_ERR_CUDA = -4,        /* CUDA runtime error (see cvpp_last_error) */
  CVPP_ERR_UNSUPPORTED = -5  /* configuration outside the compiled kernels */
};

/* batched_nms arithmetic branch (torchvision/ops/boxes.py:51-120) */
enum {
  CVPP_NMS_RULE_TORCHVISION_CPU = 0, /* per image: n > 1000 -> per-class ("vanilla") else coordinate trick */
  CVPP_NMS_RULE_COORD_TRICK = 1,     /* boxes + cls * (max_coord + 1), one class-agnostic pass          */
  CVPP_NMS_RULE_PER_CLASS = 2,       /* per-class nms on raw boxes (also what YOLOv7/SSD/YOLOv3 do)      */
  CVPP_NMS_RULE_TORCHVISION_CUDA = 3 /* batched_nms's switch for CUDA tensors: n > 5000 -> per-class else trick
                                        (boxes.numel() > 20000, torchvision/ops/boxes.py:80); same IoU test    */
};

/* output order of cvpp_nms */
enum {
  CVPP_ORDER_SCORE_DESC = 0, /* all classes merged, score descending, capped at max_det (YOLOv8)   */
  CVPP_ORDER_CLASS_MAJOR = 1 /* class ascending then score descending, no cap (YOLOv7/SSD/YOLOv3)  */
};

/* row layouts written by cvpp_detection_epilogue */
enum {
  CVPP_ROWS_YOLOV8 = 0, /* x1,y1,x2,y2,conf,cls            (non_max_suppression, ultralytics_ops.py:226) */
  CVPP_ROWS_SSD = 1,    /* x1,y1,x2,y2,label,conf          (Ssd.decode_boxes, ssd.py:275-278)            */
  CVPP_ROWS_YOLOV7 = 2, /* x1,y1,x2,y2,obj,class_conf,cls  (YOLOv7._nms, yolo_v7.py:391)                 */
  CVPP_ROWS_FULL = 3,   /* x1,y1,x2,y2,score,cls,anchor    (the all-gather payload)        */
  CVPP_ROWS_COCO = 4,   /* x,y,w,h,score,cls               (COCO json rows, yolo_v8.py:364-372)          */
  CVPP_ROWS_VOC = 5     /* cls,score,int(l),int(t),int(r),int(b) (VOC txt lines, yolo_v8.py:286-296)     */
};

/* box transform of cvpp_detection_epilogue */
enum {
  CVPP_BOX_KEEP = 0,             /* copy det_box                                                          */
  CVPP_BOX_CORRECT = 1,          /* normalised xyxy -> centre/size -> yolo_correct_boxes (original pixels) */
  CVPP_BOX_NORMALISE_CORRECT = 2 /* input-pixel xyxy: / (in_w, in_h) first (yolo_v8.py:233-234), then 1  */
};

CVPP_API int cvpp_version(void);
CVPP_API const char* cvpp_last_error(void);
CVPP_API const char* cvpp_error_name(int code);

/* ---------------------------------------------------------------------------------------------
 * YOLOv8 head decode + confidence filter, fused (kernel 1).
 * Replaces: Detect.forward eval tail  core/models/yolov8/modules.py:434-445
 *           DFL.forward               core/models/yolov8/modules.py:80-82
 *           make_anchors              core/utils/anchor.py:126-145
 *           dist2bbox                 core/utils/bboxes.py:213-222
 *           the candidate filter of non_max_suppression  core/utils/ultralytics_ops.py:190,204,220-226
 *           xywh2xyxy                 core/utils/ultralytics_ops.py:360-375
 * Input: num_levels (<= 4) head levels; element (b, c, cell) of level l is
 *        level_ptr[l][b*batch_stride[l] + c*chan_stride[l] + cell], c in [0, 4*reg_max + nc),
 *        cell = y*W_l + x.  (NCHW level tensors: chan_stride = H*W, batch_stride = C*H*W; the
 *        concatenated x_cat (B,C,A): chan_stride = A, batch_stride = C*A, pointers offset.)
 *        level_ptr / batch_stride / chan_stride / level_h / level_w / level_stride are HOST arrays.
 * Output: cand_key[b*max_cand + i], i < min(cand_count[b], max_cand): keys of the anchors whose
 *         best class score is > conf_thres (arbitrary order); cand_count[b] counts ALL of them
 *         (> max_cand means overflow); box_dense[(b*A + anchor)*4 ..] = x1,y1,x2,y2 in input
 *         pixels for candidate anchors only (other entries untouched).
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_yolov8_decode_filter(const float* const* level_ptr, const int64_t* batch_stride,
                              const int64_t* chan_stride, const int* level_h, const int* level_w,
                              const float* level_stride, int num_levels, int B, int nc, int reg_max,
                              float conf_thres, uint64_t* cand_key, int32_t* cand_count, float* box_dense,
                              int max_cand, cvpp_stream_t stream);

/* The same candidates straight from the INPUTS of the head's last 1x1 convolutions: the
 * convolutions  nn.Conv2d(c2, 4 * reg_max, 1)  and  nn.Conv2d(c3, nc, 1)   core/models/yolov8/modules.py:423-425
 * and their concatenation (:431) run on the tensor cores inside the decode kernel (tcgen05.mma kind::tf32, fp32
 * accumulators in TMEM, the decode as the TMEM epilogue), so the (B, 4*reg_max + nc, A) head tensor never exists in HBM.
 * box_feat[l] (B, c2, H_l, W_l) / cls_feat[l] (B, c3, H_l, W_l): contiguous NCHW fp32 device tensors, 16-byte aligned,
 * H*W a multiple of 4; box_w[l] (4*reg_max, c2), box_b[l] (4*reg_max), cls_w[l] (nc, c3), cls_b[l] (nc): device
 * pointers to the conv weights / biases (the pointer ARRAYS are host arrays).  c2, c3 multiples of 16; nc <= 192.
 * Arithmetic: products of TF32-truncated operands accumulated in fp32 (what cuDNN computes for the reference's fp32
 * conv under torch's default allow_tf32), bias added in fp32, then bit for bit the code of cvpp_yolov8_decode_filter.
 * Outputs exactly as cvpp_yolov8_decode_filter. */
CVPP_API int cvpp_yolov8_head_decode_filter(const float* const* box_feat, const float* const* cls_feat,
                              const float* const* box_w, const float* const* box_b, const float* const* cls_w,
                              const float* const* cls_b, const int* level_h, const int* level_w,
                              const float* level_stride, int num_levels, int B, int c2, int c3, int nc, int reg_max,
                              float conf_thres, uint64_t* cand_key, int32_t* cand_count, float* box_dense,
                              int max_cand, cvpp_stream_t stream);

/* ... and the variant that ALSO materialises the head, head_out (B, 4*reg_max + nc, A) = the reference's x_cat
 * (modules.py:438), for callers that need the raw logits as well (Detect.forward returns (y, x), the loss consumes x). */
CVPP_API int cvpp_yolov8_head_decode_filter_x(const float* const* box_feat, const float* const* cls_feat,
                              const float* const* box_w, const float* const* box_b, const float* const* cls_w,
                              const float* const* cls_b, const int* level_h, const int* level_w,
                              const float* level_stride, int num_levels, int B, int c2, int c3, int nc, int reg_max,
                              float conf_thres, uint64_t* cand_key, int32_t* cand_count, float* box_dense,
                              int max_cand, float* head_out, cvpp_stream_t stream);

/* Same decode, but writes the full y (B, 4+nc, A) = [cx,cy,w,h, sigmoid(cls)] like
 * Detect.forward (modules.py:444) for callers that want the dense tensor. */
CVPP_API int cvpp_yolov8_decode_full(const float* const* level_ptr, const int64_t* batch_stride,
                            const int64_t* chan_stride, const int* level_h, const int* level_w,
                            const float* level_stride, int num_levels, int B, int nc, int reg_max, float* y,
                            cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Confidence filter on an already decoded prediction tensor (B, 4+nc+nm, A), rows
 * cx,cy,w,h,scores...  Replaces non_max_suppression's candidate stage
 * (core/utils/ultralytics_ops.py:190,204,220-226 + xywh2xyxy :360-375).  Outputs as above.
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_pred_filter(const float* pred, int B, int channels, int nc, int64_t A, float conf_thres,
                     uint64_t* cand_key, int32_t* cand_count, float* box_dense, int max_cand,
                     cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Class filter on candidate keys, in place: keeps the keys whose class id is in `classes` (HOST array of
 * n_classes ids) and rewrites cand_count.  A count above max_cand (an overflowed filter) is clamped first.
 * Replaces: x = x[(x[:, 5:6] == torch.tensor(classes)).any(1)]   core/utils/ultralytics_ops.py:229-230
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_keep_classes(uint64_t* cand_key, int32_t* cand_count, int B, int max_cand, const int32_t* classes,
                               int n_classes, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Segmented sort of candidate keys, one segment per image (kernel 2).
 * Replaces: x[:, 4].argsort(descending=True)[:max_nms]   core/utils/ultralytics_ops.py:240
 *           the stable descending sort inside torchvision nms (per class)
 *           torch.unique(class) + per-class gathers        core/algorithms/yolo_v7.py:396-400,
 *                                                          torchvision/ops/boxes.py:112-116
 * keys[b*max_cand ..] (first min(cand_count[b], max_cand) entries) are re-packed SCORE-MAJOR,
 *     [63:33] 0x7fffffff - bits(score) | [32:12] anchor | [11:0] class,
 * and sorted ascending in place: score descending, lower anchor first on ties - the one global order
 * every consumer needs; the per-class partition is a stable counting split inside cvpp_nms.
 * max_nms (>0) truncates each image's list to its max_nms best scores (ultralytics :240) and
 * rewrites cand_count[b].  `rule` is accepted for symmetry with cvpp_nms and ignored.
 * workspace: cvpp_sort_workspace_bytes(B, max_cand) bytes, only touched by segments too large
 * for shared memory.
 * ------------------------------------------------------------------------------------------- */
CVPP_API size_t cvpp_sort_workspace_bytes(int B, int max_cand);
CVPP_API int cvpp_segmented_sort(uint64_t* keys, int32_t* cand_count, int B, int max_cand, int rule, int max_nms,
                        void* workspace, size_t workspace_bytes, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Class-aware greedy NMS + gather of the survivors (kernel 3).
 * Replaces: torchvision.ops.batched_nms / nms (torchvision/ops/boxes.py:20-120, csrc/ops/cpu/nms_kernel.cpp)
 *           at call sites core/utils/ultralytics_ops.py:247-248,257; core/utils/nms.py:69,134;
 *           core/algorithms/yolo_v7.py:407; core/algorithms/ssd.py:267.
 * Input: score-major sorted keys (from cvpp_segmented_sort), counts, box_dense (B, A, 4); class ids
 *        in the keys are < nc.
 * Suppression test: IoU computed in fp32 exactly as torchvision's CPU kernel; j is suppressed by a
 * kept i iff (double)iou > iou_thres.
 * Output rows k < min(det_count[b], max_out), for image b at index b*max_out + k:
 *   det_box (x1,y1,x2,y2), det_score, det_cls, det_anchor.  det_count[b] is the uncapped-by-max_out
 *   number of survivors (after the max_det cap when order == CVPP_ORDER_SCORE_DESC).
 * workspace: cvpp_nms_workspace_bytes(B, max_cand, nc).
 * ------------------------------------------------------------------------------------------- */
CVPP_API size_t cvpp_nms_workspace_bytes(int B, int max_cand, int nc);
CVPP_API int cvpp_nms(const uint64_t* sorted_key, const int32_t* cand_count, const float* box_dense, int B,
             int max_cand, int64_t A, int nc, double iou_thres, int rule, int order, int max_det, int max_out,
             float* det_box, float* det_score, int32_t* det_cls, int32_t* det_anchor, int32_t* det_count,
             void* workspace, size_t workspace_bytes, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused sort + NMS: the same result as cvpp_segmented_sort followed by cvpp_nms, in ONE kernel per
 * batch for every image whose candidates fit shared memory (no global sort: per-class warp sorts, greedy
 * suppression, and - for CVPP_ORDER_SCORE_DESC - a histogram selection of the max_det best survivors).
 * cand_key holds the UNSORTED keys as the filter kernels emit them.  Images that do not fit (or have more
 * than max_nms candidates) are finished by the two-kernel path inside the same call (sorting into the
 * workspace); cand_key / cand_count are never modified.
 * workspace: cvpp_sort_nms_workspace_bytes(B, max_cand, nc).
 * ------------------------------------------------------------------------------------------- */
CVPP_API size_t cvpp_sort_nms_workspace_bytes(int B, int max_cand, int nc);
CVPP_API int cvpp_sort_nms(const uint64_t* cand_key, const int32_t* cand_count, const float* box_dense, int B, int max_cand,
                           int64_t A, int nc, double iou_thres, int rule, int order, int max_det, int max_nms,
                           int max_out, float* det_box, float* det_score, int32_t* det_cls, int32_t* det_anchor,
                           int32_t* det_count, void* workspace, size_t workspace_bytes, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The whole YOLOv8 path in one call: decode+filter -> fused sort+NMS (cvpp_sort_nms) on `stream`.
 * Replaces Detect.forward eval tail + non_max_suppression as chained by YOLOv8.decode_box
 * (core/algorithms/yolo_v8.py:222-227).  Scratch buffers are carved from `workspace`
 * (cvpp_yolov8_workspace_bytes).  max_cand must be >= A (one key per anchor at most): a smaller buffer could
 * overflow and drop keys in a nondeterministic order, so it is rejected with CVPP_ERR_INVALID_ARG.
 * ------------------------------------------------------------------------------------------- */
CVPP_API size_t cvpp_yolov8_workspace_bytes(int B, int64_t A, int max_cand, int nc);
CVPP_API int cvpp_yolov8_postprocess(const float* const* level_ptr, const int64_t* batch_stride,
                            const int64_t* chan_stride, const int* level_h, const int* level_w,
                            const float* level_stride, int num_levels, int B, int nc, int reg_max,
                            float conf_thres, double iou_thres, int rule, int max_det, int max_nms,
                            int max_cand, float* det_box, float* det_score, int32_t* det_cls,
                            int32_t* det_anchor, int32_t* det_count, int32_t* cand_count_out,
                            void* workspace, size_t workspace_bytes, cvpp_stream_t stream);

/* The same call for pipelined callers (several batches in flight on different streams): `inputs_consumed`
 * (nullable) is recorded on `stream` right after the decode+filter kernel - the last reader of the head
 * tensors - so the producer of the NEXT batch may overwrite them while this batch's sort+NMS still runs.
 * Under stream capture the record becomes an external event-record node (cudaEventRecordExternal), so the
 * event can be waited on from outside the replayed graph. */
CVPP_API int cvpp_yolov8_postprocess_ev(const float* const* level_ptr, const int64_t* batch_stride,
                            const int64_t* chan_stride, const int* level_h, const int* level_w,
                            const float* level_stride, int num_levels, int B, int nc, int reg_max,
                            float conf_thres, double iou_thres, int rule, int max_det, int max_nms,
                            int max_cand, float* det_box, float* det_score, int32_t* det_cls,
                            int32_t* det_anchor, int32_t* det_count, int32_t* cand_count_out,
                            void* workspace, size_t workspace_bytes, cvpp_event_t inputs_consumed,
                            cvpp_stream_t stream);

/* cvpp_yolov8_postprocess_ev + the evaluation all-gather with NO extra launch: the CTA of the fused
 * sort+NMS kernel that finished image b also writes its rows (CVPP_ROWS_FULL: x1, y1, x2, y2, score, class, anchor;
 * zero rows past the count) and its count into the gather buffer of every rank - peer_dst[d] (peer mappings, HOST array
 * of n_ranks device pointers) or, when mc_dst is not NULL, ONE multimem.st per 16 bytes to the NVSwitch multicast
 * address of the same buffer.  Layout and visibility exactly as cvpp_detection_epilogue_allgather
 * ([n_ranks][B][max_det][7] rows | [n_ranks][B] counts, this rank at `rank`); the local det_* arrays are written too. */
CVPP_API int cvpp_yolov8_postprocess_gather(const float* const* level_ptr, const int64_t* batch_stride,
                                            const int64_t* chan_stride, const int* level_h, const int* level_w,
                                            const float* level_stride, int num_levels, int B, int nc, int reg_max,
                                            float conf_thres, double iou_thres, int rule, int max_det, int max_nms,
                                            int max_cand, float* det_box, float* det_score, int32_t* det_cls,
                                            int32_t* det_anchor, int32_t* det_count, int32_t* cand_count_out,
                                            void* workspace, size_t workspace_bytes, cvpp_event_t inputs_consumed,
                                            float* const* peer_dst, float* mc_dst, int n_ranks, int rank,
                                            cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * CenterNet decode (kernel 4): heatmap peaks + top-K + box assembly + score mask + optional
 * class-agnostic DIoU-NMS + letterbox inverse, per image.
 * Replaces: CenterNetA.decode_boxes        core/algorithms/centernet.py:271-314
 *           _suppress_redundant_centers    core/algorithms/centernet.py:316-326
 *           _top_k                         core/algorithms/centernet.py:328-338
 *           RegL1Loss.gather_feat          core/loss/centernet_loss.py:37-43
 *           xywh_to_xyxy_torch             core/utils/bboxes.py:29-49
 *           diou_nms / box_diou / box_iou  core/utils/nms.py:9-31, core/utils/iou.py:8-64
 *           reverse_letter_box             core/utils/image_process.py:100-129
 * pred: (B, H, W, nc + 4) NHWC = heat logits | reg (2) | wh (2).  pool_mode 0 reproduces the
 * reference, whose MaxPool2d runs on the NHWC tensor and therefore pools over (x, class).
 * Equal scores are ordered by the lower flat index (y*W + x)*nc + c.  Suppressed cells are never
 * emitted, so with conf_thres <= 0 and fewer than K peaks the reference's zero-score filler rows are
 * absent.  letterbox: NULL or (B, 5) fp32 rows in_w, in_h, left, top, scale (host-computed in double
 * like the reference, then cast).  Outputs: capacity K rows per image, det_count[b] valid.
 * det_box is normalised xyxy in [0,1] when letterbox is NULL.
 * ------------------------------------------------------------------------------------------- */
CVPP_API size_t cvpp_centernet_workspace_bytes(int B, int H, int W, int nc, int K);
CVPP_API int cvpp_centernet_decode(const float* pred, int B, int H, int W, int nc, int K, float conf_thres,
                                   int pool_mode, int use_nms, float nms_thres, const float* letterbox,
                                   float* det_box, float* det_score, int32_t* det_cls, int32_t* det_pixel,
                                   int32_t* det_count, void* workspace, size_t workspace_bytes,
                                   cvpp_stream_t stream);

/* diou_nms(boxes, scores, iou_threshold) (core/utils/nms.py:9-31): greedy, class-agnostic, a later box
 * survives a kept one iff DIoU <= thr (fp32).  keep receives int64 indices in descending score order
 * (ties: lower index first); n <= 16384. */
CVPP_API int cvpp_diou_nms(const float* boxes, const float* scores, int n, float thr, int64_t* keep,
                           int32_t* keep_count, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * SSD prior decode + softmax + per-(prior, class) confidence filter.
 * Replaces: Ssd.decode_boxes   core/algorithms/ssd.py:236-264 (softmax :248, class masks :256-264)
 *           Ssd._parse_mbox_loc core/algorithms/ssd.py:290-325
 * loc (B, P, 4), conf (B, P, nc + 1) logits with column 0 = background, priors (P, 4) ltrb fp32 (the
 * table of Ssd._get_ssd_anchors, ssd.py:482-541).  One key per (prior, class >= 1) whose softmax
 * probability is > conf_thres: class field = class - 1 (the reference's label), anchor field = prior
 * index; box_dense (B, P, 4) holds the decoded, clamped, normalised xyxy of every prior that has a key.
 * Follow with cvpp_segmented_sort and cvpp_nms(rule = CVPP_NMS_RULE_PER_CLASS, order =
 * CVPP_ORDER_CLASS_MAJOR, A = P) for ssd.py:267-278.
 * cvpp_ssd_parse_loc is _parse_mbox_loc alone: out (B, P, 4).
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_ssd_decode_filter(const float* loc, const float* conf, const float* priors, int B, int P, int nc,
                                    float conf_thres, uint64_t* cand_key, int32_t* cand_count, float* box_dense,
                                    int max_cand, cvpp_stream_t stream);
CVPP_API int cvpp_ssd_parse_loc(const float* loc, const float* priors, int B, int P, float* out,
                                cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * YOLOv7 3-scale anchor decode + confidence filter, fused.
 * Replaces: YOLOv7.decode_box  core/algorithms/yolo_v7.py:245-344 (sigmoids, grid/anchor box math,
 *           normalisation, concatenation of the levels) and the candidate stage of YOLOv7._nms
 *           :361 (xywh_to_xyxy_torch core/utils/bboxes.py:29-49), :370 (class max, first index),
 *           :377 (obj * class_conf >= conf).
 * Level l is (B, 3*(5+nc), H_l, W_l) addressed like cvpp_yolov8_decode_filter; channel a*(5+nc)+k is
 * attribute k (tx,ty,tw,th,obj,classes) of anchor a.  level_anchors: HOST (num_levels*3, 2) fp32 anchor
 * (w, h) in input pixels, row 3*l + a (the reference's anchors[anchors_mask[l]]).  Anchor index within an
 * image: level offset + a*H*W + y*W + x (levels in the order given: 20x20, 40x40, 80x80 for the reference).
 * Outputs: one key per anchor with obj*class_conf >= conf_thres (score = that product, class = first
 * argmax), box_dense (B, A, 4) normalised xyxy and aux_dense (B, A, 2) = (obj, class_conf) at candidate
 * anchors.  Follow with cvpp_segmented_sort + cvpp_nms(PER_CLASS, CLASS_MAJOR) for yolo_v7.py:396-413.
 * cvpp_yolov7_pred_filter is the same candidate stage on an already decoded (B, A, 5+nc) tensor, the
 * argument of YOLOv7._nms (yolo_v7.py:348) and yolo7_nms (core/utils/nms.py:87).
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_yolov7_decode_filter(const float* const* level_ptr, const int64_t* batch_stride,
                                       const int64_t* chan_stride, const int* level_h, const int* level_w,
                                       const float* level_anchors, int num_levels, int B, int nc, int input_h,
                                       int input_w, float conf_thres, uint64_t* cand_key, int32_t* cand_count,
                                       float* box_dense, float* aux_dense, int max_cand, cvpp_stream_t stream);
CVPP_API int cvpp_yolov7_pred_filter(const float* pred, int B, int64_t A, int nc, float conf_thres,
                                     uint64_t* cand_key, int32_t* cand_count, float* box_dense, float* aux_dense,
                                     int max_cand, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * YOLOv3 3-scale decode + per-(anchor, class) confidence filter, fused.
 * Replaces: predict_bounding_bbox      core/predict/yolov3_decode.py:12-29
 *           Decoder._yolo_post_process core/predict/yolov3_decode.py:40-51
 *           Decoder.__call__           core/predict/yolov3_decode.py:53-63 (scale concatenation)
 *           the mask of yolo3_nms      core/utils/nms.py:60
 *           generate_yolo3_anchor      core/utils/anchor.py:102-117 (anchor / input size)
 * Level l is (B, 3*(5+nc), H_l, W_l); level_anchors as above (the rows 3*l..3*l+2 of the reshaped
 * cfg.arch.anchor).  Anchor index within a level: (y*W + x)*3 + a.  xy is divided by H for both axes like
 * the reference (:22).  merge_batch != 0 reproduces Decoder's flattening of the batch (:47-50): ONE
 * output image whose anchor index is B*level_offset + b*3*H*W + (y*W + x)*3 + a; cand_count has one entry
 * and box_dense is (1, B*A, 4).  One key per (anchor, class) with sigmoid(obj)*sigmoid(cls) >= conf_thres.
 * cvpp_yolov3_predict_bbox is predict_bounding_bbox alone (dense): box_xy (B,H,W,3,2), box_wh (B,H,W,3,2),
 * confidence (B,H,W,3,1), class_prob (B,H,W,3,nc) for one level; anchors: HOST (3, 2) already normalised.
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_yolov3_decode_filter(const float* const* level_ptr, const int64_t* batch_stride,
                                       const int64_t* chan_stride, const int* level_h, const int* level_w,
                                       const float* level_anchors, int num_levels, int B, int nc, int input_h,
                                       int input_w, float conf_thres, int merge_batch, uint64_t* cand_key,
                                       int32_t* cand_count, float* box_dense, int max_cand, cvpp_stream_t stream);
CVPP_API int cvpp_yolov3_predict_bbox(const float* feature, int B, int nc, int H, int W, const float* anchors,
                                      float* box_xy, float* box_wh, float* confidence, float* class_prob,
                                      cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Candidate mask of the standalone yolo3_nms (core/utils/nms.py:54-60): scores (M, nc) fp32 row-major ->
 * one key (class c, score, row m) per element >= conf_thres; single image (cand_count is one int32).
 * The caller's (M, 4) xyxy boxes serve directly as box_dense for cvpp_nms (A = M).
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_score_matrix_filter(const float* scores, int64_t M, int nc, float conf_thres, uint64_t* cand_key,
                                      int32_t* cand_count, int max_cand, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Row gather: out[b, k, :] = feat[b, ind[b, k], :] for k < (count ? count[b] : K).
 * Replaces: RegL1Loss.gather_feat  core/loss/centernet_loss.py:37-43
 *           gather_op              core/utils/nms.py:34-51 (B = 1)
 * feat (B, N, C) fp32, ind (B, K) int32 or int64 (ind_is_int64), out (B, K, C).  err_flag (nullable,
 * one int32) is set to 1 when an index falls outside [0, N) (torch.gather would raise).
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_gather_feat(const float* feat, const void* ind, int ind_is_int64, const int32_t* count, int B,
                              int64_t N, int C, int K, float* out, int32_t* err_flag, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Detection epilogue: assemble the caller-facing rows from cvpp_nms outputs, with the box mapped back to
 * the original image, batched (one (h, w) per image) and on the device.
 * Replaces: row assembly            ultralytics_ops.py:226,257; yolo_v7.py:391,410; ssd.py:275-278
 *           YOLOv8 normalisation    core/algorithms/yolo_v8.py:233-234
 *           centre/size round trip  yolo_v8.py:237-238; yolo_v7.py:416-417; ssd.py:284-285
 *           yolo_correct_boxes / reverse_letter_box_numpy  core/utils/image_process.py:69-97,161-181
 * letterbox: (B, 5) fp32 rows in_w, in_h, left, top, scale (host-computed in double like the reference,
 * then cast; for letterbox_image == False pass image_w, image_h, 0, 0, 1).  rows: (B, max_out, 6 or 7)
 * by layout; rows >= det_count[b] are zero-filled.  aux_dense / A only for CVPP_ROWS_YOLOV7.  count_out
 * (nullable, [B] fp32) receives min(det_count[b], max_out), so that rows and counts can share one buffer
 * (the single all-gather payload of SURVEY.md 8e).
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_detection_epilogue(const float* det_box, const float* det_score, const int32_t* det_cls,
                                     const int32_t* det_anchor, const int32_t* det_count, const float* aux_dense,
                                     int B, int max_out, int64_t A, int layout, int box_mode, const float* letterbox,
                                     float* rows, float* count_out, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Compact detection epilogue: the same rows, WITHOUT padding - image b's rows start at
 * row_offset[b] = sum over b' < b of min(det_count[b'], max_out); row_offset has B + 1 entries (the last one
 * is the total).  This is the all-gather payload of the heads that have no max_det cap (YOLOv7._nms
 * core/algorithms/yolo_v7.py:348-422, Ssd.decode_boxes ssd.py:236-288, yolo3_nms core/utils/nms.py:54-84):
 * ranks exchange row_offset[B] first, or pass a fixed row_capacity and check `overflow` (nullable, one int32:
 * 1 when the total exceeds row_capacity; rows beyond the capacity are not written).  B <= 65535.
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_detection_epilogue_compact(const float* det_box, const float* det_score, const int32_t* det_cls,
                                             const int32_t* det_anchor, const int32_t* det_count,
                                             const float* aux_dense, int B, int max_out, int64_t A, int layout,
                                             int box_mode, const float* letterbox, float* rows, int64_t row_capacity,
                                             int32_t* row_offset, int32_t* overflow, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Detection epilogue fused with the evaluation all-gather: the same rows as
 * cvpp_detection_epilogue, but every row (and the per-image count) is stored straight into the gather
 * buffer of EVERY rank through NVLink peer mappings - one kernel, no NCCL launch, no staging copy.
 * peer_dst: HOST array of n_peers (<= 16) device pointers, peer_dst[d] = rank d's gather buffer mapped into
 * this process (CUDA IPC / symmetric memory; own rank included), each laid out as
 *     [n_peers][B][max_out][W] fp32 rows | [n_peers][B] fp32 counts,
 * this rank writing slot `rank`.  The stores are ordinary global stores: they are visible to the peers once
 * the kernel has completed, so consumers synchronise on a cross-rank barrier / signal issued after it in
 * stream order (torch symmetric-memory `barrier`, or any later collective).
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_detection_epilogue_allgather(const float* det_box, const float* det_score, const int32_t* det_cls,
                                               const int32_t* det_anchor, const int32_t* det_count,
                                               const float* aux_dense, int B, int max_out, int64_t A, int layout,
                                               int box_mode, const float* letterbox, float* const* peer_dst,
                                               int n_peers, int rank, cvpp_stream_t stream);

/* Same gather, through the NVSwitch: mc_dst is the MULTICAST address of the symmetric gather buffer (one
 * CUmulticastObject bound to every rank's copy - torch symmetric memory exposes it as `multicast_ptr`); every 16
 * bytes leave this GPU once as a `multimem.st` and the switch writes them into all n_ranks copies (own rank
 * included).  Same layout, same visibility rule as cvpp_detection_epilogue_allgather.  Needs NVLS-capable
 * hardware (an NVSwitch box); CVPP_ERR_INVALID_ARG when mc_dst is NULL. */
CVPP_API int cvpp_detection_epilogue_multicast(const float* det_box, const float* det_score, const int32_t* det_cls,
                                               const int32_t* det_anchor, const int32_t* det_count,
                                               const float* aux_dense, int B, int max_out, int64_t A, int layout,
                                               int box_mode, const float* letterbox, float* mc_dst, int n_ranks,
                                               int rank, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * reverse_letter_box (core/utils/image_process.py:100-129) on n boxes (n, 4): xywh != 0 converts (cx,cy,w,h)
 * to corners first; then * (in_w, in_h), - (left, top), * scale, one fp32 rounding per step.  The five scalars
 * are the reference's Python doubles cast to fp32 by the caller.  out may alias boxes.
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_letterbox_reverse(const float* boxes, int64_t n, int xywh, float in_w, float in_h, float left,
                                    float top, float scale, float* out, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * CenterNetA._suppress_redundant_centers (core/algorithms/centernet.py:316-326), dense form:
 * out = heat * (heat == maxpool3x3(heat)) with the pool over (x, channel) of each image row of the
 * (B, H, W, C) NHWC tensor, exactly as the reference's MaxPool2d sees it.  (The fused
 * cvpp_centernet_decode never materialises this tensor.)
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_centernet_suppress(const float* heat, int B, int H, int W, int C, float* out,
                                     cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Per-row top-K of a (B, N) fp32 score matrix, sorted: scores descending, EQUAL scores by the lower
 * flat index (the rule of cvpp_centernet_decode; torch.topk leaves it unspecified), NaN first.
 * Replaces: CenterNetA._top_k   core/algorithms/centernet.py:328-338
 *           (torch.topk(scores.view(B, -1), K) + the index split into class / y / x / y*W+x).
 * out_val (B, K) fp32, out_idx (B, K) int64 flat indices.  With C > 0 the index is also split like the
 * reference, each output nullable: out_cls = idx % C, pixel = idx / C, out_y = pixel / W,
 * out_x = pixel % W (int64, like torch) and out_pixel = y*W + x (int32, reference :337).
 * 1 <= K <= min(N, 4096), N < 2^32.  workspace: cvpp_topk_workspace_bytes(B, K).
 * ------------------------------------------------------------------------------------------- */
CVPP_API size_t cvpp_topk_workspace_bytes(int B, int K);
CVPP_API int cvpp_topk(const float* scores, int B, int64_t N, int K, int C, int W, float* out_val, int64_t* out_idx,
                       int64_t* out_cls, int64_t* out_y, int64_t* out_x, int32_t* out_pixel, void* workspace,
                       size_t workspace_bytes, cvpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Matching of detections to ground truth for the VOC-style mAP, one CTA per image.
 * Replaces the per-detection loop of get_map   core/metrics/mAP.py:441-520 (best-overlap search :486-503 with the
 * "+1" pixel convention in double precision, first maximum wins; MINOVERLAP / difficult / `used` logic :508-520).
 * det_rows (N, 6) fp32 in the CVPP_ROWS_VOC layout (cls, score, l, t, r, b), image b owning rows
 * det_offset[b] .. det_offset[b+1] (what cvpp_detection_epilogue_compact writes), every (image, class) group in
 * descending score order; gt_box (G, 4) l,t,r,b fp32 (16-byte aligned), gt_cls / gt_difficult (G) int32, image b
 * owning gt_offset[b] .. gt_offset[b+1].  Outputs per detection: flag 1 = true positive, 2 = false positive,
 * 0 = matched a difficult box (counts as neither); best_gt = index of the best box within the image's list (-1:
 * no box of the class overlaps); ovmax (double, -1 when none).  claim_ws: G int32 of scratch.
 * ------------------------------------------------------------------------------------------- */
CVPP_API int cvpp_voc_match(const float* det_rows, const int32_t* det_offset, const float* gt_box, const int32_t* gt_cls,
                            const int32_t* gt_difficult, const int32_t* gt_offset, int B, double min_overlap,
                            int32_t* flag, int32_t* best_gt, double* ovmax, int32_t* claim_ws, cvpp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CVPP_H */

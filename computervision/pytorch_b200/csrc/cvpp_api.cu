// cvpp_api.cu — the extern "C" surface of libcvpp.so (declared in include/cvpp.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "cvpp_common.cuh"

namespace cvpp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return CVPP_ERR_CUDA;
}

int device_info(DeviceInfo* out) {
  static DeviceInfo cache[64];
  static bool have[64] = {};
  int dev = 0;
  CVPP_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && have[dev]) {
    *out = cache[dev];
    return CVPP_OK;
  }
  DeviceInfo d;
  d.device = dev;
  CVPP_CUDA_TRY(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
  CVPP_CUDA_TRY(cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (dev >= 0 && dev < 64) {
    cache[dev] = d;  // benign race: every thread writes the same values
    have[dev] = true;
  }
  *out = d;
  return CVPP_OK;
}

int ensure_smem_attr(const void* func, int bytes, int device, unsigned long long* done) {
  const unsigned long long bit = (device >= 0 && device < 64) ? (1ull << device) : 0ull;
  if (bit && (__atomic_load_n(done, __ATOMIC_ACQUIRE) & bit)) return CVPP_OK;
  CVPP_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (bit) __atomic_fetch_or(done, bit, __ATOMIC_RELEASE);
  return CVPP_OK;
}

// launchers (one per .cu)
int yolov8_decode_launch(const float* const* level_ptr, const int64_t* batch_stride, const int64_t* chan_stride,
                         const int* level_h, const int* level_w, const float* level_stride, int num_levels, int B,
                         int nc, int reg_max, float conf_thres, uint64_t* cand_key, int32_t* cand_count,
                         float* box_dense, int max_cand, float* y, int force_generic, cudaStream_t stream);
int pred_filter_launch(const float* pred, int B, int channels, int nc, int64_t A, float conf_thres, uint64_t* cand_key,
                       int32_t* cand_count, float* box_dense, int max_cand, cudaStream_t stream);
int keep_classes_launch(uint64_t* cand_key, int32_t* cand_count, int B, int max_cand, const int32_t* classes, int n_classes,
                        cudaStream_t stream);
size_t segsort_workspace_bytes(int B, int max_cand);
int segsort_launch(uint64_t* keys, int32_t* cand_count, int B, int max_cand, int rule, int max_nms, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream);
size_t nms_workspace_bytes(int B, int max_cand, int nc);
int nms_launch(const uint64_t* sorted_key, const int32_t* cand_count, const float* box_dense, int B, int max_cand,
               int64_t A, int nc, double iou_thres, int rule, int order, int max_det, int max_out, float* det_box,
               float* det_score, int32_t* det_cls, int32_t* det_anchor, int32_t* det_count, void* workspace,
               size_t workspace_bytes, cudaStream_t stream);

size_t sort_nms_workspace_bytes(int B, int max_cand, int nc);
int sort_nms_launch(const uint64_t* cand_key, const int32_t* cand_count, const float* box_dense, int B, int max_cand, int64_t A,
                    int nc, double iou_thres, int rule, int order, int max_det, int max_nms, int max_out, float* det_box,
                    float* det_score, int32_t* det_cls, int32_t* det_anchor, int32_t* det_count, void* workspace,
                    size_t workspace_bytes, cudaStream_t stream, int32_t* cand_count_out = nullptr,
                    float* const* gather_dst = nullptr, float* gather_mc = nullptr, int gather_n = 0, int gather_rank = 0);

size_t centernet_workspace_bytes(int B, int H, int W, int nc, int K);
int centernet_launch(const float* pred, int B, int H, int W, int nc, int K, float conf, int pool_mode, int use_nms,
                     float nms_thr, const float* letterbox, float* det_box, float* det_score, int32_t* det_cls,
                     int32_t* det_pixel, int32_t* det_count, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream);
int diou_nms_launch(const float* boxes, const float* scores, int n, float thr, long long* keep, int32_t* keep_count,
                    cudaStream_t stream);

int ssd_decode_filter_launch(const float* loc, const float* conf, const float* priors, int B, int P, int nc,
                             float conf_thres, uint64_t* cand_key, int32_t* cand_count, float* box_dense, int max_cand,
                             cudaStream_t stream);
int ssd_parse_loc_launch(const float* loc, const float* priors, int B, int P, float* out, cudaStream_t stream);

int yolo_anchor_decode_launch(int mode, const float* const* level_ptr, const int64_t* batch_stride,
                              const int64_t* chan_stride, const int* level_h, const int* level_w,
                              const float* level_anchors, int num_levels, int B, int nc, int input_h, int input_w,
                              float conf_thres, int merge_batch, uint64_t* cand_key, int32_t* cand_count, float* box_dense,
                              float* aux_dense, int max_cand, int force_generic, cudaStream_t stream);
int yolov3_predict_bbox_launch(const float* feature, int B, int nc, int H, int W, const float* anchors, float* box_xy,
                               float* box_wh, float* confidence, float* class_prob, cudaStream_t stream);
int yolov7_pred_filter_launch(const float* pred, int B, int64_t A, int nc, float conf_thres, uint64_t* cand_key,
                              int32_t* cand_count, float* box_dense, float* aux_dense, int max_cand,
                              cudaStream_t stream);
int score_matrix_filter_launch(const float* scores, int64_t M, int nc, float conf_thres, uint64_t* cand_key,
                               int32_t* cand_count, int max_cand, cudaStream_t stream);
int gather_feat_launch(const float* feat, const void* ind, int ind_is_int64, const int32_t* count, int B, int64_t N, int C,
                       int K, float* out, int32_t* err_flag, cudaStream_t stream);
int detection_epilogue_launch(const float* det_box, const float* det_score, const int32_t* det_cls,
                              const int32_t* det_anchor, const int32_t* det_count, const float* aux_dense, int B,
                              int max_out, int64_t A, int layout, int box_mode, const float* letterbox, float* rows,
                              float* count_out, float* const* peer_dst, float* mc_dst, int n_peers, int slot, cudaStream_t stream);

int detection_epilogue_compact_launch(const float* det_box, const float* det_score, const int32_t* det_cls,
                                      const int32_t* det_anchor, const int32_t* det_count, const float* aux_dense, int B,
                                      int max_out, int64_t A, int layout, int box_mode, const float* letterbox, float* rows,
                                      int64_t row_capacity, int32_t* row_offset, int32_t* overflow, cudaStream_t stream);
int letterbox_reverse_launch(const float* boxes, int64_t n, int xywh, float in_w, float in_h, float left, float top,
                             float scale, float* out, cudaStream_t stream);
int centernet_suppress_launch(const float* heat, int B, int H, int W, int C, float* out, cudaStream_t stream);

size_t topk_workspace_bytes(int B, int K);
int topk_launch(const float* scores, int B, long long N, int K, int C, int W, float* out_val, long long* out_idx,
                long long* out_cls, long long* out_y, long long* out_x, int32_t* out_pixel, void* workspace,
                size_t workspace_bytes, cudaStream_t stream);

int voc_match_launch(const float* det_rows, const int32_t* det_offset, const float* gt_box, const int32_t* gt_cls,
                     const int32_t* gt_difficult, const int32_t* gt_offset, int B, double min_overlap, int32_t* flag,
                     int32_t* best_gt, double* ovmax, int32_t* claim_ws, cudaStream_t stream);

int yolov8_head_fused_launch(const float* const* box_feat, const float* const* cls_feat, const float* const* box_w,
                             const float* const* box_b, const float* const* cls_w, const float* const* cls_b,
                             const int* level_h, const int* level_w, const float* level_stride, int num_levels, int B, int c2,
                             int c3, int nc, int reg_max, float conf_thres, uint64_t* cand_key, int32_t* cand_count,
                             float* box_dense, int max_cand, float* head_out, cudaStream_t stream);

static int force_generic() {
  const char* e = getenv("CVPP_FORCE_GENERIC");
  return e && e[0] == '1';
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace cvpp

using namespace cvpp;

extern "C" {

int cvpp_version(void) { return 100; }

const char* cvpp_last_error(void) { return g_err; }

const char* cvpp_error_name(int code) {
  switch (code) {
    case CVPP_OK: return "CVPP_OK";
    case CVPP_ERR_INVALID_ARG: return "CVPP_ERR_INVALID_ARG";
    case CVPP_ERR_ALIGNMENT: return "CVPP_ERR_ALIGNMENT";
    case CVPP_ERR_WORKSPACE: return "CVPP_ERR_WORKSPACE";
    case CVPP_ERR_CUDA: return "CVPP_ERR_CUDA";
    case CVPP_ERR_UNSUPPORTED: return "CVPP_ERR_UNSUPPORTED";
    default: return "CVPP_ERR_UNKNOWN";
  }
}

int cvpp_yolov8_decode_filter(const float* const* level_ptr, const int64_t* batch_stride, const int64_t* chan_stride,
                              const int* level_h, const int* level_w, const float* level_stride, int num_levels,
                              int B, int nc, int reg_max, float conf_thres, uint64_t* cand_key, int32_t* cand_count,
                              float* box_dense, int max_cand, cvpp_stream_t stream) {
  if (!(conf_thres >= 0.0f && conf_thres <= 1.0f)) {
    set_error("Invalid Confidence threshold %f, valid values are between 0.0 and 1.0", conf_thres);
    return CVPP_ERR_INVALID_ARG;
  }
  return yolov8_decode_launch(level_ptr, batch_stride, chan_stride, level_h, level_w, level_stride, num_levels, B, nc,
                              reg_max, conf_thres, cand_key, cand_count, box_dense, max_cand, nullptr, force_generic(),
                              (cudaStream_t)stream);
}

int cvpp_yolov8_head_decode_filter(const float* const* box_feat, const float* const* cls_feat, const float* const* box_w,
                                   const float* const* box_b, const float* const* cls_w, const float* const* cls_b,
                                   const int* level_h, const int* level_w, const float* level_stride, int num_levels, int B,
                                   int c2, int c3, int nc, int reg_max, float conf_thres, uint64_t* cand_key,
                                   int32_t* cand_count, float* box_dense, int max_cand, cvpp_stream_t stream) {
  return yolov8_head_fused_launch(box_feat, cls_feat, box_w, box_b, cls_w, cls_b, level_h, level_w, level_stride, num_levels,
                                  B, c2, c3, nc, reg_max, conf_thres, cand_key, cand_count, box_dense, max_cand, nullptr,
                                  (cudaStream_t)stream);
}

int cvpp_yolov8_head_decode_filter_x(const float* const* box_feat, const float* const* cls_feat, const float* const* box_w,
                                     const float* const* box_b, const float* const* cls_w, const float* const* cls_b,
                                     const int* level_h, const int* level_w, const float* level_stride, int num_levels, int B,
                                     int c2, int c3, int nc, int reg_max, float conf_thres, uint64_t* cand_key,
                                     int32_t* cand_count, float* box_dense, int max_cand, float* head_out,
                                     cvpp_stream_t stream) {
  return yolov8_head_fused_launch(box_feat, cls_feat, box_w, box_b, cls_w, cls_b, level_h, level_w, level_stride, num_levels,
                                  B, c2, c3, nc, reg_max, conf_thres, cand_key, cand_count, box_dense, max_cand, head_out,
                                  (cudaStream_t)stream);
}

int cvpp_yolov8_decode_full(const float* const* level_ptr, const int64_t* batch_stride, const int64_t* chan_stride,
                            const int* level_h, const int* level_w, const float* level_stride, int num_levels, int B,
                            int nc, int reg_max, float* y, cvpp_stream_t stream) {
  if (!y) {
    set_error("yolov8_decode_full: y is NULL");
    return CVPP_ERR_INVALID_ARG;
  }
  return yolov8_decode_launch(level_ptr, batch_stride, chan_stride, level_h, level_w, level_stride, num_levels, B, nc,
                              reg_max, 0.0f, nullptr, nullptr, nullptr, 0, y, force_generic(), (cudaStream_t)stream);
}

int cvpp_pred_filter(const float* pred, int B, int channels, int nc, int64_t A, float conf_thres, uint64_t* cand_key,
                     int32_t* cand_count, float* box_dense, int max_cand, cvpp_stream_t stream) {
  if (!(conf_thres >= 0.0f && conf_thres <= 1.0f)) {
    set_error("Invalid Confidence threshold %f, valid values are between 0.0 and 1.0", conf_thres);
    return CVPP_ERR_INVALID_ARG;
  }
  return pred_filter_launch(pred, B, channels, nc, A, conf_thres, cand_key, cand_count, box_dense, max_cand,
                            (cudaStream_t)stream);
}

int cvpp_keep_classes(uint64_t* cand_key, int32_t* cand_count, int B, int max_cand, const int32_t* classes, int n_classes,
                      cvpp_stream_t stream) {
  return keep_classes_launch(cand_key, cand_count, B, max_cand, classes, n_classes, (cudaStream_t)stream);
}

size_t cvpp_sort_workspace_bytes(int B, int max_cand) {
  if (B < 0 || max_cand < 1) return 0;
  return segsort_workspace_bytes(B, max_cand);
}

int cvpp_segmented_sort(uint64_t* keys, int32_t* cand_count, int B, int max_cand, int rule, int max_nms,
                        void* workspace, size_t workspace_bytes, cvpp_stream_t stream) {
  return segsort_launch(keys, cand_count, B, max_cand, rule, max_nms, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t cvpp_nms_workspace_bytes(int B, int max_cand, int nc) {
  if (B < 0 || max_cand < 1 || nc < 1) return 0;
  return nms_workspace_bytes(B, max_cand, nc);
}

int cvpp_nms(const uint64_t* sorted_key, const int32_t* cand_count, const float* box_dense, int B, int max_cand,
             int64_t A, int nc, double iou_thres, int rule, int order, int max_det, int max_out, float* det_box,
             float* det_score, int32_t* det_cls, int32_t* det_anchor, int32_t* det_count, void* workspace,
             size_t workspace_bytes, cvpp_stream_t stream) {
  return nms_launch(sorted_key, cand_count, box_dense, B, max_cand, A, nc, iou_thres, rule, order, max_det, max_out,
                    det_box, det_score, det_cls, det_anchor, det_count, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}

size_t cvpp_sort_nms_workspace_bytes(int B, int max_cand, int nc) {
  if (B < 0 || max_cand < 1 || nc < 1) return 0;
  return sort_nms_workspace_bytes(B, max_cand, nc);
}

int cvpp_sort_nms(const uint64_t* cand_key, const int32_t* cand_count, const float* box_dense, int B, int max_cand, int64_t A, int nc,
                  double iou_thres, int rule, int order, int max_det, int max_nms, int max_out, float* det_box,
                  float* det_score, int32_t* det_cls, int32_t* det_anchor, int32_t* det_count, void* workspace,
                  size_t workspace_bytes, cvpp_stream_t stream) {
  return sort_nms_launch(cand_key, cand_count, box_dense, B, max_cand, A, nc, iou_thres, rule, order, max_det, max_nms,
                         max_out, det_box, det_score, det_cls, det_anchor, det_count, workspace, workspace_bytes,
                         (cudaStream_t)stream);
}

size_t cvpp_yolov8_workspace_bytes(int B, int64_t A, int max_cand, int nc) {
  if (B < 0 || A < 1 || max_cand < 1 || nc < 1) return 0;
  size_t s = 256;
  s += align256((size_t)B * max_cand * sizeof(uint64_t));  // cand_key
  s += align256((size_t)B * sizeof(int32_t));              // cand_count
  s += align256((size_t)B * (size_t)A * 16);               // box_dense
  s += align256(sort_nms_workspace_bytes(B, max_cand, nc));
  return s;
}

static int yolov8_postprocess_impl(const float* const* level_ptr, const int64_t* batch_stride, const int64_t* chan_stride,
                                   const int* level_h, const int* level_w, const float* level_stride, int num_levels, int B,
                                   int nc, int reg_max, float conf_thres, double iou_thres, int rule, int max_det, int max_nms,
                                   int max_cand, float* det_box, float* det_score, int32_t* det_cls, int32_t* det_anchor,
                                   int32_t* det_count, int32_t* cand_count_out, void* workspace, size_t workspace_bytes,
                                   cvpp_event_t inputs_consumed, cvpp_stream_t stream, float* const* gather_dst,
                                   float* gather_mc, int gather_n, int gather_rank) {
  if (!level_h || !level_w || num_levels < 1 || num_levels > CVPP_MAX_LEVELS) {
    set_error("yolov8_postprocess: bad level description");
    return CVPP_ERR_INVALID_ARG;
  }
  int64_t A = 0;
  for (int l = 0; l < num_levels; ++l) A += (int64_t)level_h[l] * level_w[l];
  if (max_det < 1 || max_cand < 1) {
    set_error("yolov8_postprocess: max_det and max_cand must be >= 1");
    return CVPP_ERR_INVALID_ARG;
  }
  if ((int64_t)max_cand < A) {
    set_error("yolov8_postprocess: max_cand=%d < A=%lld could overflow the candidate buffer (keys would be dropped in a "
              "nondeterministic order); pass max_cand >= A", max_cand, (long long)A);
    return CVPP_ERR_INVALID_ARG;
  }
  if (!workspace || workspace_bytes < cvpp_yolov8_workspace_bytes(B, A, max_cand, nc)) {
    set_error("yolov8_postprocess: workspace of %zu bytes needed, got %zu", cvpp_yolov8_workspace_bytes(B, A, max_cand, nc),
              workspace_bytes);
    return CVPP_ERR_WORKSPACE;
  }
  uintptr_t p = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  uint64_t* cand_key = reinterpret_cast<uint64_t*>(p);
  p += align256((size_t)B * max_cand * sizeof(uint64_t));
  int32_t* cand_count = reinterpret_cast<int32_t*>(p);
  p += align256((size_t)B * sizeof(int32_t));
  float* box_dense = reinterpret_cast<float*>(p);
  p += align256((size_t)B * (size_t)A * 16);
  void* sn_ws = reinterpret_cast<void*>(p);
  const size_t sn_bytes = sort_nms_workspace_bytes(B, max_cand, nc);

  int rc = cvpp_yolov8_decode_filter(level_ptr, batch_stride, chan_stride, level_h, level_w, level_stride, num_levels,
                                     B, nc, reg_max, conf_thres, cand_key, cand_count, box_dense, max_cand, stream);
  if (rc != CVPP_OK) return rc;
  if (inputs_consumed) {  // the head tensors have no reader after the decode kernel
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CVPP_CUDA_TRY(cudaStreamIsCapturing((cudaStream_t)stream, &cap));
    CVPP_CUDA_TRY(cudaEventRecordWithFlags((cudaEvent_t)inputs_consumed, (cudaStream_t)stream,
                                           cap == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault));
  }
  // (the fused kernel also copies the candidate counts out: no separate device-to-device copy node)
  return sort_nms_launch(cand_key, cand_count, box_dense, B, max_cand, A, nc, iou_thres, rule, CVPP_ORDER_SCORE_DESC,
                         max_det, max_nms, max_det, det_box, det_score, det_cls, det_anchor, det_count, sn_ws, sn_bytes,
                         (cudaStream_t)stream, cand_count_out, gather_dst, gather_mc, gather_n, gather_rank);
}

int cvpp_yolov8_postprocess_ev(const float* const* level_ptr, const int64_t* batch_stride, const int64_t* chan_stride,
                            const int* level_h, const int* level_w, const float* level_stride, int num_levels, int B,
                            int nc, int reg_max, float conf_thres, double iou_thres, int rule, int max_det, int max_nms,
                            int max_cand, float* det_box, float* det_score, int32_t* det_cls, int32_t* det_anchor,
                            int32_t* det_count, int32_t* cand_count_out, void* workspace, size_t workspace_bytes,
                            cvpp_event_t inputs_consumed, cvpp_stream_t stream) {
  return yolov8_postprocess_impl(level_ptr, batch_stride, chan_stride, level_h, level_w, level_stride, num_levels, B, nc,
                                 reg_max, conf_thres, iou_thres, rule, max_det, max_nms, max_cand, det_box, det_score, det_cls,
                                 det_anchor, det_count, cand_count_out, workspace, workspace_bytes, inputs_consumed, stream,
                                 nullptr, nullptr, 0, 0);
}

int cvpp_yolov8_postprocess_gather(const float* const* level_ptr, const int64_t* batch_stride, const int64_t* chan_stride,
                                   const int* level_h, const int* level_w, const float* level_stride, int num_levels, int B,
                                   int nc, int reg_max, float conf_thres, double iou_thres, int rule, int max_det, int max_nms,
                                   int max_cand, float* det_box, float* det_score, int32_t* det_cls, int32_t* det_anchor,
                                   int32_t* det_count, int32_t* cand_count_out, void* workspace, size_t workspace_bytes,
                                   cvpp_event_t inputs_consumed, float* const* peer_dst, float* mc_dst, int n_ranks, int rank,
                                   cvpp_stream_t stream) {
  if (n_ranks < 1 || (!peer_dst && !mc_dst)) {
    set_error("yolov8_postprocess_gather: n_ranks >= 1 and a peer list or a multicast address are needed");
    return CVPP_ERR_INVALID_ARG;
  }
  return yolov8_postprocess_impl(level_ptr, batch_stride, chan_stride, level_h, level_w, level_stride, num_levels, B, nc,
                                 reg_max, conf_thres, iou_thres, rule, max_det, max_nms, max_cand, det_box, det_score, det_cls,
                                 det_anchor, det_count, cand_count_out, workspace, workspace_bytes, inputs_consumed, stream,
                                 peer_dst, mc_dst, n_ranks, rank);
}

int cvpp_yolov8_postprocess(const float* const* level_ptr, const int64_t* batch_stride, const int64_t* chan_stride,
                            const int* level_h, const int* level_w, const float* level_stride, int num_levels, int B,
                            int nc, int reg_max, float conf_thres, double iou_thres, int rule, int max_det, int max_nms,
                            int max_cand, float* det_box, float* det_score, int32_t* det_cls, int32_t* det_anchor,
                            int32_t* det_count, int32_t* cand_count_out, void* workspace, size_t workspace_bytes,
                            cvpp_stream_t stream) {
  return cvpp_yolov8_postprocess_ev(level_ptr, batch_stride, chan_stride, level_h, level_w, level_stride, num_levels, B, nc,
                                    reg_max, conf_thres, iou_thres, rule, max_det, max_nms, max_cand, det_box, det_score,
                                    det_cls, det_anchor, det_count, cand_count_out, workspace, workspace_bytes, nullptr,
                                    stream);
}

size_t cvpp_centernet_workspace_bytes(int B, int H, int W, int nc, int K) {
  if (B < 0 || H < 1 || W < 1 || nc < 1 || K < 1) return 0;
  return centernet_workspace_bytes(B, H, W, nc, K);
}

int cvpp_centernet_decode(const float* pred, int B, int H, int W, int nc, int K, float conf_thres, int pool_mode,
                          int use_nms, float nms_thres, const float* letterbox, float* det_box, float* det_score,
                          int32_t* det_cls, int32_t* det_pixel, int32_t* det_count, void* workspace,
                          size_t workspace_bytes, cvpp_stream_t stream) {
  return centernet_launch(pred, B, H, W, nc, K, conf_thres, pool_mode, use_nms, nms_thres, letterbox, det_box,
                          det_score, det_cls, det_pixel, det_count, workspace, workspace_bytes, (cudaStream_t)stream);
}

int cvpp_diou_nms(const float* boxes, const float* scores, int n, float thr, int64_t* keep, int32_t* keep_count,
                  cvpp_stream_t stream) {
  return diou_nms_launch(boxes, scores, n, thr, reinterpret_cast<long long*>(keep), keep_count, (cudaStream_t)stream);
}

int cvpp_ssd_decode_filter(const float* loc, const float* conf, const float* priors, int B, int P, int nc,
                           float conf_thres, uint64_t* cand_key, int32_t* cand_count, float* box_dense, int max_cand,
                           cvpp_stream_t stream) {
  return ssd_decode_filter_launch(loc, conf, priors, B, P, nc, conf_thres, cand_key, cand_count, box_dense, max_cand,
                                  (cudaStream_t)stream);
}

int cvpp_ssd_parse_loc(const float* loc, const float* priors, int B, int P, float* out, cvpp_stream_t stream) {
  return ssd_parse_loc_launch(loc, priors, B, P, out, (cudaStream_t)stream);
}

int cvpp_yolov7_decode_filter(const float* const* level_ptr, const int64_t* batch_stride, const int64_t* chan_stride,
                              const int* level_h, const int* level_w, const float* level_anchors, int num_levels, int B,
                              int nc, int input_h, int input_w, float conf_thres, uint64_t* cand_key,
                              int32_t* cand_count, float* box_dense, float* aux_dense, int max_cand,
                              cvpp_stream_t stream) {
  return yolo_anchor_decode_launch(0, level_ptr, batch_stride, chan_stride, level_h, level_w, level_anchors, num_levels, B,
                                   nc, input_h, input_w, conf_thres, 0, cand_key, cand_count, box_dense, aux_dense,
                                   max_cand, force_generic(), (cudaStream_t)stream);
}

int cvpp_yolov7_pred_filter(const float* pred, int B, int64_t A, int nc, float conf_thres, uint64_t* cand_key,
                            int32_t* cand_count, float* box_dense, float* aux_dense, int max_cand,
                            cvpp_stream_t stream) {
  return yolov7_pred_filter_launch(pred, B, A, nc, conf_thres, cand_key, cand_count, box_dense, aux_dense, max_cand,
                                   (cudaStream_t)stream);
}

int cvpp_yolov3_decode_filter(const float* const* level_ptr, const int64_t* batch_stride, const int64_t* chan_stride,
                              const int* level_h, const int* level_w, const float* level_anchors, int num_levels, int B,
                              int nc, int input_h, int input_w, float conf_thres, int merge_batch, uint64_t* cand_key,
                              int32_t* cand_count, float* box_dense, int max_cand, cvpp_stream_t stream) {
  return yolo_anchor_decode_launch(1, level_ptr, batch_stride, chan_stride, level_h, level_w, level_anchors, num_levels, B,
                                   nc, input_h, input_w, conf_thres, merge_batch, cand_key, cand_count, box_dense, nullptr,
                                   max_cand, force_generic(), (cudaStream_t)stream);
}

int cvpp_yolov3_predict_bbox(const float* feature, int B, int nc, int H, int W, const float* anchors, float* box_xy,
                             float* box_wh, float* confidence, float* class_prob, cvpp_stream_t stream) {
  return yolov3_predict_bbox_launch(feature, B, nc, H, W, anchors, box_xy, box_wh, confidence, class_prob,
                                    (cudaStream_t)stream);
}

int cvpp_score_matrix_filter(const float* scores, int64_t M, int nc, float conf_thres, uint64_t* cand_key,
                             int32_t* cand_count, int max_cand, cvpp_stream_t stream) {
  return score_matrix_filter_launch(scores, M, nc, conf_thres, cand_key, cand_count, max_cand, (cudaStream_t)stream);
}

int cvpp_gather_feat(const float* feat, const void* ind, int ind_is_int64, const int32_t* count, int B, int64_t N, int C,
                     int K, float* out, int32_t* err_flag, cvpp_stream_t stream) {
  return gather_feat_launch(feat, ind, ind_is_int64, count, B, N, C, K, out, err_flag, (cudaStream_t)stream);
}

int cvpp_detection_epilogue(const float* det_box, const float* det_score, const int32_t* det_cls,
                            const int32_t* det_anchor, const int32_t* det_count, const float* aux_dense, int B,
                            int max_out, int64_t A, int layout, int box_mode, const float* letterbox, float* rows,
                            float* count_out, cvpp_stream_t stream) {
  return detection_epilogue_launch(det_box, det_score, det_cls, det_anchor, det_count, aux_dense, B, max_out, A, layout,
                                   box_mode, letterbox, rows, count_out, nullptr, nullptr, 0, 0, (cudaStream_t)stream);
}

int cvpp_detection_epilogue_compact(const float* det_box, const float* det_score, const int32_t* det_cls,
                                    const int32_t* det_anchor, const int32_t* det_count, const float* aux_dense, int B,
                                    int max_out, int64_t A, int layout, int box_mode, const float* letterbox, float* rows,
                                    int64_t row_capacity, int32_t* row_offset, int32_t* overflow, cvpp_stream_t stream) {
  return detection_epilogue_compact_launch(det_box, det_score, det_cls, det_anchor, det_count, aux_dense, B, max_out, A,
                                           layout, box_mode, letterbox, rows, row_capacity, row_offset, overflow,
                                           (cudaStream_t)stream);
}

int cvpp_detection_epilogue_allgather(const float* det_box, const float* det_score, const int32_t* det_cls,
                                      const int32_t* det_anchor, const int32_t* det_count, const float* aux_dense, int B,
                                      int max_out, int64_t A, int layout, int box_mode, const float* letterbox,
                                      float* const* peer_dst, int n_peers, int rank, cvpp_stream_t stream) {
  if (n_peers < 1) {
    set_error("detection_epilogue_allgather: n_peers must be >= 1");
    return CVPP_ERR_INVALID_ARG;
  }
  return detection_epilogue_launch(det_box, det_score, det_cls, det_anchor, det_count, aux_dense, B, max_out, A, layout,
                                   box_mode, letterbox, nullptr, nullptr, peer_dst, nullptr, n_peers, rank, (cudaStream_t)stream);
}

int cvpp_detection_epilogue_multicast(const float* det_box, const float* det_score, const int32_t* det_cls,
                                      const int32_t* det_anchor, const int32_t* det_count, const float* aux_dense, int B,
                                      int max_out, int64_t A, int layout, int box_mode, const float* letterbox,
                                      float* mc_dst, int n_ranks, int rank, cvpp_stream_t stream) {
  if (n_ranks < 1 || !mc_dst) {
    set_error("detection_epilogue_multicast: n_ranks must be >= 1 and mc_dst a multicast address");
    return CVPP_ERR_INVALID_ARG;
  }
  if (reinterpret_cast<uintptr_t>(mc_dst) & 15u) {
    set_error("detection_epilogue_multicast: mc_dst must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  return detection_epilogue_launch(det_box, det_score, det_cls, det_anchor, det_count, aux_dense, B, max_out, A, layout,
                                   box_mode, letterbox, nullptr, nullptr, nullptr, mc_dst, n_ranks, rank, (cudaStream_t)stream);
}

int cvpp_letterbox_reverse(const float* boxes, int64_t n, int xywh, float in_w, float in_h, float left, float top,
                           float scale, float* out, cvpp_stream_t stream) {
  return letterbox_reverse_launch(boxes, n, xywh, in_w, in_h, left, top, scale, out, (cudaStream_t)stream);
}

int cvpp_centernet_suppress(const float* heat, int B, int H, int W, int C, float* out, cvpp_stream_t stream) {
  return centernet_suppress_launch(heat, B, H, W, C, out, (cudaStream_t)stream);
}

int cvpp_voc_match(const float* det_rows, const int32_t* det_offset, const float* gt_box, const int32_t* gt_cls,
                   const int32_t* gt_difficult, const int32_t* gt_offset, int B, double min_overlap, int32_t* flag,
                   int32_t* best_gt, double* ovmax, int32_t* claim_ws, cvpp_stream_t stream) {
  return voc_match_launch(det_rows, det_offset, gt_box, gt_cls, gt_difficult, gt_offset, B, min_overlap, flag, best_gt, ovmax,
                          claim_ws, (cudaStream_t)stream);
}

size_t cvpp_topk_workspace_bytes(int B, int K) {
  if (B < 0 || K < 1) return 0;
  return topk_workspace_bytes(B, K);
}

int cvpp_topk(const float* scores, int B, int64_t N, int K, int C, int W, float* out_val, int64_t* out_idx, int64_t* out_cls,
              int64_t* out_y, int64_t* out_x, int32_t* out_pixel, void* workspace, size_t workspace_bytes,
              cvpp_stream_t stream) {
  return topk_launch(scores, B, (long long)N, K, C, W, out_val, reinterpret_cast<long long*>(out_idx),
                     reinterpret_cast<long long*>(out_cls), reinterpret_cast<long long*>(out_y),
                     reinterpret_cast<long long*>(out_x), out_pixel, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"

// rows_ops.cu — row-layout helpers around the NMS kernel (sm_100a):
//
//   yolov7_pred_filter  candidate stage of YOLOv7._nms / yolo7_nms on an already decoded (B, A, 5 + nc)
//                       tensor: core/algorithms/yolo_v7.py:361-377 == core/utils/nms.py:93-113
//                       (xywh_to_xyxy_torch core/utils/bboxes.py:29-49; class max first index; obj *
//                       class_conf >= conf).
//   gather_feat         RegL1Loss.gather_feat core/loss/centernet_loss.py:37-43 and the row gathers after
//                       NMS (gather_op core/utils/nms.py:34-51, `detections_class[keep]` yolo_v7.py:410).
//   detection_epilogue  what every decoder does with the kept rows before handing them to the caller:
//                       row assembly (ultralytics_ops.py:226,257; yolo_v7.py:391; ssd.py:275-278),
//                       YOLOv8's normalisation (yolo_v8.py:233-234), centre/size round trip
//                       (yolo_v8.py:237-238, yolo_v7.py:416-417, ssd.py:284-285) and yolo_correct_boxes /
//                       reverse_letter_box_numpy (core/utils/image_process.py:69-97,161-181), batched with
//                       one (h, w) per image instead of one D2H + numpy pass per image.
//
// All three are bandwidth-trivial next to the decode kernels (<= 85 floats per anchor streamed once, or a
// few KB of gathered rows); they exist so that no step of the path runs on the host or in eager ops.
#include "cvpp_common.cuh"

namespace cvpp {

// ---------------------------------------------------------------------------------------------------
// yolov7_pred_filter: rows are contiguous (5 + nc floats per anchor), so a CTA stages 128 rows with
// coalesced loads and each thread then scans its own row from shared memory (row pitch odd or padded to
// odd -> conflict-free).
// ---------------------------------------------------------------------------------------------------
constexpr int kRowThreads = 128;

__global__ void __launch_bounds__(kRowThreads)
yolov7_pred_filter_kernel(const float* __restrict__ pred, int64_t A, int nc, float conf_thres,
                          uint64_t* __restrict__ cand_key, int32_t* __restrict__ cand_count,
                          float4* __restrict__ box_dense, float2* __restrict__ aux_dense, int max_cand, int pitch) {
  extern __shared__ float sm[];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int attrs = 5 + nc;
  const int64_t a0 = (int64_t)blockIdx.x * kRowThreads;
  const int rows = (int)min((int64_t)kRowThreads, A - a0);
  const float* src = pred + ((int64_t)b * A + a0) * attrs;
  for (int i = tid; i < rows * attrs; i += kRowThreads) {
    const int r = i / attrs, c = i - r * attrs;
    sm[r * pitch + c] = __ldg(src + i);
  }
  __syncthreads();
  bool cand = false;
  float best = 0.0f, obj = 0.0f, score = 0.0f;
  int arg = 0;
  const float* x = sm + tid * pitch;
  if (tid < rows) {
    best = x[5];
    for (int k = 1; k < nc; ++k) {
      const float v = x[5 + k];
      if (v > best) {  // strict: first index wins ties, like torch.max(dim)
        best = v;
        arg = k;
      }
    }
    obj = x[4];
    score = fmul(obj, best);
    cand = score >= conf_thres;
  }
  const unsigned mask = __ballot_sync(0xffffffffu, cand);
  if (mask == 0) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(cand_count + b, __popc(mask));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (cand) {
    const int64_t a = a0 + tid;
    const float hw = fmul(x[2], 0.5f), hh = fmul(x[3], 0.5f);
    const int slot = base + __popc(mask & ((1u << lane) - 1u));
    if (slot < max_cand) cand_key[(int64_t)b * max_cand + slot] = key_pack((uint32_t)arg, __float_as_uint(score), (uint32_t)a);
    box_dense[(int64_t)b * A + a] = make_float4(fsub(x[0], hw), fsub(x[1], hh), fadd(x[0], hw), fadd(x[1], hh));
    aux_dense[(int64_t)b * A + a] = make_float2(obj, best);
  }
}

int yolov7_pred_filter_launch(const float* pred, int B, int64_t A, int nc, float conf_thres, uint64_t* cand_key,
                              int32_t* cand_count, float* box_dense, float* aux_dense, int max_cand,
                              cudaStream_t stream) {
  if (!pred || !cand_key || !cand_count || !box_dense || !aux_dense) {
    set_error("yolov7_pred_filter: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || nc < 1 || nc > CVPP_MAX_CLASSES || A < 1 || A > CVPP_MAX_ANCHORS || max_cand < 1) {
    set_error("yolov7_pred_filter: bad sizes (B=%d nc=%d A=%lld max_cand=%d)", B, nc, (long long)A, max_cand);
    return CVPP_ERR_INVALID_ARG;
  }
  if (!(conf_thres >= 0.0f && conf_thres <= 1.0f)) {
    set_error("yolov7_pred_filter: confidence threshold %f outside [0, 1]", conf_thres);
    return CVPP_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(box_dense) & 15u) || (reinterpret_cast<uintptr_t>(aux_dense) & 7u)) {
    set_error("yolov7_pred_filter: box_dense / aux_dense misaligned");
    return CVPP_ERR_ALIGNMENT;
  }
  CVPP_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(int32_t) * (size_t)B, stream));
  if (B == 0) return CVPP_OK;
  const int pitch = (5 + nc) | 1;
  const size_t smem = (size_t)kRowThreads * pitch * sizeof(float);
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != CVPP_OK) return rc;
  if (smem > (size_t)di.max_smem) {
    set_error("yolov7_pred_filter: nc=%d needs %zu bytes of staging shared memory", nc, smem);
    return CVPP_ERR_UNSUPPORTED;
  }
  static unsigned long long attr_done = 0;
  static int attr_bytes = 0;
  if ((int)smem > attr_bytes) {
    attr_done = 0;
    attr_bytes = (int)smem;
  }
  rc = ensure_smem_attr(reinterpret_cast<const void*>(yolov7_pred_filter_kernel), attr_bytes, di.device, &attr_done);
  if (rc != CVPP_OK) return rc;
  dim3 grid((unsigned)((A + kRowThreads - 1) / kRowThreads), (unsigned)B);
  yolov7_pred_filter_kernel<<<grid, kRowThreads, smem, stream>>>(pred, A, nc, conf_thres, cand_key, cand_count,
                                                                 reinterpret_cast<float4*>(box_dense),
                                                                 reinterpret_cast<float2*>(aux_dense), max_cand, pitch);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// score_matrix_filter: the mask of yolo3_nms (core/utils/nms.py:60) on a (M, nc) score matrix:
// one key (class c, score, row m) per element >= conf_thres.  Thread per element, coalesced.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
score_matrix_filter_kernel(const float* __restrict__ scores, int64_t total, int nc, float conf_thres,
                           uint64_t* __restrict__ cand_key, int32_t* __restrict__ cand_count, int max_cand) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  float s = 0.0f;
  bool hit = false;
  if (i < total) {
    s = __ldg(scores + i);
    hit = s >= conf_thres;
  }
  const unsigned mk = __ballot_sync(0xffffffffu, hit);
  if (mk == 0) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(cand_count, __popc(mk));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (hit) {
    const int slot = base + __popc(mk & ((1u << lane) - 1u));
    const int64_t m = i / nc;
    const int c = (int)(i - m * nc);
    if (slot < max_cand) cand_key[slot] = key_pack((uint32_t)c, __float_as_uint(s), (uint32_t)m);
  }
}

int score_matrix_filter_launch(const float* scores, int64_t M, int nc, float conf_thres, uint64_t* cand_key,
                               int32_t* cand_count, int max_cand, cudaStream_t stream) {
  if (!scores || !cand_key || !cand_count || M < 0 || M > CVPP_MAX_ANCHORS || nc < 1 || nc > CVPP_MAX_CLASSES ||
      max_cand < 1) {
    set_error("score_matrix_filter: NULL pointer or bad sizes (M=%lld nc=%d max_cand=%d)", (long long)M, nc, max_cand);
    return CVPP_ERR_INVALID_ARG;
  }
  CVPP_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(int32_t), stream));
  const int64_t total = M * nc;
  if (total == 0) return CVPP_OK;
  score_matrix_filter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(scores, total, nc, conf_thres, cand_key,
                                                                                 cand_count, max_cand);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// gather_feat: out[b, k, :] = feat[b, ind[b, k], :]   (rows of C floats; one warp per (b, k))
// ---------------------------------------------------------------------------------------------------
template <typename IndexT>
__global__ void __launch_bounds__(256)
gather_feat_kernel(const float* __restrict__ feat, const IndexT* __restrict__ ind, const int32_t* __restrict__ count,
                   int B, int64_t N, int C, int K, float* __restrict__ out, int* __restrict__ err) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)B * K) return;
  const int b = (int)(w / K), k = (int)(w - (int64_t)b * K);
  if (count && k >= count[b]) return;
  const int64_t i = (int64_t)ind[w];
  if (i < 0 || i >= N) {  // torch.gather raises; report instead of reading out of bounds
    if (lane == 0 && err) atomicExch(err, 1);
    return;
  }
  const float* src = feat + ((int64_t)b * N + i) * C;
  float* dst = out + w * C;
  for (int c = lane; c < C; c += 32) dst[c] = __ldg(src + c);
}

int gather_feat_launch(const float* feat, const void* ind, int ind_is_int64, const int32_t* count, int B, int64_t N, int C,
                       int K, float* out, int32_t* err_flag, cudaStream_t stream) {
  if (!feat || !ind || !out || B < 0 || N < 1 || C < 1 || K < 0) {
    set_error("gather_feat: NULL pointer or bad sizes (B=%d N=%lld C=%d K=%d)", B, (long long)N, C, K);
    return CVPP_ERR_INVALID_ARG;
  }
  if (err_flag) CVPP_CUDA_TRY(cudaMemsetAsync(err_flag, 0, sizeof(int32_t), stream));
  const int64_t warps = (int64_t)B * K;
  if (warps == 0) return CVPP_OK;
  const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
  if (ind_is_int64)
    gather_feat_kernel<long long><<<grid, 256, 0, stream>>>(feat, reinterpret_cast<const long long*>(ind), count, B, N, C, K,
                                                            out, err_flag);
  else
    gather_feat_kernel<int32_t><<<grid, 256, 0, stream>>>(feat, reinterpret_cast<const int32_t*>(ind), count, B, N, C, K, out,
                                                          err_flag);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// detection_epilogue: one thread per output row
// ---------------------------------------------------------------------------------------------------
struct EpilogueParams {
  const float4* det_box;
  const float* det_score;
  const int32_t* det_cls;
  const int32_t* det_anchor;
  const int32_t* det_count;
  const float2* aux_dense;
  const float* letterbox;  // (B, 5): in_w, in_h, left, top, scale
  float* rows;
  float* count_out;  // optional [B]: min(det_count, max_out) as fp32 (so rows + counts travel in one buffer)
  // peer scatter (fused all-gather over NVLink peer memory): when n_dst > 0 the row and the count are stored
  // into EVERY destination buffer dst[d] (peer-mapped device pointers, one per rank, own rank included) at
  // this rank's slot: rows at dst[d] + (slot * B + b) * max_out * width, counts at dst[d] + count_off + slot * B.
  float* dst[CVPP_MAX_PEERS];
  // NVSwitch multicast form: ONE multimem.st per 16 bytes to the multicast address of the same buffer; the switch
  // replicates it into every rank's copy, so a row leaves this GPU once instead of n_dst times.  n_dst still gives the layout.
  float* mc;
  int n_dst, slot;
  int64_t count_off;
  int B, max_out, layout, box_mode, width;
  int64_t A;
};

// one caller-facing row of detection t = b * max_out + k (k < det_count[b]) in the layout / box mode of p
__device__ __forceinline__ void epilogue_row(const EpilogueParams& p, int b, int64_t t, float (&row)[7]) {
  float4 bx = p.det_box[t];
  if (p.box_mode != CVPP_BOX_KEEP) {
    const float* L = p.letterbox + 5 * b;
    const float in_w = L[0], in_h = L[1], left = L[2], top = L[3], scale = L[4];
    if (p.box_mode == CVPP_BOX_NORMALISE_CORRECT) {  // yolo_v8.py:233-234
      bx.x = fdiv(bx.x, in_w);
      bx.z = fdiv(bx.z, in_w);
      bx.y = fdiv(bx.y, in_h);
      bx.w = fdiv(bx.w, in_h);
    }
    // (x1y1 + x2y2) / 2, x2y2 - x1y1 (yolo_v7.py:416-417), then c -/+ wh / 2 (image_process.py:80-83)
    const float cx = fmul(fadd(bx.x, bx.z), 0.5f), cy = fmul(fadd(bx.y, bx.w), 0.5f);
    const float hw = fmul(fsub(bx.z, bx.x), 0.5f), hh = fmul(fsub(bx.w, bx.y), 0.5f);
    // * input size, - padding, * scale (image_process.py:85-96); without letterbox the table holds
    // (image_w, image_h, 0, 0, 1) and the same sequence is image_process.py:178-181 bit for bit
    bx.x = fmul(fsub(fmul(fsub(cx, hw), in_w), left), scale);
    bx.z = fmul(fsub(fmul(fadd(cx, hw), in_w), left), scale);
    bx.y = fmul(fsub(fmul(fsub(cy, hh), in_h), top), scale);
    bx.w = fmul(fsub(fmul(fadd(cy, hh), in_h), top), scale);
  }
  row[0] = bx.x;
  row[1] = bx.y;
  row[2] = bx.z;
  row[3] = bx.w;
  const float cls = (float)p.det_cls[t];
  if (p.layout == CVPP_ROWS_YOLOV8) {  // x1,y1,x2,y2,conf,cls
    row[4] = p.det_score[t];
    row[5] = cls;
  } else if (p.layout == CVPP_ROWS_SSD) {  // x1,y1,x2,y2,label,conf
    row[4] = cls;
    row[5] = p.det_score[t];
  } else if (p.layout == CVPP_ROWS_YOLOV7) {  // x1,y1,x2,y2,obj,class_conf,class_pred
    const float2 a = p.aux_dense[(int64_t)b * p.A + p.det_anchor[t]];
    row[4] = a.x;
    row[5] = a.y;
    row[6] = cls;
  } else if (p.layout == CVPP_ROWS_COCO) {  // x,y,w,h,score,cls (yolo_v8.py:364-372: right - left, bottom - top in fp32)
    row[2] = fsub(bx.z, bx.x);
    row[3] = fsub(bx.w, bx.y);
    row[4] = p.det_score[t];
    row[5] = cls;
  } else if (p.layout == CVPP_ROWS_VOC) {  // cls,score,int(left),int(top),int(right),int(bottom) (yolo_v8.py:286-296)
    row[0] = cls;
    row[1] = p.det_score[t];
    row[2] = truncf(bx.x);
    row[3] = truncf(bx.y);
    row[4] = truncf(bx.z);
    row[5] = truncf(bx.w);
  } else {  // CVPP_ROWS_FULL: x1,y1,x2,y2,score,cls,anchor
    row[4] = p.det_score[t];
    row[5] = cls;
    row[6] = (float)p.det_anchor[t];
  }
}

__global__ void __launch_bounds__(256) detection_epilogue_kernel(const __grid_constant__ EpilogueParams p) {
  // The rows of a CTA are one contiguous run of the output: they are staged in shared memory and leave as
  // coalesced 128-bit stores - into each peer's buffer over NVLink in the gather form (per-element 4-byte
  // stores at a 28-byte stride made a million tiny NVLink writes per step and dragged the step at 8 GPUs).
  __shared__ __align__(16) float sh_rows[256 * 7];
  pdl_wait();  // programmatic dependent of the NMS kernel when it directly follows it in the stream
  const int64_t total = (int64_t)p.B * p.max_out;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x;
  const int64_t t = t0 + threadIdx.x;
  const bool valid = t < total;
  const int b = valid ? (int)(t / p.max_out) : 0, k = valid ? (int)(t - (int64_t)b * p.max_out) : 0;
  const int n = valid ? min(p.det_count[b], p.max_out) : 0;
  float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (valid && k < n) epilogue_row(p, b, t, row);
  for (int c = 0; c < p.width; ++c) sh_rows[threadIdx.x * p.width + c] = row[c];  // odd width: conflict-free
  __syncthreads();
  const int rows_here = (int)min((int64_t)blockDim.x, total - t0);
  const int floats = rows_here * p.width;
  if (p.mc) {
    float* o = p.mc + ((int64_t)p.slot * total + t0) * p.width;
    if ((reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
      const int nv = floats >> 2;
      const float4* s4 = reinterpret_cast<const float4*>(sh_rows);
      for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const float4 q = s4[i];
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * i), "f"(q.x), "f"(q.y), "f"(q.z),
                     "f"(q.w)
                     : "memory");
      }
      for (int i = (nv << 2) + threadIdx.x; i < floats; i += blockDim.x)
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(o + i), "f"(sh_rows[i]) : "memory");
    } else {
      for (int i = threadIdx.x; i < floats; i += blockDim.x)
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(o + i), "f"(sh_rows[i]) : "memory");
    }
    if (valid && k == 0)
      asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.mc + p.count_off + (int64_t)p.slot * p.B + b), "f"((float)n)
                   : "memory");
    return;
  }
  const int n_out = p.n_dst == 0 ? 1 : p.n_dst;
  for (int d = 0; d < n_out; ++d) {
    float* o = p.n_dst == 0 ? p.rows + t0 * p.width : p.dst[d] + ((int64_t)p.slot * total + t0) * p.width;
    if ((reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
      const int nv = floats >> 2;
      float4* o4 = reinterpret_cast<float4*>(o);
      const float4* s4 = reinterpret_cast<const float4*>(sh_rows);
      for (int i = threadIdx.x; i < nv; i += blockDim.x) o4[i] = s4[i];
      for (int i = (nv << 2) + threadIdx.x; i < floats; i += blockDim.x) o[i] = sh_rows[i];
    } else {
      for (int i = threadIdx.x; i < floats; i += blockDim.x) o[i] = sh_rows[i];
    }
    if (valid && k == 0) {
      if (p.n_dst == 0) {
        if (p.count_out) p.count_out[b] = (float)n;
      } else {
        p.dst[d][p.count_off + (int64_t)p.slot * p.B + b] = (float)n;
      }
    }
  }
}

int detection_epilogue_launch(const float* det_box, const float* det_score, const int32_t* det_cls,
                              const int32_t* det_anchor, const int32_t* det_count, const float* aux_dense, int B,
                              int max_out, int64_t A, int layout, int box_mode, const float* letterbox, float* rows,
                              float* count_out, float* const* peer_dst, float* mc_dst, int n_peers, int slot, cudaStream_t stream) {
  if (n_peers < 0 || n_peers > CVPP_MAX_PEERS || (n_peers > 0 && ((!peer_dst && !mc_dst) || slot < 0 || slot >= n_peers))) {
    set_error("detection_epilogue: bad peer list (n_peers=%d slot=%d)", n_peers, slot);
    return CVPP_ERR_INVALID_ARG;
  }
  if (!det_box || !det_score || !det_cls || !det_anchor || !det_count || (!rows && n_peers == 0)) {
    set_error("detection_epilogue: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || max_out < 1 || layout < CVPP_ROWS_YOLOV8 || layout > CVPP_ROWS_VOC || box_mode < CVPP_BOX_KEEP ||
      box_mode > CVPP_BOX_NORMALISE_CORRECT) {
    set_error("detection_epilogue: bad sizes / layout %d / box_mode %d", layout, box_mode);
    return CVPP_ERR_INVALID_ARG;
  }
  if (layout == CVPP_ROWS_YOLOV7 && (!aux_dense || A < 1)) {
    set_error("detection_epilogue: the YOLOv7 layout needs aux_dense (B, A, 2)");
    return CVPP_ERR_INVALID_ARG;
  }
  if (box_mode != CVPP_BOX_KEEP && !letterbox) {
    set_error("detection_epilogue: box_mode %d needs the (B, 5) letterbox table", box_mode);
    return CVPP_ERR_INVALID_ARG;
  }
  if (reinterpret_cast<uintptr_t>(det_box) & 15u) {
    set_error("detection_epilogue: det_box must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  if (B == 0) return CVPP_OK;
  EpilogueParams p;
  p.det_box = reinterpret_cast<const float4*>(det_box);
  p.det_score = det_score;
  p.det_cls = det_cls;
  p.det_anchor = det_anchor;
  p.det_count = det_count;
  p.aux_dense = reinterpret_cast<const float2*>(aux_dense);
  p.letterbox = letterbox;
  p.rows = rows;
  p.count_out = count_out;
  p.n_dst = n_peers;
  p.slot = slot;
  p.mc = mc_dst;
  for (int d = 0; d < n_peers && !mc_dst; ++d) {
    if (!peer_dst[d]) {
      set_error("detection_epilogue: peer destination %d is NULL", d);
      return CVPP_ERR_INVALID_ARG;
    }
    p.dst[d] = peer_dst[d];
  }
  p.B = B;
  p.max_out = max_out;
  p.layout = layout;
  p.box_mode = box_mode;
  p.width = (layout == CVPP_ROWS_YOLOV7 || layout == CVPP_ROWS_FULL) ? 7 : 6;
  p.count_off = (int64_t)n_peers * B * max_out * p.width;
  p.A = A;
  const int64_t total = (int64_t)B * max_out;
  CVPP_CUDA_TRY(launch_pdl(detection_epilogue_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, p));
  return CVPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Compact epilogue (the all-gather payload of heads WITHOUT a max_det cap - YOLOv7 / SSD / YOLOv3):
// the rows of image b start at row_offset[b] = sum over b' < b of min(det_count[b'], max_out), no padding.
// Kernel 1 (one CTA): exclusive scan of the clamped counts -> row_offset[0..B], overflow flag when the total
// exceeds the row capacity.  Kernel 2: grid (ceil(max_out / 256), B); a CTA past the image's count exits at once.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) row_offsets_kernel(const int32_t* __restrict__ det_count, int B, int max_out,
                                                           int64_t row_capacity, int32_t* __restrict__ row_offset,
                                                           int32_t* __restrict__ overflow) {
  __shared__ int sh_warp[32];
  __shared__ int sh_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) sh_base = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + tid;
    const int v = b < B ? min(max(det_count[b], 0), max_out) : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += u;
    }
    if (lane == 31) sh_warp[warp] = incl;
    __syncthreads();
    int woff = 0, total = 0;
    for (int q = 0; q < 32; ++q) {
      const int u = sh_warp[q];
      if (q < warp) woff += u;
      total += u;
    }
    const int base = sh_base;
    if (b < B) row_offset[b] = base + woff + incl - v;
    __syncthreads();
    if (tid == 0) sh_base = base + total;
    __syncthreads();
  }
  if (tid == 0) {
    row_offset[B] = sh_base;
    if (overflow) *overflow = (int64_t)sh_base > row_capacity ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256) detection_epilogue_compact_kernel(const __grid_constant__ EpilogueParams p,
                                                                         const int32_t* __restrict__ row_offset,
                                                                         int64_t row_capacity) {
  __shared__ __align__(16) float sh_rows[256 * 7];
  const int b = blockIdx.y;
  const int n = min(max(p.det_count[b], 0), p.max_out);
  const int k0 = blockIdx.x * 256;
  if (k0 >= n) return;
  const int k = k0 + threadIdx.x;
  float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (k < n) epilogue_row(p, b, (int64_t)b * p.max_out + k, row);
  for (int c = 0; c < p.width; ++c) sh_rows[threadIdx.x * p.width + c] = row[c];
  __syncthreads();
  const int64_t first = (int64_t)row_offset[b] + k0;
  int64_t rows_here = min(256, n - k0);
  if (first + rows_here > row_capacity) rows_here = row_capacity > first ? row_capacity - first : 0;  // overflow: truncated
  const int floats = (int)rows_here * p.width;
  float* o = p.rows + first * p.width;
  // the run starts at an arbitrary row: peel to the first 16-byte boundary, then 128-bit stores
  const int head = min(floats, (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(o) & 15u)) & 15u) >> 2));
  for (int i = threadIdx.x; i < head; i += 256) o[i] = sh_rows[i];
  const int nv = (floats - head) >> 2;
  float4* o4 = reinterpret_cast<float4*>(o + head);
  for (int i = threadIdx.x; i < nv; i += 256) {
    const float* s = sh_rows + head + 4 * i;
    o4[i] = make_float4(s[0], s[1], s[2], s[3]);
  }
  for (int i = head + (nv << 2) + threadIdx.x; i < floats; i += 256) o[i] = sh_rows[i];
}

int detection_epilogue_compact_launch(const float* det_box, const float* det_score, const int32_t* det_cls,
                                      const int32_t* det_anchor, const int32_t* det_count, const float* aux_dense, int B,
                                      int max_out, int64_t A, int layout, int box_mode, const float* letterbox, float* rows,
                                      int64_t row_capacity, int32_t* row_offset, int32_t* overflow, cudaStream_t stream) {
  if (!det_box || !det_score || !det_cls || !det_anchor || !det_count || !rows || !row_offset) {
    set_error("detection_epilogue_compact: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || B > 65535 || max_out < 1 || row_capacity < 0 || layout < CVPP_ROWS_YOLOV8 || layout > CVPP_ROWS_VOC ||
      box_mode < CVPP_BOX_KEEP || box_mode > CVPP_BOX_NORMALISE_CORRECT) {
    set_error("detection_epilogue_compact: bad sizes (B=%d, B <= 65535) / layout %d / box_mode %d", B, layout, box_mode);
    return CVPP_ERR_INVALID_ARG;
  }
  if ((int64_t)B * max_out > 0x7fffffffll) {
    set_error("detection_epilogue_compact: B * max_out must fit 31 bits");
    return CVPP_ERR_INVALID_ARG;
  }
  if (layout == CVPP_ROWS_YOLOV7 && (!aux_dense || A < 1)) {
    set_error("detection_epilogue_compact: the YOLOv7 layout needs aux_dense (B, A, 2)");
    return CVPP_ERR_INVALID_ARG;
  }
  if (box_mode != CVPP_BOX_KEEP && !letterbox) {
    set_error("detection_epilogue_compact: box_mode %d needs the (B, 5) letterbox table", box_mode);
    return CVPP_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(det_box) & 15u) || (reinterpret_cast<uintptr_t>(rows) & 3u)) {
    set_error("detection_epilogue_compact: det_box must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  row_offsets_kernel<<<1, 1024, 0, stream>>>(det_count, B, max_out, row_capacity, row_offset, overflow);
  CVPP_CUDA_TRY(cudaGetLastError());
  if (B == 0) return CVPP_OK;
  EpilogueParams p{};
  p.det_box = reinterpret_cast<const float4*>(det_box);
  p.det_score = det_score;
  p.det_cls = det_cls;
  p.det_anchor = det_anchor;
  p.det_count = det_count;
  p.aux_dense = reinterpret_cast<const float2*>(aux_dense);
  p.letterbox = letterbox;
  p.rows = rows;
  p.B = B;
  p.max_out = max_out;
  p.layout = layout;
  p.box_mode = box_mode;
  p.width = (layout == CVPP_ROWS_YOLOV7 || layout == CVPP_ROWS_FULL) ? 7 : 6;
  p.A = A;
  const dim3 grid((unsigned)((max_out + 255) / 256), (unsigned)B);
  detection_epilogue_compact_kernel<<<grid, 256, 0, stream>>>(p, row_offset, row_capacity);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// letterbox_reverse: reverse_letter_box (core/utils/image_process.py:100-129) on n boxes, one thread each.
// The scalars (in_w, in_h, left, top, scale) are computed by the caller in double like the reference and
// passed as fp32 (the reference multiplies fp32 tensors by Python floats, i.e. fp32 scalars).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
letterbox_reverse_kernel(const float4* __restrict__ boxes, int64_t n, int xywh, float in_w, float in_h, float left, float top,
                         float scale, float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 b = boxes[i];
  if (xywh) {  // (cx, cy, w, h) -> c -/+ wh / 2 (:112-113)
    const float hw = fmul(b.z, 0.5f), hh = fmul(b.w, 0.5f);
    b = make_float4(fsub(b.x, hw), fsub(b.y, hh), fadd(b.x, hw), fadd(b.y, hh));
  }
  b.x = fmul(fsub(fmul(b.x, in_w), left), scale);  // * input size, - padding, * scale (:116-128)
  b.z = fmul(fsub(fmul(b.z, in_w), left), scale);
  b.y = fmul(fsub(fmul(b.y, in_h), top), scale);
  b.w = fmul(fsub(fmul(b.w, in_h), top), scale);
  out[i] = b;
}

int letterbox_reverse_launch(const float* boxes, int64_t n, int xywh, float in_w, float in_h, float left, float top,
                             float scale, float* out, cudaStream_t stream) {
  if (n < 0 || (n > 0 && (!boxes || !out))) {
    set_error("letterbox_reverse: NULL pointer or negative n");
    return CVPP_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(boxes) | reinterpret_cast<uintptr_t>(out)) & 15u) {
    set_error("letterbox_reverse: boxes / out must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  if (n == 0) return CVPP_OK;
  letterbox_reverse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(boxes), n, xywh, in_w,
                                                                          in_h, left, top, scale, reinterpret_cast<float4*>(out));
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// centernet_suppress: CenterNetA._suppress_redundant_centers (core/algorithms/centernet.py:316-326), dense:
// out = heat * (heat == maxpool3x3(heat)) with the pool applied to the NHWC tensor as the reference does, i.e.
// over (x, channel) of every image row, -inf padding.  One thread per element.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
centernet_suppress_kernel(const float* __restrict__ heat, int64_t rows, int W, int C, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_row = (int64_t)W * C;
  if (i >= rows * per_row) return;
  const int64_t r = i / per_row;
  const int e = (int)(i - r * per_row);
  const int x = e / C, c = e - x * C;
  const float* q = heat + r * per_row;
  const float v = q[e];
  float m = v;
  for (int dx = -1; dx <= 1; ++dx) {
    const int xx = x + dx;
    if (xx < 0 || xx >= W) continue;
    for (int dc = -1; dc <= 1; ++dc) {
      const int cc = c + dc;
      if (cc < 0 || cc >= C) continue;
      m = fmaxf(m, __ldg(q + xx * C + cc));
    }
  }
  out[i] = v == m ? v : fmul(v, 0.0f);  // heatmap * keep.float(): v * 1 or v * 0 (keeps the sign of zero / NaN rules)
}

int centernet_suppress_launch(const float* heat, int B, int H, int W, int C, float* out, cudaStream_t stream) {
  if (B < 0 || H < 1 || W < 1 || C < 1 || !heat || !out) {
    set_error("centernet_suppress: NULL pointer or bad sizes");
    return CVPP_ERR_INVALID_ARG;
  }
  const int64_t total = (int64_t)B * H * W * C;
  if (total == 0) return CVPP_OK;
  centernet_suppress_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(heat, (int64_t)B * H, W, C, out);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

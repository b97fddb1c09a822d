/*
 * cvpp_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's detection post-processing path
 * (calmiLovesAI/ComputerVision.pytorch) in IEEE fp32, one rounding per
 * operation exactly as the reference's eager PyTorch ops perform them.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (computervision/pytorch_b200) never links or calls it.
 *
 * Pinning: the reference ships no tests or golden vectors, so
 * this file is pinned by the .npz fixtures under tests/golden/, produced by running the real
 * reference functions (+ the installed torchvision 0.26.0 CPU nms) through
 * tests/golden/make_golden.py in the build container.
 *
 * Third-party arithmetic restated here (not under /root/reference):
 *   torchvision 0.26.0 (reference README pins 0.14.1; CPU behaviour identical):
 *     torchvision/ops/boxes.py:51-120  batched_nms / _coordinate_trick / _vanilla
 *     torchvision/csrc/ops/cpu/nms_kernel.cpp  nms_kernel_impl<float>
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math, so
 * no FMA contraction and no reassociation).  Images are independent, so the
 * batch loops run on a small pthread pool (libgomp is not in this image).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_version(void) { return 1; }

static int g_threads = 0; /* 0 = all online cores */

ORC_API int orc_max_threads(void) {
  if (g_threads > 0) return g_threads;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

ORC_API void orc_set_threads(int n) { g_threads = n > 0 ? n : 0; }

/* parallel-for over [0, count): dynamic scheduling, one item at a time */
typedef void (*orc_body_fn)(int item, void* ctx);
typedef struct {
  orc_body_fn fn;
  void* ctx;
  int count;
  int next; /* atomically incremented */
} orc_pf;

static void* orc_pf_worker(void* arg) {
  orc_pf* pf = (orc_pf*)arg;
  for (;;) {
    int i = __atomic_fetch_add(&pf->next, 1, __ATOMIC_RELAXED);
    if (i >= pf->count) break;
    pf->fn(i, pf->ctx);
  }
  return NULL;
}

static void orc_parallel_for(int count, orc_body_fn fn, void* ctx) {
  int nt = orc_max_threads();
  if (nt > count) nt = count;
  orc_pf pf = {fn, ctx, count, 0};
  if (nt <= 1) {
    orc_pf_worker(&pf);
    return;
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
  int started = 0;
  for (int t = 0; t < nt - 1; ++t)
    if (pthread_create(&th[started], NULL, orc_pf_worker, &pf) == 0) ++started;
  orc_pf_worker(&pf);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
  free(th);
}

/* ------------------------------------------------------------------------- */
/* helpers                                                                   */
/* ------------------------------------------------------------------------- */

static inline float sigmoidf_ref(float x) {
  /* torch.sigmoid on fp32: 1 / (1 + exp(-x)) */
  return 1.0f / (1.0f + expf(-x));
}

typedef struct {
  float score;
  int64_t idx;
} orc_si;

/* descending score, ascending index: the total order of a *stable*
 * descending sort (torchvision nms_kernel.cpp sorts with stable=true). */
static int cmp_desc_stable(const void* a, const void* b) {
  const orc_si* x = (const orc_si*)a;
  const orc_si* y = (const orc_si*)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  if (x->idx < y->idx) return -1;
  if (x->idx > y->idx) return 1;
  return 0;
}

/* order[k] = index of the k-th element in stable descending score order */
static void argsort_desc_stable(const float* scores, int64_t n, int64_t* order) {
  orc_si* tmp = (orc_si*)malloc(sizeof(orc_si) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) {
    tmp[i].score = scores[i];
    tmp[i].idx = i;
  }
  qsort(tmp, (size_t)n, sizeof(orc_si), cmp_desc_stable);
  for (int64_t i = 0; i < n; ++i) order[i] = tmp[i].idx;
  free(tmp);
}

/* ------------------------------------------------------------------------- */
/* torchvision.ops.nms — torchvision/csrc/ops/cpu/nms_kernel.cpp             */
/*   called from: core/utils/ultralytics_ops.py:247 (via batched_nms),       */
/*   core/utils/nms.py:69,134, core/algorithms/yolo_v7.py:407,               */
/*   core/algorithms/ssd.py:267                                              */
/* boxes (n,4) xyxy fp32, scores (n) fp32, thr is a C double (python float). */
/* keep receives the kept indices in decreasing-score order; returns count.  */
/* ------------------------------------------------------------------------- */
ORC_API int64_t orc_nms(const float* boxes, const float* scores, int64_t n, double iou_threshold,
                        int64_t* keep) {
  if (n <= 0) return 0;
  float* areas = (float*)malloc(sizeof(float) * (size_t)n);
  int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  uint8_t* suppressed = (uint8_t*)calloc((size_t)n, 1);
  for (int64_t i = 0; i < n; ++i) {
    const float* b = boxes + 4 * i;
    areas[i] = (b[2] - b[0]) * (b[3] - b[1]);
  }
  argsort_desc_stable(scores, n, order);
  int64_t num_to_keep = 0;
  for (int64_t _i = 0; _i < n; ++_i) {
    int64_t i = order[_i];
    if (suppressed[i]) continue;
    keep[num_to_keep++] = i;
    float ix1 = boxes[4 * i + 0], iy1 = boxes[4 * i + 1];
    float ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    float iarea = areas[i];
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      int64_t j = order[_j];
      if (suppressed[j]) continue;
      float xx1 = fmaxf(ix1, boxes[4 * j + 0]);
      float yy1 = fmaxf(iy1, boxes[4 * j + 1]);
      float xx2 = fminf(ix2, boxes[4 * j + 2]);
      float yy2 = fminf(iy2, boxes[4 * j + 3]);
      float w = fmaxf(0.0f, xx2 - xx1);
      float h = fmaxf(0.0f, yy2 - yy1);
      float inter = w * h;
      float ovr = inter / (iarea + areas[j] - inter);
      /* float ovr is promoted to double for the compare */
      if ((double)ovr > iou_threshold) suppressed[j] = 1;
    }
  }
  free(areas);
  free(order);
  free(suppressed);
  return num_to_keep;
}

/* ------------------------------------------------------------------------- */
/* torchvision.ops.batched_nms — torchvision/ops/boxes.py:51-120             */
/* idxs are class ids stored as float (the reference passes x[:, 5]).        */
/* mode: 0 = torchvision's CPU rule (numel > 4000 -> vanilla else trick),    */
/*       1 = coordinate trick, 2 = vanilla, 3 = the switch torchvision uses  */
/*       for CUDA tensors (numel > 20000 -> vanilla), CPU kernel arithmetic. */
/* ------------------------------------------------------------------------- */
static int cmp_float_asc(const void* a, const void* b) {
  float x = *(const float*)a, y = *(const float*)b;
  return (x > y) - (x < y);
}

ORC_API int64_t orc_batched_nms(const float* boxes, const float* scores, const float* idxs, int64_t n,
                                double iou_threshold, int mode, int64_t* keep) {
  if (n <= 0) return 0;
  int vanilla = (mode == 2) || (mode == 0 && n * 4 > 4000) || (mode == 3 && n * 4 > 20000);
  if (!vanilla) {
    /* _batched_nms_coordinate_trick (boxes.py:83-100) */
    float maxc = boxes[0];
    for (int64_t i = 1; i < 4 * n; ++i)
      if (boxes[i] > maxc) maxc = boxes[i];
    float mult = maxc + 1.0f;
    float* shifted = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      float off = idxs[i] * mult;
      for (int k = 0; k < 4; ++k) shifted[4 * i + k] = boxes[4 * i + k] + off;
    }
    int64_t r = orc_nms(shifted, scores, n, iou_threshold, keep);
    free(shifted);
    return r;
  }
  /* _batched_nms_vanilla (boxes.py:103-120) */
  uint8_t* keep_mask = (uint8_t*)calloc((size_t)n, 1);
  float* uniq = (float*)malloc(sizeof(float) * (size_t)n);
  memcpy(uniq, idxs, sizeof(float) * (size_t)n);
  qsort(uniq, (size_t)n, sizeof(float), cmp_float_asc);
  int64_t nu = 0;
  for (int64_t i = 0; i < n; ++i)
    if (i == 0 || uniq[i] != uniq[nu - 1]) uniq[nu++] = uniq[i];
  float* cb = (float*)malloc(sizeof(float) * 4 * (size_t)n);
  float* cs = (float*)malloc(sizeof(float) * (size_t)n);
  int64_t* ci = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  int64_t* ck = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  for (int64_t u = 0; u < nu; ++u) {
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i)
      if (idxs[i] == uniq[u]) {
        memcpy(cb + 4 * m, boxes + 4 * i, sizeof(float) * 4);
        cs[m] = scores[i];
        ci[m] = i;
        ++m;
      }
    int64_t k = orc_nms(cb, cs, m, iou_threshold, ck);
    for (int64_t t = 0; t < k; ++t) keep_mask[ci[ck[t]]] = 1;
  }
  /* keep_indices[scores[keep_indices].sort(descending=True)[1]] */
  int64_t nk = 0;
  for (int64_t i = 0; i < n; ++i)
    if (keep_mask[i]) ci[nk++] = i;
  for (int64_t t = 0; t < nk; ++t) cs[t] = scores[ci[t]];
  argsort_desc_stable(cs, nk, ck);
  for (int64_t t = 0; t < nk; ++t) keep[t] = ci[ck[t]];
  free(keep_mask);
  free(uniq);
  free(cb);
  free(cs);
  free(ci);
  free(ck);
  return nk;
}

/* ------------------------------------------------------------------------- */
/* YOLOv8 head decode — core/models/yolov8/modules.py:434-445 (Detect tail), */
/* DFL modules.py:80-82, make_anchors core/utils/anchor.py:126-145,          */
/* dist2bbox core/utils/bboxes.py:213-222.                                   */
/* levels[l]: (B, 4*reg_max + nc, H_l, W_l) contiguous NCHW.                 */
/* y: (B, 4 + nc, A) with rows cx, cy, w, h (input pixels), sigmoid scores.  */
/* ------------------------------------------------------------------------- */
typedef struct {
  const float* const* levels;
  int num_levels;
  const int* level_h;
  const int* level_w;
  const float* level_stride;
  int nc, reg_max;
  int64_t A;
  float* y;
} orc_dec8_ctx;

static void orc_dec8_body(int b, void* vctx) {
  const orc_dec8_ctx* c = (const orc_dec8_ctx*)vctx;
  const int nc = c->nc, reg_max = c->reg_max;
  const int64_t A = c->A;
  const int no = 4 * reg_max + nc;
  int64_t a0 = 0;
  float* yb = c->y + (int64_t)b * (4 + nc) * A;
  for (int l = 0; l < c->num_levels; ++l) {
    const int H = c->level_h[l], W = c->level_w[l];
    const int64_t HW = (int64_t)H * W;
    const float* base = c->levels[l] + (int64_t)b * no * HW;
    const float s = c->level_stride[l];
    for (int64_t i = 0; i < HW; ++i) {
      float d[4];
      for (int side = 0; side < 4; ++side) {
        /* softmax over the reg_max bins (dim=1 after the transpose), then the
         * frozen 1x1 conv with weights arange(reg_max) */
        const float* p = base + (int64_t)side * reg_max * HW + i;
        float m = p[0];
        for (int k = 1; k < reg_max; ++k) {
          float v = p[(int64_t)k * HW];
          if (v > m) m = v;
        }
        float e[64];
        float sum = 0.0f;
        for (int k = 0; k < reg_max; ++k) {
          e[k] = expf(p[(int64_t)k * HW] - m);
          sum += e[k];
        }
        float acc = 0.0f;
        for (int k = 0; k < reg_max; ++k) acc += (float)k * (e[k] / sum);
        d[side] = acc;
      }
      float ax = (float)(i % W) + 0.5f;
      float ay = (float)(i / W) + 0.5f;
      float x1 = ax - d[0], y1 = ay - d[1];
      float x2 = ax + d[2], y2 = ay + d[3];
      float cx = (x1 + x2) / 2.0f, cy = (y1 + y2) / 2.0f;
      float w = x2 - x1, h = y2 - y1;
      int64_t a = a0 + i;
      yb[0 * A + a] = cx * s;
      yb[1 * A + a] = cy * s;
      yb[2 * A + a] = w * s;
      yb[3 * A + a] = h * s;
      const float* cl = base + (int64_t)4 * reg_max * HW + i;
      for (int k = 0; k < nc; ++k) yb[(int64_t)(4 + k) * A + a] = sigmoidf_ref(cl[(int64_t)k * HW]);
    }
    a0 += HW;
  }
}

ORC_API void orc_yolov8_decode(const float* const* levels, int num_levels, const int* level_h,
                               const int* level_w, const float* level_stride, int B, int nc, int reg_max,
                               float* y) {
  orc_dec8_ctx c = {levels, num_levels, level_h, level_w, level_stride, nc, reg_max, 0, y};
  if (reg_max > 64) return;
  for (int l = 0; l < num_levels; ++l) c.A += (int64_t)level_h[l] * level_w[l];
  orc_parallel_for(B, orc_dec8_body, &c);
}

/* ------------------------------------------------------------------------- */
/* Candidate filter shared by the two YOLOv8 entry points below:             */
/*   xc = prediction[:, 4:mi].amax(1) > conf_thres   (ultralytics_ops.py:190)*/
/*   box = xywh2xyxy(box)                            (:220, :360-375)        */
/*   conf, j = cls.max(1)  -> first index on ties    (:225)                  */
/*   keep rows with conf > conf_thres                (:226)                  */
/* Candidates come out in anchor order.  Returns the count.                  */
/* ------------------------------------------------------------------------- */
static int64_t yolov8_filter_image(const float* p, int nc, int64_t A, float conf_thres, float* boxes,
                                   float* conf, int32_t* cls, int32_t* anc) {
  int64_t n = 0;
  for (int64_t a = 0; a < A; ++a) {
    float best = p[(int64_t)4 * A + a];
    int bj = 0;
    for (int k = 1; k < nc; ++k) {
      float v = p[(int64_t)(4 + k) * A + a];
      if (v > best) {
        best = v;
        bj = k;
      }
    }
    if (!(best > conf_thres)) continue;
    float cx = p[0 * A + a], cy = p[1 * A + a], w = p[2 * A + a], h = p[3 * A + a];
    boxes[4 * n + 0] = cx - w / 2.0f;
    boxes[4 * n + 1] = cy - h / 2.0f;
    boxes[4 * n + 2] = cx + w / 2.0f;
    boxes[4 * n + 3] = cy + h / 2.0f;
    conf[n] = best;
    cls[n] = bj;
    anc[n] = (int32_t)a;
    ++n;
  }
  return n;
}

/* ------------------------------------------------------------------------- */
/* non_max_suppression — core/utils/ultralytics_ops.py:131-264 for the live  */
/* configuration (multi_label=False, labels=(), agnostic=False, merge=False, */
/* classes=None; wall-clock abort disabled).                                 */
/* pred: (B, 4 + nc + nm, A).  Outputs per image b (capacity max_det rows):  */
/*   det[b][k] = x1,y1,x2,y2,conf,cls  ; det_anchor[b][k] ; det_count[b]     */
/*   cand_count[b] = number of candidates that entered NMS (may be NULL).    */
/* nms_mode as in orc_batched_nms.                                           */
/* ------------------------------------------------------------------------- */
typedef struct {
  const float* pred;
  int nc, nm;
  int64_t A;
  float conf_thres;
  double iou_thres;
  int max_det, max_nms, nms_mode;
  float* det;
  int32_t* det_anchor;
  int32_t* det_count;
  int32_t* cand_count;
} orc_nms8_ctx;

static void orc_nms8_body(int b, void* vctx) {
  const orc_nms8_ctx* c = (const orc_nms8_ctx*)vctx;
  const int64_t A = c->A;
  const int ch = 4 + c->nc + c->nm;
  const float* p = c->pred + (int64_t)b * ch * A;
  float* boxes = (float*)malloc(sizeof(float) * 4 * (size_t)A);
  float* conf = (float*)malloc(sizeof(float) * (size_t)A);
  float* clsf = (float*)malloc(sizeof(float) * (size_t)A);
  int32_t* cls = (int32_t*)malloc(sizeof(int32_t) * (size_t)A);
  int32_t* anc = (int32_t*)malloc(sizeof(int32_t) * (size_t)A);
  int64_t n = yolov8_filter_image(p, c->nc, A, c->conf_thres, boxes, conf, cls, anc);
  c->det_count[b] = 0;
  if (c->cand_count) c->cand_count[b] = (int32_t)n;
  if (n > 0) {
    /* x = x[x[:, 4].argsort(descending=True)[:max_nms]] (:240) */
    int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    argsort_desc_stable(conf, n, order);
    int64_t m = n < c->max_nms ? n : c->max_nms;
    float* sb = (float*)malloc(sizeof(float) * 4 * (size_t)m);
    float* ss = (float*)malloc(sizeof(float) * (size_t)m);
    int32_t* sa = (int32_t*)malloc(sizeof(int32_t) * (size_t)m);
    for (int64_t t = 0; t < m; ++t) {
      memcpy(sb + 4 * t, boxes + 4 * order[t], sizeof(float) * 4);
      ss[t] = conf[order[t]];
      clsf[t] = (float)cls[order[t]]; /* j.float() (:226) */
      sa[t] = anc[order[t]];
    }
    int64_t* keep = (int64_t*)malloc(sizeof(int64_t) * (size_t)m);
    int64_t k = orc_batched_nms(sb, ss, clsf, m, c->iou_thres, c->nms_mode, keep); /* (:247) */
    if (k > c->max_det) k = c->max_det;                                            /* (:248) */
    for (int64_t t = 0; t < k; ++t) {
      float* row = c->det + ((int64_t)b * c->max_det + t) * 6;
      memcpy(row, sb + 4 * keep[t], sizeof(float) * 4);
      row[4] = ss[keep[t]];
      row[5] = clsf[keep[t]];
      c->det_anchor[(int64_t)b * c->max_det + t] = sa[keep[t]];
    }
    c->det_count[b] = (int32_t)k;
    free(order);
    free(sb);
    free(ss);
    free(sa);
    free(keep);
  }
  free(boxes);
  free(conf);
  free(clsf);
  free(cls);
  free(anc);
}

ORC_API void orc_yolov8_nms(const float* pred, int B, int nc, int nm, int64_t A, float conf_thres,
                            double iou_thres, int max_det, int max_nms, int nms_mode, float* det,
                            int32_t* det_anchor, int32_t* det_count, int32_t* cand_count) {
  orc_nms8_ctx c = {pred, nc, nm, A, conf_thres, iou_thres, max_det, max_nms, nms_mode,
                    det,  det_anchor, det_count, cand_count};
  orc_parallel_for(B, orc_nms8_body, &c);
}

/* Candidate stage only (what the fused CUDA decode+filter kernel emits), in  */
/* anchor order.  cand_box (B,A,4) xyxy ; cand_score/cls/anchor (B,A) ;       */
/* cand_count (B).                                                            */
typedef struct {
  const float* pred;
  int nc, nm;
  int64_t A;
  float conf_thres;
  float* cand_box;
  float* cand_score;
  int32_t* cand_cls;
  int32_t* cand_anchor;
  int32_t* cand_count;
} orc_cand8_ctx;

static void orc_cand8_body(int b, void* vctx) {
  const orc_cand8_ctx* c = (const orc_cand8_ctx*)vctx;
  const int64_t A = c->A;
  const float* p = c->pred + (int64_t)b * (4 + c->nc + c->nm) * A;
  c->cand_count[b] = (int32_t)yolov8_filter_image(p, c->nc, A, c->conf_thres, c->cand_box + (int64_t)b * A * 4,
                                                  c->cand_score + (int64_t)b * A, c->cand_cls + (int64_t)b * A,
                                                  c->cand_anchor + (int64_t)b * A);
}

ORC_API void orc_yolov8_candidates(const float* pred, int B, int nc, int nm, int64_t A, float conf_thres,
                                   float* cand_box, float* cand_score, int32_t* cand_cls,
                                   int32_t* cand_anchor, int32_t* cand_count) {
  orc_cand8_ctx c = {pred, nc, nm, A, conf_thres, cand_box, cand_score, cand_cls, cand_anchor, cand_count};
  orc_parallel_for(B, orc_cand8_body, &c);
}

/* ------------------------------------------------------------------------- */
/* Per-class NMS over one image's candidate list, the shape shared by        */
/* YOLOv7._nms (core/algorithms/yolo_v7.py:396-413), Ssd.decode_boxes        */
/* (core/algorithms/ssd.py:256-278) and yolo3_nms (core/utils/nms.py:66-76): */
/* for each class id ascending, torchvision nms on that class's boxes (in    */
/* candidate order), results concatenated class-major, score-descending.     */
/* keep receives indices into the candidate list; returns count.             */
/* ------------------------------------------------------------------------- */
ORC_API int64_t orc_nms_per_class(const float* boxes, const float* scores, const int32_t* cls, int64_t n,
                                  int nc, double iou_threshold, int64_t* keep) {
  if (n <= 0) return 0;
  float* cb = (float*)malloc(sizeof(float) * 4 * (size_t)n);
  float* cs = (float*)malloc(sizeof(float) * (size_t)n);
  int64_t* ci = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  int64_t* ck = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  int64_t total = 0;
  for (int c = 0; c < nc; ++c) {
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i)
      if (cls[i] == c) {
        memcpy(cb + 4 * m, boxes + 4 * i, sizeof(float) * 4);
        cs[m] = scores[i];
        ci[m] = i;
        ++m;
      }
    if (m == 0) continue;
    int64_t k = orc_nms(cb, cs, m, iou_threshold, ck);
    for (int64_t t = 0; t < k; ++t) keep[total++] = ci[ck[t]];
  }
  free(cb);
  free(cs);
  free(ci);
  free(ck);
  return total;
}

/* ========================================================================= */
/* CenterNet — core/algorithms/centernet.py:271-338, core/utils/nms.py:9-31, */
/* core/utils/iou.py:8-64, core/loss/centernet_loss.py:37-43,                */
/* core/utils/image_process.py:100-129                                       */
/* ========================================================================= */

/* box_diou(boxes1, boxes2) (iou.py:41-64) on two xyxy boxes, op for op in fp32 */
static float diou_ref(const float* a, const float* b) {
  const float eps = 1e-6f;
  float area1 = (a[2] - a[0]) * (a[3] - a[1]);
  float area2 = (b[2] - b[0]) * (b[3] - b[1]);
  float iw = fminf(a[2], b[2]) - fmaxf(a[0], b[0]);
  float ih = fminf(a[3], b[3]) - fmaxf(a[1], b[1]);
  if (iw < 0.0f) iw = 0.0f; /* clamp(min=0) */
  if (ih < 0.0f) ih = 0.0f;
  float inter = iw * ih;
  float uni = area1 + area2 - inter;
  float iou = inter / (uni < eps ? eps : uni);
  float c1x = (a[0] + a[2]) / 2.0f, c1y = (a[1] + a[3]) / 2.0f;
  float c2x = (b[0] + b[2]) / 2.0f, c2y = (b[1] + b[3]) / 2.0f;
  float ew = fmaxf(a[2], b[2]) - fminf(a[0], b[0]);
  float eh = fmaxf(a[3], b[3]) - fminf(a[1], b[1]);
  if (ew < 0.0f) ew = 0.0f;
  if (eh < 0.0f) eh = 0.0f;
  float c_sq = ew * ew + eh * eh;
  float dx = c1x - c2x, dy = c1y - c2y;
  float d_sq = dx * dx + dy * dy;
  return iou - d_sq / (c_sq < eps ? eps : c_sq);
}

/* diou_nms (nms.py:9-31): greedy in descending score order (ties: lower index first), a later box */
/* survives a kept one iff diou <= thr (fp32 compare).  Returns the kept indices in that order.   */
ORC_API int64_t orc_diou_nms(const float* boxes, const float* scores, int64_t n, float thr, int64_t* keep) {
  if (n <= 0) return 0;
  int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  uint8_t* dead = (uint8_t*)calloc((size_t)n, 1);
  argsort_desc_stable(scores, n, order);
  int64_t k = 0;
  for (int64_t _i = 0; _i < n; ++_i) {
    int64_t i = order[_i];
    if (dead[i]) continue;
    keep[k++] = i;
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      int64_t j = order[_j];
      if (dead[j]) continue;
      if (!(diou_ref(boxes + 4 * i, boxes + 4 * j) <= thr)) dead[j] = 1;
    }
  }
  free(order);
  free(dead);
  return k;
}

typedef struct {
  float score;
  int64_t idx;
} orc_peak;

static int cmp_peak(const void* a, const void* b) {
  const orc_peak* x = (const orc_peak*)a;
  const orc_peak* y = (const orc_peak*)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* One image of CenterNetA.decode_boxes up to (and including) the score mask (:274-304).             */
/* pred: (H, W, nc + 4) NHWC.  pool_mode 0 = the reference's behaviour (MaxPool2d applied to the      */
/* NHWC tensor, i.e. a 3x3 window over (x, class)), 1 = spatial 3x3 over (y, x).      */
/* Outputs (capacity K): box xyxy normalised+clamped, score, cls, pixel index y*W+x.  Returns count.  */
static int centernet_image(const float* pred, int H, int W, int nc, int K, float conf, int pool_mode, float* box,
                           float* score, int32_t* cls, int32_t* pix) {
  const int Cf = nc + 4;
  const int64_t total = (int64_t)H * W * nc;
  float* s = (float*)malloc(sizeof(float) * (size_t)total);
  for (int64_t p = 0; p < (int64_t)H * W; ++p)
    for (int c = 0; c < nc; ++c) s[p * nc + c] = sigmoidf_ref(pred[p * Cf + c]);
  orc_peak* peaks = (orc_peak*)malloc(sizeof(orc_peak) * (size_t)total);
  int64_t np_ = 0;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < nc; ++c) {
        const int64_t idx = ((int64_t)y * W + x) * nc + c;
        const float v = s[idx];
        float m = v;
        for (int d0 = -1; d0 <= 1; ++d0)
          for (int d1 = -1; d1 <= 1; ++d1) {
            int yy = y, xx = x, cc = c;
            if (pool_mode == 0) {
              xx = x + d0;
              cc = c + d1;
            } else {
              yy = y + d0;
              xx = x + d1;
            }
            if (yy < 0 || yy >= H || xx < 0 || xx >= W || cc < 0 || cc >= nc) continue; /* -inf padding */
            float t = s[((int64_t)yy * W + xx) * nc + cc];
            if (t > m) m = t;
          }
        /* heatmap * (heatmap == hmax): non-peaks become 0 and still take part in topk */
        peaks[np_].score = (v == m) ? v : 0.0f;
        peaks[np_].idx = idx;
        ++np_;
      }
  /* topk(k, largest, sorted): descending score, ties resolved by lower flat index (our stated rule) */
  qsort(peaks, (size_t)np_, sizeof(orc_peak), cmp_peak);
  int n = 0;
  for (int k = 0; k < K && k < np_; ++k) {
    const int64_t idx = peaks[k].idx;
    const int c = (int)(idx % nc);
    const int64_t pixel = idx / nc;
    const int y = (int)(pixel / W), x = (int)(pixel % W);
    const float* pp = pred + pixel * Cf;
    float cx = (float)x + pp[nc], cy = (float)y + pp[nc + 1]; /* xs + reg[...,0], ys + reg[...,1] */
    float w = pp[Cf - 2], h = pp[Cf - 1];                     /* wh = pred[..., -2:] */
    cx = cx / (float)W;
    w = w / (float)W;
    cy = cy / (float)H;
    h = h / (float)H;
    cx = fminf(fmaxf(cx, 0.0f), 1.0f);
    cy = fminf(fmaxf(cy, 0.0f), 1.0f);
    w = fminf(fmaxf(w, 0.0f), 1.0f);
    h = fminf(fmaxf(h, 0.0f), 1.0f);
    if (!(peaks[k].score >= conf)) continue; /* score mask (:300) */
    box[4 * n + 0] = cx - w / 2.0f;
    box[4 * n + 1] = cy - h / 2.0f;
    box[4 * n + 2] = cx + w / 2.0f;
    box[4 * n + 3] = cy + h / 2.0f;
    score[n] = peaks[k].score;
    cls[n] = c;
    pix[n] = (int32_t)pixel;
    ++n;
  }
  free(s);
  free(peaks);
  return n;
}

typedef struct {
  const float* pred;
  int H, W, nc, K, pool_mode, use_nms;
  float conf, nms_thr;
  const float* letterbox; /* per image: in_w, in_h, left, top, scale (already fp32) or NULL */
  float* box;
  float* score;
  int32_t* cls;
  int32_t* pix;
  int32_t* count;
} orc_cn_ctx;

static void orc_cn_body(int b, void* vctx) {
  const orc_cn_ctx* c = (const orc_cn_ctx*)vctx;
  const int K = c->K;
  float* box = c->box + (int64_t)b * K * 4;
  float* score = c->score + (int64_t)b * K;
  int32_t* cls = c->cls + (int64_t)b * K;
  int32_t* pix = c->pix + (int64_t)b * K;
  int n = centernet_image(c->pred + (int64_t)b * c->H * c->W * (c->nc + 4), c->H, c->W, c->nc, K, c->conf,
                          c->pool_mode, box, score, cls, pix);
  if (c->use_nms && n > 0) {
    int64_t* keep = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    int64_t k = orc_diou_nms(box, score, n, c->nms_thr, keep);
    float* tb = (float*)malloc(sizeof(float) * 4 * (size_t)k);
    float* ts = (float*)malloc(sizeof(float) * (size_t)k);
    int32_t* tc = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
    int32_t* tp = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
    for (int64_t t = 0; t < k; ++t) {
      memcpy(tb + 4 * t, box + 4 * keep[t], sizeof(float) * 4);
      ts[t] = score[keep[t]];
      tc[t] = cls[keep[t]];
      tp[t] = pix[keep[t]];
    }
    memcpy(box, tb, sizeof(float) * 4 * (size_t)k);
    memcpy(score, ts, sizeof(float) * (size_t)k);
    memcpy(cls, tc, sizeof(int32_t) * (size_t)k);
    memcpy(pix, tp, sizeof(int32_t) * (size_t)k);
    free(keep);
    free(tb);
    free(ts);
    free(tc);
    free(tp);
    n = (int)k;
  }
  if (c->letterbox) { /* reverse_letter_box(xywh=False) (image_process.py:112-129) */
    const float* L = c->letterbox + 5 * b;
    for (int i = 0; i < n; ++i) {
      float* q = box + 4 * i;
      q[0] = (q[0] * L[0] - L[2]) * L[4];
      q[2] = (q[2] * L[0] - L[2]) * L[4];
      q[1] = (q[1] * L[1] - L[3]) * L[4];
      q[3] = (q[3] * L[1] - L[3]) * L[4];
    }
  }
  c->count[b] = n;
}

/* Per-image CenterNet decode (the reference merges the images of a batch before its NMS; callers     */
/* that need that behaviour run B = 1 or merge themselves).                                           */
ORC_API void orc_centernet_decode(const float* pred, int B, int H, int W, int nc, int K, float conf, int pool_mode,
                                  int use_nms, float nms_thr, const float* letterbox, float* box, float* score,
                                  int32_t* cls, int32_t* pix, int32_t* count) {
  orc_cn_ctx c = {pred, H, W, nc, K, pool_mode, use_nms, conf, nms_thr, letterbox, box, score, cls, pix, count};
  orc_parallel_for(B, orc_cn_body, &c);
}

/* ========================================================================= */
/* SSD — core/algorithms/ssd.py:236-288 (decode_boxes), :290-325             */
/* (_parse_mbox_loc).  Priors come from _get_ssd_anchors (:482-541), computed */
/* in float64 numpy by the caller and passed in as float32.                   */
/* ========================================================================= */
static void ssd_parse_loc(const float* loc, const float* priors, int64_t P, float* box) {
  const float v0 = 0.1f, v1 = 0.2f; /* variance[::2] */
  for (int64_t i = 0; i < P; ++i) {
    const float* a = priors + 4 * i;
    const float* l = loc + 4 * i;
    float aw = a[2] - a[0], ah = a[3] - a[1];
    float acx = 0.5f * (a[2] + a[0]), acy = 0.5f * (a[3] + a[1]);
    float cx = l[0] * aw * v0;
    cx += acx;
    float cy = l[1] * ah * v0;
    cy += acy;
    float w = expf(l[2] * v1);
    w *= aw;
    float h = expf(l[3] * v1);
    h *= ah;
    float x1 = cx - 0.5f * w, y1 = cy - 0.5f * h, x2 = cx + 0.5f * w, y2 = cy + 0.5f * h;
    box[4 * i + 0] = fminf(fmaxf(x1, 0.0f), 1.0f);
    box[4 * i + 1] = fminf(fmaxf(y1, 0.0f), 1.0f);
    box[4 * i + 2] = fminf(fmaxf(x2, 0.0f), 1.0f);
    box[4 * i + 3] = fminf(fmaxf(y2, 0.0f), 1.0f);
  }
}

ORC_API void orc_ssd_parse_mbox_loc(const float* loc, const float* priors, int64_t P, float* box) {
  ssd_parse_loc(loc, priors, P, box);
}

typedef struct {
  const float* loc;
  const float* conf;
  const float* priors;
  int64_t P;
  int nc; /* foreground classes; conf has nc + 1 columns, column 0 = background */
  float conf_thres;
  double nms_thres;
  int cap; /* rows per image in the outputs */
  float* rows;        /* (B, cap, 6): x1,y1,x2,y2,label,conf (normalised coordinates) */
  int32_t* row_prior; /* (B, cap) */
  int32_t* count;     /* (B) true number of rows (may exceed cap) */
} orc_ssd_ctx;

static void orc_ssd_body(int b, void* vctx) {
  const orc_ssd_ctx* c = (const orc_ssd_ctx*)vctx;
  const int64_t P = c->P;
  const int nc1 = c->nc + 1;
  float* box = (float*)malloc(sizeof(float) * 4 * (size_t)P);
  float* prob = (float*)malloc(sizeof(float) * (size_t)P * nc1);
  ssd_parse_loc(c->loc + (int64_t)b * P * 4, c->priors, P, box);
  const float* logits = c->conf + (int64_t)b * P * nc1;
  for (int64_t i = 0; i < P; ++i) { /* torch.softmax(preds[1], dim=-1) (:248) */
    const float* x = logits + i * nc1;
    float m = x[0];
    for (int k = 1; k < nc1; ++k)
      if (x[k] > m) m = x[k];
    float sum = 0.0f;
    for (int k = 0; k < nc1; ++k) {
      prob[i * nc1 + k] = expf(x[k] - m);
      sum += prob[i * nc1 + k];
    }
    for (int k = 0; k < nc1; ++k) prob[i * nc1 + k] = prob[i * nc1 + k] / sum;
  }
  float* cb = (float*)malloc(sizeof(float) * 4 * (size_t)P);
  float* cs = (float*)malloc(sizeof(float) * (size_t)P);
  int64_t* ci = (int64_t*)malloc(sizeof(int64_t) * (size_t)P);
  int64_t* keep = (int64_t*)malloc(sizeof(int64_t) * (size_t)P);
  int64_t total = 0;
  for (int cls = 1; cls <= c->nc; ++cls) { /* (:256-278) */
    int64_t m = 0;
    for (int64_t i = 0; i < P; ++i) {
      float p = prob[i * nc1 + cls];
      if (p > c->conf_thres) {
        memcpy(cb + 4 * m, box + 4 * i, sizeof(float) * 4);
        cs[m] = p;
        ci[m] = i;
        ++m;
      }
    }
    if (m == 0) continue;
    int64_t k = orc_nms(cb, cs, m, c->nms_thres, keep);
    for (int64_t t = 0; t < k; ++t, ++total) {
      if (total >= c->cap) continue;
      float* row = c->rows + ((int64_t)b * c->cap + total) * 6;
      memcpy(row, cb + 4 * keep[t], sizeof(float) * 4);
      row[4] = (float)(cls - 1);
      row[5] = cs[keep[t]];
      c->row_prior[(int64_t)b * c->cap + total] = (int32_t)ci[keep[t]];
    }
  }
  c->count[b] = (int32_t)total;
  free(box);
  free(prob);
  free(cb);
  free(cs);
  free(ci);
  free(keep);
}

ORC_API void orc_ssd_decode(const float* loc, const float* conf, const float* priors, int B, int64_t P, int nc,
                            float conf_thres, double nms_thres, int cap, float* rows, int32_t* row_prior,
                            int32_t* count) {
  orc_ssd_ctx c = {loc, conf, priors, P, nc, conf_thres, nms_thres, cap, rows, row_prior, count};
  orc_parallel_for(B, orc_ssd_body, &c);
}

/* ========================================================================= */
/* YOLOv7 — core/algorithms/yolo_v7.py:234-346 (decode_box), :348-422 (_nms)  */
/* ========================================================================= */
typedef struct {
  const float* const* levels; /* level i: (B, 3*(5+nc), H_i, W_i) NCHW */
  int num_levels;
  const int* level_h;
  const int* level_w;
  const float* anchors;  /* (num_levels*3, 2) pixels, already ordered by anchors_mask: row 3*i + a */
  int input_h, input_w;
  int nc;
  int64_t A;
  float* out; /* (B, A, 5 + nc): cx, cy, w, h (normalised), obj, cls... */
} orc_v7dec_ctx;

static void orc_v7dec_body(int b, void* vctx) {
  const orc_v7dec_ctx* c = (const orc_v7dec_ctx*)vctx;
  const int nc = c->nc, attrs = 5 + nc;
  float* ob = c->out + (int64_t)b * c->A * attrs;
  int64_t off = 0;
  for (int l = 0; l < c->num_levels; ++l) {
    const int H = c->level_h[l], W = c->level_w[l];
    const int64_t HW = (int64_t)H * W;
    /* stride = input / feature size; anchors scaled to grid units (:254-262) in fp32 */
    const float stride_h = (float)((double)c->input_h / H), stride_w = (float)((double)c->input_w / W);
    const float* base = c->levels[l] + (int64_t)b * 3 * attrs * HW;
    for (int a = 0; a < 3; ++a) {
      const float aw = c->anchors[(3 * l + a) * 2 + 0] / stride_w;
      const float ah = c->anchors[(3 * l + a) * 2 + 1] / stride_h;
      for (int64_t cell = 0; cell < HW; ++cell) {
        const float* p = base + (int64_t)a * attrs * HW + cell;
        const float gx = (float)(cell % W), gy = (float)(cell / W);
        float x = sigmoidf_ref(p[0 * HW]), y = sigmoidf_ref(p[1 * HW]);
        float w = sigmoidf_ref(p[2 * HW]), h = sigmoidf_ref(p[3 * HW]);
        float bx = x * 2.0f - 0.5f + gx;
        float by = y * 2.0f - 0.5f + gy;
        float tw = w * 2.0f, th = h * 2.0f;
        float bw = tw * tw * aw;
        float bh = th * th * ah;
        float* o = ob + (off + (int64_t)a * HW + cell) * attrs;
        o[0] = bx / (float)W;
        o[1] = by / (float)H;
        o[2] = bw / (float)W;
        o[3] = bh / (float)H;
        o[4] = sigmoidf_ref(p[4 * HW]);
        for (int k = 0; k < nc; ++k) o[5 + k] = sigmoidf_ref(p[(int64_t)(5 + k) * HW]);
      }
    }
    off += 3 * HW;
  }
}

ORC_API void orc_yolov7_decode(const float* const* levels, int num_levels, const int* level_h, const int* level_w,
                               const float* anchors, int input_h, int input_w, int B, int nc, float* out) {
  orc_v7dec_ctx c = {levels, num_levels, level_h, level_w, anchors, input_h, input_w, nc, 0, out};
  for (int l = 0; l < num_levels; ++l) c.A += 3 * (int64_t)level_h[l] * level_w[l];
  orc_parallel_for(B, orc_v7dec_body, &c);
}

typedef struct {
  const float* pred; /* (B, A, 5 + nc) decoded */
  int64_t A;
  int nc;
  float conf_thres;
  double nms_thres;
  int cap;
  float* rows;         /* (B, cap, 7): x1,y1,x2,y2,obj,class_conf,class_pred (normalised) */
  int32_t* row_anchor; /* (B, cap) */
  int32_t* count;
  int32_t* cand_count;
} orc_v7nms_ctx;

static void orc_v7nms_body(int b, void* vctx) {
  const orc_v7nms_ctx* c = (const orc_v7nms_ctx*)vctx;
  const int64_t A = c->A;
  const int attrs = 5 + c->nc;
  const float* p = c->pred + (int64_t)b * A * attrs;
  float* box = (float*)malloc(sizeof(float) * 4 * (size_t)A);
  float* sc = (float*)malloc(sizeof(float) * (size_t)A);
  float* obj = (float*)malloc(sizeof(float) * (size_t)A);
  float* cc = (float*)malloc(sizeof(float) * (size_t)A);
  int32_t* cls = (int32_t*)malloc(sizeof(int32_t) * (size_t)A);
  int32_t* anc = (int32_t*)malloc(sizeof(int32_t) * (size_t)A);
  int64_t n = 0;
  for (int64_t a = 0; a < A; ++a) {
    const float* q = p + a * attrs;
    float best = q[5];
    int bj = 0;
    for (int k = 1; k < c->nc; ++k)
      if (q[5 + k] > best) {
        best = q[5 + k];
        bj = k;
      }
    if (!(q[4] * best >= c->conf_thres)) continue; /* (:377) non-strict */
    box[4 * n + 0] = q[0] - q[2] / 2.0f;             /* xywh_to_xyxy_torch (:361) */
    box[4 * n + 1] = q[1] - q[3] / 2.0f;
    box[4 * n + 2] = q[0] + q[2] / 2.0f;
    box[4 * n + 3] = q[1] + q[3] / 2.0f;
    obj[n] = q[4];
    cc[n] = best;
    sc[n] = q[4] * best;
    cls[n] = bj;
    anc[n] = (int32_t)a;
    ++n;
  }
  if (c->cand_count) c->cand_count[b] = (int32_t)n;
  int64_t* keep = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  int64_t k = orc_nms_per_class(box, sc, cls, n, c->nc, c->nms_thres, keep);
  for (int64_t t = 0; t < k && t < c->cap; ++t) {
    float* row = c->rows + ((int64_t)b * c->cap + t) * 7;
    memcpy(row, box + 4 * keep[t], sizeof(float) * 4);
    row[4] = obj[keep[t]];
    row[5] = cc[keep[t]];
    row[6] = (float)cls[keep[t]];
    c->row_anchor[(int64_t)b * c->cap + t] = anc[keep[t]];
  }
  c->count[b] = (int32_t)k;
  free(box);
  free(sc);
  free(obj);
  free(cc);
  free(cls);
  free(anc);
  free(keep);
}

ORC_API void orc_yolov7_nms(const float* pred, int B, int64_t A, int nc, float conf_thres, double nms_thres, int cap,
                            float* rows, int32_t* row_anchor, int32_t* count, int32_t* cand_count) {
  orc_v7nms_ctx c = {pred, A, nc, conf_thres, nms_thres, cap, rows, row_anchor, count, cand_count};
  orc_parallel_for(B, orc_v7nms_body, &c);
}

/* ========================================================================= */
/* YOLOv3 — core/predict/yolov3_decode.py:12-66 (predict_bounding_bbox,       */
/* Decoder), core/utils/nms.py:54-84 (yolo3_nms), core/utils/anchor.py:102-117 */
/* ========================================================================= */
/* One scale, dense: feature (N, 3*(5+nc), H, W) NCHW -> boxes (N*H*W*3, 4) xyxy and      */
/* scores (N*H*W*3, nc) in the reference's flattened order ((n*H + y)*W + x)*3 + a.        */
/* anchors_norm (3, 2): anchor / input size, already divided in fp32 by the caller exactly */
/* like generate_yolo3_anchor.                                                             */
ORC_API void orc_yolov3_scale(const float* feature, int N, int nc, int H, int W, const float* anchors_norm,
                              float* boxes, float* scores) {
  const int attrs = 5 + nc;
  const int64_t HW = (int64_t)H * W;
  for (int n = 0; n < N; ++n)
    for (int64_t cell = 0; cell < HW; ++cell)
      for (int a = 0; a < 3; ++a) {
        const float* p = feature + ((int64_t)n * 3 * attrs + (int64_t)a * attrs) * HW + cell;
        const int64_t row = ((int64_t)n * HW + cell) * 3 + a;
        const float gx = (float)(cell % W), gy = (float)(cell / W);
        /* (:22) box_xy = (sigmoid(t) + grid) / H  — H for both coordinates */
        const float bx = (sigmoidf_ref(p[0]) + gx) / (float)H;
        const float by = (sigmoidf_ref(p[HW]) + gy) / (float)H;
        /* (:23) box_wh = exp(t) * anchors */
        const float bw = expf(p[2 * HW]) * anchors_norm[2 * a + 0];
        const float bh = expf(p[3 * HW]) * anchors_norm[2 * a + 1];
        const float conf = sigmoidf_ref(p[4 * HW]);
        float* o = boxes + 4 * row; /* (:47) xy -/+ wh / 2 */
        o[0] = bx - bw / 2.0f;
        o[1] = by - bh / 2.0f;
        o[2] = bx + bw / 2.0f;
        o[3] = by + bh / 2.0f;
        for (int k = 0; k < nc; ++k) scores[row * nc + k] = conf * sigmoidf_ref(p[(int64_t)(5 + k) * HW]); /* (:49) */
      }
}

/* yolo3_nms (core/utils/nms.py:54-84) on boxes (M, 4), scores (M, nc): for each class ascending, the rows with
 * score >= conf (float32 compare) go through nms; out rows (K): box, score, class, source row.  Returns K
 * (rows beyond cap are counted but not written). */
ORC_API int64_t orc_yolo3_nms(const float* boxes, const float* scores, int64_t M, int nc, float conf_thres,
                              double iou_thres, int64_t cap, float* out_boxes, float* out_scores, int32_t* out_cls,
                              int64_t* out_row, int64_t* cand_count) {
  float* cb = (float*)malloc(sizeof(float) * 4 * (size_t)(M > 0 ? M : 1));
  float* cs = (float*)malloc(sizeof(float) * (size_t)(M > 0 ? M : 1));
  int64_t* ci = (int64_t*)malloc(sizeof(int64_t) * (size_t)(M > 0 ? M : 1));
  int64_t* ck = (int64_t*)malloc(sizeof(int64_t) * (size_t)(M > 0 ? M : 1));
  int64_t total = 0, cands = 0;
  for (int c = 0; c < nc; ++c) {
    int64_t m = 0;
    for (int64_t i = 0; i < M; ++i)
      if (scores[i * nc + c] >= conf_thres) {
        memcpy(cb + 4 * m, boxes + 4 * i, sizeof(float) * 4);
        cs[m] = scores[i * nc + c];
        ci[m] = i;
        ++m;
      }
    cands += m;
    if (m == 0) continue;
    int64_t k = orc_nms(cb, cs, m, iou_thres, ck);
    for (int64_t t = 0; t < k; ++t) {
      if (total < cap) {
        memcpy(out_boxes + 4 * total, cb + 4 * ck[t], sizeof(float) * 4);
        out_scores[total] = cs[ck[t]];
        out_cls[total] = c;
        out_row[total] = ci[ck[t]];
      }
      ++total;
    }
  }
  if (cand_count) *cand_count = cands;
  free(cb);
  free(cs);
  free(ci);
  free(ck);
  return total;
}

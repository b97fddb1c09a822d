// yolov8_cell.cuh — per-cell arithmetic of the YOLOv8 head decode, shared by the streaming decode kernel
// (yolov8_decode.cu) and the head-fused tcgen05 kernel (yolov8_head_fused.cu): DFL softmax-integral
// (core/models/yolov8/modules.py:80-82), running class argmax, make_anchors + dist2bbox + xywh2xyxy
// (core/utils/anchor.py:126-145, core/utils/bboxes.py:213-222, core/utils/ultralytics_ops.py:360-375).
#pragma once

#include "cvpp_common.cuh"

namespace cvpp {

constexpr int kRegMax = 16;      // DFL bins (reference hard-codes 16, modules.py:413)

// ---- per-cell arithmetic (Appendix B of SURVEY.md: one fp32 rounding per reference op) ----------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// softmax over the 16 bins followed by the arange(16) 1x1 conv: sum_k k * softmax(x)_k
__device__ __forceinline__ float dfl16(const float (&v)[kRegMax]) {
  float m = v[0];
#pragma unroll
  for (int k = 1; k < kRegMax; ++k) m = fmaxf(m, v[k]);
  const float kLog2e = 1.4426950408889634f;
  const float mb = -m * kLog2e;
  float sum = 0.0f, wsum = 0.0f;
#pragma unroll
  for (int k = 0; k < kRegMax; ++k) {
    float e = ex2_approx(fmaf(v[k], kLog2e, mb));  // exp(v - m)
    sum += e;
    wsum = fmaf((float)k, e, wsum);
  }
  return fdiv(wsum, sum);
}

// one step of the running class scan on logits.  `prev` is the largest logit seen BEFORE the current
// best (= the old best at the last improvement): only an earlier index can steal the reference's
// first-index-of-max-sigmoid when two logits round to the same sigmoid, so it is all the tie check needs.
__device__ __forceinline__ void class_step(float x, int c, float& best, int& arg, float& prev) {
  const bool gt = x > best;
  prev = gt ? best : prev;
  arg = gt ? c : arg;
  best = fmaxf(best, x);
}

struct CellBox {
  float cx, cy, w, h;      // Detect output (xywh, input pixels)
  float x1, y1, x2, y2;    // after xywh2xyxy
};

__device__ __forceinline__ CellBox cell_box(int cell, int W, float stride, float dl, float dt, float dr, float db) {
  int iy = cell / W, ix = cell - iy * W;
  float ax = (float)ix + 0.5f, ay = (float)iy + 0.5f;  // make_anchors: arange + 0.5
  float x1 = fsub(ax, dl), y1 = fsub(ay, dt);          // dist2bbox: anchor - lt
  float x2 = fadd(ax, dr), y2 = fadd(ay, db);          //            anchor + rb
  CellBox o;
  o.cx = fmul(fmul(fadd(x1, x2), 0.5f), stride);       // ((x1+x2)/2) * stride
  o.cy = fmul(fmul(fadd(y1, y2), 0.5f), stride);
  o.w = fmul(fsub(x2, x1), stride);
  o.h = fmul(fsub(y2, y1), stride);
  float hw = fmul(o.w, 0.5f), hh = fmul(o.h, 0.5f);    // xywh2xyxy: x -/+ w/2
  o.x1 = fsub(o.cx, hw);
  o.y1 = fsub(o.cy, hh);
  o.x2 = fadd(o.cx, hw);
  o.y2 = fadd(o.cy, hh);
  return o;
}

}  // namespace cvpp

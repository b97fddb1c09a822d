// map_match.cu — on-device matching of detections to ground truth for the VOC-style mAP (sm_100a).
//
// Replaces the inner loops of get_map (reference core/metrics/mAP.py:302-835): for every detection of a class, in
// confidence order (:441), the ground-truth box of the same image and class with the highest overlap is looked up
// (:486-503, "+1" pixel convention, Python doubles, strict `>` so the FIRST maximum wins); the detection is a true
// positive when that overlap is >= MINOVERLAP, the box is not "difficult" and no earlier detection has used it
// (:508-515), a false positive otherwise - except that a match with a difficult box counts as neither (:510).
//
// The `used` flag is the only sequential dependency, and it never leaves one (image, class) group: the first
// detection - in processing order - whose best box is g (with enough overlap) takes it, every later one with the
// same best box is a false positive.  Processing order is confidence descending with ties in file / line order;
// inside one image that is the LINE order, because the evaluators emit every (image, class) group in descending
// score order and the reference's 6-character score truncation is monotone (for scores >= 1e-4, where numpy prints
// positional notation).  So: one CTA per image, a thread per detection finds its best box in double precision
// (boxes arrive as integer-valued floats: every product and sum is exact, the one division is IEEE like
// Python's), then a shared-memory pass marks, for every ground-truth box, the first claimant.
//
// Input rows are the CVPP_ROWS_VOC layout of cvpp_detection_epilogue(_compact): cls, score, int(l), int(t), int(r),
// int(b).  Output: flag[d] = 1 true positive, 2 false positive, 0 neither (difficult match / class without
// ground truth is the caller's business), best_gt[d] = index of the matched box in the image's list or -1, ovmax[d].
#include "cvpp_common.cuh"

namespace cvpp {

constexpr int kMmThreads = 256;

__global__ void __launch_bounds__(kMmThreads)
voc_match_kernel(const float* __restrict__ det_rows, const int32_t* __restrict__ det_offset, const float4* __restrict__ gt_box,
                 const int32_t* __restrict__ gt_cls, const int32_t* __restrict__ gt_difficult,
                 const int32_t* __restrict__ gt_offset, double min_overlap, int32_t* __restrict__ flag,
                 int32_t* __restrict__ best_gt, double* __restrict__ ovmax_out, int32_t* __restrict__ claim /*[G]*/) {
  const int b = blockIdx.x;
  const int d0 = det_offset[b], d1 = det_offset[b + 1];
  const int g0 = gt_offset[b], g1 = gt_offset[b + 1];
  // pass 1: best ground-truth box of every detection; claim[g] = lowest detection index that wants g
  for (int g = g0 + threadIdx.x; g < g1; g += kMmThreads) claim[g] = 0x7fffffff;
  __syncthreads();
  for (int d = d0 + threadIdx.x; d < d1; d += kMmThreads) {
    const float* r = det_rows + (int64_t)d * 6;
    const int c = (int)r[0];
    const double bl = (double)r[2], bt = (double)r[3], br = (double)r[4], bb = (double)r[5];
    double ovmax = -1.0;
    int match = -1;
    for (int g = g0; g < g1; ++g) {
      if (gt_cls[g] != c) continue;
      const float4 q = gt_box[g];
      // explicit round-to-nearest intrinsics: one IEEE rounding per Python operation, never contracted into FMAs
      const double iw = __dadd_rn(__dsub_rn(fmin(br, (double)q.z), fmax(bl, (double)q.x)), 1.0);
      const double ih = __dadd_rn(__dsub_rn(fmin(bb, (double)q.w), fmax(bt, (double)q.y)), 1.0);
      if (iw > 0.0 && ih > 0.0) {
        const double a_det = __dmul_rn(__dadd_rn(__dsub_rn(br, bl), 1.0), __dadd_rn(__dsub_rn(bb, bt), 1.0));
        const double a_gt = __dmul_rn(__dadd_rn(__dsub_rn((double)q.z, (double)q.x), 1.0),
                                      __dadd_rn(__dsub_rn((double)q.w, (double)q.y), 1.0));
        const double inter = __dmul_rn(iw, ih);
        const double ua = __dsub_rn(__dadd_rn(a_det, a_gt), inter);
        const double ov = __ddiv_rn(inter, ua);
        if (ov > ovmax) {
          ovmax = ov;
          match = g;
        }
      }
    }
    best_gt[d] = match >= 0 ? match - g0 : -1;
    ovmax_out[d] = ovmax;
    int f = 2;  // false positive unless shown otherwise
    if (ovmax >= min_overlap) {  // (match >= 0 here: ovmax starts at -1 and min_overlap >= 0)
      if (gt_difficult[match]) {
        f = 0;
      } else {
        atomicMin(&claim[match], d);
        f = 3;  // candidate true positive: decided in pass 2
      }
    }
    flag[d] = f;
  }
  __syncthreads();
  // pass 2: the first claimant of a box is the true positive, later ones are repeated matches
  for (int d = d0 + threadIdx.x; d < d1; d += kMmThreads) {
    if (flag[d] == 3) flag[d] = claim[g0 + best_gt[d]] == d ? 1 : 2;
  }
}

int voc_match_launch(const float* det_rows, const int32_t* det_offset, const float* gt_box, const int32_t* gt_cls,
                     const int32_t* gt_difficult, const int32_t* gt_offset, int B, double min_overlap, int32_t* flag,
                     int32_t* best_gt, double* ovmax, int32_t* claim_ws, cudaStream_t stream) {
  if (!det_rows || !det_offset || !gt_box || !gt_cls || !gt_difficult || !gt_offset || !flag || !best_gt || !ovmax || !claim_ws) {
    set_error("voc_match: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || !(min_overlap >= 0.0 && min_overlap <= 1.0)) {
    set_error("voc_match: bad arguments (B=%d min_overlap=%f)", B, min_overlap);
    return CVPP_ERR_INVALID_ARG;
  }
  if (reinterpret_cast<uintptr_t>(gt_box) & 15u) {
    set_error("voc_match: gt_box must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  if (B == 0) return CVPP_OK;
  voc_match_kernel<<<B, kMmThreads, 0, stream>>>(det_rows, det_offset, reinterpret_cast<const float4*>(gt_box), gt_cls,
                                                 gt_difficult, gt_offset, min_overlap, flag, best_gt, ovmax, claim_ws);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

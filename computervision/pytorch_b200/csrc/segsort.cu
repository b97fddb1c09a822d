// segsort.cu — kernel 2: segmented (per-image) sort of candidate keys (sm_100a).
//
// Replaces: x[x[:, 4].argsort(descending=True)[:max_nms]]  core/utils/ultralytics_ops.py:240,
// the stable descending score sort inside torchvision nms (csrc/ops/cpu/nms_kernel.cpp), and the
// per-class partition done with torch.unique / boolean gathers (torchvision/ops/boxes.py:112-116,
// core/algorithms/yolo_v7.py:396-400, core/algorithms/ssd.py:256-264, core/utils/nms.py:66-68).
//
// One CTA per image.  Keys are re-packed score-major ([inv_score | anchor | class]) and sorted
// ascending = score descending, lower anchor first on ties: the one global order every consumer needs
// (max_nms cut, trick-branch NMS, final score-ordered output); the per-class partition is a stable
// counting split inside the NMS kernel.  Segments that fit shared memory (<= 16384 keys, 128 KB) use
// a hybrid bitonic network whose warp-local substeps run in registers with shuffles (27 block
// barriers for 4096 keys instead of 78); larger ones run a plain network on a global scratch row.
// Latency-bound, not bandwidth-bound: ~16 B per candidate of traffic.
#include "cvpp_common.cuh"

namespace cvpp {

constexpr int kSortThreads = 1024;

__global__ void __launch_bounds__(kSortThreads, 1)
segsort_kernel(uint64_t* __restrict__ keys, int32_t* __restrict__ count, int max_cand, int max_nms,
               uint64_t* __restrict__ ws, int64_t ws_stride, int smem_elems, const int32_t* __restrict__ skip,
               uint64_t* __restrict__ out_keys, int32_t* __restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* sbuf = reinterpret_cast<uint64_t*>(smem_raw);
  const int b = blockIdx.x;
  if (skip && skip[b]) return;  // already done by the fused sort+NMS kernel
  int n = count[b];
  if (n > max_cand) n = max_cand;
  if (n <= 0) {
    if (out_count && threadIdx.x == 0) out_count[b] = n;
    return;
  }
  uint64_t* kb = keys + (int64_t)b * max_cand;
  uint64_t* ko = out_keys ? out_keys + (int64_t)b * max_cand : kb;  // out of place keeps the caller's keys intact
  const int P = pow2_ceil(n < 32 ? 32 : n);
  const bool in_smem = P <= smem_elems;
  uint64_t* a = in_smem ? sbuf : (ws + (int64_t)b * ws_stride);
  for (int i = threadIdx.x; i < P; i += blockDim.x) a[i] = i < n ? key_to_score_major(kb[i]) : ~0ull;
  __syncthreads();
  if (in_smem)
    block_sort_smem(a, P);
  else
    bitonic_sort_u64_generic(a, P);
  // ultralytics :240 keeps only the max_nms best scores (over all classes) before batched_nms
  const int n_final = (max_nms > 0 && n > max_nms) ? max_nms : n;
  for (int i = threadIdx.x; i < n_final; i += blockDim.x) ko[i] = a[i];
  if (threadIdx.x == 0) {
    if (out_count) out_count[b] = n_final;
    else if (n_final != n) count[b] = n_final;
  }
}

size_t segsort_workspace_bytes(int B, int max_cand) {
  int P = 32;
  while (P < max_cand) P <<= 1;
  return (size_t)B * (size_t)P * sizeof(uint64_t);
}

int segsort_launch_skip(uint64_t* keys, int32_t* cand_count, int B, int max_cand, int rule, int max_nms, void* workspace,
                        size_t workspace_bytes, const int32_t* skip, uint64_t* out_keys, int32_t* out_count,
                        cudaStream_t stream);

int segsort_launch(uint64_t* keys, int32_t* cand_count, int B, int max_cand, int rule, int max_nms,
                   void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  return segsort_launch_skip(keys, cand_count, B, max_cand, rule, max_nms, workspace, workspace_bytes, nullptr, nullptr,
                             nullptr, stream);
}

int segsort_launch_skip(uint64_t* keys, int32_t* cand_count, int B, int max_cand, int rule, int max_nms, void* workspace,
                        size_t workspace_bytes, const int32_t* skip, uint64_t* out_keys, int32_t* out_count,
                        cudaStream_t stream) {
  if (!keys || !cand_count || B < 0 || max_cand < 1) {
    set_error("segmented_sort: NULL pointer or bad sizes (B=%d, max_cand=%d)", B, max_cand);
    return CVPP_ERR_INVALID_ARG;
  }
  (void)rule;  // the sort order no longer depends on the batched_nms branch
  if (B == 0) return CVPP_OK;
  int P = 32;
  while (P < max_cand) P <<= 1;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != CVPP_OK) return rc;
  const int max_smem = di.max_smem;
  int smem_elems = 1;
  while ((size_t)smem_elems * 2 * sizeof(uint64_t) <= (size_t)max_smem && smem_elems * 2 <= 16384) smem_elems <<= 1;
  if (smem_elems > P) smem_elems = P;
  if (P > smem_elems) {
    if (!workspace || workspace_bytes < segsort_workspace_bytes(B, max_cand)) {
      set_error("segmented_sort: workspace of %zu bytes needed, got %zu", segsort_workspace_bytes(B, max_cand),
                workspace_bytes);
      return CVPP_ERR_WORKSPACE;
    }
  }
  size_t smem = (size_t)smem_elems * sizeof(uint64_t);
  static unsigned long long attr_done = 0;
  static int attr_bytes = 0;  // the attribute must cover the largest request seen so far
  if ((int)smem > attr_bytes) {
    attr_done = 0;
    attr_bytes = (int)smem;
  }
  rc = ensure_smem_attr(reinterpret_cast<const void*>(segsort_kernel), attr_bytes, di.device, &attr_done);
  if (rc != CVPP_OK) return rc;
  segsort_kernel<<<B, kSortThreads, smem, stream>>>(keys, cand_count, max_cand, max_nms,
                                                    reinterpret_cast<uint64_t*>(workspace), (int64_t)P, smem_elems, skip, out_keys, out_count);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

// pred_filter.cu — confidence filter on a decoded prediction tensor (B, 4+nc+nm, A) (sm_100a).
//
// Replaces the candidate stage of non_max_suppression, core/utils/ultralytics_ops.py:190
// (amax over class scores > conf), :204 (transpose + boolean gather), :220 (xywh2xyxy :360-375),
// :225-226 (conf, j = cls.max(1); conf > conf_thres) for callers that hold the dense `y`
// (the Detect output) instead of the raw head.  Memory-bound on (4+nc) x A fp32 per image;
// thread-per-anchor with coalesced row reads (consecutive lanes = consecutive anchors).
#include "cvpp_common.cuh"

namespace cvpp {

__global__ void __launch_bounds__(256)
pred_filter_kernel(const float* __restrict__ pred, int channels, int nc, int A, float conf_thres,
                   uint64_t* __restrict__ cand_key, int32_t* __restrict__ cand_count, float4* __restrict__ box_dense,
                   int max_cand) {
  const int b = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const float* p = pred + (int64_t)b * channels * A;
  bool cand = false;
  float best = -INFINITY;
  int arg = 0;
  if (a < A) {
    const float* c = p + (int64_t)4 * A + a;
#pragma unroll 8
    for (int k = 0; k < nc; ++k) {
      float v = __ldg(c + (int64_t)k * A);
      if (v > best) {  // strict: first index wins ties, like torch.max(dim)
        best = v;
        arg = k;
      }
    }
    cand = best > conf_thres;
  }
  const unsigned mask = __ballot_sync(0xffffffffu, cand);
  if (mask == 0) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(cand_count + b, __popc(mask));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (cand) {
    const float cx = p[a], cy = p[(int64_t)A + a], w = p[2 * (int64_t)A + a], h = p[3 * (int64_t)A + a];
    const float hw = fmul(w, 0.5f), hh = fmul(h, 0.5f);
    const float4 box = make_float4(fsub(cx, hw), fsub(cy, hh), fadd(cx, hw), fadd(cy, hh));
    const int slot = base + __popc(mask & ((1u << lane) - 1u));
    if (slot < max_cand) cand_key[(int64_t)b * max_cand + slot] = key_pack((uint32_t)arg, __float_as_uint(best), (uint32_t)a);
    box_dense[(int64_t)b * A + a] = box;
  }
}

int pred_filter_launch(const float* pred, int B, int channels, int nc, int64_t A, float conf_thres, uint64_t* cand_key,
                       int32_t* cand_count, float* box_dense, int max_cand, cudaStream_t stream) {
  if (!pred || !cand_key || !cand_count || !box_dense) {
    set_error("pred_filter: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || nc < 1 || nc > CVPP_MAX_CLASSES || channels < 4 + nc || A < 1 || A > CVPP_MAX_ANCHORS || max_cand < 1) {
    set_error("pred_filter: bad sizes (B=%d channels=%d nc=%d A=%lld max_cand=%d)", B, channels, nc, (long long)A,
              max_cand);
    return CVPP_ERR_INVALID_ARG;
  }
  if (reinterpret_cast<uintptr_t>(box_dense) & 15u) {
    set_error("pred_filter: box_dense must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  CVPP_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(int32_t) * (size_t)B, stream));
  if (B == 0) return CVPP_OK;
  dim3 grid((unsigned)((A + 255) / 256), (unsigned)B);
  pred_filter_kernel<<<grid, 256, 0, stream>>>(pred, channels, nc, (int)A, conf_thres, cand_key, cand_count,
                                              reinterpret_cast<float4*>(box_dense), max_cand);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// keep_classes: `x = x[(x[:, 5:6] == classes).any(1)]` (core/utils/ultralytics_ops.py:229-230) on the candidate
// keys: keys whose class bit is not set in the mask are dropped, in place, one CTA per image.  The keys are an
// unordered set (the sort comes later), so the compaction need not be stable; a chunk is read into registers
// before anything of it is overwritten, and the write cursor never passes the read cursor.
// ---------------------------------------------------------------------------------------------------
struct ClassMask {
  uint32_t w[CVPP_MAX_CLASSES / 32];
};

__global__ void __launch_bounds__(1024) keep_classes_kernel(uint64_t* __restrict__ cand_key, int32_t* __restrict__ cand_count,
                                                            int max_cand, const __grid_constant__ ClassMask mask) {
  __shared__ int sh_warp[32];
  __shared__ int sh_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t* keys = cand_key + (int64_t)b * max_cand;
  const int n = min(max(cand_count[b], 0), max_cand);
  if (tid == 0) sh_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    uint64_t k = 0;
    bool keep = false;
    if (i < n) {
      k = keys[i];
      const uint32_t c = key_cls(k);
      keep = (mask.w[c >> 5] >> (c & 31)) & 1u;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) sh_warp[warp] = __popc(bal);
    __syncthreads();   // every key of the chunk is in a register now
    int woff = 0, total = 0;
    for (int q = 0; q < 32; ++q) {
      const int u = sh_warp[q];
      if (q < warp) woff += u;
      total += u;
    }
    const int base = sh_base;
    if (keep) keys[base + woff + __popc(bal & ((1u << lane) - 1u))] = k;
    __syncthreads();
    if (tid == 0) sh_base = base + total;
    __syncthreads();
  }
  if (tid == 0) cand_count[b] = sh_base;
}

int keep_classes_launch(uint64_t* cand_key, int32_t* cand_count, int B, int max_cand, const int32_t* classes, int n_classes,
                        cudaStream_t stream) {
  if (!cand_key || !cand_count || (n_classes > 0 && !classes)) {
    set_error("keep_classes: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || max_cand < 1 || n_classes < 0) {
    set_error("keep_classes: bad sizes (B=%d max_cand=%d n_classes=%d)", B, max_cand, n_classes);
    return CVPP_ERR_INVALID_ARG;
  }
  ClassMask m{};
  for (int i = 0; i < n_classes; ++i) {
    const int c = classes[i];
    if (c >= 0 && c < CVPP_MAX_CLASSES) m.w[c >> 5] |= 1u << (c & 31);  // ids outside the key range match nothing
  }
  if (B == 0) return CVPP_OK;
  keep_classes_kernel<<<B, 1024, 0, stream>>>(cand_key, cand_count, max_cand, m);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

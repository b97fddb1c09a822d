// nms.cu — kernel 3: class-aware greedy NMS on sorted candidate keys + gather of survivors (sm_100a).
//
// Replaces torchvision.ops.nms / batched_nms (torchvision/ops/boxes.py:20-120,
// csrc/ops/cpu/nms_kernel.cpp) at the reference call sites core/utils/ultralytics_ops.py:247-257,
// core/utils/nms.py:69,134, core/algorithms/yolo_v7.py:407, core/algorithms/ssd.py:267.
//
// One CTA per image.  Candidates arrive sorted (class asc, score desc, anchor asc), so every class is
// a contiguous segment already in torchvision's processing order.
//   * suppression state is a shared-memory bitmask (`alive`, one bit per sorted position);
//   * a 32-candidate word is resolved by one warp: lane j holds box j, kept boxes are broadcast with
//     shuffles, the IoU test runs in all lanes and the verdicts are collected with __ballot_sync and
//     cleared from the mask; survivors of a word are then applied to the later words of the segment;
//   * small segments are handled one per warp (classes in parallel), large ones by the whole CTA
//     (word resolve by warp 0, application to later words spread over all warps);
//   * IoU arithmetic reproduces torchvision's CPU kernel in fp32 op for op (cvpp_common.cuh), in
//     both batched_nms branches: per-class on raw boxes, or class-agnostic on boxes shifted by
//     cls * (max_coord + 1) (the "coordinate trick", taken when an image has <= 1000 candidates);
//   * survivors are emitted in class-major order, or merged in score order (second bitonic sort in
//     shared memory) and capped at max_det.
// Latency / SM-bound; the only HBM traffic is 8 B keys + gathered 16 B boxes (L2 resident).
#include "cvpp_common.cuh"

namespace cvpp {

constexpr int kNmsThreads = 512;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kCoopMinWords = 12;  // segments spanning more words than this use the whole CTA


struct NmsParams {
  const uint64_t* sorted_key;
  const int32_t* cand_count;
  const float4* box_dense;
  int max_cand;
  int64_t A;
  int nc;
  float thr_eff;
  int rule;
  int order;
  int max_det;
  int max_out;
  float4* det_box;
  float* det_score;
  int32_t* det_cls;
  int32_t* det_anchor;
  int32_t* det_count;
  // scratch
  float4* ws_box;      // [B][max_cand] sorted boxes for images that do not fit shared memory
  float* ws_area;      // [B][max_cand]
  uint64_t* ws_sort;   // [B][pow2(max_cand)] second sort for images that do not fit shared memory
  int64_t ws_sort_stride;
  int smem_boxes;      // boxes that fit in shared memory
  int alive_words;     // words reserved for the alive mask
};

struct BoxView {  // sorted boxes + areas of one image (shared or global)
  const float4* box;
  const float* area;
};

__device__ __forceinline__ float4 shfl_box(const float4& b, int src) {
  float4 o;
  o.x = __shfl_sync(0xffffffffu, b.x, src);
  o.y = __shfl_sync(0xffffffffu, b.y, src);
  o.z = __shfl_sync(0xffffffffu, b.z, src);
  o.w = __shfl_sync(0xffffffffu, b.w, src);
  return o;
}

// keep only the lowest `room` set bits of m
__device__ __forceinline__ uint32_t lowest_bits(uint32_t m, int room) {
  uint32_t o = 0;
  while (m && room > 0) {
    uint32_t low = m & (0u - m);
    o |= low;
    m ^= low;
    --room;
  }
  return o;
}

// Resolve word w of segment [s, e): returns the kept mask (identical in all lanes).  All 32 lanes call.
__device__ __forceinline__ uint32_t resolve_word(int w, int s, int e, const BoxView& bv, uint32_t* alive,
                                                 float thr_eff, float4& box, float& area) {
  const int lane = threadIdx.x & 31;
  const int r = (w << 5) + lane;
  const bool in = r >= s && r < e;
  const uint32_t aw = alive[w];
  bool my_alive = in && ((aw >> lane) & 1u);
  box = make_float4(0.f, 0.f, 0.f, 0.f);
  area = 0.f;
  if (in) {
    box = bv.box[r];
    area = bv.area[r];
  }
  uint32_t am = __ballot_sync(0xffffffffu, my_alive);
  uint32_t rem = am;
  while (rem) {
    const int i = __ffs(rem) - 1;
    rem &= rem - 1;
    const float4 bi = shfl_box(box, i);
    const float ai = __shfl_sync(0xffffffffu, area, i);
    const bool sup = my_alive && lane > i && iou_suppresses(bi, ai, box, area, thr_eff);
    const uint32_t sm = __ballot_sync(0xffffffffu, sup);
    am &= ~sm;
    rem &= ~sm;
    if (sup) my_alive = false;
  }
  return am;
}

// Apply the kept boxes `am` of word w (lane i of the calling warp holds box i in box/area when
// from_regs, otherwise they are read from bv) to word w2 of the same segment.
template <bool FROM_REGS>
__device__ __forceinline__ void apply_word(uint32_t am, int w, int w2, int e, const BoxView& bv, uint32_t* alive,
                                           float thr_eff, const float4& box, float area) {
  const int lane = threadIdx.x & 31;
  const int r2 = (w2 << 5) + lane;
  const uint32_t aw2 = alive[w2];
  const bool al2 = r2 < e && ((aw2 >> lane) & 1u);
  if (!__any_sync(0xffffffffu, al2)) return;
  float4 b2 = make_float4(0.f, 0.f, 0.f, 0.f);
  float a2 = 0.f;
  if (al2) {
    b2 = bv.box[r2];
    a2 = bv.area[r2];
  }
  bool sup2 = false;
  uint32_t km = am;
  while (km) {
    const int i = __ffs(km) - 1;
    km &= km - 1;
    float4 bi;
    float ai;
    if (FROM_REGS) {
      bi = shfl_box(box, i);
      ai = __shfl_sync(0xffffffffu, area, i);
    } else {
      bi = bv.box[(w << 5) + i];
      ai = bv.area[(w << 5) + i];
    }
    if (al2 && !sup2) sup2 = iou_suppresses(bi, ai, b2, a2, thr_eff);
  }
  const uint32_t sm2 = __ballot_sync(0xffffffffu, sup2);
  if (lane == 0 && sm2) atomicAnd(&alive[w2], ~sm2);
}

__device__ __forceinline__ uint32_t seg_mask_of_word(int w, int s, int e) {
  const int lo = max(s - (w << 5), 0), hi = min(e - (w << 5), 32);
  if (hi <= lo) return 0u;
  const uint32_t upto_hi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
  return upto_hi & ~((1u << lo) - 1u);
}

// One warp runs greedy NMS over segment [s, e).  cap > 0 stops after `cap` survivors.
__device__ void nms_segment_warp(int s, int e, int cap, const BoxView& bv, uint32_t* alive, float thr_eff) {
  const int lane = threadIdx.x & 31;
  const int w0 = s >> 5, w1 = (e - 1) >> 5;
  int kept_total = 0;
  for (int w = w0; w <= w1; ++w) {
    float4 box;
    float area;
    uint32_t am = resolve_word(w, s, e, bv, alive, thr_eff, box, area);
    if (cap > 0) am = lowest_bits(am, cap - kept_total);
    kept_total += __popc(am);
    const uint32_t inmask = seg_mask_of_word(w, s, e);
    if (lane == 0) atomicAnd(&alive[w], ~(inmask & ~am));
    if (cap > 0 && kept_total >= cap) {
      for (int w2 = w + 1 + lane; w2 <= w1; w2 += 32) atomicAnd(&alive[w2], ~seg_mask_of_word(w2, s, e));
      break;
    }
    if (am)
      for (int w2 = w + 1; w2 <= w1; ++w2) apply_word<true>(am, w, w2, e, bv, alive, thr_eff, box, area);
    __syncwarp();
  }
}

// The whole CTA runs greedy NMS over one (large) segment.  All threads call.
__device__ void nms_segment_cta(int s, int e, int cap, const BoxView& bv, uint32_t* alive, float thr_eff,
                                uint32_t* sh_mask, int* sh_kept) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w0 = s >> 5, w1 = (e - 1) >> 5;
  if (threadIdx.x == 0) *sh_kept = 0;
  __syncthreads();
  for (int w = w0; w <= w1; ++w) {
    if (warp == 0) {
      float4 box;
      float area;
      uint32_t am = resolve_word(w, s, e, bv, alive, thr_eff, box, area);
      int kept_total = *sh_kept;
      if (cap > 0) am = lowest_bits(am, cap - kept_total);
      const uint32_t inmask = seg_mask_of_word(w, s, e);
      if (lane == 0) {
        atomicAnd(&alive[w], ~(inmask & ~am));
        *sh_mask = am;
        *sh_kept = kept_total + __popc(am);
      }
    }
    __syncthreads();
    const uint32_t am = *sh_mask;
    const bool done = cap > 0 && *sh_kept >= cap;
    if (done) {
      for (int w2 = w + 1 + threadIdx.x; w2 <= w1; w2 += blockDim.x) atomicAnd(&alive[w2], ~seg_mask_of_word(w2, s, e));
    } else if (am) {
      float4 dummy = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int w2 = w + 1 + warp; w2 <= w1; w2 += kNmsWarps) apply_word<false>(am, w, w2, e, bv, alive, thr_eff, dummy, 0.f);
    }
    __syncthreads();
    if (done) break;
  }
}

__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(const __grid_constant__ NmsParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [boxes float4 x smem_boxes][areas float x smem_boxes][alive u32 x alive_words][seg int x 2*nc]
  float4* sh_box = reinterpret_cast<float4*>(smem_raw);
  float* sh_area = reinterpret_cast<float*>(sh_box + p.smem_boxes);
  uint32_t* alive = reinterpret_cast<uint32_t*>(sh_area + p.smem_boxes);
  int* seg_begin = reinterpret_cast<int*>(alive + p.alive_words);
  int* seg_end = seg_begin + p.nc;
  __shared__ float sh_red[kNmsWarps];
  __shared__ int sh_scan[kNmsWarps];
  __shared__ int sh_running;
  __shared__ uint32_t sh_mask;
  __shared__ int sh_kept;
  __shared__ int sh_next;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  int n = p.cand_count[b];
  if (n > p.max_cand) n = p.max_cand;
  if (n <= 0) {
    if (tid == 0) p.det_count[b] = 0;
    return;
  }
  const uint64_t* keys = p.sorted_key + (int64_t)b * p.max_cand;
  const float4* dense = p.box_dense + (int64_t)b * p.A;
  const bool trick = rule_uses_trick(p.rule, n);
  const int nwords = (n + 31) >> 5;

  // ---- coordinate trick: offset = cls * (max over every coordinate + 1) (boxes.py:95-97) --------
  float mult = 0.f;
  if (trick) {
    float m = -INFINITY;
    for (int r = tid; r < n; r += kNmsThreads) {
      const float4 bx = dense[key_anchor(key_from_score_major(keys[r]))];
      m = fmaxf(m, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
    }
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) sh_red[warp] = m;
    __syncthreads();
    m = sh_red[0];
    for (int q = 1; q < kNmsWarps; ++q) m = fmaxf(m, sh_red[q]);
    mult = fadd(m, 1.0f);
  }

  // ---- stage sorted boxes (+ areas), alive mask, class segments ---------------------------------
  const bool in_smem = n <= p.smem_boxes;
  float4* wbox = in_smem ? sh_box : p.ws_box + (int64_t)b * p.max_cand;
  float* warea = in_smem ? sh_area : p.ws_area + (int64_t)b * p.max_cand;
  for (int c = tid; c < p.nc; c += kNmsThreads) {
    seg_begin[c] = 0;
    seg_end[c] = 0;
  }
  for (int w = tid; w < nwords; w += kNmsThreads) {
    const int rem = n - (w << 5);
    alive[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
  }
  __syncthreads();
  for (int r = tid; r < n; r += kNmsThreads) {
    uint64_t k = keys[r];
    if (trick) k = key_from_score_major(k);
    const int cls = (int)key_cls(k);
    float4 bx = dense[key_anchor(k)];
    if (trick) {
      const float off = fmul((float)cls, mult);
      bx.x = fadd(bx.x, off);
      bx.y = fadd(bx.y, off);
      bx.z = fadd(bx.z, off);
      bx.w = fadd(bx.w, off);
    } else if (cls < p.nc) {
      const int prev = r > 0 ? (int)key_cls(keys[r - 1]) : -1;
      if (prev != cls) {
        seg_begin[cls] = r;
        if (prev >= 0 && prev < p.nc) seg_end[prev] = r;
      }
      if (r == n - 1) seg_end[cls] = n;
    }
    wbox[r] = bx;
    warea[r] = box_area(bx);
  }
  if (tid == 0) sh_next = 0;
  __syncthreads();
  BoxView bv{wbox, warea};

  // ---- greedy suppression -------------------------------------------------------------------
  const int cap = (p.order == CVPP_ORDER_SCORE_DESC && p.max_det > 0) ? p.max_det : 0;
  if (trick) {
    if (nwords > kCoopMinWords) {
      nms_segment_cta(0, n, cap, bv, alive, p.thr_eff, &sh_mask, &sh_kept);
    } else {
      if (warp == 0) nms_segment_warp(0, n, cap, bv, alive, p.thr_eff);
      __syncthreads();
    }
  } else {
    // small segments: one warp each, classes claimed dynamically; large ones afterwards by the CTA
    for (;;) {
      int c = 0;
      if (lane == 0) c = atomicAdd(&sh_next, 1);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= p.nc) break;
      const int s = seg_begin[c], e = seg_end[c];
      if (e <= s) continue;
      if (((e - 1) >> 5) - (s >> 5) + 1 > kCoopMinWords) continue;
      nms_segment_warp(s, e, cap, bv, alive, p.thr_eff);
    }
    __syncthreads();
    for (int c = 0; c < p.nc; ++c) {
      const int s = seg_begin[c], e = seg_end[c];
      if (e <= s) continue;
      if (((e - 1) >> 5) - (s >> 5) + 1 <= kCoopMinWords) continue;
      nms_segment_cta(s, e, cap, bv, alive, p.thr_eff, &sh_mask, &sh_kept);
    }
  }
  __syncthreads();

  // ---- ordered compaction of the survivors ----------------------------------------------------
  // class-major (or trick: already global score order): emit directly.  Otherwise collect the
  // survivors' keys in score-major packing, sort them, emit the best max_det.
  const bool resort = (p.order == CVPP_ORDER_SCORE_DESC) && !trick;
  uint64_t* sortbuf = nullptr;
  int P2 = 0;
  if (resort) {
    P2 = pow2_ceil(n < 2 ? 2 : n);
    // the box staging area is dead now; reuse it when the keys fit
    sortbuf = ((size_t)P2 * sizeof(uint64_t) <= (size_t)p.smem_boxes * (sizeof(float4) + sizeof(float)))
                  ? reinterpret_cast<uint64_t*>(smem_raw)
                  : p.ws_sort + (int64_t)b * p.ws_sort_stride;
  }
  float4* ob = p.det_box + (int64_t)b * p.max_out;
  float* os = p.det_score + (int64_t)b * p.max_out;
  int32_t* oc = p.det_cls + (int64_t)b * p.max_out;
  int32_t* oa = p.det_anchor + (int64_t)b * p.max_out;

  auto emit = [&](int pos, uint64_t k_class_major) {
    if (pos >= p.max_out) return;
    const uint32_t anchor = key_anchor(k_class_major);
    ob[pos] = dense[anchor];
    os[pos] = __uint_as_float(key_score_bits(k_class_major));
    oc[pos] = (int32_t)key_cls(k_class_major);
    oa[pos] = (int32_t)anchor;
  };

  if (tid == 0) sh_running = 0;
  __syncthreads();
  for (int base = 0; base < nwords; base += kNmsThreads) {
    const int wi = base + tid;
    uint32_t m = wi < nwords ? alive[wi] : 0u;
    const int cnt = __popc(m);
    int incl = cnt;
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) sh_scan[warp] = incl;
    __syncthreads();
    int woff = 0, total = 0;
    for (int q = 0; q < kNmsWarps; ++q) {
      const int v = sh_scan[q];
      if (q < warp) woff += v;
      total += v;
    }
    int pos = sh_running + woff + incl - cnt;
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      uint64_t k = keys[(wi << 5) + bit];
      if (resort) {
        sortbuf[pos] = key_to_score_major(k);
      } else {
        emit(pos, trick ? key_from_score_major(k) : k);
      }
      ++pos;
    }
    __syncthreads();
    if (tid == 0) sh_running += total;
    __syncthreads();
  }
  const int n_kept = sh_running;
  if (!resort) {
    if (tid == 0) p.det_count[b] = n_kept;
    return;
  }
  // NOTE: sortbuf may alias the box staging area: every warp passed the barrier above, boxes are dead.
  const int P3 = pow2_ceil(n_kept < 2 ? 2 : n_kept);
  for (int i = n_kept + tid; i < P3; i += kNmsThreads) sortbuf[i] = ~0ull;
  __syncthreads();
  bitonic_sort_u64(sortbuf, P3);
  const int n_out = (p.max_det > 0 && n_kept > p.max_det) ? p.max_det : n_kept;
  for (int i = tid; i < n_out; i += kNmsThreads) emit(i, key_from_score_major(sortbuf[i]));
  if (tid == 0) p.det_count[b] = n_out;
}

static int pow2_ceil_host(int n) {
  int p = 2;
  while (p < n) p <<= 1;
  return p;
}

size_t nms_workspace_bytes(int B, int max_cand) {
  size_t per = (size_t)max_cand * (sizeof(float4) + sizeof(float)) + (size_t)pow2_ceil_host(max_cand) * sizeof(uint64_t);
  per = (per + 255) & ~(size_t)255;
  return per * (size_t)B + 256;
}

int nms_launch(const uint64_t* sorted_key, const int32_t* cand_count, const float* box_dense, int B, int max_cand,
               int64_t A, int nc, double iou_thres, int rule, int order, int max_det, int max_out, float* det_box,
               float* det_score, int32_t* det_cls, int32_t* det_anchor, int32_t* det_count, void* workspace,
               size_t workspace_bytes, cudaStream_t stream) {
  if (!sorted_key || !cand_count || !box_dense || !det_box || !det_score || !det_cls || !det_anchor || !det_count) {
    set_error("nms: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || max_cand < 1 || A < 1 || nc < 1 || nc > CVPP_MAX_CLASSES || max_out < 1) {
    set_error("nms: bad sizes (B=%d max_cand=%d A=%lld nc=%d max_out=%d)", B, max_cand, (long long)A, nc, max_out);
    return CVPP_ERR_INVALID_ARG;
  }
  if (!(iou_thres >= 0.0 && iou_thres <= 1.0)) {
    set_error("nms: Invalid IoU %f, valid values are between 0.0 and 1.0", iou_thres);
    return CVPP_ERR_INVALID_ARG;
  }
  if (rule < 0 || rule > 2 || order < 0 || order > 1) {
    set_error("nms: unknown rule %d / order %d", rule, order);
    return CVPP_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(box_dense) & 15u) || (reinterpret_cast<uintptr_t>(det_box) & 15u)) {
    set_error("nms: box arrays must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  if (B == 0) return CVPP_OK;
  if (!workspace || workspace_bytes < nms_workspace_bytes(B, max_cand)) {
    set_error("nms: workspace of %zu bytes needed, got %zu", nms_workspace_bytes(B, max_cand), workspace_bytes);
    return CVPP_ERR_WORKSPACE;
  }
  // (double)ovr > thr  <=>  ovr > largest float <= thr   (ovr is a float)
  float thr_eff = (float)iou_thres;
  if ((double)thr_eff > iou_thres) thr_eff = nextafterf(thr_eff, -INFINITY);

  int dev = 0, max_smem = 0;
  CVPP_CUDA_TRY(cudaGetDevice(&dev));
  CVPP_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

  NmsParams p{};
  p.sorted_key = sorted_key;
  p.cand_count = cand_count;
  p.box_dense = reinterpret_cast<const float4*>(box_dense);
  p.max_cand = max_cand;
  p.A = A;
  p.nc = nc;
  p.thr_eff = thr_eff;
  p.rule = rule;
  p.order = order;
  p.max_det = max_det;
  p.max_out = max_out;
  p.det_box = reinterpret_cast<float4*>(det_box);
  p.det_score = det_score;
  p.det_cls = det_cls;
  p.det_anchor = det_anchor;
  p.det_count = det_count;
  // workspace carve-up
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  p.ws_box = reinterpret_cast<float4*>(base);
  p.ws_area = reinterpret_cast<float*>(base + (size_t)B * max_cand * sizeof(float4));
  uintptr_t sort_base = (base + (size_t)B * max_cand * (sizeof(float4) + sizeof(float)) + 15) & ~(uintptr_t)15;
  p.ws_sort = reinterpret_cast<uint64_t*>(sort_base);
  p.ws_sort_stride = pow2_ceil_host(max_cand);
  // shared memory: boxes+areas (<= 8192), alive mask, class segments
  p.alive_words = (max_cand + 31) / 32;
  size_t fixed = (size_t)p.alive_words * 4 + (size_t)nc * 8 + 64;
  int smem_boxes = max_cand < 8192 ? max_cand : 8192;
  smem_boxes = (smem_boxes + 3) & ~3;
  while (smem_boxes > 0 && fixed + (size_t)smem_boxes * 20 > (size_t)max_smem - 1024) smem_boxes -= 256;
  if (smem_boxes < 0) smem_boxes = 0;
  if (fixed > (size_t)max_smem - 1024) {
    set_error("nms: max_cand=%d / nc=%d need more shared memory than the device has", max_cand, nc);
    return CVPP_ERR_UNSUPPORTED;
  }
  p.smem_boxes = smem_boxes;
  size_t smem = fixed + (size_t)smem_boxes * 20;
  CVPP_CUDA_TRY(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_kernel<<<B, kNmsThreads, smem, stream>>>(p);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

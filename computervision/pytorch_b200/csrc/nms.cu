// nms.cu — kernel 3: class-aware greedy NMS on score-sorted candidate keys + gather of survivors (sm_100a).
//
// Replaces torchvision.ops.nms / batched_nms (torchvision/ops/boxes.py:20-120,
// csrc/ops/cpu/nms_kernel.cpp) at the reference call sites core/utils/ultralytics_ops.py:247-257,
// core/utils/nms.py:69,134, core/algorithms/yolo_v7.py:407, core/algorithms/ssd.py:267.
//
// One CTA per image.  Candidates arrive in global score order (cvpp_segmented_sort).
//   * a STABLE counting split by class (warp match_any + per-warp class counters + a scan) lays the
//     boxes out class-major in shared memory, each class still in score order — torchvision's
//     processing order — without a second sort;
//   * suppression state is a shared-memory bitmask (`alive`, one bit per class-major position);
//   * a 32-candidate word is resolved by one warp: lane j holds box j, kept boxes are broadcast with
//     shuffles, the IoU test runs in all lanes and the verdicts are collected with __ballot_sync and
//     cleared from the mask; survivors of a word are then applied to the later words of the segment;
//   * small segments are handled one per warp (classes in parallel), large ones by the whole CTA
//     (word resolve by warp 0, application to later words spread over all warps);
//   * IoU arithmetic reproduces torchvision's CPU kernel in fp32 op for op (cvpp_common.cuh), in
//     both batched_nms branches: per-class on raw boxes, or class-agnostic on boxes shifted by
//     cls * (max_coord + 1) (the "coordinate trick", taken when an image has <= 1000 candidates);
//   * survivors are emitted class-major, or - through a second bitmask indexed by global score
//     rank - in score order capped at max_det.  Both are ordered ballot/popc compactions.
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
  float4* ws_box;      // [B][max_cand] class-major boxes for images that do not fit shared memory
  float* ws_area;      // [B][max_cand]
  int32_t* ws_rank;    // [B][max_cand] class-major position -> global score rank
  int smem_boxes;      // boxes that fit in shared memory
  int alive_words;     // words reserved for each of the two bitmasks
  int count_warps;     // warps taking part in the counting split (per-warp class counters)
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
  // layout: [boxes float4 x N][areas float x N][rank int x N][alive u32 x W][galive u32 x W]
  //         [seg_begin int x nc][seg_end int x nc][cnt int x count_warps x nc]
  float4* sh_box = reinterpret_cast<float4*>(smem_raw);
  float* sh_area = reinterpret_cast<float*>(sh_box + p.smem_boxes);
  int32_t* sh_rank = reinterpret_cast<int32_t*>(sh_area + p.smem_boxes);
  uint32_t* alive = reinterpret_cast<uint32_t*>(sh_rank + p.smem_boxes);
  uint32_t* galive = alive + p.alive_words;
  int* seg_begin = reinterpret_cast<int*>(galive + p.alive_words);
  int* seg_end = seg_begin + p.nc;
  int* cnt = seg_end + p.nc;
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
  const uint64_t* keys = p.sorted_key + (int64_t)b * p.max_cand;  // score-major: [inv_score | anchor | class]
  const float4* dense = p.box_dense + (int64_t)b * p.A;
  const bool trick = rule_uses_trick(p.rule, n);
  const int nwords = (n + 31) >> 5;
  const int nc = p.nc;

  const bool in_smem = n <= p.smem_boxes;
  float4* wbox = in_smem ? sh_box : p.ws_box + (int64_t)b * p.max_cand;
  float* warea = in_smem ? sh_area : p.ws_area + (int64_t)b * p.max_cand;
  int32_t* wrank = in_smem ? sh_rank : p.ws_rank + (int64_t)b * p.max_cand;

  for (int w = tid; w < nwords; w += kNmsThreads) {
    const int rem = n - (w << 5);
    alive[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    galive[w] = 0u;
  }
  if (tid == 0) sh_next = 0;

  if (trick) {
    // ---- coordinate trick: one class-agnostic segment in score order; boxes shifted by
    //      cls * (max over every coordinate + 1)  (boxes.py:95-97)
    float m = -INFINITY;
    for (int r = tid; r < n; r += kNmsThreads) {
      const float4 bx = dense[(uint32_t)(keys[r] >> 12) & 0x1fffffu];
      m = fmaxf(m, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
    }
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) sh_red[warp] = m;
    __syncthreads();
    m = sh_red[0];
    for (int q = 1; q < kNmsWarps; ++q) m = fmaxf(m, sh_red[q]);
    const float mult = fadd(m, 1.0f);
    for (int r = tid; r < n; r += kNmsThreads) {
      const uint64_t k = keys[r];
      const float off = fmul((float)(uint32_t)(k & 0xfffu), mult);
      float4 bx = dense[(uint32_t)(k >> 12) & 0x1fffffu];
      bx.x = fadd(bx.x, off);
      bx.y = fadd(bx.y, off);
      bx.z = fadd(bx.z, off);
      bx.w = fadd(bx.w, off);
      wbox[r] = bx;
      warea[r] = box_area(bx);
    }
    __syncthreads();
  } else {
    // ---- stable counting split by class: position = class start + #earlier ranks of that class
    const int cw = p.count_warps;
    for (int i = tid; i < cw * nc; i += kNmsThreads) cnt[i] = 0;
    __syncthreads();
    const int per_warp = (((n + cw - 1) / cw) + 31) & ~31;  // contiguous rank range per counting warp
    const int r_begin = warp * per_warp, r_end = min(n, r_begin + per_warp);
    if (warp < cw) {
      int* my = cnt + warp * nc;
      for (int r0 = r_begin; r0 < r_end; r0 += 32) {
        const int r = r0 + lane;
        const bool v = r < r_end;
        const int c = v ? (int)(keys[r] & 0xfffu) : -1 - lane;  // invalid lanes match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (v && c < nc && (peers & ((1u << lane) - 1u)) == 0) my[c] += __popc(peers);
        __syncwarp();
      }
    }
    __syncthreads();
    // per class: exclusive offsets over the counting warps, class totals into seg_end (temporarily)
    for (int c = tid; c < nc; c += kNmsThreads) {
      int run = 0;
      for (int w = 0; w < cw; ++w) {
        const int t = cnt[w * nc + c];
        cnt[w * nc + c] = run;
        run += t;
      }
      seg_end[c] = run;
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the class totals -> segment starts
      int running = 0;
      for (int c0 = 0; c0 < nc; c0 += 32) {
        const int c = c0 + lane;
        const int t = c < nc ? seg_end[c] : 0;
        int incl = t;
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += v;
        }
        if (c < nc) {
          seg_begin[c] = running + incl - t;
          seg_end[c] = running + incl;
        }
        running += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) sh_kept = running;  // candidates with a class id < nc (all of them, by contract)
    }
    __syncthreads();
    {
      const int nvalid = sh_kept;
      if (nvalid < n)
        for (int w = tid; w < nwords; w += kNmsThreads) {
          const int rem = nvalid - (w << 5);
          if (rem < 32) alive[w] &= rem <= 0 ? 0u : ((1u << rem) - 1u);
        }
    }
    if (warp < cw) {
      int* my = cnt + warp * nc;
      for (int r0 = r_begin; r0 < r_end; r0 += 32) {
        const int r = r0 + lane;
        const bool v = r < r_end;
        uint64_t k = 0;
        int c = -1 - lane;
        if (v) {
          k = keys[r];
          c = (int)(k & 0xfffu);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (v && c < nc) {
          const int pos = seg_begin[c] + my[c] + __popc(peers & ((1u << lane) - 1u));
          const float4 bx = dense[(uint32_t)(k >> 12) & 0x1fffffu];
          wbox[pos] = bx;
          warea[pos] = box_area(bx);
          wrank[pos] = r;
        }
        __syncwarp();
        if (v && c < nc && (peers & ((1u << lane) - 1u)) == 0) my[c] += __popc(peers);
        __syncwarp();
      }
    }
    __syncthreads();
  }
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
      if (c >= nc) break;
      const int s = seg_begin[c], e = seg_end[c];
      if (e <= s) continue;
      if (((e - 1) >> 5) - (s >> 5) + 1 > kCoopMinWords) continue;
      nms_segment_warp(s, e, cap, bv, alive, p.thr_eff);
    }
    __syncthreads();
    for (int c = 0; c < nc; ++c) {
      const int s = seg_begin[c], e = seg_end[c];
      if (e <= s) continue;
      if (((e - 1) >> 5) - (s >> 5) + 1 <= kCoopMinWords) continue;
      nms_segment_cta(s, e, cap, bv, alive, p.thr_eff, &sh_mask, &sh_kept);
    }
  }
  __syncthreads();

  // ---- ordered compaction of the survivors ----------------------------------------------------
  // The bitmask to compact is `alive` itself when its positions are already in output order
  // (trick: global score order; class-major output), otherwise the survivors are first scattered
  // into `galive`, indexed by global score rank.
  const bool by_rank = (p.order == CVPP_ORDER_SCORE_DESC) && !trick;
  if (by_rank) {
    for (int w = tid; w < nwords; w += kNmsThreads) {
      uint32_t m = alive[w];
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        const int r = wrank[(w << 5) + bit];
        atomicOr(&galive[r >> 5], 1u << (r & 31));
      }
    }
    __syncthreads();
  }
  const uint32_t* mask = by_rank ? galive : alive;
  const bool pos_is_rank = by_rank || trick;
  float4* ob = p.det_box + (int64_t)b * p.max_out;
  float* os = p.det_score + (int64_t)b * p.max_out;
  int32_t* oc = p.det_cls + (int64_t)b * p.max_out;
  int32_t* oa = p.det_anchor + (int64_t)b * p.max_out;
  const int out_cap = (p.order == CVPP_ORDER_SCORE_DESC && p.max_det > 0 && p.max_det < p.max_out) ? p.max_det : p.max_out;

  if (tid == 0) sh_running = 0;
  __syncthreads();
  for (int base = 0; base < nwords; base += kNmsThreads) {
    const int wi = base + tid;
    uint32_t m = wi < nwords ? mask[wi] : 0u;
    const int c = __popc(m);
    int incl = c;
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
    int pos = sh_running + woff + incl - c;
    while (m && pos < out_cap) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      const int at = (wi << 5) + bit;
      const uint64_t k = keys[pos_is_rank ? at : wrank[at]];
      const uint32_t anchor = (uint32_t)(k >> 12) & 0x1fffffu;
      ob[pos] = dense[anchor];
      os[pos] = __uint_as_float(0x7fffffffu - (uint32_t)(k >> 33));
      oc[pos] = (int32_t)(k & 0xfffu);
      oa[pos] = (int32_t)anchor;
      ++pos;
    }
    __syncthreads();
    if (tid == 0) sh_running += total;
    __syncthreads();
  }
  if (tid == 0) {
    int n_kept = sh_running;
    if (p.order == CVPP_ORDER_SCORE_DESC && p.max_det > 0 && n_kept > p.max_det) n_kept = p.max_det;
    p.det_count[b] = n_kept;
  }
}

size_t nms_workspace_bytes(int B, int max_cand) {
  size_t per = (size_t)max_cand * (sizeof(float4) + sizeof(float) + sizeof(int32_t));
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
  p.ws_rank = reinterpret_cast<int32_t*>(base + (size_t)B * max_cand * (sizeof(float4) + sizeof(float)));
  // shared memory: two bitmasks, class segments, per-warp class counters (<= 32 KB), then as many
  // box slots (24 B each) as fit, capped at max_cand
  p.alive_words = (max_cand + 31) / 32;
  int cw = (int)((32 * 1024) / ((size_t)nc * 4));
  if (cw > kNmsWarps) cw = kNmsWarps;
  if (cw < 1) cw = 1;
  p.count_warps = cw;
  size_t fixed = (size_t)p.alive_words * 8 + (size_t)nc * 8 + (size_t)cw * nc * 4 + 64;
  if (fixed + 4096 > (size_t)max_smem) {
    set_error("nms: max_cand=%d / nc=%d need more shared memory than the device has", max_cand, nc);
    return CVPP_ERR_UNSUPPORTED;
  }
  size_t avail = (size_t)max_smem - fixed - 1024;
  int smem_boxes = (int)(avail / 24);
  if (smem_boxes > max_cand) smem_boxes = max_cand;
  smem_boxes = (smem_boxes + 3) & ~3;
  if ((size_t)smem_boxes * 24 > avail) smem_boxes -= 4;
  p.smem_boxes = smem_boxes;
  size_t smem = fixed + (size_t)smem_boxes * 24;
  CVPP_CUDA_TRY(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_kernel<<<B, kNmsThreads, smem, stream>>>(p);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

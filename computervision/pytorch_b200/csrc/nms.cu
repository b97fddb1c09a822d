// nms.cu — kernel 3: class-aware greedy NMS on score-sorted candidate keys + gather of survivors (sm_100a).
//
// Replaces torchvision.ops.nms / batched_nms (torchvision/ops/boxes.py:20-120,
// csrc/ops/cpu/nms_kernel.cpp) at the reference call sites core/utils/ultralytics_ops.py:247-257,
// core/utils/nms.py:69,134, core/algorithms/yolo_v7.py:407, core/algorithms/ssd.py:267.
//
// One CTA per image.  Candidates arrive in global score order (cvpp_segmented_sort).
//   * a STABLE counting split by class (warp match_any + per-warp class counters + a scan) lays the
//     boxes out class-major in shared memory, each class still in score order — torchvision's
//     processing order — without a second sort; every class starts on a 32-position boundary so a
//     typical (~30 box) class is exactly one word;
//   * suppression state is a shared-memory bitmask (`alive`, one bit per class-major position);
//   * a 32-candidate word is resolved by one warp: lane i builds the bitmask of later boxes it would
//     suppress (independent IoU tests against boxes broadcast from shared memory), then the greedy
//     order is replayed with one shuffle + two logic ops per surviving box; survivors of a word are
//     applied to the later words of the segment (ballot collects the verdicts);
//   * small segments are handled one per warp (classes in parallel), large ones by the whole CTA
//     (word resolve by warp 0, application to later words spread over all warps);
//   * the suppression test is bit-identical to torchvision's CPU kernel, `(double)(inter / union) >
//     thr` with the fp32 division rounded to nearest, but evaluated without the division: it is
//     equivalent to comparing inter against (midpoint of thr_eff and the next float) * union, which
//     is exact in fp64 (24-bit x 25-bit significands); both batched_nms branches are implemented:
//     per-class on raw boxes, or class-agnostic on boxes shifted by cls * (max_coord + 1) (the
//     "coordinate trick", taken when an image has <= 1000 candidates);
//   * survivors are emitted class-major, or - through a second bitmask indexed by global score
//     rank - in score order capped at max_det.  Both are ordered popc/scan compactions.
// Latency / SM-bound; the only HBM traffic is 8 B keys + gathered 16 B boxes (L2 resident).
#include <cstring>

#include "cvpp_common.cuh"

namespace cvpp {

constexpr int kNmsThreads = 512;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kCoopMinWords = 8;  // segments spanning more words than this use the whole CTA

struct NmsParams {
  const uint64_t* sorted_key;
  const int32_t* cand_count;
  const float4* box_dense;
  int max_cand;
  int64_t A;
  int nc;
  float thr_eff;   // largest float <= iou_thres
  double thr_mid;  // midpoint of thr_eff and the next float above it
  int tie_up;      // a quotient exactly at thr_mid rounds (to even) to the float ABOVE thr_eff
  int rule;
  int order;
  int max_det;
  int max_out;
  float4* det_box;
  float* det_score;
  int32_t* det_cls;
  int32_t* det_anchor;
  int32_t* det_count;
  // scratch for images whose padded class-major layout does not fit shared memory
  float4* ws_box;    // [B][pos_cap]
  float* ws_area;    // [B][pos_cap]
  int32_t* ws_rank;  // [B][pos_cap] class-major position -> global score rank
  int pos_cap;       // max_cand + 32 * nc (class starts are 32-aligned)
  int smem_boxes;    // positions that fit in shared memory
  int mask_words;    // words reserved for each of the two bitmasks
  int count_warps;   // warps taking part in the counting split (per-warp class counters)
  const int32_t* skip;  // optional [B]: images the fused sort+NMS kernel already finished
};

struct SupTest {
  float thr_eff;
  double thr_mid;
  bool tie_up;
};

// class-major boxes + areas of one image: shared memory (explicit LDS) or a global scratch row
template <bool SMEM>
struct BoxView;
template <>
struct BoxView<true> {
  uint32_t box, area;  // shared-window byte addresses
  __device__ __forceinline__ float4 load_box(int i) const {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(box + 16u * (uint32_t)i));
    return v;
  }
  __device__ __forceinline__ float load_area(int i) const {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(area + 4u * (uint32_t)i));
    return v;
  }
};
template <>
struct BoxView<false> {
  const float4* box;
  const float* area;
  __device__ __forceinline__ float4 load_box(int i) const { return box[i]; }
  __device__ __forceinline__ float load_area(int i) const { return area[i]; }
};

// Is box j suppressed by kept box i (symmetric in i, j)?  torchvision: ovr = inter / (iarea + jarea -
// inter) in fp32, suppressed iff (double)ovr > thr.  RN(inter/uni) > thr_eff  <=>  inter/uni > thr_mid
// (>= when the tie rounds up), and inter vs thr_mid*uni is exact in fp64.  Degenerate unions (<= 0, inf,
// NaN) fall back to the division.  This is the exact (slow) form.
__device__ __noinline__ bool suppresses_exact(float inter, float uni, const SupTest& t) {
  const double lhs = (double)inter, rhs = t.thr_mid * (double)uni;
  const bool cmp = t.tie_up ? lhs >= rhs : lhs > rhs;
  const bool regular = uni > 0.0f && uni < INFINITY && inter < INFINITY;
  bool sup = inter > 0.0f && regular && cmp;  // inter == 0 / NaN: ovr is 0 or NaN, never above thr
  if (inter > 0.0f && !regular) sup = fdiv(inter, uni) > t.thr_eff;
  return sup;
}

__device__ __forceinline__ void inter_union(const float4& bi, float ai, const float4& bj, float aj, float& inter, float& uni) {
  const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
  const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
  const float w = fmaxf(0.0f, fsub(xx2, xx1));
  const float h = fmaxf(0.0f, fsub(yy2, yy1));
  inter = fmul(w, h);
  uni = fsub(fadd(ai, aj), inter);
}

// Fast classification of the same test in fp32: d = inter - thr_eff * uni (one rounding, fma).  The decision
// boundary inter = thr_mid * uni lies within thr_eff * 2^-24 * uni of d = 0, and the fma is off by at most
// 2^-24 |d|, so |d| > 2^-21 * uni decides the exact test with a wide margin: d above -> suppressed, below ->
// kept.  Anything else (near-threshold pairs, non-positive / non-finite unions, NaN) is `ambiguous` and must
// go through suppresses_exact.  A pair with inter == 0 is never suppressed (d = -thr_eff * uni <= 0).
__device__ __forceinline__ bool suppresses_fast(float inter, float uni, const SupTest& t, bool& ambiguous) {
  const float d = fmaf(-t.thr_eff, uni, inter);
  const float m = uni * 4.76837158203125e-07f;  // 2^-21
  const bool clear = fabsf(d) > m && m > 0.0f && fabsf(d) < INFINITY;  // false for NaN, uni <= 0, infinities
  ambiguous = !clear && inter > 0.0f;  // inter == 0 (or NaN): never suppressed
  return clear && d > 0.0f && inter > 0.0f;
}

// fast classification, exact test only for the (rare) ambiguous pairs
__device__ __forceinline__ bool suppresses(const float4& bi, float ai, const float4& bj, float aj, const SupTest& t) {
  float inter, uni;
  inter_union(bi, ai, bj, aj, inter, uni);
  bool amb;
  bool sup = suppresses_fast(inter, uni, t, amb);
  if (amb) sup = suppresses_exact(inter, uni, t);
  return sup;
}

// two independent pair tests in one basic block: a warp running these loops is a latency chain (shared-memory load -> min / max
// -> fma -> compares: ~0.1 instructions per cycle per warp, issue slots half idle), so interleaving two tests gives the
// scheduler two chains to overlap
__device__ __forceinline__ void suppresses_x2(const float4& a1, float aa1, const float4& b1, float ab1, const float4& a2, float aa2,
                                              const float4& b2, float ab2, const SupTest& t, bool& s1, bool& s2) {
  float i1, u1, i2, u2;
  inter_union(a1, aa1, b1, ab1, i1, u1);
  inter_union(a2, aa2, b2, ab2, i2, u2);
  bool m1, m2;
  s1 = suppresses_fast(i1, u1, t, m1);
  s2 = suppresses_fast(i2, u2, t, m2);
  if (m1) s1 = suppresses_exact(i1, u1, t);
  if (m2) s2 = suppresses_exact(i2, u2, t);
}

__device__ __forceinline__ uint32_t seg_mask_of_word(int w, int s, int e) {
  const int lo = max(s - (w << 5), 0), hi = min(e - (w << 5), 32);
  if (hi <= lo) return 0u;
  const uint32_t upto_hi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
  return upto_hi & ~((1u << lo) - 1u);
}

// keep only the lowest `room` set bits of m
__device__ __forceinline__ uint32_t lowest_bits(uint32_t m, int room) {
  if (__popc(m) <= room) return m;
  uint32_t o = 0;
  while (m && room > 0) {
    const uint32_t low = m & (0u - m);
    o |= low;
    m ^= low;
    --room;
  }
  return o;
}

constexpr int kPairMaxBoxes = 11;  // 55 pairs: two ballot words

// Resolve word w of segment [s, e): returns the kept mask (identical in all lanes).  All 32 lanes call.
// Pair (i, j) of the word is tested once, by lane i at rotation t = (j - i) mod 32, t = 1..16, so the
// 496 pairs cost 16 rounds of 32 parallel tests; a ballot per round routes each verdict to the
// column mask of the pair's EARLIER box.
template <bool SMEM>
__device__ __forceinline__ uint32_t resolve_word(int w, int s, int e, const BoxView<SMEM>& bv, const uint32_t* alive,
                                                 const SupTest& t) {
  const int lane = threadIdx.x & 31;
  const int base = w << 5;
  const uint32_t inmask = seg_mask_of_word(w, s, e);
  uint32_t am = alive[w] & inmask;
  if (am == 0 || (am & (am - 1)) == 0) return am;  // zero or one live box: nothing to resolve
  {
    // few live boxes in one contiguous run (the usual state of a freshly sorted small class): lane = PAIR.
    // The k (k - 1) / 2 <= 55 pairs are enumerated row-major over the upper triangle (row a holds b = a + 1 ..
    // k - 1), tested in at most two 32-lane rounds, and lane a cuts its column mask out of the ballot words.
    const int lo = __ffs(am) - 1, k = __popc(am);
    const uint32_t run = am >> lo;
    if (k <= kPairMaxBoxes && (run & (run + 1u)) == 0u) {
      const int P = (k * (k - 1)) >> 1;
      uint64_t ww = 0;
      for (int r = 0; (r << 5) < P; ++r) {
        const int q = lane + (r << 5);
        bool sup = false;
        if (q < P) {
          int a = 0, rem = q;
          while (rem >= k - 1 - a) {
            rem -= k - 1 - a;
            ++a;
          }
          const int i = base + lo + a, j = i + 1 + rem;
          sup = suppresses(bv.load_box(i), bv.load_area(i), bv.load_box(j), bv.load_area(j), t);
        }
        ww |= (uint64_t)__ballot_sync(0xffffffffu, sup) << (r << 5);
      }
      if (ww == 0ull) return am;  // nobody suppresses anybody
      uint32_t col = 0;
      if (lane < k - 1) {
        const int off = lane * (k - 1) - ((lane * (lane - 1)) >> 1);
        const int len = k - 1 - lane;
        col = ((uint32_t)(ww >> off) & ((1u << len) - 1u)) << (lo + lane + 1);
      }
      uint32_t it = __ballot_sync(0xffffffffu, col != 0u);  // rows (ranks) that suppress somebody
      while (it) {
        const int a = __ffs(it) - 1;
        it &= it - 1;
        const uint32_t c = __shfl_sync(0xffffffffu, col, a);
        if ((am >> (lo + a)) & 1u) am &= ~c;
      }
      return am;
    }
  }
  const float4 bi = bv.load_box(base + lane);
  const float ai = bv.load_area(base + lane);
  const bool me = (am >> lane) & 1u;
  // pairs are at most `span` lanes apart; rotations beyond min(span, 16) have nothing to test
  const int span = (31 - __clz(am)) - (__ffs(am) - 1);
  const int rounds = span < 16 ? span : 16;
  // lane i owns two bit rows: col = LATER live boxes of the word that box i suppresses (pairs it tested as
  // the earlier box), row = EARLIER boxes that suppress box i (pairs it tested as the later box, i.e. the
  // wrapped rotations).  The greedy replay below combines them, so no per-round routing of verdicts.
  uint32_t col = 0, row = 0;
  int r = 1;
  for (; r + 1 <= rounds; r += 2) {  // two rotations per iteration (independent tests)
    const int j1 = (lane + r) & 31, j2 = (lane + r + 1) & 31;
    const bool p1 = me && ((am >> j1) & 1u) && (r < 16 || lane < 16);
    const bool p2 = me && ((am >> j2) & 1u) && (r + 1 < 16 || lane < 16);
    bool s1 = false, s2 = false;
    if (p1 || p2) {
      suppresses_x2(bi, ai, bv.load_box(base + j1), bv.load_area(base + j1), bi, ai, bv.load_box(base + j2), bv.load_area(base + j2), t,
                    s1, s2);
      s1 = s1 && p1;
      s2 = s2 && p2;
    }
    const uint32_t bit1 = s1 ? (1u << j1) : 0u, bit2 = s2 ? (1u << j2) : 0u;
    if (j1 > lane) col |= bit1;
    else row |= bit1;
    if (j2 > lane) col |= bit2;
    else row |= bit2;
  }
  for (; r <= rounds; ++r) {
    const int j = (lane + r) & 31;
    const bool pair = me && ((am >> j) & 1u) && (r < 16 || lane < 16);
    bool sup = false;
    if (pair) sup = suppresses(bi, ai, bv.load_box(base + j), bv.load_area(base + j), t);
    const uint32_t bit = sup ? (1u << j) : 0u;
    if (j > lane) col |= bit;
    else row |= bit;
  }
  if (__ballot_sync(0xffffffffu, (col | row) != 0u) == 0u) return am;  // nobody suppresses anybody
  // replay the greedy order on bitmasks: a live box kills the later boxes in its column mask and every
  // later box whose row mask names it
  uint32_t rem = am;
  while (rem) {
    const int i = __ffs(rem) - 1;
    rem &= rem - 1;
    const uint32_t c = __shfl_sync(0xffffffffu, col, i) | __ballot_sync(0xffffffffu, (row >> i) & 1u);
    am &= ~c;
    rem &= ~c;
  }
  return am;
}

// Apply the kept boxes `am` of word w to word w2 of the same segment (all 32 lanes call).
// Lanes take the side with more boxes and loop over the other one, so a nearly empty word costs a
// few rounds instead of 32.
template <bool SMEM>
__device__ __forceinline__ void apply_word(uint32_t am, int w, int w2, int e, const BoxView<SMEM>& bv, uint32_t* alive,
                                           const SupTest& t) {
  const int lane = threadIdx.x & 31;
  const uint32_t live2 = alive[w2] & seg_mask_of_word(w2, 0, e);
  if (live2 == 0) return;
  uint32_t sm2 = 0;
  if (__popc(live2) >= __popc(am)) {
    // lane = candidate of w2, loop over the kept boxes of w
    const bool al2 = (live2 >> lane) & 1u;
    bool sup2 = false;
    if (al2) {
      const float4 b2 = bv.load_box((w2 << 5) + lane);
      const float a2 = bv.load_area((w2 << 5) + lane);
      uint32_t km = am;
      while (km && !sup2) {
        const int i1 = __ffs(km) - 1;
        km &= km - 1;
        if (km) {  // two kept boxes per iteration (independent tests)
          const int i2 = __ffs(km) - 1;
          km &= km - 1;
          bool s1, s2;
          suppresses_x2(bv.load_box((w << 5) + i1), bv.load_area((w << 5) + i1), b2, a2, bv.load_box((w << 5) + i2),
                        bv.load_area((w << 5) + i2), b2, a2, t, s1, s2);
          sup2 = s1 || s2;
        } else {
          sup2 = suppresses(bv.load_box((w << 5) + i1), bv.load_area((w << 5) + i1), b2, a2, t);
        }
      }
    }
    sm2 = __ballot_sync(0xffffffffu, sup2);
  } else {
    // lane = kept box of w, loop over the candidates of w2
    const bool kept = (am >> lane) & 1u;
    float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
    float ai = 0.f;
    if (kept) {
      bi = bv.load_box((w << 5) + lane);
      ai = bv.load_area((w << 5) + lane);
    }
    uint32_t lm = live2;
    while (lm) {
      const int j1 = __ffs(lm) - 1;
      lm &= lm - 1;
      if (lm) {  // two candidates per iteration (independent tests)
        const int j2 = __ffs(lm) - 1;
        lm &= lm - 1;
        bool s1 = false, s2 = false;
        if (kept)
          suppresses_x2(bi, ai, bv.load_box((w2 << 5) + j1), bv.load_area((w2 << 5) + j1), bi, ai, bv.load_box((w2 << 5) + j2),
                        bv.load_area((w2 << 5) + j2), t, s1, s2);
        if (__any_sync(0xffffffffu, s1)) sm2 |= 1u << j1;
        if (__any_sync(0xffffffffu, s2)) sm2 |= 1u << j2;
      } else {
        bool sup = false;
        if (kept) sup = suppresses(bi, ai, bv.load_box((w2 << 5) + j1), bv.load_area((w2 << 5) + j1), t);
        if (__any_sync(0xffffffffu, sup)) sm2 |= 1u << j1;
      }
    }
  }
  if (lane == 0 && sm2) atomicAnd(&alive[w2], ~sm2);
}

// One warp runs greedy NMS over segment [s, e).  cap > 0 stops after `cap` survivors.
template <bool SMEM>
__device__ void nms_segment_warp(int s, int e, int cap, const BoxView<SMEM>& bv, uint32_t* alive, const SupTest& t) {
  const int lane = threadIdx.x & 31;
  const int w0 = s >> 5, w1 = (e - 1) >> 5;
  int kept_total = 0;
  for (int w = w0; w <= w1; ++w) {
    uint32_t am = resolve_word(w, s, e, bv, alive, t);
    if (cap > 0) am = lowest_bits(am, cap - kept_total);
    kept_total += __popc(am);
    const uint32_t inmask = seg_mask_of_word(w, s, e);
    if (lane == 0) atomicAnd(&alive[w], ~(inmask & ~am));
    if (cap > 0 && kept_total >= cap) {
      for (int w2 = w + 1 + lane; w2 <= w1; w2 += 32) atomicAnd(&alive[w2], ~seg_mask_of_word(w2, s, e));
      break;
    }
    if (am)
      for (int w2 = w + 1; w2 <= w1; ++w2) apply_word(am, w, w2, e, bv, alive, t);
    __syncwarp();
  }
}

// The whole CTA runs greedy NMS over one (large) segment.  All threads call.
template <bool SMEM>
__device__ void nms_segment_cta(int s, int e, int cap, const BoxView<SMEM>& bv, uint32_t* alive, const SupTest& t,
                                uint32_t* sh_mask, int* sh_kept) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w0 = s >> 5, w1 = (e - 1) >> 5;
  if (threadIdx.x == 0) *sh_kept = 0;
  __syncthreads();
  for (int w = w0; w <= w1; ++w) {
    if (warp == 0) {
      uint32_t am = resolve_word(w, s, e, bv, alive, t);
      const int kept_total = *sh_kept;
      if (cap > 0) am = lowest_bits(am, cap - kept_total);
      const uint32_t inmask = seg_mask_of_word(w, s, e);
      __syncwarp();
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
      for (int w2 = w + 1 + warp; w2 <= w1; w2 += (int)(blockDim.x >> 5)) apply_word(am, w, w2, e, bv, alive, t);
    }
    __syncthreads();
    if (done) break;
  }
}

// greedy suppression over every segment of the image (all threads call)
template <bool SMEM>
__device__ void suppress_all(const BoxView<SMEM>& bv, bool trick, int n, int nwords, int nc, int cap, uint32_t* alive,
                             const int* seg_begin, const int* seg_end, const SupTest& st, int* sh_next,
                             uint32_t* sh_mask, int* sh_kept, bool any_large = true) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (trick) {
    if (nwords > kCoopMinWords) {
      nms_segment_cta(0, n, cap, bv, alive, st, sh_mask, sh_kept);
    } else {
      if (warp == 0) nms_segment_warp(0, n, cap, bv, alive, st);
      __syncthreads();
    }
    return;
  }
  // small segments: one warp each, classes claimed dynamically; large ones afterwards by the CTA
  for (;;) {
    int c = 0;
    if (lane == 0) c = atomicAdd(sh_next, 1);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= nc) break;
    const int s = seg_begin[c], e = seg_end[c];
    if (e <= s) continue;
    if (((e - 1) >> 5) - (s >> 5) + 1 > kCoopMinWords) continue;
    nms_segment_warp(s, e, cap, bv, alive, st);
  }
  __syncthreads();
  if (!any_large) return;  // (a scan over all classes by every thread costs ~5 us at nc = 80)
  for (int c = 0; c < nc; ++c) {
    const int s = seg_begin[c], e = seg_end[c];
    if (e <= s) continue;
    if (((e - 1) >> 5) - (s >> 5) + 1 <= kCoopMinWords) continue;
    nms_segment_cta(s, e, cap, bv, alive, st, sh_mask, sh_kept);
  }
}

// The body of nms_kernel for a CTA of T threads working on image b (also the in-kernel fallback of nms2_kernel).
template <int T>
__device__ __forceinline__ void nms_body(const NmsParams& p, const int b, unsigned char* smem_raw) {
  constexpr int kThreads = T, kWarps = T / 32;
  // layout: [boxes float4 x N][areas float x N][rank int x N][alive u32 x W][galive u32 x W]
  //         [seg_begin int x nc][seg_end int x nc][cnt int x count_warps x nc]
  float4* sh_box = reinterpret_cast<float4*>(smem_raw);
  float* sh_area = reinterpret_cast<float*>(sh_box + p.smem_boxes);
  int32_t* sh_rank = reinterpret_cast<int32_t*>(sh_area + p.smem_boxes);
  uint32_t* alive = reinterpret_cast<uint32_t*>(sh_rank + p.smem_boxes);
  uint32_t* galive = alive + p.mask_words;
  int* seg_begin = reinterpret_cast<int*>(galive + p.mask_words);
  int* seg_end = seg_begin + p.nc;
  int* cnt = seg_end + p.nc;
  __shared__ float sh_red[kWarps];
  __shared__ int sh_scan[kWarps];
  __shared__ int sh_running;
  __shared__ uint32_t sh_mask;
  __shared__ int sh_kept;
  __shared__ int sh_next;
  __shared__ int sh_npos;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (p.skip && p.skip[b]) return;  // already done by the fused kernel
  int n = p.cand_count[b];
  if (n > p.max_cand) n = p.max_cand;
  if (n <= 0) {
    if (tid == 0) p.det_count[b] = 0;
    return;
  }
  const uint64_t* keys = p.sorted_key + (int64_t)b * p.max_cand;  // score-major: [inv_score | anchor | class]
  const float4* dense = p.box_dense + (int64_t)b * p.A;
  const bool trick = rule_uses_trick(p.rule, n);
  const int nc = p.nc;
  const SupTest st{p.thr_eff, p.thr_mid, p.tie_up != 0};
  const int rank_words = (n + 31) >> 5;

  float4* wbox;
  float* warea;
  int32_t* wrank;
  int npos;  // size of the position space (== n for the trick branch, 32-aligned class segments otherwise)

  for (int w = tid; w < p.mask_words; w += kThreads) {
    alive[w] = 0u;
    galive[w] = 0u;
  }
  if (tid == 0) sh_next = 0;
  __syncthreads();

  if (trick) {
    // ---- coordinate trick: one class-agnostic segment in score order; boxes shifted by
    //      cls * (max over every coordinate + 1)  (boxes.py:95-97)
    npos = n;
    const bool in_smem = npos <= p.smem_boxes;
    wbox = in_smem ? sh_box : p.ws_box + (int64_t)b * p.pos_cap;
    warea = in_smem ? sh_area : p.ws_area + (int64_t)b * p.pos_cap;
    wrank = in_smem ? sh_rank : p.ws_rank + (int64_t)b * p.pos_cap;
    float m = -INFINITY;
    for (int r = tid; r < n; r += kThreads) {
      const float4 bx = dense[(uint32_t)(keys[r] >> 12) & 0x1fffffu];
      m = fmaxf(m, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
    }
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) sh_red[warp] = m;
    __syncthreads();
    m = sh_red[0];
    for (int q = 1; q < kWarps; ++q) m = fmaxf(m, sh_red[q]);
    const float mult = fadd(m, 1.0f);
    for (int r = tid; r < n; r += kThreads) {
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
    for (int w = tid; w < rank_words; w += kThreads) {
      const int rem = n - (w << 5);
      alive[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    }
    __syncthreads();
  } else {
    // ---- stable counting split by class: position = class start + #earlier ranks of that class
    const int cw = p.count_warps;
    for (int i = tid; i < cw * nc; i += kThreads) cnt[i] = 0;
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
    for (int c = tid; c < nc; c += kThreads) {
      int run = 0;
      for (int w = 0; w < cw; ++w) {
        const int t = cnt[w * nc + c];
        cnt[w * nc + c] = run;
        run += t;
      }
      seg_end[c] = run;
    }
    __syncthreads();
    if (warp == 0) {  // scan of the 32-aligned class sizes -> segment starts
      int running = 0;
      for (int c0 = 0; c0 < nc; c0 += 32) {
        const int c = c0 + lane;
        const int t = c < nc ? seg_end[c] : 0;
        const int ta = (t + 31) & ~31;
        int incl = ta;
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += v;
        }
        if (c < nc) {
          const int s0 = running + incl - ta;
          seg_begin[c] = s0;
          seg_end[c] = s0 + t;
          for (int w = 0; w < (ta >> 5); ++w) {  // live bits of this class
            const int rem = t - (w << 5);
            alive[(s0 >> 5) + w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
          }
        }
        running += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) sh_npos = running;
    }
    __syncthreads();
    npos = sh_npos;
    const bool in_smem = npos <= p.smem_boxes;
    wbox = in_smem ? sh_box : p.ws_box + (int64_t)b * p.pos_cap;
    warea = in_smem ? sh_area : p.ws_area + (int64_t)b * p.pos_cap;
    wrank = in_smem ? sh_rank : p.ws_rank + (int64_t)b * p.pos_cap;
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
  const int nwords = (npos + 31) >> 5;

  // ---- greedy suppression -------------------------------------------------------------------
  const int cap = (p.order == CVPP_ORDER_SCORE_DESC && p.max_det > 0) ? p.max_det : 0;
  if (npos <= p.smem_boxes) {
    const BoxView<true> bv{smem_u32(sh_box), smem_u32(sh_area)};
    suppress_all(bv, trick, n, nwords, nc, cap, alive, seg_begin, seg_end, st, &sh_next, &sh_mask, &sh_kept);
  } else {
    const BoxView<false> bv{wbox, warea};
    suppress_all(bv, trick, n, nwords, nc, cap, alive, seg_begin, seg_end, st, &sh_next, &sh_mask, &sh_kept);
  }
  __syncthreads();

  // ---- ordered compaction of the survivors ----------------------------------------------------
  // The bitmask to compact is `alive` itself when its positions are already in output order
  // (trick: global score order; class-major output), otherwise the survivors are first scattered
  // into `galive`, indexed by global score rank.
  const bool by_rank = (p.order == CVPP_ORDER_SCORE_DESC) && !trick;
  if (by_rank) {
    for (int w = tid; w < nwords; w += kThreads) {
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
  const int mwords = by_rank ? rank_words : nwords;
  const bool pos_is_rank = by_rank || trick;
  float4* ob = p.det_box + (int64_t)b * p.max_out;
  float* os = p.det_score + (int64_t)b * p.max_out;
  int32_t* oc = p.det_cls + (int64_t)b * p.max_out;
  int32_t* oa = p.det_anchor + (int64_t)b * p.max_out;
  const int out_cap = (p.order == CVPP_ORDER_SCORE_DESC && p.max_det > 0 && p.max_det < p.max_out) ? p.max_det : p.max_out;

  // pass 1: ordered positions -> index list (the area array is dead now and is reused for it)
  int* out_idx = reinterpret_cast<int*>(warea);
  if (tid == 0) sh_running = 0;
  __syncthreads();
  for (int base = 0; base < mwords; base += kThreads) {
    const int wi = base + tid;
    uint32_t m = wi < mwords ? mask[wi] : 0u;
    const int c = __popc(m);
    int incl = c;
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) sh_scan[warp] = incl;
    __syncthreads();
    int woff = 0, total = 0;
    for (int q = 0; q < kWarps; ++q) {
      const int v = sh_scan[q];
      if (q < warp) woff += v;
      total += v;
    }
    int pos = sh_running + woff + incl - c;
    while (m && pos < out_cap) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      out_idx[pos++] = (wi << 5) + bit;
    }
    __syncthreads();
    if (tid == 0) sh_running += total;
    __syncthreads();
  }
  int n_kept = sh_running;
  if (p.order == CVPP_ORDER_SCORE_DESC && p.max_det > 0 && n_kept > p.max_det) n_kept = p.max_det;
  // pass 2: one output row per thread, all gathers in flight at once
  const int n_rows = min(n_kept, out_cap);
  for (int o = tid; o < n_rows; o += kThreads) {
    const int at = out_idx[o];
    const uint64_t k = keys[pos_is_rank ? at : wrank[at]];
    const uint32_t anchor = (uint32_t)(k >> 12) & 0x1fffffu;
    ob[o] = dense[anchor];
    os[o] = __uint_as_float(0x7fffffffu - (uint32_t)(k >> 33));
    oc[o] = (int32_t)(k & 0xfffu);
    oa[o] = (int32_t)anchor;
  }
  if (tid == 0) p.det_count[b] = n_kept;
}

__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(const __grid_constant__ NmsParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  nms_body<kNmsThreads>(p, (int)blockIdx.x, smem_raw);
}

static size_t pos_capacity(int max_cand, int nc) { return (size_t)max_cand + 32u * (size_t)nc; }

size_t nms_workspace_bytes(int B, int max_cand, int nc) {
  size_t per = pos_capacity(max_cand, nc) * (sizeof(float4) + sizeof(float) + sizeof(int32_t));
  per = (per + 255) & ~(size_t)255;
  return per * (size_t)B + 256;
}

int nms_launch_skip(const uint64_t* sorted_key, const int32_t* cand_count, const float* box_dense, int B, int max_cand,
                    int64_t A, int nc, double iou_thres, int rule, int order, int max_det, int max_out, float* det_box,
                    float* det_score, int32_t* det_cls, int32_t* det_anchor, int32_t* det_count, void* workspace,
                    size_t workspace_bytes, const int32_t* skip, cudaStream_t stream);

int nms_launch(const uint64_t* sorted_key, const int32_t* cand_count, const float* box_dense, int B, int max_cand,
               int64_t A, int nc, double iou_thres, int rule, int order, int max_det, int max_out, float* det_box,
               float* det_score, int32_t* det_cls, int32_t* det_anchor, int32_t* det_count, void* workspace,
               size_t workspace_bytes, cudaStream_t stream) {
  return nms_launch_skip(sorted_key, cand_count, box_dense, B, max_cand, A, nc, iou_thres, rule, order, max_det, max_out,
                         det_box, det_score, det_cls, det_anchor, det_count, workspace, workspace_bytes, nullptr, stream);
}

// Validates the arguments, fills the kernel parameters and reports the dynamic shared memory nms_body needs.
static int nms_prepare(NmsParams& p, size_t& smem, const uint64_t* sorted_key, const int32_t* cand_count,
                       const float* box_dense, int B, int max_cand, int64_t A, int nc, double iou_thres, int rule, int order,
                       int max_det, int max_out, float* det_box, float* det_score, int32_t* det_cls, int32_t* det_anchor,
                       int32_t* det_count, void* workspace, size_t workspace_bytes, const int32_t* skip, int max_smem) {
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
  if (rule < 0 || rule > 3 || order < 0 || order > 1) {
    set_error("nms: unknown rule %d / order %d", rule, order);
    return CVPP_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(box_dense) & 15u) || (reinterpret_cast<uintptr_t>(det_box) & 15u)) {
    set_error("nms: box arrays must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  if (!workspace || workspace_bytes < nms_workspace_bytes(B, max_cand, nc)) {
    set_error("nms: workspace of %zu bytes needed, got %zu", nms_workspace_bytes(B, max_cand, nc), workspace_bytes);
    return CVPP_ERR_WORKSPACE;
  }
  p = NmsParams{};
  // (double)ovr > thr  <=>  ovr > thr_eff := largest float <= thr   (ovr is a float)
  float thr_eff = (float)iou_thres;
  if ((double)thr_eff > iou_thres) thr_eff = nextafterf(thr_eff, -INFINITY);
  const float thr_next = nextafterf(thr_eff, INFINITY);
  p.thr_eff = thr_eff;
  p.thr_mid = ((double)thr_eff + (double)thr_next) * 0.5;  // exact: 25 significant bits
  uint32_t next_bits;
  memcpy(&next_bits, &thr_next, sizeof(next_bits));
  p.tie_up = (next_bits & 1u) == 0;  // round-half-to-even picks thr_next when its significand is even
  p.sorted_key = sorted_key;
  p.cand_count = cand_count;
  p.box_dense = reinterpret_cast<const float4*>(box_dense);
  p.max_cand = max_cand;
  p.A = A;
  p.nc = nc;
  p.rule = rule;
  p.order = order;
  p.max_det = max_det;
  p.max_out = max_out;
  p.det_box = reinterpret_cast<float4*>(det_box);
  p.det_score = det_score;
  p.det_cls = det_cls;
  p.det_anchor = det_anchor;
  p.det_count = det_count;
  p.skip = skip;
  // workspace carve-up
  const size_t cap = pos_capacity(max_cand, nc);
  p.pos_cap = (int)cap;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  p.ws_box = reinterpret_cast<float4*>(base);
  p.ws_area = reinterpret_cast<float*>(base + (size_t)B * cap * sizeof(float4));
  p.ws_rank = reinterpret_cast<int32_t*>(base + (size_t)B * cap * (sizeof(float4) + sizeof(float)));
  // shared memory: two bitmasks, class segments, per-warp class counters (<= 32 KB), then as many
  // box slots (24 B each) as fit, capped at the padded position capacity
  p.mask_words = (int)((cap + 31) / 32);
  int cw = (int)((32 * 1024) / ((size_t)nc * 4));
  if (cw > kNmsWarps) cw = kNmsWarps;
  if (cw < 1) cw = 1;
  p.count_warps = cw;
  size_t fixed = (size_t)p.mask_words * 8 + (size_t)nc * 8 + (size_t)cw * nc * 4 + 64;
  if (fixed + 4096 > (size_t)max_smem) {
    set_error("nms: max_cand=%d / nc=%d need more shared memory than the device has", max_cand, nc);
    return CVPP_ERR_UNSUPPORTED;
  }
  size_t avail = (size_t)max_smem - fixed - 1024;
  size_t smem_boxes = avail / 24;
  if (smem_boxes > cap) smem_boxes = cap;
  smem_boxes &= ~(size_t)3;
  p.smem_boxes = (int)smem_boxes;
  smem = fixed + smem_boxes * 24;
  return CVPP_OK;
}

int nms_launch_skip(const uint64_t* sorted_key, const int32_t* cand_count, const float* box_dense, int B, int max_cand,
                    int64_t A, int nc, double iou_thres, int rule, int order, int max_det, int max_out, float* det_box,
                    float* det_score, int32_t* det_cls, int32_t* det_anchor, int32_t* det_count, void* workspace,
                    size_t workspace_bytes, const int32_t* skip, cudaStream_t stream) {
  if (B == 0) return CVPP_OK;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != CVPP_OK) return rc;
  NmsParams p;
  size_t smem = 0;
  rc = nms_prepare(p, smem, sorted_key, cand_count, box_dense, B, max_cand, A, nc, iou_thres, rule, order, max_det, max_out,
                   det_box, det_score, det_cls, det_anchor, det_count, workspace, workspace_bytes, skip, di.max_smem);
  if (rc != CVPP_OK) return rc;
  static unsigned long long attr_done = 0;
  static int attr_bytes = 0;  // the attribute must cover the largest request seen so far
  if ((int)smem > attr_bytes) {
    attr_done = 0;
    attr_bytes = (int)smem;
  }
  rc = ensure_smem_attr(reinterpret_cast<const void*>(nms_kernel), attr_bytes, di.device, &attr_done);
  if (rc != CVPP_OK) return rc;
  nms_kernel<<<B, kNmsThreads, smem, stream>>>(p);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

namespace cvpp {

// =====================================================================================================
// Fused sort + NMS (v2): one kernel per image, no global sort.
//
// The separate segmented sort (a full bitonic sort of ~2 500 keys padded to 4 096, 28 us for 64 images)
// plus nms_kernel (66 us) made the latency-bound stages longer than the HBM-bound decode (70 us).  Only two
// orders are actually needed: score order WITHIN a class (torchvision's processing order) and the global
// score order of the max_det SURVIVORS.  So:
//   1. histogram of the (unsorted) candidate keys by class, 32-aligned class segments (shared-memory atomics);
//   2. scatter into the segments, then every class segment is sorted on its own - a typical ~30-box class is
//      ONE warp-register bitonic sort of 32 keys (15 shuffle steps), classes run in parallel over 32 warps;
//      the warp that sorted a class also gathers its boxes/areas into shared memory;
//   3. greedy suppression exactly as in nms_kernel (same device functions, same bit-exact IoU test);
//   4. class-major output: ordered compaction of the alive bitmask; score-ordered output: histogram of the
//      survivors' score bits (4 096 bins), the bin where the cumulative count reaches max_det, and a sort of
//      only the selected <= max_det + boundary-bin keys.
// The coordinate-trick branch (<= 1000 candidates) needs the global score order for its class-agnostic pass
// and sorts its <= 1 024 keys in shared memory.  Images that do not fit shared memory (or exceed max_nms) take
// the slow path inside the same CTA (nms2_fallback: global score-order sort + the body of nms_kernel), so the
// whole stage is always ONE launch.
// =====================================================================================================
constexpr int kN2Threads = 1024;
constexpr int kN2Warps = kN2Threads / 32;
constexpr int kN2Bins = 4096;
constexpr int kN2DirectSel = 1024;   // survivors up to this count skip the bin selection: one block sort
constexpr int kN2SmallMax = 256;     // coordinate-trick images up to this size are handled by one warp
constexpr int kN2WarpSortMax = 256;  // keys per class a single warp sorts in registers (E <= 8)

struct Nms2Params {
  const uint64_t* cand_key;  // UNSORTED class-major keys
  const int32_t* cand_count;
  const float4* box_dense;
  int max_cand;
  int64_t A;
  int nc;
  float thr_eff;
  double thr_mid;
  int tie_up;
  int rule, order, max_det, max_nms, max_out;
  float4* det_box;
  float* det_score;
  int32_t* det_cls;
  int32_t* det_anchor;
  int32_t* det_count;
  int32_t* cand_count_out;  // optional [B]: copy of cand_count (saves the caller a device-to-device copy node)
  // optional fused row epilogue + all-gather (the CVPP_ROWS_FULL / CVPP_BOX_KEEP form of detection_epilogue_kernel, done
  // by the image's own CTA once its detections are written: the step has no third launch).  Destination layout as in
  // cvpp_detection_epilogue_allgather: [n_ranks][B][max_out][7] rows | [n_ranks][B] counts, this rank at slot gather_rank.
  float* gather_dst[CVPP_MAX_PEERS];  // unicast form: every rank's buffer (peer mappings)
  float* gather_mc;                   // multicast form: the NVSwitch multicast address of the buffer (one multimem.st)
  int gather_n, gather_rank;          // gather_n = 0: no fused gather
  int cap_pos;     // class-major positions that fit shared memory (multiple of 32)
  int mask_words;  // cap_pos / 32
  // in-kernel fallback (images that do not fit shared memory, or with more than max_nms candidates): global
  // score-order sort into fb_keys, then the body of nms_kernel on them
  NmsParams fb;
  uint64_t* fb_keys;    // [B][max_cand]
  int32_t* fb_count;    // [B]
  uint64_t* fb_sort_ws; // [B][fb_sort_stride] scratch for sorts larger than shared memory
  int64_t fb_sort_stride;
  int fb_smem_keys;     // keys that fit the dynamic shared memory for the fallback sort (power of two)
};

template <int E>
__device__ __forceinline__ void warp_sort_segment(uint64_t* seg, int t) {
  const int lane = threadIdx.x & 31;
  uint64_t x[E];
#pragma unroll
  for (int e = 0; e < E; ++e) x[e] = (e * 32 + lane) < t ? seg[e * 32 + lane] : ~0ull;
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1) warp_bitonic_steps<E>(x, 0, k, k >> 1);
#pragma unroll
  for (int e = 0; e < E; ++e)
    if ((e * 32 + lane) < t) seg[e * 32 + lane] = x[e];
  __syncwarp();
}

// t <= 32 distinct keys: lane i counts the keys below its own (broadcast shared-memory reads) and stores its
// key at that rank - about 4 t instructions instead of the 15-stage shuffle network (classes of the score
// prefix hold ~8 keys).
__device__ __forceinline__ void warp_rank_sort_32(uint64_t* seg, int t) {
  const int lane = threadIdx.x & 31;
  const uint64_t mine = lane < t ? seg[lane] : ~0ull;
  int rank = 0;
  for (int j = 0; j < t; ++j) rank += seg[j] < mine ? 1 : 0;
  __syncwarp();
  if (lane < t) seg[rank] = mine;
  __syncwarp();
}

__device__ __forceinline__ void block_sort_smem_max8(uint64_t* a, int P) {
  const int T = blockDim.x;
  const int E = P > T ? P / T : 1;
  switch (E) {
    case 1: block_sort_smem_e<1>(a, P); break;
    case 2: block_sort_smem_e<2>(a, P); break;
    case 4: block_sort_smem_e<4>(a, P); break;
    case 8: block_sort_smem_e<8>(a, P); break;
    default: bitonic_sort_u64_generic(a, P); break;
  }
}

// exclusive block scan helper over one int per thread (blockDim = kN2Threads); returns the exclusive prefix,
// *total receives the block total.  Two __syncthreads.
__device__ __forceinline__ int block_excl_scan(int v, int* sh_warp /*[kN2Warps]*/, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += u;
  }
  if (lane == 31) sh_warp[warp] = incl;
  __syncthreads();
  int woff = 0, tot = 0;
  for (int q = 0; q < kN2Warps; ++q) {
    const int u = sh_warp[q];
    if (q < warp) woff += u;
    tot += u;
  }
  __syncthreads();
  *total = tot;
  return woff + incl - v;
}

// The slow path of nms2_kernel, all threads of the CTA: what cvpp_segmented_sort + cvpp_nms do for one image.
__device__ __noinline__ void nms2_fallback(const Nms2Params& p, int b, unsigned char* smem_raw) {
  int n = p.cand_count[b];
  if (n > p.max_cand) n = p.max_cand;
  const uint64_t* kb = p.cand_key + (int64_t)b * p.max_cand;
  uint64_t* ko = p.fb_keys + (int64_t)b * p.max_cand;
  const int P = pow2_ceil(n < 32 ? 32 : n);
  const bool in_smem = P <= p.fb_smem_keys;
  uint64_t* a = in_smem ? reinterpret_cast<uint64_t*>(smem_raw) : p.fb_sort_ws + (int64_t)b * p.fb_sort_stride;
  for (int i = threadIdx.x; i < P; i += blockDim.x) a[i] = i < n ? key_to_score_major(kb[i]) : ~0ull;
  __syncthreads();
  if (in_smem)
    block_sort_smem_max8(a, P);
  else
    bitonic_sort_u64_generic(a, P);
  const int n_final = (p.max_nms > 0 && n > p.max_nms) ? p.max_nms : n;  // ultralytics_ops.py:240
  for (int i = threadIdx.x; i < n_final; i += blockDim.x) ko[i] = a[i];
  if (threadIdx.x == 0) p.fb_count[b] = n_final;
  __threadfence_block();
  __syncthreads();
  nms_body<kN2Threads>(p.fb, b, smem_raw);
}

#ifdef CVPP_NMS_TIMING
__device__ long long g_n2_t[16 * 64];
#define N2_MARK(i) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x < 64) g_n2_t[blockIdx.x * 16 + (i)] = clock64(); } while (0)
#else
#define N2_MARK(i) do {} while (0)
#endif

__device__ __forceinline__ void nms2_image(const Nms2Params& p, unsigned char* smem_raw) {
  // layout: [box float4 x cap][key u64 x cap][area f32 x cap][alive u32 x words][seg_begin, seg_end, fill int x nc]
  float4* sh_box = reinterpret_cast<float4*>(smem_raw);
  uint64_t* ckey = reinterpret_cast<uint64_t*>(sh_box + p.cap_pos);
  float* sh_area = reinterpret_cast<float*>(ckey + p.cap_pos);
  uint32_t* alive = reinterpret_cast<uint32_t*>(sh_area + p.cap_pos);
  int* seg_begin = reinterpret_cast<int*>(alive + p.mask_words);
  int* seg_end = seg_begin + p.nc;
  int* fill = seg_end + p.nc;
  __shared__ float sh_red[kN2Warps];
  __shared__ int sh_scan[kN2Warps];
  __shared__ uint32_t sh_mask;
  __shared__ int sh_kept, sh_next, sh_npos, sh_i0, sh_maxt;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  // launched with programmatic stream serialisation: nothing the producer kernel wrote is read above this line; the
  // kernel that follows (the row epilogue) may be scheduled as soon as every CTA of this grid is resident
  pdl_trigger();
  pdl_wait();
  int n = p.cand_count[b];
  if (tid == 0 && p.cand_count_out) p.cand_count_out[b] = n;
  if (n > p.max_cand) n = p.max_cand;
  if (n <= 0) {
    if (tid == 0) p.det_count[b] = 0;
    return;
  }
  const uint64_t* keys = p.cand_key + (int64_t)b * p.max_cand;
  const float4* dense = p.box_dense + (int64_t)b * p.A;
  const bool trick = rule_uses_trick(p.rule, n);
  const int nc = p.nc;
  const SupTest st{p.thr_eff, p.thr_mid, p.tie_up != 0};
  const int cap = (p.order == CVPP_ORDER_SCORE_DESC && p.max_det > 0) ? p.max_det : 0;
  bool fits = p.cap_pos > 0 && !(p.max_nms > 0 && n > p.max_nms);
  if (trick && pow2_ceil(n < 32 ? 32 : n) > p.cap_pos) fits = false;
  if (!fits) {  // uniform
    nms2_fallback(p, b, smem_raw);
    return;
  }
  if (trick && n <= kN2SmallMax) {
    // ---- few candidates on the coordinate-trick branch (the bs = 1, conf .25 case): ONE warp does the whole
    //      image - sort, shifted boxes, greedy pass, ordered output - without a single CTA barrier
    if (warp != 0) return;
    for (int i = lane; i < n; i += 32) ckey[i] = key_to_score_major(keys[i]);
    __syncwarp();
    if (n <= 32) warp_sort_segment<1>(ckey, n);
    else if (n <= 64) warp_sort_segment<2>(ckey, n);
    else if (n <= 128) warp_sort_segment<4>(ckey, n);
    else warp_sort_segment<8>(ckey, n);
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) {
      const float4 bx = dense[(uint32_t)(ckey[i] >> 12) & 0x1fffffu];
      sh_box[i] = bx;
      m = fmaxf(m, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
    }
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    const float mult = fadd(m, 1.0f);  // boxes.py:95-97: offsets = idxs * (max_coordinate + 1)
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const float off = fmul((float)(uint32_t)(ckey[i] & 0xfffu), mult);
      float4 bx = sh_box[i];
      bx.x = fadd(bx.x, off);
      bx.y = fadd(bx.y, off);
      bx.z = fadd(bx.z, off);
      bx.w = fadd(bx.w, off);
      sh_box[i] = bx;
      sh_area[i] = box_area(bx);
    }
    const int words = (n + 31) >> 5;
    for (int w = lane; w < words; w += 32) {
      const int rem = n - (w << 5);
      alive[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    }
    __syncwarp();
    {
      const BoxView<true> bv{smem_u32(sh_box), smem_u32(sh_area)};
      nms_segment_warp(0, n, cap, bv, alive, st);
    }
    __syncwarp();
    float4* ob = p.det_box + (int64_t)b * p.max_out;
    float* os = p.det_score + (int64_t)b * p.max_out;
    int32_t* oc = p.det_cls + (int64_t)b * p.max_out;
    int32_t* oa = p.det_anchor + (int64_t)b * p.max_out;
    int running = 0;
    for (int w = 0; w < words; ++w) {
      const uint32_t mw = alive[w];
      if ((mw >> lane) & 1u) {
        const int o = running + __popc(mw & ((1u << lane) - 1u));
        if (o < p.max_out) {
          const uint64_t k = ckey[(w << 5) + lane];
          const uint32_t anchor = (uint32_t)(k >> 12) & 0x1fffffu;
          ob[o] = dense[anchor];
          os[o] = __uint_as_float(0x7fffffffu - (uint32_t)(k >> 33));
          oc[o] = (int32_t)(k & 0xfffu);
          oa[o] = (int32_t)anchor;
        }
      }
      running += __popc(mw);
    }
    if (lane == 0) p.det_count[b] = (cap > 0 && running > cap) ? cap : running;
    return;
  }

  N2_MARK(0);
  int npos = 0, nwords = 0, n_surv = 0;
  // attempt 0: score-prefix subset (only when the output is score-ordered, capped, and the image is large);
  // attempt 1: every candidate
  int attempt = (!trick && cap > 0 && n > 2 * cap + 256) ? 0 : 1;
  int tb = kN2Bins - 1;  // candidates with score bin <= tb take part
  for (;;) {
  for (int w = tid; w < p.mask_words; w += kN2Threads) alive[w] = 0u;
  if (tid == 0) sh_next = 0;
  __syncthreads();

  if (trick) {
    // ---- coordinate trick: global score order, one class-agnostic segment on shifted boxes (boxes.py:95-97)
    const int P = pow2_ceil(n < 32 ? 32 : n);
    for (int i = tid; i < P; i += kN2Threads) ckey[i] = i < n ? key_to_score_major(keys[i]) : ~0ull;
    __syncthreads();
    block_sort_smem_max8(ckey, P);
    npos = n;
    float m = -INFINITY;
    for (int r = tid; r < n; r += kN2Threads) {
      const float4 bx = dense[(uint32_t)(ckey[r] >> 12) & 0x1fffffu];
      sh_box[r] = bx;
      m = fmaxf(m, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
    }
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) sh_red[warp] = m;
    __syncthreads();
    m = sh_red[0];
    for (int q = 1; q < kN2Warps; ++q) m = fmaxf(m, sh_red[q]);
    const float mult = fadd(m, 1.0f);
    for (int r = tid; r < n; r += kN2Threads) {
      const float off = fmul((float)(uint32_t)(ckey[r] & 0xfffu), mult);
      float4 bx = sh_box[r];
      bx.x = fadd(bx.x, off);
      bx.y = fadd(bx.y, off);
      bx.z = fadd(bx.z, off);
      bx.w = fadd(bx.w, off);
      sh_box[r] = bx;
      sh_area[r] = box_area(bx);
    }
    for (int w = tid; w < ((n + 31) >> 5); w += kN2Threads) {
      const int rem = n - (w << 5);
      alive[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    }
    __syncthreads();
  } else {
    if (attempt == 0) {
      // ---- score-prefix shortcut (score-ordered output capped at max_det): a candidate's fate depends only on
      //      better-scored candidates of its class, so greedy NMS restricted to "every candidate with score bin
      //      <= tb" yields exactly the true survivors of that score range.  If at least max_det of them
      //      survive, they contain the global top max_det and the remaining ~80 % of the candidates never
      //      need to be sorted, gathered or tested; otherwise the image is redone in full (attempt 1).
      int* hist = reinterpret_cast<int*>(sh_box);  // boxes are not loaded yet
      for (int i = tid; i < kN2Bins; i += kN2Threads) hist[i] = 0;
      if (tid == 0) sh_i0 = kN2Bins - 1;
      __syncthreads();
      for (int r = tid; r < n; r += kN2Threads) atomicAdd(&hist[(int)((keys[r] >> 40) & 0xfffu)], 1);
      __syncthreads();
      constexpr int kPer = kN2Bins / kN2Threads;
      int local[kPer], lsum = 0;
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        local[q] = hist[tid * kPer + q];
        lsum += local[q];
      }
      int total = 0;
      int run = block_excl_scan(lsum, sh_scan, &total);
      const int want = 2 * cap;
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        if (run < want && run + local[q] >= want) sh_i0 = tid * kPer + q;
        run += local[q];
      }
      __syncthreads();
      tb = sh_i0;
      __syncthreads();
      N2_MARK(10);
    }
    // ---- class histogram -> 32-aligned segments -> scatter -> per-class sort
    for (int c = tid; c < nc; c += kN2Threads) {
      seg_end[c] = 0;
      fill[c] = 0;
    }
    __syncthreads();
    for (int r = tid; r < n; r += kN2Threads) {
      const uint64_t k = keys[r];
      const int c = (int)(k >> 52);
      if (c < nc && (int)((k >> 40) & 0xfffu) <= tb) atomicAdd(&seg_end[c], 1);
    }
    __syncthreads();
    if (warp == 0) {
      int running = 0, tmax = 0;
      for (int c0 = 0; c0 < nc; c0 += 32) {
        const int c = c0 + lane;
        const int t = c < nc ? seg_end[c] : 0;
        tmax = max(tmax, t);
        const int ta = (t + 31) & ~31;
        int incl = ta;
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += v;
        }
        if (c < nc) {
          const int s0 = running + incl - ta;
          seg_begin[c] = s0;
          seg_end[c] = s0 + t;
        }
        running += __shfl_sync(0xffffffffu, incl, 31);
      }
      for (int d = 16; d > 0; d >>= 1) tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, d));
      if (lane == 0) {
        sh_npos = running;
        sh_maxt = tmax;
      }
    }
    __syncthreads();
    N2_MARK(1);
    npos = sh_npos;
    if (npos > p.cap_pos) {  // does not fit shared memory (uniform): global-memory path
      __syncthreads();
      nms2_fallback(p, b, smem_raw);
      return;
    }
    for (int r = tid; r < n; r += kN2Threads) {
      const uint64_t k = keys[r];
      const int c = (int)(k >> 52);
      if (c < nc && (int)((k >> 40) & 0xfffu) <= tb) ckey[seg_begin[c] + atomicAdd(&fill[c], 1)] = k;
    }
    __syncthreads();
    N2_MARK(2);
    // warps claim classes; the warp that sorts a class also gathers its boxes and sets its live bits
    for (;;) {
      int c = 0;
      if (lane == 0) c = atomicAdd(&sh_next, 1);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= nc) break;
      const int s = seg_begin[c], t = seg_end[c] - s;
      if (t <= 0) continue;
      if (t <= kN2WarpSortMax) {
        uint64_t* seg = ckey + s;
        if (t <= 32) warp_rank_sort_32(seg, t);
        else if (t <= 64) warp_sort_segment<2>(seg, t);
        else if (t <= 128) warp_sort_segment<4>(seg, t);
        else warp_sort_segment<8>(seg, t);
        for (int w = lane; w < ((t + 31) >> 5); w += 32) {
          const int rem = t - (w << 5);
          alive[(s >> 5) + w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
        }
      }
    }
    __syncthreads();
    N2_MARK(8);
    // classes too large for one warp: CTA-wide sort in a scratch area (the box array of the positions
    // behind npos is free only if it fits; otherwise sort in place through the generic network)
    for (int c = 0; c < nc && sh_maxt > kN2WarpSortMax; ++c) {
      const int s = seg_begin[c], t = seg_end[c] - s;
      if (t <= kN2WarpSortMax) continue;
      const int P = pow2_ceil(t);
      uint64_t* scratch = reinterpret_cast<uint64_t*>(sh_box + s);  // 16 B per position >= 8 B * 2 t
      for (int i = tid; i < P; i += kN2Threads) scratch[i] = i < t ? ckey[s + i] : ~0ull;
      __syncthreads();
      block_sort_smem_max8(scratch, P);
      for (int i = tid; i < t; i += kN2Threads) ckey[s + i] = scratch[i];
      __syncthreads();
      for (int w = tid; w < ((t + 31) >> 5); w += kN2Threads) {
        const int rem = t - (w << 5);
        alive[(s >> 5) + w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
      }
      __syncthreads();
    }
    N2_MARK(9);
    if (tid == 0) sh_next = 0;
    // gather the boxes of every live position, all threads, all loads in flight at once
    for (int i = tid; i < npos; i += kN2Threads) {
      if ((alive[i >> 5] >> (i & 31)) & 1u) {
        const float4 bx = dense[(uint32_t)(ckey[i] & 0x1fffffu)];
        sh_box[i] = bx;
        sh_area[i] = box_area(bx);
      }
    }
    __syncthreads();
  }
  nwords = (npos + 31) >> 5;
  N2_MARK(3);

  // ---- greedy suppression (shared with nms_kernel) -------------------------------------------------
  {
    const BoxView<true> bv{smem_u32(sh_box), smem_u32(sh_area)};
    suppress_all(bv, trick, trick ? n : npos, nwords, nc, cap, alive, seg_begin, seg_end, st, &sh_next, &sh_mask, &sh_kept,
                 trick || sh_maxt > 32 * kCoopMinWords);
  }
  __syncthreads();
  N2_MARK(11);
  if (p.order == CVPP_ORDER_SCORE_DESC && !trick) {
    // compact the survivors' score-major keys behind the (dead) boxes: their count decides attempt 0, and up
    // to kN2DirectSel of them are sorted directly by the output stage (a warp covers exactly one alive word)
    uint64_t* sel = reinterpret_cast<uint64_t*>(reinterpret_cast<int*>(sh_box) + kN2Bins);
    if (tid == 0) sh_kept = 0;
    __syncthreads();
    for (int i0 = warp << 5; i0 < npos; i0 += kN2Threads) {
      const uint32_t word = alive[i0 >> 5];
      if (word == 0u) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(&sh_kept, __popc(word));
      base = __shfl_sync(0xffffffffu, base, 0);
      if ((word >> lane) & 1u) {
        const int pos = base + __popc(word & ((1u << lane) - 1u));
        if (pos < kN2DirectSel) sel[pos] = key_to_score_major(ckey[i0 + lane]);
      }
    }
    __syncthreads();
    n_surv = sh_kept;
    if (attempt == 0 && n_surv < cap) {  // the subset does not hold the global top max_det survivors: redo in full
      attempt = 1;
      tb = kN2Bins - 1;
      __syncthreads();
      continue;
    }
  }
  break;
  }
  N2_MARK(4);

  float4* ob = p.det_box + (int64_t)b * p.max_out;
  float* os = p.det_score + (int64_t)b * p.max_out;
  int32_t* oc = p.det_cls + (int64_t)b * p.max_out;
  int32_t* oa = p.det_anchor + (int64_t)b * p.max_out;

  if (p.order == CVPP_ORDER_SCORE_DESC && !trick) {
    // ---- top-max_det survivors in score order: bin selection + a small sort ------------------------
    // (boxes and areas are dead: the histogram and the selected keys reuse their memory)
    int* hist = reinterpret_cast<int*>(sh_box);                               // kN2Bins ints
    uint64_t* sel = reinterpret_cast<uint64_t*>(hist + kN2Bins);              // up to cap_pos keys (16 B/pos - 16 KB)
    if (n_surv <= kN2DirectSel) {
      // few survivors (the usual case after the score-prefix shortcut): they are already compacted - one sort
      const int P = pow2_ceil(n_surv < 32 ? 32 : n_surv);
      for (int i = n_surv + tid; i < P; i += kN2Threads) sel[i] = ~0ull;
      __syncthreads();
      N2_MARK(5);
      // (2, 4 or 8 keys per thread on fewer warps - fewer block barriers - were measured: 4.7 / 5.0 / 6.0 us vs 5.0)
      block_sort_smem_max8(sel, P);
      N2_MARK(6);
      const int n_kept = (p.max_det > 0 && p.max_det < n_surv) ? p.max_det : n_surv;
      const int n_rows = min(n_kept, p.max_out);
      for (int o = tid; o < n_rows; o += kN2Threads) {
        const uint64_t k = sel[o];
        const uint32_t anchor = (uint32_t)(k >> 12) & 0x1fffffu;
        ob[o] = dense[anchor];
        os[o] = __uint_as_float(0x7fffffffu - (uint32_t)(k >> 33));
        oc[o] = (int32_t)(k & 0xfffu);
        oa[o] = (int32_t)anchor;
      }
      N2_MARK(7);
      if (tid == 0) p.det_count[b] = n_kept;
      return;
    }
    for (int i = tid; i < kN2Bins; i += kN2Threads) hist[i] = 0;
    if (tid == 0) {
      sh_i0 = kN2Bins - 1;
      sh_kept = 0;
    }
    __syncthreads();
    for (int i = tid; i < npos; i += kN2Threads)
      if ((alive[i >> 5] >> (i & 31)) & 1u) atomicAdd(&hist[(int)(key_to_score_major(ckey[i]) >> 52)], 1);
    __syncthreads();
    // first bin whose inclusive cumulative count reaches max_det (all bins when fewer survivors)
    constexpr int kPer = kN2Bins / kN2Threads;
    int local[kPer], lsum = 0;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      local[q] = hist[tid * kPer + q];
      lsum += local[q];
    }
    int total = 0;
    int run = block_excl_scan(lsum, sh_scan, &total);
    const int want = (p.max_det > 0 && p.max_det < total) ? p.max_det : total;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      if (run < want && run + local[q] >= want) sh_i0 = tid * kPer + q;  // exactly one thread/bin matches
      run += local[q];
    }
    __syncthreads();
    const int tbin = sh_i0;
    for (int i = tid; i < npos; i += kN2Threads) {
      if ((alive[i >> 5] >> (i & 31)) & 1u) {
        const uint64_t k = key_to_score_major(ckey[i]);
        if ((int)(k >> 52) <= tbin) sel[atomicAdd(&sh_kept, 1)] = k;
      }
    }
    __syncthreads();
    const int nsel = sh_kept;
    const int P = pow2_ceil(nsel < 32 ? 32 : nsel);
    for (int i = nsel + tid; i < P; i += kN2Threads) sel[i] = ~0ull;
    __syncthreads();
    N2_MARK(5);
    block_sort_smem_max8(sel, P);
    N2_MARK(6);
    const int n_kept = want;
    const int n_rows = min(n_kept, p.max_out);
    for (int o = tid; o < n_rows; o += kN2Threads) {
      const uint64_t k = sel[o];
      const uint32_t anchor = (uint32_t)(k >> 12) & 0x1fffffu;
      ob[o] = dense[anchor];
      os[o] = __uint_as_float(0x7fffffffu - (uint32_t)(k >> 33));
      oc[o] = (int32_t)(k & 0xfffu);
      oa[o] = (int32_t)anchor;
    }
    N2_MARK(7);
    if (tid == 0) p.det_count[b] = n_kept;
    return;
  }

  // ---- positions already in output order (class-major, or global score order for the trick) ---------
  int* out_idx = reinterpret_cast<int*>(sh_area);
  const int out_cap = (cap > 0 && cap < p.max_out) ? cap : p.max_out;
  int running = 0;
  for (int base = 0; base < nwords; base += kN2Threads) {
    const int wi = base + tid;
    uint32_t m = wi < nwords ? alive[wi] : 0u;
    const int c = __popc(m);
    int total = 0;
    int pos = running + block_excl_scan(c, sh_scan, &total);
    while (m && pos < out_cap) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      out_idx[pos++] = (wi << 5) + bit;
    }
    running += total;
  }
  __syncthreads();
  int n_kept = running;
  if (cap > 0 && n_kept > cap) n_kept = cap;
  const int n_rows = min(n_kept, out_cap);
  for (int o = tid; o < n_rows; o += kN2Threads) {
    const uint64_t k = trick ? ckey[out_idx[o]] : key_to_score_major(ckey[out_idx[o]]);
    const uint32_t anchor = (uint32_t)(k >> 12) & 0x1fffffu;
    ob[o] = dense[anchor];
    os[o] = __uint_as_float(0x7fffffffu - (uint32_t)(k >> 33));
    oc[o] = (int32_t)(k & 0xfffu);
    oa[o] = (int32_t)anchor;
  }
  N2_MARK(7);
  if (tid == 0) p.det_count[b] = n_kept;
}

#ifdef CVPP_NMS_TIMING
extern "C" __attribute__((visibility("default"))) int cvpp_debug_n2_timing(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_n2_t, sizeof(long long) * 16 * 64);
}
#endif

size_t segsort_workspace_bytes(int B, int max_cand);
int segsort_launch_skip(uint64_t* keys, int32_t* cand_count, int B, int max_cand, int rule, int max_nms, void* workspace,
                        size_t workspace_bytes, const int32_t* skip, uint64_t* out_keys, int32_t* out_count,
                        cudaStream_t stream);

__global__ void __launch_bounds__(kN2Threads, 1) nms2_kernel(const __grid_constant__ Nms2Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  nms2_image(p, smem_raw);
  if (p.gather_n == 0) return;  // (uniform)
  // ---- fused row epilogue + all-gather: rows (x1, y1, x2, y2, score, class, anchor) of this image, zero rows past its
  //      count, staged in shared memory and stored as 128-bit words into every rank's gather buffer
  __syncthreads();  // every detection of the image has been written by this CTA
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = min(__ldcg(p.det_count + b), p.max_out);
  float* sh = reinterpret_cast<float*>(smem_raw);
  const int64_t t0 = (int64_t)b * p.max_out;
  for (int k = tid; k < p.max_out; k += kN2Threads) {
    float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (k < n) {
      const float4 bx = __ldcg(p.det_box + t0 + k);
      row[0] = bx.x;
      row[1] = bx.y;
      row[2] = bx.z;
      row[3] = bx.w;
      row[4] = __ldcg(p.det_score + t0 + k);
      row[5] = (float)__ldcg(p.det_cls + t0 + k);
      row[6] = (float)__ldcg(p.det_anchor + t0 + k);
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) sh[k * 7 + c] = row[c];
  }
  __syncthreads();
  const int floats = p.max_out * 7;
  const int64_t off = ((int64_t)p.gather_rank * gridDim.x + b) * floats;                       // rows of (rank, image b)
  const int64_t cnt_off = (int64_t)p.gather_n * gridDim.x * floats + (int64_t)p.gather_rank * gridDim.x + b;
  const bool vec = (floats & 3) == 0;  // every image's block is then 16-byte aligned (the buffers are)
  if (p.gather_mc) {
    float* o = p.gather_mc + off;
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(sh);
      for (int i = tid; i < (floats >> 2); i += kN2Threads) {
        const float4 q = s4[i];
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * i), "f"(q.x), "f"(q.y), "f"(q.z),
                     "f"(q.w)
                     : "memory");
      }
    } else {
      for (int i = tid; i < floats; i += kN2Threads)
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(o + i), "f"(sh[i]) : "memory");
    }
    if (tid == 0) asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.gather_mc + cnt_off), "f"((float)n) : "memory");
    return;
  }
  for (int d = 0; d < p.gather_n; ++d) {
    float* o = p.gather_dst[d] + off;
    if (vec) {
      float4* o4 = reinterpret_cast<float4*>(o);
      const float4* s4 = reinterpret_cast<const float4*>(sh);
      for (int i = tid; i < (floats >> 2); i += kN2Threads) o4[i] = s4[i];
    } else {
      for (int i = tid; i < floats; i += kN2Threads) o[i] = sh[i];
    }
    if (tid == 0) p.gather_dst[d][cnt_off] = (float)n;
  }
}

static size_t align256_(size_t x) { return (x + 255) & ~(size_t)255; }

size_t sort_nms_workspace_bytes(int B, int max_cand, int nc) {
  return 256 + 2 * align256_((size_t)B * sizeof(int32_t)) + align256_((size_t)B * (size_t)max_cand * sizeof(uint64_t)) +
         align256_(segsort_workspace_bytes(B, max_cand)) + align256_(nms_workspace_bytes(B, max_cand, nc));
}

// cand_key: UNSORTED class-major keys (as the filter kernels emit them); never modified (the fallback sorts
// into the workspace), so the call can be repeated on the same candidates.
int sort_nms_launch(const uint64_t* cand_key, const int32_t* cand_count, const float* box_dense, int B, int max_cand, int64_t A,
                    int nc, double iou_thres, int rule, int order, int max_det, int max_nms, int max_out, float* det_box,
                    float* det_score, int32_t* det_cls, int32_t* det_anchor, int32_t* det_count, void* workspace,
                    size_t workspace_bytes, cudaStream_t stream, int32_t* cand_count_out, float* const* gather_dst,
                    float* gather_mc, int gather_n, int gather_rank) {
  if (!cand_key || !cand_count || !box_dense || !det_box || !det_score || !det_cls || !det_anchor || !det_count) {
    set_error("sort_nms: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || max_cand < 1 || A < 1 || nc < 1 || nc > CVPP_MAX_CLASSES || max_out < 1) {
    set_error("sort_nms: bad sizes (B=%d max_cand=%d A=%lld nc=%d max_out=%d)", B, max_cand, (long long)A, nc, max_out);
    return CVPP_ERR_INVALID_ARG;
  }
  if (!(iou_thres >= 0.0 && iou_thres <= 1.0)) {
    set_error("sort_nms: Invalid IoU %f, valid values are between 0.0 and 1.0", iou_thres);
    return CVPP_ERR_INVALID_ARG;
  }
  if (rule < 0 || rule > 3 || order < 0 || order > 1) {
    set_error("sort_nms: unknown rule %d / order %d", rule, order);
    return CVPP_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(box_dense) & 15u) || (reinterpret_cast<uintptr_t>(det_box) & 15u)) {
    set_error("sort_nms: box arrays must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  if (B == 0) return CVPP_OK;
  if (!workspace || workspace_bytes < sort_nms_workspace_bytes(B, max_cand, nc)) {
    set_error("sort_nms: workspace of %zu bytes needed, got %zu", sort_nms_workspace_bytes(B, max_cand, nc), workspace_bytes);
    return CVPP_ERR_WORKSPACE;
  }
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != CVPP_OK) return rc;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  base += align256_((size_t)B * sizeof(int32_t));  // (formerly the per-image done flags)
  int32_t* fb_count = reinterpret_cast<int32_t*>(base);  // fallback: counts after the max_nms cut
  base += align256_((size_t)B * sizeof(int32_t));
  uint64_t* fb_keys = reinterpret_cast<uint64_t*>(base);  // fallback: sorted keys
  base += align256_((size_t)B * (size_t)max_cand * sizeof(uint64_t));
  void* sort_ws = reinterpret_cast<void*>(base);
  const size_t sort_bytes = segsort_workspace_bytes(B, max_cand);
  base += align256_(sort_bytes);
  void* nms_ws = reinterpret_cast<void*>(base);
  const size_t nms_bytes = nms_workspace_bytes(B, max_cand, nc);

  Nms2Params p{};
  // the in-kernel fallback runs the body of nms_kernel on the keys it has sorted into fb_keys
  size_t smem_fb = 0;
  rc = nms_prepare(p.fb, smem_fb, fb_keys, fb_count, box_dense, B, max_cand, A, nc, iou_thres, rule, order, max_det, max_out,
                   det_box, det_score, det_cls, det_anchor, det_count, nms_ws, nms_bytes, nullptr, di.max_smem);
  if (rc != CVPP_OK) return rc;
  p.fb_keys = fb_keys;
  p.fb_count = fb_count;
  p.fb_sort_ws = reinterpret_cast<uint64_t*>(sort_ws);
  {
    int P = 32;
    while (P < max_cand) P <<= 1;
    p.fb_sort_stride = P;
  }
  p.thr_eff = p.fb.thr_eff;
  p.thr_mid = p.fb.thr_mid;
  p.tie_up = p.fb.tie_up;
  p.cand_key = cand_key;
  p.cand_count = cand_count;
  p.box_dense = reinterpret_cast<const float4*>(box_dense);
  p.max_cand = max_cand;
  p.A = A;
  p.nc = nc;
  p.rule = rule;
  p.order = order;
  p.max_det = max_det;
  p.max_nms = max_nms;
  p.max_out = max_out;
  p.det_box = reinterpret_cast<float4*>(det_box);
  p.det_score = det_score;
  p.det_cls = det_cls;
  p.det_anchor = det_anchor;
  p.det_count = det_count;
  p.cand_count_out = cand_count_out;
  p.gather_n = 0;
  p.gather_mc = nullptr;
  p.gather_rank = 0;
  if (gather_n > 0) {
    if (gather_n > CVPP_MAX_PEERS || gather_rank < 0 || gather_rank >= gather_n || (!gather_dst && !gather_mc)) {
      set_error("sort_nms: bad gather description (n=%d rank=%d)", gather_n, gather_rank);
      return CVPP_ERR_INVALID_ARG;
    }
    if (gather_mc && (reinterpret_cast<uintptr_t>(gather_mc) & 15u)) {
      set_error("sort_nms: the multicast gather address must be 16-byte aligned");
      return CVPP_ERR_ALIGNMENT;
    }
    p.gather_n = gather_n;
    p.gather_rank = gather_rank;
    p.gather_mc = gather_mc;
    for (int d = 0; d < gather_n && !gather_mc; ++d) {
      if (!gather_dst[d] || (reinterpret_cast<uintptr_t>(gather_dst[d]) & 15u)) {
        set_error("sort_nms: gather destination %d is NULL or not 16-byte aligned", d);
        return CVPP_ERR_INVALID_ARG;
      }
      p.gather_dst[d] = gather_dst[d];
    }
  }
  // shared memory of the fused path: 28 B per position + 1 bit, 3 ints per class; the selection stage needs
  // 16 KB of histogram inside the 16 B/position box array and 8 B/position of selected keys behind it
  const size_t fixed = (size_t)nc * 12 + 256;
  size_t cap_pos = 0;
  if (fixed + 64 * 1024 <= (size_t)di.max_smem) {
    cap_pos = ((size_t)di.max_smem - fixed - 1024) * 8 / (28 * 8 + 1);
    const size_t worst = (size_t)max_cand + 32u * (size_t)nc;  // never need more positions than this
    if (cap_pos > worst) cap_pos = worst;
    cap_pos &= ~(size_t)31;
    if (cap_pos < 2048) cap_pos = 2048;  // 16 B * cap_pos must hold the 16 KB histogram + 8 B * cap_pos of keys
    if (fixed + cap_pos * 28 + cap_pos / 8 + 1024 > (size_t)di.max_smem) cap_pos = 0;
  }
  p.cap_pos = (int)cap_pos;   // 0: nothing fits the fused layout, every image takes the fallback
  p.mask_words = (int)(cap_pos / 32);
  size_t smem = cap_pos * 28 + (size_t)p.mask_words * 4 + (size_t)nc * 12;
  if (smem_fb > smem) smem = smem_fb;
  if (p.gather_n > 0 && (size_t)max_out * 28 > smem) smem = (size_t)max_out * 28;  // staging of the image's rows
  // the fallback sort uses whatever dynamic shared memory the launch has (power-of-two key count)
  int fb_keys_smem = 32;
  while ((size_t)fb_keys_smem * 2 * sizeof(uint64_t) <= smem && fb_keys_smem < 8192) fb_keys_smem <<= 1;
  p.fb_smem_keys = fb_keys_smem;
  static unsigned long long attr_done = 0;
  static int attr_bytes = 0;
  if ((int)smem > attr_bytes) {
    attr_done = 0;
    attr_bytes = (int)smem;
  }
  rc = ensure_smem_attr(reinterpret_cast<const void*>(nms2_kernel), attr_bytes, di.device, &attr_done);
  if (rc != CVPP_OK) return rc;
  {
    // programmatic dependent launch: when the previous operation in the stream is a kernel that executes
    // griddepcontrol.launch_dependents (the decode kernels do), this grid is scheduled while that one still runs and
    // waits at griddepcontrol.wait; otherwise this is an ordinary stream-ordered launch
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)B);
    cfg.blockDim = dim3(kN2Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CVPP_CUDA_TRY(cudaLaunchKernelEx(&cfg, nms2_kernel, p));
  }
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

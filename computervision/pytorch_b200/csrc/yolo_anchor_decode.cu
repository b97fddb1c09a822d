// yolo_anchor_decode.cu — anchor-based YOLO head decode + confidence filter, fused (sm_100a).
// One streaming kernel, two arithmetic modes:
//
// MODE_V7 replaces (reference file:line): YOLOv7.decode_box core/algorithms/yolo_v7.py:245-344 (per-level
// reshape/permute, six sigmoids, grid + anchor box math, normalisation, concatenation) and the
// candidate stage of YOLOv7._nms :361 (xywh_to_xyxy_torch, core/utils/bboxes.py:29-49), :370 (class max,
// first index), :377 (obj * class_conf >= conf).  One key per surviving anchor.
//
// MODE_V3 replaces: predict_bounding_bbox core/predict/yolov3_decode.py:12-29, Decoder._yolo_post_process
// :40-51 (xyxy, conf * prob), Decoder.__call__ :53-63 (concatenate scales) and the mask of yolo3_nms
// core/utils/nms.py:60 (score >= conf per (anchor, class)).  One key per surviving (anchor, class).
//
// The per-class NMS that follows (yolo_v7.py:396-413, nms.py:66-71) is cvpp_segmented_sort + cvpp_nms with
// CVPP_NMS_RULE_PER_CLASS / CVPP_ORDER_CLASS_MAJOR.
//
// Memory-bound: (5 + nc) fp32 per anchor, read once: 25 200 x 85 x 4 = 8 568 000 B per YOLOv7 image,
// 10 647 x 25 x 4 = 1 064 700 B per YOLOv3-VOC image.  Same machinery as the YOLOv8 kernel: persistent
// CTAs, independent warps, a tile = 128 cells of one (image, level, anchor), streamed as 16-channel x
// 128-cell chunks by ONE 3-D tensor-map TMA load each into a private 2-stage ring.  V7: class argmax on
// logits with the exact first-index tie repair.  V3: every class logit is compared with a per-cell
// conservative logit cut derived from the objectness (sigma(o) * sigma(c) >= thr  =>  c >= logit(thr /
// sigma(o)) - slack) into a per-lane bit mask (no branch); the flagged logits of a chunk are compacted across the lanes and
// the exact fp32 product is evaluated 32 at a time from the staged chunk (round 2: 127.8 -> 95 us per 256 images).  Levels whose rows
// are not 16-byte aligned (13 x 13 = 169 cells) go through the thread-per-anchor kernel instead.
#include <cuda.h>

#include "cvpp_common.cuh"

namespace cvpp {

constexpr int kYaChunkRows = 16;
constexpr int kYaStages = 2;
// CPL = cells per lane: tiles of 32 * CPL cells, chunks of 16 rows x 32 * CPL cells (4 KB at CPL 2, 8 KB at CPL 4).
// The per-chunk arithmetic is latency-bound per warp (fewer warps are slower for both modes, unlike the YOLOv8
// kernel), so the default is CPL 2: half the per-thread state and twice the warps.
constexpr int ya_max_warps(int mode, int cpl) { return cpl == 4 ? (mode == 0 ? 14 : 12) : 24; }
constexpr int kYaHitCap = 128;    // V3: per-warp staging of candidate keys (8 B each, see V3Stage)
constexpr int kYaMaxLevels = 4;
constexpr int MODE_V7 = 0;
constexpr int MODE_V3 = 1;

struct YaLevel {
  const float* ptr;
  int64_t batch_stride;
  int64_t chan_stride;
  int hw, w, h;
  float aw[3], ah[3];   // V7: anchors in grid units (anchor / stride); V3: anchors / input size (fp32, like the reference)
  int anchor_off;       // first output anchor index of this level (times B when the batch is merged)
  int image_stride;     // merged batch: 3 * hw (anchor index advance per image), else 0
  int tile_off;         // first tile (within an image) of this level
  int tiles_per_anchor;
  float inv_tiles_per_anchor;
};

struct YaParams {
  CUtensorMap tmap[kYaMaxLevels];       // box tile_a cells x 16 rows x 1 image
  CUtensorMap tmap_tail[kYaMaxLevels];  // box tile_a cells x tail_rows x 1 image: the last chunk of an anchor reads
                                        // exactly the rows that are left ((5 + nc) mod 16), not 16
  int tail_rows;
  int tile_a;           // cells per tile (32 * CPL)
  YaLevel lv[kYaMaxLevels];
  int num_levels, B, nc, tiles_per_image, total_tiles;
  float inv_tiles_per_image;
  int merged;           // V3: Decoder flattens the batch (yolov3_decode.py:47-50): one output "image"
  int64_t A;            // anchors per OUTPUT image
  float conf_thres;
  uint64_t* cand_key;
  int32_t* cand_count;
  float4* box_dense;
  float2* aux_dense;    // V7: (obj, class_conf) per anchor
  int max_cand;
};

// single-use input: L2 evict-first (see yolov8_decode.cu)
__device__ __forceinline__ uint64_t ya_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ void ya_tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar,
                                               uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}

// n / d for 0 <= n < 2^24 through a float reciprocal estimate corrected by at most one (a 32-bit integer division is ~25
// instructions; the tile bookkeeping did four per tile and the box decode one per cell: 6.7 % of the V7 kernel)
__device__ __forceinline__ int ya_div(int n, int d, float inv) {
  int q = (int)((float)n * inv);
  const int r = n - q * d;
  q += r >= d ? 1 : (r < 0 ? -1 : 0);
  return q;
}

// ---- YOLOv7 arithmetic ----------------------------------------------------------------------------
__device__ __forceinline__ void v7_class_step(float x, int c, float& best, int& arg, float& prev) {
  const bool gt = x > best;
  prev = gt ? best : prev;
  arg = gt ? c : arg;
  best = fmaxf(best, x);
}

// exact `class_conf, class_pred = max(sigmoid(cls), 1)`: first index of the maximum sigmoid
__device__ __noinline__ void v7_class_argmax_sigmoid(const float* col, int64_t cs, int nc, float& s, int& arg) {
  float bs = -1.0f;
  int ba = 0;
  for (int c = 0; c < nc; ++c) {
    const float v = sigmoid_precise(col[(int64_t)c * cs]);
    if (v > bs) {
      bs = v;
      ba = c;
    }
  }
  s = bs;
  arg = ba;
}

struct V7Cell {
  float4 box;  // normalised xyxy
  float obj, cconf, score;
  int cls;
  bool cand;
};

// yolo_v7.py:329-342 + :361: t = (tx, ty, tw, th) logits of one cell of a W x H grid
__device__ __forceinline__ float4 v7_box(float tx, float ty, float tw, float th, int cell, int W, int H, float aw, float ah) {
  const int iy = ya_div(cell, W, 1.0f / (float)W), ix = cell - iy * W;
  const float sx = sigmoid_precise(tx), sy = sigmoid_precise(ty);
  const float sw = sigmoid_precise(tw), sh = sigmoid_precise(th);
  const float bx = fadd(fsub(fmul(sx, 2.0f), 0.5f), (float)ix);
  const float by = fadd(fsub(fmul(sy, 2.0f), 0.5f), (float)iy);
  const float w2 = fmul(sw, 2.0f), h2 = fmul(sh, 2.0f);
  const float bw = fmul(fmul(w2, w2), aw), bh = fmul(fmul(h2, h2), ah);
  const float cx = fdiv(bx, (float)W), cy = fdiv(by, (float)H);
  const float w = fdiv(bw, (float)W), h = fdiv(bh, (float)H);
  const float hw = fmul(w, 0.5f), hh = fmul(h, 0.5f);
  return make_float4(fsub(cx, hw), fsub(cy, hh), fadd(cx, hw), fadd(cy, hh));
}

__device__ __forceinline__ void v7_finalize(float tobj, float best, int arg_in, float prev, const float* cls_col, int64_t cs,
                                            int nc, float conf_thres, V7Cell& o) {
  o.obj = sigmoid_precise(tobj);
  o.cconf = sigmoid_precise(best);
  o.cls = arg_in;
  o.score = fmul(o.obj, o.cconf);
  o.cand = o.score >= conf_thres;
  if (o.cand && sigmoid_precise(prev) >= o.cconf) {  // an earlier class with the same rounded sigmoid
    v7_class_argmax_sigmoid(cls_col, cs, nc, o.cconf, o.cls);
    o.score = fmul(o.obj, o.cconf);
    o.cand = o.score >= conf_thres;
  }
}

// ---- YOLOv3 arithmetic ----------------------------------------------------------------------------
// yolov3_decode.py:22-23,47: xy = (sigmoid(t) + grid) / H for BOTH coordinates, wh = exp(t) * anchor_norm
__device__ __forceinline__ float4 v3_box(float tx, float ty, float tw, float th, int cell, int W, int H, float aw, float ah) {
  const int iy = ya_div(cell, W, 1.0f / (float)W), ix = cell - iy * W;
  const float bx = fdiv(fadd(sigmoid_precise(tx), (float)ix), (float)H);
  const float by = fdiv(fadd(sigmoid_precise(ty), (float)iy), (float)H);
  const float bw = fmul(expf(tw), aw), bh = fmul(expf(th), ah);
  const float hw = fmul(bw, 0.5f), hh = fmul(bh, 0.5f);
  return make_float4(fsub(bx, hw), fsub(by, hh), fadd(bx, hw), fadd(by, hh));
}

// per-cell logit cut: a class can only reach sigmoid(o) * sigmoid(c) >= thr when c >= cut.  Conservative
// (relative margin 1e-5 on the probability, 0.05 on the logit): survivors are re-tested exactly.
__device__ __forceinline__ float v3_logit_cut(float so, float thr) {
  if (so < thr) return INFINITY;  // sigmoid(c) <= 1 and fmul rounds monotonically: no class can pass
  if (thr <= 0.0f) return -INFINITY;
  const float r = fminf(__fdividef(thr, so), 1.0f) * (1.0f - 1e-5f);
  return __logf(__fdividef(r, 1.0f - r)) - 0.05f;
}

template <int MODE>
__device__ __forceinline__ int ya_local_index(int a, int cell, int hw) {
  return MODE == MODE_V7 ? a * hw + cell : cell * 3 + a;
}

__device__ __forceinline__ void ya_tile_info(const YaParams& p, int g, int& b, int& l, int& a, int& cell0, int& nA) {
  b = ya_div(g, p.tiles_per_image, p.inv_tiles_per_image);
  const int j = g - b * p.tiles_per_image;
  l = 0;
#pragma unroll
  for (int q = 1; q < kYaMaxLevels; ++q)
    if (q < p.num_levels && j >= p.lv[q].tile_off) l = q;
  const int t = j - p.lv[l].tile_off;
  a = ya_div(t, p.lv[l].tiles_per_anchor, p.lv[l].inv_tiles_per_anchor);
  cell0 = (t - a * p.lv[l].tiles_per_anchor) * p.tile_a;
  nA = min(p.tile_a, p.lv[l].hw - cell0);
}

// ---- V3 candidate staging ---------------------------------------------------------------------------
// Round 2 (the first form pushed a RECORD per passing logit inside the class loop - a divergent region with a shared-memory
// atomic per (row, cell), 40 per tile, two thirds of them taken at eval thresholds: 30 % of the kernel's instructions were
// BSSY / BRA / BSYNC / ISETP, 1 630 warp-instructions per 64-cell tile).  Now the class loop only ORs a bit per passing
// logit into a per-lane mask (no branch); after the chunk's rows, while the chunk is still in the ring stage, the
// warp pops one bit per lane per round, re-reads that logit from shared memory, evaluates the exact fp32 product at up to
// 32 hits per round and stages the surviving KEYS per warp; they leave with one global atomic per tile.
struct V3Stage {
  uint64_t* keys;   // [kYaHitCap] per warp
  uint16_t* list;   // [kYaListCap] per warp: (lane << 5 | bit) of the chunk's hits, compacted across the lanes
};

// all lanes; n = staged keys (warp-uniform register)
__device__ __forceinline__ void v3_flush(const V3Stage& st, const YaParams& p, int ob, int& n) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (n == 0) return;
  pdl_wait();  // the counts have been zeroed by the programmatic predecessor
  int gbase = 0;
  if (lane == 0) gbase = atomicAdd(p.cand_count + ob, n);
  gbase = __shfl_sync(0xffffffffu, gbase, 0);
  uint64_t* dst = p.cand_key + (int64_t)ob * p.max_cand;
  for (int i = lane; i < n; i += 32)
    if (gbase + i < p.max_cand) dst[gbase + i] = st.keys[i];
  __syncwarp();
  n = 0;
}

// exact evaluation of the logits flagged in `hm` (bit r * CPL + k = row r of the staged chunk, cell k of the lane).
// A cell with a high objectness has most of its classes above the cut, so popping the bits lane by lane is imbalanced (10
// rounds per tile measured, most lanes idle): the hits are first COMPACTED across the lanes into a per-warp list (a shuffle
// scan of the per-lane counts, then a cheap per-lane write loop) and evaluated 32 at a time at full lane efficiency.
constexpr int kYaListCap = 128;  // (lane, bit) entries per chunk; denser chunks take the per-lane loop
template <int CPL>
__device__ __forceinline__ void v3_eval_hits(unsigned hm, const float* stage, int tile_a, int c_of_row0, const float (&so)[CPL],
                                             unsigned& boxed, const V3Stage& st, int& wk_n, const YaParams& p, int ob,
                                             int anchor_base, int a, int cell0, int hw) {
  static_assert(CPL == 2, "the V3 hit mask holds 16 rows x 2 cells");
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int cnt = __popc(hm);
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += u;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;  // (uniform)
  if (total <= kYaListCap) {
    int pos = incl - cnt;
    while (hm) {
      st.list[pos++] = (uint16_t)((lane << 5) | (__ffs(hm) - 1));
      hm &= hm - 1u;
    }
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
      const int i = base + lane;
      const bool has = i < total;
      const int e = has ? (int)st.list[i] : 0;
      const int src = e >> 5, bit = e & 31, r = bit >> 1, k = bit & 1;
      const float so0 = __shfl_sync(0xffffffffu, so[0], src), so1 = __shfl_sync(0xffffffffu, so[1], src);
      const float x = stage[r * tile_a + CPL * src + k];
      const float score = fmul(k ? so1 : so0, sigmoid_precise(x));  // confidence * class_prob (yolov3_decode.py:49)
      const bool hit = has && score >= p.conf_thres;                  // nms.py:60
      const unsigned mk = __ballot_sync(0xffffffffu, hit);
      if (mk) {
        const unsigned b0 = __reduce_or_sync(0xffffffffu, (hit && k == 0) ? (1u << src) : 0u);
        const unsigned b1 = __reduce_or_sync(0xffffffffu, (hit && k == 1) ? (1u << src) : 0u);
        boxed |= ((b0 >> lane) & 1u) | (((b1 >> lane) & 1u) << 1);
        if (wk_n + __popc(mk) > kYaHitCap) v3_flush(st, p, ob, wk_n);
        if (hit) {
          const int anchor = anchor_base + ya_local_index<MODE_V3>(a, cell0 + CPL * src + k, hw);
          st.keys[wk_n + __popc(mk & lt)] = key_pack((uint32_t)(c_of_row0 + r), __float_as_uint(score), (uint32_t)anchor);
        }
        wk_n += __popc(mk);
      }
    }
    __syncwarp();
    return;
  }
  while (__any_sync(0xffffffffu, hm != 0u)) {
    const bool has = hm != 0u;
    const int bit = has ? __ffs(hm) - 1 : 0;
    hm &= hm - 1u;  // (0 stays 0)
    const int r = bit >> 1, k = bit & 1;
    const float x = stage[r * tile_a + CPL * lane + k];
    const float score = fmul(k ? so[1] : so[0], sigmoid_precise(x));
    const bool hit = has && score >= p.conf_thres;
    const unsigned mk = __ballot_sync(0xffffffffu, hit);
    if (mk) {
      if (wk_n + __popc(mk) > kYaHitCap) v3_flush(st, p, ob, wk_n);
      if (hit) {
        const int anchor = anchor_base + ya_local_index<MODE_V3>(a, cell0 + CPL * lane + k, hw);
        st.keys[wk_n + __popc(mk & lt)] = key_pack((uint32_t)(c_of_row0 + r), __float_as_uint(score), (uint32_t)anchor);
        boxed |= 1u << k;
      }
      wk_n += __popc(mk);
    }
  }
  __syncwarp();
}

// class rows [RBEGIN, rows) of one 16-row chunk; FULL: rows == 16 at compile time
template <int MODE, int CPL, int RBEGIN, bool FULL>
__device__ __forceinline__ void ya_class_rows(const float (&v)[kYaChunkRows][CPL], int rows, int c_first, float (&best)[CPL],
                                              int (&arg)[CPL], float (&prev)[CPL], const float (&cut)[CPL], unsigned& hm) {
#pragma unroll
  for (int r = RBEGIN; r < kYaChunkRows; ++r) {
    if (FULL || r < rows) {
      const int c = c_first + (r - RBEGIN);
      const float(&x)[CPL] = v[r];
      if (MODE == MODE_V7) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) v7_class_step(x[k], c, best[k], arg[k], prev[k]);
      } else {
#pragma unroll
        for (int k = 0; k < CPL; ++k) hm |= (x[k] >= cut[k]) ? (1u << ((r * CPL + k) & 31)) : 0u;
      }
    }
  }
}

template <int CPL>
struct YaVec;
template <>
struct YaVec<4> {
  typedef float4 type;
  static __device__ __forceinline__ void unpack(const float4& q, float (&o)[4]) {
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
  }
};
template <>
struct YaVec<2> {
  typedef float2 type;
  static __device__ __forceinline__ void unpack(const float2& q, float (&o)[2]) {
    o[0] = q.x; o[1] = q.y;
  }
};

template <int MODE, int CPL>
__global__ void __launch_bounds__(ya_max_warps(MODE, CPL) * 32, 1)
yolo_anchor_stream_kernel(const __grid_constant__ YaParams p) {
  constexpr int kYaTileA = 32 * CPL;
  constexpr int kYaChunkFloats = kYaChunkRows * kYaTileA;
  const int kYaWarps = blockDim.x >> 5;  // picked by the host (ya_pick_warps)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ring = reinterpret_cast<float*>(smem_raw) + (size_t)warp * (kYaStages * kYaChunkFloats);
  unsigned char* after_rings = smem_raw + (size_t)kYaWarps * kYaStages * kYaChunkFloats * sizeof(float);
  uint64_t* bar = reinterpret_cast<uint64_t*>(after_rings) + warp * kYaStages;
  V3Stage st;  // V3 only: [keys u64 x cap] per warp
  st.keys = reinterpret_cast<uint64_t*>(after_rings + (size_t)kYaWarps * kYaStages * sizeof(uint64_t)) + (size_t)warp * kYaHitCap;
  st.list = reinterpret_cast<uint16_t*>(after_rings + (size_t)kYaWarps * (kYaStages * sizeof(uint64_t) + kYaHitCap * sizeof(uint64_t))) +
            (size_t)warp * 128;
  int wk_n = 0;  // V3: keys staged by this warp (warp-uniform)
  const int nc = p.nc;
  const int attrs = 5 + nc;
  const int nchunks = (attrs + kYaChunkRows - 1) / kYaChunkRows;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kYaStages; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  __syncwarp();

  const int stride_tiles = gridDim.x * kYaWarps;
  const int first = blockIdx.x * kYaWarps + warp;  // the warps of a CTA stream ADJACENT tiles: their 512 B row pieces share DRAM pages
  const int n_tiles = first < p.total_tiles ? (p.total_tiles - first + stride_tiles - 1) / stride_tiles : 0;
  const int total_q = n_tiles * nchunks;

  const uint64_t policy = ya_evict_first_policy();
  int pq = 0, pj = 0, pg = first, pb = 0, pl = 0, pa = 0, pcell0 = 0;
  auto issue = [&]() {
    if (pj == 0) {
      int nA_unused;
      ya_tile_info(p, pg, pb, pl, pa, pcell0, nA_unused);
    }
    if (lane == 0) {
      uint64_t* fb = &bar[pq & (kYaStages - 1)];
      const bool tail = pj == nchunks - 1;
      mbar_arrive_expect_tx(fb, (uint32_t)((tail ? p.tail_rows : kYaChunkRows) * kYaTileA * sizeof(float)));
      ya_tma_load_3d(ring + (pq & (kYaStages - 1)) * kYaChunkFloats, tail ? &p.tmap_tail[pl] : &p.tmap[pl], pcell0,
                     pa * attrs + kYaChunkRows * pj, pb, fb, policy);
    }
    ++pq;
    if (++pj == nchunks) {
      pj = 0;
      pg += stride_tiles;
    }
  };
  for (int q = 0; q < kYaStages && q < total_q; ++q) issue();

  float t5[5][CPL];            // tx, ty, tw, th, tobj logits of the lane's CPL cells
  float best[CPL], prev[CPL];  // V7: running class scan
  int arg[CPL];
  float so[CPL], cut[CPL];     // V3: sigmoid(obj), per-cell class logit cut
  unsigned boxed = 0;      // V3: cells that pushed a record (their box is written at the end of the tile)
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    best[k] = prev[k] = -INFINITY;
    arg[k] = 0;
    so[k] = 0.0f;
    cut[k] = INFINITY;
  }
  int b = 0, l = 0, a = 0, cell0 = 0, nA = 0;
  int j = 0, g = first;
  for (int q = 0; q < total_q; ++q) {
    if (j == 0) ya_tile_info(p, g, b, l, a, cell0, nA);
    const YaLevel& L = p.lv[l];
    const int s = q & (kYaStages - 1);
    mbar_wait(&bar[s], (uint32_t)(q / kYaStages) & 1u);
    float v[kYaChunkRows][CPL];
    {
      typedef typename YaVec<CPL>::type vec_t;
      const vec_t* src = reinterpret_cast<const vec_t*>(ring + s * kYaChunkFloats) + lane;
#pragma unroll
      for (int r = 0; r < kYaChunkRows; ++r) YaVec<CPL>::unpack(src[r * 32], v[r]);
    }
    __syncwarp();
    if (MODE == MODE_V7 && pq < total_q) issue();  // V3 re-arms the stage after its hits were re-read from it (below)

    // rows of this chunk are attributes [16 j, 16 j + 16) of the anchor: 0..4 box/obj, 5.. classes
    const bool active = CPL * lane < nA;
    const int ob = p.merged ? 0 : b;
    const int anchor_base = L.anchor_off + b * L.image_stride;
    unsigned hm = 0;  // V3: logits of this chunk above the lane's cut, bit r * CPL + k
    if (j == 0) {
#pragma unroll
      for (int r = 0; r < 5; ++r) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) t5[r][k] = v[r][k];
      }
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        best[k] = prev[k] = -INFINITY;
        arg[k] = 0;
      }
      boxed = 0;
      if (MODE == MODE_V3) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          so[k] = sigmoid_precise(t5[4][k]);
          cut[k] = active ? v3_logit_cut(so[k], p.conf_thres) : INFINITY;
        }
      }
      if (attrs >= kYaChunkRows)
        ya_class_rows<MODE, CPL, 5, true>(v, kYaChunkRows, 0, best, arg, prev, cut, hm);
      else
        ya_class_rows<MODE, CPL, 5, false>(v, attrs, 0, best, arg, prev, cut, hm);
    } else if (j < nchunks - 1) {
      ya_class_rows<MODE, CPL, 0, true>(v, kYaChunkRows, kYaChunkRows * j - 5, best, arg, prev, cut, hm);
    } else {
      ya_class_rows<MODE, CPL, 0, false>(v, attrs - kYaChunkRows * j, kYaChunkRows * j - 5, best, arg, prev, cut, hm);
    }
    if constexpr (MODE == MODE_V3 && CPL == 2) {
      // class of row 0 of this chunk: rows are attributes 16 j .., classes start at attribute 5
      v3_eval_hits<CPL>(hm, ring + s * kYaChunkFloats, kYaTileA, kYaChunkRows * j - 5, so, boxed, st, wk_n, p, ob, anchor_base, a,
                        cell0, L.hw);
      if (pq < total_q) issue();
    }

    if (++j == nchunks) {  // tile complete
      j = 0;
      g += stride_tiles;
      if (MODE == MODE_V3) {
        v3_flush(st, p, ob, wk_n);
        // (compacting the boxed cells across the lanes before v3_box was measured: 97.2 vs 95.0 us - not in)
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          if (boxed & (1u << k)) {
            const int cell = cell0 + CPL * lane + k;
            p.box_dense[(int64_t)ob * p.A + anchor_base + ya_local_index<MODE_V3>(a, cell, L.hw)] =
                v3_box(t5[0][k], t5[1][k], t5[2][k], t5[3][k], cell, L.w, L.h, L.aw[a], L.ah[a]);
          }
        }
      } else {
        const float* col = L.ptr + (int64_t)b * L.batch_stride + (int64_t)(a * attrs + 5) * L.chan_stride + cell0 + CPL * lane;
        V7Cell cell[CPL];
        unsigned m[CPL];
        int total = 0;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          cell[k].cand = false;
          if (active) {
            v7_finalize(t5[4][k], best[k], arg[k], prev[k], col + k, L.chan_stride, nc, p.conf_thres, cell[k]);
            if (cell[k].cand)
              cell[k].box = v7_box(t5[0][k], t5[1][k], t5[2][k], t5[3][k], cell0 + CPL * lane + k, L.w, L.h, L.aw[a], L.ah[a]);
          }
          m[k] = __ballot_sync(0xffffffffu, cell[k].cand);
          total += __popc(m[k]);
        }
        if (total) {
          pdl_wait();  // the counts have been zeroed by the programmatic predecessor
          int base = 0;
          if (lane == 0) base = atomicAdd(p.cand_count + ob, total);
          base = __shfl_sync(0xffffffffu, base, 0);
          const unsigned lt = (1u << lane) - 1u;
#pragma unroll
          for (int k = 0; k < CPL; ++k) {
            if (cell[k].cand) {
              const int slot = base + __popc(m[k] & lt);
              const int anchor = anchor_base + ya_local_index<MODE_V7>(a, cell0 + CPL * lane + k, L.hw);
              if (slot < p.max_cand)
                p.cand_key[(int64_t)ob * p.max_cand + slot] =
                    key_pack((uint32_t)cell[k].cls, __float_as_uint(cell[k].score), (uint32_t)anchor);
              p.box_dense[(int64_t)ob * p.A + anchor] = cell[k].box;
              p.aux_dense[(int64_t)ob * p.A + anchor] = make_float2(cell[k].obj, cell[k].cconf);
            }
            base += __popc(m[k]);
          }
        }
      }
    }
  }
}

// generic kernel: one thread per anchor, no alignment requirements
// Units of blockDim.x consecutive anchors of one image, walked grid-stride (the launch uses one CTA per unit).
template <int MODE>
__global__ void __launch_bounds__(128) yolo_anchor_generic_kernel(const __grid_constant__ YaParams p, int anchors_in) {
  const int nc = p.nc, attrs = 5 + nc;
  const int lane = threadIdx.x & 31;
  const int units_per_image = (anchors_in + (int)blockDim.x - 1) / (int)blockDim.x;
  const int total_units = units_per_image * p.B;
  for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
  const int b = unit / units_per_image;
  const int idx = (unit - b * units_per_image) * blockDim.x + threadIdx.x;  // position among the anchors of the listed levels
  const int ob = p.merged ? 0 : b;
  V7Cell o;
  o.cand = false;
  int anchor = 0;
  if (idx < anchors_in) {
    int l = 0, start = 0, acc = 0;
#pragma unroll
    for (int q = 0; q < kYaMaxLevels; ++q) {
      if (q < p.num_levels) {
        if (idx >= acc) {
          l = q;
          start = acc;
        }
        acc += 3 * p.lv[q].hw;
      }
    }
    const YaLevel& L = p.lv[l];
    const int rel = idx - start;
    const int a = rel / L.hw, cell = rel - a * L.hw;
    anchor = L.anchor_off + b * L.image_stride + ya_local_index<MODE>(a, cell, L.hw);
    const float* base = L.ptr + (int64_t)b * L.batch_stride + (int64_t)(a * attrs) * L.chan_stride + cell;
    const int64_t cs = L.chan_stride;
    if (MODE == MODE_V7) {
      float best = -INFINITY, prev = -INFINITY;
      int arg = 0;
      for (int c = 0; c < nc; ++c) v7_class_step(base[(int64_t)(5 + c) * cs], c, best, arg, prev);
      v7_finalize(base[4 * cs], best, arg, prev, base + 5 * cs, cs, nc, p.conf_thres, o);
      if (o.cand) o.box = v7_box(base[0], base[cs], base[2 * cs], base[3 * cs], cell, L.w, L.h, L.aw[a], L.ah[a]);
    }
  }
  if (MODE == MODE_V3) {
    // Classes in groups of 32: the group's logits are loaded together (eight independent loads in flight per thread:
    // the first version's one dependent load per class was latency-bound, 25 us for 5 % of the data), hits are
    // evaluated exactly, and the group's keys leave with ONE global atomic per warp (a shuffle scan of the per-lane
    // counts) instead of one per hit.
    const bool live = idx < anchors_in;
    int l = 0, start = 0, acc = 0;
#pragma unroll
    for (int q = 0; q < kYaMaxLevels; ++q) {
      if (q < p.num_levels) {
        if (idx >= acc) {
          l = q;
          start = acc;
        }
        acc += 3 * p.lv[q].hw;
      }
    }
    const YaLevel& L = p.lv[l];
    const int rel = live ? idx - start : 0;
    const int a = rel / L.hw, cell = rel - a * L.hw;
    const float* base = L.ptr + (int64_t)b * L.batch_stride + (int64_t)(a * attrs) * L.chan_stride + cell;
    const int64_t cs = L.chan_stride;
    float so = 0.0f, cut = INFINITY;
    if (live) {
      so = sigmoid_precise(base[4 * cs]);
      cut = v3_logit_cut(so, p.conf_thres);
    }
    bool boxed = false;
    for (int c0 = 0; c0 < nc; c0 += 32) {
      uint32_t mask = 0;   // classes c0 + i of this anchor that pass the exact test
      const int cn = min(32, nc - c0);
      for (int i0 = 0; i0 < cn; i0 += 8) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = (live && i0 + i < cn) ? __ldg(base + (int64_t)(5 + c0 + i0 + i) * cs) : -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (x[i] >= cut && fmul(so, sigmoid_precise(x[i])) >= p.conf_thres) mask |= 1u << (i0 + i);
      }
      const int mine = __popc(mask);
      int incl = mine;   // inclusive scan over the warp
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total == 0) continue;
      int slot0 = 0;
      if (lane == 31) slot0 = atomicAdd(p.cand_count + ob, total);
      slot0 = __shfl_sync(0xffffffffu, slot0, 31);
      int slot = slot0 + incl - mine;
      while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        const int c = c0 + i;
        const float score = fmul(so, sigmoid_precise(__ldg(base + (int64_t)(5 + c) * cs)));  // the same fp32 value as above
        if (slot < p.max_cand)
          p.cand_key[(int64_t)ob * p.max_cand + slot] = key_pack((uint32_t)c, __float_as_uint(score), (uint32_t)anchor);
        ++slot;
        boxed = true;
      }
    }
    if (boxed)
      p.box_dense[(int64_t)ob * p.A + anchor] = v3_box(base[0], base[cs], base[2 * cs], base[3 * cs], cell, L.w, L.h, L.aw[a], L.ah[a]);
  }
  if (MODE == MODE_V7) {
    const unsigned mk = __ballot_sync(0xffffffffu, o.cand);
    if (mk == 0) continue;
    int slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(p.cand_count + ob, __popc(mk));
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    if (o.cand) {
      const int slot = slot0 + __popc(mk & ((1u << lane) - 1u));
      if (slot < p.max_cand)
        p.cand_key[(int64_t)ob * p.max_cand + slot] = key_pack((uint32_t)o.cls, __float_as_uint(o.score), (uint32_t)anchor);
      p.box_dense[(int64_t)ob * p.A + anchor] = o.box;
      p.aux_dense[(int64_t)ob * p.A + anchor] = make_float2(o.obj, o.cconf);
    }
  }
  }  // unit loop
}

typedef CUresult (*YaEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static YaEncodeTiledFn ya_encode_fn() {
  static YaEncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<YaEncodeTiledFn>(f);
  }();
  return fn;
}

static int ya_pick_cpl() {  // CVPP_YA_CPL=4 selects the 128-cell-tile variant (kept for comparison)
  const char* e = getenv("CVPP_YA_CPL");
  return (e && e[0] == '4') ? 4 : 2;
}

static int ya_pick_warps(int mode, int cpl) {
  const char* e = getenv("CVPP_YA_WARPS");
  const int cap = ya_max_warps(mode, cpl);
  if (e && atoi(e) >= 1 && atoi(e) <= cap) return atoi(e);
  return cap;
}

static size_t ya_smem_bytes(int mode, int warps, int cpl) {
  size_t s = (size_t)warps * kYaStages * kYaChunkRows * 32 * cpl * sizeof(float) + (size_t)warps * kYaStages * sizeof(uint64_t);
  if (mode == MODE_V3) s += (size_t)warps * (kYaHitCap * sizeof(uint64_t) + 128 * sizeof(uint16_t));
  return s;
}

template <int MODE, int CPL>
static int ya_launch_mode(YaParams& stream_p, bool have_stream, YaParams& gen_p, int gen_anchors, int B, size_t smem,
                          const DeviceInfo& di, cudaStream_t stream) {
  const int kYaWarps = ya_pick_warps(MODE, CPL);
  if (have_stream) {
    static unsigned long long attr_done = 0;
    int rc = ensure_smem_attr(reinterpret_cast<const void*>(yolo_anchor_stream_kernel<MODE, CPL>), di.max_smem, di.device, &attr_done);
    if (rc != CVPP_OK) return rc;
    const int want = (stream_p.total_tiles + kYaWarps - 1) / kYaWarps;
    const int grid = want < di.sms ? want : di.sms;
    CVPP_CUDA_TRY(launch_pdl(yolo_anchor_stream_kernel<MODE, CPL>, dim3(grid), dim3(kYaWarps * 32), smem, stream, stream_p));
  }
  if (gen_anchors > 0) {
    // One CTA per unit.  (Running this kernel NEXT TO the streaming one - a programmatic dependent of one 64-thread CTA per SM,
    // which is what fits beside its 768 threads - was measured: 239 us instead of 95; a thread needs ~16 us per anchor, four
    // dependent batches of DRAM loads, so the level needs the 130 K threads of a launch of its own.)
    const int units = ((gen_anchors + 127) / 128) * B;
    yolo_anchor_generic_kernel<MODE><<<(unsigned)units, 128, 0, stream>>>(gen_p, gen_anchors);
    CVPP_CUDA_TRY(cudaGetLastError());
  }
  return CVPP_OK;
}

// mode 0: YOLOv7, mode 1: YOLOv3
int yolo_anchor_decode_launch(int mode, const float* const* level_ptr, const int64_t* batch_stride,
                              const int64_t* chan_stride, const int* level_h, const int* level_w,
                              const float* level_anchors, int num_levels, int B, int nc, int input_h, int input_w,
                              float conf_thres, int merge_batch, uint64_t* cand_key, int32_t* cand_count, float* box_dense,
                              float* aux_dense, int max_cand, int force_generic, cudaStream_t stream) {
  const char* name = mode == MODE_V7 ? "yolov7 decode" : "yolov3 decode";
  if (!level_ptr || !batch_stride || !chan_stride || !level_h || !level_w || !level_anchors || !cand_key || !cand_count ||
      !box_dense || (mode == MODE_V7 && !aux_dense)) {
    set_error("%s: NULL pointer argument", name);
    return CVPP_ERR_INVALID_ARG;
  }
  if (num_levels < 1 || num_levels > kYaMaxLevels || B < 0 || nc < 1 || nc > CVPP_MAX_CLASSES || max_cand < 1 ||
      input_h < 1 || input_w < 1) {
    set_error("%s: bad num_levels=%d / B=%d / nc=%d / max_cand=%d", name, num_levels, B, nc, max_cand);
    return CVPP_ERR_INVALID_ARG;
  }
  if (!(conf_thres >= 0.0f && conf_thres <= 1.0f)) {
    set_error("%s: confidence threshold %f outside [0, 1]", name, conf_thres);
    return CVPP_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(box_dense) & 15u) || (reinterpret_cast<uintptr_t>(aux_dense) & 7u)) {
    set_error("%s: box_dense / aux_dense misaligned", name);
    return CVPP_ERR_ALIGNMENT;
  }
  const bool merged = mode == MODE_V3 && merge_batch;
  const int cpl = mode == MODE_V3 ? 2 : ya_pick_cpl(), tile_a = 32 * cpl;  // (the V3 hit mask is 16 rows x 2 cells)
  YaLevel lv[kYaMaxLevels];
  bool tma_level[kYaMaxLevels];
  int64_t A = 0;
  const int attrs = 5 + nc;
  for (int l = 0; l < num_levels; ++l) {
    if (!level_ptr[l] || level_h[l] < 1 || level_w[l] < 1) {
      set_error("%s: level %d is empty", name, l);
      return CVPP_ERR_INVALID_ARG;
    }
    YaLevel& L = lv[l];
    L.ptr = level_ptr[l];
    L.batch_stride = batch_stride[l];
    L.chan_stride = chan_stride[l];
    L.h = level_h[l];
    L.w = level_w[l];
    L.hw = level_h[l] * level_w[l];
    for (int a = 0; a < 3; ++a) {
      const float pw = level_anchors[(3 * l + a) * 2 + 0], ph = level_anchors[(3 * l + a) * 2 + 1];
      if (mode == MODE_V7) {
        // stride = input / feature size (Python doubles), anchors / stride evaluated in fp32 (yolo_v7.py:254-262)
        const float stride_h = (float)((double)input_h / level_h[l]), stride_w = (float)((double)input_w / level_w[l]);
        L.aw[a] = pw / stride_w;
        L.ah[a] = ph / stride_h;
      } else {
        // generate_yolo3_anchor (core/utils/anchor.py:102-117): fp32 tensor divided in place by w, h
        L.aw[a] = pw / (float)input_w;
        L.ah[a] = ph / (float)input_h;
      }
    }
    L.anchor_off = (int)(merged ? A * B : A);
    L.image_stride = merged ? 3 * L.hw : 0;
    L.tile_off = 0;
    L.tiles_per_anchor = (L.hw + tile_a - 1) / tile_a;
    L.inv_tiles_per_anchor = 1.0f / (float)L.tiles_per_anchor;
    A += 3 * (int64_t)L.hw;
    tma_level[l] = !force_generic && !((reinterpret_cast<uintptr_t>(L.ptr) & 15u) || (L.batch_stride & 3) ||
                                       (L.chan_stride & 3) || (L.hw & 3));
  }
  const int64_t A_out = merged ? A * B : A;
  if (A_out > CVPP_MAX_ANCHORS) {
    set_error("%s: %lld anchors exceed the %d-anchor key field", name, (long long)A_out, CVPP_MAX_ANCHORS);
    return CVPP_ERR_UNSUPPORTED;
  }
  const int n_out = merged ? (B > 0 ? 1 : 0) : B;
  CVPP_CUDA_TRY(zero_counts_async(cand_count, n_out, stream));  // the stream kernel is its programmatic dependent (cvpp_common.cuh)
  if (B == 0) return CVPP_OK;

  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != CVPP_OK) return rc;
  const size_t smem = ya_smem_bytes(mode, ya_pick_warps(mode, cpl), cpl);
  const bool tma_avail = smem <= (size_t)di.max_smem && ya_encode_fn();

  alignas(64) YaParams sp{}, gp{};
  for (YaParams* q : {&sp, &gp}) {
    q->B = B;
    q->nc = nc;
    q->merged = merged ? 1 : 0;
    q->A = A_out;
    q->conf_thres = conf_thres;
    q->cand_key = cand_key;
    q->cand_count = cand_count;
    q->box_dense = reinterpret_cast<float4*>(box_dense);
    q->aux_dense = reinterpret_cast<float2*>(aux_dense);
    q->max_cand = max_cand;
    q->tail_rows = attrs % kYaChunkRows ? attrs % kYaChunkRows : kYaChunkRows;
    q->tile_a = tile_a;
  }
  int gen_anchors = 0, tiles = 0;
  for (int l = 0; l < num_levels; ++l) {
    bool ok = tma_avail && tma_level[l];
    if (ok) {
      const YaLevel& L = lv[l];
      cuuint64_t dims[3] = {(cuuint64_t)L.hw, (cuuint64_t)(3 * attrs), (cuuint64_t)B};
      cuuint64_t strides[2] = {(cuuint64_t)L.chan_stride * 4u, (cuuint64_t)L.batch_stride * 4u};
      if (B == 1) strides[1] = (cuuint64_t)L.chan_stride * 4u * (cuuint64_t)(3 * attrs);
      cuuint32_t box[3] = {(cuuint32_t)tile_a, (cuuint32_t)kYaChunkRows, 1u};
      cuuint32_t estr[3] = {1u, 1u, 1u};
      ok = ya_encode_fn()(&sp.tmap[sp.num_levels], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(L.ptr), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      if (ok) {
        box[1] = (cuuint32_t)sp.tail_rows;
        ok = ya_encode_fn()(&sp.tmap_tail[sp.num_levels], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(L.ptr), dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      }
    }
    if (ok) {
      YaLevel& D = sp.lv[sp.num_levels++];
      D = lv[l];
      D.tile_off = tiles;
      tiles += 3 * D.tiles_per_anchor;
    } else {
      gp.lv[gp.num_levels++] = lv[l];
      gen_anchors += 3 * lv[l].hw;
    }
  }
  sp.tiles_per_image = tiles;
  sp.inv_tiles_per_image = tiles > 0 ? 1.0f / (float)tiles : 0.0f;
  if ((int64_t)tiles * B >= (1 << 24)) {
    set_error("%s: %lld tiles exceed the 2^24 the tile bookkeeping supports: split the batch", name, (long long)tiles * B);
    return CVPP_ERR_UNSUPPORTED;
  }
  sp.total_tiles = tiles * B;
  if (cpl == 4) {
    return ya_launch_mode<MODE_V7, 4>(sp, sp.num_levels > 0, gp, gen_anchors, B, smem, di, stream);
  }
  if (mode == MODE_V7) return ya_launch_mode<MODE_V7, 2>(sp, sp.num_levels > 0, gp, gen_anchors, B, smem, di, stream);
  return ya_launch_mode<MODE_V3, 2>(sp, sp.num_levels > 0, gp, gen_anchors, B, smem, di, stream);
}

// ---- predict_bounding_bbox (dense), core/predict/yolov3_decode.py:12-29 -----------------------------
// One CTA transposes a 32-cell x 3*(5+nc)-channel tile through shared memory: channel rows are read
// coalesced along the cells, the four NHWC outputs are written coalesced along (anchor, attribute).
__global__ void __launch_bounds__(256)
yolov3_predict_bbox_kernel(const float* __restrict__ feature, int nc, int H, int W, float aw0, float ah0, float aw1,
                           float ah1, float aw2, float ah2, float* __restrict__ box_xy, float* __restrict__ box_wh,
                           float* __restrict__ confidence, float* __restrict__ class_prob) {
  extern __shared__ float tile[];  // [3 * attrs][33]
  const int attrs = 5 + nc, C = 3 * attrs, HW = H * W;
  const int b = blockIdx.y, cell0 = blockIdx.x * 32;
  const int ncell = min(32, HW - cell0);
  const float* src = feature + (int64_t)b * C * HW + cell0;
  for (int i = threadIdx.x; i < C * 32; i += blockDim.x) {
    const int ch = i >> 5, x = i & 31;
    if (x < ncell) tile[ch * 33 + x] = __ldg(src + (int64_t)ch * HW + x);
  }
  __syncthreads();
  const float aw[3] = {aw0, aw1, aw2}, ah[3] = {ah0, ah1, ah2};
  const int64_t base = (int64_t)b * HW + cell0;  // first output cell
  // box_xy / box_wh: (cell, a, 2)
  for (int i = threadIdx.x; i < ncell * 6; i += blockDim.x) {
    const int x = i / 6, r = i - x * 6, a = r >> 1, d = r & 1;
    const int cell = cell0 + x;
    const float g = d == 0 ? (float)(cell % W) : (float)(cell / W);
    box_xy[base * 6 + i] = fdiv(fadd(sigmoid_precise(tile[(a * attrs + d) * 33 + x]), g), (float)H);
    box_wh[base * 6 + i] = fmul(expf(tile[(a * attrs + 2 + d) * 33 + x]), d == 0 ? aw[a] : ah[a]);
  }
  for (int i = threadIdx.x; i < ncell * 3; i += blockDim.x) {
    const int x = i / 3, a = i - x * 3;
    confidence[base * 3 + i] = sigmoid_precise(tile[(a * attrs + 4) * 33 + x]);
  }
  const int per_cell = 3 * nc;
  for (int i = threadIdx.x; i < ncell * per_cell; i += blockDim.x) {
    const int x = i / per_cell, r = i - x * per_cell, a = r / nc, c = r - a * nc;
    class_prob[base * per_cell + i] = sigmoid_precise(tile[(a * attrs + 5 + c) * 33 + x]);
  }
}

int yolov3_predict_bbox_launch(const float* feature, int B, int nc, int H, int W, const float* anchors, float* box_xy,
                               float* box_wh, float* confidence, float* class_prob, cudaStream_t stream) {
  if (!feature || !anchors || !box_xy || !box_wh || !confidence || !class_prob) {
    set_error("yolov3_predict_bbox: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || nc < 1 || nc > CVPP_MAX_CLASSES || H < 1 || W < 1) {
    set_error("yolov3_predict_bbox: bad sizes (B=%d nc=%d H=%d W=%d)", B, nc, H, W);
    return CVPP_ERR_INVALID_ARG;
  }
  if (B == 0) return CVPP_OK;
  const size_t smem = (size_t)3 * (5 + nc) * 33 * sizeof(float);
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != CVPP_OK) return rc;
  if (smem > (size_t)di.max_smem) {
    set_error("yolov3_predict_bbox: nc=%d needs %zu bytes of shared memory", nc, smem);
    return CVPP_ERR_UNSUPPORTED;
  }
  static unsigned long long attr_done = 0;
  static int attr_bytes = 0;
  if ((int)smem > attr_bytes) {
    attr_done = 0;
    attr_bytes = (int)smem;
  }
  rc = ensure_smem_attr(reinterpret_cast<const void*>(yolov3_predict_bbox_kernel), attr_bytes, di.device, &attr_done);
  if (rc != CVPP_OK) return rc;
  dim3 grid((unsigned)((H * W + 31) / 32), (unsigned)B);
  yolov3_predict_bbox_kernel<<<grid, 256, smem, stream>>>(feature, nc, H, W, anchors[0], anchors[1], anchors[2], anchors[3],
                                                        anchors[4], anchors[5], box_xy, box_wh, confidence, class_prob);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

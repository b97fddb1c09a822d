// yolov8_decode.cu — kernel 1: fused YOLOv8 head decode + confidence filter (sm_100a).
//
// Replaces (reference file:line): Detect.forward eval tail core/models/yolov8/modules.py:434-445,
// DFL.forward modules.py:80-82, make_anchors core/utils/anchor.py:126-145, dist2bbox
// core/utils/bboxes.py:213-222 and the candidate stage of non_max_suppression
// core/utils/ultralytics_ops.py:190,204,220-226 (+ xywh2xyxy :360-375).
//
// Memory-bound: every image reads (4*reg_max + nc) x A fp32 once (4 838 400 B for the 8400-anchor,
// 80-class head) and writes 8 B + 16 B per surviving candidate.  Design (v2, after the first ncu
// capture showed the CTA-wide-barrier version stalled on barriers at 25 % of HBM peak):
//   * one persistent CTA per SM, ~12 fully independent warps; a warp owns whole tiles of 64
//     consecutive cells of one level of one image (2 cells per lane) and never meets a CTA barrier;
//   * a tile is streamed as chunks of 16 channel rows (one DFL side, or 16 classes) x 64 cells:
//     one elected lane issues ONE 3-D tensor-map TMA load (cp.async.bulk.tensor -> UTMALDG, box
//     64 cells x 16 channels x 1 image; out-of-range cells are zero-filled by the engine) into the
//     warp's private 2-stage shared-memory ring, each stage guarded by an mbarrier armed with the
//     box's byte count (v3: 16 per-row UBLKCPs cost ~160 issue slots per chunk in the v2 capture);
//   * a landed chunk is pulled into registers with conflict-free LDS.64, the stage is re-armed at
//     once for the chunk after next, and the arithmetic (softmax-integral over the 16 DFL bins,
//     running class argmax on logits) runs out of registers while two chunks are in flight;
//   * v4 (launch-shape sweep on B200): the delivered bandwidth PEAKS at about 96 KB of
//     requests in flight per SM (2 stages x 12 warps x 4 KB: 5.7 TB/s) and falls when more is kept in flight
//     (224 KB: 4.9 TB/s - the requests only queue in the memory system), and the last partial round of
//     tiles costs a whole round, so the host picks the warp count that fills the rounds (pick_shape);
//   * sigmoid is evaluated once per cell (it is monotone); exact first-index tie semantics of
//     `cls.max(1)` on sigmoid values are restored on a (rare) slow path;
//   * survivors are compacted with one warp-aggregated atomic per tile into the per-image key list (deferring the use
//     of the atomic's result to the end of the NEXT tile, so that its ~1 us round trip is never waited for, was
//     measured in round 2: no change - 54.0 us either way - and 14 more registers, so it is not in).
#include <cuda.h>
#include <cstdlib>

#include "cvpp_common.cuh"
#include "yolov8_cell.cuh"

namespace cvpp {

constexpr int kChunkRows = 16;   // channel rows per chunk
constexpr int kMaxWarps = 16;    // independent warps per CTA (the host picks 9..14, see pick_warps)

struct LevelDesc {
  const float* ptr;
  int64_t batch_stride;
  int64_t chan_stride;
  int hw;
  int w;
  float stride;
  int anchor_off;  // first anchor index of this level
  int tile_off;    // first tile index (within an image) of this level
};

struct DecodeParams {
  CUtensorMap tmap[CVPP_MAX_LEVELS];  // (cell, channel, image) fp32, box 128 x 16 x 1; first: 64-byte aligned
  LevelDesc lv[CVPP_MAX_LEVELS];
  int num_levels;
  int B;
  int nc;
  int A;
  int tiles_per_image;
  int total_tiles;
  float conf_thres;
  uint64_t* cand_key;
  int32_t* cand_count;
  float4* box_dense;
  int max_cand;
  float* y;  // FULL mode only
};

// exact `conf, j = cls.max(1)` over sigmoid values: first index of the maximum sigmoid
__device__ __noinline__ void class_argmax_sigmoid(const float* col, int64_t cs, int nc, float& s, int& arg) {
  float bs = -1.0f;
  int ba = 0;
  for (int c = 0; c < nc; ++c) {
    float v = sigmoid_precise(col[(int64_t)c * cs]);
    if (v > bs) {
      bs = v;
      ba = c;
    }
  }
  s = bs;
  arg = ba;
}

// warp-aggregated append of one candidate per flagged lane (all 32 lanes must call)
__device__ __forceinline__ void emit_candidate(bool flag, int b, uint64_t key, int anchor, const float4& box,
                                               const DecodeParams& p) {
  unsigned mask = __ballot_sync(0xffffffffu, flag);
  if (mask == 0) return;
  int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(p.cand_count + b, __popc(mask));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (flag) {
    int slot = base + __popc(mask & ((1u << lane) - 1u));
    if (slot < p.max_cand) p.cand_key[(int64_t)b * p.max_cand + slot] = key;
    p.box_dense[(int64_t)b * p.A + anchor] = box;
  }
}

__device__ __forceinline__ void tile_info(const DecodeParams& p, int tile_a, int g, int& b, int& l, int& cell0, int& nA) {
  b = g / p.tiles_per_image;
  int j = g - b * p.tiles_per_image;
  l = 0;
#pragma unroll
  for (int q = 1; q < CVPP_MAX_LEVELS; ++q)
    if (q < p.num_levels && j >= p.lv[q].tile_off) l = q;
  cell0 = (j - p.lv[l].tile_off) * tile_a;
  nA = min(tile_a, p.lv[l].hw - cell0);
}

// score / class / candidate decision for one cell, shared by both kernels.
// `col` points at the cell's first class logit in GLOBAL memory (only touched on the tie path).
__device__ __forceinline__ bool finalize_cell(float best, int arg_in, float prev, const float* col, int64_t cs, int nc,
                                              float conf_thres, float& score, int& arg) {
  score = sigmoid_precise(best);
  arg = arg_in;
  const bool cand = score > conf_thres;
  // an EARLIER class whose sigmoid rounds to the same float: the reference takes the FIRST index of
  // the maximum sigmoid value, which need not be the first maximum logit.
  if (cand && sigmoid_precise(prev) >= score) class_argmax_sigmoid(col, cs, nc, score, arg);
  return cand;
}

// -----------------------------------------------------------------------------------------------
// TMA-streamed persistent kernel: independent warps, private tensor-map TMA rings
// -----------------------------------------------------------------------------------------------
// The head outputs are read exactly once: the loads carry an L2 evict-first policy so that 310 MB of streamed
// input does not push the candidate keys, the dense boxes and the NMS kernel's CODE out of L2.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}

// CPL = cells per lane.  CPL = 2 (default): 64-cell tiles, 4 KB chunks, LDS.64, ~90 registers.  CPL = 4
// (CVPP_DECODE_CPL=4, kept for comparison): 128-cell tiles, 8 KB chunks, LDS.128, ~125 registers.
template <int CPL>
struct VecOf;
template <>
struct VecOf<4> {
  typedef float4 type;
};
template <>
struct VecOf<2> {
  typedef float2 type;
};
template <int CPL>
__device__ __forceinline__ void vec_unpack(const typename VecOf<CPL>::type& q, float (&o)[CPL]);
template <>
__device__ __forceinline__ void vec_unpack<4>(const float4& q, float (&o)[4]) {
  o[0] = q.x, o[1] = q.y, o[2] = q.z, o[3] = q.w;
}
template <>
__device__ __forceinline__ void vec_unpack<2>(const float2& q, float (&o)[2]) {
  o[0] = q.x, o[1] = q.y;
}
template <int CPL>
__device__ __forceinline__ typename VecOf<CPL>::type vec_pack(const float (&o)[CPL]);
template <>
__device__ __forceinline__ float4 vec_pack<4>(const float (&o)[4]) {
  return make_float4(o[0], o[1], o[2], o[3]);
}
template <>
__device__ __forceinline__ float2 vec_pack<2>(const float (&o)[2]) {
  return make_float2(o[0], o[1]);
}

template <bool FULL, int CPL, int STAGES>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) yolov8_decode_stream_kernel(const __grid_constant__ DecodeParams p) {
  constexpr int kStages = STAGES;
  // programmatic dependent launch: the sort+NMS kernel that follows in the stream may be scheduled now (its CTAs park on
  // griddepcontrol.wait until this grid has completed and flushed): its launch latency and prologue leave the critical path
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int Warps = blockDim.x >> 5;  // chosen by the host so that the tiles of an SM fill whole rounds
  constexpr int TileA = 32 * CPL;            // cells per tile
  constexpr int ChunkFloats = kChunkRows * TileA;
  typedef typename VecOf<CPL>::type vec_t;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ring = reinterpret_cast<float*>(smem_raw) + (size_t)warp * (kStages * ChunkFloats);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)Warps * kStages * ChunkFloats * sizeof(float)) +
                  warp * kStages;
  const int nc = p.nc;
  const int C = 4 * kRegMax + nc;
  const int nchunks = (C + kChunkRows - 1) / kChunkRows;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  __syncwarp();

  // tiles of this warp: first, first + stride, ...  (CTA-major so small batches spread over all SMs)
  const int stride_tiles = gridDim.x * Warps;
  const int first = blockIdx.x * Warps + warp;  // the warps of a CTA stream ADJACENT tiles: their row pieces share DRAM pages
  // ... except in the last, partial round, whose tiles are dealt SM-minor so that every SM keeps streaming
  const int full_rounds = p.total_tiles / stride_tiles;
  const int last_slot = warp * gridDim.x + blockIdx.x;
  const int n_tiles = full_rounds + (full_rounds * stride_tiles + last_slot < p.total_tiles ? 1 : 0);
  auto tile_of = [&](int r) { return r * stride_tiles + (r < full_rounds ? first : last_slot); };
  const int total_q = n_tiles * nchunks;

  const uint64_t policy = l2_evict_first_policy();
  // producer cursor (runs kStages chunks ahead of the consumer cursor)
  int pq = 0, pj = 0, pr = 0, pb = 0, pl = 0, pcell0 = 0;
  auto issue = [&]() {
    if (pj == 0) {
      int nA_unused;
      tile_info(p, TileA, tile_of(pr), pb, pl, pcell0, nA_unused);
    }
    if (lane == 0) {
      uint64_t* fb = &bar[pq % kStages];
      mbar_arrive_expect_tx(fb, (uint32_t)(ChunkFloats * sizeof(float)));
      tma_load_3d(ring + (pq % kStages) * ChunkFloats, &p.tmap[pl], pcell0, kChunkRows * pj, pb, fb, policy);
    }
    ++pq;
    if (++pj == nchunks) {
      pj = 0;
      ++pr;
    }
  };
  for (int q = 0; q < kStages && q < total_q; ++q) issue();

  // per-lane state of the current tile: CPL cells
  float d[4][CPL];  // [side][cell]
  float best[CPL], prev[CPL];
  int arg[CPL];
  int b = 0, l = 0, cell0 = 0, nA = 0;
  int j = 0, r = 0;
  for (int q = 0; q < total_q; ++q) {
    if (j == 0) {
      tile_info(p, TileA, tile_of(r), b, l, cell0, nA);
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        best[k] = -INFINITY;
        prev[k] = -INFINITY;
        arg[k] = 0;
      }
    }
    const LevelDesc& L = p.lv[l];
    const int s = q % kStages;
    mbar_wait(&bar[s], (uint32_t)(q / kStages) & 1u);
    float v[kChunkRows][CPL];
    {
      const vec_t* src = reinterpret_cast<const vec_t*>(ring + s * ChunkFloats) + lane;
#pragma unroll
      for (int r = 0; r < kChunkRows; ++r) vec_unpack<CPL>(src[r * 32], v[r]);
    }
    __syncwarp();  // every lane holds its copy: the stage may be refilled
    if (pq < total_q) issue();

    const bool active = CPL * lane < nA;
    const int anchor0 = L.anchor_off + cell0 + CPL * lane;
    if (j < 4) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        float t[kRegMax];
#pragma unroll
        for (int r = 0; r < kRegMax; ++r) t[r] = v[r][k];
        const float dk = dfl16(t);
#pragma unroll
        for (int side = 0; side < 4; ++side)
          if (j == side) d[side][k] = dk;
      }
    } else {
      const int c0 = kChunkRows * (j - 4);
      const int rows = min(kChunkRows, nc - c0);
      if (FULL) {
        if (active) {
          float* yc = p.y + ((int64_t)b * (4 + nc) + 4 + c0) * p.A + anchor0;
#pragma unroll
          for (int r = 0; r < kChunkRows; ++r) {
            if (r < rows) {
              float o[CPL];
#pragma unroll
              for (int k = 0; k < CPL; ++k) o[k] = sigmoid_precise(v[r][k]);
              *reinterpret_cast<vec_t*>(yc + (int64_t)r * p.A) = vec_pack<CPL>(o);
            }
          }
        }
      } else if (rows == kChunkRows) {
#pragma unroll
        for (int r = 0; r < kChunkRows; ++r) {
#pragma unroll
          for (int k = 0; k < CPL; ++k) class_step(v[r][k], c0 + r, best[k], arg[k], prev[k]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < kChunkRows; ++r) {
          if (r < rows) {
#pragma unroll
            for (int k = 0; k < CPL; ++k) class_step(v[r][k], c0 + r, best[k], arg[k], prev[k]);
          }
        }
      }
    }

    if (++j == nchunks) {  // tile complete
      j = 0;
      ++r;
      CellBox box[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) box[k] = cell_box(cell0 + CPL * lane + k, L.w, L.stride, d[0][k], d[1][k], d[2][k], d[3][k]);
      if (FULL) {
        if (active) {
          float* yb = p.y + (int64_t)b * (4 + nc) * p.A + anchor0;
          float o[CPL];
#pragma unroll
          for (int k = 0; k < CPL; ++k) o[k] = box[k].cx;
          *reinterpret_cast<vec_t*>(yb) = vec_pack<CPL>(o);
#pragma unroll
          for (int k = 0; k < CPL; ++k) o[k] = box[k].cy;
          *reinterpret_cast<vec_t*>(yb + (int64_t)p.A) = vec_pack<CPL>(o);
#pragma unroll
          for (int k = 0; k < CPL; ++k) o[k] = box[k].w;
          *reinterpret_cast<vec_t*>(yb + 2 * (int64_t)p.A) = vec_pack<CPL>(o);
#pragma unroll
          for (int k = 0; k < CPL; ++k) o[k] = box[k].h;
          *reinterpret_cast<vec_t*>(yb + 3 * (int64_t)p.A) = vec_pack<CPL>(o);
        }
      } else {
        const float* col = L.ptr + (int64_t)b * L.batch_stride + (int64_t)(4 * kRegMax) * L.chan_stride + cell0 + CPL * lane;
        bool cand[CPL];
        float score[CPL];
        int cls[CPL];
        unsigned m[CPL];
        int total = 0;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          cand[k] = false;
          score[k] = 0.f;
          cls[k] = 0;
          if (active) cand[k] = finalize_cell(best[k], arg[k], prev[k], col + k, L.chan_stride, nc, p.conf_thres, score[k], cls[k]);
          m[k] = __ballot_sync(0xffffffffu, cand[k]);
          total += __popc(m[k]);
        }
        if (total) {
          pdl_wait();  // the counts have been zeroed (cvpp_zero_i32_kernel, the programmatic predecessor); free after the first time
          int base = 0;
          if (lane == 0) base = atomicAdd(p.cand_count + b, total);
          base = __shfl_sync(0xffffffffu, base, 0);
          const unsigned lt = (1u << lane) - 1u;
#pragma unroll
          for (int k = 0; k < CPL; ++k) {
            if (cand[k]) {
              const int slot = base + __popc(m[k] & lt);
              const int anchor = anchor0 + k;
              if (slot < p.max_cand)
                p.cand_key[(int64_t)b * p.max_cand + slot] = key_pack((uint32_t)cls[k], __float_as_uint(score[k]), (uint32_t)anchor);
              p.box_dense[(int64_t)b * p.A + anchor] = make_float4(box[k].x1, box[k].y1, box[k].x2, box[k].y2);
            }
            base += __popc(m[k]);
          }
        }
      }
    }
  }
}

// -----------------------------------------------------------------------------------------------
// Generic kernel (no alignment / shape requirements): one thread per cell, straight from global.
// Used when a level's H*W or strides are not multiples of 4 floats or the tile ring would not fit.
// -----------------------------------------------------------------------------------------------
template <bool FULL>
__global__ void __launch_bounds__(128) yolov8_decode_generic_kernel(const __grid_constant__ DecodeParams p) {
  const int nc = p.nc;
  const int b = blockIdx.y;
  const int anchor = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = anchor < p.A;
  int l = 0;
#pragma unroll
  for (int q = 1; q < CVPP_MAX_LEVELS; ++q)
    if (q < p.num_levels && anchor >= p.lv[q].anchor_off) l = q;
  const LevelDesc& L = p.lv[l];
  const int cell = anchor - L.anchor_off;
  bool cand = false;
  float score = 0.0f;
  int arg = 0;
  CellBox box;
  if (valid) {
    const float* col = L.ptr + (int64_t)b * L.batch_stride + cell;
    const int64_t cs = L.chan_stride;
    float d[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
      float t[kRegMax];
#pragma unroll
      for (int k = 0; k < kRegMax; ++k) t[k] = col[(int64_t)(side * kRegMax + k) * cs];
      d[side] = dfl16(t);
    }
    box = cell_box(cell, L.w, L.stride, d[0], d[1], d[2], d[3]);
    const float* ccol = col + (int64_t)4 * kRegMax * cs;
    if (FULL) {
      float* yb = p.y + (int64_t)b * (4 + nc) * p.A + anchor;
      yb[0] = box.cx;
      yb[(int64_t)p.A] = box.cy;
      yb[2 * (int64_t)p.A] = box.w;
      yb[3 * (int64_t)p.A] = box.h;
      for (int c = 0; c < nc; ++c) yb[(int64_t)(4 + c) * p.A] = sigmoid_precise(ccol[(int64_t)c * cs]);
    } else {
      float best = -INFINITY, prev = -INFINITY;
      int a0 = 0;
      for (int c = 0; c < nc; ++c) class_step(ccol[(int64_t)c * cs], c, best, a0, prev);
      cand = finalize_cell(best, a0, prev, ccol, cs, nc, p.conf_thres, score, arg);
    }
  }
  if (!FULL) {
    uint64_t key = key_pack((uint32_t)arg, __float_as_uint(score), (uint32_t)anchor);
    emit_candidate(cand, b, key, anchor, make_float4(box.x1, box.y1, box.x2, box.y2), p);
  }
}

// -----------------------------------------------------------------------------------------------
// host launcher
// -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry-point lookup: no link-time libcuda dependency.
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// (cell, channel, image) view of one level; box = 128 cells x 16 channels x 1 image, zero fill out of range
static bool make_level_tmap(CUtensorMap* tm, const LevelDesc& L, int C, int B, int tile_a) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)L.hw, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)L.chan_stride * 4u, (cuuint64_t)L.batch_stride * 4u};
  cuuint32_t box[3] = {(cuuint32_t)tile_a, (cuuint32_t)kChunkRows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  if (B == 1) strides[1] = (cuuint64_t)L.chan_stride * 4u * (cuuint64_t)C;  // unused dimension: any valid stride
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(L.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static int decode_cpl() {  // cells per lane of the streaming kernel (CVPP_DECODE_CPL=4 selects the 14-warp variant)
  const char* e = getenv("CVPP_DECODE_CPL");
  return (e && e[0] == '4') ? 4 : 2;
}

// Launch shape of the streaming kernel.  A warp streams whole tiles, so an SM with T tiles and W warps runs
// ceil(T / W) rounds and the last one may be mostly idle: W is chosen in 9..14 to fill the rounds (measured at
// C2: 2 stages x 12 warps x 4 KB = 96 KB in flight per SM is the sweet spot - more requests in flight only queue
// in the memory system and lower the delivered bandwidth).  Small batches get few warps per CTA, spread over
// every SM, and a deeper ring (latency- rather than bandwidth-bound).  CVPP_DECODE_STAGES / CVPP_DECODE_WARPS
// override the choice (tuning knobs, tools/bench_paths.py).
static void pick_shape(int total_tiles, int sms, int cpl, int* stages, int* warps, int* grid) {
  const char* es = getenv("CVPP_DECODE_STAGES");
  const char* ew = getenv("CVPP_DECODE_WARPS");
  if (cpl == 4) {
    *stages = 2;
    *warps = 14;
  } else if (total_tiles <= 4 * sms) {
    *stages = 6;
    *warps = (total_tiles + sms - 1) / sms;
  } else {
    const int T = (total_tiles + sms - 1) / sms;
    int best_w = 12, best_slots = ((T + 11) / 12) * 12;
    for (int w = 9; w <= 14; ++w) {
      const int slots = ((T + w - 1) / w) * w;
      if (slots < best_slots) {
        best_slots = slots;
        best_w = w;
      }
    }
    *stages = 2;
    *warps = best_w;
  }
  if (es && cpl != 4) *stages = atoi(es) == 3 ? 3 : atoi(es) == 6 ? 6 : 2;
  if (ew && atoi(ew) >= 1 && atoi(ew) <= kMaxWarps) *warps = atoi(ew);
  const int g = (total_tiles + *warps - 1) / *warps;
  *grid = g < sms ? g : sms;
}

template <bool FULL, int CPL, int STAGES>
static int launch_stream(DecodeParams& p, const DeviceInfo& di, int grid, int warps, size_t smem, cudaStream_t stream) {
  auto kern = yolov8_decode_stream_kernel<FULL, CPL, STAGES>;
  static unsigned long long attr_done = 0;  // one per template instantiation
  int rc = ensure_smem_attr(reinterpret_cast<const void*>(kern), di.max_smem, di.device, &attr_done);
  if (rc != CVPP_OK) return rc;
  if (FULL) {
    kern<<<grid, warps * 32, smem, stream>>>(p);   // ordinary launch: the kernel reads its inputs at once
  } else {
    // programmatic dependent of zero_counts_kernel (which was launched normally, i.e. after everything earlier in the stream
    // had completed, and triggers at once): the inputs may be read before griddepcontrol.wait, the counts may not
    CVPP_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(warps * 32), smem, stream, p));
  }
  return CVPP_OK;
}

template <bool FULL>
static int launch_decode(DecodeParams& p, bool tma_ok, cudaStream_t stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != CVPP_OK) return rc;
  const int max_smem = di.max_smem;
  const int cpl = decode_cpl();
  const int tile_a = 32 * cpl;
  // tiling of every level for the chosen tile width
  int tiles = 0;
  for (int l = 0; l < p.num_levels; ++l) {
    p.lv[l].tile_off = tiles;
    tiles += (p.lv[l].hw + tile_a - 1) / tile_a;
  }
  p.tiles_per_image = tiles;
  p.total_tiles = tiles * p.B;
  int stages, warps, grid;
  pick_shape(p.total_tiles, di.sms, cpl, &stages, &warps, &grid);
  const size_t smem = (size_t)warps * stages * kChunkRows * tile_a * sizeof(float) + (size_t)warps * stages * sizeof(uint64_t);
  if (tma_ok && smem <= (size_t)max_smem) {
    const int C = 4 * kRegMax + p.nc;
    for (int l = 0; l < p.num_levels && tma_ok; ++l) tma_ok = make_level_tmap(&p.tmap[l], p.lv[l], C, p.B, tile_a);
  } else {
    tma_ok = false;
  }
  if (tma_ok) {
    if (cpl == 4) rc = launch_stream<FULL, 4, 2>(p, di, grid, warps, smem, stream);
    else if (stages == 2) rc = launch_stream<FULL, 2, 2>(p, di, grid, warps, smem, stream);
    else if (stages == 3) rc = launch_stream<FULL, 2, 3>(p, di, grid, warps, smem, stream);
    else rc = launch_stream<FULL, 2, 6>(p, di, grid, warps, smem, stream);
    if (rc != CVPP_OK) return rc;
  } else {
    dim3 grid((p.A + 127) / 128, p.B);
    yolov8_decode_generic_kernel<FULL><<<grid, 128, 0, stream>>>(p);
  }
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

int yolov8_decode_launch(const float* const* level_ptr, const int64_t* batch_stride, const int64_t* chan_stride,
                         const int* level_h, const int* level_w, const float* level_stride, int num_levels, int B,
                         int nc, int reg_max, float conf_thres, uint64_t* cand_key, int32_t* cand_count,
                         float* box_dense, int max_cand, float* y, int force_generic, cudaStream_t stream) {
  if (!level_ptr || !batch_stride || !chan_stride || !level_h || !level_w || !level_stride) {
    set_error("yolov8 decode: NULL level description");
    return CVPP_ERR_INVALID_ARG;
  }
  if (num_levels < 1 || num_levels > CVPP_MAX_LEVELS || B < 0 || nc < 1 || nc > CVPP_MAX_CLASSES) {
    set_error("yolov8 decode: bad num_levels=%d / B=%d / nc=%d", num_levels, B, nc);
    return CVPP_ERR_INVALID_ARG;
  }
  if (reg_max != kRegMax) {
    set_error("yolov8 decode: reg_max=%d is not compiled in (reference hard-codes 16, modules.py:413)", reg_max);
    return CVPP_ERR_UNSUPPORTED;
  }
  const bool full = y != nullptr;
  if (!full && (!cand_key || !cand_count || !box_dense || max_cand < 1)) {
    set_error("yolov8 decode: NULL output / max_cand < 1");
    return CVPP_ERR_INVALID_ARG;
  }
  alignas(64) DecodeParams p{};
  p.num_levels = num_levels;
  p.B = B;
  p.nc = nc;
  p.conf_thres = conf_thres;
  p.cand_key = cand_key;
  p.cand_count = cand_count;
  p.box_dense = reinterpret_cast<float4*>(box_dense);
  p.max_cand = max_cand;
  p.y = y;
  bool tma_ok = !force_generic;
  int64_t A = 0;
  int tiles = 0;
  for (int l = 0; l < num_levels; ++l) {
    if (!level_ptr[l] || level_h[l] < 1 || level_w[l] < 1) {
      set_error("yolov8 decode: level %d is empty", l);
      return CVPP_ERR_INVALID_ARG;
    }
    LevelDesc& L = p.lv[l];
    L.ptr = level_ptr[l];
    L.batch_stride = batch_stride[l];
    L.chan_stride = chan_stride[l];
    L.hw = level_h[l] * level_w[l];
    L.w = level_w[l];
    L.stride = level_stride[l];
    L.anchor_off = (int)A;
    L.tile_off = tiles;
    A += L.hw;
    if ((reinterpret_cast<uintptr_t>(L.ptr) & 15u) || (L.batch_stride & 3) || (L.chan_stride & 3) || (L.hw & 3))
      tma_ok = false;  // bulk copies need 16-byte aligned rows
  }
  if (A > CVPP_MAX_ANCHORS) {
    set_error("yolov8 decode: %lld anchors exceed the %d-anchor key field", (long long)A, CVPP_MAX_ANCHORS);
    return CVPP_ERR_UNSUPPORTED;
  }
  if (!full && (reinterpret_cast<uintptr_t>(box_dense) & 15u)) {
    set_error("yolov8 decode: box_dense must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  p.A = (int)A;
  p.tiles_per_image = tiles;
  p.total_tiles = tiles * B;
  if (B == 0) return CVPP_OK;
  if (!full) {
    CVPP_CUDA_TRY(zero_counts_async(cand_count, B, stream));   // (see cvpp_common.cuh: the decode kernel is its programmatic dependent)
  }
  return full ? launch_decode<true>(p, tma_ok, stream) : launch_decode<false>(p, tma_ok, stream);
}

}  // namespace cvpp

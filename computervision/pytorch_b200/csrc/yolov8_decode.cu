// yolov8_decode.cu — kernel 1: fused YOLOv8 head decode + confidence filter (sm_100a).
//
// Replaces (reference file:line): Detect.forward eval tail core/models/yolov8/modules.py:434-445,
// DFL.forward modules.py:80-82, make_anchors core/utils/anchor.py:126-145, dist2bbox
// core/utils/bboxes.py:213-222 and the candidate stage of non_max_suppression
// core/utils/ultralytics_ops.py:190,204,220-226 (+ xywh2xyxy :360-375).
//
// Memory-bound: every image reads (4*reg_max + nc) x A fp32 once (4 838 400 B for the 8400-anchor,
// 80-class head) and writes 8 B + 16 B per surviving candidate.  Design:
//   * persistent CTAs (one per SM) walk tiles of TILE_A consecutive cells of one level of one image;
//   * a tile is C rows of TILE_A*4 contiguous bytes; warp 0 issues one 1-D bulk async copy
//     (cp.async.bulk -> UBLKCP, the TMA engine) per row into a STAGES-deep shared-memory ring, each
//     stage guarded by an mbarrier armed with the tile's byte count, so 2 tiles are always in
//     flight while one is being consumed;
//   * 16 consumer warps: thread (part, a) handles DFL side `part` (softmax-integral over 16 bins)
//     and a quarter of the class logits of cell a; class argmax is done on logits (sigmoid is
//     monotone) and sigmoid is evaluated once per cell; exact first-index tie semantics of
//     `cls.max(1)` on the sigmoid values are restored on a (rare) slow path;
//   * survivors are compacted with warp-aggregated atomics into the per-image key list.
#include "cvpp_common.cuh"

namespace cvpp {

constexpr int kTileA = 128;     // cells per tile
constexpr int kRegMax = 16;     // DFL bins (reference hard-codes 16, modules.py:413)
constexpr int kThreads = 512;   // 4 parts x 128 cells
constexpr int kParts = 4;

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

// ---- per-cell arithmetic (Appendix B of SURVEY.md: one fp32 rounding per reference op) ----------
// softmax over the 16 bins followed by the arange(16) 1x1 conv: sum_k k * softmax(x)_k
__device__ __forceinline__ float dfl_expectation(const float* col, int64_t cs) {
  float v[kRegMax];
#pragma unroll
  for (int k = 0; k < kRegMax; ++k) v[k] = col[k * cs];
  float m = v[0];
#pragma unroll
  for (int k = 1; k < kRegMax; ++k) m = fmaxf(m, v[k]);
  float sum = 0.0f, wsum = 0.0f;
#pragma unroll
  for (int k = 0; k < kRegMax; ++k) {
    float e = __expf(v[k] - m);
    sum += e;
    wsum = fmaf((float)k, e, wsum);
  }
  return fdiv(wsum, sum);
}

// running (best, first-argmax, runner-up) over class logits [c0, c1)
__device__ __forceinline__ void class_scan(const float* col, int64_t cs, int c0, int c1, float& best, int& arg,
                                           float& sec) {
#pragma unroll 4
  for (int c = c0; c < c1; ++c) {
    float x = col[(int64_t)(c - c0) * cs];
    if (x > best) {
      sec = best;
      best = x;
      arg = c;
    } else {
      sec = fmaxf(sec, x);
    }
  }
}

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

__device__ __forceinline__ void tile_info(const DecodeParams& p, int g, int& b, int& l, int& cell0, int& nA) {
  b = g / p.tiles_per_image;
  int j = g - b * p.tiles_per_image;
  l = 0;
#pragma unroll
  for (int q = 1; q < CVPP_MAX_LEVELS; ++q)
    if (q < p.num_levels && j >= p.lv[q].tile_off) l = q;
  cell0 = (j - p.lv[l].tile_off) * kTileA;
  nA = min(kTileA, p.lv[l].hw - cell0);
}

// -----------------------------------------------------------------------------------------------
// TMA-staged persistent kernel
// -----------------------------------------------------------------------------------------------
template <bool FULL>
__global__ void __launch_bounds__(kThreads, 1)
yolov8_decode_tma_kernel(const __grid_constant__ DecodeParams p, const int stages) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nc = p.nc;
  const int C = 4 * kRegMax + nc;
  const int tile_floats = C * kTileA;
  float* tiles = reinterpret_cast<float*>(smem_raw);       // [stages][C][kTileA]
  float* part_d = tiles + (size_t)stages * tile_floats;    // [4][kTileA]
  float* part_max = part_d + kParts * kTileA;
  float* part_sec = part_max + kParts * kTileA;
  int* part_arg = reinterpret_cast<int*>(part_sec + kParts * kTileA);
  uint64_t* full = reinterpret_cast<uint64_t*>(part_arg + kParts * kTileA);  // [stages]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = warp >> 2;
  const int a = ((warp & 3) << 5) | lane;
  const int ncq = (nc + kParts - 1) / kParts;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int n_my = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  auto issue = [&](int i) {  // warp 0 only
    int g = blockIdx.x + i * gridDim.x;
    int s = i % stages;
    int b, l, cell0, nA;
    tile_info(p, g, b, l, cell0, nA);
    const LevelDesc& L = p.lv[l];
    const float* src = L.ptr + (int64_t)b * L.batch_stride + cell0;
    const uint32_t row_bytes = (uint32_t)nA * 4u;
    if (lane == 0) mbar_arrive_expect_tx(&full[s], row_bytes * (uint32_t)C);
    __syncwarp();
    float* dst = tiles + (size_t)s * tile_floats;
    for (int c = lane; c < C; c += 32) bulk_g2s(dst + c * kTileA, src + (int64_t)c * L.chan_stride, row_bytes, &full[s]);
  };

  if (warp == 0) {
    int pre = n_my < stages ? n_my : stages;
    for (int i = 0; i < pre; ++i) issue(i);
  }

  for (int i = 0; i < n_my; ++i) {
    const int s = i % stages;
    const uint32_t parity = (uint32_t)(i / stages) & 1u;
    int b, l, cell0, nA;
    tile_info(p, blockIdx.x + i * gridDim.x, b, l, cell0, nA);
    const LevelDesc& L = p.lv[l];
    const bool valid = a < nA;
    const int anchor = L.anchor_off + cell0 + a;

    mbar_wait(&full[s], parity);
    const float* T = tiles + (size_t)s * tile_floats;

    float d = 0.0f, best = -INFINITY, sec = -INFINITY;
    int arg = 0;
    if (valid) {
      d = dfl_expectation(T + (part * kRegMax) * kTileA + a, kTileA);
      const int c0 = part * ncq, c1 = min(nc, c0 + ncq);
      const float* ccol = T + (4 * kRegMax + c0) * kTileA + a;
      class_scan(ccol, kTileA, c0, c1, best, arg, sec);
      if (FULL) {
        float* yc = p.y + ((int64_t)b * (4 + nc) + 4 + c0) * p.A + anchor;
        for (int c = c0; c < c1; ++c) yc[(int64_t)(c - c0) * p.A] = sigmoid_precise(ccol[(c - c0) * kTileA]);
      }
    }
    part_d[part * kTileA + a] = d;
    if (!FULL) {
      part_max[part * kTileA + a] = best;
      part_sec[part * kTileA + a] = sec;
      part_arg[part * kTileA + a] = arg;
    }
    __syncthreads();

    if (part == 0) {
      CellBox box;
      if (valid)
        box = cell_box(cell0 + a, L.w, L.stride, part_d[a], part_d[kTileA + a], part_d[2 * kTileA + a],
                       part_d[3 * kTileA + a]);
      if (FULL) {
        if (valid) {
          float* yb = p.y + (int64_t)b * (4 + nc) * p.A + anchor;
          yb[0] = box.cx;
          yb[(int64_t)p.A] = box.cy;
          yb[2 * (int64_t)p.A] = box.w;
          yb[3 * (int64_t)p.A] = box.h;
        }
      } else {
        bool cand = false;
        float score = 0.0f;
        if (valid) {
#pragma unroll
          for (int q = 1; q < kParts; ++q) {
            float bq = part_max[q * kTileA + a];
            if (bq > best) {
              sec = fmaxf(sec, best);
              best = bq;
              arg = part_arg[q * kTileA + a];
            } else {
              sec = fmaxf(sec, bq);
            }
            sec = fmaxf(sec, part_sec[q * kTileA + a]);
          }
          score = sigmoid_precise(best);
          cand = score > p.conf_thres;
          // another class whose sigmoid rounds to the same float: the reference takes the FIRST
          // index of the maximum sigmoid value, which need not be the first maximum logit.
          if (cand && sigmoid_precise(sec) >= score)
            class_argmax_sigmoid(T + (4 * kRegMax) * kTileA + a, kTileA, nc, score, arg);
        }
        uint64_t key = key_pack((uint32_t)arg, __float_as_uint(score), (uint32_t)anchor);
        emit_candidate(cand, b, key, anchor, make_float4(box.x1, box.y1, box.x2, box.y2), p);
      }
    }
    __syncthreads();  // every read of stage s (and of the part_* arrays) is done
    if (warp == 0 && i + stages < n_my) issue(i + stages);
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
    for (int side = 0; side < 4; ++side) d[side] = dfl_expectation(col + (int64_t)side * kRegMax * cs, cs);
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
      float best = -INFINITY, sec = -INFINITY;
      class_scan(ccol, cs, 0, nc, best, arg, sec);
      score = sigmoid_precise(best);
      cand = score > p.conf_thres;
      if (cand && sigmoid_precise(sec) >= score) class_argmax_sigmoid(ccol, cs, nc, score, arg);
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
static int sm_count_of_current_device(int* sms, int* max_smem) {
  int dev = 0;
  CVPP_CUDA_TRY(cudaGetDevice(&dev));
  CVPP_CUDA_TRY(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  CVPP_CUDA_TRY(cudaDeviceGetAttribute(max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return CVPP_OK;
}

template <bool FULL>
static int launch_decode(DecodeParams& p, bool tma_ok, cudaStream_t stream) {
  int sms = 0, max_smem = 0;
  int rc = sm_count_of_current_device(&sms, &max_smem);
  if (rc != CVPP_OK) return rc;
  const int C = 4 * kRegMax + p.nc;
  const size_t tile_bytes = (size_t)C * kTileA * sizeof(float);
  const size_t fixed = (size_t)4 * kParts * kTileA * sizeof(float) + 8 * sizeof(uint64_t);
  int stages = 0;
  if (tma_ok && (size_t)max_smem > fixed + 2 * tile_bytes) {
    stages = (int)(((size_t)max_smem - fixed) / tile_bytes);
    if (stages > 4) stages = 4;
  }
  if (stages >= 2) {
    const size_t smem = fixed + (size_t)stages * tile_bytes;
    auto kern = yolov8_decode_tma_kernel<FULL>;
    CVPP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = p.total_tiles < sms ? p.total_tiles : sms;
    kern<<<grid, kThreads, smem, stream>>>(p, stages);
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
  DecodeParams p{};
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
    tiles += (L.hw + kTileA - 1) / kTileA;
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
  if (!full) CVPP_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(int32_t) * (size_t)B, stream));
  if (B == 0) return CVPP_OK;
  return full ? launch_decode<true>(p, tma_ok, stream) : launch_decode<false>(p, tma_ok, stream);
}

}  // namespace cvpp

// yolov8_head_fused.cu — SURVEY.md §8(f) rank 3: the LAST 1x1 convolutions of the YOLOv8 Detect head fused with the
// decode + confidence filter, so that the (B, 144, A) head tensor is never written to / re-read from HBM (sm_100a).
//
// Replaces (reference file:line): the final layers of Detect.cv2[i] / Detect.cv3[i],
//     nn.Conv2d(c2, 4 * reg_max, 1)   and   nn.Conv2d(c3, nc, 1)          core/models/yolov8/modules.py:423-425,
// their concatenation  x[i] = torch.cat((cv2[i](x[i]), cv3[i](x[i])), 1)  modules.py:431, and everything
// yolov8_decode.cu replaces downstream (DFL :80-82, make_anchors, dist2bbox, the candidate stage of
// non_max_suppression).
//
// This is the one contraction next to the path, so it is the one kernel of the library on the tensor cores:
//   D[cell, n] = sum_k X[k, cell] * W[n, k] + bias[n]        M = 128 cells per tile, N = 64 (box) | nc (class), K = c2 | c3
// as tcgen05.mma.kind::tf32 (fp32 activations and weights are read as TF32, fp32 accumulate in TMEM - the precision
// cuDNN uses for the reference's fp32 convolution on the GPU with torch's default allow_tf32).
//   * A operand = the activations exactly as the previous layer left them, NCHW: cells are contiguous, so a tile is
//     "MN-major"; 3-D tensor-map TMA loads of [32 cells x 16 channels] with the 32-byte-atom 128-byte swizzle
//     (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B = UMMA SWIZZLE_128B_BASE32B, the one layout the tensor core takes for an
//     MN-major 32-bit operand) land in the canonical UMMA layout (4 atoms of 32 cells per 128-cell tile), 8-stage
//     ring of 8 KB chunks, L2 evict-first;
//   * B operand = the conv weights (n, k), K-major, written ONCE per level into shared memory in the swizzled
//     canonical layout by the epilogue warps;
//   * one elected thread issues the MMAs (two K = 8 steps per chunk), tcgen05.commit releases ring stages and
//     publishes the accumulator (two TMEM accumulator stages of 64 + nc columns);
//   * epilogue = two groups of four warps alternating over tiles: thread t owns TMEM lane t = cell t of the tile,
//     pulls its 64 box logits and nc class logits with tcgen05.ld (32x32b.x16), adds the bias and runs EXACTLY the
//     per-cell code of the streaming decode kernel (yolov8_cell.cuh) - DFL integral, class argmax with first-index
//     tie repair, sigmoid, threshold, warp-aggregated candidate append.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = epilogue.  Persistent, one CTA per SM;
// tiles are dealt level by level (the weights of one level are resident at a time).
// HBM-bound like the decode: (c2 + c3) * A * 4 bytes per image (4 838 400 B for the n model) - but the producing
// convolution's 4.8 MB/image write and the decode's re-read of it are gone.
#include <cuda.h>

#include "cvpp_common.cuh"
#include "yolov8_cell.cuh"

namespace cvpp {

constexpr int kHfTileM = 128;        // cells per tile (UMMA M)
constexpr int kHfAtomCells = 32;     // cells per 128-byte swizzle atom row
constexpr int kHfChunkK = 16;        // channels per ring stage (two K = 8 MMA steps)
constexpr int kHfStages = 8;
constexpr int kHfStageBytes = kHfTileM * kHfChunkK * 4;  // 8192
constexpr int kHfAtomBytes = kHfChunkK * 128;            // one 32-cell atom of a stage: 2048
constexpr int kHfBoxN = 4 * kRegMax;                     // 64
constexpr int kHfThreads = 320;                          // 2 + 8 warps
constexpr int kHfAccCols = 256;                          // TMEM columns per accumulator stage (64 + nc_pad <= 256)

struct HeadLevel {
  const float* box_w;   // (64, c2) row-major
  const float* box_b;   // (64)
  const float* cls_w;   // (nc, c3)
  const float* cls_b;   // (nc)
  int hw, w;
  float stride;
  int anchor_off;
  int tiles_per_image;
  int tile_base;        // tiles of the earlier levels (all images): rotates the deal so that every SM stays busy
};

struct HeadParams {
  CUtensorMap tmap_box[CVPP_MAX_LEVELS];  // (cell, channel, image) of the box-branch features, box 32 x 16 x 1, SWIZZLE_128B
  CUtensorMap tmap_cls[CVPP_MAX_LEVELS];
  HeadLevel lv[CVPP_MAX_LEVELS];
  int num_levels, B, c2, c3, nc, nc_pad, A;
  float conf_thres;
  uint64_t* cand_key;
  int32_t* cand_count;
  float4* box_dense;
  int max_cand;
  int w_box_bytes, w_cls_bytes;  // shared-memory footprint of one level's weights
  float* head_out;               // optional (B, 64 + nc, A): the materialised head x_cat (modules.py:438), for callers that want it
};

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void hf_mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a protocol bug traps (the launch fails) instead of hanging the device
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 28)) __trap();
}
__device__ __forceinline__ void hf_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, fp32 accumulate; one thread issues for the CTA
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcgen05_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void hf_tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets in 16-byte
// units, version 1 (Blackwell), layout type 2 = SWIZZLE_128B (16-byte chunks), 1 = SWIZZLE_128B_BASE32B (32-byte chunks:
// the ONLY layout the tensor core accepts for an MN-major 32-bit operand - with type 2 the MMA silently yields zeros,
// tools/probes/umma_probe.cu).
constexpr int kUmmaSw128 = 2, kUmmaSw128Base32 = 1;
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, int layout_type) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 @ bit 4), A = B = TF32 (2 @ bits 7, 10),
// A MN-major (bit 15), B K-major, N >> 3 @ bit 17, M >> 4 @ bit 24.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// weights (n_rows, k) row-major in global -> K-major canonical layout with the 128-byte swizzle:
// [ceil(k / 32) atoms][n_pad rows][128 B], 16-byte chunk index XOR (row % 8); rows >= n_rows and columns >= k are zero.
__device__ __forceinline__ void hf_stage_weights(float* dst, const float* __restrict__ w, int n_rows, int n_pad, int k, int tid, int nthreads) {
  const int atoms = (k + 31) >> 5;
  const int total = atoms * n_pad * 32;
  for (int i = tid; i < total; i += nthreads) {
    const int kk = i & 31, n = (i >> 5) % n_pad, a = (i >> 5) / n_pad;
    const int kg = (a << 5) + kk;
    const float v = (n < n_rows && kg < k) ? __ldg(w + (int64_t)n * k + kg) : 0.0f;
    const int off = (a * n_pad + n) * 32 + ((((kk >> 2) ^ (n & 7)) << 2) | (kk & 3));
    dst[off] = v;
  }
}

__global__ void __launch_bounds__(kHfThreads, 1) yolov8_head_fused_kernel(const __grid_constant__ HeadParams p) {
  extern __shared__ unsigned char smem_dyn[];
  // the 128-byte swizzle pattern is anchored at 1024-byte boundaries: align the carve-up by hand (1 KB of slack is allocated)
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  // layout: [ring kHfStages x 8 KB | box W | cls W | bias (64 + nc_pad) | barriers | tmem slot]
  unsigned char* ring = smem_raw;
  float* w_box = reinterpret_cast<float*>(smem_raw + kHfStages * kHfStageBytes);
  float* w_cls = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(w_box) + p.w_box_bytes);
  float* bias = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(w_cls) + p.w_cls_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias + kHfBoxN + p.nc_pad);
  uint64_t* full = bars;                      // [kHfStages] TMA -> MMA
  uint64_t* empty = bars + kHfStages;         // [kHfStages] MMA -> TMA
  uint64_t* acc_full = bars + 2 * kHfStages;  // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;         // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunk_box = p.c2 / kHfChunkK, nchunk_cls = p.c3 / kHfChunkK;
  const int nchunks = nchunk_box + nchunk_cls;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kHfStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);  // one arrive per epilogue warp of the group
    }
    mbar_fence_init();
  }
  if (warp == 1) {  // 512 TMEM columns: two accumulator stages of 256
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // running pipeline state (identical sequences in every role)
  uint32_t q = 0;     // ring chunk counter
  uint32_t tile_i = 0;  // CTA-local tile counter (accumulator stage = tile_i & 1)
  const uint64_t policy = l2_policy_evict_first();
  const uint32_t idesc_box = umma_idesc_tf32(kHfTileM, kHfBoxN);
  const uint32_t idesc_cls = umma_idesc_tf32(kHfTileM, p.nc_pad);

  for (int l = 0; l < p.num_levels; ++l) {
    const HeadLevel& L = p.lv[l];
    // (A) every MMA of the previous level has completed (the epilogue waited for each accumulator)
    __syncthreads();
    if (warp >= 2) {
      const int t = tid - 64;
      hf_stage_weights(w_box, L.box_w, kHfBoxN, kHfBoxN, p.c2, t, kHfThreads - 64);
      hf_stage_weights(w_cls, L.cls_w, p.nc, p.nc_pad, p.c3, t, kHfThreads - 64);
      for (int i = t; i < kHfBoxN + p.nc_pad; i += kHfThreads - 64)
        bias[i] = i < kHfBoxN ? __ldg(L.box_b + i) : (i - kHfBoxN < p.nc ? __ldg(L.cls_b + i - kHfBoxN) : 0.0f);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA's async proxy
    }
    __syncthreads();  // (B)

    const int n_tiles = L.tiles_per_image * p.B;
    // deal: tile t of the level goes to CTA (t + tile_base) % grid
    int first = (int)blockIdx.x - (L.tile_base % (int)gridDim.x);
    if (first < 0) first += gridDim.x;

    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        for (int t = first; t < n_tiles; t += gridDim.x) {
          const int b = t / L.tiles_per_image, cell0 = (t - b * L.tiles_per_image) * kHfTileM;
          for (int c = 0; c < nchunks; ++c, ++q) {
            const uint32_t s = q % kHfStages;
            hf_mbar_wait(&empty[s], ((q / kHfStages) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&full[s], kHfStageBytes);
            const CUtensorMap* tm = c < nchunk_box ? &p.tmap_box[l] : &p.tmap_cls[l];
            const int ch = (c < nchunk_box ? c : c - nchunk_box) * kHfChunkK;
            unsigned char* dst = ring + s * kHfStageBytes;
#pragma unroll
            for (int a = 0; a < kHfTileM / kHfAtomCells; ++a)
              hf_tma_load_3d(dst + a * kHfAtomBytes, tm, cell0 + a * kHfAtomCells, ch, b, &full[s], policy);
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        const uint32_t wb = smem_u32(w_box), wc = smem_u32(w_cls);
        for (int t = first; t < n_tiles; t += gridDim.x, ++tile_i) {
          const uint32_t acc = tile_i & 1u;
          hf_mbar_wait(&acc_empty[acc], ((tile_i >> 1) & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t d_box = tmem_base + acc * kHfAccCols, d_cls = d_box + kHfBoxN;
          for (int c = 0; c < nchunks; ++c, ++q) {
            const uint32_t s = q % kHfStages;
            hf_mbar_wait(&full[s], (q / kHfStages) & 1u);
            tcgen05_fence_after();
            const uint32_t a_base = smem_u32(ring + s * kHfStageBytes);
            const bool is_box = c < nchunk_box;
            const int cc = is_box ? c : c - nchunk_box;
#pragma unroll
            for (int ks = 0; ks < kHfChunkK / 8; ++ks) {
              const int kk = cc * (kHfChunkK / 8) + ks;  // K = 8 step index within the branch
              // A: MN-major, 32-byte-chunk swizzle: atoms of 32 cells [16 rows x 128 B] 2048 B apart (LBO), the 4-row K
              // groups of the swizzle atom 512 B apart (SBO); a K = 8 step is two of them = 1024 B
              const uint64_t a_desc = umma_desc(a_base + ks * 1024, kHfAtomBytes, 512, kUmmaSw128Base32);
              // B: K-major, [atom][row][128 B]: 8-row groups 1024 B apart (SBO), K step = 32 B inside the 128-byte row
              const uint32_t n_pad = is_box ? kHfBoxN : p.nc_pad;
              const uint32_t b_addr = (is_box ? wb : wc) + (kk >> 2) * n_pad * 128 + (kk & 3) * 32;
              const uint64_t b_desc = umma_desc(b_addr, 16, 1024, kUmmaSw128);
              tcgen05_mma_tf32(is_box ? d_box : d_cls, a_desc, b_desc, is_box ? idesc_box : idesc_cls, kk > 0 ? 1u : 0u);
            }
            tcgen05_commit(&empty[s]);  // the stage is free once these MMAs have read it
          }
          tcgen05_commit(&acc_full[acc]);  // accumulator complete
        }
      }
      __syncwarp();
    } else {
      // ===================== epilogue: 2 groups x 4 warps, group g takes the tiles with tile_i % 2 == g ==========
      const int group = (warp - 2) >> 2;
      const int quarter = warp & 3;  // TMEM lanes [32 * quarter, +32) are the ones this warp may read
      for (int t = first; t < n_tiles; t += gridDim.x, ++tile_i) {
        if ((int)(tile_i & 1u) != group) continue;
        const uint32_t acc = tile_i & 1u;
        const int b = t / L.tiles_per_image, cell0 = (t - b * L.tiles_per_image) * kHfTileM;
        hf_mbar_wait(&acc_full[acc], (tile_i >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t trow = tmem_base + acc * kHfAccCols + ((uint32_t)(quarter * 32) << 16);
        const int cell = cell0 + quarter * 32 + lane;
        float v[16];
        float d[4];
#pragma unroll
        for (int side = 0; side < 4; ++side) {
          tcgen05_ld16(trow + side * 16, v);
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = fadd(v[k], bias[side * 16 + k]);
          if (p.head_out && cell < L.hw) {
#pragma unroll
            for (int k = 0; k < 16; ++k)
              p.head_out[((int64_t)b * (kHfBoxN + p.nc) + side * 16 + k) * p.A + L.anchor_off + cell] = v[k];
          }
          d[side] = dfl16(v);
        }
        float best = -INFINITY, prev = -INFINITY;
        int arg = 0;
        for (int c0 = 0; c0 < p.nc; c0 += 16) {
          tcgen05_ld16(trow + kHfBoxN + c0, v);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (c0 + k < p.nc) {
              const float x = fadd(v[k], bias[kHfBoxN + c0 + k]);
              if (p.head_out && cell < L.hw) p.head_out[((int64_t)b * (kHfBoxN + p.nc) + kHfBoxN + c0 + k) * p.A + L.anchor_off + cell] = x;
              class_step(x, c0 + k, best, arg, prev);
            }
          }
        }
        const bool active = cell < L.hw;
        float score = sigmoid_precise(best);
        bool cand = active && score > p.conf_thres;
        // an EARLIER class whose sigmoid rounds to the same float takes the reference's first-index argmax: redo the
        // scan on sigmoid values (warp-collective TMEM loads, so the whole warp joins when any lane needs it)
        if (__any_sync(0xffffffffu, cand && sigmoid_precise(prev) >= score)) {
          float bs = -1.0f;
          int ba = 0;
          for (int c0 = 0; c0 < p.nc; c0 += 16) {
            tcgen05_ld16(trow + kHfBoxN + c0, v);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              if (c0 + k < p.nc) {
                const float sv = sigmoid_precise(fadd(v[k], bias[kHfBoxN + c0 + k]));
                if (sv > bs) {
                  bs = sv;
                  ba = c0 + k;
                }
              }
            }
          }
          if (cand && sigmoid_precise(prev) >= score) {
            score = bs;
            arg = ba;
          }
        }
        // the accumulator stage can be overwritten: every lane of this warp has its values in registers
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) hf_mbar_arrive(&acc_empty[acc]);

        const CellBox bx = cell_box(active ? cell : 0, L.w, L.stride, d[0], d[1], d[2], d[3]);
        const unsigned m = __ballot_sync(0xffffffffu, cand);
        if (m) {
          int base = 0;
          if (lane == 0) base = atomicAdd(p.cand_count + b, __popc(m));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (cand) {
            const int slot = base + __popc(m & ((1u << lane) - 1u));
            const int anchor = L.anchor_off + cell;
            if (slot < p.max_cand)
              p.cand_key[(int64_t)b * p.max_cand + slot] = key_pack((uint32_t)arg, __float_as_uint(score), (uint32_t)anchor);
            p.box_dense[(int64_t)b * p.A + anchor] = make_float4(bx.x1, bx.y1, bx.x2, bx.y2);
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*HfEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static HfEncodeTiledFn hf_encode_fn() {
  static HfEncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<HfEncodeTiledFn>(f);
  }();
  return fn;
}

// (cell, channel, image) view of a contiguous (B, C, H, W) feature map; box = 32 cells x 16 channels x 1 image with the
// 32-byte-atom 128-byte swizzle (the UMMA canonical MN-major atom for 32-bit operands), zero fill past the last cell
static bool hf_make_tmap(CUtensorMap* tm, const float* ptr, int hw, int C, int B) {
  HfEncodeTiledFn enc = hf_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)hw, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)hw * 4u, (cuuint64_t)hw * 4u * (cuuint64_t)C};
  cuuint32_t box[3] = {(cuuint32_t)kHfAtomCells, (cuuint32_t)kHfChunkK, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int yolov8_head_fused_launch(const float* const* box_feat, const float* const* cls_feat, const float* const* box_w,
                             const float* const* box_b, const float* const* cls_w, const float* const* cls_b,
                             const int* level_h, const int* level_w, const float* level_stride, int num_levels, int B, int c2,
                             int c3, int nc, int reg_max, float conf_thres, uint64_t* cand_key, int32_t* cand_count,
                             float* box_dense, int max_cand, float* head_out, cudaStream_t stream) {
  if (!box_feat || !cls_feat || !box_w || !box_b || !cls_w || !cls_b || !level_h || !level_w || !level_stride) {
    set_error("yolov8 head: NULL level description");
    return CVPP_ERR_INVALID_ARG;
  }
  if (!cand_key || !cand_count || !box_dense || max_cand < 1) {
    set_error("yolov8 head: NULL output / max_cand < 1");
    return CVPP_ERR_INVALID_ARG;
  }
  if (num_levels < 1 || num_levels > CVPP_MAX_LEVELS || B < 0 || nc < 1 || nc > 192) {
    set_error("yolov8 head: bad num_levels=%d / B=%d / nc=%d (nc <= 192: two 64 + nc accumulators share the 512 TMEM columns)",
              num_levels, B, nc);
    return CVPP_ERR_INVALID_ARG;
  }
  if (reg_max != kRegMax) {
    set_error("yolov8 head: reg_max=%d is not compiled in (reference hard-codes 16, modules.py:413)", reg_max);
    return CVPP_ERR_UNSUPPORTED;
  }
  if (c2 < kHfChunkK || c3 < kHfChunkK || (c2 % kHfChunkK) || (c3 % kHfChunkK) || c2 > 1024 || c3 > 1024) {
    set_error("yolov8 head: c2=%d / c3=%d must be multiples of %d (the reference's widths are 64 / 80..320)", c2, c3, kHfChunkK);
    return CVPP_ERR_UNSUPPORTED;
  }
  if (!(conf_thres >= 0.0f && conf_thres <= 1.0f)) {
    set_error("Invalid Confidence threshold %f, valid values are between 0.0 and 1.0", conf_thres);
    return CVPP_ERR_INVALID_ARG;
  }
  if (reinterpret_cast<uintptr_t>(box_dense) & 15u) {
    set_error("yolov8 head: box_dense must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != CVPP_OK) return rc;
  alignas(64) HeadParams p{};
  p.num_levels = num_levels;
  p.B = B;
  p.c2 = c2;
  p.c3 = c3;
  p.nc = nc;
  p.nc_pad = (nc + 15) & ~15;
  p.conf_thres = conf_thres;
  p.cand_key = cand_key;
  p.cand_count = cand_count;
  p.box_dense = reinterpret_cast<float4*>(box_dense);
  p.max_cand = max_cand;
  p.head_out = head_out;
  p.w_box_bytes = ((c2 + 31) / 32) * kHfBoxN * 128;
  p.w_cls_bytes = ((c3 + 31) / 32) * p.nc_pad * 128;
  int64_t A = 0;
  int tiles = 0;
  for (int l = 0; l < num_levels; ++l) {
    if (!box_feat[l] || !cls_feat[l] || !box_w[l] || !box_b[l] || !cls_w[l] || !cls_b[l] || level_h[l] < 1 || level_w[l] < 1) {
      set_error("yolov8 head: level %d is empty", l);
      return CVPP_ERR_INVALID_ARG;
    }
    HeadLevel& L = p.lv[l];
    L.box_w = box_w[l];
    L.box_b = box_b[l];
    L.cls_w = cls_w[l];
    L.cls_b = cls_b[l];
    L.hw = level_h[l] * level_w[l];
    L.w = level_w[l];
    L.stride = level_stride[l];
    L.anchor_off = (int)A;
    L.tiles_per_image = (L.hw + kHfTileM - 1) / kHfTileM;
    L.tile_base = tiles;
    tiles += L.tiles_per_image * B;
    A += L.hw;
    if ((reinterpret_cast<uintptr_t>(box_feat[l]) & 15u) || (reinterpret_cast<uintptr_t>(cls_feat[l]) & 15u) || (L.hw & 3)) {
      set_error("yolov8 head: level %d feature maps must be 16-byte aligned with H*W a multiple of 4 (TMA rows)", l);
      return CVPP_ERR_ALIGNMENT;
    }
  }
  if (A > CVPP_MAX_ANCHORS) {
    set_error("yolov8 head: %lld anchors exceed the %d-anchor key field", (long long)A, CVPP_MAX_ANCHORS);
    return CVPP_ERR_UNSUPPORTED;
  }
  p.A = (int)A;
  CVPP_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(int32_t) * (size_t)B, stream));
  if (B == 0) return CVPP_OK;
  for (int l = 0; l < num_levels; ++l) {
    if (!hf_make_tmap(&p.tmap_box[l], box_feat[l], p.lv[l].hw, c2, B) || !hf_make_tmap(&p.tmap_cls[l], cls_feat[l], p.lv[l].hw, c3, B)) {
      set_error("yolov8 head: cuTensorMapEncodeTiled failed for level %d", l);
      return CVPP_ERR_CUDA;
    }
  }
  const size_t smem = (size_t)kHfStages * kHfStageBytes + p.w_box_bytes + p.w_cls_bytes + (size_t)(kHfBoxN + p.nc_pad) * 4 +
                      (2 * kHfStages + 4) * sizeof(uint64_t) + 16 + 1024;
  if (smem > (size_t)di.max_smem) {
    set_error("yolov8 head: c2=%d c3=%d nc=%d need %zu bytes of shared memory, the device has %d", c2, c3, nc, smem, di.max_smem);
    return CVPP_ERR_UNSUPPORTED;
  }
  static unsigned long long attr_done = 0;
  rc = ensure_smem_attr(reinterpret_cast<const void*>(yolov8_head_fused_kernel), di.max_smem, di.device, &attr_done);
  if (rc != CVPP_OK) return rc;
  const int grid = tiles < di.sms ? tiles : di.sms;
  yolov8_head_fused_kernel<<<grid, kHfThreads, smem, stream>>>(p);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

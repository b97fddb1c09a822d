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
//     MN-major 32-bit operand) land in the canonical UMMA layout (4 atoms of 32 cells per 128-cell tile), a ring of
//     16 KB stages (32 channels) as deep as shared memory allows, L2 evict-first;
//   * B operand = the conv weights (n, k), K-major: 2-D tensor-map TMA loads of [32 k x N rows] with the plain 128-byte
//     swizzle are the canonical layout as they land; two weight buffers alternate by level, so there is no CTA-wide
//     barrier after the set-up (the first version staged weights by hand behind __syncthreads: 10 us of the kernel);
//   * two MMA warps (box branch N = 64, class branch N = nc_pad), each running its loop warp-uniformly with one elected
//     lane issuing (four K = 8 steps per stage); tcgen05.commit releases ring stages and publishes the branch's
//     accumulator columns (two TMEM accumulator stages of 64 + nc columns); each loads its own weights by TMA;
//   * epilogue = two groups of eight warps alternating over tiles; in a group, TMEM lane t = cell t of the tile is
//     served by TWO threads: one of a "box warp" pulls the 64 box logits with tcgen05.ld (32x32b.x16), adds the bias and
//     runs the DFL integral + anchor arithmetic (it starts when the box branch is done, under the class MMAs), one of
//     a "class warp" pulls the nc class logits and runs the argmax with first-index tie repair, sigmoid, threshold and
//     the warp-aggregated candidate append (the box reaches it through shared memory + a 64-thread named barrier) -
//     EXACTLY the per-cell code of the streaming decode kernel (yolov8_cell.cuh).
// Warp roles: 0 / 3 = TMA producers of the box / class branch (activations only: one 4-D UTMALDG per 16 KB stage into the
// branch's own ring), 1 / 2 = MMA issuers of the box / class branch (1 also owns the TMEM allocation), 4..19 = epilogue.  Persistent, one CTA per SM; tiles are dealt level by level.
// Measured steps (C2 shape, 64 images, us): one thread issuing under `if (lane == 0)` 80 (every UTMALDG / UTCHMMA wrapped in
// an elect + R2UR waterfall) -> warp-uniform loops with elect.sync 66.7 -> weights off the producer's path 65.5 -> two MMA
// warps, one ring per branch 58.2 -> one producer per ring + the class scan shared with the box warps 56.8 (0.85 of the
// measured HBM peak); the same rings with the MMAs and the epilogue switched off (CVPP_HEAD_DEBUG=3) stream in 49.2, with
// the MMAs on and the epilogue off (=2) in 50.8: what is left is issue-slot contention between the sixteen epilogue warps
// and the four single-lane roles that share their schedulers.
// HBM-bound like the decode: (c2 + c3) * A * 4 bytes per image (4 838 400 B for the n model) - but the producing
// convolution's 4.8 MB/image write and the decode's re-read of it are gone.
#include <cuda.h>
#include <cstdlib>

#include "cvpp_common.cuh"
#include "yolov8_cell.cuh"

namespace cvpp {

constexpr int kHfTileM = 128;        // cells per tile (UMMA M)
constexpr int kHfAtomCells = 32;     // cells per 128-byte swizzle atom row
constexpr int kHfChunkK = 32;        // channels per ring stage (up to four K = 8 MMA steps) = one 128-byte K atom of the weights
constexpr int kHfMaxStages = 12;                        // ring stages actually used: as many as shared memory holds (host)
constexpr int kHfStageBytes = kHfTileM * kHfChunkK * 4;  // 16384
constexpr int kHfAtomBytes = kHfChunkK * 128;            // one 32-cell atom of a stage: 4096
constexpr int kHfBoxN = 4 * kRegMax;                     // 64
constexpr int kHfThreads = 640;                          // 4 + 16 warps (the epilogue warps keep warp % 4 = TMEM lane quarter)
constexpr int kHfAccCols = 256;                          // TMEM columns per accumulator stage (64 + nc_pad <= 256)

struct HfSlot {   // what a box warp hands to the class warp of the same cells
  float4 box;     // decoded xyxy
  float4 part;    // its share of the class scan: (best logit, class index as bits, largest logit before it, -)
};

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
  int one_tma;          // H*W is a multiple of 32: the feature maps have the 4-D (cell in atom, channel, atom, image) tensor map
};

struct HeadParams {
  CUtensorMap tmap_box[CVPP_MAX_LEVELS];  // (cell, channel, image) of the box-branch features, box 32 x 16 x 1, SWIZZLE_128B
  CUtensorMap tmap_cls[CVPP_MAX_LEVELS];
  CUtensorMap tmap_wbox[CVPP_MAX_LEVELS];  // (k, n) conv weights, box 32 k x 64 rows, SWIZZLE_128B: lands in the K-major UMMA layout
  CUtensorMap tmap_wcls[CVPP_MAX_LEVELS];  // box 32 k x nc_pad rows (rows >= nc and k >= c3 zero-filled)
  HeadLevel lv[CVPP_MAX_LEVELS];
  int num_levels, B, c2, c3, nc, nc_pad, A;
  float conf_thres;
  uint64_t* cand_key;
  int32_t* cand_count;
  float4* box_dense;
  int max_cand;
  int w_box_bytes, w_cls_bytes;  // shared-memory footprint of one level's weights
  int stages_box;                // of which the box branch's ring (the rest is the class branch's)
  int stages;                    // ring depth (8 KB each): the bytes in flight per SM that keep HBM busy
  int w_level_bytes;             // footprint of one level's weights: box W + cls W
  int cls_split;                 // classes [cls_split, nc) are scanned by the box warps (multiple of 16, or nc = none)
  int debug;                     // CVPP_HEAD_DEBUG (measurement only): 1 = no MMAs (pure TMA stream), 2 = epilogue releases at once, 4 = no candidate append
  int w_bufs;                    // weight buffers: 2 (alternating by level) when shared memory allows, else 1 (the wide models)
  float* head_out;               // optional (B, 64 + nc, A): the materialised head x_cat (modules.py:438), for callers that want it
};

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
#ifdef CVPP_NMS_TIMING   // the instrumented build (make timing): a timeline of CTA 0 (clock64 stamps, fire-and-forget stores)
__device__ long long g_hf_tl[8][512];   // [0] producer: chunk issued  [1] mma: chunk's operands landed  [2] mma: chunk committed
                                        // [3] box warp (quarter 0): accumulator seen full  [4] ... released  [5] kernel start / end
__device__ unsigned long long g_hf_cta[256][4];   // per CTA: globaltimer at start / first operands landed / last tile's accumulator / end
__device__ __forceinline__ unsigned long long hf_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define HF_STAMP(row, idx) do { if (blockIdx.x == 0 && (idx) < 512) g_hf_tl[row][idx] = clock64(); } while (0)
#define HF_CTA(k) do { if (blockIdx.x < 256) g_hf_cta[blockIdx.x][k] = hf_gtime(); } while (0)
#else
#define HF_CTA(k) do {} while (0)
#define HF_STAMP(row, idx) do {} while (0)
#endif
#define HF_WAIT(slot, stmt) do { stmt; } while (0)
#ifndef HF_SLEEP_NS
#define HF_SLEEP_NS 160   // (40 / 160 / 400 ns measured: 56.7 / 56.2 / 56.1 us)
#endif
__device__ __forceinline__ void hf_mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a protocol bug traps (the launch fails) instead of hanging the device
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 28)) __trap();
}
// the epilogue warps wait for most of a tile period: poll with a back-off.  (A third of the instructions this kernel executes
// are these polls - try_wait returns after a short hardware limit whatever suspend-time hint it is given, measured - but
// they are not what bounds it: the same polls run with the epilogue arithmetic switched off, 50.8 us.)
__device__ __forceinline__ void hf_mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    __nanosleep(HF_SLEEP_NS);
    if (spins > (1u << 26)) __trap();
  }
}
// one lane of a CONVERGED warp: the warp runs the role's loop uniformly (descriptors and addresses live in uniform registers)
// and only the issue is predicated; `if (lane == 0)` around the whole loop makes the compiler wrap every UTMALDG / UTCHMMA in
// an elect + R2UR waterfall (~25 instructions, ~100 cycles per MMA on a scheduler shared with four epilogue warps)
__device__ __forceinline__ bool hf_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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

__device__ __forceinline__ void hf_tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar,
                                               uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
      : "memory");
}

__device__ __forceinline__ void hf_tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
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

template <bool HEAD_OUT>
__global__ void __launch_bounds__(kHfThreads, 1) yolov8_head_fused_kernel(const __grid_constant__ HeadParams p) {
  extern __shared__ unsigned char smem_dyn[];
  // the 128-byte swizzle pattern is anchored at 1024-byte boundaries: align the carve-up by hand (1 KB of slack is allocated)
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  // layout: [ring p.stages x 16 KB | 2 x (box W | cls W) | barriers | tmem slot | box hand-over slots]
  unsigned char* ring = smem_raw;
  const uint32_t n_stages = (uint32_t)p.stages;
  unsigned char* w_region = smem_raw + n_stages * kHfStageBytes;  // weight buffer of level l: w_region + (l % w_bufs) * w_level_bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_region + (size_t)p.w_bufs * p.w_level_bytes);
  uint64_t* full = bars;                         // [kHfMaxStages] TMA -> MMA
  uint64_t* empty = bars + kHfMaxStages;         // [kHfMaxStages] MMA -> TMA
  // per branch (0 = box, 1 = class) x 2 stages / buffers:
  uint64_t* acc_full = bars + 2 * kHfMaxStages;  // [2][2] MMA -> epilogue: the branch's accumulator columns are complete
  uint64_t* acc_empty = acc_full + 4;            // [2][2] epilogue -> MMA
  uint64_t* w_full = acc_empty + 4;              // [2][2] TMA -> MMA: the level's weights of the branch have landed
  uint64_t* w_empty = w_full + 4;                // [2][2] MMA -> TMA: every MMA that reads the buffer has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + 4);
  HfSlot* box_sh = reinterpret_cast<HfSlot*>(reinterpret_cast<unsigned char*>(tmem_slot) + 16);  // [2 groups][2 slots][128 cells]

  if (threadIdx.x == 0) { HF_STAMP(5, 0); HF_CTA(0); }
  uint32_t dbg_chunk = 0;  // chunk counter of the instrumented build
  (void)dbg_chunk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunk_box = (p.c2 + kHfChunkK - 1) / kHfChunkK, nchunk_cls = (p.c3 + kHfChunkK - 1) / kHfChunkK;
  const int nchunks = nchunk_box + nchunk_cls;

  if ((warp == 0 || warp == 3) && lane == 1) {
    for (int l = 0; l < p.num_levels; ++l)
      asm volatile("prefetch.tensormap [%0];" ::"l"(warp == 0 ? &p.tmap_box[l] : &p.tmap_cls[l]) : "memory");
  }
  if (warp == 1 && lane == 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmap_wbox[0]) : "memory");
  if (warp == 2 && lane == 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmap_wcls[0]) : "memory");
  if (warp == 0 && lane == 0) {
    for (uint32_t s = 0; s < n_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 4; ++a) {
      mbar_init(&acc_full[a], 1);
      // one arrive per epilogue warp of the group that reads the branch's columns (the box warps also read class columns)
      mbar_init(&acc_empty[a], (a >= 2 && p.cls_split < p.nc) ? 8 : 4);
      mbar_init(&w_full[a], 1);
      mbar_init(&w_empty[a], 1);
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
  // Two rings, one per branch: box stages [0, n_box), class stages [n_box, n_stages).  Each MMA warp then sees EVERY phase of
  // the barriers it waits on, in order (with one shared ring a warp skips the other branch's fills, and a parity wait
  // cannot tell a stage that is one fill behind from one that is one fill ahead: measured as rare launch failures).
  const uint32_t n_box = (uint32_t)p.stages_box, n_cls = n_stages - n_box;
  uint32_t stage = 0, phase = 0;    // producers and MMA warps: position in their branch's ring
  uint32_t tile_i = 0;  // CTA-local tile counter (accumulator stage = tile_i & 1)
  const uint64_t policy = l2_policy_evict_first();
  const uint32_t idesc_box = umma_idesc_tf32(kHfTileM, kHfBoxN);
  const uint32_t idesc_cls = umma_idesc_tf32(kHfTileM, p.nc_pad);

  for (int l = 0; l < p.num_levels; ++l) {
    const HeadLevel& L = p.lv[l];
    // The conv weights of level l live in buffer l % w_bufs, loaded by the producer's TMA (no CTA-wide barrier anywhere
    // after the set-up: the roles only meet on mbarriers).
    const int wbuf_i = p.w_bufs == 2 ? (l & 1) : 0;
    const uint32_t wbuf_par = (uint32_t)(p.w_bufs == 2 ? (l >> 1) : l) & 1u;
    unsigned char* w_buf = w_region + (size_t)wbuf_i * p.w_level_bytes;
    const float* __restrict__ bias_box = L.box_b;   // biases straight from global memory (L1-resident broadcast loads)
    const float* __restrict__ bias_cls = L.cls_b;

    const int n_tiles = L.tiles_per_image * p.B;
    // deal: tile t of the level goes to CTA (t + tile_base) % grid
    int first = (int)blockIdx.x - (L.tile_base % (int)gridDim.x);
    if (first < 0) first += gridDim.x;

    if (warp == 0 || warp == 3) {
      // ===================== TMA producers: warp 0 streams the box branch's activations into the box ring, warp 3 the
      // class branch's into the class ring.  (One in-order producer over both rings blocked on whichever ring was full
      // while the other had free stages: 3400 cycles per tile instead of the 2800 the memory system delivers.)
      // (ring position = (stage, parity) counters carried across tiles and levels: no division on the issue path)
      {
        const bool is_box = warp == 0;
        const CUtensorMap* tm = is_box ? &p.tmap_box[l] : &p.tmap_cls[l];
        const int my_chunks = is_box ? nchunk_box : nchunk_cls;
        const uint32_t ring0 = is_box ? 0u : n_box, my_stages = is_box ? n_box : n_cls;
        for (int t = first; t < n_tiles; t += gridDim.x) {
          const int b = t / L.tiles_per_image, cell0 = (t - b * L.tiles_per_image) * kHfTileM;
          for (int c = 0; c < my_chunks; ++c) {
            const uint32_t st = ring0 + stage;
            hf_mbar_wait(&empty[st], phase ^ 1u);
            const int ch = c * kHfChunkK;
            unsigned char* dst = ring + st * kHfStageBytes;
            if (hf_elect_one()) {
              mbar_arrive_expect_tx(&full[st], kHfStageBytes);   // channels past the last one are zero-filled and still counted
              if (L.one_tma) {
                // one instruction per stage: (cell in atom, channel, atom, image) view, box 32 x 32 x 4 x 1
                hf_tma_load_4d(dst, tm, 0, ch, cell0 / kHfAtomCells, b, &full[st], policy);
              } else {
#pragma unroll
                for (int a = 0; a < kHfTileM / kHfAtomCells; ++a)
                  hf_tma_load_3d(dst + a * kHfAtomBytes, tm, cell0 + a * kHfAtomCells, ch, b, &full[st], policy);
              }
            }
            __syncwarp();
            if (lane == 0) HF_STAMP(0, dbg_chunk + (is_box ? 0 : nchunk_box) + c);
            if (++stage == my_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          dbg_chunk += nchunks;
        }
      }
    } else if (warp == 1 || warp == 2) {
      // ===================== MMA issuers: warp 1 = box branch (N = 64), warp 2 = class branch (N = nc_pad) ===========
      // Two issuing warps on two schedulers: one warp's serial loop (wait, fence, elect, 4 UTCHMMA, UTCBAR: ~100
      // instructions per stage, sharing its scheduler with four epilogue warps) took 750 cycles per stage - more than
      // the 530 the memory stream needs.  Each branch has its own accumulator columns, weights and barriers, so the box
      // epilogue also starts while the class MMAs of the same tile are still running.
      //   A (MN-major, 32-byte-chunk swizzle): atoms of 32 cells [32 rows x 128 B] 4096 B apart (LBO), the 4-row K groups of
      //     the swizzle atom 512 B apart (SBO); a K = 8 step = 1024 B further into every atom;
      //   B (K-major, 128-byte swizzle): [K atom = chunk][row][128 B], 8-row groups 1024 B apart (SBO), K step = 32 B.
      {
        const bool is_box = warp == 1;
        const int role = is_box ? 0 : 1;
        const int my_chunks = is_box ? nchunk_box : nchunk_cls;
        const int skip_before = is_box ? 0 : nchunk_box;   // (chunk numbering of the instrumented build)
        (void)skip_before;
        const uint32_t ring0 = is_box ? 0u : n_box, my_stages = is_box ? n_box : n_cls;
        const int ksteps = (is_box ? p.c2 : p.c3) >> 3;
        const int n_rows = is_box ? kHfBoxN : p.nc_pad;
        const uint32_t my_w_off = is_box ? 0u : (uint32_t)p.w_box_bytes, my_w_bytes = (uint32_t)(is_box ? p.w_box_bytes : p.w_cls_bytes);
        const uint32_t idesc = is_box ? idesc_box : idesc_cls;
        const uint32_t atom16 = (uint32_t)(n_rows * 128) >> 4;
        const uint64_t a_desc0 = umma_desc(smem_u32(ring), kHfAtomBytes, 512, kUmmaSw128Base32);
        const uint64_t b_desc0 = umma_desc(smem_u32(w_buf + my_w_off), 16, 1024, kUmmaSw128);
        uint64_t* my_w_full = w_full + 2 * role;
        uint64_t* my_w_empty = w_empty + 2 * role;
        uint64_t* my_acc_full = acc_full + 2 * role;
        uint64_t* my_acc_empty = acc_empty + 2 * role;
        // The conv weights are this warp's own TMA loads (the producer only streams activations): level 0 (and every level
        // when there is one buffer) on entry; with two buffers, level l + 1 once the first tile of level l is in flight,
        // after the MMAs of level l - 1 - the last readers of that buffer - have completed (w_empty, committed below).
        auto load_weights = [&](int wl) {
          const int buf = p.w_bufs == 2 ? (wl & 1) : 0;
          const uint32_t use = (uint32_t)(p.w_bufs == 2 ? (wl >> 1) : wl);   // how many times the buffer was filled before
          if (use) hf_mbar_wait(&my_w_empty[buf], (use & 1u) ^ 1u);
          unsigned char* wb = w_region + (size_t)buf * p.w_level_bytes + my_w_off;
          if (hf_elect_one()) {
            mbar_arrive_expect_tx(&my_w_full[buf], my_w_bytes);
            const CUtensorMap* tm = is_box ? &p.tmap_wbox[wl] : &p.tmap_wcls[wl];
            for (int a = 0; a < my_chunks; ++a) hf_tma_load_2d(wb + a * (n_rows * 128), tm, a * 32, 0, &my_w_full[buf]);
          }
          __syncwarp();
        };
        if (l == 0 || p.w_bufs == 1) load_weights(l);
        bool next_pending = p.w_bufs == 2 && l + 1 < p.num_levels;
        hf_mbar_wait(&my_w_full[wbuf_i], wbuf_par);   // this level's weights have landed
        tcgen05_fence_after();
        for (int t = first; t < n_tiles; t += gridDim.x, ++tile_i) {
          const uint32_t acc = tile_i & 1u;
          hf_mbar_wait(&my_acc_empty[acc], ((tile_i >> 1) & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t d = tmem_base + acc * kHfAccCols + (is_box ? 0u : (uint32_t)kHfBoxN);
          for (int cc = 0; cc < my_chunks; ++cc) {
            const int nks = min(4, ksteps - 4 * cc);   // K = 8 steps with real channels
            const uint32_t st = ring0 + stage;
            const uint64_t a_desc = a_desc0 + (uint64_t)(st * (kHfStageBytes >> 4));
            const uint64_t b_desc = b_desc0 + (uint64_t)(cc * atom16);
            hf_mbar_wait(&full[st], phase);
            if (lane == 0) HF_STAMP(1, dbg_chunk + skip_before + cc);
            if (dbg_chunk == 0 && cc == 0 && is_box && lane == 0) HF_CTA(1);
            tcgen05_fence_after();
            if (hf_elect_one()) {
              if (p.debug & 1) {
                hf_mbar_arrive(&empty[st]);
              } else {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (ks < nks) tcgen05_mma_tf32(d, a_desc + (uint64_t)(ks * 64), b_desc + (uint64_t)(ks * 2), idesc, (cc | ks) ? 1u : 0u);
                tcgen05_commit(&empty[st]);  // the stage is free once these MMAs have read it
              }
            }
            __syncwarp();
            if (lane == 0) HF_STAMP(2, dbg_chunk + skip_before + cc);
            if (++stage == my_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          dbg_chunk += nchunks;
          if (hf_elect_one()) tcgen05_commit(&my_acc_full[acc]);  // this branch's accumulator columns are complete
          __syncwarp();
          if (lane == 0 && !is_box) HF_CTA(2);
          if (next_pending) {
            load_weights(l + 1);
            next_pending = false;
          }
        }
        if (next_pending) load_weights(l + 1);
        if (hf_elect_one()) tcgen05_commit(&my_w_empty[wbuf_i]);     // every MMA reading this level's weights is done once this arrives
        __syncwarp();
      }
      __syncwarp();
    } else if (warp >= 4) {
      // ===================== epilogue: 2 groups x (4 box warps + 4 class warps); group g takes the tiles with
      //                       tile_i % 2 == g; the box warp and the class warp of a TMEM lane quarter meet on a named barrier
      const int e = warp - 4;
      const int group = e >> 3;
      const bool cls_role = (e >> 2) & 1;
      const int quarter = warp & 3;  // TMEM lanes [32 * quarter, +32) are the ones this warp may read
      const int pair_bar = 1 + group * 4 + quarter;
      for (int t = first; t < n_tiles; t += gridDim.x, ++tile_i) {
        if ((int)(tile_i & 1u) != group) continue;
        const uint32_t acc = tile_i & 1u;
        const int b = t / L.tiles_per_image, cell0 = (t - b * L.tiles_per_image) * kHfTileM;
        HfSlot* slot = box_sh + ((group * 2 + ((tile_i >> 1) & 1u)) * kHfTileM + quarter * 32 + lane);
        uint64_t* my_acc_empty = acc_empty + 2 * (int)cls_role + acc;
        hf_mbar_wait_relaxed(&acc_full[2 * (int)cls_role + acc], (tile_i >> 1) & 1u);
        if (!cls_role && quarter == 0 && lane == 0) HF_STAMP(3, tile_i);
        tcgen05_fence_after();
        if (p.debug & 2) {
          __syncwarp();
          if (lane == 0) hf_mbar_arrive(my_acc_empty);
          if (!cls_role && p.cls_split < p.nc) {
            hf_mbar_wait_relaxed(&acc_full[2 + acc], (tile_i >> 1) & 1u);
            __syncwarp();
            if (lane == 0) hf_mbar_arrive(&acc_empty[2 + acc]);
          }
          continue;
        }
        const uint32_t trow = tmem_base + acc * kHfAccCols + ((uint32_t)(quarter * 32) << 16);
        const int cell = cell0 + quarter * 32 + lane;
        const bool active = cell < L.hw;
        float v[16];
        // running class argmax over the class columns [c_begin, c_end) of this lane's cell (bias added in fp32)
        auto scan_classes = [&](int c_begin, int c_end, float& best, int& arg, float& prev) {
          for (int c0 = c_begin; c0 < c_end; c0 += 16) {
            tcgen05_ld16(trow + kHfBoxN + c0, v);
            if (c0 + 16 <= p.nc) {
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                v[k] = fadd(v[k], __ldg(bias_cls + c0 + k));
                class_step(v[k], c0 + k, best, arg, prev);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                if (c0 + k < p.nc) {
                  v[k] = fadd(v[k], __ldg(bias_cls + c0 + k));
                  class_step(v[k], c0 + k, best, arg, prev);
                }
              }
            }
            if (HEAD_OUT) {
              if (active) {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                  if (c0 + k < p.nc) p.head_out[((int64_t)b * (kHfBoxN + p.nc) + kHfBoxN + c0 + k) * p.A + L.anchor_off + cell] = v[k];
              }
            }
          }
        };
        const int c_split = p.cls_split;  // classes [c_split, nc) are scanned by the box warp (multiple of 16; nc = none)
        if (!cls_role) {
          // ---- box warp: 4 x (16 bins -> DFL integral), anchor + stride -> xyxy; then its share of the class scan (the
          //      class warp alone was the bottleneck: ~5500 cycles per tile against ~2000 here); box and partial scan are
          //      handed to the class warp through shared memory
          float d[4];
#pragma unroll
          for (int side = 0; side < 4; ++side) {
            tcgen05_ld16(trow + side * 16, v);
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = fadd(v[k], __ldg(bias_box + side * 16 + k));
            if (HEAD_OUT) {
              if (active) {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                  p.head_out[((int64_t)b * (kHfBoxN + p.nc) + side * 16 + k) * p.A + L.anchor_off + cell] = v[k];
              }
            }
            d[side] = dfl16(v);
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) hf_mbar_arrive(my_acc_empty);
          if (quarter == 0 && lane == 0) HF_STAMP(4, tile_i);
          const CellBox bx = cell_box(active ? cell : 0, L.w, L.stride, d[0], d[1], d[2], d[3]);
          float pbest = -INFINITY, pprev = -INFINITY;
          int parg = 0;
          if (c_split < p.nc) {
            hf_mbar_wait_relaxed(&acc_full[2 + acc], (tile_i >> 1) & 1u);  // the class branch's columns
            tcgen05_fence_after();
            scan_classes(c_split, p.nc, pbest, parg, pprev);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) hf_mbar_arrive(&acc_empty[2 + acc]);
          }
          slot->box = make_float4(bx.x1, bx.y1, bx.x2, bx.y2);
          slot->part = make_float4(pbest, __int_as_float(parg), pprev, 0.0f);
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");  // the class warp may read the slot
        } else {
          // ---- class warp: running argmax over its class logits, merge with the box warp's share, sigmoid, threshold,
          //      candidate append
          float best = -INFINITY, prev = -INFINITY;
          int arg = 0;
          scan_classes(0, min(c_split, p.nc), best, arg, prev);
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");  // the box warp has written the slot
          if (c_split < p.nc) {
            // first-index argmax over the union: the later range wins only when strictly greater; `prev` = the largest
            // logit BEFORE the winner (all of the earlier range, and the later range's own running maximum before it)
            const float4 part = slot->part;
            if (part.x > best) {
              prev = fmaxf(best, part.z);
              best = part.x;
              arg = __float_as_int(part.y);
            }
          }
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
                  const float sv = sigmoid_precise(fadd(v[k], __ldg(bias_cls + c0 + k)));
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
          if (lane == 0) hf_mbar_arrive(my_acc_empty);
          if (quarter == 0 && lane == 0) HF_STAMP(6, tile_i);
          const unsigned m = (p.debug & 4) ? 0u : __ballot_sync(0xffffffffu, cand);
          if (m) {
            pdl_wait();  // the counts have been zeroed by the programmatic predecessor
            int base = 0;
            if (lane == 0) base = atomicAdd(p.cand_count + b, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (cand) {
              const int slot_i = base + __popc(m & ((1u << lane) - 1u));
              const int anchor = L.anchor_off + cell;
              if (slot_i < p.max_cand)
                p.cand_key[(int64_t)b * p.max_cand + slot_i] = key_pack((uint32_t)arg, __float_as_uint(score), (uint32_t)anchor);
              p.box_dense[(int64_t)b * p.A + anchor] = slot->box;
            }
          }
          if (quarter == 0 && lane == 0) HF_STAMP(7, tile_i);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) { HF_STAMP(5, 1); HF_CTA(3); }
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

#ifdef CVPP_NMS_TIMING
extern "C" __attribute__((visibility("default"))) int cvpp_debug_hf_timing(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_hf_tl, sizeof(long long) * 8 * 512);
}
extern "C" __attribute__((visibility("default"))) int cvpp_debug_hf_cta(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_hf_cta, sizeof(unsigned long long) * 256 * 4);
}
#endif

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

// (cell, channel, image) view of a contiguous (B, C, H, W) feature map; box = 32 cells x 32 channels x 1 image with the
// 32-byte-atom 128-byte swizzle (the UMMA canonical MN-major atom for 32-bit operands), zero fill past the last cell
static bool hf_make_tmap(CUtensorMap* tm, const float* ptr, int hw, int C, int B) {
  HfEncodeTiledFn enc = hf_encode_fn();
  if (!enc) return false;
  if (hw % kHfAtomCells == 0) {
    // whole atoms only: split the cell index into (cell in atom, atom) so that ONE box of 32 x 32 channels x 4 atoms lands as
    // [atom][channel][32 cells] - the stage layout - with a single instruction; atoms past the last one are zero-filled
    cuuint64_t dims[4] = {(cuuint64_t)kHfAtomCells, (cuuint64_t)C, (cuuint64_t)(hw / kHfAtomCells), (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)hw * 4u, (cuuint64_t)kHfAtomCells * 4u, (cuuint64_t)hw * 4u * (cuuint64_t)C};
    cuuint32_t box[4] = {(cuuint32_t)kHfAtomCells, (cuuint32_t)kHfChunkK, (cuuint32_t)(kHfTileM / kHfAtomCells), 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  cuuint64_t dims[3] = {(cuuint64_t)hw, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)hw * 4u, (cuuint64_t)hw * 4u * (cuuint64_t)C};
  cuuint32_t box[3] = {(cuuint32_t)kHfAtomCells, (cuuint32_t)kHfChunkK, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// (k, n) view of row-major conv weights (n_rows, k); box = 32 k x box_rows with the 16-byte-chunk 128-byte swizzle: exactly
// the K-major canonical UMMA layout [row][128 B]; rows past n_rows and columns past k are zero-filled
static bool hf_make_wmap(CUtensorMap* tm, const float* ptr, int k, int n_rows, int box_rows) {
  HfEncodeTiledFn enc = hf_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)k * 4u};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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
  if (c2 < kHfChunkK || c3 < kHfChunkK || (c2 % 16) || (c3 % 16) || c2 > 1024 || c3 > 1024) {
    set_error("yolov8 head: c2=%d / c3=%d must be multiples of 16, at least %d (the reference's widths are 64 / 80..320)", c2, c3,
              kHfChunkK);
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
  {
    const int chunks16 = (nc + 15) / 16, share = (2 * chunks16) / 5;   // ~2/5 of the class columns go to the box warps
    p.cls_split = share > 0 ? 16 * (chunks16 - share) : nc;
    if (const char* e = getenv("CVPP_HEAD_SPLIT")) {  // tuning knob: chunks of 16 classes taken by the box warps
      const int v = atoi(e);
      if (v >= 0 && v < chunks16) p.cls_split = v > 0 ? 16 * (chunks16 - v) : nc;
    }
  }
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
    L.one_tma = (L.hw % kHfAtomCells) == 0;
    L.w = level_w[l];
    L.stride = level_stride[l];
    L.anchor_off = (int)A;
    L.tiles_per_image = (L.hw + kHfTileM - 1) / kHfTileM;
    L.tile_base = tiles;
    tiles += L.tiles_per_image * B;
    A += L.hw;
    if ((reinterpret_cast<uintptr_t>(box_feat[l]) & 15u) || (reinterpret_cast<uintptr_t>(cls_feat[l]) & 15u) || (L.hw & 3) ||
        (reinterpret_cast<uintptr_t>(box_w[l]) & 15u) || (reinterpret_cast<uintptr_t>(cls_w[l]) & 15u)) {
      set_error("yolov8 head: level %d feature maps / weights must be 16-byte aligned with H*W a multiple of 4 (TMA rows)", l);
      return CVPP_ERR_ALIGNMENT;
    }
  }
  if (A > CVPP_MAX_ANCHORS) {
    set_error("yolov8 head: %lld anchors exceed the %d-anchor key field", (long long)A, CVPP_MAX_ANCHORS);
    return CVPP_ERR_UNSUPPORTED;
  }
  p.A = (int)A;
  CVPP_CUDA_TRY(zero_counts_async(cand_count, B, stream));  // the kernel below is its programmatic dependent (cvpp_common.cuh)
  if (B == 0) return CVPP_OK;
  for (int l = 0; l < num_levels; ++l) {
    if (!hf_make_tmap(&p.tmap_box[l], box_feat[l], p.lv[l].hw, c2, B) || !hf_make_tmap(&p.tmap_cls[l], cls_feat[l], p.lv[l].hw, c3, B) ||
        !hf_make_wmap(&p.tmap_wbox[l], box_w[l], c2, kHfBoxN, kHfBoxN) || !hf_make_wmap(&p.tmap_wcls[l], cls_w[l], c3, nc, p.nc_pad)) {
      set_error("yolov8 head: cuTensorMapEncodeTiled failed for level %d", l);
      return CVPP_ERR_CUDA;
    }
  }
  p.w_level_bytes = p.w_box_bytes + p.w_cls_bytes;   // multiples of 1024 each: the swizzle atoms stay aligned
  const size_t misc = (2 * kHfMaxStages + 16) * sizeof(uint64_t) + 16 + 4 * kHfTileM * sizeof(HfSlot) + 1024;
  // two weight buffers (the next level's weights land while this level computes) when that leaves a ring of >= 4 stages,
  // else one (the m/l/x widths: 127 KB of weights per level at c3 = 320)
  p.w_bufs = 2 * (size_t)p.w_level_bytes + misc + 4 * kHfStageBytes <= (size_t)di.max_smem ? 2 : 1;
  const size_t fixed = (size_t)p.w_bufs * p.w_level_bytes + misc;
  // ring depth: whatever shared memory is left, at most kHfMaxStages (HBM wants ~100+ KB in flight per SM), at least 4
  int stages = fixed + 4 * kHfStageBytes <= (size_t)di.max_smem ? (int)(((size_t)di.max_smem - fixed) / kHfStageBytes) : 0;
  if (stages > kHfMaxStages) stages = kHfMaxStages;
  if (const char* e = getenv("CVPP_HEAD_STAGES")) {  // tuning knob
    const int v = atoi(e);
    if (v >= 4 && v <= stages) stages = v;
  }
  if (stages < 4) {
    set_error("yolov8 head: c2=%d c3=%d nc=%d leave no room for the activation ring in %d bytes of shared memory", c2, c3, nc,
              di.max_smem);
    return CVPP_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  {  // split in proportion to the branches' chunks per tile, at least two stages each
    const int cb = (c2 + kHfChunkK - 1) / kHfChunkK, cc = (c3 + kHfChunkK - 1) / kHfChunkK;
    int sb = (stages * cb + (cb + cc) / 2) / (cb + cc);
    if (sb < 2) sb = 2;
    if (sb > stages - 2) sb = stages - 2;
    p.stages_box = sb;
  }
  if (const char* e = getenv("CVPP_HEAD_DEBUG")) p.debug = atoi(e);
  const size_t smem = fixed + (size_t)stages * kHfStageBytes;
  static unsigned long long attr_done[2] = {0, 0};
  const int grid = tiles < di.sms ? tiles : di.sms;
  if (head_out) {
    rc = ensure_smem_attr(reinterpret_cast<const void*>(yolov8_head_fused_kernel<true>), di.max_smem, di.device, &attr_done[1]);
    if (rc != CVPP_OK) return rc;
    CVPP_CUDA_TRY(launch_pdl(yolov8_head_fused_kernel<true>, dim3(grid), dim3(kHfThreads), smem, stream, p));
  } else {
    rc = ensure_smem_attr(reinterpret_cast<const void*>(yolov8_head_fused_kernel<false>), di.max_smem, di.device, &attr_done[0]);
    if (rc != CVPP_OK) return rc;
    CVPP_CUDA_TRY(launch_pdl(yolov8_head_fused_kernel<false>, dim3(grid), dim3(kHfThreads), smem, stream, p));
  }
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

// ssd_decode.cu — SSD prior decode + softmax + per-(prior, class) confidence filter (sm_100a).
//
// Replaces (reference file:line): Ssd.decode_boxes core/algorithms/ssd.py:236-264 (softmax :248,
// per-class threshold mask :256-264) and Ssd._parse_mbox_loc :290-325.  The per-class NMS that follows
// (:267) is cvpp_segmented_sort + cvpp_nms with CVPP_NMS_RULE_PER_CLASS / CVPP_ORDER_CLASS_MAJOR.
//
// Memory-bound on loc (16 B) + conf (4*(nc+1) B) per prior = 100 B at nc = 20 (873 200 B per 8732-prior
// image); the priors table (140 KB) stays L2-resident.  A CTA stages the contiguous conf rows of 256
// priors through shared memory with coalesced loads (a prior's nc+1 logits are contiguous, so a
// thread-per-prior global read would be strided); thread-per-prior then reads its row conflict-free
// ((nc+1) odd), runs the softmax, decodes the box once, and emits one key per class above the threshold
// (one ballot + warp-aggregated atomic per class with any hit).
#include <cstdlib>

#include "cvpp_common.cuh"

namespace cvpp {

constexpr int kSsdThreads = 256;
constexpr int kSsdKeyStage = 512;  // keys staged per block before the single global reservation

// Ssd._parse_mbox_loc for one prior: ltrb prior + (dx, dy, dw, dh) -> clamped xyxy (ssd.py:293-324)
__device__ __forceinline__ float4 ssd_decode_box(const float4& a, const float4& l) {
  const float v0 = 0.1f, v1 = 0.2f;  // variance[::2]
  const float aw = fsub(a.z, a.x), ah = fsub(a.w, a.y);
  const float acx = fmul(0.5f, fadd(a.z, a.x)), acy = fmul(0.5f, fadd(a.w, a.y));
  const float cx = fadd(fmul(fmul(l.x, aw), v0), acx);
  const float cy = fadd(fmul(fmul(l.y, ah), v0), acy);
  const float w = fmul(expf(fmul(l.z, v1)), aw);
  const float h = fmul(expf(fmul(l.w, v1)), ah);
  const float hw = fmul(0.5f, w), hh = fmul(0.5f, h);
  float4 o;
  o.x = fminf(fmaxf(fsub(cx, hw), 0.0f), 1.0f);
  o.y = fminf(fmaxf(fsub(cy, hh), 0.0f), 1.0f);
  o.z = fminf(fmaxf(fadd(cx, hw), 0.0f), 1.0f);
  o.w = fminf(fmaxf(fadd(cy, hh), 0.0f), 1.0f);
  return o;
}

__device__ __forceinline__ float ssd_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Two phases per 256-prior block (first version: 21 + 20 precise expf and 20 IEEE divisions for EVERY prior made
// the kernel SFU/ALU-bound at 12 % of the HBM roofline, although only ~1 % of the (prior, class) pairs survive):
//   1. thread per prior, approximate: softmax denominator with ex2.approx, then one logit cut per prior -
//      prob_c > thr  <=>  x_c > m + ln(thr * sum) - and a compare per class; priors with any class above the
//      (slightly lowered) cut are compacted into a shared list;
//   2. thread per LISTED prior, exact: the reference's softmax arithmetic (max-subtract, expf, sequential sum,
//      IEEE division) for the classes above the cut, the threshold test on the exact probability, the key, and
//      the prior's decoded box.
__global__ void __launch_bounds__(kSsdThreads)
ssd_decode_filter_kernel(const float4* __restrict__ loc, const float* __restrict__ conf,
                         const float4* __restrict__ priors, int P, int nc, float conf_thres,
                         uint64_t* __restrict__ cand_key, int32_t* __restrict__ cand_count,
                         float4* __restrict__ box_dense, int max_cand) {
  extern __shared__ float sm[];  // [kSsdThreads][nc + 1]
  __shared__ int sh_list[kSsdThreads];
  __shared__ float sh_cut[kSsdThreads];
  __shared__ int sh_n, sh_k, sh_gbase;
  __shared__ uint64_t sh_key[kSsdKeyStage];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int p0 = blockIdx.x * kSsdThreads;
  const int rows = min(kSsdThreads, P - p0);
  const int nc1 = nc + 1;
  const float* src = conf + ((int64_t)b * P + p0) * nc1;
  if (tid == 0) {
    sh_n = 0;
    sh_k = 0;
  }
  {  // the block's conf rows are one contiguous run: 128-bit loads when its start is 16-byte aligned
    const int total = rows * nc1;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
      const int n4 = total >> 2;
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(sm);
      const uint64_t policy = l2_policy_evict_first();
      for (int i = tid; i < n4; i += kSsdThreads) d4[i] = ldg_stream_f4(s4 + i, policy);
      for (int i = (n4 << 2) + tid; i < total; i += kSsdThreads) sm[i] = __ldg(src + i);
    } else {
      for (int i = tid; i < total; i += kSsdThreads) sm[i] = __ldg(src + i);
    }
  }
  __syncthreads();

  // ---- phase 1: approximate per-prior cut
  {
    const float* x = sm + tid * nc1;
    bool cand = false;
    float cut = INFINITY;
    if (tid < rows) {
      float m = x[0];
      for (int k = 1; k < nc1; ++k) m = fmaxf(m, x[k]);
      const float kLog2e = 1.4426950408889634f;
      float sum = 0.0f;
      for (int k = 0; k < nc1; ++k) sum += ssd_ex2((x[k] - m) * kLog2e);
      // exact: prob_c = exp(x_c - m) / sum > thr.  The approximations (ex2.approx, fast log, different
      // summation order) are far below the 0.02 slack taken off the cut.
      cut = conf_thres > 0.0f ? m + __logf(conf_thres * sum) - 0.02f : -INFINITY;
      float top = -INFINITY;
      for (int k = 1; k < nc1; ++k) top = fmaxf(top, x[k]);
      cand = top >= cut;
    }
    const unsigned mk = __ballot_sync(0xffffffffu, cand);
    if (mk) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&sh_n, __popc(mk));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (cand) {
        const int slot = base + __popc(mk & ((1u << lane) - 1u));
        sh_list[slot] = tid;
        sh_cut[slot] = cut;
      }
    }
  }
  __syncthreads();

  // ---- phase 2: exact evaluation of the listed priors; keys are staged in shared memory and leave with ONE
  //      global reservation per block (a per-hit atomicAdd with its L2 round trip serialised the threads)
  const int n_list = sh_n;
  if (n_list == 0) return;  // uniform
  // 8 lanes per listed prior (a thread per prior left 6 of the 8 warps waiting at the barrier while 2 ran 21
  // precise expf each): the lanes split the nc + 1 logits, combine max and sum with three shuffles, and each
  // tests its own classes.  (The sum is taken as per-lane partial sums + a butterfly instead of sequentially;
  // scores stay within the 1e-5 tolerance, which is all the fp32 softmax is pinned to.)
  {
    const int grp = tid >> 3, gl = tid & 7;
    for (int base = 0; base < n_list; base += kSsdThreads / 8) {  // uniform trip count: shuffles inside
      const int i = base + grp;
      const bool valid = i < n_list;
      const int row = valid ? sh_list[i] : 0;
      const float cut = valid ? sh_cut[i] : INFINITY;
      const float* x = sm + row * nc1;
      const int pr = p0 + row;
      float m = -INFINITY;
      for (int k = gl; k < nc1; k += 8) m = fmaxf(m, x[k]);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
      float sum = 0.0f;
      for (int k = gl; k < nc1; k += 8) sum = fadd(sum, expf(fsub(x[k], m)));  // torch.softmax: exp(x - max) / sum
      sum = fadd(sum, __shfl_xor_sync(0xffffffffu, sum, 1));
      sum = fadd(sum, __shfl_xor_sync(0xffffffffu, sum, 2));
      sum = fadd(sum, __shfl_xor_sync(0xffffffffu, sum, 4));
      bool any = false;
      if (valid) {
        for (int c = gl; c <= nc; c += 8) {
          if (c == 0 || !(x[c] >= cut)) continue;
          const float prob = fdiv(expf(fsub(x[c], m)), sum);
          if (!(prob > conf_thres)) continue;
          const uint64_t key = key_pack((uint32_t)(c - 1), __float_as_uint(prob), (uint32_t)pr);
          const int pos = atomicAdd(&sh_k, 1);
          if (pos < kSsdKeyStage) {
            sh_key[pos] = key;
          } else {  // stage full (a dense block): reserve directly
            const int slot = atomicAdd(cand_count + b, 1);
            if (slot < max_cand) cand_key[(int64_t)b * max_cand + slot] = key;
          }
          any = true;
        }
      }
      const unsigned am = __ballot_sync(0xffffffffu, any);
      if (gl == 0 && ((am >> (lane & 24)) & 0xffu))
        box_dense[(int64_t)b * P + pr] = ssd_decode_box(priors[pr], loc[(int64_t)b * P + pr]);
    }
  }
  __syncthreads();
  const int n_stage = min(sh_k, kSsdKeyStage);
  if (n_stage == 0) return;  // uniform
  if (tid == 0) sh_gbase = atomicAdd(cand_count + b, n_stage);
  __syncthreads();
  const int gbase = sh_gbase;
  for (int i = tid; i < n_stage; i += kSsdThreads)
    if (gbase + i < max_cand) cand_key[(int64_t)b * max_cand + gbase + i] = sh_key[i];
}

// -----------------------------------------------------------------------------------------------
// Streaming version (default): persistent CTAs of independent warps, as in yolov8_decode.cu.  A warp owns tiles of
// 32 consecutive priors of one image; the tile's conf rows are ONE contiguous run of 128 (nc + 1) bytes that a
// single 1-D bulk copy (TMA engine, L2 evict-first) lands in the warp's private 3-stage ring, so two tiles are in
// flight while one is evaluated and no CTA barrier is ever met.  (The block-synchronous version above loads,
// waits and then computes: 2.1 TB/s.)  Same two phases, same arithmetic; keys are staged per warp and leave
// with one reservation per tile.
constexpr int kSsdTile = 32;        // priors per tile (one per lane)
constexpr int kSsdStages = 2;
constexpr int kSsdWarpKeys = 96;    // per-warp key stage: at most 32 keys are appended per step
constexpr int kSsdPairCap = 128;    // per-warp (listed prior, class) pairs awaiting exact evaluation
constexpr int kSsdMaxWarps = 32;    // the per-tile arithmetic is a latency chain: many warps, small tiles

struct SsdParams {
  const float4* loc;
  const float* conf;
  const float4* priors;
  int P, nc, B;
  float conf_thres;
  uint64_t* cand_key;
  int32_t* cand_count;
  float4* box_dense;
  int max_cand;
  int tiles_per_image, total_tiles;
};

// NC1 > 0: nc + 1 known at compile time (21 for VOC) - the per-prior loops unroll and their loads overlap
template <int NC1>
__global__ void __launch_bounds__(kSsdMaxWarps * 32, 1) ssd_stream_kernel(const __grid_constant__ SsdParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int nc = NC1 > 0 ? NC1 - 1 : p.nc, nc1 = nc + 1, P = p.P;
  const int chunk_floats = kSsdTile * nc1;
  float* ring = reinterpret_cast<float*>(smem_raw) + (size_t)warp * kSsdStages * chunk_floats;
  unsigned char* q = smem_raw + (size_t)warps * kSsdStages * chunk_floats * sizeof(float);
  uint64_t* bar = reinterpret_cast<uint64_t*>(q) + warp * kSsdStages;
  q += (size_t)warps * kSsdStages * sizeof(uint64_t);
  uint64_t* keys = reinterpret_cast<uint64_t*>(q) + warp * kSsdWarpKeys;
  q += (size_t)warps * kSsdWarpKeys * sizeof(uint64_t);
  int* list = reinterpret_cast<int*>(q) + warp * kSsdTile;
  q += (size_t)warps * kSsdTile * sizeof(int);
  float* cuts = reinterpret_cast<float*>(q) + warp * kSsdTile;
  q += (size_t)warps * kSsdTile * sizeof(float);
  float2* msum = reinterpret_cast<float2*>(q) + warp * kSsdTile;
  q += (size_t)warps * kSsdTile * sizeof(float2);
  uint16_t* pairs = reinterpret_cast<uint16_t*>(q) + warp * kSsdPairCap;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kSsdStages; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  __syncwarp();
  const uint64_t policy = l2_policy_evict_first();

  // tiles of this warp (the last, partial round is dealt SM-minor so that every SM keeps streaming)
  const int stride_tiles = gridDim.x * warps;
  const int first = blockIdx.x * warps + warp;
  const int full_rounds = p.total_tiles / stride_tiles;
  const int last_slot = warp * gridDim.x + blockIdx.x;
  const int n_tiles = full_rounds + (full_rounds * stride_tiles + last_slot < p.total_tiles ? 1 : 0);
  auto tile_of = [&](int r) { return r * stride_tiles + (r < full_rounds ? first : last_slot); };
  // g / tiles_per_image without the ~25-instruction integer division (two per tile were 5.6 % of the kernel): float
  // reciprocal estimate, corrected by at most one (g < 2^24 tiles is checked on the host)
  const float inv_tpi = 1.0f / (float)p.tiles_per_image;
  auto image_of = [&](int g) {
    int b = (int)((float)g * inv_tpi);
    const int rem = g - b * p.tiles_per_image;
    b += rem >= p.tiles_per_image ? 1 : (rem < 0 ? -1 : 0);
    return b;
  };
  int pr_round = 0;
  auto issue = [&]() {
    const int g = tile_of(pr_round);
    const int b = image_of(g);
    const int p0 = (g - b * p.tiles_per_image) * kSsdTile;
    const int rows = min(kSsdTile, P - p0);
    if (lane == 0) {
      uint64_t* fb = &bar[pr_round % kSsdStages];
      const uint32_t bytes = (uint32_t)(rows * nc1) * (uint32_t)sizeof(float);
      mbar_arrive_expect_tx(fb, bytes);
      bulk_g2s_hint(ring + (pr_round % kSsdStages) * chunk_floats, p.conf + ((int64_t)b * P + p0) * nc1, bytes, fb, policy);
    }
    ++pr_round;
  };
  for (int r = 0; r < kSsdStages && r < n_tiles; ++r) issue();

  const unsigned lt = (1u << lane) - 1u;
  const int grp = lane >> 3, gl = lane & 7;
  for (int r = 0; r < n_tiles; ++r) {
    const int g = tile_of(r);
    const int b = image_of(g);
    const int p0 = (g - b * p.tiles_per_image) * kSsdTile;
    const int rows = min(kSsdTile, P - p0);
    const int s = r % kSsdStages;
    mbar_wait(&bar[s], (uint32_t)(r / kSsdStages) & 1u);
    const float* x0 = ring + s * chunk_floats;

    // ---- phase 1: approximate per-prior cut (one prior per lane; odd nc + 1 makes the rows conflict-free)
    int n_list = 0;
#pragma unroll
    for (int u = 0; u < kSsdTile / 32; ++u) {
      const int row = lane + 32 * u;
      bool cand = false;
      float cut = INFINITY, m_exact = 0.0f;
      uint32_t cmask = 0;
      float xv_keep[NC1 > 0 ? NC1 : 1];
      if (row < rows) {
        const float* x = x0 + row * nc1;
        float top = -INFINITY;
        const float kLog2e = 1.4426950408889634f;
        float sum = 0.0f, m;
        if (NC1 > 0) {
          float xv[NC1 > 0 ? NC1 : 1];
#pragma unroll
          for (int k = 0; k < NC1; ++k) xv[k] = x[k];
#pragma unroll
          for (int k = 0; k < NC1; ++k) xv_keep[k] = xv[k];
#pragma unroll
          for (int k = 1; k < NC1; ++k) top = fmaxf(top, xv[k]);
          m = fmaxf(xv[0], top);
          float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
          for (int k = 0; k < NC1; ++k) {
            const float e = ssd_ex2((xv[k] - m) * kLog2e);
            if (k & 1) s1 += e;
            else s0 += e;
          }
          sum = s0 + s1;
        } else {
          for (int k = 1; k < nc1; ++k) top = fmaxf(top, x[k]);
          m = fmaxf(x[0], top);
          for (int k = 0; k < nc1; ++k) sum += ssd_ex2((x[k] - m) * kLog2e);
        }
        cut = p.conf_thres > 0.0f ? m + __logf(p.conf_thres * sum) - 0.02f : -INFINITY;
        cand = top >= cut;
        m_exact = m;  // fmaxf does not round: phase 1's maximum IS the reference's softmax maximum
        if (NC1 > 0 && NC1 <= 32) {
          // classes above the cut, as a bit mask, while the logits are in registers (phase 2 re-scanned the listed priors'
          // rows with 8 lanes, a ballot and a conditional store per 8 classes: 16 % of the kernel)
#pragma unroll
          for (int k = 1; k < (NC1 > 0 ? NC1 : 1); ++k) cmask |= (xv_keep[k] >= cut) ? (1u << k) : 0u;
        }
      }
      const unsigned mk = __ballot_sync(0xffffffffu, cand);
      if (cand) {
        const int slot = n_list + __popc(mk & lt);
        list[slot] = row;
        if (NC1 > 0 && NC1 <= 32) reinterpret_cast<uint32_t*>(cuts)[slot] = cmask;
        else cuts[slot] = cut;
        msum[slot].x = m_exact;
      }
      n_list += __popc(mk);
    }
    __syncwarp();

    // ---- phase 2: exact evaluation of the listed priors.  (a) 8 lanes per prior: the reference's softmax
    //      maximum and denominator (per-lane partial sums + butterfly, as in the block version); (b) the same lanes
    //      list the (prior, class) pairs above the cut; (c) the pairs are evaluated 32 at a time at full lane
    //      efficiency; (d) the box of every prior with a surviving class is decoded once, a lane per prior.
    int wk_n = 0;
    auto flush = [&]() {
      __syncwarp();
      pdl_wait();  // the counts have been zeroed by the programmatic predecessor
      int base = 0;
      if (lane == 0) base = atomicAdd(p.cand_count + b, wk_n);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int i = lane; i < wk_n; i += 32)
        if (base + i < p.max_cand) p.cand_key[(int64_t)b * p.max_cand + base + i] = keys[i];
      __syncwarp();
      wk_n = 0;
    };
    int n_pairs = 0;
    unsigned boxed = 0;  // listed priors with at least one surviving class
    auto eval_pairs = [&]() {
      __syncwarp();
      for (int base = 0; base < n_pairs; base += 32) {
        const int j = base + lane;
        bool hit = false;
        uint64_t key = 0;
        int i = 0;
        if (j < n_pairs) {
          const int pc = pairs[j];
          i = pc >> 8;
          const int c = pc & 0xff, row = list[i];
          const float2 ms = msum[i];
          const float prob = fdiv(expf(fsub(x0[row * nc1 + c], ms.x)), ms.y);  // torch.softmax: exp(x - max) / sum
          if (prob > p.conf_thres) {
            hit = true;
            key = key_pack((uint32_t)(c - 1), __float_as_uint(prob), (uint32_t)(p0 + row));
          }
        }
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hm) {
          boxed |= __reduce_or_sync(0xffffffffu, hit ? (1u << i) : 0u);
          if (wk_n + __popc(hm) > kSsdWarpKeys) flush();
          if (hit) keys[wk_n + __popc(hm & lt)] = key;
          wk_n += __popc(hm);
        }
      }
      __syncwarp();
      n_pairs = 0;
    };
    if (NC1 > 0 && NC1 <= 32) {
      // (a) exact denominators, 8 lanes per listed prior (the maximum comes from phase 1)
      for (int base = 0; base < n_list; base += 4) {
        const int i = base + grp;
        const bool valid = i < n_list;
        const int row = valid ? list[i] : 0;
        const float m = valid ? msum[i].x : 0.0f;
        const float* x = x0 + row * nc1;
        float sum = 0.0f;
#pragma unroll
        for (int k0 = 0; k0 < (NC1 > 0 ? NC1 : 1); k0 += 8)
          if (k0 + gl < nc1) sum = fadd(sum, expf(fsub(x[k0 + gl], m)));
        sum = fadd(sum, __shfl_xor_sync(0xffffffffu, sum, 1));
        sum = fadd(sum, __shfl_xor_sync(0xffffffffu, sum, 2));
        sum = fadd(sum, __shfl_xor_sync(0xffffffffu, sum, 4));
        if (valid && gl == 0) msum[i].y = sum;
      }
      // (b) the (prior, class) pairs of the masks, compacted across the lanes: lane i owns listed prior i
      uint32_t my = lane < n_list ? reinterpret_cast<const uint32_t*>(cuts)[lane] : 0u;
      while (__any_sync(0xffffffffu, my != 0u)) {
        const int take = min(__popc(my), kSsdPairCap / 32);   // at most 4 per lane per round: a round never overflows the list
        int incl = take;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += u;
        }
        int pos = incl - take;
        for (int q = 0; q < take; ++q) {
          const int c = __ffs(my) - 1;
          my &= my - 1u;
          pairs[pos++] = (uint16_t)((lane << 8) | c);
        }
        n_pairs = __shfl_sync(0xffffffffu, incl, 31);
        eval_pairs();
      }
    } else {
    for (int base = 0; base < n_list; base += 4) {
      const int i = base + grp;
      const bool valid = i < n_list;
      const int row = valid ? list[i] : 0;
      const float cut = valid ? cuts[i] : INFINITY;
      const float* x = x0 + row * nc1;
      float m = -INFINITY;
      for (int k = gl; k < nc1; k += 8) m = fmaxf(m, x[k]);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
      float sum = 0.0f;
      for (int k = gl; k < nc1; k += 8) sum = fadd(sum, expf(fsub(x[k], m)));
      sum = fadd(sum, __shfl_xor_sync(0xffffffffu, sum, 1));
      sum = fadd(sum, __shfl_xor_sync(0xffffffffu, sum, 2));
      sum = fadd(sum, __shfl_xor_sync(0xffffffffu, sum, 4));
      if (valid && gl == 0) msum[i] = make_float2(m, sum);
      for (int c0 = 0; c0 <= nc; c0 += 8) {  // uniform trip count: ballots inside
        const int c = c0 + gl;
        const bool above = valid && c >= 1 && c <= nc && x[c] >= cut;
        const unsigned am = __ballot_sync(0xffffffffu, above);
        if (am) {
          if (n_pairs + __popc(am) > kSsdPairCap) eval_pairs();  // (msum of this pass is visible: __syncwarp inside)
          if (above) pairs[n_pairs + __popc(am & lt)] = (uint16_t)((i << 8) | c);
          n_pairs += __popc(am);
        }
      }
    }
    }
    if (n_pairs) eval_pairs();
    if (lane < n_list && ((boxed >> lane) & 1u)) {
      const int pr = p0 + list[lane];
      p.box_dense[(int64_t)b * P + pr] = ssd_decode_box(p.priors[pr], p.loc[(int64_t)b * P + pr]);
    }
    if (wk_n) flush();
    __syncwarp();  // every lane is done with the stage: it may be refilled
    if (pr_round < n_tiles) issue();
  }
}

__global__ void __launch_bounds__(256)
ssd_parse_loc_kernel(const float4* __restrict__ loc, const float4* __restrict__ priors, int P, int64_t total,
                     float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = ssd_decode_box(priors[i % P], loc[i]);
}

int ssd_decode_filter_launch(const float* loc, const float* conf, const float* priors, int B, int P, int nc,
                             float conf_thres, uint64_t* cand_key, int32_t* cand_count, float* box_dense, int max_cand,
                             cudaStream_t stream) {
  if (!loc || !conf || !priors || !cand_key || !cand_count || !box_dense) {
    set_error("ssd_decode_filter: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || P < 1 || P > CVPP_MAX_ANCHORS || nc < 1 || nc > CVPP_MAX_CLASSES || max_cand < 1) {
    set_error("ssd_decode_filter: bad sizes (B=%d P=%d nc=%d max_cand=%d)", B, P, nc, max_cand);
    return CVPP_ERR_INVALID_ARG;
  }
  if (!(conf_thres >= 0.0f && conf_thres <= 1.0f)) {
    set_error("ssd_decode_filter: confidence threshold %f outside [0, 1]", conf_thres);
    return CVPP_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(loc) | reinterpret_cast<uintptr_t>(priors) | reinterpret_cast<uintptr_t>(box_dense)) & 15u) {
    set_error("ssd_decode_filter: loc, priors and box_dense must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  CVPP_CUDA_TRY(zero_counts_async(cand_count, B, stream));  // the stream kernel is its programmatic dependent (cvpp_common.cuh)
  if (B == 0) return CVPP_OK;
  {
    // streaming kernel: needs 16-byte aligned conf rows per image and tile (bulk copies) and room for >= 4 warps
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc != CVPP_OK) return rc;
    const int nc1 = nc + 1;
    const size_t per_warp = (size_t)kSsdStages * kSsdTile * nc1 * sizeof(float) + kSsdStages * sizeof(uint64_t) +
                            kSsdWarpKeys * sizeof(uint64_t) + kSsdTile * (sizeof(int) + sizeof(float) + sizeof(float2)) +
                            kSsdPairCap * sizeof(uint16_t);
    int wmax = (int)(((size_t)di.max_smem - 1024) / per_warp);
    if (wmax > kSsdMaxWarps) wmax = kSsdMaxWarps;
    const bool aligned = (reinterpret_cast<uintptr_t>(conf) & 15u) == 0 && (((int64_t)P * nc1) & 3) == 0;
    const char* force = getenv("CVPP_SSD_BLOCK_KERNEL");
    // (the kernel's reciprocal tile -> image division is exact below 2^24 tiles)
    const bool few_tiles = (int64_t)((P + kSsdTile - 1) / kSsdTile) * B < (1 << 24);
    if (aligned && wmax >= 4 && nc <= 255 && few_tiles && !(force && force[0] == '1')) {
      SsdParams sp{};
      sp.loc = reinterpret_cast<const float4*>(loc);
      sp.conf = conf;
      sp.priors = reinterpret_cast<const float4*>(priors);
      sp.P = P;
      sp.nc = nc;
      sp.B = B;
      sp.conf_thres = conf_thres;
      sp.cand_key = cand_key;
      sp.cand_count = cand_count;
      sp.box_dense = reinterpret_cast<float4*>(box_dense);
      sp.max_cand = max_cand;
      sp.tiles_per_image = (P + kSsdTile - 1) / kSsdTile;
      sp.total_tiles = sp.tiles_per_image * B;
      // warps per CTA: fill whole rounds of tiles (see pick_shape in yolov8_decode.cu)
      int warps = wmax;
      if (sp.total_tiles <= di.sms * wmax) {
        warps = (sp.total_tiles + di.sms - 1) / di.sms;
      } else {
        const int T = (sp.total_tiles + di.sms - 1) / di.sms;
        int best = ((T + wmax - 1) / wmax) * wmax;
        for (int w = wmax - 1; w >= wmax - 4 && w >= 4; --w) {
          const int slots = ((T + w - 1) / w) * w;
          if (slots < best) {
            best = slots;
            warps = w;
          }
        }
      }
      const char* ew = getenv("CVPP_SSD_WARPS");
      if (ew && atoi(ew) >= 1 && atoi(ew) <= wmax) warps = atoi(ew);
      const int want = (sp.total_tiles + warps - 1) / warps;
      const int grid = want < di.sms ? want : di.sms;
      static unsigned long long attr_done = 0;
      static unsigned long long attr_done21 = 0;
      if (nc1 == 21) {
        rc = ensure_smem_attr(reinterpret_cast<const void*>(ssd_stream_kernel<21>), di.max_smem, di.device, &attr_done21);
        if (rc != CVPP_OK) return rc;
        CVPP_CUDA_TRY(launch_pdl(ssd_stream_kernel<21>, dim3(grid), dim3(warps * 32), per_warp * warps, stream, sp));
      } else {
        rc = ensure_smem_attr(reinterpret_cast<const void*>(ssd_stream_kernel<0>), di.max_smem, di.device, &attr_done);
        if (rc != CVPP_OK) return rc;
        CVPP_CUDA_TRY(launch_pdl(ssd_stream_kernel<0>, dim3(grid), dim3(warps * 32), per_warp * warps, stream, sp));
      }
      CVPP_CUDA_TRY(cudaGetLastError());
      return CVPP_OK;
    }
  }
  const size_t smem = (size_t)kSsdThreads * (nc + 1) * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("ssd_decode_filter: nc=%d needs more than 48 KB of staging shared memory", nc);
    return CVPP_ERR_UNSUPPORTED;
  }
  dim3 grid((unsigned)((P + kSsdThreads - 1) / kSsdThreads), (unsigned)B);
  ssd_decode_filter_kernel<<<grid, kSsdThreads, smem, stream>>>(
      reinterpret_cast<const float4*>(loc), conf, reinterpret_cast<const float4*>(priors), P, nc, conf_thres, cand_key,
      cand_count, reinterpret_cast<float4*>(box_dense), max_cand);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

int ssd_parse_loc_launch(const float* loc, const float* priors, int B, int P, float* out, cudaStream_t stream) {
  if (!loc || !priors || !out || B < 0 || P < 1) {
    set_error("ssd_parse_loc: NULL pointer or bad sizes");
    return CVPP_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(loc) | reinterpret_cast<uintptr_t>(priors) | reinterpret_cast<uintptr_t>(out)) & 15u) {
    set_error("ssd_parse_loc: pointers must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  const int64_t total = (int64_t)B * P;
  if (total == 0) return CVPP_OK;
  ssd_parse_loc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(loc), reinterpret_cast<const float4*>(priors), P, total,
      reinterpret_cast<float4*>(out));
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

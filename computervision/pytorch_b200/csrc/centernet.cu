// centernet.cu — kernel 4: CenterNet heatmap peak extraction + top-K + box assembly + DIoU-NMS (sm_100a).
//
// Replaces (reference file:line): CenterNetA.decode_boxes core/algorithms/centernet.py:271-314,
// _suppress_redundant_centers :316-326, _top_k :328-338, RegL1Loss.gather_feat
// core/loss/centernet_loss.py:37-43, xywh_to_xyxy_torch core/utils/bboxes.py:29-49, diou_nms
// core/utils/nms.py:9-31 (+ box_diou / box_iou core/utils/iou.py:8-64), reverse_letter_box
// core/utils/image_process.py:100-129.
//
// Reference quirk kept on purpose: the reference applies MaxPool2d(3,1,1) to the
// NHWC tensor, so the 3x3 window spans (x, class) of one image row y, not (y, x).  Rows are therefore
// independent, which is what makes the single streaming pass below possible.
//
// Pass A (memory-bound, reads the (B,H,W,nc+4) tensor exactly once): persistent CTAs walk image rows;
// a row (W*(nc+4) contiguous floats, 43 KB at 128x84) is one bulk async copy (UBLKCP) into a 2-stage
// shared-memory ring.  Threads test each heat cell against its 8 neighbours in the LOGIT domain
// (sigmoid is monotone; equality of rounded sigmoids is re-checked only when logits nearly tie),
// turn peaks into 64-bit keys [inv_score | flat index] and keep those not worse than the image's
// running K-th best key `tau` (atomicMin, only ever tightened by rows that hold >= K peaks, which
// they sort with the hybrid bitonic network).  Each row appends at most K keys to the image's list.
// Pass B (one CTA per image): sort the list, take K, gather reg/wh, assemble + clamp boxes, score
// mask, optional class-agnostic DIoU-NMS (greedy, one barrier per candidate), letterbox inverse.
#include <cstdlib>

#include "cvpp_common.cuh"

namespace cvpp {

constexpr int kCnThreads = 1024;

struct CnParams {
  const float* pred;
  int B, H, W, nc, K;
  unsigned long long* tau;  // [B] running upper bound of the K-th best key
  int* tau_logit;           // [B] the same bound as a heat LOGIT (conservative, cn_enc-encoded), for the cheap in-loop test
  uint32_t* hist;           // [B][kCnBins] scores of every emitted peak, best bin first
  uint64_t* list;           // [B][list_cap]
  int32_t* list_count;      // [B]
  int list_cap;             // H * K
  int key_cap;              // shared key buffer (power of two >= W * nc)
  // pass B
  float conf;
  int use_nms;
  float nms_thr;
  const float* letterbox;  // [B][5]: in_w, in_h, left, top, scale  (or NULL)
  float4* det_box;
  float* det_score;
  int32_t* det_cls;
  int32_t* det_pixel;
  int32_t* det_count;
  uint64_t* ws_sort;  // [B][sort_cap] for lists too large for shared memory
  int sort_cap;
  // tile kernel (primary pass A) + exact redo of the images whose list overflowed
  int tile_cols, tiles_per_row, total_tiles, tile_stage_floats;
  int tile_lo;        // the launch covers units [tile_lo, total_tiles)
  int row_lo, row_hi; // the row kernel covers image rows [row_lo, row_hi)
  int32_t* redo;      // [B] 1 = the image is redone by the row kernel
  int32_t* redo_any;  // [1]
};

__device__ __forceinline__ uint64_t cn_key(float score, uint32_t flat) {
  return ((uint64_t)(0x7fffffffu - (__float_as_uint(score) & 0x7fffffffu)) << 32) | (uint64_t)flat;
}

// ---------------------------------------------------------------------------------------------
// pass A
// ---------------------------------------------------------------------------------------------
constexpr int kCnPad = 32;          // per-image scalars (tau, tau_logit, list_count) sit 128 B apart: they are
                                    // hammered by atomics and bound reads from every CTA, and 64 of them packed into
                                    // two cache lines serialise on one L2 slice
constexpr int kCnBins = 1024;       // per-image histogram of emitted peak scores (float-bit bins, best first)
constexpr int kCnGroup = 30;        // classes per warp item (lanes 1..30; lanes 0 and 31 are halo classes)
constexpr float kCnTieEps = 1e-3f;  // logits closer than this may round to the same sigmoid

__device__ __forceinline__ int cn_score_bin(uint64_t key) {
  // inv_score = 0x7fffffff - bits(score); scores in (0, 1] -> bits <= 0x3f800000
  const uint32_t bits = 0x7fffffffu - (uint32_t)(key >> 32);
  const int bin = (int)((0x3f800000u - min(bits, 0x3f800000u)) >> 17);
  return bin < kCnBins - 1 ? bin : kCnBins - 1;
}

// conservative logit bound of "score < lower edge of bin e": a peak whose logit is below it is worse than
// every peak counted in bins 0..e
__device__ __forceinline__ float cn_bin_logit_bound(int e) {
  if (e >= kCnBins - 1) return -INFINITY;
  const float s = __uint_as_float(0x3f800000u - ((uint32_t)(e + 1) << 17));
  if (!(s > 0.0f) || !(s < 1.0f)) return -INFINITY;
  return logf(s / (1.0f - s)) - 4.0f * kCnTieEps;
}

// The logit bound is stored as a signed int whose order is the floats' order, so it can be raised with a
// fire-and-forget atomicMax instead of a compare-and-swap loop (two L2 round trips on the critical path).
__device__ __forceinline__ int cn_enc(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float cn_dec(int e) { return __int_as_float(e >= 0 ? e : e ^ 0x7fffffff); }
__device__ __forceinline__ void atomic_max_float(int* addr, float v) { atomicMax(addr, cn_enc(v)); }
// L2 read of a value other CTAs keep raising: never cached in L1, never merged with an earlier read, and - unlike
// a `volatile` access - free to be scheduled among the streaming loads (ld.volatile serialised them: 2x slower)
__device__ __forceinline__ int ld_cg_s32(const int* p) {
  int v;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

constexpr int kCnAThreads = 512;  // pass A: 16 warps, two CTAs per SM (four 43 KB rows in flight per SM)
constexpr int kCnAWarps = kCnAThreads / 32;
constexpr int kCnStage = 2048;    // staged keys per row before they are cut to the best K

// Row scan.  FAST PATH: a warp owns whole columns x; lane l holds the 4 consecutive classes 4l..4l+3 of the
// column (one LDS.128) and only asks "is any of them >= thr" (thr = the image's running logit bound): 3 max,
// 1 compare, 1 vote per 4*32 cells.  Once the bound is tight almost no column passes.  SLOW PATH (a lane
// with a cell above the bound): the 3x3 (x, class) window maximum is read from the staged row, the
// reference's `heatmap == maxpool(heatmap)` test is evaluated on the rounded sigmoids, and the peak is
// staged as a key.  Returns nothing; sh_cnt counts every push (also the ones that did not fit).
__device__ __forceinline__ void cn_scan_columns(const float* row, int x_begin, int x_end, int y, int W, int nc, int Cf,
                                                float thr, unsigned long long tau, uint64_t* keys, int* sh_cnt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec = (Cf & 3) == 0;
  for (int x = x_begin + warp; x < x_end; x += kCnAWarps) {
    const float* colx = row + x * Cf;
    for (int c0 = 4 * lane; c0 < nc; c0 += 128) {  // one pass for nc <= 128 (uniform trip count + 1 for the vote)
      float v[4];
      if (vec && c0 + 3 < nc) {
        const float4 q = *reinterpret_cast<const float4*>(colx + c0);
        v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = c0 + k < nc ? colx[c0 + k] : -INFINITY;
      }
      const float mx = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
      if (!(mx >= thr)) continue;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + k;
        if (!(v[k] >= thr) || c >= nc) continue;
        float m = v[k];  // 3x3 window maximum over (x-1..x+1, c-1..c+1), -inf outside the map
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= W) continue;
          const float* q = row + xx * Cf + c;
          if (c > 0) m = fmaxf(m, q[-1]);
          m = fmaxf(m, q[0]);
          if (c + 1 < nc) m = fmaxf(m, q[1]);
        }
        // heatmap == maxpool(heatmap) on the ROUNDED sigmoids: v == m implies it; below the maximum the
        // rounded sigmoids can still coincide (nearly equal logits, or saturation)
        const float sc = sigmoid_precise(v[k]);
        bool peak = v[k] >= m;
        if (!peak && (m - v[k] < kCnTieEps || v[k] > 8.0f)) peak = sc == sigmoid_precise(m);
        if (peak) {
          const uint64_t key = cn_key(sc, (uint32_t)((y * W + x) * nc + c));
          if (key <= tau) {
            const int pos = atomicAdd(sh_cnt, 1);
            if (pos < kCnStage) keys[pos] = key;
          }
        }
      }
    }
  }
}

// keep the best K of the staged keys (all threads call); returns the new count
__device__ __forceinline__ int cn_truncate(uint64_t* keys, int n, int K, int* sh_cnt) {
  if (n <= K) return n;
  const int P = pow2_ceil(n < 32 ? 32 : n);
  for (int t = n + threadIdx.x; t < P; t += kCnAThreads) keys[t] = ~0ull;
  __syncthreads();
  block_sort_smem(keys, P);
  if (threadIdx.x == 0) *sh_cnt = K;
  __syncthreads();
  return K;
}

// ---------------------------------------------------------------------------------------------
// pass A, primary: independent warps streaming column tiles (the design of the YOLO decode kernels)
// ---------------------------------------------------------------------------------------------
// The row kernel below synchronises a whole CTA twice per image row and pays several L2 round trips per row
// (bounds, list reservation, histogram): ~4 us per 43 KB row, 27 % of the HBM roofline.  Here a WARP owns a
// tile = tile_cols columns of one image row plus one halo column each side (contiguous in the NHWC tensor ->
// one bulk async copy into the warp's private 2-stage ring), never meets a CTA barrier, and touches global
// memory only when the tile holds a cell above the image's running bound (13 % of the tiles once the bound
// is tight).  Tiles are handed out image-fastest, so every image's bound tightens after its first few tiles.
// There is no per-row cap here: an image that overflows its K*H list (dense maps before a bound exists,
// massive score ties) is flagged and redone exactly by the row kernel.
constexpr int kCtKeyStage = 256;  // staged keys per warp

__device__ __forceinline__ void ct_tile_info(const CnParams& p, int g, int& b, int& y, int& x0, int& x1, int& xlo, int& xhi) {
  // units are handed out image-fastest (every image's bound tightens early), but each image walks its units
  // from a different starting point: the same (row, column) of different images lies a multiple of the image
  // size apart (5 505 024 B = 21 * 2^18 at 128x128x84), and thousands of concurrent requests at such strides
  // pile up on the same HBM channels / banks
  const int g0 = g - p.tile_lo;
  b = g0 % p.B;
  const int per_image = (p.total_tiles - p.tile_lo) / p.B;
  const int k = g0 / p.B;
  const int t = p.tile_lo / p.B + (int)(((unsigned)k + (unsigned)b * 37u) % (unsigned)per_image);
  y = t / p.tiles_per_row;
  x0 = (t - y * p.tiles_per_row) * p.tile_cols;
  x1 = min(p.W, x0 + p.tile_cols);
  xlo = max(x0 - 1, 0);
  xhi = min(x1 + 1, p.W);
}

// all lanes of the warp; n = staged keys (warp-uniform)
__device__ __forceinline__ void ct_flush(const CnParams& p, int b, uint64_t* keys, int* wcnt) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  const int n = min(*wcnt, kCtKeyStage);
  if (n == 0) return;
  uint64_t* dst = p.list + (size_t)b * p.list_cap;
  uint32_t* hist = p.hist + (size_t)b * kCnBins;
  // the list reservation and the read of the histogram's best 128 bins are issued together: ONE L2 round trip
  // for the whole flush (the bound below is raised with a fire-and-forget atomicMax)
  int base = 0;
  if (lane == 0) base = atomicAdd(p.list_count + b * kCnPad, n);
  uint4 h0 = __ldcg(reinterpret_cast<const uint4*>(hist + 4 * lane));
  base = __shfl_sync(0xffffffffu, base, 0);
  for (int t = lane; t < n; t += 32) {
    const uint64_t k = keys[t];
    if (base + t < p.list_cap) dst[base + t] = k;
    atomicAdd(hist + cn_score_bin(k), 1u);
  }
  if (lane == 0 && base + n > p.list_cap) atomicExch(p.redo_any, 1);  // list overflow: the image is redone exactly
  // tighten the image's bound: first score bin (best first) at which the cumulative count of emitted peaks
  // reaches K (counts only grow, so a stale read - it misses this flush's own keys - is conservative)
  int run = 0, edge = -1;
  for (int b0 = 0; b0 < kCnBins && edge < 0; b0 += 32 * 4) {
    const uint4 h = b0 == 0 ? h0 : __ldcg(reinterpret_cast<const uint4*>(hist + b0 + 4 * lane));
    const int mine = (int)(h.x + h.y + h.z + h.w);
    int incl = mine;
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += u;
    }
    const int before = run + incl - mine;
    int e = 0x7fffffff;
    if (before < p.K && before + mine >= p.K) {
      int acc = before;
      const int hv[4] = {(int)h.x, (int)h.y, (int)h.z, (int)h.w};
      for (int q = 0; q < 4; ++q) {
        acc += hv[q];
        if (acc >= p.K) {
          e = b0 + 4 * lane + q;
          break;
        }
      }
    }
    for (int d = 16; d > 0; d >>= 1) e = min(e, __shfl_xor_sync(0xffffffffu, e, d));
    if (e != 0x7fffffff) edge = e;
    run += __shfl_sync(0xffffffffu, incl, 31);
    if (b0 == 0 && run < p.K / 4) break;  // far from K peaks so far: do not walk the whole histogram
  }
  if (lane == 0 && edge >= 0) {
    const float bound = cn_bin_logit_bound(edge);
    if (bound > -INFINITY) atomic_max_float(p.tau_logit + b * kCnPad, bound);
  }
  __syncwarp();
  if (lane == 0) *wcnt = 0;
  __syncwarp();
}

// exact test of one cell above the bound (same arithmetic as the row kernel's slow path), key -> warp stage
__device__ __noinline__ void ct_test_cell(const float* row, int x, int c, float v, int y, int W, int nc, int Cf,
                                          uint64_t* keys, int* wcnt) {
  // 3x3 window maximum over (x-1..x+1, c-1..c+1) with -inf padding == the maximum over the window clamped to
  // the map (a clamped index only repeats a cell that is in the window anyway).  Eight independent loads: one
  // L1/L2 latency instead of eight.
  const int xm = max(x - 1, 0), xp = min(x + 1, W - 1), cm = max(c - 1, 0), cp = min(c + 1, nc - 1);
  const float* r0 = row + xm * Cf;
  const float* r1 = row + x * Cf;
  const float* r2 = row + xp * Cf;
  const float a0 = r0[cm], a1 = r0[c], a2 = r0[cp], b0 = r1[cm], b2 = r1[cp], d0 = r2[cm], d1 = r2[c], d2 = r2[cp];
  const float m = fmaxf(fmaxf(fmaxf(fmaxf(a0, a1), fmaxf(a2, b0)), fmaxf(fmaxf(b2, d0), fmaxf(d1, d2))), v);
  bool peak = v >= m;
  if (!peak && !(m - v < kCnTieEps || v > 8.0f)) return;  // clearly below a neighbour: the common rejection
  const float sc = sigmoid_precise(v);
  if (!peak) peak = sc == sigmoid_precise(m);
  if (!peak) return;
  const int pos = atomicAdd(wcnt, 1);
  if (pos < kCtKeyStage) keys[pos] = cn_key(sc, (uint32_t)((y * W + x) * nc + c));
}

// Streaming scan: no shared-memory staging at all.  A warp owns a unit = tile_cols consecutive columns of one
// image row = a contiguous run of float4 groups in the NHWC tensor, and reads it with plain coalesced 128-bit
// loads, kCtUnroll x 512 B in flight per warp.  (A per-warp shared-memory ring fed by one TMA load per unit - 1-D
// bulk copy or 2-D tensor map, both measured, 12 warps x 3 stages x 5 KB - stays at 2.0-2.6 TB/s: with so few
// warps the per-unit serial chain, not the memory system, sets the pace.)  A float4 holds 4 consecutive classes of one column (or its reg/wh group,
// which is skipped); the FAST PATH only asks "any of the 4 >= the image's running logit bound".  A cell above
// the bound takes the SLOW PATH: its 3x3 (x, class) window is read back through L1/L2 (the neighbouring
// columns are +-336 B away) and tested exactly.
constexpr int kCtThreads = 256;
constexpr int kCtUnroll = 3;   // (x occupancy 4 CTAs/SM at 64 registers; measured at C3: unroll 2 / 3 / 4 / 6 = 98.1 / 87.7 / 90.6 / 109 us; 3 CTAs x unroll 4 = 94.9)

__global__ void __launch_bounds__(kCtThreads, 4) centernet_tiles_kernel(const __grid_constant__ CnParams p) {
  pdl_trigger();
  pdl_wait();  // the sampled bounds / zeroed lists of centernet_sample_bound_kernel
  __shared__ uint64_t sh_keys[kCtThreads / 32][kCtKeyStage];
  __shared__ int sh_wcnt[kCtThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kCtThreads / 32;
  const int W = p.W, nc = p.nc, Cf = p.nc + 4;
  const int cf4 = Cf >> 2, nc4 = nc >> 2;  // float4 groups per column; group nc4 is reg/wh
  const int r32 = 32 % cf4;
  uint64_t* keys = sh_keys[warp];
  int* wcnt = &sh_wcnt[warp];
  if (lane == 0) *wcnt = 0;
  __syncwarp();
  const uint64_t policy = l2_policy_evict_first();

  const int stride = gridDim.x * kWarps;
  const int first = p.tile_lo + blockIdx.x * kWarps + warp;
  const int n_my = first < p.total_tiles ? (p.total_tiles - first + stride - 1) / stride : 0;

  // the image's bound is fetched one unit ahead (its L2 round trip hides behind the current unit); the first
  // units read it directly so that only ONE wave runs without a bound
  int thr_pf = cn_enc(-INFINITY);
  for (int i = 0; i < n_my; ++i) {
    int b, y, x0, x1, xlo, xhi;
    ct_tile_info(p, first + i * stride, b, y, x0, x1, xlo, xhi);
    if (i < 3) thr_pf = ld_cg_s32(p.tau_logit + b * kCnPad);
    const float thr = cn_dec(thr_pf) - kCnTieEps;
    if (i + 1 < n_my) thr_pf = ld_cg_s32(p.tau_logit + ((first + (i + 1) * stride) % p.B) * kCnPad);
    const float* row = p.pred + ((size_t)b * p.H + y) * (size_t)W * Cf;  // column x of the row: row + x * Cf
    const float4* src = reinterpret_cast<const float4*>(row + (size_t)x0 * Cf);
    const int n4 = (x1 - x0) * cf4;
    int mod = lane % cf4;  // (group index within its column) of this lane's next float4
    const float4 kNone = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    float4 cur[kCtUnroll], nxt[kCtUnroll];
#pragma unroll
    for (int u = 0; u < kCtUnroll; ++u) {
      const int idx = u * 32 + lane;
      cur[u] = idx < n4 ? ldg_stream_f4(src + idx, policy) : kNone;
    }
    float thr_b = thr;
    int thr_bpf = cn_enc(-INFINITY);
    for (int it = 0; it < n4; it += 32 * kCtUnroll) {
      // software pipeline: the next batch is in flight while this one is tested; the bound is refreshed the same
      // way every batch (a fresher bound means fewer cells on the slow path and fewer keys to flush)
      const float thr = fmaxf(thr_b, cn_dec(thr_bpf) - kCnTieEps);
      thr_b = thr;
      thr_bpf = ld_cg_s32(p.tau_logit + b * kCnPad);
#pragma unroll
      for (int u = 0; u < kCtUnroll; ++u) {
        const int idx = it + 32 * kCtUnroll + u * 32 + lane;
        nxt[u] = idx < n4 ? ldg_stream_f4(src + idx, policy) : kNone;
      }
#pragma unroll
      for (int u = 0; u < kCtUnroll; ++u) {
        const float4 t = cur[u];
        const bool hit = mod != nc4 && fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)) >= thr;
        if (__any_sync(0xffffffffu, hit)) {  // rare once the bound is tight
          if (hit) {
            const int idx = it + u * 32 + lane;
            const int x = x0 + idx / cf4, c0 = 4 * mod;
            if (t.x >= thr) ct_test_cell(row, x, c0, t.x, y, W, nc, Cf, keys, wcnt);
            if (t.y >= thr) ct_test_cell(row, x, c0 + 1, t.y, y, W, nc, Cf, keys, wcnt);
            if (t.z >= thr) ct_test_cell(row, x, c0 + 2, t.z, y, W, nc, Cf, keys, wcnt);
            if (t.w >= thr) ct_test_cell(row, x, c0 + 3, t.w, y, W, nc, Cf, keys, wcnt);
          }
          __syncwarp();
          if (*wcnt > kCtKeyStage - 128) ct_flush(p, b, keys, wcnt);  // room for one more load of <= 128 cells
        }
        mod += r32;
        if (mod >= cf4) mod -= cf4;
      }
#pragma unroll
      for (int u = 0; u < kCtUnroll; ++u) cur[u] = nxt[u];
    }
    ct_flush(p, b, keys, wcnt);
  }
}

// Cold start of the streaming scan: a GUESS of the image's bound from a 4096-cell sample of the heat logits - the
// logit that about 6 of the samples exceed, i.e. ~0.15 % of the map (~2 000 cells at 128x128x80), far more than K
// peaks.  The guess is only a filter: every peak at or above it is emitted, so the K best emitted peaks are the
// image's K best whenever at least K were emitted, and centernet_redo_mark_kernel sends the image to the exact
// row kernel otherwise.  (An exact bootstrap on the first rows cost 40 us; this costs ~3 us.)
constexpr int kCsSamples = 4096;
constexpr int kCsTop = 6;

__global__ void __launch_bounds__(256) centernet_sample_bound_kernel(const CnParams p) {
  pdl_trigger();  // the tile kernel may be scheduled; it waits for this grid in pdl_wait()
  __shared__ int sh_hist[1024];  // logit bins of 1/16 over [-32, 32)
  __shared__ int sh_coarse[16];  // sums of 64 bins
  const int b = blockIdx.x;
  const int Cf = p.nc + 4;
  const unsigned cells = (unsigned)p.H * p.W * p.nc;  // < 2^32 (checked at launch)
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh_hist[i] = 0;
  // per-image state of the whole decode (replaces three memset / fill launches)
  for (int i = threadIdx.x; i < kCnBins; i += blockDim.x) p.hist[(size_t)b * kCnBins + i] = 0u;
  if (threadIdx.x == 0) {
    p.tau[b * kCnPad] = ~0ull;
    p.list_count[b * kCnPad] = 0;
    p.redo[b] = 0;
    if (b == 0) *p.redo_any = 0;
  }
  __syncthreads();
  unsigned step = cells / kCsSamples;
  if (step < 1) step = 1;
  step |= 1;  // odd stride: walks through all classes
  const float* base = p.pred + (size_t)b * p.H * p.W * Cf;
  float v[kCsSamples / 256];
#pragma unroll
  for (int s = 0; s < kCsSamples / 256; ++s) {  // all loads of a thread in flight at once
    const unsigned idx = (unsigned)(threadIdx.x + 256 * s) * step;
    const unsigned pixel = idx / (unsigned)p.nc;
    v[s] = idx < cells ? __ldg(base + (size_t)pixel * Cf + (idx - pixel * p.nc)) : -INFINITY;
  }
#pragma unroll
  for (int s = 0; s < kCsSamples / 256; ++s) {
    if (v[s] > -INFINITY) {
      int bin = (int)((fminf(fmaxf(v[s], -32.0f), 31.9f) + 32.0f) * 16.0f);
      bin = bin < 0 ? 0 : (bin > 1023 ? 1023 : bin);
      atomicAdd(&sh_hist[bin], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    int t = 0;
    for (int i = 0; i < 64; ++i) t += sh_hist[threadIdx.x * 64 + i];
    sh_coarse[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0, blk = 15;
    for (; blk > 0 && acc + sh_coarse[blk] < kCsTop; --blk) acc += sh_coarse[blk];
    int bin = blk * 64 + 63;
    for (; bin > blk * 64; --bin) {
      acc += sh_hist[bin];
      if (acc >= kCsTop) break;
    }
    // lower edge of that bin; when even all samples are fewer than kCsTop this ends at bin 0 = -32: no bound
    p.tau_logit[b * kCnPad] = cn_enc(bin > 0 ? (float)bin * 0.0625f - 32.0f : -INFINITY);
  }
}

// images whose list overflowed in the tile kernel are reset and flagged for the exact row kernel
__global__ void centernet_redo_mark_kernel(const CnParams p) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  // overflowed list, or fewer than K keys above the guessed bound (and the map has at least K cells at all)
  const int n_keys = p.list_count[b * kCnPad];
  const long long cells = (long long)p.H * p.W * p.nc;
  const bool flag = n_keys > p.list_cap || (n_keys < p.K && cells >= p.K);
  if (flag && threadIdx.x == 0) atomicExch(p.redo_any, 1);
  if (threadIdx.x == 0) p.redo[b] = flag ? 1 : 0;
  if (!flag) return;
  for (int i = threadIdx.x; i < kCnBins; i += blockDim.x) p.hist[(size_t)b * kCnBins + i] = 0u;
  if (threadIdx.x == 0) {
    p.list_count[b * kCnPad] = 0;
    p.tau[b * kCnPad] = ~0ull;
    p.tau_logit[b * kCnPad] = cn_enc(-INFINITY);
  }
}

__global__ void __launch_bounds__(kCnAThreads, 2) centernet_peaks_kernel(const __grid_constant__ CnParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int W = p.W, nc = p.nc, Cf = p.nc + 4;
  const int row_floats = W * Cf;
  const uint32_t row_bytes = (uint32_t)row_floats * 4u;
  float* ring = reinterpret_cast<float*>(smem_raw);                                 // [2][row_floats]
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw + (((size_t)2 * row_bytes + 127) & ~(size_t)127));  // [kCnStage]
  uint64_t* bar = keys + kCnStage;                                                  // [2]
  __shared__ int sh_cnt2[2];  // per ring stage: the counter of row i is never reset while row i is being read
  __shared__ int sh_base;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (p.redo_any && *p.redo_any == 0) return;  // exact redo pass: nothing overflowed (the common case)
  const int rows_total = p.B * (p.row_hi - p.row_lo);
  const int n_my = ((int)blockIdx.x < rows_total) ? (rows_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int i) {  // thread 0 only; rows are handed out image-fastest so that concurrently
    const int r = blockIdx.x + i * gridDim.x;  // running CTAs work on different images
    const int b = r % p.B, y = p.row_lo + r / p.B;
    mbar_arrive_expect_tx(&bar[i & 1], row_bytes);
    bulk_g2s(ring + (size_t)(i & 1) * row_floats, p.pred + ((size_t)b * p.H + y) * row_floats, row_bytes, &bar[i & 1]);
  };
  if (tid == 0) {
    if (n_my > 0) issue(0);
    if (n_my > 1) issue(1);
  }

  for (int i = 0; i < n_my; ++i) {
    const int r = blockIdx.x + i * gridDim.x;
    const int b = r % p.B, y = p.row_lo + r / p.B;
    const float* row = ring + (size_t)(i & 1) * row_floats;
    int& sh_cnt = sh_cnt2[i & 1];
    if (tid == 0) sh_cnt = 0;
    unsigned long long tau = *reinterpret_cast<volatile unsigned long long*>(p.tau + b * kCnPad);
    const float thr = cn_dec(*reinterpret_cast<volatile int*>(p.tau_logit + b * kCnPad)) - kCnTieEps;
    mbar_wait(&bar[i & 1], (uint32_t)(i >> 1) & 1u);
    __syncthreads();

    if (!p.redo || p.redo[b]) cn_scan_columns(row, 0, W, y, W, nc, Cf, thr, tau, keys, &sh_cnt);  // uniform
    __syncthreads();
    int n_r = sh_cnt;
    if (n_r > kCnStage) {
      // more peaks above the bound than the stage holds (no bound yet on a dense map): redo the row in column
      // chunks that cannot overflow, cutting to the best K after each
      __syncthreads();
      if (tid == 0) sh_cnt = 0;
      __syncthreads();
      int xc = (kCnStage - p.K) / nc;
      if (xc < 1) xc = 1;  // (launch checks K + nc <= kCnStage)
      n_r = 0;
      for (int xb = 0; xb < W; xb += xc) {
        cn_scan_columns(row, xb, min(W, xb + xc), y, W, nc, Cf, thr, tau, keys, &sh_cnt);
        __syncthreads();
        n_r = cn_truncate(keys, sh_cnt, p.K, &sh_cnt);
        if (n_r == p.K) tau = min(tau, (unsigned long long)keys[p.K - 1]);
      }
    }
    if (tid == 0 && i + 2 < n_my) issue(i + 2);  // every read of the row is behind a barrier: refill the stage
    const bool cut = n_r > p.K;
    const int n_emit = cn_truncate(keys, n_r, p.K, &sh_cnt);
    if (tid == 0) {
      if (cut || n_emit == p.K) {
        atomicMin(p.tau + b * kCnPad, (unsigned long long)keys[p.K - 1]);
        // the same bound in the logit domain: score of the row's K-th best peak
        const float s = __uint_as_float(0x7fffffffu - (uint32_t)(keys[p.K - 1] >> 32));
        if (s > 0.0f && s < 1.0f) atomic_max_float(p.tau_logit + b * kCnPad, logf(s / (1.0f - s)) - 4.0f * kCnTieEps);
      }
      if (n_emit > 0) sh_base = atomicAdd(p.list_count + b * kCnPad, n_emit);
    }
    if (n_emit == 0) continue;  // (uniform) nothing above the bound in this row: the common case
    __syncthreads();
    const int base = sh_base;
    uint64_t* dst = p.list + (size_t)b * p.list_cap;
    uint32_t* hist = p.hist + (size_t)b * kCnBins;
    for (int t = tid; t < n_emit; t += kCnAThreads) {
      const uint64_t k = keys[t];
      if (base + t < p.list_cap) dst[base + t] = k;
      atomicAdd(hist + cn_score_bin(k), 1u);
    }
    // tighten the image's bound from the histogram of everything emitted so far (any row, any CTA): the
    // first bin (best scores first) at which the cumulative count reaches K
    if (n_emit >= 4 && warp == 0) {
      __threadfence();
      int run = 0, edge = -1;
      for (int b0 = 0; b0 < kCnBins && edge < 0; b0 += 32 * 4) {
        const uint4 h = __ldcg(reinterpret_cast<const uint4*>(hist + b0 + 4 * lane));  // L2, never a stale L1 line
        const int mine = (int)(h.x + h.y + h.z + h.w);
        int incl = mine;
        for (int d = 1; d < 32; d <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += u;
        }
        const int before = run + incl - mine;
        int e = 0x7fffffff;
        if (before < p.K && before + mine >= p.K) {  // the crossing is inside this lane's 4 bins
          int acc = before;
          const int hv[4] = {(int)h.x, (int)h.y, (int)h.z, (int)h.w};
          for (int q = 0; q < 4; ++q) {
            acc += hv[q];
            if (acc >= p.K) {
              e = b0 + 4 * lane + q;
              break;
            }
          }
        }
        for (int d = 16; d > 0; d >>= 1) e = min(e, __shfl_xor_sync(0xffffffffu, e, d));
        if (e != 0x7fffffff) edge = e;
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0 && edge >= 0) {
        const float bound = cn_bin_logit_bound(edge);
        if (bound > -INFINITY) atomic_max_float(p.tau_logit + b * kCnPad, bound);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// box_diou (iou.py:41-64) of two xyxy boxes, one fp32 rounding per reference op
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float diou_exact(const float4& a, const float4& b) {
  const float eps = 1e-6f;
  const float area1 = fmul(fsub(a.z, a.x), fsub(a.w, a.y));
  const float area2 = fmul(fsub(b.z, b.x), fsub(b.w, b.y));
  const float iw = fmaxf(fsub(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
  const float ih = fmaxf(fsub(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
  const float inter = fmul(iw, ih);
  const float uni = fsub(fadd(area1, area2), inter);
  const float iou = fdiv(inter, fmaxf(uni, eps));
  const float c1x = fmul(fadd(a.x, a.z), 0.5f), c1y = fmul(fadd(a.y, a.w), 0.5f);
  const float c2x = fmul(fadd(b.x, b.z), 0.5f), c2y = fmul(fadd(b.y, b.w), 0.5f);
  const float ew = fmaxf(fsub(fmaxf(a.z, b.z), fminf(a.x, b.x)), 0.0f);
  const float eh = fmaxf(fsub(fmaxf(a.w, b.w), fminf(a.y, b.y)), 0.0f);
  const float c_sq = fadd(fmul(ew, ew), fmul(eh, eh));
  const float dx = fsub(c1x, c2x), dy = fsub(c1y, c2y);
  const float d_sq = fadd(fmul(dx, dx), fmul(dy, dy));
  return fsub(iou, fdiv(d_sq, fmaxf(c_sq, eps)));
}

// greedy DIoU-NMS over n (<= blockDim.x) score-ordered boxes in shared memory; alive[] in/out
__device__ __forceinline__ void diou_greedy_block(const float4* box, int n, float thr, unsigned char* alive) {
  const int j = threadIdx.x;
  float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < n) mine = box[j];
  for (int i = 0; i < n; ++i) {
    __syncthreads();
    if (alive[i] && j > i && j < n && alive[j]) {
      if (!(diou_exact(box[i], mine) <= thr)) alive[j] = 0;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// pass B
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCnThreads, 1) centernet_finalize_kernel(const __grid_constant__ CnParams p) {
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);  // [key_cap]
  __shared__ float4 sbox[kCnThreads];
  __shared__ float sscore[kCnThreads];
  __shared__ int scls[kCnThreads];
  __shared__ int spix[kCnThreads];
  __shared__ unsigned char alive[kCnThreads];
  __shared__ int sh_cnt;
  __shared__ int sh_warp[kCnThreads / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int W = p.W, H = p.H, nc = p.nc, Cf = p.nc + 4;
  int m = p.list_count[b * kCnPad];
  if (m > p.list_cap) m = p.list_cap;
  const unsigned long long tau = p.tau[b * kCnPad];
  const uint64_t* src = p.list + (size_t)b * p.list_cap;

  // keep the keys that can still be among the K best, compacted into the sort buffer.  tau alone is loose (it is only
  // tightened by rows that hold K peaks): the image's histogram counts EVERY emitted key (best score bin first), so the
  // first bin at which the cumulative count reaches K bounds the K best exactly - a few hundred keys are sorted instead
  // of the whole list (the sort was most of this kernel's 16 us).
  static_assert(kCnThreads >= kCnBins, "one histogram bin per thread");
  __shared__ int sh_bin_cut;
  {
    const int mine = tid < kCnBins ? (int)__ldcg(p.hist + (size_t)b * kCnBins + tid) : 0;
    int incl = mine;
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += u;
    }
    if (lane == 31) sh_warp[warp] = incl;
    if (tid == 0) {
      sh_bin_cut = kCnBins - 1;  // fewer than K keys in all: take everything
      sh_cnt = 0;
    }
    __syncthreads();
    for (int q = 0; q < warp; ++q) incl += sh_warp[q];
    if (incl >= p.K && incl - mine < p.K) sh_bin_cut = tid;  // one thread at most
    __syncthreads();
  }
  const int bin_cut = sh_bin_cut;
  uint64_t* buf = (p.sort_cap <= p.key_cap) ? keys : p.ws_sort + (size_t)b * p.sort_cap;
  for (int t0 = 0; t0 < m; t0 += kCnThreads) {
    const int t = t0 + tid;
    uint64_t k = 0;
    bool take = false;
    if (t < m) {
      k = src[t];
      take = k <= tau && cn_score_bin(k) <= bin_cut;
    }
    const unsigned mk = __ballot_sync(0xffffffffu, take);
    if (mk) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&sh_cnt, __popc(mk));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (take) buf[base + __popc(mk & ((1u << lane) - 1u))] = k;
    }
  }
  __syncthreads();
  const int m2 = sh_cnt;
  const int P = pow2_ceil(m2 < 32 ? 32 : m2);
  for (int t = m2 + tid; t < P; t += kCnThreads) buf[t] = ~0ull;
  __syncthreads();
  if (buf == keys)
    block_sort_smem(buf, P);
  else
    bitonic_sort_u64_generic(buf, P);

  // top K: decode, gather reg / wh, assemble (centernet.py:282-304)
  const int n_top = m2 < p.K ? m2 : p.K;
  bool pass = false;
  if (tid < n_top) {
    const uint64_t k = buf[tid];
    const uint32_t flat = (uint32_t)k;
    const float score = __uint_as_float(0x7fffffffu - (uint32_t)(k >> 32));
    const int c = (int)(flat % (uint32_t)nc);
    const int pixel = (int)(flat / (uint32_t)nc);
    const int y = pixel / W, x = pixel - y * W;
    const float* pp = p.pred + ((size_t)b * H * W + pixel) * Cf;
    float cx = fadd((float)x, pp[nc]), cy = fadd((float)y, pp[nc + 1]);
    float w = pp[Cf - 2], h = pp[Cf - 1];
    cx = fdiv(cx, (float)W);
    w = fdiv(w, (float)W);
    cy = fdiv(cy, (float)H);
    h = fdiv(h, (float)H);
    cx = fminf(fmaxf(cx, 0.0f), 1.0f);
    cy = fminf(fmaxf(cy, 0.0f), 1.0f);
    w = fminf(fmaxf(w, 0.0f), 1.0f);
    h = fminf(fmaxf(h, 0.0f), 1.0f);
    const float hw = fmul(w, 0.5f), hh = fmul(h, 0.5f);
    sbox[tid] = make_float4(fsub(cx, hw), fsub(cy, hh), fadd(cx, hw), fadd(cy, hh));
    sscore[tid] = score;
    scls[tid] = c;
    spix[tid] = pixel;
    pass = score >= p.conf;  // scores are descending: the survivors are a prefix
  }
  const unsigned pm = __ballot_sync(0xffffffffu, pass);
  if (lane == 0) sh_warp[warp] = __popc(pm);
  __syncthreads();
  int n_conf = 0;
  for (int q = 0; q < kCnThreads / 32; ++q) n_conf += sh_warp[q];
  alive[tid] = tid < n_conf ? 1 : 0;
  __syncthreads();
  if (p.use_nms && n_conf > 1) diou_greedy_block(sbox, n_conf, p.nms_thr, alive);

  // ordered compaction + letterbox inverse (xywh=False branch, image_process.py:112-129)
  const bool keep = tid < n_conf && alive[tid];
  const unsigned km = __ballot_sync(0xffffffffu, keep);
  __syncthreads();
  if (lane == 0) sh_warp[warp] = __popc(km);
  __syncthreads();
  int off = 0, total = 0;
  for (int q = 0; q < kCnThreads / 32; ++q) {
    if (q < warp) off += sh_warp[q];
    total += sh_warp[q];
  }
  if (keep) {
    const int o = off + __popc(km & ((1u << lane) - 1u));
    float4 bx = sbox[tid];
    if (p.letterbox) {
      const float* L = p.letterbox + 5 * b;
      bx.x = fmul(fsub(fmul(bx.x, L[0]), L[2]), L[4]);
      bx.z = fmul(fsub(fmul(bx.z, L[0]), L[2]), L[4]);
      bx.y = fmul(fsub(fmul(bx.y, L[1]), L[3]), L[4]);
      bx.w = fmul(fsub(fmul(bx.w, L[1]), L[3]), L[4]);
    }
    const size_t at = (size_t)b * p.K + o;
    p.det_box[at] = bx;
    p.det_score[at] = sscore[tid];
    p.det_cls[at] = scls[tid];
    p.det_pixel[at] = spix[tid];
  }
  if (tid == 0) p.det_count[b] = total;
}

// ---------------------------------------------------------------------------------------------
// standalone diou_nms(boxes, scores, thr) (core/utils/nms.py:9-31), one CTA
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCnThreads, 1)
diou_nms_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, int n, float thr,
                long long* __restrict__ keep, int32_t* __restrict__ keep_count, int P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);                     // [P]
  unsigned char* alive = reinterpret_cast<unsigned char*>(keys + P);          // [P]
  __shared__ int sh_warp[kCnThreads / 32];
  __shared__ int sh_running;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int t = tid; t < P; t += kCnThreads) {
    keys[t] = t < n ? cn_key(scores[t], (uint32_t)t) : ~0ull;
    alive[t] = t < n ? 1 : 0;
  }
  __syncthreads();
  block_sort_smem(keys, P);
  // greedy in sorted order; alive[] is indexed by sorted position
  for (int i = 0; i < n; ++i) {
    __syncthreads();
    if (!alive[i]) continue;
    const float4 bi = boxes[(uint32_t)keys[i]];
    for (int j = i + 1 + tid; j < n; j += kCnThreads)
      if (alive[j] && !(diou_exact(bi, boxes[(uint32_t)keys[j]]) <= thr)) alive[j] = 0;
  }
  __syncthreads();
  if (tid == 0) sh_running = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kCnThreads) {
    const int t = base + tid;
    const bool k = t < n && alive[t];
    const unsigned km = __ballot_sync(0xffffffffu, k);
    if (lane == 0) sh_warp[warp] = __popc(km);
    __syncthreads();
    int off = 0, total = 0;
    for (int q = 0; q < kCnThreads / 32; ++q) {
      if (q < warp) off += sh_warp[q];
      total += sh_warp[q];
    }
    if (k) keep[sh_running + off + __popc(km & ((1u << lane) - 1u))] = (long long)(uint32_t)keys[t];
    __syncthreads();
    if (tid == 0) sh_running += total;
    __syncthreads();
  }
  if (tid == 0) *keep_count = sh_running;
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
__global__ void cn_fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
static cudaError_t cudaMemsetD32Async_compat(void* p, uint32_t v, size_t n, cudaStream_t stream) {
  cn_fill_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<uint32_t*>(p), v, n);
  return cudaGetLastError();
}

static int cn_sort_cap(int H, int K) {
  int P = 32;
  while (P < H * K) P <<= 1;
  return P;
}

size_t centernet_workspace_bytes(int B, int H, int W, int nc, int K) {
  (void)W;
  (void)nc;
  size_t s = 256;
  s += ((size_t)B * 32 * 8 + 255) & ~(size_t)255;                     // tau (padded, kCnPad)
  s += ((size_t)B * 32 * 4 + 255) & ~(size_t)255;                     // list_count (padded)
  s += ((size_t)(B + 1) * 4 + 255) & ~(size_t)255;                    // redo flags + redo_any
  s += ((size_t)B * 32 * 4 + 255) & ~(size_t)255;                     // tau_logit (padded)
  s += ((size_t)B * 1024 * 4 + 255) & ~(size_t)255;                   // hist
  s += ((size_t)B * (size_t)H * K * 8 + 255) & ~(size_t)255;          // list
  s += ((size_t)B * (size_t)cn_sort_cap(H, K) * 8 + 255) & ~(size_t)255;  // sort scratch
  return s;
}

int centernet_launch(const float* pred, int B, int H, int W, int nc, int K, float conf, int pool_mode, int use_nms,
                     float nms_thr, const float* letterbox, float* det_box, float* det_score, int32_t* det_cls,
                     int32_t* det_pixel, int32_t* det_count, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream) {
  if (!pred || !det_box || !det_score || !det_cls || !det_pixel || !det_count) {
    set_error("centernet: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || H < 1 || W < 1 || nc < 1 || K < 1 || (int64_t)H * W * nc >= (1ll << 32)) {
    set_error("centernet: bad sizes (B=%d H=%d W=%d nc=%d K=%d)", B, H, W, nc, K);
    return CVPP_ERR_INVALID_ARG;
  }
  if (pool_mode != 0) {
    set_error("centernet: pool_mode %d is not compiled in (0 = the reference's (x, class) window)", pool_mode);
    return CVPP_ERR_UNSUPPORTED;
  }
  if (K > kCnThreads) {
    set_error("centernet: K=%d exceeds %d", K, kCnThreads);
    return CVPP_ERR_UNSUPPORTED;
  }
  const size_t row_bytes = (size_t)W * (nc + 4) * 4;
  if ((reinterpret_cast<uintptr_t>(pred) & 15u) || (row_bytes & 15u) || (reinterpret_cast<uintptr_t>(det_box) & 15u)) {
    set_error("centernet: pred rows and det_box must be 16-byte aligned (W*(nc+4) %% 4 == 0)");
    return CVPP_ERR_ALIGNMENT;
  }
  if (B == 0) return CVPP_OK;
  if (!workspace || workspace_bytes < centernet_workspace_bytes(B, H, W, nc, K)) {
    set_error("centernet: workspace of %zu bytes needed, got %zu", centernet_workspace_bytes(B, H, W, nc, K),
              workspace_bytes);
    return CVPP_ERR_WORKSPACE;
  }
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != CVPP_OK) return rc;

  CnParams p{};
  p.pred = pred;
  p.B = B;
  p.H = H;
  p.W = W;
  p.nc = nc;
  p.K = K;
  p.conf = conf;
  p.use_nms = use_nms;
  p.nms_thr = nms_thr;
  p.letterbox = letterbox;
  p.det_box = reinterpret_cast<float4*>(det_box);
  p.det_score = det_score;
  p.det_cls = det_cls;
  p.det_pixel = det_pixel;
  p.det_count = det_count;
  p.list_cap = H * K;
  p.sort_cap = cn_sort_cap(H, K);
  uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  p.tau = reinterpret_cast<unsigned long long*>(w);
  w += ((size_t)B * kCnPad * 8 + 255) & ~(size_t)255;
  p.list_count = reinterpret_cast<int32_t*>(w);
  w += ((size_t)B * kCnPad * 4 + 255) & ~(size_t)255;
  p.redo = reinterpret_cast<int32_t*>(w);
  p.redo_any = p.redo + B;
  w += ((size_t)(B + 1) * 4 + 255) & ~(size_t)255;
  p.tau_logit = reinterpret_cast<int*>(w);
  w += ((size_t)B * kCnPad * 4 + 255) & ~(size_t)255;
  p.hist = reinterpret_cast<uint32_t*>(w);
  w += ((size_t)B * 1024 * 4 + 255) & ~(size_t)255;
  p.list = reinterpret_cast<uint64_t*>(w);
  w += ((size_t)B * (size_t)H * K * 8 + 255) & ~(size_t)255;
  p.ws_sort = reinterpret_cast<uint64_t*>(w);

  const bool tile_path = ((nc + 4) & 3) == 0;
  if (!tile_path) {  // (the tile path initialises everything in centernet_sample_bound_kernel)
    CVPP_CUDA_TRY(cudaMemsetAsync(p.tau, 0xff, sizeof(unsigned long long) * (size_t)B * kCnPad, stream));
    // list_count, tau_logit (-inf in the ordered-int encoding is written below), hist: one contiguous region
    CVPP_CUDA_TRY(cudaMemsetAsync(p.list_count, 0, reinterpret_cast<uintptr_t>(p.list) - reinterpret_cast<uintptr_t>(p.list_count), stream));
    CVPP_CUDA_TRY(cudaMemsetD32Async_compat(p.tau_logit, 0xff800000u ^ 0x7fffffffu, (size_t)B * kCnPad, stream));
  }

  // ---- pass A, primary: independent warps over column tiles (needs 16-byte aligned columns: (nc + 4) % 4 == 0)
  bool tiles_done = false;
  if (tile_path) {
    const int cf = nc + 4;
    // unit = a quarter row or ~10 KB, whichever is smaller (enough units for the image-fastest hand-out)
    int tc = (W + 3) / 4;
    const int tc_cap = (int)(12288 / ((size_t)cf * 4));
    if (tc > tc_cap) tc = tc_cap;
    if (tc < 1) tc = 1;
    p.tile_cols = tc;
    p.tiles_per_row = (W + tc - 1) / tc;
    p.total_tiles = B * H * p.tiles_per_row;
    p.tile_stage_floats = 0;
    int ctas_per_sm = 1;
    CVPP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, centernet_tiles_kernel, kCtThreads, 0));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    constexpr int kW = kCtThreads / 32;
    // Cold start: without a bound every cell takes the slow path and every peak is emitted; the sampled guess
    // (centernet_sample_bound_kernel) gives every image a filter before the first unit is read.
    centernet_sample_bound_kernel<<<B, 256, 0, stream>>>(p);
    CVPP_CUDA_TRY(cudaGetLastError());
    {
      CnParams pb2 = p;
      pb2.tile_lo = 0;
      const int warps_needed = (p.total_tiles + kW - 1) / kW;
      const int grid_t = warps_needed < ctas_per_sm * di.sms ? warps_needed : ctas_per_sm * di.sms;
      CVPP_CUDA_TRY(launch_pdl(centernet_tiles_kernel, dim3(grid_t), dim3(kCtThreads), 0, stream, pb2));
    }
    // flags (and resets) the images whose list overflowed
    CVPP_CUDA_TRY(launch_pdl(centernet_redo_mark_kernel, dim3(B), dim3(256), 0, stream, p));
    tiles_done = true;
  }
  CnParams pr = p;  // the exact row kernel: everything when the tile kernel could not run, else only flagged images
  pr.row_lo = 0;
  pr.row_hi = H;
  if (!tiles_done) {
    pr.redo = nullptr;
    pr.redo_any = nullptr;
  }

  // pass A, exact (per-row cap K): 2 row stages + the key stage + 2 barriers; two CTAs per SM when they fit
  const size_t smem_a = ((2 * row_bytes + 127) & ~(size_t)127) + (size_t)kCnStage * 8 + 16;
  if (smem_a > (size_t)di.max_smem || K + nc > kCnStage) {
    set_error("centernet: a row of %d x %d cells (K=%d) does not fit the shared-memory pipeline", W, nc + 4, K);
    return CVPP_ERR_UNSUPPORTED;
  }
  p.key_cap = kCnStage;
  pr.key_cap = kCnStage;
  static unsigned long long done_a = 0;
  static int bytes_a = 0;
  if ((int)smem_a > bytes_a) {
    done_a = 0;
    bytes_a = (int)smem_a;
  }
  rc = ensure_smem_attr(reinterpret_cast<const void*>(centernet_peaks_kernel), bytes_a, di.device, &done_a);
  if (rc != CVPP_OK) return rc;
  const int rows = B * H;
  const int ctas_per_sm = (2 * (smem_a + 1024) <= (size_t)di.max_smem + 1024) ? 2 : 1;
  const int grid_a = rows < ctas_per_sm * di.sms ? rows : ctas_per_sm * di.sms;
  CVPP_CUDA_TRY(launch_pdl(centernet_peaks_kernel, dim3(grid_a), dim3(kCnAThreads), smem_a, stream, pr));

  // pass B shared memory: sort buffer when it fits (<= 16384 keys), else the global scratch rows
  CnParams pb = p;
  pb.key_cap = p.sort_cap <= 16384 ? p.sort_cap : 32;
  const size_t smem_b = (size_t)pb.key_cap * 8;
  static unsigned long long done_b = 0;
  static int bytes_b = 0;
  if ((int)smem_b > bytes_b) {
    done_b = 0;
    bytes_b = (int)smem_b;
  }
  rc = ensure_smem_attr(reinterpret_cast<const void*>(centernet_finalize_kernel), bytes_b, di.device, &done_b);
  if (rc != CVPP_OK) return rc;
  CVPP_CUDA_TRY(launch_pdl(centernet_finalize_kernel, dim3(B), dim3(kCnThreads), smem_b, stream, pb));
  return CVPP_OK;
}

int diou_nms_launch(const float* boxes, const float* scores, int n, float thr, long long* keep, int32_t* keep_count,
                    cudaStream_t stream) {
  if (!keep_count || n < 0 || (n > 0 && (!boxes || !scores || !keep))) {
    set_error("diou_nms: NULL pointer or negative n");
    return CVPP_ERR_INVALID_ARG;
  }
  if (n == 0) {
    CVPP_CUDA_TRY(cudaMemsetAsync(keep_count, 0, sizeof(int32_t), stream));
    return CVPP_OK;
  }
  if (reinterpret_cast<uintptr_t>(boxes) & 15u) {
    set_error("diou_nms: boxes must be 16-byte aligned");
    return CVPP_ERR_ALIGNMENT;
  }
  int P = 32;
  while (P < n) P <<= 1;
  if (P > 16384) {
    set_error("diou_nms: n=%d exceeds the 16384-box single-CTA limit", n);
    return CVPP_ERR_UNSUPPORTED;
  }
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != CVPP_OK) return rc;
  const size_t smem = (size_t)P * 9;
  static unsigned long long done = 0;
  static int bytes = 0;
  if ((int)smem > bytes) {
    done = 0;
    bytes = (int)smem;
  }
  rc = ensure_smem_attr(reinterpret_cast<const void*>(diou_nms_kernel), bytes, di.device, &done);
  if (rc != CVPP_OK) return rc;
  diou_nms_kernel<<<1, kCnThreads, smem, stream>>>(reinterpret_cast<const float4*>(boxes), scores, n, thr, keep,
                                                 keep_count, P);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

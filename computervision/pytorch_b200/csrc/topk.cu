// topk.cu — standalone per-image top-K over a flat score row (sm_100a).
//
// Replaces torch.topk inside CenterNetA._top_k (reference core/algorithms/centernet.py:328-338):
//     topk_scores, topk_inds = torch.topk(scores.view(B, -1), K, largest=True, sorted=True)
// torch.topk leaves the order of equal scores unspecified; this kernel pins it to the product's rule
// (the one the fused cvpp_centernet_decode uses): equal scores are ordered by the LOWER flat index.
//
// Exact radix select on the 64-bit key [~ordered(score) : 32 | flat index : 32] (ascending key = score
// descending, lower index first; NaN sorts first like torch.topk).  A level = one streaming histogram pass
// over the row restricted to the current boundary prefix (12-bit digits, shared-memory bins flushed with
// one global atomic per non-empty bin) + a one-CTA-per-image scan that picks the digit where the cumulative
// count reaches K.  The descent stops as soon as the boundary group fits a shared-memory sort (<= 4096
// keys) - ONE level for continuous scores; tie-heavy rows (e.g. the exact zeros a suppressed heat map is
// full of) descend through the score digits into the index digits, which always terminates because flat
// indices are unique.  A collect pass then appends every key better than the boundary group plus the
// boundary group itself to per-image lists, and one CTA per image sorts <= K + 4096 keys and writes the K
// best in order.  Images that are already resolved skip the remaining levels (their CTAs exit at once), so
// the level launches need no host round trip.  HBM-bound: (levels + 1) reads of the row.
#include "cvpp_common.cuh"

namespace cvpp {

constexpr int kTkThreads = 512;
constexpr int kTkBins = 4096;
constexpr int kTkSortCap = 4096;  // boundary groups up to this size are finished by the final sort
constexpr int kTkMaxK = 4096;
constexpr int kTkMaxLevels = 8;

struct TkState {  // one per image
  unsigned long long prefix;  // key >> shift of the boundary group
  int shift;                  // 64 = nothing resolved yet
  int k_rem;                  // keys still to take from the boundary group
  int n_in;                   // keys strictly better than the boundary group
  int boundary;               // size of the boundary group
  int done;                   // boundary <= kTkSortCap
  int in_fill, bd_fill;       // list cursors of the collect pass
  int pad[6];
};
static_assert(sizeof(TkState) == 64, "TkState is one 64-byte record per image");

struct TkParams {
  const float* scores;
  long long N;
  int B, K;
  TkState* state;
  uint32_t* hist;    // [B][kTkBins]
  uint64_t* in_list; // [B][K]
  uint64_t* bd_list; // [B][kTkSortCap]
  int shift, bins;   // this level's digit
  float* out_val;
  long long* out_idx;
  int C, W;          // optional index split (C > 0): cls = idx % C, pixel = idx / C, y = pixel / W, x = pixel % W
  long long *out_cls, *out_y, *out_x;
  int32_t* out_pixel;
};

__device__ __forceinline__ uint64_t tk_key(float v, uint32_t idx) {
  uint32_t b = __float_as_uint(v);
  if (b == 0x80000000u) b = 0u;  // -0.0 == +0.0 (torch compares values)
  const uint32_t ordered = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // ascending unsigned == ascending float
  return ((uint64_t)(~ordered) << 32) | (uint64_t)idx;
}

// visits every element of image b assigned to this CTA: f(key)
template <typename F>
__device__ __forceinline__ void tk_for_each(const TkParams& p, int b, F f) {
  const float* row = p.scores + (long long)b * p.N;
  const long long n4 = ((reinterpret_cast<uintptr_t>(row) & 15u) == 0) ? (p.N >> 2) : 0;
  const float4* row4 = reinterpret_cast<const float4*>(row);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(row4 + i);
    const uint32_t base = (uint32_t)(i << 2);
    f(tk_key(v.x, base));
    f(tk_key(v.y, base + 1));
    f(tk_key(v.z, base + 2));
    f(tk_key(v.w, base + 3));
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.N; i += stride)
    f(tk_key(row[i], (uint32_t)i));
}

__global__ void topk_init_kernel(TkParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  TkState s{};
  s.shift = 64;
  s.k_rem = p.K;
  s.boundary = (int)min(p.N, (long long)0x7fffffff);
  p.state[b] = s;
}

__global__ void __launch_bounds__(kTkThreads) topk_hist_kernel(const __grid_constant__ TkParams p) {
  __shared__ uint32_t sh_hist[kTkBins];
  const int b = blockIdx.y;
  const TkState st = p.state[b];
  if (st.done) return;
  for (int i = threadIdx.x; i < p.bins; i += kTkThreads) sh_hist[i] = 0u;
  __syncthreads();
  const uint32_t mask = (uint32_t)p.bins - 1u;
  // run-length accumulation per thread: a row of equal scores costs one shared atomic per thread, not per element
  int last = -1;
  uint32_t run = 0;
  tk_for_each(p, b, [&](uint64_t key) {
    if (st.shift < 64 && (key >> st.shift) != st.prefix) return;
    const int d = (int)((uint32_t)(key >> p.shift) & mask);
    if (d != last) {
      if (run) atomicAdd(&sh_hist[last], run);
      last = d;
      run = 0;
    }
    ++run;
  });
  if (run) atomicAdd(&sh_hist[last], run);
  __syncthreads();
  uint32_t* gh = p.hist + (size_t)b * kTkBins;
  for (int i = threadIdx.x; i < p.bins; i += kTkThreads) {
    const uint32_t v = sh_hist[i];
    if (v) atomicAdd(&gh[i], v);
  }
}

// one CTA per image: the digit where the cumulative count reaches k_rem becomes the new boundary group
__global__ void __launch_bounds__(1024) topk_select_kernel(const __grid_constant__ TkParams p) {
  __shared__ int sh_warp[32];
  __shared__ int sh_digit, sh_before, sh_count;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  TkState st = p.state[b];
  if (st.done) return;
  uint32_t* gh = p.hist + (size_t)b * kTkBins;
  constexpr int kPer = kTkBins / 1024;
  int local[kPer], lsum = 0;
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int i = tid * kPer + q;
    local[q] = i < p.bins ? (int)gh[i] : 0;
    if (i < p.bins) gh[i] = 0u;  // ready for the next level
    lsum += local[q];
  }
  int incl = lsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += u;
  }
  if (lane == 31) sh_warp[warp] = incl;
  __syncthreads();
  int run = incl - lsum;
  for (int q = 0; q < warp; ++q) run += sh_warp[q];
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    if (run < st.k_rem && run + local[q] >= st.k_rem) {  // exactly one (thread, bin) matches
      sh_digit = tid * kPer + q;
      sh_before = run;
      sh_count = local[q];
    }
    run += local[q];
  }
  __syncthreads();
  if (tid == 0) {
    const int up = st.shift >= 64 ? 0 : st.shift - p.shift;  // bits between the old and the new boundary prefix
    st.prefix = (st.shift >= 64 ? 0ull : (st.prefix << up)) | (unsigned long long)sh_digit;
    st.shift = p.shift;
    st.n_in += sh_before;
    st.k_rem -= sh_before;
    st.boundary = sh_count;
    st.done = (sh_count <= kTkSortCap || p.shift == 0) ? 1 : 0;
    p.state[b] = st;
  }
}

__global__ void __launch_bounds__(kTkThreads) topk_collect_kernel(const __grid_constant__ TkParams p) {
  const int b = blockIdx.y;
  TkState* sp = p.state + b;
  const unsigned long long prefix = sp->prefix;
  const int shift = sp->shift;
  uint64_t* in_list = p.in_list + (size_t)b * p.K;
  uint64_t* bd_list = p.bd_list + (size_t)b * kTkSortCap;
  tk_for_each(p, b, [&](uint64_t key) {
    const unsigned long long hi = key >> shift;
    if (hi < prefix) {
      const int at = atomicAdd(&sp->in_fill, 1);
      if (at < p.K) in_list[at] = key;
    } else if (hi == prefix) {
      const int at = atomicAdd(&sp->bd_fill, 1);
      if (at < kTkSortCap) bd_list[at] = key;
    }
  });
}

__global__ void __launch_bounds__(1024, 1) topk_final_kernel(const __grid_constant__ TkParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* a = reinterpret_cast<uint64_t*>(smem_raw);
  const int b = blockIdx.x, tid = threadIdx.x;
  const TkState st = p.state[b];
  const int n_in = min(st.n_in, p.K), n_bd = min(st.boundary, kTkSortCap);
  const int n = n_in + n_bd;
  const int P = pow2_ceil(n < 32 ? 32 : n);
  const uint64_t* in_list = p.in_list + (size_t)b * p.K;
  const uint64_t* bd_list = p.bd_list + (size_t)b * kTkSortCap;
  for (int i = tid; i < P; i += 1024) a[i] = i < n_in ? in_list[i] : (i < n ? bd_list[i - n_in] : ~0ull);
  __syncthreads();
  block_sort_smem(a, P);
  for (int k = tid; k < p.K; k += 1024) {
    const uint64_t key = a[k];
    const long long idx = (long long)(uint32_t)key;
    const size_t o = (size_t)b * p.K + k;
    p.out_val[o] = p.scores[(long long)b * p.N + idx];
    p.out_idx[o] = idx;
    if (p.C > 0) {
      const long long pixel = idx / p.C;
      if (p.out_cls) p.out_cls[o] = idx % p.C;
      if (p.out_y) p.out_y[o] = pixel / p.W;
      if (p.out_x) p.out_x[o] = pixel % p.W;
      if (p.out_pixel) p.out_pixel[o] = (int32_t)pixel;
    }
  }
}

static size_t tk_align(size_t x) { return (x + 255) & ~(size_t)255; }

size_t topk_workspace_bytes(int B, int K) {
  return 256 + tk_align((size_t)B * sizeof(TkState)) + tk_align((size_t)B * kTkBins * sizeof(uint32_t)) +
         tk_align((size_t)B * (size_t)K * sizeof(uint64_t)) + tk_align((size_t)B * kTkSortCap * sizeof(uint64_t));
}

int topk_launch(const float* scores, int B, long long N, int K, int C, int W, float* out_val, long long* out_idx,
                long long* out_cls, long long* out_y, long long* out_x, int32_t* out_pixel, void* workspace,
                size_t workspace_bytes, cudaStream_t stream) {
  if (!scores || !out_val || !out_idx) {
    set_error("topk: NULL pointer argument");
    return CVPP_ERR_INVALID_ARG;
  }
  if (B < 0 || N < 1 || N > 0xffffffffll || K < 1 || K > kTkMaxK || (long long)K > N) {
    set_error("topk: bad sizes (B=%d N=%lld K=%d; 1 <= K <= min(N, %d), N < 2^32)", B, N, K, kTkMaxK);
    return CVPP_ERR_INVALID_ARG;
  }
  if (C < 0 || (C > 0 && W < 1)) {
    set_error("topk: the index split needs C > 0 and W > 0 (got C=%d W=%d)", C, W);
    return CVPP_ERR_INVALID_ARG;
  }
  if (B == 0) return CVPP_OK;
  if (!workspace || workspace_bytes < topk_workspace_bytes(B, K)) {
    set_error("topk: workspace of %zu bytes needed, got %zu", topk_workspace_bytes(B, K), workspace_bytes);
    return CVPP_ERR_WORKSPACE;
  }
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != CVPP_OK) return rc;
  TkParams p{};
  p.scores = scores;
  p.N = N;
  p.B = B;
  p.K = K;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  p.state = reinterpret_cast<TkState*>(base);
  base += tk_align((size_t)B * sizeof(TkState));
  p.hist = reinterpret_cast<uint32_t*>(base);
  base += tk_align((size_t)B * kTkBins * sizeof(uint32_t));
  p.in_list = reinterpret_cast<uint64_t*>(base);
  base += tk_align((size_t)B * (size_t)K * sizeof(uint64_t));
  p.bd_list = reinterpret_cast<uint64_t*>(base);
  p.out_val = out_val;
  p.out_idx = out_idx;
  p.C = C;
  p.W = W;
  p.out_cls = out_cls;
  p.out_y = out_y;
  p.out_x = out_x;
  p.out_pixel = out_pixel;

  // digit schedule: 32 score bits as 12 + 12 + 8, then the significant index bits in 12-bit digits
  int shifts[kTkMaxLevels], bins[kTkMaxLevels], levels = 0;
  shifts[levels] = 52, bins[levels++] = 4096;
  shifts[levels] = 40, bins[levels++] = 4096;
  shifts[levels] = 32, bins[levels++] = 256;
  int idx_bits = 1;
  while (idx_bits < 32 && (1ll << idx_bits) < N) ++idx_bits;
  for (int hi = idx_bits; hi > 0;) {
    const int lo = hi > 12 ? hi - 12 : 0;
    shifts[levels] = lo, bins[levels++] = 1 << (hi - lo);
    hi = lo;
  }

  CVPP_CUDA_TRY(cudaMemsetAsync(p.hist, 0, (size_t)B * kTkBins * sizeof(uint32_t), stream));
  topk_init_kernel<<<(B + 127) / 128, 128, 0, stream>>>(p);
  // CTAs per image: enough to fill the machine, each with at least ~16 K elements
  long long per_img = (N + 16383) / 16384;
  long long want = (2ll * di.sms + B - 1) / B;
  int gx = (int)(per_img < want ? per_img : want);
  if (gx < 1) gx = 1;
  const dim3 grid((unsigned)gx, (unsigned)B);
  for (int l = 0; l < levels; ++l) {
    p.shift = shifts[l];
    p.bins = bins[l];
    topk_hist_kernel<<<grid, kTkThreads, 0, stream>>>(p);
    topk_select_kernel<<<B, 1024, 0, stream>>>(p);
  }
  topk_collect_kernel<<<grid, kTkThreads, 0, stream>>>(p);
  const int smem = (kTkMaxK + kTkSortCap) * (int)sizeof(uint64_t);
  static unsigned long long attr_done = 0;
  rc = ensure_smem_attr(reinterpret_cast<const void*>(topk_final_kernel), smem, di.device, &attr_done);
  if (rc != CVPP_OK) return rc;
  topk_final_kernel<<<B, 1024, smem, stream>>>(p);
  CVPP_CUDA_TRY(cudaGetLastError());
  return CVPP_OK;
}

}  // namespace cvpp

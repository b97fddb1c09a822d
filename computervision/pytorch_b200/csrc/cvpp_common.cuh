// cvpp_common.cuh — shared device helpers for libcvpp (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "cvpp.h"

#define CVPP_KEY_ANCHOR_BITS 21
#define CVPP_KEY_SCORE_BITS 31
#define CVPP_KEY_CLASS_BITS 12
#define CVPP_MAX_ANCHORS (1 << CVPP_KEY_ANCHOR_BITS)
#define CVPP_MAX_CLASSES (1 << CVPP_KEY_CLASS_BITS)
#define CVPP_MAX_LEVELS 4
#define CVPP_MAX_PEERS 16

// torchvision.ops.batched_nms on CPU switches to the per-class branch when boxes.numel() > 4000
// (torchvision/ops/boxes.py:80), i.e. more than 1000 boxes.
#define CVPP_TRICK_MAX_BOXES 1000
// ... and on CUDA tensors when boxes.numel() > 20000, i.e. more than 5000 boxes (same line).
#define CVPP_TRICK_MAX_BOXES_CUDA 5000

namespace cvpp {

// ---- error plumbing (host) -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
// cached per-device facts (SM count, opt-in shared memory); returns CVPP_OK or CVPP_ERR_CUDA
struct DeviceInfo {
  int device;
  int sms;
  int max_smem;
};
int device_info(DeviceInfo* out);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): `done` is a per-kernel bitmask
int ensure_smem_attr(const void* func, int bytes, int device, unsigned long long* done);

#define CVPP_CUDA_TRY(expr)                                     \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return ::cvpp::cuda_fail(_e, #expr); \
  } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// A chain of short dependent kernels pays ~2 us of launch latency per link.  launch_pdl() marks a launch as a
// programmatic dependent of the previous kernel in the stream: it is scheduled as soon as every CTA of that kernel has
// executed pdl_trigger() (or exited), and its threads block in pdl_wait() until that kernel has completed and its
// writes are visible - so the launch latency (and whatever runs above pdl_wait) overlaps the predecessor.  Without
// a trigger in the predecessor, or after a non-kernel stream operation, the launch is an ordinary ordered launch.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// Counters zeroed by a KERNEL that triggers its dependents at once (instead of a memset node): the filter kernel that follows,
// launched with launch_pdl(), sets up and streams its first tiles while this runs and calls pdl_wait() only before its first
// append.  Because this kernel itself is an ordinary launch (it starts after everything earlier in the stream has completed),
// the dependent may read its INPUTS before pdl_wait(); only the counters must not be touched before it.
#ifdef __CUDACC__
static __global__ void __launch_bounds__(256) cvpp_zero_i32_kernel(int32_t* __restrict__ p, int n) {
  pdl_trigger();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0;
}
static inline cudaError_t zero_counts_async(int32_t* p, int n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  cvpp_zero_i32_kernel<<<(n + 255) / 256, 256, 0, stream>>>(p, n);
  return cudaGetLastError();
}
#endif

// ---- single-use (streamed) global loads ------------------------------------------------------
// Inputs that are read exactly once carry an L2 evict-first policy: hundreds of MB of streamed predictions
// then do not displace what the next kernel needs from L2 (candidate keys, dense boxes, its code).
#ifdef __CUDACC__
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(policy));
  return v;
}
#endif

// ---- candidate keys --------------------------------------------------------------------------
// class-major packing: [class:12 | inv_score:31 | anchor:21]
__host__ __device__ __forceinline__ uint64_t key_pack(uint32_t cls, uint32_t score_bits, uint32_t anchor) {
  return ((uint64_t)cls << 52) | ((uint64_t)(0x7fffffffu - (score_bits & 0x7fffffffu)) << 21) | (uint64_t)anchor;
}
__host__ __device__ __forceinline__ uint32_t key_cls(uint64_t k) { return (uint32_t)(k >> 52); }
__host__ __device__ __forceinline__ uint32_t key_score_bits(uint64_t k) {
  return 0x7fffffffu - (uint32_t)((k >> 21) & 0x7fffffffu);
}
__host__ __device__ __forceinline__ uint32_t key_anchor(uint64_t k) { return (uint32_t)(k & 0x1fffffu); }
// score-major packing used for the coordinate-trick branch: [inv_score:31 | anchor:21 | class:12]
__host__ __device__ __forceinline__ uint64_t key_to_score_major(uint64_t k) {
  return ((k & 0x000fffffffffffffull) << 12) | (k >> 52);
}
__host__ __device__ __forceinline__ uint64_t key_from_score_major(uint64_t k) {
  return (k >> 12) | ((k & 0xfffull) << 52);
}

__host__ __device__ __forceinline__ bool rule_uses_trick(int rule, int n) {
  return rule == CVPP_NMS_RULE_COORD_TRICK || (rule == CVPP_NMS_RULE_TORCHVISION_CPU && n <= CVPP_TRICK_MAX_BOXES) ||
         (rule == CVPP_NMS_RULE_TORCHVISION_CUDA && n <= CVPP_TRICK_MAX_BOXES_CUDA);
}

#ifdef __CUDACC__
// ---- exact fp32 arithmetic -------------------------------------------------------------------
// The reference runs every step as a separate eager op: one IEEE rounding each, never fused.
// The explicit round-to-nearest intrinsics are never contracted into FMAs by nvcc.
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// torch.sigmoid restated as 1 / (1 + exp(-x)) with the full-precision expf (<= 1 ulp)
__device__ __forceinline__ float sigmoid_precise(float x) { return fdiv(1.0f, fadd(1.0f, expf(-x))); }

// torchvision nms_kernel.cpp inner loop: is box j suppressed by kept box i?
//   ovr = inter / (iarea + jarea - inter);  suppressed iff (double)ovr > thr  <=>  ovr > thr_eff
// thr_eff = largest float <= thr (host-computed, thr in [0,1]).  inter == 0 (or NaN) gives ovr == 0 or
// NaN, which never exceeds a non-negative threshold, so the division is skipped for disjoint boxes.
__device__ __forceinline__ bool iou_suppresses(const float4& bi, float ai, const float4& bj, float aj, float thr_eff) {
  float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
  float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
  float w = fmaxf(0.0f, fsub(xx2, xx1));
  float h = fmaxf(0.0f, fsub(yy2, yy1));
  float inter = fmul(w, h);
  if (!(inter > 0.0f)) return false;
  float ovr = fdiv(inter, fsub(fadd(ai, aj), inter));
  return ovr > thr_eff;
}
__device__ __forceinline__ float box_area(const float4& b) { return fmul(fsub(b.z, b.x), fsub(b.w, b.y)); }

// ---- mbarrier / bulk-copy (TMA engine, 1-D) wrappers -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// (mbarrier.try_wait with a suspend-time hint was measured in the head-fused kernel: the polling loops execute just as many
// instructions - the hardware returns after its own short limit whatever the hint - so the plain form is used.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk async copy; completion is signalled on `bar` in bytes.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// the same with an L2 cache policy (single-use input streams: l2_policy_evict_first)
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// ---- block-wide sort of 64-bit keys --------------------------------------------------------------
__host__ __device__ __forceinline__ int pow2_ceil(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

// Plain bitonic network through a generic pointer (shared OR global): one __syncthreads per substep.
// Only used for segments that do not fit shared memory.  Ends with a __syncthreads().
__device__ __forceinline__ void bitonic_sort_u64_generic(uint64_t* a, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int l = i | j;
        bool up = (i & k) == 0;
        uint64_t x = a[i], y = a[l];
        if ((x > y) == up) {
          a[i] = y;
          a[l] = x;
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ void cmpswap_u64(uint64_t& a, uint64_t& b, bool up) {
  const bool sw = (a > b) == up;
  const uint64_t t = a;
  a = sw ? b : a;
  b = sw ? t : b;
}

// Substeps j = j_start, j_start/2, ..., 1 of the bitonic merge of size k on a warp-private block of
// 32*E keys held in registers, striped: x[e] is element base + e*32 + lane.  Lane-crossing substeps
// use shuffles, element-crossing ones stay in registers.
template <int E>
__device__ __forceinline__ void warp_bitonic_steps(uint64_t (&x)[E], int base, int k, int j_start) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 16 * E; j >= 1; j >>= 1) {
    if (j > j_start) continue;
    if (j >= 32) {
      const int je = j >> 5;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if ((e & je) == 0) {
          const bool up = ((base + e * 32 + lane) & k) == 0;
          cmpswap_u64(x[e], x[e | je], up);
        }
      }
    } else {
      const bool lower = (lane & j) == 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const uint64_t o = __shfl_xor_sync(0xffffffffu, x[e], j);
        const bool up = ((base + e * 32 + lane) & k) == 0;
        const bool take_min = lower == up;
        const bool o_lt = o < x[e];
        x[e] = (take_min == o_lt) ? o : x[e];
      }
    }
  }
}

// Hybrid bitonic sort of P keys in SHARED memory by the whole CTA (blockDim.x = T, P = E*T for
// P >= T, else E = 1 and only P/32 warps work; P is a power of two >= 32).  Every merge level runs
// its warp-local substeps (j < 32*E) in registers, so a sort of 4096 keys takes 27 block barriers
// instead of 78.  Ends with a __syncthreads().
template <int E>
__device__ __forceinline__ void block_sort_smem_e(uint64_t* a, int P) {
  constexpr int J0 = 32 * E;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nblocks = P / J0;
  const int base = warp * J0;
  uint64_t x[E];
  if (warp < nblocks) {
#pragma unroll
    for (int e = 0; e < E; ++e) x[e] = a[base + e * 32 + lane];
    for (int k = 2; k <= J0; k <<= 1) warp_bitonic_steps<E>(x, base, k, k >> 1);
#pragma unroll
    for (int e = 0; e < E; ++e) a[base + e * 32 + lane] = x[e];
  }
  __syncthreads();
  for (int k = 2 * J0; k <= P; k <<= 1) {
    for (int j = k >> 1; j >= J0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const bool up = (i & k) == 0;
        const uint64_t u = a[i], v = a[l];
        if ((u > v) == up) {
          a[i] = v;
          a[l] = u;
        }
      }
      __syncthreads();
    }
    if (warp < nblocks) {
#pragma unroll
      for (int e = 0; e < E; ++e) x[e] = a[base + e * 32 + lane];
      warp_bitonic_steps<E>(x, base, k, J0 >> 1);
#pragma unroll
      for (int e = 0; e < E; ++e) a[base + e * 32 + lane] = x[e];
    }
    __syncthreads();
  }
}

// dispatch on keys per thread; `a` MUST be a shared-memory array of P keys
__device__ __forceinline__ void block_sort_smem(uint64_t* a, int P) {
  const int T = blockDim.x;
  const int E = P > T ? P / T : 1;
  switch (E) {
    case 1: block_sort_smem_e<1>(a, P); break;
    case 2: block_sort_smem_e<2>(a, P); break;
    case 4: block_sort_smem_e<4>(a, P); break;
    case 8: block_sort_smem_e<8>(a, P); break;
    case 16: block_sort_smem_e<16>(a, P); break;
    default: bitonic_sort_u64_generic(a, P); break;
  }
}
#endif  // __CUDACC__

}  // namespace cvpp

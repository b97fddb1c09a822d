// Timing-only probe: cycles per tcgen05.mma kind::tf32 as a function of N, operand major-ness and accumulator rotation.
// nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate_probe umma_rate_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct V { int m, n, a_mn_major, nacc, reps, kind_f16; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int swz) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
         (1ull << 46) | ((uint64_t)swz << 61);
}

__global__ void __launch_bounds__(128) probe(V v, long long* out) {
  extern __shared__ unsigned char dyn[];
  unsigned char* base = dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u);
  float* sa = reinterpret_cast<float*>(base);            // 16 KB: 128 cells x 32 k
  float* sb = reinterpret_cast<float*>(base + 16384);    // 32 KB: 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(base + 49152 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 12288; i += 128) sa[i] = 0.f;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)v.a_mn_major << 15) | ((uint32_t)(v.n >> 3) << 17) | ((uint32_t)(v.m >> 4) << 24);
    const uint32_t idesc_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(v.n >> 3) << 17) | ((uint32_t)(v.m >> 4) << 24);
    const uint64_t da0 = v.a_mn_major ? desc(smem_u32(sa), 4096, 512, 1) : desc(smem_u32(sa), 16, 1024, 2);
    const uint64_t db0 = desc(smem_u32(sb), 16, 1024, 2);
    const long long t0 = clock64();
    // 8 MMAs per iteration, descriptors precomputed: the loop body is 8 x (MMA) + a counter
    uint64_t da[4], db[4];
    for (int ks = 0; ks < 4; ++ks) {
      da[ks] = da0 + (uint64_t)(v.a_mn_major ? ks * 64 : ks * 2);
      db[ks] = db0 + (uint64_t)(ks * 2);
    }
    const uint32_t d0 = tmem, d1 = tmem + (v.nacc > 1 ? 256u : 0u);
    const uint32_t idesc = v.kind_f16 ? idesc_bf16 : idesc_tf32;
#define MMA(d, a, b) \
    if (v.kind_f16) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1u) : "memory"); \
    else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1u) : "memory");
    for (int rep = 0; rep < v.reps; rep += 8) {
      MMA(d0, da[0], db[0]) MMA(d0, da[1], db[1]) MMA(d0, da[2], db[2]) MMA(d0, da[3], db[3])
      MMA(d1, da[0], db[0]) MMA(d1, da[1], db[1]) MMA(d1, da[2], db[2]) MMA(d1, da[3], db[3])
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    uint32_t ok = 0;
    for (int spin = 0; !ok && spin < (1 << 26); ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  V vs[] = {
      {128, 64, 0, 1, 400, 0}, {128, 64, 1, 1, 400, 0}, {128, 64, 1, 2, 400, 0},
      {128, 80, 1, 1, 400, 0}, {128, 128, 1, 1, 400, 0}, {128, 144, 1, 1, 400, 0}, {128, 256, 1, 1, 400, 0}, {128, 256, 0, 1, 400, 0},
      {64, 64, 1, 1, 400, 0}, {64, 256, 1, 1, 400, 0}, {64, 256, 0, 1, 400, 0},
      {128, 64, 0, 1, 400, 1}, {128, 256, 0, 1, 400, 1},
  };
  for (auto& v : vs) {
    for (int it = 0; it < 2; ++it) {
      probe<<<1, 128, 65536>>>(v, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (it == 1)
        printf("M=%3d N=%3d A %s nacc=%d %s: %s  issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", v.m, v.n, v.a_mn_major ? "MN-major" : "K-major ", v.nacc,
               v.kind_f16 ? "bf16 K=16" : "tf32 K=8 ", cudaGetErrorString(e), (double)h[0] / v.reps, (double)h[1] / v.reps);
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}

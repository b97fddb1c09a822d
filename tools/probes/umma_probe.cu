// Standalone probe of tcgen05.mma kind::tf32 descriptor conventions (MN-major A from NCHW activations, K-major B).
// Builds: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu ; prints max error per variant.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 16;

struct Variant {
  int a_lbo, a_sbo, a_kstep;   // bytes
  int b_lbo, b_sbo, b_kstep;
  int a_major_bit;             // instruction descriptor bit 15
  int a_layout;                // 0: [4 atoms][K rows][128 B] swizzled (MN-major); 1: K-major swizzled [M rows][K*4 B] (needs transposed data)
  int reps;
  int commit_each;             // 1: tcgen05.commit to a scratch mbarrier after every pair of MMAs (as a pipelined kernel does per stage)
  int swz;                     // layout type field of A (2 = 128B, 1 = 128B base 32B); B always uses 2
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int swz) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
         (1ull << 46) | ((uint64_t)swz << 61);
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ X /*[K][M]*/, const float* __restrict__ W /*[N][K]*/,
                                              float* __restrict__ out /*[M][N]*/, Variant v) {
  extern __shared__ unsigned char dyn[];
  unsigned char* base = dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u);
  float* sa = reinterpret_cast<float*>(base);            // 8 KB
  float* sb = reinterpret_cast<float*>(base + 16384);    // N * 128 B = 8 KB (one K atom of 32, only 16 used)
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 24576);
  uint32_t* slot = reinterpret_cast<uint32_t*>(base + 24576 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 4096; i += 128) { sa[i] = 0.f; sb[i >> 1] = 0.f; }
  __syncthreads();
  if (v.a_layout == 0) {
    // MN-major: element (k, m): atom = m / 32, row k, 16-byte chunk (m % 32) / 4 XOR (k % 8)
    for (int i = tid; i < K * M; i += 128) {
      const int k = i / M, m = i % M;
      const int atom = m >> 5, mm = m & 31;
      const int off = atom * (K * 32) + k * 32 + ((((mm >> 2) ^ (k & 7)) << 2) | (mm & 3));
      sa[off] = X[k * M + m];
    }
  } else if (v.a_layout == 2) {
    // MN-major with the 32-byte-atom 128B swizzle (Swizzle<2,5,2> on the byte address): 32-byte chunk (m % 32) / 8 XOR (k % 4)
    for (int i = tid; i < K * M; i += 128) {
      const int k = i / M, m = i % M;
      const int atom = m >> 5, mm = m & 31;
      const int off = atom * (K * 32) + k * 32 + ((((mm >> 3) ^ (k & 3)) << 3) | (mm & 7));
      sa[off] = X[k * M + m];
    }
  } else {
    // K-major: row m (128 B: 32 floats, 16 used), chunk (k / 4) XOR (m % 8)
    for (int i = tid; i < K * M; i += 128) {
      const int k = i / M, m = i % M;
      const int off = m * 32 + ((((k >> 2) ^ (m & 7)) << 2) | (k & 3));
      sa[off] = X[k * M + m];
    }
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const int off = n * 32 + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3));
    sb[off] = W[n * K + k];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  long long t0 = 0;
  if (tid == 0) {
    t0 = clock64();
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)v.a_major_bit << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int rep = 0; rep < v.reps; ++rep) {
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t da = desc(smem_u32(sa) + ks * v.a_kstep, v.a_lbo, v.a_sbo, v.swz);
      const uint64_t db = desc(smem_u32(sb) + ks * v.b_kstep, v.b_lbo, v.b_sbo, 2);
      const uint32_t acc = ks > 0 && rep == 0 ? 1u : (rep > 0 ? (ks > 0) : 0u);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    if (v.commit_each) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 1)) : "memory");
    }
    const long long t_issue = clock64() - t0;
    if (v.reps > 1) printf("   issue loop alone: %lld cycles (%.1f per pair)\n", t_issue, (double)t_issue / v.reps);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // wait
  {
    uint32_t ok = 0;
    for (int spin = 0; !ok && spin < (1 << 24); ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    if (!ok && tid == 0) printf("TIMEOUT waiting for the MMA commit\n");
  }
  if (tid == 0 && v.reps > 1) printf("   %d x %d MMAs (M=128 N=%d K=8): %lld cycles -> %.1f cycles per MMA\n", v.reps, K / 8, N, clock64() - t0, (double)(clock64() - t0) / (v.reps * (K / 8)));
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(trow + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[tid * N + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  std::vector<float> X(K * M), W(N * K), ref(M * N);
  srand(1);
  for (auto& x : X) x = (rand() % 65 - 32) / 8.0f;
  for (auto& w : W) w = (rand() % 65 - 32) / 64.0f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)X[k * M + m] * W[n * K + k];
      ref[m * N + n] = (float)s;
    }
  float *dX, *dW, *dO;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dO, ref.size() * 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  Variant vs[] = {
      // a_lbo a_sbo a_kstep  b_lbo b_sbo b_kstep a_major a_layout swz
      {K * 128, 1024, 1024, 16, 1024, 32, 1, 0, 1, 0, 2},   // the design: MN-major A
      {1024, K * 128, 1024, 16, 1024, 32, 1, 0, 1, 0, 2},   // LBO / SBO swapped for A
      {K * 128, 1024, 1024, 0, 1024, 32, 1, 0, 1, 0, 2},    // B LBO = 0
      {16, 1024, 32, 16, 1024, 32, 0, 1, 1, 0, 2},          // control: both K-major (classic layout)
      {0, 1024, 32, 0, 1024, 32, 0, 1, 1, 0, 2},            // control with LBO 0
      {K * 128, 512, 1024, 16, 1024, 32, 1, 2, 1, 0, 1},    // MN-major A, SWIZZLE_128B_BASE32B (layout type 1), 4-row K groups 512 B apart
      {512, K * 128, 1024, 16, 1024, 32, 1, 2, 1, 0, 1},    // ... LBO / SBO swapped
      {K * 128, 1024, 1024, 16, 1024, 32, 1, 2, 1, 0, 1},   // ... SBO = 1024
      {16, 1024, 32, 16, 1024, 32, 0, 1, 200, 0, 2},          // timing: K-major A
      {K * 128, 512, 1024, 16, 1024, 32, 1, 2, 200, 0, 1},    // timing: MN-major A (BASE32B)
      {K * 128, 512, 1024, 16, 1024, 32, 1, 2, 200, 1, 1},    // timing: MN-major A + a commit after every pair
  };
  for (size_t i = 0; i < sizeof(vs) / sizeof(vs[0]); ++i) {
    cudaMemset(dO, 0xff, ref.size() * 4);
    probe<<<1, 128, 32768>>>(dX, dW, dO, vs[i]);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(ref.size());
    cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost);
    double me = 0; int nz = 0;
    for (size_t j = 0; j < o.size(); ++j) { me = fmax(me, fabs((double)o[j] - ref[j])); nz += o[j] != 0.f; }
    printf("variant %zu: %s  max_err %.6g  nonzero %d/%zu   out[0][0..3] = %g %g %g %g   ref = %g %g %g %g   out[1][0] %g ref %g  out[33][0] %g ref %g\n", i,
           cudaGetErrorString(e), me, nz, o.size(), o[0], o[1], o[2], o[3], ref[0], ref[1], ref[2], ref[3], o[N], ref[N], o[33 * N], ref[33 * N]);
    if (e != cudaSuccess) break;
  }
  return 0;
}

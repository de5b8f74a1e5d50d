// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128 or 64, K = 16) as a function of N, with both
// operands resident in shared memory (128-byte swizzle, K-major), one CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ELECT: the issuing lane is chosen with elect.sync inside warp 0 (all 32 lanes reach it) instead of `threadIdx.x == 0`
template <bool ELECT>
__global__ void __launch_bounds__(128, 1) rate_kernel(int M, int N, int iters, int a_blocks, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  bool issuer = threadIdx.x == 0;
  if (ELECT) issuer = (threadIdx.x < 32) && elect_one();
  if (issuer) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t b_base = base + 128 * 1024;                     // B: up to 256 rows x 128 B
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t a = base + (uint32_t)(it & (a_blocks - 1)) * 16384u;   // rotate over a_blocks (a power of two) different A tiles
#pragma unroll
      for (int k = 0; k < 4; ++k) umma(tm, desc_sw128(a) + 2u * k, desc_sw128(b_base) + 2u * k, idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_a) : "memory");
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// Second experiment: does the tensor pipe lose time when consecutive k-blocks (4 MMAs each) change their N or their TMEM
// window?  (conv_up4_kernel interleaves N = 256 / 128 / 64 MMAs on different 64-column windows of one accumulator.)
__global__ void __launch_bounds__(128, 1) mix_kernel(int n0, int d0, int n1, int d1, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t id0 = (1u << 4) | ((uint32_t)(n0 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t id1 = (1u << 4) | ((uint32_t)(n1 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t b_base = base + 128 * 1024;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t a = base + (uint32_t)(it % 8) * 16384u;
      const bool odd = it & 1;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma(tm + (odd ? d1 : d0), desc_sw128(a) + 2u * k, desc_sw128(b_base) + 2u * k, odd ? id1 : id0, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_a) : "memory");
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4000;
  for (int elect : {0, 1})
  for (int M : {128, 64})
  for (int ctas : {148}) {
    for (int N : {16, 32, 64, 96, 128, 192, 256}) {
      if (elect) rate_kernel<true><<<ctas, 128, 200 * 1024>>>(M, N, iters, 8, d);
      else rate_kernel<false><<<ctas, 128, 200 * 1024>>>(M, N, iters, 8, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d failed: %s\n", N, cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
      const double cyc = (double)mx / (iters * 4.0);
      printf("%s M=%3d ctas=%3d N=%3d  %.1f cycles per MMA (K=16)  -> %.0f MAC/clk/SM, smem operand read %.0f B/clk\n", elect ? "elect.sync " : "thread 0   ", M, ctas, N, cyc,
             (double)M * N * 16 / cyc, (M * 16 * 2 + N * 16 * 2) / cyc);
    }
  }
  cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Mix { int n0, d0, n1, d1; const char* what; };
  const Mix mixes[] = {{64, 0, 64, 0, "N=64 same window"}, {64, 0, 64, 64, "N=64, window alternates"}, {64, 0, 64, 192, "N=64, windows 0 / 192"},
                       {256, 0, 64, 64, "N=256 / N=64 alternate"}, {128, 0, 64, 128, "N=128 / N=64 alternate"}, {256, 0, 128, 0, "N=256 / N=128 alternate"},
                       {192, 0, 192, 64, "N=192, window shifts by 64"}, {256, 0, 256, 256, "N=256, two accumulator stages"}};
  for (const Mix& m : mixes) {
    mix_kernel<<<148, 128, 200 * 1024>>>(m.n0, m.d0, m.n1, m.d1, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mix failed: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per_pair = (double)mx / (iters / 2.0);      // cycles per (4 MMAs of shape 0 + 4 MMAs of shape 1)
    const double ideal = 4.0 * ((m.n0 / 2 > 71 ? m.n0 / 2 : 71) + (m.n1 / 2 > 71 ? m.n1 / 2 : 71));
    printf("mix %-32s %.0f cycles per pair of k-blocks, %.0f if every MMA ran at max(71, N/2)\n", m.what, per_pair, ideal);
  }
  return 0;
}

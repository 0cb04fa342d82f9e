// Raw tcgen05.mma issue-rate probe (SS mode, bf16 -> fp32, K = 16 per instruction): cycles per MMA for (M, N, accumulators in rotation).
// One CTA per SM, zero operands in SWIZZLE_128B K-major tiles, one elected lane of a warp-uniform loop issues ITER x 4 MMAs, commits and waits.
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_bench tools/umma_bench.cu -lcuda && /tmp/umma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128, 1) bench(int M, int N, int nacc, int iters, int a_sbo, long long* out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* p0 = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < (64 + 100) * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(p0)[i] = make_uint4(0, 0, 0, 0);
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tptr;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_hi = ((uint32_t)a_sbo >> 4) | (1u << 14) | (2u << 29), b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo = ((base & 0x3FFFFu) >> 4) | (1u << 16), b_lo = (((base + 64 * 1024) & 0x3FFFFu) >> 4) | (1u << 16);
    long long t0 = clock64();
    if (mode == 0) {
      for (int i = 0; i < iters; i++) {
        const uint32_t d = tmem + (uint32_t)((i % nacc) * N);
        if (leader) {
          umma(d, pack64(a_lo, a_hi), pack64(b_lo, b_hi), idesc);
          umma(d, pack64(a_lo + 2, a_hi), pack64(b_lo + 2, b_hi), idesc);
          umma(d, pack64(a_lo + 4, a_hi), pack64(b_lo + 4, b_hi), idesc);
          umma(d, pack64(a_lo + 6, a_hi), pack64(b_lo + 6, b_hi), idesc);
        }
      }
    } else {
      // conv-like: tap t reads A shifted by ((t / 3) * 10 + t % 3) pixel rows of 128 B and its own 8 KB weight chunk (mode 2: same weight chunk)
      for (int i = 0; i < iters; i++) {
        const int t = i % 9;
        const uint32_t d = tmem + (uint32_t)((i % nacc) * N);
        const uint32_t al = a_lo + (uint32_t)((t / 3) * 10 + t % 3) * 8u, bl = b_lo + (mode == 1 ? (uint32_t)t * (uint32_t)(N * 8) : 0u);
        if (leader) {
          umma(d, pack64(al, a_hi), pack64(bl, b_hi), idesc);
          umma(d, pack64(al + 2, a_hi), pack64(bl + 2, b_hi), idesc);
          umma(d, pack64(al + 4, a_hi), pack64(bl + 4, b_hi), idesc);
          umma(d, pack64(al + 6, a_hi), pack64(bl + 6, b_hi), idesc);
        }
      }
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (leader && blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
  const int iters = 2000;
  struct { int M, N, nacc, sbo; } cfg[] = {{128, 64, 1, 1024}, {128, 64, 2, 1024}, {128, 64, 4, 1024}, {128, 64, 1, 1280}, {128, 128, 1, 1024}, {128, 128, 2, 1024},
                                           {128, 256, 1, 1024}, {128, 256, 2, 1024}, {128, 32, 1, 1024}, {128, 16, 1, 1024}, {64, 64, 1, 1024}, {64, 128, 1, 1024},
                                           {64, 256, 1, 1024}, {64, 256, 2, 1024}, {64, 160, 1, 1024}, {128, 160, 1, 1024}, {128, 192, 1, 1024}};
  for (auto& c : cfg) {
    for (int grid : {1, 148}) {
      bench<<<grid, 128, 180 * 1024>>>(c.M, c.N, c.nacc, iters, c.sbo, out, 0);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
      double per = (double)cyc / (iters * 4.0);
      printf("{\"M\": %d, \"N\": %d, \"nacc\": %d, \"a_sbo\": %d, \"grid\": %d, \"cycles_per_mma\": %.1f, \"mac_per_clk\": %.0f, \"err\": \"%s\"}\n", c.M, c.N, c.nacc, c.sbo, grid, per,
             (double)c.M * c.N * 16 / per, cudaGetErrorString(e));
    }
  }
  for (int mode : {1, 2})
    for (int sbo : {1024, 1280}) {
      bench<<<148, 128, 180 * 1024>>>(128, 64, 1, iters, sbo, out, mode);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
      printf("{\"conv_like_mode\": %d, \"M\": 128, \"N\": 64, \"a_sbo\": %d, \"cycles_per_mma\": %.1f, \"err\": \"%s\"}\n", mode, sbo, (double)cyc / (iters * 4.0), cudaGetErrorString(e));
    }
  return 0;
}

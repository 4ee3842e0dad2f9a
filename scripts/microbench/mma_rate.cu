// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16 -> fp32), M = 128, with operand A read
// from shared memory (SS) or from tensor memory (TS), for N = 64 / 128 / 256, accumulating into
// one TMEM tile or alternating between two. One CTA per SM, operands are whatever the memories
// hold (only timing matters). Prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu -I../../keds_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace keds;

__device__ __forceinline__ void umma_bf16_ts_(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}

template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) k_rate(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar = sbase + 96 * 1024;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(gbase + 96 * 1024 + 64);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(slot)), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_bf16_f32(128, N);
    const uint64_t adesc = smem_desc_sw128(sbase);               // 128 rows x 128 B
    const uint64_t bdesc = smem_desc_sw128(sbase + 16 * 1024);   // N rows x 128 B
    const uint32_t a_tmem = tm + 384;                            // TS: 32 columns of "A" per k-block
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = tm + static_cast<uint32_t>(((i * 4 + k) % NACC) * N);
        if (TS) umma_bf16_ts_(d, a_tmem + k * 8, bdesc + 2u * k, idesc, 1u);
        else umma_bf16(d, adesc + 2u * k, bdesc + 2u * k, idesc, 1u);
      }
    }
    umma_commit(bar);
    while (!mbar_try_wait(bar, 0)) {}
    const long long t1 = clock64();
    if (blockIdx.x == 0) *out = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

template <int N, bool TS, int NACC>
void run(const char* name, long long* d_out) {
  const int iters = 4000;
  const size_t smem = 100 * 1024;
  cudaFuncSetAttribute(k_rate<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_rate<N, TS, NACC><<<148, 128, smem>>>(iters, d_out);
  k_rate<N, TS, NACC><<<148, 128, smem>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %8.1f cycles per MMA (floor %d)  %s\n", name, (double)cyc / (iters * 4.0), N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  run<256, false, 1>("SS N=256 one accumulator", d_out);
  run<256, true, 1>("TS N=256 one accumulator", d_out);
  run<128, false, 1>("SS N=128 one accumulator", d_out);
  run<128, true, 1>("TS N=128 one accumulator", d_out);
  run<64, false, 1>("SS N=64  one accumulator", d_out);
  run<64, true, 1>("TS N=64  one accumulator", d_out);
  run<64, false, 2>("SS N=64  two accumulators", d_out);
  run<64, true, 2>("TS N=64  two accumulators", d_out);
  run<128, true, 2>("TS N=128 two accumulators", d_out);
  return 0;
}

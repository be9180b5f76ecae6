// Micro-probe: how long does a chain of tcgen05.mma (kind::f16, M = 128, SS operands, SWIZZLE_128B K-major) take per
// instruction as a function of N and of how the accumulators are rotated?  One CTA, one issuing thread, operands
// resident in shared memory (contents irrelevant).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I neutts_air_b200/csrc
#include "common.cuh"
#include <cstdio>
using namespace nt;

template <bool ELECT>
__global__ void __launch_bounds__(128, 1) probe(int N, int nacc_log2, int nmma, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (8 * 16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  bool issuer;
  if (ELECT) issuer = warp == 0 && elect_one();
  else issuer = threadIdx.x == 0;
  if (issuer) {
    const uint32_t idesc = umma_idesc(1, 128, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 8 * 16384);
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < nmma; ++i) {
        const int kb = (i >> 2) & 7, k = i & 3;
        const uint64_t ad = umma_desc_sw128(a0 + kb * 16384) + 2 * k;
        const uint64_t bd = umma_desc_sw128(b0) + 2 * k;
        umma_bf16(tm + (i & ((1 << nacc_log2) - 1)) * N, ad, bd, idesc, (i >> nacc_log2) ? 1u : 0u);
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      out[2 * rep] = t1 - t0, out[2 * rep + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int smem = 8 * 16384 + 32768 + 1024;
  cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int Ns[] = {16, 64, 128, 256};
  for (int el = 0; el < 2; ++el)
  for (int N : Ns)
    for (int nl : {0, 2})
      for (int nmma : {8, 56}) {
        const int nacc = 1 << nl;
        if (nacc * N > 512) continue;
        if (el) probe<true><<<1, 128, smem>>>(N, nl, nmma, d);
        else probe<false><<<1, 128, smem>>>(N, nl, nmma, d);
        long long h[8];
        cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("MMA-PROBE %s M128 N%-3d nacc %d nmma %2d: issue %5lld cyc, complete %6lld cyc -> %.1f cyc/mma (last rep)\n", el ? "elect.sync" : "tid==0    ", N, nacc, nmma, h[4], h[5],
               double(h[5]) / nmma);
      }
  return 0;
}

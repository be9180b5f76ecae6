// Shared device helpers for the sm_100a kernels: mbarrier, bulk/TMA copies, tcgen05
// (UMMA + TMEM), programmatic dependent launch, small math utilities.
// Everything is inline PTX; no CUTLASS/CuTe headers are included.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

#define NT_DEVINL __device__ __forceinline__

// Spin bound for every mbarrier wait: a protocol bug becomes a trap (the launch fails with
// an error the host reports) instead of a hung GPU box.  try_wait itself suspends the thread for
// a hardware-defined slice per call, so 2^22 tries is seconds — far beyond any legitimate wait
// in these kernels (the longest is one decode step, ~1 ms).
#ifndef NT_SPIN_LIMIT
#define NT_SPIN_LIMIT (1u << 22)
#endif

namespace nt {

constexpr int kWarp = 32;

NT_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
NT_DEVINL int lane_id() { return threadIdx.x & 31; }
// One lane of a fully converged warp (the same lane on every call).  Code issuing uniform-datapath instructions
// (tcgen05.mma / commit, TMA) must sit under THIS predicate inside warp-uniform control flow: behind a plain
// `if (lane == 0)` the compiler cannot prove that a single thread is active and wraps every such instruction in an
// ELECT / BRA.U.ANY loop with R2UR moves -- measured ~160 cycles per tcgen05.mma issue instead of back-to-back issue.
// Call it AT the use site, after any data-dependent wait loop: elect.sync (full mask) is also the reconvergence point,
// and the compiler emits the guarded UTMALDG / UTCHMMA unpredicated (only the operand moves carry the predicate), so
// a diverged lane group without the leader would execute it with stale uniform registers (memcheck: out-of-range
// shared address; found with a cached `leader` flag behind an mbarrier spin).
NT_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// value of lane 0, which the compiler then knows to be warp-uniform
NT_DEVINL int uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }
NT_DEVINL int warp_id() { return threadIdx.x >> 5; }

// ---------------------------------------------------------------- PDL (griddepcontrol)
// launch_dependents: lets the next kernel in the stream start its prologue (weight
// prefetch) early; wait: blocks until the previous kernel has fully completed and its
// writes are visible.  Rule used throughout: before pdl_wait() a kernel may only touch
// immutable weight memory and its own shared memory.
NT_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
NT_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier
NT_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
NT_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
NT_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
NT_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
NT_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
NT_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
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
// out of line on purpose: the report is cold code, inlined it would sit (printf argument set-up and all) in the
// instruction stream of every wait loop of kernels that already fight for the instruction cache
__device__ __noinline__ inline void nt_timeout(const char* what) {
  printf("neutts_b200: timed out waiting for %s (block %d,%d thread %d)\n", what, blockIdx.x, blockIdx.y, threadIdx.x);
  __trap();
}
NT_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > NT_SPIN_LIMIT) nt_timeout("an mbarrier");
  }
}

// ---------------------------------------------------------------- 1-D bulk copy (TMA engine, no tensor map)
// global -> shared, completion counted on an mbarrier.  dst/src 16-byte aligned, bytes % 16 == 0.
NT_DEVINL void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
NT_DEVINL void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- 2-D tiled TMA load
NT_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
NT_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// Same with an L2 eviction-priority hint (createpolicy): weights that stream through once per decode step are loaded
// evict_first so that they do not push the small hot set (KV pages, hand-off buffers, logits) out of the 126 MB L2.
NT_DEVINL uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
NT_DEVINL void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
NT_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
NT_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Whole warp.  Writes the TMEM base address to *smem_slot.  ncols: power of two in [32, 512].
NT_DEVINL void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
NT_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// K-major operand tile in shared memory written by TMA with CU_TENSOR_MAP_SWIZZLE_128B:
// rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), descriptor version 1 (sm_100),
// layout type 2 (SWIZZLE_128B).  Field layout follows the UMMA shared-memory descriptor
// (start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version [46,48), layout [61,64)).
NT_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (ignored for swizzled K-major), canonical value 1
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO = 1024 B between 8-row core groups
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16 / kind::tf32, fp32 accumulate, both operands K-major.
// fmt: 0 = f16, 1 = bf16, 2 = tf32.
__host__ __device__ constexpr uint32_t umma_idesc(int fmt, int M, int N) {
  return (1u << 4)                      // C format = F32
         | (uint32_t(fmt) << 7)         // A format
         | (uint32_t(fmt) << 10)        // B format
         | (0u << 15) | (0u << 16)      // A, B K-major
         | (uint32_t(N >> 3) << 17)     // N / 8
         | (uint32_t(M >> 4) << 24);    // M / 16
}

NT_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
NT_DEVINL void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
NT_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns.
NT_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
NT_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- math / conversion
NT_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
NT_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// 8 bf16 packed in a uint4 -> 8 floats (bf16 -> fp32 is a 16-bit shift)
NT_DEVINL void bf16x8_to_f32(const uint4& p, float (&f)[8]) {
  f[0] = __uint_as_float(p.x << 16);
  f[1] = __uint_as_float(p.x & 0xffff0000u);
  f[2] = __uint_as_float(p.y << 16);
  f[3] = __uint_as_float(p.y & 0xffff0000u);
  f[4] = __uint_as_float(p.z << 16);
  f[5] = __uint_as_float(p.z & 0xffff0000u);
  f[6] = __uint_as_float(p.w << 16);
  f[7] = __uint_as_float(p.w & 0xffff0000u);
}
NT_DEVINL float silu(float x) { return x / (1.0f + __expf(-x)); }
NT_DEVINL uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace nt

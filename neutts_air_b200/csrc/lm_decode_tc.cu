// Persistent tcgen05 decode kernel: ALL layers, the lm_head, the sampler and the whole multi-step decode loop in
// ONE cooperative launch of one CTA per SM, for every batch size 1..64.
//
//   GEMM phases (qkv | o_proj | gate/up | down | lm_head) run on the 5th-generation tensor cores:
//     * A operand = WEIGHTS, exactly as they lie in HBM (K-major): 128 rows x 64 k tiles (16 KB) fetched by 2-D TMA
//       (SWIZZLE_128B) through per-matrix tensor maps built once at nt_lm_create; one warp streams this CTA's tiles
//       of the WHOLE step into a deep mbarrier ring and runs ahead across phase boundaries (weights are immutable);
//     * B operand = ACTIVATIONS, K-major [tokens x 64 k] chunks in shared memory: tokens sit on the UMMA N axis
//       (N = 16 | 32 | 64).  Batch <= 8 feeds every activation as a bf16 hi + lo pair on two N columns (~16 mantissa
//       bits, the decode path keeps fp32-grade activations); larger batches use plain bf16 like the prefill path;
//     * accumulators in TMEM (128 lanes = weight rows, N fp32 columns), double-buffered: one elected thread issues
//       tcgen05.mma, four epilogue warps tcgen05.ld their 32 lanes and run the fused epilogues.
//   Work split: every weight matrix is cut into (128-row tile, K slice) items spread over the CTAs so that each
//   SM streams the same number of bytes per layer; matrices with few row tiles (qkv 9, o 7, down 7) split K and
//   write raw partial sums, folded IN SLICE ORDER by their consumer (bit-reproducible, no atomics).  gate/up keeps
//   K whole (SwiGLU is not linear) on its own set of CTAs.
//   Phases of a layer (grid barrier between them):
//       [fold+RMSNorm] qkv -> RoPE/KV-append + split-KV attention -> merge + o_proj -> [fold+RMSNorm] gate/up+SwiGLU
//       -> down
//     batch <= 4: the consumers fold the split-K slices and normalise while staging their B operand (5 barriers per
//     layer); larger batches: token-owner CTAs fold + normalise into bf16 rows that the consumers fetch by TMA
//     (7 barriers per layer, no per-CTA re-reading of the whole batch).
//   lm_head epilogue: logits -> HBM once, plus the processed maximum of every 128-row tile; the sampler then needs
//   only the top_k tiles with the largest maxima (provably a superset of the top-k logits), so selection costs
//   ~10 us on one CTA per sequence instead of a pass over the vocabulary.
//
// Replaces transformers generation/utils.py:2743-2805 + modeling_qwen2.py:280-309,353-413 for the decode loop
// (SURVEY.md §8a rows A1, A3-A12); supersedes the CUDA-core megakernel (lm_mega.cu) and the 196-launch chain.
#include "lm_device.cuh"
#include "lm_decode_tc.cuh"

#include <cuda.h>

#include <cstdlib>
#include <vector>

namespace nt {

// ------------------------------------------------------------------------------------------ small device helpers
NT_DEVINL unsigned tc_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
NT_DEVINL long long tc_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
NT_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
NT_DEVINL void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
NT_DEVINL void bar_epi() { asm volatile("bar.sync 2, 128;" ::: "memory"); }  // the four epilogue warps

struct TcProf {
  long long* buf;
  int n;
  NT_DEVINL void mark() {
    if (buf && n < 1024) buf[n++] = tc_ns();
  }
};

// bf16 hi/lo split of an fp32 value: hi = rn(x), lo = rn(x - hi); hi + lo carries ~16 mantissa bits
NT_DEVINL void split_hilo(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(x);
  lo = __float2bfloat16(x - __bfloat162float(hi));
}
// element (row n, k) of a K-major SWIZZLE_128B chunk [rows][64 bf16]: 16-byte group g = k / 8 sits at g ^ (n & 7)
NT_DEVINL __nv_bfloat16* chunk_elem(uint8_t* chunk, int n, int k) {
  return reinterpret_cast<__nv_bfloat16*>(chunk + n * 128 + ((((k >> 3) ^ (n & 7)) << 4) | ((k & 7) << 1)));
}

constexpr int kPhQ = 0, kPhO = 1, kPhG = 2, kPhD = 3;

// shared-memory misc block (after the ring and the union region)
struct TcMisc {
  uint64_t full_bar[16];
  uint64_t empty_bar[16];
  uint64_t bop_bar;
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  AttnSync attn;
  uint32_t tmem_slot;
  int go;                 // steps released to the stream / MMA warps so far, -1 = stop
  int pos[kTcMaxBatch];   // this step's seq_lens snapshot
  int mask_eos[kTcMaxBatch];
  float red[64];
  float tile_max[2][4][kTcMaxBatch];
  float rstd[8];
  int sel[8];
  TcPlan plan;
};

// Attention staging (fp32 CUDA-core path): one 64-token K page + V page, the group's queries, running softmax state.
struct TcAttnSmem {
  __nv_bfloat16 k[64 * 64];
  __nv_bfloat16 v[64 * 64];
  float q[8][64];
  float s[8][64];
  float ml[8][2];
  float corr[8];
  float red[4][8][64];
  float knew[64], vnew[64];
};

// ------------------------------------------------------------------------------------------ the kernel
template <int NT, bool HILO>
__global__ void __launch_bounds__(kTcThreads, 1) decode_tc_kernel(const __grid_constant__ TcParams P) {
  constexpr int CHUNK = NT * 128;            // bytes of one B-operand k-block
  constexpr int NTOK = HILO ? 8 : NT;        // token slots on the N axis
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;
  uint8_t* uni = smem + P.uni_off;
  TcMisc* ms = reinterpret_cast<TcMisc*>(smem + P.misc_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = P.nstages;
  const int L = P.n_layers;
  const int B = P.B;
  const int H = P.hidden;
  const int KBH = H >> 6;                    // k-blocks of a hidden-sized K
  const CUtensorMap* wmaps = reinterpret_cast<const CUtensorMap*>(P.wmaps);
  const CUtensorMap* xmap = reinterpret_cast<const CUtensorMap*>(P.xmap);
  const CUtensorMap* amap = reinterpret_cast<const CUtensorMap*>(P.amap);

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&ms->full_bar[s], 1);
      mbar_init(&ms->empty_bar[s], 1);
    }
    mbar_init(&ms->bop_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ms->acc_full[i], 1);
      mbar_init(&ms->acc_empty[i], 4);
    }
    mbar_init(&ms->attn.bar, 1);
    ms->attn.uses = 0;
    fence_barrier_init();
    ms->go = 1;
  }
  for (int i = tid; i < static_cast<int>(sizeof(TcPlan) / 4); i += kTcThreads)
    reinterpret_cast<int*>(&ms->plan)[i] = reinterpret_cast<const int*>(P.plan + blockIdx.x)[i];
  if (warp == 9) tmem_alloc(&ms->tmem_slot, 2 * NT < 32 ? 32 : 2 * NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ms->tmem_slot;
  const TcPlan& plan = ms->plan;
  const int n_head_tiles = plan.head_t1 - plan.head_t0;

  auto wait_go = [&](int step) -> bool {  // stream / MMA warps: released one step at a time (early-exit safety)
    uint32_t spins = 0;
    int go;
    while ((go = *reinterpret_cast<volatile int*>(&ms->go)) >= 0 && go <= step) {
      __nanosleep(64);
      if (++spins > (1u << 25)) {
        printf("neutts_b200: decode_tc step gate timed out (block %d, warp %d)\n", blockIdx.x, warp);
        __trap();
      }
    }
    return go >= 0;
  };

  if (warp == 8) {
    // ================================================================== weight stream (one thread)
    if (lane == 0) {
      int slot = 0;
      uint32_t par = 0;       // parity of the slot's NEXT completion of empty_bar that we must have seen
      bool wrapped = false;   // ring used at least once
      auto push = [&](const CUtensorMap* m, int kcol, int row) {
        if (wrapped) mbar_wait(&ms->empty_bar[slot], par ^ 1);
        mbar_arrive_expect_tx(&ms->full_bar[slot], 16384);
        tma_load_2d(ring + static_cast<size_t>(slot) * 16384, m, kcol, row, &ms->full_bar[slot]);
        if (++slot == NS) slot = 0, par ^= 1, wrapped = true;
      };
      for (int step = 0; step < P.n_steps; ++step) {
        if (!wait_go(step)) break;
        for (int l = 0; l < L; ++l)
          for (int ph = 0; ph < 4; ++ph) {
            const CUtensorMap* m = wmaps + 4 * l + ph;
            for (int i = 0; i < plan.n[ph]; ++i) {
              const TcItem it = plan.it[ph][i];
              for (int kb = 0; kb < it.nkb; ++kb) push(m, (it.kb0 + kb) * 64, it.tile * 128);
            }
          }
        const CUtensorMap* hm = wmaps + 4 * P.total_layers;
        for (int t = plan.head_t0; t < plan.head_t1; ++t)
          for (int kb = 0; kb < KBH; ++kb) push(hm, kb * 64, t * 128);
      }
    }
  } else if (warp == 9) {
    // ================================================================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(1, 128, NT);
      int slot = 0;
      uint32_t par = 0;
      uint32_t bop_n = 0, acc_n = 0;
      const uint32_t bop_addr = smem_u32(uni);
      // one item: nkb ring tiles against B chunks chunk0, chunk0 + 1, ...
      auto run_item = [&](int nkb, int chunk0) {
        const uint32_t buf = acc_n & 1;
        mbar_wait(&ms->acc_empty[buf], ((acc_n >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dst = tmem_base + buf * NT;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&ms->full_bar[slot], par);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(ring + static_cast<size_t>(slot) * 16384));
          const uint64_t bdesc = umma_desc_sw128(bop_addr + static_cast<uint32_t>(chunk0 + kb) * CHUNK);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(dst, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&ms->empty_bar[slot]);
          if (++slot == NS) slot = 0, par ^= 1;
        }
        umma_commit(&ms->acc_full[buf]);
        ++acc_n;
      };
      for (int step = 0; step < P.n_steps; ++step) {
        if (!wait_go(step)) break;
        for (int l = 0; l < L; ++l)
          for (int ph = 0; ph < 4; ++ph) {
            if (plan.n[ph] == 0) continue;
            mbar_wait(&ms->bop_bar, bop_n & 1);
            ++bop_n;
            tc_fence_after();
            int chunk = 0;
            for (int i = 0; i < plan.n[ph]; ++i) {
              const TcItem it = plan.it[ph][i];
              if (ph == kPhG) {
                run_item(it.nkb, 0);   // whole K, all items share the staged input
              } else {
                run_item(it.nkb, chunk);
                chunk += it.nkb;
              }
            }
          }
        if (n_head_tiles > 0) {
          mbar_wait(&ms->bop_bar, bop_n & 1);
          ++bop_n;
          tc_fence_after();
          for (int t = 0; t < n_head_tiles; ++t) run_item(KBH, 0);
        }
      }
    }
  } else {
    // ================================================================== worker warps 0..7
    const SyncConsumers csync;
    const unsigned G = gridDim.x;
    unsigned target = 0;
    uint32_t acc_n = 0;
    TcProf prof{nullptr, 0};
    const int n_rep = P.n_heads / P.n_kv;
    const int split_cap = P.split_cap;
    const bool fold_cta = P.fold_in_cta != 0;
    float* xf = reinterpret_cast<float*>(uni + 14 * CHUNK);   // fold_in_cta: fp32 folded rows [B][H] behind the B chunks
    TcAttnSmem* asmem = reinterpret_cast<TcAttnSmem*>(uni);

    // ---- grid barrier; `post` runs on thread 0 between the release and the trailing CTA barrier
    auto grid_sync = [&](auto post) {
      csync();
      if (tid == 0) {
        target += G;
        __threadfence();
        atomicAdd(P.gbar, 1u);
        uint32_t spins = 0;
        while (tc_ld_acquire(P.gbar) < target) {
          if (++spins > (1u << 24)) {
            printf("neutts_b200: decode_tc grid barrier timed out (block %d, target %u, seen %u)\n", blockIdx.x, target, *P.gbar);
            __trap();
          }
        }
        prof.mark();
        post();
      }
      csync();
    };
    auto no_post = [] {};

    // ---- B operand by TMA from global bf16 rows (thread 0, after the barrier that published them)
    auto load_bop_split = [&](const CUtensorMap* m, int ph) {   // chunks of the items' own k ranges, item after item
      const int n = plan.n[ph];
      if (n == 0) return;
      int total = 0;
      for (int i = 0; i < n; ++i) total += plan.it[ph][i].nkb;
      fence_proxy_async_all();
      mbar_arrive_expect_tx(&ms->bop_bar, static_cast<uint32_t>(total) * CHUNK);
      int c = 0;
      for (int i = 0; i < n; ++i)
        for (int kb = 0; kb < plan.it[ph][i].nkb; ++kb, ++c) tma_load_2d(uni + c * CHUNK, m, (plan.it[ph][i].kb0 + kb) * 64, 0, &ms->bop_bar);
    };
    auto load_bop_full = [&](const CUtensorMap* m, bool need) {  // all KBH chunks of the hidden-sized K
      if (!need) return;
      fence_proxy_async_all();
      mbar_arrive_expect_tx(&ms->bop_bar, static_cast<uint32_t>(KBH) * CHUNK);
      for (int kb = 0; kb < KBH; ++kb) tma_load_2d(uni + kb * CHUNK, m, kb * 64, 0, &ms->bop_bar);
    };
    // thread-staged B operand is complete: every writer fenced its writes towards the async proxy
    auto bop_ready = [&] {
      fence_proxy_async();
      csync();
      if (tid == 0) mbar_arrive(&ms->bop_bar);
    };

    // ---- accumulator of the next item -> registers (epilogue warps 0..3; lane = weight row of the tile)
    auto acc_take = [&](float (&v)[NT]) {
      const uint32_t buf = acc_n & 1;
      mbar_wait(&ms->acc_full[buf], (acc_n >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + buf * NT;
      if constexpr (NT == 16) {
        uint32_t r[16];
        tmem_ld16(taddr, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      } else {
#pragma unroll
        for (int c = 0; c < NT / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[c * 32 + j] = __uint_as_float(r[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->acc_empty[buf]);
      ++acc_n;
      if constexpr (HILO) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += v[j + 8];
      }
    };

    // ---- epilogue of the split-K phases: raw partial sums, one row per lane, coalesced over the warp
    auto epi_partials = [&](int ph, float* part, int rows) {
      if (warp >= 4) return;
      for (int i = 0; i < plan.n[ph]; ++i) {
        const TcItem it = plan.it[ph][i];
        float v[NT];
        acc_take(v);
        const int row = it.tile * 128 + warp * 32 + lane;
        if (row < rows) {
          float* dst = part + (static_cast<long long>(it.slice) * B) * rows + row;
#pragma unroll
          for (int n = 0; n < NTOK; ++n)
            if (n < B) dst[static_cast<long long>(n) * rows] = v[n];
        }
      }
    };

    // ---- fold + RMSNorm of ONE token row by its owner CTA -> residual stream (fp32) + normalised bf16 rows
    auto fold_phase = [&](const float* parts, int nparts, int rows, const float* norm_w) {
      const int b = blockIdx.x;
      if (b >= B) return;
      float* hb = P.h + static_cast<long long>(b) * H;
      float ss = 0.f;
      for (int i = tid; i < H; i += kConsumerThreads) {
        float v = __ldcg(hb + i);
        for (int s = 0; s < nparts; ++s) v += __ldcg(parts + (static_cast<long long>(s) * B + b) * rows + i);
        hb[i] = v;
        ss += v * v;
      }
      ss = warp_sum(ss);
      if (lane == 0) ms->red[warp] = ss;
      csync();
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) t += ms->red[w];
      const float sc = rsqrtf(t / static_cast<float>(H) + P.eps);
      for (int i = tid; i < H; i += kConsumerThreads) {
        const float xn = __ldg(norm_w + i) * (hb[i] * sc);   // hb[i]: this thread's own store above
        if constexpr (HILO) {
          __nv_bfloat16 hi, lo;
          split_hilo(xn, hi, lo);
          P.xa[static_cast<long long>(b) * H + i] = hi;
          P.xa[static_cast<long long>(8 + b) * H + i] = lo;
        } else {
          P.xa[static_cast<long long>(b) * H + i] = __float2bfloat16(xn);
        }
      }
    };

    // ---- batch <= 4: fold ALL rows in this CTA (h + slices, slice order), normalise, stage `nchunk` B chunks
    //      (k-blocks listed per item for split phases, or all of them) -- no fold phase, no barrier
    auto fold_stage = [&](const float* parts, int nparts, int rows, const float* norm_w, int ph /* -1: all k-blocks */) {
      for (int e = tid; e < B * H; e += kConsumerThreads) {
        const int b = e / H, i = e - b * H;
        float v = __ldcg(P.h + e);
        for (int s = 0; s < nparts; ++s) v += __ldcg(parts + (static_cast<long long>(s) * B + b) * rows + i);
        xf[e] = v;
      }
      csync();
      if (warp < B) {  // warp b: sum of squares of row b
        float ss = 0.f;
        for (int i = lane; i < H; i += 32) ss += xf[warp * H + i] * xf[warp * H + i];
        ss = warp_sum(ss);
        if (lane == 0) ms->rstd[warp] = rsqrtf(ss / static_cast<float>(H) + P.eps);
      }
      csync();
      auto stage_chunk = [&](int chunk, int kb) {
        uint8_t* cb = uni + chunk * CHUNK;
        for (int e = tid; e < B * 64; e += kConsumerThreads) {
          const int b = e >> 6, k = e & 63;
          const int i = kb * 64 + k;
          const float xn = __ldg(norm_w + i) * (xf[b * H + i] * ms->rstd[b]);
          __nv_bfloat16 hi, lo;
          split_hilo(xn, hi, lo);
          *chunk_elem(cb, b, k) = hi;
          *chunk_elem(cb, 8 + b, k) = lo;
        }
      };
      if (ph < 0) {
        for (int kb = 0; kb < KBH; ++kb) stage_chunk(kb, kb);
      } else {
        int c = 0;
        for (int i = 0; i < plan.n[ph]; ++i)
          for (int kb = 0; kb < plan.it[ph][i].nkb; ++kb, ++c) stage_chunk(c, plan.it[ph][i].kb0 + kb);
      }
      bop_ready();
    };
    auto write_back_h = [&] {  // the designated CTA publishes the folded residual stream (after the phase's barrier)
      for (int e = tid; e < B * H; e += kConsumerThreads) P.h[e] = xf[e];
      csync();  // xf lies in the union region: nobody may reuse it before every thread has read its part
    };

    // ---- attention item of this CTA: (sequence, kv head, split)
    const int per_b = P.n_kv * split_cap;
    const int my_b = blockIdx.x / per_b, my_kvh = (blockIdx.x % per_b) / split_cap, my_split = blockIdx.x % split_cap;

    auto attention_phase = [&](int l) {
      if (my_b >= B) return;
      const int pos = ms->pos[my_b];
      const SplitGeom geo = split_geom(pos, P.kv.max_ctx, split_cap);
      if (my_split >= geo.nsplit) return;
      const int b = my_b, kvh = my_kvh;
      const int p0 = my_split * geo.pps, p1 = min(p0 + geo.pps, geo.npages);
      const bool appends = pos < P.kv.max_ctx && (pos >> 6) >= p0 && (pos >> 6) < p1;
      // prologue: fold the qkv slices (slice order) + bias, RoPE; q of the group -> shared; new K/V row -> page
      const float* bias = P.bqkv[l];
      const int QN = P.qkv_n;
      const float* pq = P.part_q + static_cast<long long>(b) * QN;
      const long long sstride = static_cast<long long>(B) * QN;
      for (int idx = tid; idx < n_rep * 32 + 64; idx += kConsumerThreads) {
        int row0;
        const int which = idx < n_rep * 32 ? 0 : (idx < n_rep * 32 + 32 ? 1 : 2);
        const int i = idx & 31;
        if (which == 0) row0 = (kvh * n_rep + (idx >> 5)) * 64 + 2 * i;
        else if (which == 1) row0 = (P.n_heads + kvh) * 64 + 2 * i;
        else row0 = (P.n_heads + P.n_kv + kvh) * 64 + 2 * i;
        if (which != 0 && !appends) continue;
        float2 a = make_float2(0.f, 0.f);
        for (int s = 0; s < P.sq; ++s) {
          const float2 t = __ldcg(reinterpret_cast<const float2*>(pq + s * sstride + row0));
          a.x += t.x, a.y += t.y;
        }
        a.x += __ldg(bias + row0), a.y += __ldg(bias + row0 + 1);
        if (which == 2) {
          asmem->vnew[2 * i] = a.x, asmem->vnew[2 * i + 1] = a.y;
        } else {
          float sn, cs;
          sincosf(static_cast<float>(pos) * __ldg(P.inv_freq + i), &sn, &cs);
          const float lo = a.x * cs - a.y * sn, hi = a.y * cs + a.x * sn;   // rows (2i, 2i+1) = dims (i, i + 32)
          if (which == 0) asmem->q[idx >> 5][i] = lo, asmem->q[idx >> 5][i + 32] = hi;
          else asmem->knew[i] = lo, asmem->knew[i + 32] = hi;
        }
      }
      if (tid < 8) asmem->ml[tid][0] = -INFINITY, asmem->ml[tid][1] = 0.f;
      csync();
      if (appends && tid < 64) {  // the new token's K/V row joins the cache (bf16) for the steps to come
        const int page = __ldcg(P.kv.page_table + b * P.kv.max_pages_per_seq + (pos >> 6));
        P.kv.page_ptr(l, 0, page, kvh)[(pos & 63) * 64 + tid] = __float2bfloat16(asmem->knew[tid]);
        P.kv.page_ptr(l, 1, page, kvh)[(pos & 63) * 64 + tid] = __float2bfloat16(asmem->vnew[tid]);
      }
      float acc[8];
#pragma unroll
      for (int h = 0; h < 8; ++h) acc[h] = 0.f;
      AttnSync* sy = &ms->attn;
      for (int pg = p0; pg < p1; ++pg) {
        const uint32_t parity = sy->uses & 1;
        csync();  // previous page fully consumed; everyone has read `uses`
        if (tid == 0) {
          const int page = __ldcg(P.kv.page_table + b * P.kv.max_pages_per_seq + pg);
          fence_proxy_async_all();
          mbar_arrive_expect_tx(&sy->bar, 2 * 8192);
          bulk_g2s(asmem->k, P.kv.page_ptr(l, 0, page, kvh), 8192, &sy->bar);
          bulk_g2s(asmem->v, P.kv.page_ptr(l, 1, page, kvh), 8192, &sy->bar);
          sy->uses += 1;
        }
        mbar_wait(&sy->bar, parity);
        if (appends && pg == (pos >> 6)) {  // patch the staged page with the new row (the copy may predate our store)
          if (tid < 64) {
            asmem->k[(pos & 63) * 64 + tid] = __float2bfloat16(asmem->knew[tid]);
            asmem->v[(pos & 63) * 64 + tid] = __float2bfloat16(asmem->vnew[tid]);
          }
          csync();
        }
        {  // scores: thread = (token, quarter of the head dim)
          const int tok = tid >> 2, part = tid & 3;
          const uint4* kr = reinterpret_cast<const uint4*>(asmem->k + tok * 64 + part * 16);
          float kf[16];
          {
            float t[8];
            bf16x8_to_f32(kr[0], t);
#pragma unroll
            for (int j = 0; j < 8; ++j) kf[j] = t[j];
            bf16x8_to_f32(kr[1], t);
#pragma unroll
            for (int j = 0; j < 8; ++j) kf[8 + j] = t[j];
          }
          const bool valid = (pg * 64 + tok) < geo.n_ctx;
          for (int h = 0; h < n_rep; ++h) {
            float d = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) d += kf[j] * asmem->q[h][part * 16 + j];
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            if (part == 0) asmem->s[h][tok] = valid ? d * P.scale_log2 : -INFINITY;
          }
        }
        csync();
        if (warp < n_rep) {  // online softmax update of head `warp`
          const float s0 = asmem->s[warp][lane], s1 = asmem->s[warp][lane + 32];
          const float m_old = asmem->ml[warp][0];
          const float m_new = fmaxf(m_old, warp_max(fmaxf(s0, s1)));  // every page of a live split has a valid token
          const float p0v = exp2f(s0 - m_new), p1v = exp2f(s1 - m_new);
          const float lsum = warp_sum(p0v + p1v);
          asmem->s[warp][lane] = p0v;
          asmem->s[warp][lane + 32] = p1v;
          if (lane == 0) {
            const float c = exp2f(m_old - m_new);  // 0 on the first page (m_old = -inf)
            asmem->corr[warp] = c;
            asmem->ml[warp][0] = m_new;
            asmem->ml[warp][1] = asmem->ml[warp][1] * c + lsum;
          }
        }
        csync();
        {  // P.V : thread = (dim, token group of 16), accumulators carried across pages
          const int d = tid & 63, g = tid >> 6;
#pragma unroll
          for (int h = 0; h < 8; ++h)
            if (h < n_rep) acc[h] *= asmem->corr[h];
          for (int t = g * 16; t < g * 16 + 16; ++t) {
            const float v = __bfloat162float(asmem->v[t * 64 + d]);
#pragma unroll
            for (int h = 0; h < 8; ++h)
              if (h < n_rep) acc[h] += asmem->s[h][t] * v;
          }
        }
      }
      {
        const int d = tid & 63, g = tid >> 6;
#pragma unroll
        for (int h = 0; h < 8; ++h)
          if (h < n_rep) asmem->red[g][h][d] = acc[h];
      }
      csync();
      for (int i = tid; i < n_rep * 64; i += kConsumerThreads) {
        const int h = i >> 6, d = i & 63;
        const float o = asmem->red[0][h][d] + asmem->red[1][h][d] + asmem->red[2][h][d] + asmem->red[3][h][d];
        const long long hh = static_cast<long long>(b) * P.n_heads + kvh * n_rep + h;
        P.att_o[(hh * P.max_splits + my_split) * 64 + d] = o;
        if (d == 0) {
          P.att_ml[(hh * P.max_splits + my_split) * 2 + 0] = asmem->ml[h][0];
          P.att_ml[(hh * P.max_splits + my_split) * 2 + 1] = asmem->ml[h][1];
        }
      }
    };

    // ---- o_proj input: merge the split-KV partials of the heads this CTA's items need, straight into B chunks
    auto stage_attn = [&] {
      int c = 0;
      for (int i = 0; i < plan.n[kPhO]; ++i)
        for (int kb = 0; kb < plan.it[kPhO][i].nkb; ++kb, ++c) {
          const int head = plan.it[kPhO][i].kb0 + kb;
          uint8_t* cb = uni + c * CHUNK;
          for (int e = tid; e < B * 64; e += kConsumerThreads) {
            const int b = e >> 6, d = e & 63;
            const SplitGeom g = split_geom(ms->pos[b], P.kv.max_ctx, split_cap);
            const long long hh = static_cast<long long>(b) * P.n_heads + head;
            const float2* ml = reinterpret_cast<const float2*>(P.att_ml) + hh * P.max_splits;
            const float* po = P.att_o + hh * P.max_splits * 64 + d;
            float M = -INFINITY;
            for (int s = 0; s < g.nsplit; ++s) M = fmaxf(M, __ldcg(ml + s).x);
            float Ls = 0.f, O = 0.f;
            for (int s = 0; s < g.nsplit; ++s) {
              const float2 t = __ldcg(ml + s);
              const float wgt = exp2f(t.x - M);
              Ls += wgt * t.y;
              O += wgt * __ldcg(po + s * 64);
            }
            const float val = O / Ls;
            if constexpr (HILO) {
              __nv_bfloat16 hi, lo;
              split_hilo(val, hi, lo);
              *chunk_elem(cb, b, d) = hi;
              *chunk_elem(cb, 8 + b, d) = lo;
            } else {
              *chunk_elem(cb, b, d) = __float2bfloat16(val);
            }
          }
        }
      bop_ready();
    };

    // ---- gate/up epilogue: rows (2j, 2j+1) = (gate_j, up_j) on adjacent lanes -> act[b][j] = silu(gate) * up
    auto epi_swiglu = [&] {
      if (warp >= 4) return;
      const int I = P.inter;
      for (int i = 0; i < plan.n[kPhG]; ++i) {
        const TcItem it = plan.it[kPhG][i];
        float v[NT];
        acc_take(v);
        const int row = it.tile * 128 + warp * 32 + lane;
        const int j = row >> 1;
#pragma unroll
        for (int n = 0; n < NTOK; ++n) {
          const float up = __shfl_down_sync(0xffffffffu, v[n], 1);
          if (n < B && !(lane & 1) && j < I) {
            const float a = silu(v[n]) * up;
            if constexpr (HILO) {
              __nv_bfloat16 hi, lo;
              split_hilo(a, hi, lo);
              P.act[static_cast<long long>(n) * I + j] = hi;
              P.act[static_cast<long long>(8 + n) * I + j] = lo;
            } else {
              P.act[static_cast<long long>(n) * I + j] = __float2bfloat16(a);
            }
          }
        }
      }
    };

    // ---- lm_head epilogue: logits -> HBM, processed maximum of the 128-row tile per sequence
    auto epi_head = [&] {
      if (warp >= 4) return;
      const int V = P.vocab;
      const float inv_t = 1.0f / P.samp.sp.temperature;
      const int eos = P.samp.sp.eos_id;
      for (int t = 0; t < n_head_tiles; ++t) {
        const int tile = plan.head_t0 + t;
        float v[NT];
        acc_take(v);
        const int row = tile * 128 + warp * 32 + lane;
        const bool ok = row < V;
        float (*tm)[kTcMaxBatch] = ms->tile_max[t & 1];
#pragma unroll
        for (int n = 0; n < NTOK; ++n) {
          if (n < B) {   // uniform across the warp
            if (ok) P.logits[static_cast<long long>(n) * V + row] = v[n];
            float pv = (!ok || (ms->mask_eos[n] && row == eos)) ? -INFINITY : v[n] * inv_t;
            pv = warp_max(pv);
            if (lane == 0) tm[warp][n] = pv;
          }
        }
        bar_epi();
        if (tid < B) P.tmax[static_cast<long long>(tid) * P.ntiles + tile] = fmaxf(fmaxf(tm[0][tid], tm[1][tid]), fmaxf(tm[2][tid], tm[3][tid]));
        // tile_max is double-buffered by tile parity: the barrier of tile t+1 orders these reads before tile t+2's writes
      }
    };

    // ---- sampler for sequence b = blockIdx.x: top-k tiles by maximum -> their logits >= threshold -> exact top-k
    auto sample_phase = [&] {
      const int b = blockIdx.x;
      if (b >= B) return;
      const int nt = P.ntiles, V = P.vocab;
      uint32_t* scratch = reinterpret_cast<uint32_t*>(uni);
      int* tiles = reinterpret_cast<int*>(scratch + kSelScratch);   // [64] chosen tiles
      int* counts = tiles + 64;                                      // [64] candidates per chosen tile, then offsets
      Cand* win = reinterpret_cast<Cand*>(counts + 64);              // [2 * kTopKeep]
      int* s_tok = reinterpret_cast<int*>(win + 2 * kTopKeep);
      uint32_t* keys = reinterpret_cast<uint32_t*>(s_tok + 4);       // the rest of the union region
      const int key_cap = static_cast<int>((P.uni_bytes - (kSelScratch + 128 + 4) * 4 - 2 * kTopKeep * sizeof(Cand)) / 4);
      const float inv_t = 1.0f / P.samp.sp.temperature;
      const bool mask_eos = ms->mask_eos[b] != 0;
      const int eos = P.samp.sp.eos_id;
      for (int i = tid; i < nt; i += kConsumerThreads) keys[i] = f2key(__ldcg(P.tmax + static_cast<long long>(b) * nt + i));
      if (tid < 64) tiles[tid] = -1, counts[tid] = 0;
      csync();
      const int k = min(min(P.samp.sp.top_k, kTopKeep), nt);
      uint32_t thr;
      int take_eq;
      radix_select_kth(keys, nt, k, scratch, thr, take_eq, csync);
      // fewer tiles than top_k: the tile maxima bound nothing, every logit of every tile is a candidate
      const uint32_t cthr = nt < P.samp.sp.top_k ? 1u : thr;
      // the k tiles: maxima above the threshold, then the first take_eq tiles (index order) that equal it
      compact_topk(keys, nt, thr, take_eq, scratch, csync, [&](int slot, int i) { tiles[slot] = i; });
      csync();
      const float* lg = P.logits + static_cast<long long>(b) * V;
      auto tile_keys = [&](int tile, uint32_t (&kk)[4]) {   // this lane's 4 logits of the tile -> processed keys
        const int r0 = tile * 128 + lane * 4;
        float4 x = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (r0 + 3 < V) {
          x = __ldcg(reinterpret_cast<const float4*>(lg + r0));
        } else {
          if (r0 < V) x.x = __ldcg(lg + r0);
          if (r0 + 1 < V) x.y = __ldcg(lg + r0 + 1);
          if (r0 + 2 < V) x.z = __ldcg(lg + r0 + 2);
        }
        kk[0] = (r0 < V) ? processed_key(x.x, r0, mask_eos, eos, inv_t) : 0u;
        kk[1] = (r0 + 1 < V) ? processed_key(x.y, r0 + 1, mask_eos, eos, inv_t) : 0u;
        kk[2] = (r0 + 2 < V) ? processed_key(x.z, r0 + 2, mask_eos, eos, inv_t) : 0u;
        kk[3] = (r0 + 3 < V) ? processed_key(x.w, r0 + 3, mask_eos, eos, inv_t) : 0u;
      };
      // pass 1: candidates (key >= threshold) per chosen tile
      for (int j = warp; j < k; j += kConsumerWarps) {
        uint32_t kk[4];
        tile_keys(tiles[j], kk);
        int c = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) c += __popc(__ballot_sync(0xffffffffu, kk[q] >= cthr && kk[q] != 0u));
        if (lane == 0) counts[j] = c;
      }
      csync();
      if (warp == 0) {  // exclusive prefix over <= 64 tiles
        const int c0 = counts[lane], c1 = counts[lane + 32];
        int s0 = c0, s1 = c1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int a = __shfl_up_sync(0xffffffffu, s0, off), bb = __shfl_up_sync(0xffffffffu, s1, off);
          if (lane >= off) s0 += a, s1 += bb;
        }
        const int tot0 = __shfl_sync(0xffffffffu, s0, 31);
        const int tot = tot0 + __shfl_sync(0xffffffffu, s1, 31);
        counts[lane] = s0 - c0;
        counts[lane + 32] = tot0 + s1 - c1;
        if (lane == 0) ms->sel[0] = tot;
      }
      csync();
      const int ncand = min(min(ms->sel[0], key_cap), 256 * kTopKeep);   // beyond: thousands of exact ties at the threshold
      // pass 2: write the candidates at their deterministic offsets (layout sample_stage2_seq expects: [b * ncand + i])
      constexpr long long kCandPitch = 256 * kTopKeep;   // row pitch of the candidate arrays (sampler_scratch_floats)
      float* cv = P.samp.cand_val + static_cast<long long>(b) * kCandPitch;
      int32_t* ci = P.samp.cand_idx + static_cast<long long>(b) * kCandPitch;
      for (int j = warp; j < k; j += kConsumerWarps) {
        uint32_t kk[4];
        const int tile = tiles[j];
        tile_keys(tile, kk);
        int base = counts[j];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool hit = kk[q] >= cthr && kk[q] != 0u;
          const uint32_t m = __ballot_sync(0xffffffffu, hit);
          if (hit) {
            const int o = base + __popc(m & ((1u << lane) - 1u));
            if (o < ncand) {
              cv[o] = key2f(kk[q]);
              ci[o] = tile * 128 + lane * 4 + q;
            }
          }
          base += __popc(m);
        }
      }
      csync();
      sample_stage2_seq(P.samp, b, ncand, keys, scratch, win, s_tok, csync, NoMark(), kCandPitch);
    };

    // =============================================================== the decode loop
    for (int step = 0; step < P.n_steps; ++step) {
      if (P.prof && tid == 0 && step == P.prof_step && blockIdx.x == 0) {
        prof.buf = P.prof;
        prof.n = 0;
        prof.mark();
      } else {
        prof.buf = nullptr;
      }
      if (tid < B) {
        ms->pos[tid] = __ldcg(P.kv.seq_lens + tid);
        ms->mask_eos[tid] = __ldcg(P.samp.n_generated + tid) < P.samp.sp.min_new_tokens ? 1 : 0;
      }
      csync();
      if (!fold_cta) {
        fold_phase(nullptr, 0, H, L > 0 ? P.ln1[0] : P.final_norm);
        if (L > 0) grid_sync([&] { load_bop_split(xmap, kPhQ); });
        else grid_sync([&] { load_bop_full(xmap, n_head_tiles > 0); });
      }
      for (int l = 0; l < L; ++l) {
        // ---- qkv
        if (fold_cta && plan.n[kPhQ] > 0) fold_stage(l > 0 ? P.part_d : nullptr, l > 0 ? P.sd : 0, H, P.ln1[l], kPhQ);
        epi_partials(kPhQ, P.part_q, P.qkv_n);
        grid_sync(no_post);
        if (fold_cta && plan.fold_q && l > 0) write_back_h();
        // ---- attention
        attention_phase(l);
        grid_sync(no_post);
        // ---- o_proj
        if (plan.n[kPhO] > 0) stage_attn();
        epi_partials(kPhO, P.part_o, H);
        if (!fold_cta) {
          grid_sync(no_post);
          fold_phase(P.part_o, P.so, H, P.ln2[l]);
          grid_sync([&] { load_bop_full(xmap, plan.n[kPhG] > 0); });
        } else {
          grid_sync(no_post);
          if (plan.n[kPhG] > 0) fold_stage(P.part_o, P.so, H, P.ln2[l], -1);
        }
        // ---- gate/up + SwiGLU
        epi_swiglu();
        grid_sync([&] { load_bop_split(amap, kPhD); });
        if (fold_cta && plan.fold_g) write_back_h();
        // ---- down
        epi_partials(kPhD, P.part_d, H);
        if (!fold_cta) {
          grid_sync(no_post);
          const bool last = l + 1 == L;
          fold_phase(P.part_d, P.sd, H, last ? P.final_norm : P.ln1[l + 1]);
          if (last) grid_sync([&] { load_bop_full(xmap, n_head_tiles > 0); });
          else grid_sync([&] { load_bop_split(xmap, kPhQ); });
        } else {
          grid_sync(no_post);
        }
      }
      // ---- lm_head
      if (fold_cta && n_head_tiles > 0) fold_stage(L > 0 ? P.part_d : nullptr, L > 0 ? P.sd : 0, H, P.final_norm, -1);
      epi_head();
      grid_sync(no_post);
      // ---- sampler (+ tests: keep every step's logits)
      if (P.logits_out) {
        const long long n = static_cast<long long>(B) * P.vocab;
        float* dst = P.logits_out + static_cast<long long>(step) * P.logits_step_stride;
        for (long long i = static_cast<long long>(blockIdx.x) * kConsumerThreads + tid; i < n; i += static_cast<long long>(G) * kConsumerThreads)
          dst[i] = __ldcg(P.logits + i);
      }
      sample_phase();
      grid_sync(no_post);
      bool all_done = true;
      for (int b = 0; b < B; ++b) all_done = all_done && (__ldcg(P.samp.done + b) != 0);
      if (all_done || step + 1 == P.n_steps) break;
      if (tid == 0) *reinterpret_cast<volatile int*>(&ms->go) = step + 2;
    }
    csync();
    if (tid == 0) *reinterpret_cast<volatile int*>(&ms->go) = -1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 2 * NT < 32 ? 32 : 2 * NT);
}

// ------------------------------------------------------------------------------------------ host side
static size_t tc_union_bytes(int nt) { return size_t(14) * nt * 128 + (nt == 16 ? 16 * 1024 : 0); }

size_t tc_smem_bytes(int nt) {
  const size_t budget = 227 * 1024;
  const size_t misc = (sizeof(TcMisc) + 127) & ~size_t(127);
  const size_t fixed = tc_union_bytes(nt) + misc + 1024;
  const int ns = int((budget - fixed) / 16384);
  return size_t(ns > 16 ? 16 : ns) * 16384 + fixed;
}

int tc_build_plan(const TcShape& s, int G, TcPlan* plan, int* sq, int* so, int* sd, int* ntiles, int* max_split_chunks) {
  if (G < 8 || G > 256) return set_error(NT_ERR_INVALID, "decode_tc: %d SMs unsupported", G);
  if (s.hidden % 64 || s.inter % 64) return set_error(NT_ERR_INVALID, "decode_tc: hidden / inter must be multiples of 64");
  const int Tq = (s.qkv_n + 127) / 128, To = (s.hidden + 127) / 128, Tg = (2 * s.inter + 127) / 128;
  const int KBh = s.hidden / 64, KBo = s.n_heads, KBi = s.inter / 64;
  if (KBh > 14) return set_error(NT_ERR_INVALID, "decode_tc: hidden %d > 896 does not fit the shared-memory plan", s.hidden);
  for (int c = 0; c < G; ++c) plan[c] = TcPlan{};
  // gate/up keeps K whole: its row tiles go to a dedicated, evenly spread subset of the CTAs
  int ngu = Tg < G ? Tg : G;
  std::vector<int> gu, rest;
  if (G - ngu < 16) {  // too few CTAs would be left for the split phases: everybody does everything
    for (int c = 0; c < G; ++c) gu.push_back(c), rest.push_back(c);
    ngu = G;
  } else {
    for (int c = 0; c < G; ++c) {
      const bool is_gu = ((c + 1) * ngu) / G > (c * ngu) / G;
      (is_gu ? gu : rest).push_back(c);
    }
  }
  auto add = [&](int cta, int ph, TcItem it) -> bool {
    TcPlan& p = plan[cta];
    if (p.n[ph] >= kTcMaxItems) return false;
    p.it[ph][p.n[ph]++] = it;
    return true;
  };
  for (int t = 0; t < Tg; ++t)
    if (!add(gu[t % gu.size()], kPhG, TcItem{short(t), 0, short(KBh), 0})) return set_error(NT_ERR_INVALID, "decode_tc: too many gate/up tiles per CTA");
  const int nr = int(rest.size());
  int rot = 0;
  auto split_phase = [&](int ph, int T, int KB, int* slices) -> bool {
    int S = nr / T;
    if (S < 1) S = 1;
    if (S > KB) S = KB;
    if (S > kTcMaxSlices) S = kTcMaxSlices;
    *slices = S;
    for (int t = 0; t < T; ++t)
      for (int z = 0; z < S; ++z) {
        const int k0 = (KB * z) / S, k1 = (KB * (z + 1)) / S;
        if (!add(rest[rot % nr], ph, TcItem{short(t), short(k0), short(k1 - k0), short(z)})) return false;
        ++rot;
      }
    return true;
  };
  if (!split_phase(kPhQ, Tq, KBh, sq) || !split_phase(kPhO, To, KBo, so) || !split_phase(kPhD, To, KBi, sd))
    return set_error(NT_ERR_INVALID, "decode_tc: too many split-K items per CTA");
  int worst = 0;
  bool fq = false, fg = false;
  for (int c = 0; c < G; ++c) {
    for (int ph : {kPhQ, kPhO, kPhD}) {
      int chunks = 0;
      for (int i = 0; i < plan[c].n[ph]; ++i) chunks += plan[c].it[ph][i].nkb;
      if (chunks > worst) worst = chunks;
    }
    if (!fq && plan[c].n[kPhQ] > 0) plan[c].fold_q = 1, fq = true;
    if (!fg && plan[c].n[kPhG] > 0) plan[c].fold_g = 1, fg = true;
  }
  *max_split_chunks = worst;
  const int nt = (s.vocab + 127) / 128;
  *ntiles = nt;
  for (int c = 0; c < G; ++c) {
    plan[c].head_t0 = int((static_cast<long long>(nt) * c) / G);
    plan[c].head_t1 = int((static_cast<long long>(nt) * (c + 1)) / G);
  }
  return NT_OK;
}

template <int NT, bool HILO>
static int launch_tc(TcParams& P, int num_sms, cudaStream_t stream) {
  auto kern = decode_tc_kernel<NT, HILO>;
  const size_t smem = tc_smem_bytes(NT);
  const size_t misc = (sizeof(TcMisc) + 127) & ~size_t(127);
  P.nstages = int((smem - tc_union_bytes(NT) - misc - 1024) / 16384);
  P.uni_off = unsigned(size_t(P.nstages) * 16384);
  P.uni_bytes = unsigned(tc_union_bytes(NT));
  P.misc_off = P.uni_off + P.uni_bytes;
  NT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  NT_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTcThreads, smem));
  if (per_sm < 1) return set_error(NT_ERR_CUDA, "decode_tc: a CTA does not fit on an SM (%zu B shared memory)", smem);
  NT_CUDA_CHECK(cudaMemsetAsync(P.gbar, 0, sizeof(unsigned) * 64, stream));
  void* args[] = {&P};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(num_sms), dim3(kTcThreads), args, smem, stream);
  if (e != cudaSuccess) return set_error(NT_ERR_CUDA, "decode_tc launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return NT_OK;
}

int launch_decode_tc(TcParams& P, int B, int num_sms, int max_split_chunks, cudaStream_t stream) {
  if (B < 1 || B > kTcMaxBatch) return set_error(NT_ERR_INVALID, "decode_tc: batch %d not in 1..%d", B, kTcMaxBatch);
  if (P.n_heads % P.n_kv || P.n_heads / P.n_kv > 8) return set_error(NT_ERR_INVALID, "decode_tc: unsupported GQA ratio");
  P.B = B;
  // attention items: (sequence, kv head, split) -> one CTA each
  int cap = num_sms / (B * P.n_kv);
  if (cap < 1) return set_error(NT_ERR_INVALID, "decode_tc: %d sequences x %d kv heads exceed %d SMs", B, P.n_kv, num_sms);
  if (cap > 16) cap = 16;
  if (cap > P.max_splits) cap = P.max_splits;
  P.split_cap = cap;
  const int nt = B <= 16 ? 16 : (B <= 32 ? 32 : 64);
  if (max_split_chunks > 14) return set_error(NT_ERR_INVALID, "decode_tc: %d k-blocks per CTA exceed the staging area", max_split_chunks);
  if (sizeof(TcAttnSmem) > tc_union_bytes(nt)) return set_error(NT_ERR_INVALID, "decode_tc: attention staging does not fit");
  const bool hilo = B <= 8;
  const char* fe = getenv("NT_TC_FOLD");   // experiments: "phase" forces the fold phases at small batch
  P.fold_in_cta = (B <= 4 && size_t(B) * P.hidden * 4 <= 16 * 1024 && !(fe && fe[0] == 'p')) ? 1 : 0;
  if (hilo) return launch_tc<16, true>(P, num_sms, stream);
  if (nt == 16) return launch_tc<16, false>(P, num_sms, stream);
  if (nt == 32) return launch_tc<32, false>(P, num_sms, stream);
  return launch_tc<64, false>(P, num_sms, stream);
}

}  // namespace nt

// Persistent decode megakernel (batch <= 4): ALL layers, the lm_head, the sampler and the whole
// multi-step decode loop run in one cooperative launch of one CTA per SM.
//
//   * one producer warp per CTA streams this CTA's slice of every weight matrix, in phase order,
//     HBM -> shared memory with cp.async.bulk into a deep mbarrier ring (~170 KB in flight per SM);
//     weights are immutable, so the stream runs ahead across phase boundaries and hides the grid
//     barriers;
//   * eight consumer warps execute the phases (fused RMSNorm + QKV GEMV + bias + RoPE + KV-page
//     append | split-KV GQA attention | o_proj + residual | RMSNorm + gate/up + SiLU*up | down +
//     residual | lm_head | radix-select top-k + multinomial draw + next embedding) and meet at a
//     grid-wide barrier (one L2 atomic + acquire spin) between dependent phases;
//   * no host involvement between tokens: stop flags, KV lengths and sampled ids live on the device.
//
// Replaces the per-token loop of transformers generation/utils.py:2743-2805 (one host sync per
// step there) and the ~25 ATen launches per layer listed in SURVEY.md §2.1.
#include "lm_device.cuh"
#include "lm_mega.cuh"

#include <cstdlib>

namespace nt {

NT_DEVINL unsigned ld_relaxed_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
NT_DEVINL unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

NT_DEVINL long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// optional in-kernel timeline (thread 0 of the first and the last CTA, one chosen step)
struct Prof {
  long long* buf;
  int n;
  NT_DEVINL void mark() {
    if (buf && n < kProfMarks) buf[n++] = global_ns();
  }
};

// Grid-wide barrier over the consumer warps of all CTAs: one relaxed L2 atomic per CTA on a
// monotonically growing counter (release fence before it), then thread 0 spins with acquire loads
// until the counter reaches this barrier's target.  Measured on B200 (profiles/mega_timeline_r1.md):
// ~0.5 us to publish + ~1.3-2.6 us until the slowest CTA has arrived.  A flag-array variant (one word
// per CTA, coalesced polling, no atomics) was tried and is 2x slower: 148 pollers hammering the flag
// lines delay the flag stores themselves.
NT_DEVINL void grid_sync(unsigned* gbar, unsigned& target, unsigned nblocks, Prof& prof) {
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (threadIdx.x == 0) {
    prof.mark();  // all consumer warps of this CTA are done
    target += nblocks;
    __threadfence();
    atomicAdd(gbar, 1u);
    prof.mark();  // arrival published
    uint32_t spins = 0;
    while (ld_acquire_gpu(gbar) < target) {
      if (++spins > (1u << 23)) {
        printf("neutts_b200: grid barrier timed out (block %d, target %u, seen %u)\n", blockIdx.x, target, *gbar);
        __trap();
      }
    }
    prof.mark();  // barrier released
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

NT_DEVINL void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
NT_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
NT_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int NB>
__global__ void __launch_bounds__((kConsumerWarps + 1) * 32, 1) decode_mega_kernel(const MegaParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* ring = smem + P.ring_off;
  __nv_bfloat16* xsb = reinterpret_cast<__nv_bfloat16*>(smem + P.x_off);  // bf16 hi(/lo) activation rows
  uint8_t* uni = smem + P.union_off;  // AttnSmem | head-phase logits + selector scratch ; x+uni together: final selection
  uint8_t* zero16 = smem + P.misc_off;                                   // 16 zero bytes: padding rows of the MMA A operand
  float2* red2 = reinterpret_cast<float2*>(zero16 + 128);                // [2 parity][2 tiles][8 warps][32 lanes]
  float* s_part = reinterpret_cast<float*>(red2 + 2 * 2 * kConsumerWarps * 32);  // [8 warps][8 rows]
  int* pos_cache = reinterpret_cast<int*>(s_part + 64);
  int* page_cache = pos_cache + 8;
  float* norm_buf = reinterpret_cast<float*>(page_cache + 8);
  float* bias_buf = norm_buf + P.hidden;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P.bar_off);
  uint64_t* empty_bar = full_bar + 8;
  AttnSync* async_ = reinterpret_cast<AttnSync*>(empty_bar + 8);
  volatile int* s_go = reinterpret_cast<volatile int*>(async_ + 1);  // producer gate: steps released so far, -1 = stop

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = P.nstages;
  const int L = P.n_layers;
  const int n_phases = 4 * L + 1;
  AttnSmem* asmem = reinterpret_cast<AttnSmem*>(uni);

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    mbar_init(&async_->bar, 1);
    async_->uses = 0;
    fence_barrier_init();
    *s_go = 1;
  }
  if (tid < 4) reinterpret_cast<uint32_t*>(zero16)[tid] = 0u;
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ============================================================ producer warp
    if (lane == 0) {
      uint32_t g = 0;  // global stage counter (ring position)
      for (int step = 0; step < P.n_steps; ++step) {
        uint32_t spins = 0;
        int go;
        while ((go = *s_go) >= 0 && go <= step) {  // released one step at a time (early exit safety)
          __nanosleep(64);
          if (++spins > (1u << 25)) {
            printf("neutts_b200: producer gate timed out (block %d)\n", blockIdx.x);
            __trap();
          }
        }
        if (go < 0) break;
        for (int ph = 0; ph < n_phases; ++ph) {
          const int pidx = (ph == n_phases - 1) ? (4 * P.total_layers) : ph;  // lm_head is the last table entry
          if (P.l2_prefetch && (ph & 3) == 0 && ph < 4 * L) {
            // entering layer l: pull this CTA's rows of layer l+1 (or the head of the lm_head slice) into L2
            const int l = ph >> 2;
            if (l + 1 < L) {
              for (int q = 0; q < 4; ++q) {
                const MegaPhase np = P.phases[4 * (l + 1) + q];
                const TileGeom ng = tile_geom(np.rows, np.K);
                if (ng.r_end > ng.r_begin)
                  bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(np.W) + static_cast<long long>(ng.r_begin) * np.K * 2,
                                   static_cast<uint32_t>(ng.r_end - ng.r_begin) * np.K * 2);
              }
            }
            if (l + 2 >= L) {
              const MegaPhase hp = P.phases[4 * P.total_layers];
              const TileGeom hg = tile_geom(hp.rows, hp.K);
              const long long total = static_cast<long long>(hg.r_end - hg.r_begin) * hp.K * 2;
              const long long chunk = P.l2_head_bytes / 2;
              const long long off = (l + 2 == L) ? 0 : chunk;
              if (off < total)
                bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(hp.W) + static_cast<long long>(hg.r_begin) * hp.K * 2 + off,
                                 static_cast<uint32_t>(min(chunk, total - off)) & ~15u);
            }
          }
          const MegaPhase mp = P.phases[pidx];
          const TileGeom tg = tile_geom(mp.rows, mp.K);
          const uint32_t row_bytes = static_cast<uint32_t>(tg.kc) * 2u;
          for (int rb = 0; rb < tg.nrb; ++rb) {
            const int row0 = tg.r_begin + rb * tg.rps;
            const int nrows = min(tg.rps, tg.r_end - row0);
            for (int kc = 0; kc < tg.nkc; ++kc, ++g) {
              const int slot = g % NS;
              if (g >= static_cast<uint32_t>(NS)) mbar_wait(&empty_bar[slot], ((g / NS) - 1) & 1);
              mbar_arrive_expect_tx(&full_bar[slot], row_bytes * nrows);
              uint8_t* dst = ring + static_cast<size_t>(slot) * P.stage_bytes;
              const uint8_t* src = reinterpret_cast<const uint8_t*>(mp.W) + (static_cast<long long>(row0) * mp.K + kc * tg.kc) * 2;
              for (int i = 0; i < nrows; ++i)  // one bulk copy per row: rows land 16 bytes apart from a 128-byte multiple
                bulk_g2s(dst + i * (row_bytes + 16), src + static_cast<long long>(i) * mp.K * 2, row_bytes, &full_bar[slot]);
            }
          }
        }
      }
    }
    return;
  }

  // ============================================================== consumer warps
  uint32_t g = 0;
  unsigned target = 0;
  const unsigned G = gridDim.x;
  Prof prof{nullptr, 0};
  const SyncConsumers csync;
  const int H = P.hidden, I = P.inter, HD = P.n_heads * 64;
  const int n_rep = P.n_heads / P.kv.n_kv_heads;
  const int split_cap = P.split_cap;
  const uint32_t xs_addr = smem_u32(xsb), zero_addr = smem_u32(zero16);

  // this CTA's slice of every layer's QKV bias (immutable) -> shared memory, once
  const TileGeom qkv_g = tile_geom(P.qkv_n, H);
  const int bias_rows = qkv_g.r_end - qkv_g.r_begin;
  const bool bias_cached = bias_rows <= P.bias_cap;
  if (bias_cached)
    for (int i = tid; i < L * bias_rows; i += kConsumerThreads) {
      const int l = i / bias_rows, r = i - l * bias_rows;
      bias_buf[l * P.bias_cap + r] = __ldg(P.bqkv[l] + qkv_g.r_begin + r);
    }

  auto prefetch_norm = [&](const float* w) {  // asynchronous global -> shared copy of one RMSNorm weight vector
    if (tid < (H >> 2)) cp_async16(norm_buf + 4 * tid, w + 4 * tid);
    cp_async_commit();
  };

  // one GEMV phase on the tensor cores: stages of (row block x K chunk), 8 warps split K, partial
  // accumulators meet in shared memory once per row block, warps 0/1 run the fused epilogues
  auto run_phase = [&](const GemvParams& gp, const TileGeom& tg, int K) {
    if (tid == 0) prof.mark();  // input vector staged
    const int xstride_b = (K + 8) * 2;
    int stage_no = 0;
    for (int rb = 0; rb < tg.nrb; ++rb) {
      float acc[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
      const int row0 = tg.r_begin + rb * tg.rps;
      const int nrows = min(tg.rps, tg.r_end - row0);
      const int ntiles = (nrows + 7) >> 3;
      for (int kc = 0; kc < tg.nkc; ++kc, ++g, ++stage_no) {
        const int slot = g % NS;
        mbar_wait(&full_bar[slot], (g / NS) & 1);
        if (tid == 0 && (stage_no == 0 || stage_no == tg.stages - 1)) prof.mark();  // first / last stage landed
        mma_consume_stage<NB>(smem_u32(ring + static_cast<size_t>(slot) * P.stage_bytes), xs_addr, zero_addr, xstride_b, tg.kc, kc * tg.kc,
                              ntiles, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
      }
      float2* rbuf = red2 + (rb & 1) * (2 * kConsumerWarps * 32);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (j < ntiles) rbuf[(j * kConsumerWarps + warp) * 32 + lane] = make_float2(acc[j][0], acc[j][1]);
      csync();
      if (warp < ntiles) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) {
          const float2 v = rbuf[(warp * kConsumerWarps + w) * 32 + lane];
          a0 += v.x, a1 += v.y;
        }
        const int b = lane >> 2, row = row0 + warp * 8 + 2 * (lane & 3);
        if (b < NB && row < tg.r_end) unit_epilogue(gp, row >> 1, b, a0, a1);
      }
    }
  };

  AttnDecParams ap;
  ap.q = P.q, ap.kv = P.kv, ap.n_heads = P.n_heads, ap.n_rep = n_rep, ap.scale_log2 = P.scale_log2;
  ap.part_o = P.part_o, ap.part_ml = P.part_ml, ap.counters = P.counters, ap.out = P.attn, ap.out_bf16 = nullptr;
  ap.max_splits = P.max_splits, ap.layer = 0;
  // the attention item of this CTA (at most one): (sequence, kv head, split)
  const int per_b = P.kv.n_kv_heads * split_cap;
  const int my_b = blockIdx.x / per_b, my_kvh = (blockIdx.x % per_b) / split_cap, my_split = blockIdx.x % split_cap;

  for (int step = 0; step < P.n_steps; ++step) {
    if (P.prof && tid == 0 && step == P.prof_step && (blockIdx.x == 0 || blockIdx.x == G - 1)) {
      prof.buf = P.prof + (blockIdx.x == 0 ? 0 : kProfMarks);
      prof.n = 0;
      prof.mark();
    } else {
      prof.buf = nullptr;
    }
    // per-step snapshot of the sequence lengths and of the page that receives the new token
    if (tid < NB) {
      const int pos = __ldcg(P.kv.seq_lens + tid);
      pos_cache[tid] = pos;
      page_cache[tid] = (pos < P.kv.max_ctx) ? __ldcg(P.kv.page_table + tid * P.kv.max_pages_per_seq + (pos >> 6)) : 0;
    }
    prefetch_norm(P.ln1[0]);
    csync();
    SplitGeom geo{0, 0, 1, 0};
    const bool has_item = my_b < NB;
    if (has_item) geo = split_geom(pos_cache[my_b], P.kv.max_ctx, split_cap);
    const bool item_live = has_item && my_split < geo.nsplit;
    const bool can_prefetch_kv = item_live && (my_split * geo.pps < geo.npages - 1);  // first page is not the one being appended to

    for (int l = 0; l < L; ++l) {
      GemvParams gp;
      // ---- QKV: fused RMSNorm + bias + RoPE + KV-page append
      gp = GemvParams{};
      gp.rows = P.qkv_n, gp.K = H, gp.eps = P.eps, gp.bias = P.bqkv[l];
      gp.epi = GEMV_QKV_ROPE, gp.q_out = P.q, gp.kv = P.kv, gp.layer = l, gp.n_heads = P.n_heads, gp.inv_freq = P.inv_freq;
      gp.pos_cache = pos_cache, gp.page_cache = page_cache;
      if (bias_cached) gp.bias_smem = bias_buf + l * P.bias_cap, gp.row0 = qkv_g.r_begin;
      cp_async_wait_all();
      stage_x_bf16<NB>(P.h, H, H, norm_buf, P.eps, xsb, s_part, csync);
      run_phase(gp, qkv_g, H);
      ap.layer = l;
      if (can_prefetch_kv && tid == 0) attn_issue_page(ap, my_b, my_kvh, my_split * geo.pps, asmem, async_);
      grid_sync(P.gbar, target, G, prof);
      // ---- split-KV attention: partial (m, l, o) per split, merged by the consumers of the output
      prefetch_norm(P.ln2[l]);  // lands long before the barrier's release fence
      if (item_live) attn_split_item(ap, my_b, my_kvh, my_split, geo.pps, geo.npages, geo.n_ctx, asmem, async_, can_prefetch_kv, csync);
      grid_sync(P.gbar, target, G, prof);
      // ---- o_proj + residual (input = merged attention output)
      gp = GemvParams{};
      gp.rows = H, gp.K = HD, gp.epi = GEMV_STORE, gp.out = P.h, gp.ldo = H, gp.residual = P.h, gp.ldr = H;
      load_attn_merged_bf16<NB>(ap, pos_cache, split_cap, xsb, reinterpret_cast<float*>(uni), csync);
      run_phase(gp, tile_geom(H, HD), HD);
      grid_sync(P.gbar, target, G, prof);
      // ---- RMSNorm + gate/up + SiLU*up
      gp = GemvParams{};
      gp.rows = 2 * I, gp.K = H, gp.eps = P.eps, gp.epi = GEMV_SWIGLU, gp.out = P.act, gp.ldo = I;
      cp_async_wait_all();
      stage_x_bf16<NB>(P.h, H, H, norm_buf, P.eps, xsb, s_part, csync);
      prefetch_norm(l + 1 < L ? P.ln1[l + 1] : P.final_norm);  // norm_buf is free again: every thread passed the staging barrier
      run_phase(gp, tile_geom(2 * I, H), H);
      grid_sync(P.gbar, target, G, prof);
      // ---- down + residual
      gp = GemvParams{};
      gp.rows = H, gp.K = I, gp.epi = GEMV_STORE, gp.out = P.h, gp.ldo = H, gp.residual = P.h, gp.ldr = H;
      stage_x_bf16<NB>(P.act, I, I, nullptr, 0.f, xsb, s_part, csync);
      run_phase(gp, tile_geom(H, I), I);
      grid_sync(P.gbar, target, G, prof);
    }
    // ---- lm_head (fused final RMSNorm); the CTA keeps its own logits in shared memory and selects its
    //      local top-64 per sequence right away (no second pass over the logits, no extra barrier)
    {
      const TileGeom hg = tile_geom(P.vocab, H);
      float* lsm = reinterpret_cast<float*>(uni);
      GemvParams gp{};
      gp.rows = P.vocab, gp.K = H, gp.eps = P.eps, gp.epi = GEMV_STORE, gp.out = P.logits, gp.ldo = P.vocab;
      gp.smem_out = lsm, gp.smem_ld = P.head_ld, gp.row0 = hg.r_begin;
      cp_async_wait_all();
      stage_x_bf16<NB>(P.h, H, H, norm_buf, P.eps, xsb, s_part, csync);
      run_phase(gp, hg, H);
      csync();
      uint32_t* scratch = reinterpret_cast<uint32_t*>(lsm + NB * P.head_ld);
      const int n_local = hg.r_end - hg.r_begin;
      const float inv_t = 1.0f / P.samp.sp.temperature;
#pragma unroll 1
      for (int b = 0; b < NB; ++b) {
        const bool mask_eos = __ldcg(P.samp.n_generated + b) < P.samp.sp.min_new_tokens;
        uint32_t* keys = reinterpret_cast<uint32_t*>(lsm + b * P.head_ld);
        for (int e = tid; e < n_local; e += kConsumerThreads)
          keys[e] = processed_key(lsm[b * P.head_ld + e], gp.row0 + e, mask_eos, P.samp.sp.eos_id, inv_t);
        csync();
        emit_local_topk(P.samp, keys, n_local, gp.row0, (static_cast<long long>(b) * G + blockIdx.x) * kTopKeep, scratch, csync);
      }
    }
    grid_sync(P.gbar, target, G, prof);
    if (P.logits_out) {  // tests: keep every step's logits
      const long long n = static_cast<long long>(NB) * P.vocab;
      float* dst = P.logits_out + static_cast<long long>(step) * n;
      for (long long i = static_cast<long long>(blockIdx.x) * kConsumerThreads + tid; i < n; i += static_cast<long long>(G) * kConsumerThreads)
        dst[i] = __ldcg(P.logits + i);
    }
    // ---- final selection: CTA b finishes sequence b (top-k over G*64 candidates, softmax, draw, state, next embedding)
    if (static_cast<int>(blockIdx.x) < NB) {
      const int ncand = static_cast<int>(G) * kTopKeep;
      uint32_t* keys = reinterpret_cast<uint32_t*>(xsb);  // x rows + union region are contiguous and idle here
      uint32_t* scratch = keys + ncand;
      Cand* win = reinterpret_cast<Cand*>(scratch + kSelScratch);
      int* s_tok = reinterpret_cast<int*>(win + 2 * kTopKeep);
      sample_stage2_seq(P.samp, blockIdx.x, ncand, keys, scratch, win, s_tok, csync, [&]() {
        if (tid == 0) prof.mark();
      });
    }
    grid_sync(P.gbar, target, G, prof);
    // ---- stop when every sequence is finished (same decision in every CTA: flags were published before the barrier)
    bool all_done = true;
    for (int b = 0; b < NB; ++b) all_done = all_done && (__ldcg(P.samp.done + b) != 0);
    if (all_done || step + 1 == P.n_steps) break;
    if (tid == 0) *s_go = step + 2;  // release the producer into the next step
  }
  cp_async_wait_all();
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (tid == 0) *s_go = -1;
}

bool mega_supported(int hidden, int inter, int n_heads, int batch) {
  auto ok = [](int K) {
    const int nkc = (K <= 1024) ? 1 : (K + 1535) / 1536;
    return K % (32 * nkc) == 0;
  };
  return batch >= 1 && batch <= 8 && ok(hidden) && ok(inter) && ok(n_heads * 64);
}

int launch_decode_mega(MegaParams& P, int nb, int num_sms, cudaStream_t stream) {
  if (!mega_supported(P.hidden, P.inter, P.n_heads, nb)) return set_error(NT_ERR_INVALID, "megakernel: unsupported shape / batch %d", nb);
  if (num_sms > 256) num_sms = 256;
  // ---- shared-memory plan
  const int HD = P.n_heads * 64;
  auto stage_of = [](int K) {
    const int nkc = (K <= 1024) ? 1 : (K + 1535) / 1536;
    return (nkc == 1 ? 16 : 8) * ((K / nkc) * 2 + 16);
  };
  int stage = stage_of(P.hidden);
  if (stage_of(P.inter) > stage) stage = stage_of(P.inter);
  if (stage_of(HD) > stage) stage = stage_of(HD);
  stage = (stage + 127) & ~127;
  const int parts = nb <= 4 ? 2 : 1;
  int kmax = P.hidden > P.inter ? P.hidden : P.inter;
  if (HD > kmax) kmax = HD;
  size_t x_bytes = size_t(parts) * nb * (kmax + 8) * 2;
  const size_t x_norm = size_t(parts) * nb * (P.hidden + 8) * 2 + size_t(nb) * P.hidden * 4;  // bf16 rows + fp32 scratch of a norm phase
  if (x_norm > x_bytes) x_bytes = x_norm;
  x_bytes = (x_bytes + 127) & ~size_t(127);
  // union region: attention staging | head-phase logits (nb rows) + selector scratch | merge weights
  P.head_ld = 2 * ((P.vocab / 2 + num_sms - 1) / num_sms) + 8;
  size_t uni = size_t(nb) * P.head_ld * 4 + kSelScratch * 4 + 64;
  if (sizeof(AttnSmem) > uni) uni = sizeof(AttnSmem);
  uni = (uni + 127) & ~size_t(127);
  // the final selection needs G*64 keys + scratch + winners inside x + union
  const size_t final_need = size_t(num_sms) * kTopKeep * 4 + kSelScratch * 4 + 2 * kTopKeep * sizeof(Cand) + 64;
  if (x_bytes + uni < final_need) uni = ((final_need - x_bytes) + 127) & ~size_t(127);
  P.bias_cap = 2 * ((P.qkv_n / 2 + num_sms - 1) / num_sms) + 2;
  if (P.bias_cap > 64) P.bias_cap = 0;  // huge slices: read the bias from global memory instead
  const size_t misc = (128 + 2 * 2 * kConsumerWarps * 32 * sizeof(float2) + 64 * sizeof(float) + 16 * sizeof(int) + size_t(P.hidden) * 4 +
                       size_t(P.total_layers) * P.bias_cap * 4 + 127) & ~size_t(127);
  const size_t bars = (16 * sizeof(uint64_t) + sizeof(AttnSync) + 16 + 127) & ~size_t(127);
  const size_t fixed = x_bytes + uni + misc + bars + 128;
  const size_t budget = 227 * 1024;
  if (fixed + 2 * size_t(stage) > budget) return set_error(NT_ERR_INVALID, "megakernel: model does not fit the shared-memory plan");
  int ns = int((budget - fixed) / stage);
  if (ns > 8) ns = 8;
  P.nstages = ns;
  P.stage_bytes = stage;
  P.split_cap = 16 / nb > 0 ? 16 / nb : 1;
  if (P.split_cap > P.max_splits) P.split_cap = P.max_splits;
  if (nb * P.kv.n_kv_heads * P.split_cap > num_sms) return set_error(NT_ERR_INVALID, "megakernel: too few SMs for the attention items");
  P.l2_prefetch = getenv("NT_NO_L2_PREFETCH") ? 0 : 1;
  P.l2_head_bytes = 256 * 1024;
  size_t off = 0;
  P.ring_off = off, off += size_t(ns) * stage;
  P.x_off = off, off += x_bytes;
  P.union_off = off, off += uni;
  P.misc_off = off, off += misc;
  P.bar_off = off, off += bars;
  const size_t smem = off + 128;
  P.samp.nchunks = num_sms;  // candidate arrays are indexed [sequence][CTA][64] in this path

  void (*kern)(const MegaParams) = nullptr;
  switch (nb) {
    case 1: kern = decode_mega_kernel<1>; break;
    case 2: kern = decode_mega_kernel<2>; break;
    case 3: kern = decode_mega_kernel<3>; break;
    case 4: kern = decode_mega_kernel<4>; break;
    case 5: kern = decode_mega_kernel<5>; break;
    case 6: kern = decode_mega_kernel<6>; break;
    case 7: kern = decode_mega_kernel<7>; break;
    default: kern = decode_mega_kernel<8>; break;
  }
  static size_t attr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (attr[nb] < smem) {
    NT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    attr[nb] = smem;
  }
  int per_sm = 0;
  NT_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (kConsumerWarps + 1) * 32, smem));
  if (per_sm < 1) return set_error(NT_ERR_CUDA, "megakernel: a CTA does not fit on an SM (%zu B shared memory)", smem);
  NT_CUDA_CHECK(cudaMemsetAsync(P.gbar, 0, sizeof(unsigned) * 256, stream));
  void* args[] = {&P};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(num_sms), dim3((kConsumerWarps + 1) * 32), args, smem, stream);
  if (e != cudaSuccess) return set_error(NT_ERR_CUDA, "megakernel launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return NT_OK;
}

}  // namespace nt

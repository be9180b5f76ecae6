// Speech-LM kernels for sm_100a.
//
// Decode (memory-bound, batch <= 4 on CUDA cores):
//   gemv_kernel        weight rows stream HBM -> shared memory through a cp.async.bulk (TMA
//                      engine) ring guarded by mbarriers; 8 consumer warps + 1 producer warp.
//                      Fused prologue: RMSNorm of the input vector (modeling_qwen2.py:258-263).
//                      Fused epilogues: bias + RoPE + KV-page append (:217-225), residual add
//                      (:302,:308), SiLU(gate)*up (:46-48).  The weight prefetch is issued
//                      BEFORE griddepcontrol.wait so it overlaps the previous kernel (PDL).
//   attn_decode_kernel split-KV GQA attention over 64-token pages staged by bulk copies;
//                      fp32 softmax (:161-183); last-arriving CTA merges the splits.
//   topk kernels       min-new-tokens EOS mask, temperature, top-k, softmax, multinomial
//                      (logits_process.py:224-233,296-299,580-586; utils.py:2789-2791).
// Prefill helpers (the GEMMs go through gemm_tc.cu): embedding gather, RMSNorm rows,
// RoPE + KV append, causal GQA attention.
#include "lm_kernels.cuh"

#include <cfloat>

namespace nt {

// =================================================================================== GEMV
constexpr int kGemvConsumerWarps = 8;
constexpr int kGemvThreads = (kGemvConsumerWarps + 1) * 32;

struct GemvSmemPlan {
  int stage_bytes, nstages, units_per_stage, wpu;
  size_t ring_off, x_off, bar_off, red_off, total;
};

static GemvSmemPlan gemv_plan(int K, int nb) {
  GemvSmemPlan p;
  const int unit_bytes = 4 * K;  // two bf16 rows
  p.wpu = (K >= 2048) ? kGemvConsumerWarps : 1;
  p.units_per_stage = (p.wpu == 1) ? kGemvConsumerWarps : 1;
  p.stage_bytes = unit_bytes * p.units_per_stage;
  p.nstages = (p.wpu == 1) ? 3 : 4;
  size_t off = 0;
  p.ring_off = off;
  off += size_t(p.stage_bytes) * p.nstages;
  p.x_off = off;
  off += size_t(nb) * K * 4;
  p.red_off = off;
  off += 2 * kGemvConsumerWarps * 2 * 4 * sizeof(float);  // [parity][warp][row][nb<=4]
  p.bar_off = off;
  off += 2 * 8 * sizeof(uint64_t) + 64;
  p.total = off + 128;  // alignment slack
  return p;
}

template <int NB>
NT_DEVINL void gemv_epilogue(const GemvParams& p, int u, float (&d0)[NB], float (&d1)[NB], int lane) {
  // all lanes hold the full sums; lane b finishes batch row b
  if (lane >= NB) return;
  const int b = lane;
  float a0 = d0[0], a1 = d1[0];
#pragma unroll
  for (int i = 1; i < NB; ++i)
    if (b == i) a0 = d0[i], a1 = d1[i];
  const int r0 = 2 * u;
  if (p.bias) {
    a0 += p.bias[r0];
    a1 += p.bias[r0 + 1];
  }
  if (p.epi == GEMV_STORE) {
    if (p.residual) {
      a0 += p.residual[b * p.ldr + r0];
      a1 += p.residual[b * p.ldr + r0 + 1];
    }
    *reinterpret_cast<float2*>(p.out + b * p.ldo + r0) = make_float2(a0, a1);
  } else if (p.epi == GEMV_SWIGLU) {
    p.out[b * p.ldo + u] = silu(a0) * a1;
  } else {  // GEMV_QKV_ROPE
    const int head = u >> 5;  // 32 units per 64-row head
    const int i = u & 31;
    const int pos = p.kv.seq_lens[b];
    const int n_kv = p.kv.n_kv_heads;
    if (head < p.n_heads + n_kv) {
      // rows (i, i+32) of a q/k head: half-split rotation (modeling_qwen2.py:116-146)
      float s, c;
      sincosf(static_cast<float>(pos) * p.inv_freq[i], &s, &c);
      const float lo = a0 * c - a1 * s;
      const float hi = a1 * c + a0 * s;
      if (head < p.n_heads) {
        float* q = p.q_out + (static_cast<long long>(b) * p.n_heads + head) * 64;
        q[i] = lo;
        q[i + 32] = hi;
      } else if (pos < p.kv.max_ctx) {
        const int page = p.kv.page_table[b * p.kv.max_pages_per_seq + (pos >> 6)];
        __nv_bfloat16* kp = p.kv.page_ptr(p.layer, 0, page, head - p.n_heads) + (pos & 63) * 64;
        kp[i] = __float2bfloat16(lo);
        kp[i + 32] = __float2bfloat16(hi);
      }
    } else if (pos < p.kv.max_ctx) {
      const int page = p.kv.page_table[b * p.kv.max_pages_per_seq + (pos >> 6)];
      __nv_bfloat16* vp = p.kv.page_ptr(p.layer, 1, page, head - p.n_heads - n_kv) + (pos & 63) * 64;
      *reinterpret_cast<__nv_bfloat162*>(vp + 2 * i) = __floats2bfloat162_rn(a0, a1);
    }
  }
}

template <int NB>
__global__ void __launch_bounds__(kGemvThreads, 1) gemv_kernel(const GemvParams p, const GemvSmemPlan plan) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* ring = smem + plan.ring_off;
  float4* xs = reinterpret_cast<float4*>(smem + plan.x_off);
  float* red = reinterpret_cast<float*>(smem + plan.red_off);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + plan.bar_off);
  uint64_t* empty_bar = full_bar + 8;
  __shared__ float s_scale[4];
  __shared__ float s_part[kGemvConsumerWarps][4];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K;
  const int nch = K >> 3;  // 16-byte chunks per row
  const int nunits = p.rows >> 1;
  const int u_begin = static_cast<int>((static_cast<long long>(nunits) * blockIdx.x) / gridDim.x);
  const int u_end = static_cast<int>((static_cast<long long>(nunits) * (blockIdx.x + 1)) / gridDim.x);
  const int my_units = u_end - u_begin;
  const int ups = plan.units_per_stage;
  const int total_stages = (my_units + ups - 1) / ups;
  const int unit_bytes = 4 * K;
  const int NS = plan.nstages;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kGemvConsumerWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_launch_dependents();

  const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.W) + static_cast<long long>(u_begin) * unit_bytes;
  auto issue_stage = [&](int it) {
    const int s = it % NS;
    const int u0 = it * ups;
    const int n = min(ups, my_units - u0);
    const uint32_t bytes = static_cast<uint32_t>(n) * unit_bytes;
    mbar_arrive_expect_tx(&full_bar[s], bytes);
    bulk_g2s(ring + static_cast<size_t>(s) * plan.stage_bytes, wbase + static_cast<long long>(u0) * unit_bytes, bytes,
             &full_bar[s]);
  };

  // weight prefetch: independent of the previous kernel, so it goes before the dependency wait
  if (warp == kGemvConsumerWarps && lane == 0) {
    const int pre = min(NS, total_stages);
    for (int it = 0; it < pre; ++it) issue_stage(it);
  }

  pdl_wait();

  // ---- input vector(s) -> shared memory planes, optional fused RMSNorm
  // plane layout: element k = 8c + j lives in xs[(2b + j/4) * nch + c] component j%4
  if (warp < kGemvConsumerWarps) {
    float ssq[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) ssq[b] = 0.f;
    const int nvec = K >> 2;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float4* src = reinterpret_cast<const float4*>(p.x + b * p.ldx);
      for (int m = tid; m < nvec; m += kGemvConsumerWarps * 32) {
        const float4 v = src[m];
        xs[(2 * b + (m & 1)) * nch + (m >> 1)] = v;
        ssq[b] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    }
    if (p.norm_w) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float t = warp_sum(ssq[b]);
        if (lane == 0) s_part[warp][b] = t;
      }
    }
  }
  __syncthreads();
  if (p.norm_w) {
    if (tid < NB) {
      float t = 0.f;
      for (int w = 0; w < kGemvConsumerWarps; ++w) t += s_part[w][tid];
      s_scale[tid] = rsqrtf(t / static_cast<float>(K) + p.eps);
    }
    __syncthreads();
    if (warp < kGemvConsumerWarps) {
      const int nvec = K >> 2;
      const float4* nw = reinterpret_cast<const float4*>(p.norm_w);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float sc = s_scale[b];
        for (int m = tid; m < nvec; m += kGemvConsumerWarps * 32) {
          float4& v = xs[(2 * b + (m & 1)) * nch + (m >> 1)];
          const float4 g = __ldg(nw + m);
          v.x = v.x * sc * g.x, v.y = v.y * sc * g.y, v.z = v.z * sc * g.z, v.w = v.w * sc * g.w;
        }
      }
    }
    __syncthreads();
  }

  if (warp == kGemvConsumerWarps) {
    // ---- producer: refill slots as the consumers release them
    if (lane == 0) {
      for (int it = NS; it < total_stages; ++it) {
        const int s = it % NS;
        const uint32_t ph = ((it / NS) - 1) & 1;  // completion of the slot's previous use
        mbar_wait(&empty_bar[s], ph);
        issue_stage(it);
      }
    }
    return;
  }

  // ---- consumers
  for (int it = 0; it < total_stages; ++it) {
    const int s = it % NS;
    const uint32_t ph = (it / NS) & 1;
    mbar_wait(&full_bar[s], ph);
    const uint8_t* st = ring + static_cast<size_t>(s) * plan.stage_bytes;
    if (plan.wpu == 1) {
      const int ul = it * ups + warp;  // this warp's unit inside the CTA slice
      float d0[NB], d1[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) d0[b] = d1[b] = 0.f;
      const bool has = ul < my_units;
      if (has) {
        const uint4* r0 = reinterpret_cast<const uint4*>(st + static_cast<size_t>(warp) * unit_bytes);
        const uint4* r1 = r0 + nch;
        for (int c = lane; c < nch; c += 32) {
          float f0[8], f1[8];
          bf16x8_to_f32(r0[c], f0);
          bf16x8_to_f32(r1[c], f1);
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const float4 xa = xs[(2 * b) * nch + c];
            const float4 xb = xs[(2 * b + 1) * nch + c];
            d0[b] += f0[0] * xa.x + f0[1] * xa.y + f0[2] * xa.z + f0[3] * xa.w + f0[4] * xb.x + f0[5] * xb.y +
                     f0[6] * xb.z + f0[7] * xb.w;
            d1[b] += f1[0] * xa.x + f1[1] * xa.y + f1[2] * xa.z + f1[3] * xa.w + f1[4] * xb.x + f1[5] * xb.y +
                     f1[6] * xb.z + f1[7] * xb.w;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      if (has) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          d0[b] = warp_sum(d0[b]);
          d1[b] = warp_sum(d1[b]);
        }
        gemv_epilogue<NB>(p, u_begin + ul, d0, d1, lane);
      }
    } else {
      // one unit per stage, the 8 warps split K
      const int c_lo = (nch * warp) / kGemvConsumerWarps, c_hi = (nch * (warp + 1)) / kGemvConsumerWarps;
      const uint4* r0 = reinterpret_cast<const uint4*>(st);
      const uint4* r1 = r0 + nch;
      float d0[NB], d1[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) d0[b] = d1[b] = 0.f;
      for (int c = c_lo + lane; c < c_hi; c += 32) {
        float f0[8], f1[8];
        bf16x8_to_f32(r0[c], f0);
        bf16x8_to_f32(r1[c], f1);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float4 xa = xs[(2 * b) * nch + c];
          const float4 xb = xs[(2 * b + 1) * nch + c];
          d0[b] += f0[0] * xa.x + f0[1] * xa.y + f0[2] * xa.z + f0[3] * xa.w + f0[4] * xb.x + f0[5] * xb.y +
                   f0[6] * xb.z + f0[7] * xb.w;
          d1[b] += f1[0] * xa.x + f1[1] * xa.y + f1[2] * xa.z + f1[3] * xa.w + f1[4] * xb.x + f1[5] * xb.y +
                   f1[6] * xb.z + f1[7] * xb.w;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      float* rbuf = red + (it & 1) * (kGemvConsumerWarps * 2 * 4);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        d0[b] = warp_sum(d0[b]);
        d1[b] = warp_sum(d1[b]);
      }
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          rbuf[(warp * 2 + 0) * 4 + b] = d0[b];
          rbuf[(warp * 2 + 1) * 4 + b] = d1[b];
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // consumer warps only
      if (warp == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          float t0 = 0.f, t1 = 0.f;
          for (int w = 0; w < kGemvConsumerWarps; ++w) {
            t0 += rbuf[(w * 2 + 0) * 4 + b];
            t1 += rbuf[(w * 2 + 1) * 4 + b];
          }
          d0[b] = t0, d1[b] = t1;
        }
        gemv_epilogue<NB>(p, u_begin + it, d0, d1, lane);
      }
    }
  }
}

int launch_gemv(const GemvParams& p, int nb, int num_sms, cudaStream_t stream) {
  if (nb < 1 || nb > 4) return set_error(NT_ERR_INVALID, "gemv: batch %d not in 1..4", nb);
  if (p.rows & 1) return set_error(NT_ERR_INVALID, "gemv: odd row count %d", p.rows);
  if (p.K % 64) return set_error(NT_ERR_INVALID, "gemv: K=%d must be a multiple of 64", p.K);
  GemvSmemPlan plan = gemv_plan(p.K, nb);
  if (plan.total > 227 * 1024) return set_error(NT_ERR_INVALID, "gemv: K=%d needs %zu B of shared memory", p.K, plan.total);
  const int grid = min(num_sms, p.rows / 2);
  static size_t attr_bytes[5] = {0, 0, 0, 0, 0};
  void (*kern)(const GemvParams, const GemvSmemPlan) = nullptr;
  switch (nb) {
    case 1: kern = gemv_kernel<1>; break;
    case 2: kern = gemv_kernel<2>; break;
    case 3: kern = gemv_kernel<3>; break;
    default: kern = gemv_kernel<4>; break;
  }
  if (attr_bytes[nb] < plan.total) {
    NT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(plan.total)));
    attr_bytes[nb] = plan.total;
  }
  return launch_kernel(kern, dim3(grid), dim3(kGemvThreads), plan.total, stream, true, p, plan);
}

// =================================================================================== decode attention
// grid (max_splits, n_kv_heads, B), 256 threads.  One 64-token page of one KV head per CTA.
__global__ void __launch_bounds__(256) attn_decode_kernel(const AttnDecParams p) {
  __shared__ __align__(128) __nv_bfloat16 sK[64 * 64];
  __shared__ __align__(128) __nv_bfloat16 sV[64 * 64];
  __shared__ float sQ[8][64];
  __shared__ float sS[8][64];
  __shared__ float sML[8][2];
  __shared__ float sRed[4][8][64];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_last;

  pdl_launch_dependents();
  pdl_wait();

  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_ctx = min(p.kv.seq_lens[b] + 1, p.kv.max_ctx);
  const int nsplit = (n_ctx + 63) >> 6;
  if (split >= nsplit) return;
  const int n_rep = p.n_rep;
  const int page = p.kv.page_table[b * p.kv.max_pages_per_seq + split];

  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(&bar, 2 * 8192);
    bulk_g2s(sK, p.kv.page_ptr(p.layer, 0, page, kvh), 8192, &bar);
    bulk_g2s(sV, p.kv.page_ptr(p.layer, 1, page, kvh), 8192, &bar);
  }
  for (int i = tid; i < n_rep * 64; i += 256)
    sQ[i >> 6][i & 63] = p.q[(static_cast<long long>(b) * p.n_heads + kvh * n_rep + (i >> 6)) * 64 + (i & 63)];
  __syncthreads();
  mbar_wait(&bar, 0);

  // ---- scores: thread = (token, quarter of the head dim)
  {
    const int tok = tid >> 2, part = tid & 3;
    const uint4* kr = reinterpret_cast<const uint4*>(sK + tok * 64 + part * 16);
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
    const bool valid = (split * 64 + tok) < n_ctx;
    for (int h = 0; h < n_rep; ++h) {
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) d += kf[j] * sQ[h][part * 16 + j];
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      if (part == 0) sS[h][tok] = valid ? d * p.scale_log2 : -INFINITY;
    }
  }
  __syncthreads();
  // ---- per-head softmax partials (fp32, base-2 exponent with log2e folded into the scale)
  if (warp < n_rep) {
    const float s0 = sS[warp][lane], s1 = sS[warp][lane + 32];
    const float m = warp_max(fmaxf(s0, s1));  // position 0 of split 0 is always valid -> finite
    const float p0 = exp2f(s0 - m), p1 = exp2f(s1 - m);
    const float l = warp_sum(p0 + p1);
    sS[warp][lane] = p0;
    sS[warp][lane + 32] = p1;
    if (lane == 0) sML[warp][0] = m, sML[warp][1] = l;
  }
  __syncthreads();
  // ---- P.V : thread = (dim, token group of 16)
  {
    const int d = tid & 63, g = tid >> 6;
    float acc[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) acc[h] = 0.f;
    for (int t = g * 16; t < g * 16 + 16; ++t) {
      const float v = __bfloat162float(sV[t * 64 + d]);
#pragma unroll
      for (int h = 0; h < 8; ++h)
        if (h < n_rep) acc[h] += sS[h][t] * v;
    }
#pragma unroll
    for (int h = 0; h < 8; ++h)
      if (h < n_rep) sRed[g][h][d] = acc[h];
  }
  __syncthreads();
  for (int i = tid; i < n_rep * 64; i += 256) {
    const int h = i >> 6, d = i & 63;
    const float o = sRed[0][h][d] + sRed[1][h][d] + sRed[2][h][d] + sRed[3][h][d];
    const long long hh = static_cast<long long>(b) * p.n_heads + kvh * n_rep + h;
    p.part_o[(hh * p.max_splits + split) * 64 + d] = o;
    if (d == 0) {
      p.part_ml[(hh * p.max_splits + split) * 2 + 0] = sML[h][0];
      p.part_ml[(hh * p.max_splits + split) * 2 + 1] = sML[h][1];
    }
  }
  // ---- last CTA of this (sequence, kv head) merges the splits in split order (deterministic)
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int old = atomicAdd(&p.counters[b * p.kv.n_kv_heads + kvh], 1);
    s_last = (old == nsplit - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = tid; i < n_rep * 64; i += 256) {
    const int h = i >> 6, d = i & 63;
    const long long hh = static_cast<long long>(b) * p.n_heads + kvh * n_rep + h;
    float M = -INFINITY;
    for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldcg(&p.part_ml[(hh * p.max_splits + s) * 2]));
    float L = 0.f, O = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float w = exp2f(__ldcg(&p.part_ml[(hh * p.max_splits + s) * 2]) - M);
      L += w * __ldcg(&p.part_ml[(hh * p.max_splits + s) * 2 + 1]);
      O += w * __ldcg(&p.part_o[(hh * p.max_splits + s) * 64 + d]);
    }
    p.out[hh * 64 + d] = O / L;
    if (p.out_bf16) p.out_bf16[hh * 64 + d] = __float2bfloat16(O / L);
  }
  if (tid == 0) p.counters[b * p.kv.n_kv_heads + kvh] = 0;
}

int launch_attn_decode(const AttnDecParams& p, int B, cudaStream_t stream) {
  if (p.n_rep < 1 || p.n_rep > 8) return set_error(NT_ERR_INVALID, "attention: %d query heads per KV head unsupported (1..8)", p.n_rep);
  return launch_kernel(attn_decode_kernel, dim3(p.max_splits, p.kv.n_kv_heads, B), dim3(256), 0, stream, true, p);
}

// =================================================================================== sampler
struct Cand {
  float v;
  int i;
};
NT_DEVINL bool cand_before(const Cand& a, const Cand& b) { return a.v > b.v || (a.v == b.v && a.i < b.i); }

// full descending bitonic sort of n (power of two) candidates in shared memory
NT_DEVINL void bitonic_sort_desc(Cand* a, int n, int tid, int nthreads) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n >> 1); t += nthreads) {
        const int i = ((t / j) * 2 * j) + (t % j);
        const int l = i + j;
        const bool desc = ((i & k) == 0);
        const Cand x = a[i], y = a[l];
        const bool swap = desc ? cand_before(y, x) : cand_before(x, y);
        if (swap) a[i] = y, a[l] = x;
      }
      __syncthreads();
    }
  }
}

NT_DEVINL void philox4x32_10(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
    const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ k1, n3 = lo0;
    ctr[0] = n0, ctr[1] = n1, ctr[2] = n2, ctr[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

constexpr int kTopChunk = 2048;
constexpr int kTopKeep = 64;
constexpr int kSelThreads = 256;

int sampler_nchunks(int V) { return (V + kTopChunk - 1) / kTopChunk; }
size_t sampler_scratch_floats(int B, int V) { return size_t(B) * sampler_nchunks(V) * kTopKeep; }

// order-preserving float -> uint key (larger float <=> larger key; -inf is the smallest finite key)
NT_DEVINL uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Block-wide radix select (4 passes of 8 bits, MSB first) over n keys in shared memory:
// finds the key of the k-th largest element and how many elements equal to it belong to the
// top-k.  All kSelThreads threads of the block must call it.  scratch: >= 258 uint32.
NT_DEVINL void radix_select_kth(const uint32_t* keys, int n, int k, uint32_t* scratch, uint32_t& thr, int& take_eq) {
  uint32_t* hist = scratch;           // [256]
  uint32_t* sel = scratch + 256;      // [2]: bin, remaining
  const int tid = threadIdx.x, lane = tid & 31;
  uint32_t prefix = 0, mask = 0;
  int remaining = k;
  const int n_pad = (n + 31) & ~31;
#pragma unroll 1
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n_pad; i += kSelThreads) {
      uint32_t bin = 0xffffffffu;
      if (i < n) {
        const uint32_t key = keys[i];
        if ((key & mask) == prefix) bin = (key >> shift) & 255u;
      }
      // one shared-memory atomic per distinct bin per warp (logits crowd into few top-byte bins)
      const uint32_t peers = __match_any_sync(0xffffffffu, bin);
      if (bin != 0xffffffffu && lane == (__ffs(peers) - 1)) atomicAdd(&hist[bin], __popc(peers));
    }
    __syncthreads();
    if (tid < 32) {
      uint32_t c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c[j] = hist[8 * lane + j];
        sum += c[j];
      }
      uint32_t suf = sum;  // elements in bins >= 8*lane
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_down_sync(0xffffffffu, suf, off);
        if (lane + off < 32) suf += t;
      }
      const uint32_t above = suf - sum;
      if (above < static_cast<uint32_t>(remaining) && static_cast<uint32_t>(remaining) <= suf) {
        uint32_t acc = above;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
          if (acc + c[j] >= static_cast<uint32_t>(remaining)) {
            sel[0] = 8 * lane + j;
            sel[1] = remaining - acc;
            break;
          }
          acc += c[j];
        }
      }
    }
    __syncthreads();
    prefix |= sel[0] << shift;
    mask |= 0xffu << shift;
    remaining = static_cast<int>(sel[1]);
    __syncthreads();
  }
  thr = prefix;
  take_eq = remaining;
}

// Deterministic compaction of the top-k winners (keys > thr, plus the first take_eq keys == thr
// in index order) into out slots [0, k).  Elements are owned in contiguous runs per thread so a
// block scan preserves index order.  scratch: >= 2*8+2 uint32.
template <typename Emit>
NT_DEVINL void compact_topk(const uint32_t* keys, int n, uint32_t thr, int take_eq, uint32_t* scratch, Emit emit) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + kSelThreads - 1) / kSelThreads;
  const int lo = min(n, tid * per), hi = min(n, lo + per);
  int ngt = 0, neq = 0;
  for (int i = lo; i < hi; ++i) {
    const uint32_t key = keys[i];
    ngt += key > thr;
    neq += key == thr;
  }
  // exclusive scans across the block (warp shuffles + one smem hop)
  int sgt = ngt, seq = neq;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, sgt, off), b = __shfl_up_sync(0xffffffffu, seq, off);
    if (lane >= off) sgt += a, seq += b;
  }
  uint32_t* wg = scratch;       // [8] per-warp totals (gt)
  uint32_t* we = scratch + 8;   // [8] per-warp totals (eq)
  __syncthreads();
  if (lane == 31) wg[warp] = sgt, we[warp] = seq;
  __syncthreads();
  int bg = 0, be = 0, total_gt = 0;
  for (int w = 0; w < kSelThreads / 32; ++w) {
    if (w < warp) bg += wg[w], be += we[w];
    total_gt += wg[w];
  }
  int pos_gt = bg + sgt - ngt;          // exclusive prefix of "greater" elements
  int idx_eq = be + seq - neq;          // exclusive prefix of "equal" elements
  for (int i = lo; i < hi; ++i) {
    const uint32_t key = keys[i];
    if (key > thr) {
      emit(pos_gt++, i);
    } else if (key == thr) {
      if (idx_eq < take_eq) emit(total_gt + idx_eq, i);
      ++idx_eq;
    }
  }
}

// stage 1: grid (nchunks, B), 256 threads: logits processors + exact top-64 of a 2048-logit chunk
__global__ void __launch_bounds__(kSelThreads) topk_stage1_kernel(const SamplerParams p) {
  __shared__ uint32_t keys[kTopChunk];
  __shared__ uint32_t scratch[260];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int ngen = p.n_generated_override ? p.n_generated_override[b] : p.n_generated[b];
  const bool mask_eos = ngen < p.sp.min_new_tokens;
  const float inv_t = 1.0f / p.sp.temperature;
  const float* lg = p.logits + static_cast<long long>(b) * p.V;
  const int base = chunk * kTopChunk;
  const int n = min(kTopChunk, p.V - base);
  for (int e = tid; e < n; e += kSelThreads) {
    float v = lg[base + e];
    if (mask_eos && base + e == p.sp.eos_id) v = -INFINITY;  // MinNewTokensLength
    keys[e] = f2key(v * inv_t);                              // Temperature
  }
  __syncthreads();
  const long long o = (static_cast<long long>(b) * p.nchunks + chunk) * kTopKeep;
  const int k = min(kTopKeep, n);
  uint32_t thr;
  int take_eq;
  radix_select_kth(keys, n, k, scratch, thr, take_eq);
  compact_topk(keys, n, thr, take_eq, scratch, [&](int slot, int i) {
    p.cand_val[o + slot] = lg[base + i];   // raw logit; stage 2 re-applies the processors
    p.cand_idx[o + slot] = base + i;
  });
  for (int s = k + tid; s < kTopKeep; s += kSelThreads) {
    p.cand_val[o + s] = -INFINITY;
    p.cand_idx[o + s] = 0x7fffffff;
  }
}

// stage 2: grid (B), 256 threads: top-k of the candidates, softmax, draw, state update, next embedding
__global__ void __launch_bounds__(kSelThreads) topk_stage2_kernel(const SamplerParams p, const int ncand) {
  extern __shared__ uint8_t smem_raw[];
  uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw);    // [ncand]
  __shared__ uint32_t scratch[260];
  __shared__ Cand win[kTopKeep];
  __shared__ int s_tok;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  const bool stateless = p.n_generated_override != nullptr;
  const int ngen = stateless ? p.n_generated_override[b] : p.n_generated[b];
  const bool is_done = stateless ? false : (p.done[b] != 0);
  const bool mask_eos = ngen < p.sp.min_new_tokens;
  const float inv_t = 1.0f / p.sp.temperature;
  const float* cv = p.cand_val + static_cast<long long>(b) * ncand;
  const int32_t* ci = p.cand_idx + static_cast<long long>(b) * ncand;
  for (int e = tid; e < ncand; e += kSelThreads) {
    float v = cv[e];
    if (mask_eos && ci[e] == p.sp.eos_id) v = -INFINITY;
    keys[e] = f2key(v * inv_t);
  }
  if (tid < kTopKeep) win[tid].v = -INFINITY, win[tid].i = 0x7fffffff;
  __syncthreads();
  const int k = min(min(p.sp.top_k, kTopKeep), ncand);
  uint32_t thr;
  int take_eq;
  radix_select_kth(keys, ncand, k, scratch, thr, take_eq);
  compact_topk(keys, ncand, thr, take_eq, scratch, [&](int slot, int i) {
    float v = cv[i];
    if (mask_eos && ci[i] == p.sp.eos_id) v = -INFINITY;
    win[slot].v = v * inv_t;
    win[slot].i = ci[i];
  });
  __syncthreads();
  bitonic_sort_desc(win, kTopKeep, tid, kSelThreads);   // 64 winners: (score desc, index asc)

  if (tid < 32) {
    // softmax over the k kept scores (TopK processor + softmax, utils.py:2789)
    const float m = win[0].v;
    const float e0 = (tid < k) ? __expf(win[tid].v - m) : 0.f;
    const float e1 = (tid + 32 < k) ? __expf(win[tid + 32].v - m) : 0.f;
    const float sum = warp_sum(e0 + e1);
    if (p.dbg_topk_val) {
      p.dbg_topk_val[b * kTopKeep + tid] = (tid < k) ? e0 / sum : 0.f;
      p.dbg_topk_val[b * kTopKeep + tid + 32] = (tid + 32 < k) ? e1 / sum : 0.f;
      p.dbg_topk_idx[b * kTopKeep + tid] = (tid < k) ? win[tid].i : -1;
      p.dbg_topk_idx[b * kTopKeep + tid + 32] = (tid + 32 < k) ? win[tid + 32].i : -1;
    }
    if (tid == 0) {
      int tok;
      if (p.sp.forced && !stateless) {
        tok = p.sp.forced[static_cast<long long>(b) * p.max_new + ngen];
      } else if (p.sp.greedy) {
        tok = win[0].i;
      } else {
        uint32_t ctr[4] = {static_cast<uint32_t>(stateless ? p.step_override : ngen), static_cast<uint32_t>(b), 0u, 0u};
        philox4x32_10(ctr, static_cast<uint32_t>(p.sp.seed), static_cast<uint32_t>(p.sp.seed >> 32));
        const float u = (ctr[0] >> 8) * (1.0f / 16777216.0f);  // [0,1)
        const float target = u * sum;
        float cum = 0.f;
        tok = win[k - 1].i;
        for (int j = 0; j < k; ++j) {
          cum += __expf(win[j].v - m);
          if (cum > target) {
            tok = win[j].i;
            break;
          }
        }
      }
      s_tok = tok;
      if (p.dbg_token) p.dbg_token[b] = tok;
      if (!stateless && !is_done) {
        p.out_tokens[static_cast<long long>(b) * p.max_new + ngen] = tok;
        p.n_generated[b] = ngen + 1;
        p.cur_token[b] = tok;
        const int cached = p.seq_lens[b] + p.advance;  // decode: this step's input token is now in the KV cache
        if (p.advance) p.seq_lens[b] = cached;
        const int total = cached + 1;  // tokens in context once `tok` is appended
        if (tok == p.sp.eos_id || ngen + 1 >= p.sp.max_new_tokens || ngen + 1 >= p.max_new || total >= p.max_ctx)
          p.done[b] = 1;
      }
    }
  }
  __syncthreads();
  if (!stateless && !is_done && p.h) {
    const int tok = s_tok;
    const __nv_bfloat16* e = p.embed + static_cast<long long>(tok) * p.hidden;
    for (int i = tid; i < p.hidden; i += kSelThreads) p.h[static_cast<long long>(b) * p.hidden + i] = __bfloat162float(e[i]);
  }
}

int launch_sampler(const SamplerParams& p, int B, cudaStream_t stream) {
  if (p.sp.top_k < 1 || p.sp.top_k > kTopKeep) return set_error(NT_ERR_INVALID, "sampler: top_k=%d not in 1..64", p.sp.top_k);
  if (!(p.sp.temperature > 0.f)) return set_error(NT_ERR_INVALID, "sampler: temperature must be > 0");
  const int ncand = p.nchunks * kTopKeep;
  const size_t smem = size_t(ncand) * sizeof(uint32_t);
  if (smem > 200 * 1024) return set_error(NT_ERR_INVALID, "sampler: vocabulary too large (%d)", p.V);
  int rc = launch_kernel(topk_stage1_kernel, dim3(p.nchunks, B), dim3(kSelThreads), 0, stream, true, p);
  if (rc) return rc;
  static size_t attr = 0;
  if (smem > 40 * 1024 && attr < smem) {
    NT_CUDA_CHECK(cudaFuncSetAttribute(topk_stage2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    attr = smem;
  }
  return launch_kernel(topk_stage2_kernel, dim3(B), dim3(kSelThreads), smem, stream, true, p, ncand);
}

// =================================================================================== prefill helpers
__global__ void embed_rows_kernel(const __nv_bfloat16* embed, const int32_t* ids, int hidden, float* h) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x;
  const __nv_bfloat16* e = embed + static_cast<long long>(ids[t]) * hidden;
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) h[static_cast<long long>(t) * hidden + i] = __bfloat162float(e[i]);
}
int launch_embed_rows(const __nv_bfloat16* embed, const int32_t* ids, int T, int hidden, float* h, cudaStream_t s) {
  return launch_kernel(embed_rows_kernel, dim3(T), dim3(256), 0, s, true, embed, ids, hidden, h);
}

// one warp per row; fp32 statistics; out = w * (x * rsqrt(mean(x^2)+eps))  (modeling_qwen2.py:258-263)
__global__ void __launch_bounds__(256) rmsnorm_rows_kernel(const float* x, const float* w, float eps, int rows, int cols,
                                                           float* out_f32, __nv_bfloat16* out_bf16) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * cols;
  float ss = 0.f;
  for (int i = lane; i < cols; i += 32) ss += xr[i] * xr[i];
  ss = warp_sum(ss);
  const float sc = rsqrtf(ss / static_cast<float>(cols) + eps);
  for (int i = lane; i < cols; i += 32) {
    const float v = w[i] * (xr[i] * sc);
    if (out_f32) out_f32[static_cast<long long>(row) * cols + i] = v;
    if (out_bf16) out_bf16[static_cast<long long>(row) * cols + i] = __float2bfloat16(v);
  }
}
int launch_rmsnorm_rows(const float* x, const float* w, float eps, int rows, int cols, float* out_f32,
                        __nv_bfloat16* out_bf16, cudaStream_t s) {
  return launch_kernel(rmsnorm_rows_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, true, x, w, eps, rows, cols, out_f32,
                       out_bf16);
}

// thread = one unit (pair of packed columns) of one token
__global__ void __launch_bounds__(256) rope_append_kernel(const float* qkv, int qkv_n, const int32_t* tok_seq,
                                                          const int32_t* tok_pos, int n_heads, const float* inv_freq,
                                                          float* q_out, const KVLayout kv, int layer) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x;
  const int u = blockIdx.y * 256 + threadIdx.x;
  if (u >= (qkv_n >> 1)) return;
  const float2 v = *reinterpret_cast<const float2*>(qkv + static_cast<long long>(t) * qkv_n + 2 * u);
  const int head = u >> 5, i = u & 31;
  const int b = tok_seq[t], pos = tok_pos[t];
  const int n_kv = kv.n_kv_heads;
  if (head < n_heads + n_kv) {
    float s, c;
    sincosf(static_cast<float>(pos) * inv_freq[i], &s, &c);
    const float lo = v.x * c - v.y * s, hi = v.y * c + v.x * s;
    if (head < n_heads) {
      float* q = q_out + (static_cast<long long>(t) * n_heads + head) * 64;
      q[i] = lo, q[i + 32] = hi;
    } else {
      const int page = kv.page_table[b * kv.max_pages_per_seq + (pos >> 6)];
      __nv_bfloat16* kp = kv.page_ptr(layer, 0, page, head - n_heads) + (pos & 63) * 64;
      kp[i] = __float2bfloat16(lo), kp[i + 32] = __float2bfloat16(hi);
    }
  } else {
    const int page = kv.page_table[b * kv.max_pages_per_seq + (pos >> 6)];
    __nv_bfloat16* vp = kv.page_ptr(layer, 1, page, head - n_heads - n_kv) + (pos & 63) * 64;
    *reinterpret_cast<__nv_bfloat162*>(vp + 2 * i) = __floats2bfloat162_rn(v.x, v.y);
  }
}
int launch_rope_append(const float* qkv, int T, int qkv_n, const int32_t* tok_seq, const int32_t* tok_pos, int n_heads,
                       const float* inv_freq, float* q_out, const KVLayout& kv, int layer, cudaStream_t s) {
  return launch_kernel(rope_append_kernel, dim3(T, ((qkv_n >> 1) + 255) / 256), dim3(256), 0, s, true, qkv, qkv_n, tok_seq,
                       tok_pos, n_heads, inv_freq, q_out, kv, layer);
}

// Causal GQA attention for prefill, fp32 math on CUDA cores (first version: correctness and a
// sane baseline; the tensor-core flash kernel replaces it for large batches).
// grid (ceil(max_len/16), n_kv_heads, B); thread = one (query, head-in-group) row.
constexpr int kPfQ = 16;
__global__ void __launch_bounds__(128) attn_prefill_kernel(const AttnPrefillParams p) {
  __shared__ __align__(16) __nv_bfloat16 sK[32 * 64];
  __shared__ __align__(16) __nv_bfloat16 sV[32 * 64];
  pdl_launch_dependents();
  pdl_wait();
  const int qb = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int t0 = p.cu_seqlens[b], len = p.cu_seqlens[b + 1] - t0;
  if (qb * kPfQ >= len) return;
  const int n_rep = p.n_rep;
  const int qi = tid / n_rep, h = tid - qi * n_rep;
  const int tq = qb * kPfQ + qi;
  const bool active = (qi < kPfQ) && (tq < len);
  float q[64], o[64];
  float m = -INFINITY, l = 0.f;
  if (active) {
    const float4* qp = reinterpret_cast<const float4*>(p.q + (static_cast<long long>(t0 + tq) * p.n_heads + kvh * n_rep + h) * 64);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 v = qp[j];
      q[4 * j] = v.x * p.scale_log2, q[4 * j + 1] = v.y * p.scale_log2, q[4 * j + 2] = v.z * p.scale_log2,
            q[4 * j + 3] = v.w * p.scale_log2;
    }
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) o[j] = 0.f;
  const int kend = min(len, (qb + 1) * kPfQ);
  const int ntiles = (kend + 31) >> 5;
  for (int tile = 0; tile < ntiles; ++tile) {
    const int k0 = tile * 32;
    const int page = p.kv.page_table[b * p.kv.max_pages_per_seq + (k0 >> 6)];
    const uint4* gk = reinterpret_cast<const uint4*>(p.kv.page_ptr(p.layer, 0, page, kvh) + (k0 & 63) * 64);
    const uint4* gv = reinterpret_cast<const uint4*>(p.kv.page_ptr(p.layer, 1, page, kvh) + (k0 & 63) * 64);
    __syncthreads();
    for (int i = tid; i < 256; i += 128) {
      reinterpret_cast<uint4*>(sK)[i] = gk[i];
      reinterpret_cast<uint4*>(sV)[i] = gv[i];
    }
    __syncthreads();
    if (!active) continue;
    float s[32];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float d = 0.f;
      const uint4* kr = reinterpret_cast<const uint4*>(sK + j * 64);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float f[8];
        bf16x8_to_f32(kr[c], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) d += q[8 * c + e] * f[e];
      }
      s[j] = (k0 + j <= tq) ? d : -INFINITY;  // causal mask (key position <= query position)
      tmax = fmaxf(tmax, s[j]);
    }
    const float mn = fmaxf(m, tmax);  // key 0 is always visible -> finite from the first tile on
    const float corr = exp2f(m - mn);
    l *= corr;
#pragma unroll
    for (int j = 0; j < 64; ++j) o[j] *= corr;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float pj = exp2f(s[j] - mn);
      l += pj;
      const uint4* vr = reinterpret_cast<const uint4*>(sV + j * 64);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float f[8];
        bf16x8_to_f32(vr[c], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[8 * c + e] += pj * f[e];
      }
    }
    m = mn;
  }
  if (active) {
    const float inv = 1.0f / l;
    __nv_bfloat16* op = p.out + (static_cast<long long>(t0 + tq) * p.n_heads + kvh * n_rep + h) * 64;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      reinterpret_cast<uint4*>(op)[j] =
          make_uint4(pack_bf16x2(o[8 * j] * inv, o[8 * j + 1] * inv), pack_bf16x2(o[8 * j + 2] * inv, o[8 * j + 3] * inv),
                     pack_bf16x2(o[8 * j + 4] * inv, o[8 * j + 5] * inv), pack_bf16x2(o[8 * j + 6] * inv, o[8 * j + 7] * inv));
  }
}
int launch_attn_prefill(const AttnPrefillParams& p, int B, cudaStream_t s) {
  if (p.n_rep * kPfQ > 128) return set_error(NT_ERR_INVALID, "prefill attention: %d query heads per KV head unsupported", p.n_rep);
  return launch_kernel(attn_prefill_kernel, dim3((p.max_len + kPfQ - 1) / kPfQ, p.kv.n_kv_heads, B), dim3(128), 0, s, true, p);
}

__global__ void gather_rows_kernel(const float* src, const int32_t* rows, int cols, float* dst) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  const float* s = src + static_cast<long long>(rows[r]) * cols;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) dst[static_cast<long long>(r) * cols + i] = s[i];
}
int launch_gather_rows(const float* src, const int32_t* rows, int n, int cols, float* dst, cudaStream_t s) {
  return launch_kernel(gather_rows_kernel, dim3(n), dim3(256), 0, s, true, src, rows, cols, dst);
}

}  // namespace nt

"""Speech-LM engine on the C-ABI: weight packing, paged KV pool, prefill + device-side decode loop.

Host-side mirror of inner seam 1 of the reference (``neutts/neutts.py:334-352``): an object with
``.device`` and ``.generate(LongTensor[1,P], max_length=, eos_token_id=, do_sample=, temperature=,
top_k=, use_cache=, min_new_tokens=) -> LongTensor[1,P+N]``, plus a batched ``generate_batch``.
All arithmetic happens in ``libneutts_b200.so``; torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


@dataclass
class LMShape:
    """Decoder shape; read from the checkpoint's config.json at load time (never hard-coded:
    SURVEY.md §8).  Defaults = NeuTTS-Air as inferred from TRAINING.md:33 / README.md:44-45."""

    vocab_size: int = 217472
    hidden_size: int = 896
    intermediate_size: int = 4864
    num_layers: int = 24
    num_heads: int = 14
    num_kv_heads: int = 2
    head_dim: int = 64
    rms_eps: float = 1e-6
    rope_theta: float = 1e6
    tie_embeddings: bool = True

    @staticmethod
    def from_hf_config(cfg: dict) -> "LMShape":
        rope = cfg.get("rope_theta")
        if rope is None and isinstance(cfg.get("rope_parameters"), dict):
            rope = cfg["rope_parameters"].get("rope_theta")
        heads = cfg["num_attention_heads"]
        return LMShape(
            vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"], intermediate_size=cfg["intermediate_size"],
            num_layers=cfg["num_hidden_layers"], num_heads=heads,
            num_kv_heads=cfg.get("num_key_value_heads", heads),
            head_dim=cfg.get("head_dim") or cfg["hidden_size"] // heads,
            rms_eps=cfg.get("rms_norm_eps", 1e-6), rope_theta=float(rope if rope is not None else 1e6),
            tie_embeddings=bool(cfg.get("tie_word_embeddings", True)))


def _rope_pair_perm(n_heads: int) -> torch.Tensor:
    """Row order that puts RoPE partners (i, i+32) of every 64-row head next to each other."""
    one = torch.stack((torch.arange(32), torch.arange(32) + 32), dim=1).reshape(-1)
    return (torch.arange(n_heads)[:, None] * 64 + one[None, :]).reshape(-1)


def pack_weights(shape: LMShape, sd: dict, device) -> dict:
    """HF Qwen2 state_dict (``model.layers.N.self_attn.q_proj.weight`` ...) -> the packed bf16/fp32
    device tensors the kernels stream (layouts documented in include/neutts_b200.h)."""
    dev = torch.device(device)
    bf = lambda t: t.to(device=dev, dtype=torch.bfloat16).contiguous()
    f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
    pq = _rope_pair_perm(shape.num_heads)
    pk = _rope_pair_perm(shape.num_kv_heads)
    g = lambda k: sd[k]
    out = dict(ln1=[], wqkv=[], bqkv=[], wo=[], ln2=[], wgu=[], wd=[])
    for i in range(shape.num_layers):
        p = f"model.layers.{i}."
        wq, wk, wv = g(p + "self_attn.q_proj.weight"), g(p + "self_attn.k_proj.weight"), g(p + "self_attn.v_proj.weight")
        zb = lambda w_: torch.zeros(w_.shape[0], dtype=torch.float32)     # Llama-style checkpoints carry no q/k/v bias
        bq = sd.get(p + "self_attn.q_proj.bias", None)
        bk = sd.get(p + "self_attn.k_proj.bias", None)
        bv = sd.get(p + "self_attn.v_proj.bias", None)
        bq, bk, bv = (bq if bq is not None else zb(wq)), (bk if bk is not None else zb(wk)), (bv if bv is not None else zb(wv))
        out["wqkv"].append(bf(torch.cat((wq[pq], wk[pk], wv), dim=0)))
        out["bqkv"].append(f32(torch.cat((bq.float()[pq], bk.float()[pk], bv.float()), dim=0)))
        out["wo"].append(bf(g(p + "self_attn.o_proj.weight")))
        wg, wu = g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight")
        out["wgu"].append(bf(torch.stack((wg, wu), dim=1).reshape(2 * wg.shape[0], wg.shape[1])))
        out["wd"].append(bf(g(p + "mlp.down_proj.weight")))
        out["ln1"].append(f32(g(p + "input_layernorm.weight")))
        out["ln2"].append(f32(g(p + "post_attention_layernorm.weight")))
    out["embed"] = bf(g("model.embed_tokens.weight"))
    if shape.tie_embeddings or "lm_head.weight" not in sd:
        out["lm_head"] = out["embed"]
    else:
        out["lm_head"] = bf(g("lm_head.weight"))
    out["final_norm"] = f32(g("model.norm.weight"))
    return out


class PagePool:
    """Free-list allocator over the KV page ids (host side; one id addresses every layer)."""

    def __init__(self, num_pages: int, shuffle_seed: int | None = None):
        order = list(range(num_pages))
        if shuffle_seed is not None:
            rng = np.random.default_rng(shuffle_seed)
            rng.shuffle(order)
        self.free = order[::-1]

    def alloc(self, n: int) -> list:
        if n > len(self.free):
            raise RuntimeError(f"KV page pool exhausted: need {n}, have {len(self.free)}")
        return [self.free.pop() for _ in range(n)]

    def release(self, pages) -> None:
        self.free.extend(reversed(list(pages)))


class SpeechLM:
    PAGE = 64

    def __init__(self, shape: LMShape, state_dict: dict, device="cuda", max_batch: int = 1, max_ctx: int = 2048,
                 max_new: int | None = None, max_prefill_tokens: int | None = None, page_shuffle_seed: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("neutts_air_b200.SpeechLM needs a CUDA device (sm_100a); there is no CPU fallback")
        self.L = _lib.lib()
        self.shape = shape
        self.device = torch.device(device)
        self.max_batch, self.max_ctx = max_batch, max_ctx
        self.max_new = max_new or max_ctx
        self.max_pages = max_ctx // self.PAGE
        self.num_pages = max_batch * self.max_pages
        self.max_prefill_tokens = max_prefill_tokens or max_batch * max_ctx
        with torch.cuda.device(self.device):
            self.w = pack_weights(shape, state_dict, self.device)
            cfg = _lib.LMConfig(shape.vocab_size, shape.hidden_size, shape.intermediate_size, shape.num_layers,
                                shape.num_heads, shape.num_kv_heads, shape.head_dim, shape.rms_eps, shape.rope_theta,
                                max_batch, max_ctx, self.PAGE, self.num_pages, self.max_prefill_tokens)
            self.cfg = cfg
            ws_bytes = self.L.nt_lm_workspace_bytes(C.byref(cfg))
            if ws_bytes == 0:
                _lib.check(-1)
            self.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            self._ptrs = {k: _lib.ptr_array(self.w[k]) for k in ("ln1", "wqkv", "bqkv", "wo", "ln2", "wgu", "wd")}
            wts = _lib.LMWeights(self.w["embed"].data_ptr(), self.w["lm_head"].data_ptr(), self.w["final_norm"].data_ptr(),
                                 self._ptrs["ln1"], self._ptrs["wqkv"], self._ptrs["bqkv"], self._ptrs["wo"],
                                 self._ptrs["ln2"], self._ptrs["wgu"], self._ptrs["wd"])
            self.handle = C.c_void_p()
            _lib.check(self.L.nt_lm_create(C.byref(cfg), C.byref(wts), self.workspace.data_ptr(), ws_bytes,
                                           C.byref(self.handle)))
            # caller-owned state (zeroed KV pool: masked keys must hold finite values)
            i32 = dict(dtype=torch.int32, device=self.device)
            self.kv = torch.zeros(shape.num_layers, 2, self.num_pages, shape.num_kv_heads, self.PAGE, 64,
                                  dtype=torch.bfloat16, device=self.device)
            self.page_table = torch.zeros(max_batch, self.max_pages, **i32)
            self.seq_lens = torch.zeros(max_batch, **i32)
            self.cur_token = torch.zeros(max_batch, **i32)
            self.out_tokens = torch.zeros(max_batch, self.max_new, **i32)
            self.n_generated = torch.zeros(max_batch, **i32)
            self.done = torch.zeros(max_batch, **i32)
            self.forced = None
        self.state = _lib.LMState(self.kv.data_ptr(), self.page_table.data_ptr(), self.seq_lens.data_ptr(),
                                  self.cur_token.data_ptr(), self.out_tokens.data_ptr(), self.n_generated.data_ptr(),
                                  self.done.data_ptr(), self.max_new)
        self.pool = PagePool(self.num_pages, page_shuffle_seed)
        self._table_host = None
        self._slot_pages = [[] for _ in range(max_batch)]

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.L.nt_lm_destroy(self.handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ low-level steps
    def sampling(self, eos_id: int, min_new_tokens: int = 50, max_new_tokens: int | None = None, top_k: int = 50,
                 temperature: float = 1.0, seed: int = 0, greedy: bool = False, forced: torch.Tensor | None = None,
                 limits=None, slot_base: int = 0):
        """``limits``: optional per-sequence caps on generated tokens (list / tensor, one per slot);
        ``slot_base``: global index of slot 0, so that chunks of a larger batch and ranks of a distributed job
        draw from independent Philox streams under one seed."""
        mnt = min(max_new_tokens or self.max_new, self.max_new)
        lptr = None
        if limits is not None:
            lt = torch.full((self.max_batch,), mnt, dtype=torch.int32)
            lt[: len(limits)] = torch.as_tensor(list(limits), dtype=torch.int32)
            self.limits = lt.to(self.device)
            lptr = self.limits.data_ptr()
        fptr = None
        if forced is not None:
            f = torch.zeros(self.max_batch, self.max_new, dtype=torch.int32, device=self.device)
            f[: forced.shape[0], : forced.shape[1]] = forced.to(self.device, torch.int32)
            self.forced = f
            fptr = f.data_ptr()
        return _lib.Sampling(int(eos_id), int(min_new_tokens), int(mnt), int(top_k), float(temperature), int(seed),
                             int(bool(greedy)), fptr, lptr, int(slot_base))

    def prefill(self, prompts, sp, return_logits: bool = False):
        """prompts: list of 1-D int sequences.  Fills the KV cache and samples the first token."""
        lens = [len(p) for p in prompts]
        if not lens or min(lens) < 1:
            raise ValueError("empty prompt")
        flat = np.concatenate([np.asarray(p, dtype=np.int64) for p in prompts])
        if flat.min() < 0 or flat.max() >= self.shape.vocab_size:
            raise ValueError("token id out of range")
        # stage through a persistent pinned buffer: the H2D copy of the prompt ids is asynchronous and DMA-able
        if getattr(self, "_ids_pinned", None) is None or self._ids_pinned.numel() < flat.size:
            self._ids_pinned = torch.empty(max(flat.size, self.max_prefill_tokens), dtype=torch.int32).pin_memory()
        staged = self._ids_pinned[: flat.size]
        staged.copy_(torch.from_numpy(flat.astype(np.int32)))
        return self.prefill_packed(staged, lens, sp, return_logits)

    def prefill_packed(self, ids_host: torch.Tensor, lens, sp, return_logits: bool = False):
        """ids_host: int32 host tensor (pinned memory makes the H2D copy asynchronous) holding the
        prompts back to back; lens: their lengths."""
        B = len(lens)
        if not 1 <= B <= self.max_batch:
            raise ValueError(f"batch {B} not in 1..{self.max_batch}")
        if ids_host.dtype != torch.int32 or ids_host.numel() != sum(lens):
            raise ValueError("ids_host must be int32 with sum(lens) elements")
        for b in range(self.max_batch):
            if self._slot_pages[b]:
                self.pool.release(self._slot_pages[b])
                self._slot_pages[b] = []
        table = np.zeros((self.max_batch, self.max_pages), dtype=np.int32)
        for b, n in enumerate(lens):
            if not 1 <= n < self.max_ctx:
                raise ValueError(f"prompt {b} has length {n}; must be in 1..{self.max_ctx - 1}")
            need = (min(n + sp.max_new_tokens, self.max_ctx) + self.PAGE - 1) // self.PAGE
            pages = self.pool.alloc(need)
            self._slot_pages[b] = pages
            table[b, :need] = pages
        cu = np.zeros(B + 1, dtype=np.int32)
        cu[1:] = np.cumsum(lens)
        with torch.cuda.device(self.device):
            ids = ids_host.to(self.device, non_blocking=True)
            if self._table_host is None or not np.array_equal(self._table_host, table):
                self.page_table.copy_(torch.from_numpy(table))
                self._table_host = table
            self.n_generated.zero_()
            self.done.zero_()
            logits = torch.empty(B, self.shape.vocab_size, dtype=torch.float32, device=self.device) if return_logits else None
            _lib.check(self.L.nt_lm_prefill(self.handle, C.byref(self.state), ids.data_ptr(),
                                            cu.ctypes.data_as(C.POINTER(C.c_int32)), B, C.byref(sp),
                                            logits.data_ptr() if return_logits else None, _lib.current_stream_ptr()))
        self._B = B
        return logits

    def decode(self, n_steps: int, sp, return_logits: bool = False):
        B = self._B
        with torch.cuda.device(self.device):
            logits = (torch.empty(n_steps, B, self.shape.vocab_size, dtype=torch.float32, device=self.device)
                      if return_logits else None)
            _lib.check(self.L.nt_lm_decode(self.handle, C.byref(self.state), B, n_steps, C.byref(sp),
                                           logits.data_ptr() if return_logits else None, _lib.current_stream_ptr()))
        return logits

    def head_gemv(self, h: torch.Tensor) -> torch.Tensor:
        """lm_head (+ final RMSNorm) GEMV alone: the kernel bench.py puts on the roofline."""
        B = h.shape[0]
        logits = torch.empty(B, self.shape.vocab_size, dtype=torch.float32, device=self.device)
        _lib.check(self.L.nt_lm_head_gemv(self.handle, h.data_ptr(), B, logits.data_ptr(), _lib.current_stream_ptr()))
        return logits

    # ------------------------------------------------------------------ per-stage parity hooks (tests)
    def debug_set_layers(self, n: int) -> None:
        _lib.check(self.L.nt_lm_debug_set_layers(self.handle, n))

    def debug_buffer(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        """View of an internal activation buffer (lives inside ``self.workspace``)."""
        ptr = self.L.nt_lm_debug_ptr(self.handle, name.encode())
        if not ptr:
            raise KeyError(name)
        off = ptr - self.workspace.data_ptr()
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        return self.workspace[off: off + n].view(dtype).view(*shape)

    # ------------------------------------------------------------------ generation
    def generate_batch(self, prompts, eos_token_id: int, max_length: int | None = None, min_new_tokens: int = 50,
                       temperature: float = 1.0, top_k: int = 50, max_new_tokens: int | None = None, seed: int = 0,
                       greedy: bool = False, forced: torch.Tensor | None = None, check_every: int = 64, slot_base: int = 0):
        """Returns a list of int64 CPU tensors with the generated ids of each prompt (EOS included
        when it was sampled), following transformers' stopping rules (stopping_criteria.py:73-84,
        467-471): stop at EOS or when prompt + generated reaches max_length."""
        max_length = max_length or self.max_ctx
        if max_length > self.max_ctx:
            raise ValueError(f"max_length {max_length} exceeds the engine context {self.max_ctx}")
        lens = [len(p) for p in prompts]
        if min(max_length - n for n in lens) < 1:
            raise ValueError("prompt already at max_length")
        # max_length counts prompt + generated PER SEQUENCE (stopping_criteria.py:73-84): a long prompt in the batch
        # must not shorten its neighbours, so every slot gets its own cap and the loop runs to the largest one
        caps = [min(max_length - n, max_new_tokens or max_length, self.max_new) for n in lens]
        limit = max(caps)
        sp = self.sampling(eos_token_id, min_new_tokens, limit, top_k, temperature, seed, greedy, forced,
                           limits=caps if min(caps) < limit else None, slot_base=slot_base)
        self.prefill(prompts, sp)
        remaining = limit - 1
        B = len(prompts)
        while remaining > 0:
            n = min(check_every, remaining)
            self.decode(n, sp)
            remaining -= n
            if remaining > 0 and bool(self.done[:B].all()):  # one small D2H read per `check_every` steps
                break
        ngen = self.n_generated[:B].cpu()
        toks = self.out_tokens[:B].cpu()
        return [toks[b, : int(ngen[b])].long() for b in range(B)]

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, max_length: int = 2048, eos_token_id: int | None = None,
                 do_sample: bool = True, temperature: float = 1.0, top_k: int = 50, use_cache: bool = True,
                 min_new_tokens: int = 0, max_new_tokens: int | None = None, seed: int | None = None, **_):
        """transformers-compatible seam used by ``NeuTTS._infer_torch`` (neutts/neutts.py:338-347)."""
        if eos_token_id is None:
            raise ValueError("eos_token_id is required")
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [B, P]")
        prompts = [row.tolist() for row in input_ids.cpu()]
        if seed is None:
            seed = int(torch.randint(0, 2**31 - 1, (1,)).item())  # reference sampling is unseeded
        outs = self.generate_batch(prompts, eos_token_id, max_length, min_new_tokens, temperature, top_k,
                                   max_new_tokens, seed, greedy=not do_sample)
        n = max(len(o) for o in outs)
        res = torch.full((len(outs), input_ids.shape[1] + n), int(eos_token_id), dtype=torch.long)
        for b, o in enumerate(outs):
            res[b, : input_ids.shape[1]] = input_ids[b].cpu()
            res[b, input_ids.shape[1]: input_ids.shape[1] + len(o)] = o
        return res.to(input_ids.device)

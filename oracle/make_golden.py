"""Generates tests/golden/*.npz.  Run in THIS container (needs `transformers`); the fixtures are
committed so the GPU box and CI never need to regenerate them.

  lm_tiny.npz     logits of the REAL transformers.Qwen2ForCausalLM (eager attention, fp32) on seeded
                  weights: a 24-token prefill, 6 cached decode steps, and the processed sampling
                  distribution of the HF logits processors -> pins oracle/lm_oracle.py.
  codec_tiny.npz  PCM + stage activations of oracle/codec_oracle.py itself on seeded weights
                  (self-pinning only: neucodec is not available offline, parity stays "unpinned").
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import codec_oracle as CO  # noqa: E402
from oracle import lm_oracle as LO  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def weights_digest(w) -> str:
    h = hashlib.sha256()
    for t in [w.embed, w.final_norm] + [L[k] for L in w.layers for k in sorted(L)]:
        h.update(t.numpy().tobytes())
    return h.hexdigest()


def main():
    import transformers
    from transformers.generation.logits_process import (MinNewTokensLengthLogitsProcessor, TemperatureLogitsWarper,
                                                          TopKLogitsWarper)

    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    cfg = LO.LMConfig.tiny(num_heads=4, num_kv_heads=2, hidden_size=256, intermediate_size=384)
    w = LO.random_weights(cfg, seed=7, std=0.06)
    model = LO.to_hf_model(cfg, w, "eager")
    g = torch.Generator().manual_seed(3)
    prompt = torch.randint(0, cfg.vocab_size, (24,), generator=g)
    steps = torch.randint(0, cfg.vocab_size, (6,), generator=g)
    with torch.no_grad():
        out = model(prompt[None], use_cache=True)
        prefill = out.logits[0].clone()
        pkv, dec = out.past_key_values, []
        for t in steps:
            o = model(t.view(1, 1), past_key_values=pkv, use_cache=True)
            pkv = o.past_key_values
            dec.append(o.logits[0, -1].clone())
    dec = torch.stack(dec)
    # HF processors in generate()'s order (generation/utils.py:1134,1214,1219) on the last prefill row
    eos, P = 17, prompt.shape[0]
    scores = prefill[-1][None].clone()
    ids = torch.cat((prompt, torch.zeros(3, dtype=torch.long)))[None]   # 3 tokens generated so far
    procs = [MinNewTokensLengthLogitsProcessor(P, 5, eos, device="cpu"), TemperatureLogitsWarper(0.8), TopKLogitsWarper(top_k=50)]
    for p in procs:
        scores = p(ids, scores)
    probs = torch.softmax(scores, -1)[0]
    np.savez_compressed(
        os.path.join(OUT, "lm_tiny.npz"), cfg=np.array([cfg.vocab_size, cfg.hidden_size, cfg.intermediate_size, cfg.num_layers,
                                                        cfg.num_heads, cfg.num_kv_heads, cfg.head_dim]),
        seed=7, std=0.06, rope_theta=cfg.rope_theta, rms_eps=cfg.rms_eps, prompt=prompt.numpy(), steps=steps.numpy(),
        prefill_logits=prefill.numpy(), decode_logits=dec.numpy(), weights_sha256=weights_digest(w),
        proc_eos=eos, proc_generated=3, proc_min_new=5, proc_temperature=0.8, proc_topk=50, proc_probs=probs.numpy(),
        transformers_version=transformers.__version__, torch_version=torch.__version__)

    ccfg = CO.CodecConfig.tiny()
    cw = CO.random_weights(ccfg, 5)
    codes = torch.randint(0, ccfg.codebook_size, (2, 1, 40), generator=torch.Generator().manual_seed(11))
    col = {}
    with torch.no_grad():
        pcm = CO.decode_code(codes, cw, ccfg, col)
    np.savez_compressed(os.path.join(OUT, "codec_tiny.npz"), codes=codes.numpy(), pcm=pcm.numpy(), seed=5,
                        fc_post_a=col["fc_post_a"].numpy(), prior=col["prior"].numpy(), final=col["final"].numpy())
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()

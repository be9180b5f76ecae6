"""Small batched decode runs for compute-sanitizer (profiles/run_sanitizer.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import lm_oracle as LO
from tests.helpers import make_lm

cfg = LO.LMConfig.tiny(vocab_size=2048, hidden_size=256, intermediate_size=512, num_layers=2, num_heads=4, num_kv_heads=2)
w = LO.random_weights(cfg, 0, std=0.05, bf16_round=True)
g = torch.Generator().manual_seed(0)
for B, impl in ((6, "tc"), (3, "tc"), (12, "perop")):
    os.environ["NT_DECODE_IMPL"] = impl
    lm = make_lm(cfg, w, max_batch=B, max_ctx=256, max_new=16)
    prompts = [torch.randint(0, cfg.vocab_size, (20 + 13 * b,), generator=g).tolist() for b in range(B)]
    sp = lm.sampling(cfg.vocab_size - 1, min_new_tokens=2, max_new_tokens=8, seed=3)
    lm.prefill(prompts, sp)
    lm.decode(6, sp)
    torch.cuda.synchronize()
    print("sanitizer run ok:", B, impl, lm.n_generated[:B].tolist())
    del lm

"""Bring-up / parity probe of the persistent tcgen05 decode kernel (lm_decode_tc.cu) against the CPU oracle.

    python profiles/probe_tc.py <cfg: small|wide|nano> <B> <steps> [layers]

Prints the relative RMS error of the teacher-forced decode logits vs the mirrored oracle and the pure-fp32 reference
per slot and per step.  One config per process: a device-side trap poisons the CUDA context.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import lm_oracle as O
from tests.helpers import make_lm, max_err, rel_err

CFGS = dict(
    small=dict(vocab_size=4096, hidden_size=256, intermediate_size=640, num_layers=3, num_heads=4, num_kv_heads=2),
    wide=dict(vocab_size=8192, hidden_size=896, intermediate_size=4864, num_layers=2, num_heads=14, num_kv_heads=2),
    nano=dict(vocab_size=16384, hidden_size=576, intermediate_size=1536, num_layers=4, num_heads=9, num_kv_heads=3),
)


def main():
    name, B, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    kw = dict(CFGS[name])
    if len(sys.argv) > 4:
        kw["num_layers"] = int(sys.argv[4])
    torch.set_num_threads(16)
    cfg = O.LMConfig.tiny(**kw)
    w = O.random_weights(cfg, 13, std=0.05, bf16_round=True)
    lm = make_lm(cfg, w, max_batch=B, max_ctx=512, page_shuffle_seed=5)
    g = torch.Generator().manual_seed(6)
    lens = [(37 * (i + 1)) % 190 + 3 for i in range(B)]
    lens[0] = 130
    prompts = [torch.randint(0, cfg.vocab_size, (n,), generator=g) for n in lens]
    forced = torch.randint(0, cfg.vocab_size, (B, steps + 1), generator=g)
    eos = cfg.vocab_size - 1
    sp = lm.sampling(eos, min_new_tokens=0, max_new_tokens=steps + 1, forced=forced)
    lm.prefill([p.tolist() for p in prompts], sp, return_logits=True)
    ls = lm.decode(steps, sp, return_logits=True)
    torch.cuda.synchronize()
    got = ls.permute(1, 0, 2).float().cpu()      # [B, steps, V]
    dm = "decode" if B <= 8 else "batched"
    worst = 0.0
    for b in range(min(B, 6)):
        _, mir = O.generate(cfg, w, prompts[b], eos, max_length=512, max_new_tokens=steps + 1, forced=forced[b], mirror=True,
                            decode_mirror=dm)
        _, ref = O.generate(cfg, w, prompts[b], eos, max_length=512, max_new_tokens=steps + 1, forced=forced[b], mirror=False)
        rm, rr = rel_err(got[b], mir[1:]), rel_err(got[b], ref[1:])
        per = [f"{rel_err(got[b, i], mir[1 + i]):.1e}" for i in range(steps)]
        print(f"TC-PROBE {name} B={B} slot {b} (P={lens[b]}): relRMS mirrored {rm:.2e} reference {rr:.2e} "
              f"max/spread {max_err(got[b], ref[1:]) / float(ref.std()):.2e} per-step {per}")
        worst = max(worst, rm)
    ok_state = (lm.out_tokens[:B, : steps + 1].cpu().tolist() == forced.tolist()
                and lm.seq_lens[:B].cpu().tolist() == [n + steps for n in lens])
    print(f"TC-PROBE {name} B={B}: worst mirrored relRMS {worst:.2e}; state machine ok: {ok_state}")


if __name__ == "__main__":
    main()

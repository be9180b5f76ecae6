"""Decode-loop timing of the full NeuTTS-Air shape at one batch size, per implementation, plus the in-kernel
timeline of the persistent tcgen05 kernel (grid-barrier release times seen by CTA 0).

    python profiles/perf_tc.py <B> [impl: tc|mega|perop] [timeline: 0|1]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
impl = sys.argv[2] if len(sys.argv) > 2 else "tc"
os.environ["NT_DECODE_IMPL"] = impl
import torch

from neutts_air_b200 import synthetic
from neutts_air_b200.lm import LMShape, SpeechLM

B = int(sys.argv[1])
timeline = len(sys.argv) > 3 and sys.argv[3] == "1"
P, N = 500, 250
shape = LMShape()
lm = SpeechLM(shape, synthetic.lm_state_dict(shape, 0), device="cuda:0", max_batch=B, max_ctx=2048, max_new=256,
              max_prefill_tokens=B * P)
g = torch.Generator().manual_seed(1)
prompts = [torch.randint(0, 151643, (P,), generator=g).tolist() for _ in range(B)]
eos = 151670
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for it in range(3):
    sp = lm.sampling(eos, min_new_tokens=N, max_new_tokens=N, seed=it)
    lm.prefill(prompts, sp)
    torch.cuda.synchronize()
    e0.record()
    lm.decode(N - 1, sp)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
ms = best / (N - 1)
bytes_step = 1105.5e6 + B * (12288 * (P + N / 2 + 1) + 1792)
print(f"PERF impl={impl} B={B}: {ms * 1e3:.1f} us/step, {B / ms * 1e3:.0f} tok/s, {bytes_step / ms / 1e6:.0f} GB/s "
      f"(frac of 6572.9: {bytes_step / ms / 1e6 / 6572.9:.3f}), n_generated {lm.n_generated[:B].tolist()[:4]}")

if timeline and impl == "tc":
    buf = torch.zeros(2 * 1024, dtype=torch.int64, device="cuda:0")
    lm.L.nt_lm_debug_set_profile(lm.handle, buf.data_ptr(), 20)
    sp = lm.sampling(eos, min_new_tokens=N, max_new_tokens=N, seed=9)
    lm.prefill(prompts, sp)
    lm.decode(40, sp)
    torch.cuda.synchronize()
    t = buf[:1024].cpu().tolist()
    t = [x for x in t if x]
    d = [(t[i + 1] - t[i]) / 1e3 for i in range(len(t) - 1)]
    L = shape.num_layers
    fold_cta = B <= 4 and os.environ.get("NT_TC_FOLD", "")[:1] != "p"
    if fold_cta:
        names = ["qkv", "attn", "o", "gu", "down"]
        per, off = 5, 0
    else:
        names = ["qkv", "attn", "o", "fold2", "gu", "down", "fold1"]
        per, off = 7, 1
    print(f"TIMELINE marks={len(t)} step total {(t[-1] - t[0]) / 1e3:.1f} us")
    if off:
        print(f"  first fold: {d[0]:.2f} us")
    acc = [0.0] * per
    for l in range(L):
        for j in range(per):
            acc[j] += d[off + l * per + j]
    print("  per layer (avg us): " + ", ".join(f"{n} {a / L:.2f}" for n, a in zip(names, acc)) + f" | sum {sum(acc) / L:.2f}")
    rest = d[off + L * per:]
    print("  tail (lm_head, sampler): " + ", ".join(f"{x:.2f}" for x in rest))

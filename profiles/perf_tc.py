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
    names = {0: "step start", 1: "fold: start", 2: "fold: polled", 4: "fold: staged", 5: "fold: bop ready", 10: "epilogue: accumulator ready",
             21: "attn: page issued", 22: "attn: qkv polled", 23: "attn: page landed", 25: "attn: stored", 31: "merge: polled+staged",
             32: "merge: bop ready", 41: "act: polled+staged", 42: "act: bop ready", 50: "swiglu: accumulator ready",
             105: "PHASE fold(o) done", 106: "PHASE fold(down) done", 120: "sampler: start", 126: "sampler: maxima loaded", 127: "sampler: maxima shared", 128: "sampler: ranked maxima", 121: "sampler: threshold", 122: "sampler: tiles listed", 123: "sampler: candidates gathered", 124: "sampler: ranked", 125: "sampler: finished", 110: "PHASE lm_head done", 111: "PHASE sampler done", 100: "PHASE qkv done", 101: "PHASE attention done", 102: "PHASE o_proj done", 103: "PHASE gate/up done", 104: "PHASE down done",
             200: "grid barrier released"}
    for cta in (0, 1):
        raw = buf[cta * 1024: cta * 1024 + 1024].cpu().tolist()
        ts, ids = raw[:512], raw[512:]
        n = sum(1 for x in ts if x)
        ts, ids = ts[:n], ids[:n]
        print(f"TIMELINE cta {cta}: {n} marks, step total {(ts[-1] - ts[0]) / 1e3:.1f} us")
        L = shape.num_layers
        # per-phase averages from the phase marks
        acc, last = {}, ts[0]
        for tt, i in zip(ts, ids):
            if i >= 100:
                acc.setdefault(i, []).append((tt - last) / 1e3)
                last = tt
        print("  avg us between phase marks: " + ", ".join(f"{names.get(i, i)} {sum(v) / len(v):.2f} (x{len(v)})" for i, v in sorted(acc.items())))
        # the fine-grained layer
        fine = [(tt, i) for tt, i in zip(ts, ids)]
        start = None
        for j, (tt, i) in enumerate(fine):
            if i in (1, 21, 31, 41, 50, 10) and start is None and any(k[1] < 100 for k in fine[j:j + 2]):
                start = j
                break
        if start is not None:
            j = start
            t0 = fine[j - 1][0] if j > 0 else fine[j][0]
            print("  fine-grained layer (us since previous phase mark):")
            while j < len(fine) and not (fine[j][1] == 104):
                print(f"    +{(fine[j][0] - t0) / 1e3:7.2f}  {names.get(fine[j][1], fine[j][1])}")
                j += 1
                if j - start > 60:
                    break
            if j < len(fine):
                print(f"    +{(fine[j][0] - t0) / 1e3:7.2f}  {names.get(fine[j][1], fine[j][1])}")

        samp = [(tt, i) for tt, i in zip(ts, ids) if 120 <= i <= 128 or i == 111]
        if samp:
            print("  sampler (us since start): " + ", ".join(f"{names.get(i, i)} +{(tt - samp[0][0]) / 1e3:.2f}" for tt, i in samp))

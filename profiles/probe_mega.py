"""In-kernel timeline of one decode step of the persistent megakernel (first and last CTA).
Marks per GEMV phase: [x staged] ... then per grid barrier: [CTA done] [arrival published] [released]."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neutts_air_b200 import synthetic, _lib
from neutts_air_b200.lm import LMShape, SpeechLM

layers = int(os.environ.get("LAYERS", "24"))
shape = LMShape(num_layers=layers)
lm = SpeechLM(shape, synthetic.lm_state_dict(shape, 0), device="cuda:0", max_batch=1, max_ctx=2048, max_new=64, max_prefill_tokens=512)
buf = torch.zeros(2, 1024, dtype=torch.int64, device="cuda:0")
_lib.check(lm.L.nt_lm_debug_set_profile(lm.handle, buf.data_ptr(), 10))
print("engine built", flush=True)
sp = lm.sampling(151670, min_new_tokens=64, max_new_tokens=40)
lm.prefill([list(range(100, 600))], sp)
torch.cuda.synchronize()
print("prefill done", flush=True)
lm.decode(30, sp)
torch.cuda.synchronize()
print("decode done", flush=True)
t = buf.cpu().numpy()
for c in range(2):
    m = t[c][t[c] > 0]
    d = np.diff(m) / 1000.0
    print(f"CTA {'first' if c == 0 else 'last'}: {len(m)} marks, step total {(m[-1]-m[0])/1000:.1f} us")
    # layout: mark0 = step start; per layer: (x, done, arr, rel) qkv | (done, arr, rel) attn | (x,done,arr,rel) o | gu | d
    i = 1
    names = []
    # marks per GEMV phase: x staged, first stage landed, (last stage landed if > 1 stage), CTA done, arrival published, released
    nst = {"qkv": 1, "o": 1, "gu": 5, "d": 3, "head": 92}
    def gemv(prefix, ph):
        out = [f"{prefix}{ph}.x", f"{prefix}{ph}.w0"]
        if nst[ph] > 1:
            out.append(f"{prefix}{ph}.wN")
        return out + [f"{prefix}{ph}.done", f"{prefix}{ph}.arr", f"{prefix}{ph}.rel"]
    for l in range(layers):
        names += gemv(f"L{l}.", "qkv") + [f"L{l}.att.done", f"L{l}.att.arr", f"L{l}.att.rel"]
        for ph in ("o", "gu", "d"):
            names += gemv(f"L{l}.", ph)
    names += gemv("", "head") + ["final.keys", "final.select", "final.gather", "final.sort", "final.done", "final.arr", "final.rel"]
    agg = {}
    for nme, dt in zip(names, d):
        key = nme.split(".", 1)[1] if nme.startswith("L") else nme
        agg.setdefault(key, []).append(dt)
    print({k: round(float(np.mean(v)), 2) for k, v in agg.items()})

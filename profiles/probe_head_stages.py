"""lm_head GEMV (390 MB of bf16 weights per launch) timed with CUDA events for ring depths NT_GEMV_STAGES=2..7.
Each depth runs in its own process (the plan is read at launch time; the attribute cache is per process)."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from neutts_air_b200 import synthetic
from neutts_air_b200.lm import LMShape, SpeechLM
shape = LMShape(num_layers=1)
lm = SpeechLM(shape, synthetic.lm_state_dict(shape, 0), device="cuda:0", max_batch=1, max_ctx=2048, max_new=8, max_prefill_tokens=64)
h = torch.randn(1, shape.hidden_size, device="cuda:0")
for _ in range(5): lm.head_gemv(h)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(50): lm.head_gemv(h)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 50
print("stages", os.environ.get("NT_GEMV_STAGES", "default"), "us/launch %%.1f  TB/s %%.2f" %% (us, shape.vocab_size * shape.hidden_size * 2 / us / 1e6))
''' % ROOT
for n in sys.argv[1:] or ["3", "4", "5", "6", "7"]:
    subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, NT_GEMV_STAGES=n), check=False)

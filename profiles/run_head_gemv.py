"""Launches the dominant decode kernel (lm_head GEMV + fused final RMSNorm, 390 MB of bf16 weights)
a few times so ncu can capture one launch:  ncu --set full -k regex:gemv_kernel -s 2 -c 1 python profiles/run_head_gemv.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neutts_air_b200 import synthetic
from neutts_air_b200.lm import LMShape, SpeechLM

shape = LMShape()
sd = synthetic.lm_state_dict(LMShape(num_layers=1), 0)   # one layer is enough: only embed / lm_head are touched
lm = SpeechLM(LMShape(num_layers=1), sd, device="cuda:0", max_batch=1, max_ctx=2048, max_new=8, max_prefill_tokens=64)
h = torch.randn(1, shape.hidden_size, device="cuda:0")
for _ in range(5):
    out = lm.head_gemv(h)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))

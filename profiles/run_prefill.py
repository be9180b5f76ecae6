"""One prefill (BATCH x 500 tokens) between cudaProfilerStart/Stop so that
`ncu --profile-from-start off --metrics gpu__time_duration.sum` lists exactly its launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neutts_air_b200 import synthetic
from neutts_air_b200.lm import LMShape, SpeechLM

B = int(os.environ.get("BATCH", "1"))
shape = LMShape()
lm = SpeechLM(shape, synthetic.lm_state_dict(shape, 0), device="cuda:0", max_batch=B, max_ctx=2048, max_new=32, max_prefill_tokens=B * 512)
g = torch.Generator().manual_seed(0)
prompts = [torch.randint(0, 151643, (500,), generator=g).tolist() for _ in range(B)]
sp = lm.sampling(151670, min_new_tokens=32, max_new_tokens=16)
lm.prefill(prompts, sp)
torch.cuda.synchronize()
torch.cuda.profiler.start()
lm.prefill(prompts, sp)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")

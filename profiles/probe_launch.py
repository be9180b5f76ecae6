"""Launch-latency probe on the B200: cost of one *dependent* kernel in a chain, eager vs CUDA graph,
with and without programmatic dependent launch (set NT_NO_PDL=1 for the latter)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neutts_air_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda:0")
ctr = torch.zeros(1, dtype=torch.int32, device=dev)
res = {"pdl": not bool(os.environ.get("NT_NO_PDL"))}
for grid, block in ((148, 288), (64, 256), (1, 32)):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            L.nt_debug_launch_chain(200, grid, block, ctr.data_ptr(), s.cuda_stream)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        L.nt_debug_launch_chain(1000, grid, block, ctr.data_ptr(), s.cuda_stream)
        e1.record(s)
        s.synchronize()
        eager = e0.elapsed_time(e1)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            L.nt_debug_launch_chain(1000, grid, block, ctr.data_ptr(), s.cuda_stream)
        g.replay(); s.synchronize()
        e0.record(s)
        g.replay()
        e1.record(s)
        s.synchronize()
        res[f"{grid}x{block}"] = {"eager_us_per_kernel": eager, "graph_us_per_kernel": e0.elapsed_time(e1)}
print(json.dumps(res))

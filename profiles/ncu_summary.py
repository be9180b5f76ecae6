"""Text summary of one `ncu --set full` capture of the persistent decode kernel (profiles/*_ncu_summary.txt).
usage: python profiles/ncu_summary.py <report.ncu-rep> <batch> <steps> <algorithmic GB per launch> > summary.txt"""
import csv
import subprocess
import sys

rep, batch, steps, alg_gb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
g = lambda k: d[k][0]
f = lambda k: float(g(k).replace(",", ""))
stalls = sorted([(float(v[0] or 0), h) for h, v in d.items() if "average_warps_issue_stalled" in h and "per_issue_active" in h], reverse=True)
rd, wr, dur = f("dram__bytes_read.sum"), f("dram__bytes_write.sum"), f("gpu__time_duration.sum")
ru, du = d["dram__bytes_read.sum"][1], d["gpu__time_duration.sum"][1]
rd_gb = rd * {"Gbyte": 1, "Mbyte": 1e-3, "Tbyte": 1e3}.get(ru, 1)
wr_gb = wr * {"Gbyte": 1, "Mbyte": 1e-3, "Tbyte": 1e3}.get(d["dram__bytes_write.sum"][1], 1)
dur_ms = dur * {"ms": 1, "us": 1e-3, "s": 1e3}.get(du, 1)
print(f"""ncu --set full --import-source on --clock-control none -k regex:decode_tc_kernel --launch-skip 1 --launch-count 1 python profiles/perf_tc.py {batch} tc 0
kernel: {g('Kernel Name') if 'Kernel Name' in d else 'decode_tc_kernel'}  (batch {batch}, persistent: {steps} decode steps of NeuTTS-Air in ONE launch)
grid {g('launch__grid_size')} x {g('launch__block_size')} threads, {g('launch__registers_per_thread')} registers/thread, {g('launch__shared_mem_per_block_dynamic')} KB dynamic shared memory, 1 CTA / SM

gpu__time_duration.sum            {dur_ms:.2f} ms   -> {dur_ms / steps * 1000:.1f} us per decode step (under ncu, clocks not locked)
dram__bytes_read.sum              {rd_gb:.2f} GB   (per launch)
dram__bytes_write.sum             {wr_gb:.3f} GB
  algorithmic bytes per launch    {alg_gb:.1f} GB  => traffic / algorithmic = {(rd_gb + wr_gb) / alg_gb:.3f}
dram read throughput vs MEASURED_PEAKS 6572.9 GB/s: {rd_gb / (dur_ms / 1e3) / 6572.9:.3f}
lts__t_sector_hit_rate            {f('lts__t_sector_hit_rate.pct'):.1f} %
sm__pipe_tensor_cycles_active     {f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.2f} %
smsp__issue_active                {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} %
smsp__inst_executed.sum           {f('smsp__inst_executed.sum'):.3e}

warp stall reasons (average warps stalled per issue-active cycle; 10 warps / SM):""")
for v, h in stalls[:10]:
    print(f"  {v:7.3f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")

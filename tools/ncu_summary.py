"""Summarise an .ncu-rep (ncu --set full), or the CSV of its raw page (`ncu -i x.ncu-rep --page raw --csv`, made on the GPU box
when the report itself is too large to bring back), into the handful of numbers DESIGN.md / bench.py quote.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep|raw.csv [label ...] > profiles/r02_prof.txt
Optional labels name the launches in order (e.g. the layer shapes tools/prof_layers_r02.py prints)."""
import csv, io, subprocess, sys

rep = sys.argv[1]
labels = sys.argv[2:]
if rep.endswith(".csv"):
    raw = "".join(l for l in open(rep) if not l.startswith("=="))
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
WANT = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "sm cycles"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma inst %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (LSU)"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/shared throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"),
]
print(f"# {rep}: {len(data)} kernel launches captured with ncu --set full --clock-control none")
for n, d in enumerate(data):
    print("-" * 100)
    if n < len(labels):
        print(f"{'launch':28s}: {labels[n]}")
    for key, label in WANT:
        if key in idx:
            print(f"{label:28s}: {d[idx[key]]} {units[idx[key]]}")

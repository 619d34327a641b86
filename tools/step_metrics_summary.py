"""Summarise an ncu metrics CSV of tools/one_step.py (2 eager steps; the second one is kept) into
profiles/<tag>_step_metrics_summary.json: per-kernel time / DRAM traffic / L2 bytes, and the causal-conv3d class that
bench.py's roofline.traffic quotes.
Capture:  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\\
              sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum \\
              --clock-control none --csv --log-file gpurun_out/step_metrics.csv python tools/one_step.py 4 2
Usage:    python tools/step_metrics_summary.py gpurun_out/step_metrics.csv profiles/r01_step_metrics_summary.json"""
import collections, csv, json, re, sys

src, dst = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per, order = {}, []
for r in rows:
    i = int(r["ID"])
    if i not in per:
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", r["Kernel Name"]).replace("mv2::", "").replace("void ", ""))
        per[i] = {"kernel": name, "grid": r["Grid Size"]}
        order.append(i)
    v = float(r["Metric Value"].replace(",", ""))
    u, n = r["Metric Unit"], r["Metric Name"]
    if "byte" in u:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    if n == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)
    per[i][n] = v
# second step only: a step starts at its (single) ingest launch; everything before the last one is set-up + step 1
starts = [k for k, i in enumerate(order) if per[i]["kernel"] == "ingest_kwpack_kernel"]
order = order[starts[-1]:]
TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
agg = collections.OrderedDict()
for i in order:
    p = per[i]
    a = agg.setdefault(p["kernel"], dict(n=0, us=0.0, rd=0.0, wr=0.0, l2=0.0))
    a["n"] += 1; a["us"] += p["gpu__time_duration.sum"]; a["rd"] += p["dram__bytes_read.sum"]
    a["wr"] += p["dram__bytes_write.sum"]; a["l2"] += p["lts__t_bytes.sum"]
out, tot = [], dict(us=0.0, b=0.0)
print(f"{'kernel':34s} {'n':>4s} {'ms':>8s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'GB/s':>8s} {'L2 MB':>9s}")
for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
    gbs = (a["rd"] + a["wr"]) / a["us"] / 1e3
    print(f"{k[:34]:34s} {a['n']:4d} {a['us'] / 1e3:8.3f} {a['rd'] / 1e6:11.1f} {a['wr'] / 1e6:11.1f} {gbs:8.0f} {a['l2'] / 1e6:9.0f}")
    tot["us"] += a["us"]; tot["b"] += a["rd"] + a["wr"]
    out.append(dict(kernel=k, launches=a["n"], ms=a["us"] / 1e3, dram_read_MB=a["rd"] / 1e6, dram_write_MB=a["wr"] / 1e6,
                    dram_GBps=gbs, l2_MB=a["l2"] / 1e6))
# the causal 3x3x3 conv launches: slab launches that keep the tensor pipe busy (pointwise / FF launches stay below 30 %)
c3 = [per[i] for i in order if per[i]["kernel"] == "tc_slab_kernel" and per[i][TP] > 30]
summary = dict(per_kernel=out, step_ms=tot["us"] / 1e3, step_dram_GB=tot["b"] / 1e9,
               conv3d=dict(launches=len(c3),
                           avg_dram_bytes=sum(p["dram__bytes_read.sum"] + p["dram__bytes_write.sum"] for p in c3) / max(len(c3), 1),
                           avg_tensor_pipe_pct=sum(p[TP] for p in c3) / max(len(c3), 1)))
print("step:", round(summary["step_ms"], 3), "ms (cold, serialised)", round(summary["step_dram_GB"], 2), "GB DRAM;  conv3d:", summary["conv3d"])
json.dump(summary, open(dst, "w"), indent=1)

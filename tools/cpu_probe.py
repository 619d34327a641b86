import os, time, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
os.system("lscpu | grep -E 'Model name|^CPU\\(s\\)|Thread|Socket' | cut -c1-200 | head -6")
x = torch.randn(1, 64, 20, 64, 64)
w = torch.randn(64, 64, 3, 3, 3)
for nt in (4, 8, 16, 32, 64, 128):
    torch.set_num_threads(nt)
    torch.nn.functional.conv3d(x, w, padding=1)
    t = time.perf_counter()
    for _ in range(3): torch.nn.functional.conv3d(x, w, padding=1)
    dt = (time.perf_counter() - t) / 3
    print(f"threads {nt:4d}: conv3d fp32 {dt*1e3:8.1f} ms  {2*64*64*27*20*64*64/dt/1e9:8.1f} GFLOP/s", flush=True)

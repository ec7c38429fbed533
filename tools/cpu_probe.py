"""What the GPU box's host really offers to the CPU legs of bench.py (run on the box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "host_threads", b.host_threads())
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/proc/loadavg"):
    try:
        print(p, open(p).read().strip())
    except OSError as e:
        print(p, "n/a")
for th in (b.host_threads(), 32, 16):
    t0 = time.perf_counter(); v, times = b.cpu_oracle_throughput(1, th, 2); print("threads", th, "2 proposals:", round(times[0], 2), "s ->", round(v, 3), "poses/s", flush=True)

import sys, os
sys.path.insert(0, os.getcwd())
import bench
if __name__ == "__main__":
    procs_list = [int(a) for a in sys.argv[1:]] or [1, 4, 8]
    for procs in procs_list:
        v, n, dt = bench.cpu_throughput(procs * 30, procs)
        print(procs, 'procs:', round(v, 1), 'frames/s', round(v / procs, 1), 'per proc', n, 'frames', round(dt, 1), 's', flush=True)
    print("os.cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "usable", bench.usable_cpus())
    for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try: print(f, open(f).read().strip())
        except Exception as e: print(f, "n/a")

import time, os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, ".")
import numpy as np
import bench
from oracle import ptr as optr
print(open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else "no cpu.max", "| loadavg", open("/proc/loadavg").read().strip(), "| cpu_count", os.cpu_count(), "| affinity", len(os.sched_getaffinity(0)), flush=True)
pb, g = bench.oracle_base_guess(100)
sc = optr.Scaling(pb)
for nproc in (1, 8, 32, 128):
    X, U, P = bench.make_seeds(g, sc.Sx, sc.Su, nproc, 0)
    its, wall, ph, st = bench.cpu_run(100, 100, pb.hs, X, U, P, nproc)
    print("nproc", nproc, "its", its, "wall", round(wall, 1), "it/s", round(its / wall, 2), "solve s/it", round(ph["solve"] / its, 2),
          "formulate s/it", round(ph["formulate"] / its, 3), "disc s/it", round(ph["discretize"] / its, 3), flush=True)

#!/usr/bin/env python
"""Worker-count sweep of bench.py's CPU arm (the reference's algorithm on the host cores): how many
worker processes give the best alignments/s on this box, and what the container's CPU quota is.

    python tools/cpu_sweep.py 8 16 32 64 128 > gpurun_out/cpu_sweep.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ffsubsync_b200.synth import BENCH_RATIOS  # noqa: E402


def read(path):
    try:
        with open(path) as fh:
            return fh.read().strip()
    except OSError:
        return None


def main():
    workers = [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64, 128]
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    info = {"os_cpu_count": os.cpu_count(), "sched_affinity": len(os.sched_getaffinity(0)),
            "cgroup_cpu_max": read("/sys/fs/cgroup/cpu.max"),
            "cgroup_v1_quota": read("/sys/fs/cgroup/cpu/cpu.cfs_quota_us"),
            "cpu_model": next((l.split(":", 1)[1].strip() for l in (read("/proc/cpuinfo") or "").splitlines()
                               if l.startswith("model name")), None),
            "loadavg": read("/proc/loadavg")}
    bench._cpu_setup(BENCH_RATIOS)
    rows = []
    for w in workers:
        if w > (os.cpu_count() or 1):
            continue
        rate, dt = bench.cpu_pass(w, w)
        rows.append({"workers": w, "pairs": w, "alignments_per_s": rate, "seconds": dt,
                     "per_worker_seconds": dt})
        print("workers=%d: %.2f alignments/s (%.1f s)" % (w, rate, dt), file=sys.stderr, flush=True)
    info["sweep"] = rows
    print(json.dumps(info, indent=1))


if __name__ == "__main__":
    main()

"""sharded_set (bench.py) on one GPU under environment variants; prints makespan and the number of graph captures."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time, torch
sys.path.insert(0, %r)
import bench
dev = torch.device("cuda:0")
r = bench.sharded_set(dev, 0, 1, None)
print("RESULT", r["makespan_s"], r["iterations"], r["object_frames"], flush=True)
''' % ROOT
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for name, env, prefix in (("default", {}, []), ("frame batch 1", {"EOSVOS_FRAME_BATCH": "1"}, []),
                          ("host labels", {"EOSVOS_DEVICE_LABELS": "0"}, []),
                          ("default, 8 cores", {}, ["taskset", "-c", "0-7"]),
                          ("host labels, 8 cores", {"EOSVOS_DEVICE_LABELS": "0"}, ["taskset", "-c", "0-7"])):
    out = subprocess.run(prefix + [sys.executable, "-c", CHILD], capture_output=True, text=True,
                         env=dict(os.environ, EOSVOS_DEBUG_GRAPHS="1", **env))
    caps = out.stdout.count("capturing graph")
    res = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    print(f"{name:24s} {res[0] if res else out.stderr[-300:]}  captures {caps}", flush=True)

"""Generate one bench workload on the GPU and print progress + memory (debugging aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
t0 = time.time()
name = sys.argv[1]
print("start", name, flush=True)
try:
    off, keys, K, info = bench.build_workload(name, "cuda:0")
    print("built", info, "%.1fs" % (time.time() - t0), "peak GB %.1f" % (torch.cuda.max_memory_allocated() / 1e9), flush=True)
except Exception as e:   # noqa: BLE001
    print("EXC", repr(e)[:500], "peak GB %.1f" % (torch.cuda.max_memory_allocated() / 1e9), flush=True)

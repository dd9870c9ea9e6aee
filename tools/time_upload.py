"""Pageable-memory upload rate of pipeline._upload_staged against a plain torch copy (FDB_UPLOAD_THREADS sets the thread count)."""
import time, sys, numpy as np, torch
sys.path.insert(0, ".")
from flashdeconv_b200 import pipeline
a = np.random.default_rng(0).random(400_000_000, dtype=np.float32)          # 1.6 GB, pageable
dev = torch.device("cuda:0")
for name, f in (("plain", lambda: torch.from_numpy(a).to(dev)), ("staged", lambda: pipeline._upload_staged(a, torch.float32, dev))):
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = f(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    ok = bool(torch.equal(out[:1000].cpu(), torch.from_numpy(a[:1000]))) and float(out[-1]) == float(a[-1])
    print(f"{name:7s} {1e3 * best:7.1f} ms  {a.nbytes / best / 1e9:5.1f} GB/s  equal={ok}", flush=True)
    del out

"""Times plan_tile_device on a C3-sized kNN graph (single GPU, any simulated world size)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flashdeconv_b200 import pipeline, tiling
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
side = int(n ** 0.5)
g = torch.Generator(device="cuda").manual_seed(0)
xy = torch.stack(torch.meshgrid(torch.arange(side, device="cuda"), torch.arange(side, device="cuda"), indexing="ij"), -1).reshape(-1, 2).double()
xy = xy[:n] + 0.1 * torch.randn(xy[:n].shape, device="cuda", dtype=torch.float64, generator=g)
graph = pipeline.build_graph(xy, "knn", 6)
bounds = tiling.tile_bounds(xy.shape[0], world)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = tiling.plan_tile_device(graph.indptr, graph.indices, graph.nnz, bounds, 1)
    torch.cuda.synchronize(); print("plan_tile_device ms", 1e3 * (time.perf_counter() - t0), "n_halo", p.n_halo, "boundary patches", int(p.n_boundary))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    p = tiling.plan_tile_device(graph.indptr, graph.indices, graph.nnz, bounds, 1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14))

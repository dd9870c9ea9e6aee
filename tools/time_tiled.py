"""torchrun entry (any N): solve-stage time of the tiled path vs the plain device path on a C3-shaped problem of
`n_spots` spots per rank-count -- separates the cost of the exchange from the cost of running on a tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from flashdeconv_b200 import genes, pipeline, tiling
from flashdeconv_b200.synth import make_dataset_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
data = make_dataset_device(n, 18000, 30, 400.0, seed=0, device=f"cuda:{local}")
csr = pipeline.DeviceCSR(data["indptr"], data["indices"], data["data"], (n, 18000))
gene_idx, lev = genes.select_informative_genes_device(csr, data["X"], 2000, 50)
tables = pipeline.build_tables(data["X"], gene_idx, lev, 512, 0, 18000)
def stage_times(path, reps=4):
    out = {}
    for _ in range(reps):
        ev = {}
        path.run_resident(events=ev)
        torch.cuda.synchronize()
        out = {k: round(a.elapsed_time(b), 3) for k, (a, b) in ev.items()}
    return out
tp = tiling.TiledPath(csr, data["coords"], tables, 30)
a = stage_times(tp)
if rank == 0:
    print(f"tiled x{world} [{tp.mode}] n={n} own={tp.plan.n_own} halo={tp.plan.n_halo}:", a)
if world == 1:
    print("device path:", stage_times(pipeline.DevicePath(csr, data["coords"], tables, 30)))
tiling.release_communicators()
dist.destroy_process_group()

"""torchrun entry: the tiled multi-GPU path must reproduce the single-GPU path (and the reference bars).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/check_tiled.py [n_spots]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashdeconv_b200 import genes, pipeline, tiling          # noqa: E402
from flashdeconv_b200.synth import make_dataset               # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ds = make_dataset(n_spots=n, n_genes=1200, n_types=9, depth=300.0, seed=4)
    gene_idx, lev = genes.select_informative_genes(ds.Y, ds.X, 2000, 50)
    tables = pipeline.build_tables(ds.X, gene_idx, lev, 128, 0, ds.Y.shape[1])
    csr = pipeline.csr_to_device(ds.Y)
    coords = torch.from_numpy(ds.coords).cuda()
    tp = tiling.TiledPath(csr, coords, tables, ds.X.shape[0])
    b64, p64, info, lam = tp.run_resident(max_iter=40, gather=True)
    torch.cuda.synchronize()
    single = pipeline.DevicePath(csr, coords, tables, ds.X.shape[0]).run(max_iter=40)
    err = float(np.max(np.abs(p64.cpu().numpy() - single.proportions)))
    berr = float(np.max(np.abs(b64.cpu().numpy() - single.beta)))
    ok = (err <= 1e-5 and info["n_iterations"] == single.info["n_iterations"] == 40
          and abs(info["final_objective"] - single.info["final_objective"]) <= 1e-6 * abs(single.info["final_objective"])
          and abs(lam - single.lambda_used) <= 1e-12 * lam)
    halo = tp.plan.n_halo
    tp.close()
    # the host-input call: every rank uploads 1/R of the rows, H rows travel to their tile's owner over NVLink
    if tp.mode == "peer":
        res = tiling.deconvolve_path_tiled(ds.Y, ds.X, ds.coords, gene_idx, lev, sketch_dim=128, max_iter=40)
        e2 = float(np.max(np.abs(res.proportions - single.proportions)))
        ok = ok and e2 <= 1e-5 and res.info["n_iterations"] == 40 and res.h2d_bytes > 0
        err = max(err, e2)
    tiling.release_communicators()
    flags = torch.tensor([int(ok), halo], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"tiled x{world} [{tp.mode}]: max|dprop|={err:.2e} max|dbeta|={berr:.2e} sweeps={info['n_iterations']} "
              f"obj={info['final_objective']:.6g} vs {single.info['final_objective']:.6g} min_halo={int(flags[1])} "
              f"{'OK' if int(flags[0]) else 'MISMATCH'}")
    dist.destroy_process_group()
    sys.exit(0 if int(flags[0]) else 1)


if __name__ == "__main__":
    main()

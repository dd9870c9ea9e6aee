#!/bin/bash
# memcheck + racecheck of the hot path: smoke-sized whole path, wide-row / bitmap sketch forms, small solves (K = 50, 9, 17),
# 3-D / large-k / grid graphs -- with the persistent sweep grid capped at 5 CTAs so that every CTA walks several patches
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, sys, os
sys.path.insert(0, '.')
import __graft_entry__ as g
g.smoke()
import torch
from scipy import sparse
from flashdeconv_b200 import pipeline as pl
from flashdeconv_b200.solver import bcd_solve
from flashdeconv_b200.graph import build_knn_graph, build_grid_graph, build_radius_graph
rng = np.random.default_rng(3)
for n, K in ((3000, 50), (2500, 9), (1000, 17), (600, 100)):
    Xs = rng.standard_normal((K, 64)) + 0.3
    Ys = (rng.random((n, K)) * (rng.random((n, K)) < 0.3)) @ Xs
    A = build_knn_graph(rng.random((n, 2)), k=6)
    b, info = bcd_solve(Ys, Xs, A, lambda_=0.05, rho=0.01, max_iter=5, tol=1e-12)
    print('solve', n, K, info['n_iterations'], float(b.sum()))
# fused sketch: table form, bitmap form, wide rows, overflowing rows, linear mode
for n, G, K, d, dens, tab in ((300, 3000, 7, 512, 0.3, ""), (400, 2500, 40, 256, 0.1, "0"), (257, 25000, 40, 512, 0.01, ""), (200, 600, 50, 512, 0.9, "")):
    if tab: os.environ["FDB_SKETCH_TAB"] = tab
    else: os.environ.pop("FDB_SKETCH_TAB", None)
    Y = sparse.random(n, G, density=dens, format="csr", random_state=np.random.RandomState(K), data_rvs=lambda s: rng.integers(1, 30, s).astype(np.float64))
    X = rng.random((K, G)) + 0.05
    gi = np.sort(rng.choice(G, size=G // 3, replace=False))
    for mode in ("log_cpm", "raw"):
        tb = pl.build_tables(X, gi, rng.random(gi.size), d, 5, G, preprocess=mode)
        p = pl.DevicePath(pl.csr_to_device(Y), torch.zeros((n, 2), dtype=torch.float64, device="cuda"), tb, K)
        p.stage_sketch(); torch.cuda.synchronize()
        print('sketch', n, G, K, mode, tab, float(p.h.sum()), float(p.ysq.sum()))
c3 = rng.random((700, 3)) * 10
print('graphs', build_knn_graph(c3, 6).nnz, build_knn_graph(rng.random((900, 2)), 40).nnz, build_grid_graph(c3[:, :2] * 3).nnz, build_radius_graph(c3, 1.0).nnz)
PY
for tool in memcheck racecheck; do
  FDB_SWEEP_MAX_CTAS=5 timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|^solve|^sketch|^graphs|Error|hazard" gpurun_out/sanitize_$tool.log | head -24
done

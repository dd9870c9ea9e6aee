#!/bin/bash
# memcheck + racecheck of the hot path on the smoke-sized problem and three small solves (K = 50, 9, 17), with the
# persistent sweep grid capped at 5 CTAs so that every CTA walks several patches (the pipelined path)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
import __graft_entry__ as g
g.smoke()
from flashdeconv_b200.solver import bcd_solve
from flashdeconv_b200.graph import build_knn_graph
rng = np.random.default_rng(3)
for n, K in ((3000, 50), (2500, 9), (1000, 17)):
    Xs = rng.standard_normal((K, 64)) + 0.3
    Ys = (rng.random((n, K)) * (rng.random((n, K)) < 0.3)) @ Xs
    A = build_knn_graph(rng.random((n, 2)), k=6)
    b, info = bcd_solve(Ys, Xs, A, lambda_=0.05, rho=0.01, max_iter=5, tol=1e-12)
    print('solve', n, K, info['n_iterations'], float(b.sum()))
PY
for tool in memcheck racecheck; do
  FDB_SWEEP_MAX_CTAS=5 timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|^solve|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done

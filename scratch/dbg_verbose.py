import numpy as np, sys
sys.path.insert(0, '.')
from flashdeconv_b200.solver import bcd_solve
from flashdeconv_b200.graph import build_knn_graph
rng = np.random.default_rng(0)
Xs, Ys = rng.standard_normal((5, 32)), rng.standard_normal((40, 32))
A = build_knn_graph(rng.random((40, 2)), k=4)
bv, iv = bcd_solve(Ys, Xs, A, lambda_=0.1, rho=0.01, max_iter=12, verbose=True)
print(iv)
b2, i2 = bcd_solve(Ys, Xs, A, lambda_=0.1, rho=0.01, max_iter=12, verbose=False)
print(i2, np.abs(bv-b2).max())

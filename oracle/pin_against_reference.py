"""
oracle/pin_against_reference.py -- pins the oracle against the REAL reference and
writes the golden fixtures under tests/golden/.  TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    NUMBA_CACHE_DIR=/tmp/numba_cache python oracle/pin_against_reference.py

For every case below the script
  1. runs the imported reference stage by stage (core/deconv.py:309-398),
  2. runs oracle/fd_oracle.py on the same inputs,
  3. asserts: gene_idx / buckets / signs / adjacency index arrays bit-exact,
     float64 quantities (leverage, weights, Y_s, X_s, lambda, beta, proportions,
     objective) within 1e-9 relative,
  4. stores inputs + reference outputs as tests/golden/<case>.npz.
The fixtures are therefore outputs of the reference itself, not of the oracle.
"""
from __future__ import annotations

import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import numpy as np
from scipy import sparse

import flashdeconv                                       # the real reference
from flashdeconv import FlashDeconv as RefFlashDeconv
from flashdeconv.core.sketching import build_countsketch_matrix, sketch_data
from flashdeconv.core.solver import bcd_solve as ref_bcd_solve, normalize_proportions
from flashdeconv.core.spatial import auto_tune_lambda
from flashdeconv.utils.genes import select_informative_genes
from flashdeconv.utils.graph import build_knn_graph, coords_to_adjacency

from oracle import fd_oracle as fo
from flashdeconv_b200.synth import make_dataset

GOLD = os.path.join(ROOT, "tests", "golden")
VERSIONS = dict(reference=flashdeconv.__version__, numpy=np.__version__,
                scipy=__import__("scipy").__version__, numba=__import__("numba").__version__)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)) if a.size else 0.0


def pin_path_case(name, *, n_spots, n_genes, n_types, depth, d, k=6, method="knn", seed=0,
                  dense=False, jitter=0.1, max_iter=100, n_hvg=2000, n_markers=50, keep_rows=None,
                  preprocess="log_cpm"):
    ds = make_dataset(n_spots=n_spots, n_genes=n_genes, n_types=n_types, depth=depth,
                      jitter=jitter, seed=seed)
    Y = ds.Y.toarray().astype(np.float64) if dense else ds.Y.astype(np.float64)   # float64 counts: the reference's sparse log-CPM keeps the input dtype (deconv.py:183-188)
    X, coords = ds.X, ds.coords

    # ---- reference, stage by stage (mirrors FlashDeconv.fit) ----
    gene_idx, lev = select_informative_genes(Y, X, n_hvg=n_hvg, n_markers_per_type=n_markers)
    model = RefFlashDeconv(sketch_dim=d, k_neighbors=k, spatial_method=method, max_iter=max_iter,
                           n_hvg=n_hvg, n_markers_per_type=n_markers, random_state=seed, preprocess=preprocess)
    Ysub = Y[:, gene_idx]
    if sparse.issparse(Ysub):
        Ysub = Ysub.tocsr()
    Yt, Xt = model._preprocess_data(Ysub, X[:, gene_idx], preprocess)
    Ys, Xs, Omega = sketch_data(Yt, Xt, sketch_dim=d, leverage_scores=lev, random_state=seed)
    Omega = Omega.tocsr()
    A = coords_to_adjacency(coords, method=method, k=k)
    A.sort_indices()
    lam = auto_tune_lambda(Ys, Xs, A)
    beta, info = ref_bcd_solve(Ys, Xs, A, lambda_=lam, rho=0.01, max_iter=max_iter, tol=1e-4)
    prop = normalize_proportions(beta)
    # end-to-end through the public class must agree with the staged run
    full = RefFlashDeconv(sketch_dim=d, k_neighbors=k, spatial_method=method, max_iter=max_iter,
                          n_hvg=n_hvg, n_markers_per_type=n_markers, random_state=seed, preprocess=preprocess)
    assert np.array_equal(full.fit_transform(Y, X, coords), prop)

    # ---- oracle on the same inputs ----
    o_idx, o_lev = fo.select_genes(Y, X, n_hvg, n_markers)
    assert np.array_equal(o_idx, gene_idx), name
    assert rel(o_lev, lev) < 1e-9, (name, rel(o_lev, lev))
    res = fo.run_path(Y, X, coords, gene_idx, lev, d=d, method=method, k=k, max_iter=max_iter, seed=seed,
                      preprocess_method=preprocess)
    ref_bucket = Omega.indices.copy()
    ref_w = Omega.data.copy()
    assert Omega.nnz == len(gene_idx) and np.all(np.diff(Omega.indptr) == 1)
    assert np.array_equal(res["bucket"], ref_bucket), name
    assert np.array_equal(res["sign"], np.sign(ref_w).astype(np.int64)), name
    assert rel(res["weight"], ref_w) < 1e-12
    assert rel(res["Y_s"], Ys) < 1e-12 and rel(res["X_s"], Xs) < 1e-12
    Ao = res["A"].tocsr()
    Ao.sort_indices()
    assert np.array_equal(Ao.indptr, A.indptr) and np.array_equal(Ao.indices, A.indices), name
    assert abs(res["lam"] - lam) <= 1e-12 * abs(lam)
    assert res["info"]["n_iterations"] == info["n_iterations"]
    assert res["info"]["converged"] == info["converged"]
    assert rel(res["beta"], beta) < 1e-9, (name, rel(res["beta"], beta))
    assert rel(res["proportions"], prop) < 1e-9
    assert abs(res["info"]["final_objective"] - info["final_objective"]) <= 1e-9 * abs(info["final_objective"])
    if sparse.issparse(Y) and preprocess == "log_cpm":      # the fused full-CSR formulation the CUDA kernel uses
        Yf = fo.sketch_full_csr(Y, gene_idx, res["bucket"], res["weight"], d)
        assert rel(Yf, Ys) < 1e-12, (name, rel(Yf, Ys))
    # tie-freeness of the k/(k+1) boundary (precondition for bit-exact kNN sets)
    if method == "knn":
        from scipy.spatial import cKDTree
        dist, _ = cKDTree(coords).query(coords, k=min(k, n_spots - 1) + 2)
        assert np.all(np.diff(dist, axis=1)[:, 1:] > 0), "kNN tie in generated coords"

    rows = np.arange(n_spots) if keep_rows is None else np.arange(0, n_spots, keep_rows)
    Yc = sparse.csr_matrix(Y)
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        Y_indptr=Yc.indptr.astype(np.int64), Y_indices=Yc.indices.astype(np.int32),
        Y_data=Yc.data.astype(np.float32), Y_shape=np.array(Yc.shape), dense_input=np.array(dense),
        X=X, coords=coords, gene_idx=gene_idx, leverage=lev, bucket=ref_bucket, weight=ref_w,
        Ys_rows=rows, Ys=Ys[rows], Xs=Xs, A_indptr=A.indptr, A_indices=A.indices, lam=np.array(lam),
        beta=beta, proportions=prop, n_iterations=np.array(info["n_iterations"]),
        converged=np.array(info["converged"]), final_objective=np.array(info["final_objective"]),
        final_change=np.array(info["final_change"]),
        params=np.array([d, k, seed, max_iter, n_hvg, n_markers]), method=np.array(method),
        preprocess=np.array(preprocess),
        versions=np.array(repr(VERSIONS)))
    print(f"[pin] {name}: N={n_spots} G={n_genes}->{len(gene_idx)} K={n_types} d={d} nnz(A)={A.nnz} "
          f"iters={info['n_iterations']} conv={info['converged']} lam={lam:.5g} OK")


def pin_solver_fixture():
    """The reference's own solver fixtures (tests/test_solver.py:67-90 and :298-313)."""
    out = {}
    for tag, (n, K, d, noise, kw) in {
        "simple": (50, 5, 32, 0.1, dict(lambda_=0.1, rho=0.01, max_iter=50, tol=1e-4)),
        "determinism": (60, 7, 48, 0.05, dict(lambda_=0.1, rho=0.01, max_iter=30, tol=1e-6)),
    }.items():
        np.random.seed(42)
        Xs = np.random.randn(K, d)
        bt = np.random.rand(n, K)
        bt = bt / bt.sum(axis=1, keepdims=True)
        Ys = bt @ Xs + noise * np.random.randn(n, d)
        coords = np.random.rand(n, 2)
        A = build_knn_graph(coords, k=4)
        A.sort_indices()
        beta, info = ref_bcd_solve(Ys, Xs, A, **kw)
        ob, oi = fo.bcd_solve(Ys, Xs, A, kw["lambda_"], kw["rho"], kw["max_iter"], kw["tol"])
        assert rel(ob, beta) < 1e-9 and oi["n_iterations"] == info["n_iterations"], tag
        assert oi["converged"] == info["converged"]
        assert abs(oi["final_objective"] - info["final_objective"]) < 1e-9 * abs(info["final_objective"])
        Ao = fo.knn_adjacency(coords, 4)
        Ao.sort_indices()
        assert np.array_equal(Ao.indices, A.indices) and np.array_equal(Ao.indptr, A.indptr)
        for key, val in dict(Ys=Ys, Xs=Xs, coords=coords, A_indptr=A.indptr, A_indices=A.indices,
                             beta=beta, n_iterations=np.array(info["n_iterations"]),
                             converged=np.array(info["converged"]),
                             final_objective=np.array(info["final_objective"]),
                             final_change=np.array(info["final_change"]),
                             kw=np.array([kw["lambda_"], kw["rho"], kw["max_iter"], kw["tol"]])).items():
            out[f"{tag}_{key}"] = val
        print(f"[pin] solver fixture {tag}: iters={info['n_iterations']} conv={info['converged']} OK")
    out["versions"] = np.array(repr(VERSIONS))
    np.savez_compressed(os.path.join(GOLD, "solver_fixtures.npz"), **out)


def pin_known_answers():
    """Known answers the reference's tests hold (test_solver.py:22-35,153-185; test_spatial.py:57-74)."""
    from flashdeconv.core.solver import soft_threshold
    assert soft_threshold(5.0, 2.0) == 3.0 and soft_threshold(-5.0, 2.0) == -3.0 and soft_threshold(1.0, 2.0) == 0.0
    b = np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0], [2.0, 2.0, 0.0]])
    assert np.array_equal(fo.normalize(b), normalize_proportions(b))
    g = np.array([[i, j] for i in range(3) for j in range(3)], dtype=float)
    assert fo.radius_adjacency(g, 1.5)[4].nnz == 8 and fo.radius_adjacency(g, 1.1)[4].nnz == 4
    for seed in (0, 1, 7):
        for G, d in ((100, 16), (3305, 512)):
            lev = np.random.RandomState(seed).rand(G)
            Om = build_countsketch_matrix(G, d, lev, seed).tocsr()
            bk, sg, w = fo.countsketch_table(G, d, lev, seed)
            assert np.array_equal(Om.indices, bk) and rel(w, Om.data) < 1e-12
    print("[pin] known answers OK")


PATH_CASES = {
    "path_sparse_small": dict(n_spots=600, n_genes=900, n_types=6, depth=300.0, d=64, seed=0),
    "path_dense_small": dict(n_spots=400, n_genes=500, n_types=5, depth=4000.0, d=128, seed=1, dense=True),
    "path_sparse_k30": dict(n_spots=2500, n_genes=3000, n_types=30, depth=400.0, d=512, seed=2, keep_rows=25),
    "path_grid": dict(n_spots=900, n_genes=700, n_types=8, depth=500.0, d=128, seed=3, method="grid", jitter=0.0),
    # the linear preprocess branches (core/deconv.py:199-229), "next" row f2
    "path_raw": dict(n_spots=500, n_genes=800, n_types=7, depth=300.0, d=64, seed=4, preprocess="raw"),
    "path_pearson": dict(n_spots=700, n_genes=900, n_types=10, depth=350.0, d=128, seed=5, preprocess="pearson"),
    # more cell types than the register-resident kernels hold (warp-per-spot path, csrc/wide.cu); a fixed number of
    # sweeps so that float32 and float64 cannot stop one sweep apart
    "path_k72": dict(n_spots=1200, n_genes=900, n_types=72, depth=3000.0, d=256, seed=6, max_iter=30, keep_rows=10),
}


if __name__ == "__main__":
    # python oracle/pin_against_reference.py [case ...]   (no names: everything)
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:]
    if not only:
        pin_known_answers()
        pin_solver_fixture()
    for case_name, kw in PATH_CASES.items():
        if not only or case_name in only:
            pin_path_case(case_name, **kw)
    print("all pins OK; versions:", VERSIONS)

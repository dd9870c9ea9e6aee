"""CPU suite: the oracle against the golden fixtures (outputs of the real reference), the
host-side logic of the package, and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
from scipy import sparse

from conftest import ROOT, Golden, PATH_CASES
from oracle import fd_oracle as fo


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / (np.max(np.abs(b)) + 1e-300))


# ------------------------------------------------------------------ oracle vs reference goldens
def test_oracle_gene_selection_matches_reference(golden):
    idx, lev = fo.select_genes(golden.Y_input(), golden.X, golden.n_hvg, golden.n_markers)
    assert np.array_equal(idx, golden.gene_idx)
    assert rel(lev, golden.leverage) < 1e-9


def test_oracle_linear_preprocess_matches_reference(golden_linear):
    """preprocess='raw' / 'pearson' (core/deconv.py:199-229): whole path against the reference's outputs; these
    cases CONVERGE (15 / 27 sweeps), so the stop test (core/solver.py:395-413) is pinned too"""
    g = golden_linear
    res = fo.run_path(g.Y_input(), g.X, g.coords, g.gene_idx, g.leverage, d=g.d, method=g.method, k=g.k,
                      max_iter=g.max_iter, seed=g.seed, preprocess_method=g.preprocess)
    assert g.converged and res["info"]["converged"] and res["info"]["n_iterations"] == g.n_iterations < g.max_iter
    assert rel(res["Y_s"][g.Ys_rows], g.Ys) < 1e-12 and rel(res["X_s"], g.Xs) < 1e-12
    assert rel(res["proportions"], g.proportions) < 1e-9 and rel(res["beta"], g.beta) < 1e-9
    assert abs(res["lam"] - g.lam) <= 1e-12 * g.lam


def test_oracle_countsketch_bit_exact(golden):
    bucket, sign, weight = fo.countsketch_table(len(golden.gene_idx), golden.d, golden.leverage, golden.seed)
    assert np.array_equal(bucket, golden.bucket)
    assert np.array_equal(sign, np.sign(golden.weight).astype(np.int64))
    assert rel(weight, golden.weight) < 1e-12


def test_oracle_path_matches_reference(golden):
    res = fo.run_path(golden.Y_input(), golden.X, golden.coords, golden.gene_idx, golden.leverage, d=golden.d,
                      method=golden.method, k=golden.k, max_iter=golden.max_iter, seed=golden.seed)
    assert rel(res["Y_s"][golden.Ys_rows], golden.Ys) < 1e-12
    assert rel(res["X_s"], golden.Xs) < 1e-12
    A = res["A"].tocsr()
    A.sort_indices()
    assert np.array_equal(A.indptr, golden.A.indptr) and np.array_equal(A.indices, golden.A.indices)
    assert abs(res["lam"] - golden.lam) <= 1e-12 * golden.lam
    assert res["info"]["n_iterations"] == golden.n_iterations
    assert res["info"]["converged"] == golden.converged
    assert rel(res["beta"], golden.beta) < 1e-9
    assert rel(res["proportions"], golden.proportions) < 1e-9
    assert abs(res["info"]["final_objective"] - golden.final_objective) <= 1e-9 * abs(golden.final_objective)
    assert abs(res["info"]["final_change"] - golden.final_change) <= 1e-6 * abs(golden.final_change)


def test_oracle_fused_csr_sketch_equals_staged():
    g = Golden("path_sparse_k30")
    Yf = fo.sketch_full_csr(g.Y, g.gene_idx, g.bucket, g.weight, g.d)
    assert rel(Yf[g.Ys_rows], g.Ys) < 1e-12


def test_oracle_solver_fixtures():
    z = np.load(os.path.join(ROOT, "tests", "golden", "solver_fixtures.npz"))
    for tag in ("simple", "determinism"):
        n = len(z[f"{tag}_A_indptr"]) - 1
        A = sparse.csr_matrix((np.ones(len(z[f"{tag}_A_indices"])), z[f"{tag}_A_indices"], z[f"{tag}_A_indptr"]),
                              shape=(n, n))
        lam, rho, max_iter, tol = z[f"{tag}_kw"]
        beta, info = fo.bcd_solve(z[f"{tag}_Ys"], z[f"{tag}_Xs"], A, lam, rho, int(max_iter), tol)
        assert rel(beta, z[f"{tag}_beta"]) < 1e-9
        assert info["n_iterations"] == int(z[f"{tag}_n_iterations"])
        assert info["converged"] == bool(z[f"{tag}_converged"])
        Ak = fo.knn_adjacency(z[f"{tag}_coords"], 4)
        Ak.sort_indices()
        assert np.array_equal(Ak.indices, z[f"{tag}_A_indices"])


def test_oracle_known_answers():
    b = np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0], [2.0, 2.0, 0.0]])
    p = fo.normalize(b)
    assert np.allclose(p[0], [1 / 6, 2 / 6, 3 / 6]) and np.allclose(p[1], 1 / 3) and np.allclose(p[2], [.5, .5, 0])
    grid = np.array([[i, j] for i in range(3) for j in range(3)], dtype=float)
    assert fo.radius_adjacency(grid, 1.5)[4].nnz == 8 and fo.radius_adjacency(grid, 1.1)[4].nnz == 4
    rng = np.random.default_rng(0)
    c = rng.random((200, 2))
    A = fo.knn_adjacency(c, 5)
    assert (A != A.T).nnz == 0 and A.diagonal().sum() == 0
    nn = fo.knn_directed_bruteforce(c, 5)
    D = sparse.csr_matrix((np.ones(nn.size), (np.repeat(np.arange(200), 5), nn.ravel())), shape=(200, 200))
    S = ((D + D.T) > 0).astype(float)
    assert (S != A).nnz == 0


def test_oracle_objective_zero_at_perfect_fit():
    rng = np.random.default_rng(1)
    Xs = rng.standard_normal((4, 16))
    beta = rng.random((30, 4))
    Ys = beta @ Xs
    A = fo.knn_adjacency(rng.random((30, 2)), 3)
    H = (Xs @ Ys.T).T
    assert abs(fo.objective(beta, H, Xs @ Xs.T, float(np.sum(Ys ** 2)), A, 0.0, 0.0)) < 1e-8


# ------------------------------------------------------------------ host logic of the package
def test_package_gene_selection_matches_reference(golden):
    from flashdeconv_b200 import genes
    idx, lev = genes.select_informative_genes(golden.Y_input(), golden.X, golden.n_hvg, golden.n_markers)
    assert np.array_equal(idx, golden.gene_idx)
    assert rel(lev, golden.leverage) < 1e-9


def test_package_tables_bit_exact(golden):
    from flashdeconv_b200.pipeline import build_tables
    t = build_tables(golden.X, golden.gene_idx, golden.leverage, golden.d, golden.seed, golden.Y.shape[1])
    assert np.array_equal(t.bucket, golden.bucket)                       # buckets bit-exact
    assert np.array_equal(t.sign, np.sign(golden.weight).astype(np.int64))   # signs bit-exact
    assert rel(t.weight, golden.weight) < 1e-12
    assert rel(t.X_sketch, golden.Xs) < 1e-12
    sel = t.gene_bucket >= 0
    assert np.array_equal(np.flatnonzero(sel), golden.gene_idx)
    assert np.array_equal(t.gene_bucket[sel], golden.bucket.astype(np.int32))


def test_countsketch_matrix_properties():
    """reference tests/test_sketching.py:16-51: shape, one entry per row, same seed -> identical."""
    from flashdeconv_b200.sketching import build_countsketch_matrix
    Om = build_countsketch_matrix(100, 16, random_state=42)
    assert Om.shape == (100, 16)
    assert np.all(np.diff(Om.tocsr().indptr) == 1)
    Om2 = build_countsketch_matrix(100, 16, random_state=42)
    assert np.array_equal(Om.toarray(), Om2.toarray())
    lev = np.random.RandomState(0).rand(100)
    assert build_countsketch_matrix(100, 16, leverage_scores=lev / lev.sum(), random_state=42).shape == (100, 16)


def test_estimator_validation_messages():
    """reference core/deconv.py:105-124 and tests/test_integration.py:254-257,387-422."""
    from flashdeconv_b200 import FlashDeconv
    with pytest.raises(ValueError, match="radius must be specified"):
        FlashDeconv(spatial_method="radius")
    with pytest.raises(ValueError, match="sketch_dim must be positive"):
        FlashDeconv(sketch_dim=0)
    with pytest.raises(ValueError, match="tol must be positive"):
        FlashDeconv(tol=0)
    m = FlashDeconv()
    with pytest.raises(RuntimeError, match="not been fitted"):
        m.get_cell_type_proportions()
    assert m.summary() == {"fitted": False}
    Y, X, c = np.ones((5, 7)), np.ones((2, 6)), np.zeros((5, 2))
    with pytest.raises(ValueError, match="Gene dimension mismatch"):
        m.fit(Y, X, c)
    with pytest.raises(ValueError, match="Spot count mismatch"):
        m.fit(Y, np.ones((2, 7)), np.zeros((4, 2)))
    with pytest.raises(ValueError, match="at least one cell type"):
        m.fit(Y, np.ones((0, 7)), c)
    assert "not fitted" in repr(m)
    # cell-type limits: 64 on the register-resident kernels, 1024 on the warp-per-spot kernels (checked before any upload)
    with pytest.raises(ValueError, match="at most 1024 cell types"):
        m.fit(np.ones((5, 7)), np.ones((1025, 7)), c)


def test_multi_gpu_path_rejects_more_than_64_types():
    """the tiled path keeps the register-resident kernels; the limit is raised before any device work"""
    from flashdeconv_b200 import tiling
    with pytest.raises(ValueError, match="at most 64 cell types"):
        tiling.TiledPath(None, None, None, 65)
    with pytest.raises(ValueError, match="download must be 'all' or 'rank0'"):
        tiling.deconvolve_path_tiled(None, None, None, None, None, download="rank1")


class _Ad:
    """the slice of AnnData the io helpers touch"""
    def __init__(self, X, var, obs=None, obsm=None):
        import pandas as pd
        self.X, self.layers, self.var_names = X, {}, np.asarray(var)
        self.obs, self.obsm, self.uns = pd.DataFrame(obs or {}), dict(obsm or {}), {}
        self.n_obs = X.shape[0]
        self.obs_names = np.asarray([f"s{i}" for i in range(X.shape[0])])


def test_io_helpers_follow_the_reference_loader():
    """flashdeconv/io/loader.py:15-318: coordinate lookup order, per-type means, first-occurrence gene alignment in sorted
    name order, DataFrame + categorical hand-off, error messages (dense inputs: host bookkeeping only)"""
    from flashdeconv_b200 import io as fio
    rng = np.random.default_rng(0)
    st = _Ad(rng.poisson(2.0, (6, 5)).astype(np.float64), ["g3", "g1", "g1", "g7", "g0"], obs={"x": np.arange(6.0), "y": np.ones(6)})
    Y, coords, genes = fio.load_spatial_data(st)
    assert Y is st.X and np.array_equal(coords, np.column_stack([np.arange(6.0), np.ones(6)])) and list(genes) == list(st.var_names)
    st.obsm["X_spatial"] = np.zeros((6, 2))
    assert np.array_equal(fio.load_spatial_data(st)[1], np.zeros((6, 2)))
    st.obsm["spatial"] = np.full((6, 2), 3.0)
    assert np.array_equal(fio.load_spatial_data(st)[1], np.full((6, 2), 3.0))
    with pytest.raises(ValueError, match="Could not find spatial coordinates"):
        fio.load_spatial_data(_Ad(st.X, st.var_names))
    cells = rng.poisson(3.0, (9, 4)).astype(np.float64)
    labels = np.array(["b", "a", "b", "c", "a", "b", "c", "c", "b"])
    rf = _Ad(cells, ["g1", "g9", "g0", "g3"], obs={"cell_type": labels})
    X, names, rgenes = fio.load_reference(rf)
    assert list(names) == ["a", "b", "c"] and np.allclose(X[1], cells[labels == "b"].mean(0)) and X.dtype == np.float64
    assert np.allclose(fio.load_reference(rf, method="sum")[0][2], cells[labels == "c"].sum(0))
    with pytest.raises(ValueError, match="Unknown aggregation method"):
        fio.load_reference(rf, method="median")
    with pytest.raises(ValueError, match="Cell type key 'nope' not found"):
        fio.load_reference(rf, cell_type_key="nope")
    Ya, Xa, common = fio.align_genes(st.X, X, st.var_names, rgenes)
    assert list(common) == ["g0", "g1", "g3"]                                   # sorted names
    assert np.array_equal(Ya, st.X[:, [4, 1, 0]]) and np.array_equal(Xa, X[:, [2, 0, 3]])    # duplicate g1: first occurrence
    with pytest.raises(ValueError, match="No common genes"):
        fio.align_genes(st.X, X, np.array(list("abcde")), np.array(list("wxyz")))
    Yp, Xp, cp, np_names, gp = fio.prepare_data(st, rf)
    assert np.array_equal(Yp, Ya) and np.array_equal(Xp, Xa) and list(gp) == list(common) and list(np_names) == ["a", "b", "c"]
    beta = rng.random((6, 3))
    out = fio.result_to_anndata(beta, st, names)
    assert out is st and list(st.obsm["flashdeconv"].columns) == ["a", "b", "c"]
    assert np.array_equal(st.obsm["flashdeconv"].to_numpy(), beta)
    assert list(st.obs["flashdeconv_dominant"]) == list(names[np.argmax(beta, axis=1)])
    assert list(fio.result_to_anndata(beta, _Ad(st.X, st.var_names)).obsm["flashdeconv"].columns) == ["CellType_0", "CellType_1", "CellType_2"]
    with pytest.raises(ValueError, match="beta must be 2D"):
        fio.result_to_anndata(beta[0], st)
    with pytest.raises(ValueError, match="beta rows must match"):
        fio.result_to_anndata(beta[:4], st)
    with pytest.raises(ValueError, match="must match beta.shape"):
        fio.result_to_anndata(beta, st, ["a", "b"])


def test_reference_import_paths_resolve():
    """flashdeconv/{core,utils,io,tl}/__init__.py: the names a user of the reference imports resolve under the same
    sub-package paths (utils.metrics is evaluation code outside the accelerated path and is not provided)"""
    import importlib
    import flashdeconv_b200 as fd
    want = {"core": ["FlashDeconv", "build_countsketch_matrix", "project_to_sketch", "compute_laplacian", "get_neighbor_indices",
                     "bcd_solve"],
            "utils": ["select_hvg", "select_markers", "compute_leverage_scores", "build_knn_graph", "build_radius_graph",
                      "coords_to_adjacency", "check_random_state"],
            "io": ["load_spatial_data", "load_reference", "align_genes", "result_to_anndata", "prepare_data"],
            "tl": ["deconvolve"]}
    for sub, names in want.items():
        mod = importlib.import_module(f"flashdeconv_b200.{sub}")
        for name in names:
            assert callable(getattr(mod, name)), (sub, name)
    for path, name in (("core.deconv", "FlashDeconv"), ("core.sketching", "sketch_data"), ("core.solver", "normalize_proportions"),
                       ("core.solver", "compute_objective"), ("core.spatial", "auto_tune_lambda"),
                       ("utils.genes", "select_informative_genes"), ("utils.graph", "build_grid_graph"),
                       ("utils.random", "check_random_state"), ("io.loader", "prepare_data")):
        assert callable(getattr(importlib.import_module(f"flashdeconv_b200.{path}"), name)), (path, name)
    assert fd.core.FlashDeconv is fd.FlashDeconv
    from flashdeconv_b200.utils import check_random_state
    rs = np.random.RandomState(4)
    assert check_random_state(rs) is rs and check_random_state(None) is np.random.mtrand._rand
    assert check_random_state(np.int64(7)).randint(0, 100) == np.random.RandomState(7).randint(0, 100)
    with pytest.raises(ValueError, match="cannot be used to seed"):
        check_random_state("seed")


def test_synth_generator_is_deterministic():
    from flashdeconv_b200.synth import make_dataset
    a = make_dataset(300, 200, 4, depth=100.0, seed=5)
    b = make_dataset(300, 200, 4, depth=100.0, seed=5)
    assert np.array_equal(a.Y.indices, b.Y.indices) and np.array_equal(a.Y.data, b.Y.data)
    assert np.array_equal(a.coords, b.coords) and a.Y.shape == (300, 200)
    assert np.allclose(a.beta_true.sum(1), 1.0)


# ------------------------------------------------------------------ C-ABI surface
def test_cabi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fdb200.h")).read()
    declared = set(re.findall(r"\b(fdb_[a-z0-9_]+)\s*\(", header))
    declared -= {"fdb_status"}
    assert len(declared) >= 15
    from flashdeconv_b200 import _native
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    so = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(so, name), name
    assert _native.lib.fdb_abi_version() == 1
    assert [_native.padded_types(k) for k in (1, 4, 5, 30, 33, 50, 64)] == [8, 8, 8, 32, 40, 56, 64]


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "flashdeconv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "fd_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f

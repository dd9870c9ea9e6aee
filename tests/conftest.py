import os
import sys

import numpy as np
import pytest
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with `-m gpu` on a B200)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One tests/golden/<name>.npz fixture: inputs + outputs of the REAL reference (see
    oracle/pin_against_reference.py)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.z = z
        shape = tuple(int(v) for v in z["Y_shape"])
        self.Y = sparse.csr_matrix((z["Y_data"].astype(np.float64), z["Y_indices"], z["Y_indptr"]), shape=shape)
        self.dense_input = bool(z["dense_input"])
        self.X, self.coords = z["X"], z["coords"]
        self.gene_idx, self.leverage = z["gene_idx"], z["leverage"]
        self.bucket, self.weight = z["bucket"], z["weight"]
        self.Ys_rows, self.Ys, self.Xs = z["Ys_rows"], z["Ys"], z["Xs"]
        self.A = sparse.csr_matrix((np.ones(len(z["A_indices"])), z["A_indices"], z["A_indptr"]),
                                   shape=(shape[0], shape[0]))
        self.lam = float(z["lam"])
        self.beta, self.proportions = z["beta"], z["proportions"]
        self.n_iterations, self.converged = int(z["n_iterations"]), bool(z["converged"])
        self.final_objective, self.final_change = float(z["final_objective"]), float(z["final_change"])
        self.d, self.k, self.seed, self.max_iter, self.n_hvg, self.n_markers = (int(v) for v in z["params"])
        self.method = str(z["method"])
        self.preprocess = str(z["preprocess"]) if "preprocess" in z.files else "log_cpm"

    def Y_input(self):
        return self.Y.toarray() if self.dense_input else self.Y


PATH_CASES = ["path_sparse_small", "path_dense_small", "path_sparse_k30", "path_grid", "path_k72"]   # k72: more types than the register-resident kernels hold


LINEAR_CASES = ["path_raw", "path_pearson"]          # preprocess="raw" / "pearson" (core/deconv.py:199-229)


@pytest.fixture(params=PATH_CASES)
def golden(request):
    return Golden(request.param)


@pytest.fixture(params=LINEAR_CASES)
def golden_linear(request):
    return Golden(request.param)


def pearson_per_type(a, b):
    out = []
    for k in range(a.shape[1]):
        x, y = a[:, k], b[:, k]
        if x.std() == 0 and y.std() == 0:
            out.append(1.0)
        else:
            out.append(float(np.corrcoef(x, y)[0, 1]))
    return np.array(out)

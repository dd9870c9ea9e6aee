"""Function-level mirror of the reference's ``flashdeconv/core/solver.py`` over libfdb200.

``bcd_solve`` (:287-428) takes the same arguments (float64 numpy sketches, scipy
adjacency) and returns ``(beta, info)`` with the same keys; H, the sweeps, the
convergence test and the objective all run on the GPU in float32 storage /
float64 objective accumulation.  Spots keep the caller's order here (the
adjacency is whatever the caller built), unlike the production path which
tile-orders them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy import sparse


def soft_threshold(x: float, threshold: float) -> float:
    return x - threshold if x > threshold else (x + threshold if x < -threshold else 0.0)


def precompute_gram_matrix(X_sketch: np.ndarray) -> np.ndarray:
    return X_sketch @ X_sketch.T


def normalize_proportions(beta: np.ndarray) -> np.ndarray:
    """Row-normalise on the device; all-zero rows become uniform (core/solver.py:431-452)."""
    from ._native import check, lib, padded_types, require_cuda
    from .pipeline import _ptr, _stream
    torch = require_cuda()
    beta = np.asarray(beta, dtype=np.float64)
    n, k = beta.shape
    if n == 0 or k == 0:
        return beta.copy()
    kp = padded_types(k)
    b = torch.zeros((n, kp), dtype=torch.float32, device="cuda")
    b[:, :k] = torch.from_numpy(beta).to("cuda", torch.float32)
    out = torch.empty((n, k), dtype=torch.float64, device="cuda")
    check(lib.fdb_finish(_ptr(b), C.c_void_p(0), n, k, C.c_void_p(0), _ptr(out), _stream(torch)), "finish")
    return out.cpu().numpy()


class _Problem:
    """Device state for one (Y_sketch, X_sketch, A) problem in the caller's spot order."""

    def __init__(self, Y_sketch, X_sketch, A):
        from ._native import check, lib, padded_types, require_cuda
        from .pipeline import _ptr, _stream
        self.torch = torch = require_cuda()
        self.check, self.lib, self._ptr, self._stream = check, lib, _ptr, _stream
        self.n, self.d = Y_sketch.shape
        self.K = X_sketch.shape[0]
        self.Kp = padded_types(self.K)
        dev = "cuda"
        ys = torch.from_numpy(np.ascontiguousarray(Y_sketch, dtype=np.float32)).to(dev)
        xs = torch.from_numpy(np.ascontiguousarray(X_sketch, dtype=np.float32)).to(dev)
        self.h = torch.zeros((self.n, self.Kp), dtype=torch.float32, device=dev)
        self.ysq = torch.zeros(self.n, dtype=torch.float32, device=dev)
        check(lib.fdb_contract(_ptr(ys), _ptr(xs), self.n, self.d, self.K, _ptr(self.h), _ptr(self.ysq),
                               _stream(torch)), "contract")
        Ac = sparse.csr_matrix(A)
        Ac.sort_indices()
        self.indptr = torch.from_numpy(Ac.indptr.astype(np.int32)).to(dev)
        self.indices = torch.from_numpy(Ac.indices.astype(np.int32)).to(dev)
        if self.indices.numel() == 0:
            self.indices = torch.zeros(1, dtype=torch.int32, device=dev)
        self.gram64 = np.asarray(X_sketch, dtype=np.float64) @ np.asarray(X_sketch, dtype=np.float64).T
        self.gram32 = np.ascontiguousarray(self.gram64, dtype=np.float32)
        self.gram_ptr = self.gram32.ctypes.data_as(C.c_void_p)
        self.a = torch.empty((self.n, self.Kp), dtype=torch.float32, device=dev)
        self.b = torch.empty((self.n, self.Kp), dtype=torch.float32, device=dev)
        self.state = torch.zeros(16, dtype=torch.int32, device=dev)
        from .pipeline import MAX_TYPES, MAX_TYPES_WIDE, build_sweep_plan
        if self.K > MAX_TYPES_WIDE:
            raise ValueError(f"at most {MAX_TYPES_WIDE} cell types are supported; got {self.K}")
        self.wide = self.K > MAX_TYPES             # warp-per-spot kernels (csrc/wide.cu): device Gram, no gather plan
        if self.wide:
            gp = np.zeros((self.Kp, self.Kp), dtype=np.float32)
            gp[: self.K, : self.K] = self.gram32
            self.gram_dev = torch.from_numpy(gp).to(dev)
            self.plan = None
        else:
            self.plan = build_sweep_plan(self.indptr, self.indices, self.n, Ac.nnz, self.K)

    def solve(self, lam, rho_scaled, max_iter, tol):
        st = self._stream(self.torch)
        a = (self._ptr(self.indptr), self._ptr(self.indices), self.n, self.K, float(lam), float(rho_scaled), int(max_iter),
             float(tol), self._ptr(self.state))
        if self.wide:
            self.check(self.lib.fdb_bcd_solve_wide(self._ptr(self.h), self._ptr(self.gram_dev), self._ptr(self.a),
                                                   self._ptr(self.b), *a, st), "bcd_solve_wide")
        else:
            self.check(self.lib.fdb_bcd_solve(self._ptr(self.h), self.gram_ptr, self._ptr(self.a), self._ptr(self.b), *a,
                                              self._ptr(self.plan), st), "bcd_solve")

    def sweep(self, cur, nxt, lam, rho_scaled, tol):
        st = self._stream(self.torch)
        a = (self._ptr(self.indptr), self._ptr(self.indices), self.n, self.K, float(lam), float(rho_scaled), float(tol), 1,
             self._ptr(self.state))
        if self.wide:
            self.check(self.lib.fdb_bcd_sweep_wide(self._ptr(self.h), self._ptr(self.gram_dev), self._ptr(cur), self._ptr(nxt),
                                                   *a, st), "bcd_sweep_wide")
        else:
            self.check(self.lib.fdb_bcd_sweep(self._ptr(self.h), self.gram_ptr, self._ptr(cur), self._ptr(nxt), *a,
                                              self._ptr(self.plan), st), "bcd_sweep")

    def read_state(self):
        st = self.state.cpu()
        return int(st[3]), bool(int(st[4])), float(st[5:6].view(self.torch.float32)[0])

    def objective(self, beta_dev, lam, rho_scaled):
        out = self.torch.zeros(5, dtype=self.torch.float64, device="cuda")
        fn, gram = (self.lib.fdb_objective_terms_wide, self._ptr(self.gram_dev)) if self.wide else \
            (self.lib.fdb_objective_terms, self.gram_ptr)
        self.check(fn(self._ptr(beta_dev), self._ptr(self.h), self._ptr(self.ysq), gram, self._ptr(self.indptr),
                      self._ptr(self.indices), self.n, self.K, self._ptr(out), self._stream(self.torch)), "objective_terms")
        cross, quad, lap, l1, yty = out.cpu().tolist()
        return 0.5 * (yty - 2.0 * cross + quad) + 0.5 * lam * lap + rho_scaled * l1


def bcd_solve(Y_sketch: np.ndarray, X_sketch: np.ndarray, A, lambda_: float = 0.1, rho: float = 0.01,
              max_iter: int = 100, tol: float = 1e-4, verbose: bool = False):
    n = Y_sketch.shape[0]
    K = X_sketch.shape[0]
    if n == 0 or K == 0:
        return np.empty((n, K), dtype=np.float64), dict(converged=True, n_iterations=0, final_objective=0.0,
                                                        objectives=[], final_change=0.0)
    P = _Problem(Y_sketch, X_sketch, A)
    rho_scaled = rho * float(np.mean(np.diag(P.gram64)))
    st = P._stream(P.torch)
    objectives = []
    if not verbose:
        P.solve(lambda_, rho_scaled, max_iter, tol)
        n_iter, conv, rel = P.read_state()
    else:
        P.check(P.lib.fdb_bcd_init(P._ptr(P.a), n, K, P._ptr(P.state), st), "bcd_init")
        cur, nxt = P.a, P.b
        n_iter, conv, rel = 0, False, 0.0
        for it in range(max_iter):
            P.sweep(cur, nxt, lambda_, rho_scaled, tol)
            n_iter, conv, rel = P.read_state()
            if it % 10 == 0 or it == max_iter - 1:
                obj = P.objective(nxt, lambda_, rho_scaled)
                objectives.append(obj)
                print(f"Iteration {it}: objective = {obj:.6f}, rel_change = {rel:.6e}")
            cur, nxt = nxt, cur
            if conv:
                print(f"Converged at iteration {it}")
                break
    if max_iter == 0:
        rel = 0.0
    final = P.a if n_iter % 2 == 0 else P.b
    info = dict(converged=conv, n_iterations=n_iter, final_objective=P.objective(final, lambda_, rho_scaled),
                objectives=objectives if verbose else [], final_change=rel)
    return final[:, :K].to(P.torch.float64).cpu().numpy(), info


def compute_objective(beta: np.ndarray, H: np.ndarray, XtX: np.ndarray, YtY: float, L, lambda_: float,
                      rho: float) -> float:
    """Objective from precomputed pieces with the reference's signature (H is K x N, L the Laplacian).

    The Laplacian term is evaluated on the device from the adjacency recovered as A = diag(L) - L."""
    from ._native import check, lib, padded_types, require_cuda
    from .pipeline import _ptr, _stream
    torch = require_cuda()
    n, K = beta.shape
    if n == 0:
        return 0.5 * float(YtY)
    kp = padded_types(K)
    Lc = sparse.csr_matrix(L)
    A = (sparse.diags(Lc.diagonal()) - Lc).tocsr()
    A.eliminate_zeros()
    A.sort_indices()
    if A.nnz and not np.allclose(A.data, 1.0):
        raise NotImplementedError("compute_objective on the device expects the Laplacian of a binary adjacency")
    pad = lambda M: torch.from_numpy(np.ascontiguousarray(
        np.pad(np.asarray(M, dtype=np.float32), ((0, 0), (0, kp - K))))).cuda()
    b, h = pad(beta), pad(np.asarray(H).T)
    ysq = torch.zeros(n, dtype=torch.float32, device="cuda")
    ip = torch.from_numpy(A.indptr.astype(np.int32)).cuda()
    ix = torch.from_numpy(np.concatenate([A.indices.astype(np.int32), np.zeros(1, np.int32)])).cuda()
    g32 = np.ascontiguousarray(XtX, dtype=np.float32)
    out = torch.zeros(5, dtype=torch.float64, device="cuda")
    check(lib.fdb_objective_terms(_ptr(b), _ptr(h), _ptr(ysq), g32.ctypes.data_as(C.c_void_p), _ptr(ip), _ptr(ix),
                                  n, K, _ptr(out), _stream(torch)), "objective_terms")
    cross, quad, lap, l1, _ = out.cpu().tolist()
    return 0.5 * (float(YtY) - 2.0 * cross + quad) + 0.5 * lambda_ * lap + rho * l1

"""N>1 host logic on CPU: world-size-2 gloo run of the tile partition + halo exchange, with the oracle's
float64 sweep standing in for the CUDA kernel.  The partitioned solve must reproduce the single-process one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flashdeconv_b200.tiling import halo_exchange, plan_tile, tile_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(n=900, K=6, d=24, k=6, seed=3):
    from oracle import fd_oracle as fo
    rng = np.random.default_rng(seed)
    coords = rng.random((n, 2)) * 30
    # a spatially coherent numbering (row-major over coarse cells), like the device tile order
    key = np.floor(coords[:, 1] / 3).astype(int) * 100 + np.floor(coords[:, 0] / 3).astype(int)
    order = np.argsort(key, kind="stable")
    coords = coords[order]
    A = fo.knn_adjacency(coords, k)
    A.sort_indices()
    Xs = rng.standard_normal((K, d)) + 0.4
    Ys = (rng.random((n, K)) * (rng.random((n, K)) < 0.4)) @ Xs + 0.05 * rng.standard_normal((n, d))
    return A, Xs, Ys


def _worker(rank, world, port, sweeps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import fd_oracle as fo
        A, Xs, Ys = _problem()
        n, K = Ys.shape[0], Xs.shape[0]
        gram = np.ascontiguousarray(Xs @ Xs.T)
        H = np.ascontiguousarray((Xs @ Ys.T).T)
        lam, rho_s = 0.8, 0.01 * float(np.mean(np.diag(gram)))
        bounds = tile_bounds(n, world, align=32)
        plan = plan_tile(torch.from_numpy(A.indptr.astype(np.int32)), torch.from_numpy(A.indices.astype(np.int32)),
                         bounds, rank)
        assert plan.n_own == bounds[rank][1] - bounds[rank][0] and plan.n_halo > 0
        ptr = plan.indptr.numpy().astype(np.int64)
        idx = plan.indices.numpy().astype(np.int64)
        cur = torch.full((plan.n_total, K), 1.0 / K, dtype=torch.float64)
        nxt = cur.clone()
        h_own = np.ascontiguousarray(H[plan.lo:plan.hi])
        diffs, absv = np.empty(plan.n_own), np.empty(plan.n_own)
        rels = []
        for _ in range(sweeps):
            b_prev, b_next = cur.numpy(), nxt.numpy()
            # sweep own rows only; the oracle reads halo rows through the local adjacency
            fo._native().fdo_bcd_sweep(fo._p(h_own, fo.ctypes.c_double), fo._p(gram, fo.ctypes.c_double),
                                       fo._p(b_prev, fo.ctypes.c_double), fo._p(b_next, fo.ctypes.c_double),
                                       fo._p(ptr, fo.ctypes.c_int64), fo._p(idx, fo.ctypes.c_int64), plan.n_own, K,
                                       lam, rho_s, fo._p(diffs, fo.ctypes.c_double), fo._p(absv, fo.ctypes.c_double))
            halo_exchange(nxt, plan, lambda b, rows: b.index_select(0, rows.to(torch.int64)).contiguous())
            stat = torch.tensor([diffs.max(), absv.max()], dtype=torch.float64)
            dist.all_reduce(stat, op=dist.ReduceOp.MAX)
            rels.append(float(stat[0] / (stat[1] + 1e-10)))
            cur, nxt = nxt, cur
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), beta=cur[:plan.n_own].numpy(), lo=plan.lo, hi=plan.hi,
                 rels=np.array(rels), halo=plan.halo_global.numpy(), n_send=sum(r.numel() for _, r in plan.send))
    finally:
        dist.destroy_process_group()


def test_tile_bounds_cover_and_align():
    for n, w in ((1000, 2), (256 * 7 + 3, 4), (100, 8), (0, 2), (5, 1)):
        b = tile_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert all(lo % 256 == 0 for lo, _ in b if lo < n)


def test_plan_tile_is_consistent_without_processes():
    A, _, _ = _problem(n=500)
    ip, ix = torch.from_numpy(A.indptr.astype(np.int32)), torch.from_numpy(A.indices.astype(np.int32))
    bounds = tile_bounds(500, 3, align=32)
    plans = [plan_tile(ip, ix, bounds, r) for r in range(3)]
    for p in plans:
        glob = np.concatenate([np.arange(p.lo, p.hi), p.halo_global.numpy()])
        for i in range(p.n_own):                                     # local adjacency maps back to the global one
            mine = np.sort(glob[p.indices.numpy()[p.indptr[i]:p.indptr[i + 1]]])
            assert np.array_equal(mine, A.indices[A.indptr[p.lo + i]:A.indptr[p.lo + i + 1]])
        for peer, first, count in p.recv:                            # what I receive is exactly what the peer sends me
            sent = [rows for q, rows in plans[peer].send if q == p.rank][0].numpy() + plans[peer].lo
            assert np.array_equal(sent, p.halo_global.numpy()[first:first + count])


@pytest.mark.timeout(300)
def test_two_rank_gloo_solve_matches_single_process(tmp_path):
    from oracle import fd_oracle as fo
    sweeps, world = 12, 2
    mp.spawn(_worker, args=(world, _free_port(), sweeps, str(tmp_path)), nprocs=world, join=True)
    A, Xs, Ys = _problem()
    trace = []
    fo.bcd_solve(Ys, Xs, A, 0.8, 0.01, sweeps, 1e-30, trace=trace)
    want = trace[-1]
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    got = np.concatenate([p["beta"] for p in parts])
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    assert np.allclose(parts[0]["rels"], parts[1]["rels"])          # both ranks see the same global stop statistic
    assert all(p["n_send"] > 0 for p in parts)


def test_device_plan_equals_reference_plan():
    """plan_tile_device (sync-light, what the fused multi-GPU sweep consumes) against plan_tile on every rank: same
    local adjacency and halo rows, and push lists that land each boundary row exactly where the peer's halo slice
    expects it (its own rows, then the rows of lower ranks, ascending position)."""
    from flashdeconv_b200 import tiling
    for n, world, seed in ((900, 2, 3), (5000, 4, 5), (2300, 8, 7), (300, 3, 1)):
        A, _, _ = _problem(n=n, seed=seed)
        ptr = torch.from_numpy(A.indptr.astype(np.int32))
        idx = torch.from_numpy(A.indices.astype(np.int32))
        bounds = tiling.tile_bounds(n, world, align=128)
        ref = [tiling.plan_tile(ptr, idx, bounds, r) for r in range(world)]
        dev = [tiling.plan_tile_device(ptr, idx, int(idx.numel()), bounds, r) for r in range(world)]
        for r in range(world):
            a, b = ref[r], dev[r]
            assert (a.n_own, a.n_halo) == (b.n_own, b.n_halo)
            assert torch.equal(a.indptr, b.indptr) and torch.equal(a.indices[: int(a.indptr[-1])], b.indices[: int(b.indptr[-1])])
            assert torch.equal(a.halo_global, b.halo_global)
            assert sorted(a.recv) == sorted(b.recv)
            assert b.cap_rows == max(x.n_total for x in ref)
            # push entries grouped by peer must reproduce the reference send lists and land in the peer's recv slice
            ent = b.push_ent[: int(b.push_ptr[-1])]
            rows = torch.repeat_interleave(torch.arange(b.n_own), (b.push_ptr[1:] - b.push_ptr[:-1]).long())
            for peer, send_rows in a.send:
                sel = ent[:, 0] == peer
                assert torch.equal(rows[sel].to(torch.int32), send_rows)
                first = [f for (q, f, c) in ref[peer].recv if q == r][0]
                want = ref[peer].n_own + first + torch.arange(int(sel.sum()))
                assert torch.equal(ent[sel, 1].long(), want)
            assert sum(len(sr) for _, sr in a.send) == int(b.push_ptr[-1])
            # patch order: a permutation with the boundary patches first
            nb = int(b.n_boundary)
            order = b.patch_order.long()
            assert sorted(order.tolist()) == list(range(order.numel()))
            cnt = b.push_ptr[1:] - b.push_ptr[:-1]
            has = [bool(cnt[q * 128:(q + 1) * 128].sum() > 0) for q in range(order.numel())]
            assert all(has[q] for q in order[:nb].tolist()) and not any(has[q] for q in order[nb:].tolist())

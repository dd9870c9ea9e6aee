"""Multi-GPU partition of the hot path: spatial tiles + per-sweep halo exchange.

The reference is single-process; this module is the B200-native scale-out of its Jacobi
sweep.  Because every sweep reads only the PREVIOUS iterate of a spot's neighbours (core/solver.py:161-166),
cutting the spots into R tiles changes nothing but the order of float additions:

  * tile order (graph.cu) is a spatially coherent numbering, so rank r owns the contiguous position range
    [lo_r, hi_r) -- a compact patch of the tissue -- and sketches / solves only those rows;
  * neighbours owned by another rank become HALO rows appended after the rank's own rows in its beta buffers
    (sorted by global position, hence grouped by owner: a peer's rows land in one contiguous slice, no unpack);
  * per sweep: sweep own rows -> pack the boundary rows each peer needs (fdb_rows_gather) -> batched
    isend/irecv over NCCL/NVLink straight into the halo slices -> MAX all-reduce of the two max-norm words
    (the reference's stop test uses global max norms, core/solver.py:395-397) -> fdb_bcd_finalize;
  * X_s, the Gram matrix, lambda, rho and the (small) graph are replicated: every rank builds the full graph
    from the replicated coordinates (0.7 ms at 1M spots) instead of exchanging k-NN lists.

`plan_tile` / `halo_exchange` are device-agnostic torch code so that the partition logic is tested on CPU with
the gloo backend (tests/test_tiling_gloo.py); `TiledPath` is the GPU driver.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import numpy as np


_COMM_CACHE = {}
_SYMM_CACHE = {}        # (floats, group id) -> (symmetric buffer, rendezvous handle, peer base pointers)
_SEQ_BASE = 1           # sweep sequence numbers of the peer-memory hand-shake; identical on every rank


def release_communicators():
    """Destroy the cached native NCCL communicators (call before torch.distributed.destroy_process_group)."""
    import torch
    from . import _native
    if _COMM_CACHE and torch.cuda.is_available():
        torch.cuda.synchronize()
    for comm in _COMM_CACHE.values():
        _native.lib.fdb_comm_destroy(comm)
    _COMM_CACHE.clear()
    _SYMM_CACHE.clear()


def tile_bounds(n: int, world: int, align: int = 256) -> List[Tuple[int, int]]:
    """Contiguous position ranges, one per rank, cut at multiples of `align` (the sweep kernel's CTA tile)."""
    if world <= 0:
        raise ValueError("world must be positive")
    blocks = -(-n // align)
    cuts = [min(n, ((blocks * r) // world) * align) for r in range(world)] + [n]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


@dataclass
class TilePlan:
    rank: int
    lo: int
    hi: int
    n_own: int
    n_halo: int
    indptr: "object"               # int32[n_own + 1], local
    indices: "object"              # int32[nnz_local], local numbering: own rows 0..n_own-1, halo rows after
    halo_global: "object"          # int64[n_halo] global positions of the halo rows (ascending)
    recv: List[Tuple[int, int, int]]        # (peer, first halo slot, count)
    send: List[Tuple[int, "object"]]        # (peer, int32 local row ids to send, ascending global position)

    @property
    def n_total(self) -> int:
        return self.n_own + self.n_halo


def plan_tile(indptr, indices, bounds: List[Tuple[int, int]], rank: int) -> TilePlan:
    """Local adjacency + halo maps for `rank` from the replicated global CSR (torch tensors, any device).

    Only the rank's own rows are scanned: the graph is undirected, so the rows a peer needs from this rank are
    exactly this rank's rows that have a neighbour owned by that peer."""
    import torch
    lo, hi = bounds[rank]
    dev = indices.device
    n_own = hi - lo
    edge_range = indptr[lo:hi + 1].to(torch.int64)
    e0, e1 = (int(v) for v in edge_range[[0, -1]].tolist()) if n_own else (0, 0)
    nbr = indices[e0:e1].to(torch.int64)
    local_ptr = (edge_range - e0)
    outside = (nbr < lo) | (nbr >= hi)
    cross = torch.nonzero(outside).flatten()                      # my edges that leave the tile
    cross_nbr = nbr[cross]
    halo_global = torch.unique(cross_nbr)                          # sorted ascending
    n_halo = int(halo_global.numel())
    local = nbr - lo
    if n_halo:
        local[cross] = n_own + torch.searchsorted(halo_global, cross_nbr)
    recv, send = [], []
    if n_halo:
        starts = torch.tensor([b[0] for b in bounds] + [bounds[-1][1]], device=dev, dtype=torch.int64)
        owner_edges = torch.searchsorted(halo_global, starts).tolist()       # halo rows are grouped by owner
        for peer in range(len(bounds)):
            a, b = owner_edges[peer], owner_edges[peer + 1]
            if peer != rank and b > a:
                recv.append((peer, a, b - a))
        # (owner of the outside neighbour, my row) pairs -> per-peer ascending send lists
        cross_row = torch.searchsorted(local_ptr, cross, right=True) - 1     # local row of each crossing edge
        cross_owner = torch.searchsorted(starts, cross_nbr, right=True) - 1
        pair = torch.unique(cross_owner * n_own + cross_row)                  # sorted by (owner, row)
        owner, row = pair // n_own, pair % n_own
        cuts = torch.searchsorted(owner, torch.arange(len(bounds) + 1, device=dev, dtype=torch.int64)).tolist()
        rows32 = row.to(torch.int32)
        for peer in range(len(bounds)):
            a, b = cuts[peer], cuts[peer + 1]
            if peer != rank and b > a:
                send.append((peer, rows32[a:b].contiguous()))
    return TilePlan(rank, lo, hi, n_own, n_halo, local_ptr.to(torch.int32).contiguous(),
                    local.to(torch.int32).contiguous(), halo_global, recv, send)


@dataclass
class DeviceTilePlan:
    """What `plan_tile` computes, built on the device with ONE small device->host read (the R x R boundary-count
    matrix and two row pointers), plus what the fused sweep kernel wants: the boundary rows as a per-row list of
    (peer, destination row) and a patch order with the boundary patches first."""
    rank: int
    lo: int
    hi: int
    n_own: int
    n_halo: int
    cap_rows: int                  # max over ranks of own + halo rows (size of the symmetric beta buffers)
    indptr: "object"               # int32[n_own + 1]
    indices: "object"              # int32[nnz_local], local numbering
    halo_global: "object"          # int64[n_halo]
    push_ptr: "object"             # int32[n_own + 1]
    push_ent: "object"             # int32[T, 2]  (peer, destination row in the peer's buffers)
    patch_order: "object"          # int32[n_patches]
    n_boundary: "object"           # int32[1] (device)
    recv: List[Tuple[int, int, int]]

    @property
    def n_total(self) -> int:
        return self.n_own + self.n_halo


def plan_tile_device(indptr, indices, nnz: int, bounds: List[Tuple[int, int]], rank: int, patch: int = 128) -> DeviceTilePlan:
    """Same partition as `plan_tile`, from the replicated global CSR (tile order) on the device.  Every rank computes the
    whole R x R matrix of boundary-row counts from its own copy of the graph, so no collective is needed."""
    import torch
    t = torch
    dev = indptr.device
    n = int(indptr.numel()) - 1
    R = len(bounds)
    lo, hi = bounds[rank]
    n_own = hi - lo
    i32, i64 = t.int32, t.int64
    starts = t.tensor([b[0] for b in bounds] + [n], device=dev, dtype=i64)
    ptr = indptr.to(i64)
    idx = indices[:nnz].to(i64)
    pos = t.arange(n, device=dev, dtype=i64)
    owner = t.bucketize(pos, starts[1:], right=True)                    # owner rank of every position
    edge_row = t.repeat_interleave(pos, ptr[1:] - ptr[:-1], output_size=nnz)
    dst_owner = owner[idx]
    cross = owner[edge_row] != dst_owner
    flags = t.zeros(n * R + 1, device=dev, dtype=i32)
    flags.scatter_(0, t.where(cross, edge_row * R + dst_owner, t.full_like(idx, n * R)),
                   t.ones(1, device=dev, dtype=i32).expand(nnz))
    flags = flags[: n * R].view(n, R)                                   # flags[i, q] = row i has a neighbour owned by q
    counts_dev = t.stack([flags[a:b].sum(0, dtype=i64) for a, b in bounds])          # counts[r, q]
    head = t.cat([counts_dev.flatten(), ptr[lo:lo + 1], ptr[hi:hi + 1]]).cpu().tolist()     # the one host read
    counts = [head[r * R:(r + 1) * R] for r in range(R)]
    e0, e1 = int(head[-2]), int(head[-1])
    n_halo_of = [sum(counts[r][q] for r in range(R)) for q in range(R)]
    n_halo = n_halo_of[rank]
    cap_rows = max(max((b[1] - b[0]) + n_halo_of[q] for q, b in enumerate(bounds)), 1)
    recv, first = [], 0
    for r in range(R):
        if r != rank and counts[r][rank] > 0:
            recv.append((r, first, counts[r][rank]))
        first += counts[r][rank]
    # where my boundary rows live in each peer's buffers: after its own rows, behind the rows of lower ranks
    base = t.tensor([(bounds[q][1] - bounds[q][0]) + sum(counts[r][q] for r in range(rank)) for q in range(R)],
                    device=dev, dtype=i64)
    my = flags[lo:hi].t().contiguous().to(i64)                          # R x n_own (scans run along the last dim)
    qrank = t.cumsum(my, 1) - my                                        # rank of row i among my rows that peer q needs
    push_cnt = my.sum(0)
    push_ptr = t.zeros(n_own + 1, device=dev, dtype=i64)
    t.cumsum(push_cnt, 0, out=push_ptr[1:])
    within = t.zeros_like(my)                                           # entries of lower peers in the same row
    for q in range(1, R):
        within[q] = within[q - 1] + my[q - 1]
    T = sum(counts[rank])
    tgt = t.where(my > 0, push_ptr[None, :-1] + within, t.full_like(my, T)).flatten()
    ent = t.zeros((T + 1, 2), device=dev, dtype=i32)
    peers = t.arange(R, device=dev, dtype=i32)[:, None].expand(R, n_own).reshape(-1)
    ent[:, 0].scatter_(0, tgt, peers)
    ent[:, 1].scatter_(0, tgt, (base[:, None] + qrank).to(i32).flatten())
    # halo rows: the outside neighbours of my rows, ascending global position (= grouped by owner)
    nbr = idx[e0:e1]
    outside = (nbr < lo) | (nbr >= hi)
    hflag = t.zeros(n + 1, device=dev, dtype=i64)
    hflag.scatter_(0, t.where(outside, nbr, t.full_like(nbr, n)), t.ones(1, device=dev, dtype=i64).expand(nbr.numel()))
    hflag = hflag[:n]
    hscan = t.cumsum(hflag, 0) - hflag
    local = t.where(outside, n_own + hscan[nbr], nbr - lo).to(i32)
    halo_global = t.zeros(n_halo + 1, device=dev, dtype=i64)
    halo_global.scatter_(0, t.where(hflag > 0, hscan, t.full_like(hscan, n_halo)), pos)
    # patches holding boundary rows first
    n_patches = max(-(-n_own // patch), 1)
    padded = t.zeros(n_patches * patch, device=dev, dtype=i64)
    padded[:n_own] = push_cnt
    bflag = (padded.view(n_patches, patch).sum(1) > 0).to(i64)
    brank = t.cumsum(bflag, 0) - bflag
    n_boundary = bflag.sum()
    pid = t.arange(n_patches, device=dev, dtype=i64)
    position = t.where(bflag > 0, brank, n_boundary + (pid - brank))
    order = t.empty(n_patches, device=dev, dtype=i32)
    order.scatter_(0, position, pid.to(i32))
    if local.numel() == 0:
        local = t.zeros(1, dtype=i32, device=dev)
    return DeviceTilePlan(rank, lo, hi, n_own, n_halo, cap_rows, (ptr[lo:hi + 1] - e0).to(i32).contiguous(),
                          local.contiguous(), halo_global[:n_halo], push_ptr.to(i32).contiguous(),
                          ent[: max(T, 1)].contiguous(), order, n_boundary.to(i32).reshape(1), recv)


def plan_tile_native(indptr, indices, nnz: int, bounds: List[Tuple[int, int]], rank: int) -> DeviceTilePlan:
    """`plan_tile_device` through libfdb200 (csrc/tile.cu): three kernels + scans, one host read of the R x R
    boundary-count matrix.  This is what TiledPath uses on the GPU."""
    import torch
    from . import _native
    from .pipeline import _ptr, _stream
    lib, check, t = _native.lib, _native.check, torch
    dev = indptr.device
    n = int(indptr.numel()) - 1
    R = len(bounds)
    lo, hi = bounds[rank]
    n_own = hi - lo
    n_own_max = max(b[1] - b[0] for b in bounds)
    hb = (C.c_int32 * (R + 1))(*([b[0] for b in bounds] + [n]))
    ws_bytes = int(lib.fdb_tile_plan_workspace_bytes(n, n_own_max, R))
    ws = t.empty(ws_bytes, dtype=t.uint8, device=dev)
    hc = (C.c_int64 * (R * R))()
    he = (C.c_int64 * 2)()
    st = _stream(t)
    check(lib.fdb_tile_plan_counts(_ptr(indptr), _ptr(indices), n, hb, R, rank, n_own_max, _ptr(ws), ws_bytes, hc, he, st),
          "tile_plan_counts")
    counts = [[int(hc[r * R + q]) for q in range(R)] for r in range(R)]
    nnz_local = int(he[1] - he[0])
    n_halo_of = [sum(counts[r][q] for r in range(R)) for q in range(R)]
    n_halo = n_halo_of[rank]
    cap_rows = max(max((b[1] - b[0]) + n_halo_of[q] for q, b in enumerate(bounds)), 1)
    recv, first = [], 0
    for r in range(R):
        if r != rank and counts[r][rank] > 0:
            recv.append((r, first, counts[r][rank]))
        first += counts[r][rank]
    base = (C.c_int64 * R)(*[(bounds[q][1] - bounds[q][0]) + sum(counts[r][q] for r in range(rank)) for q in range(R)])
    T = sum(counts[rank])
    n_patches = max(-(-n_own // 128), 1)
    i32 = t.int32
    local_ptr = t.empty(n_own + 1, dtype=i32, device=dev)
    local_idx = t.empty(max(nnz_local, 1), dtype=i32, device=dev)
    halo_global = t.empty(max(n_halo, 1), dtype=t.int64, device=dev)
    push_ptr = t.empty(n_own + 1, dtype=i32, device=dev)
    push_ent = t.empty((max(T, 1), 2), dtype=i32, device=dev)
    order = t.empty(n_patches, dtype=i32, device=dev)
    n_boundary = t.empty(1, dtype=i32, device=dev)
    check(lib.fdb_tile_plan_build(_ptr(indptr), _ptr(indices), n, hb, R, rank, n_own_max, base, _ptr(ws), ws_bytes,
                                  _ptr(local_ptr), _ptr(local_idx), _ptr(halo_global), _ptr(push_ptr), _ptr(push_ent),
                                  _ptr(order), _ptr(n_boundary), st), "tile_plan_build")
    plan = DeviceTilePlan(rank, lo, hi, n_own, n_halo, cap_rows, local_ptr, local_idx, halo_global[:n_halo], push_ptr,
                          push_ent, order, n_boundary, recv)
    plan._keepalive = ws                                  # the kernels above are still in flight
    return plan


def halo_exchange(beta, plan: TilePlan, pack: Callable, group=None, tag: int = 0):
    """Fill the halo rows of `beta` (n_total x row_floats) with the owners' current rows.

    pack(beta, row_ids) -> contiguous (len(row_ids) x row_floats) tensor of the rows to send."""
    import torch.distributed as dist
    ops, keep = [], []
    for peer, first, count in plan.recv:
        ops.append(dist.P2POp(dist.irecv, beta[plan.n_own + first: plan.n_own + first + count], peer, group))
    for peer, rows in plan.send:
        buf = pack(beta, rows)
        keep.append(buf)
        ops.append(dist.P2POp(dist.isend, buf, peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return keep


# ----------------------------------------------------------------------------------------------------
# GPU driver
# ----------------------------------------------------------------------------------------------------
class TiledPath:
    """One rank's share of a deconvolution partitioned over the ranks of a torch.distributed group.

    Every rank passes the SAME full inputs (replicated CSR / coords) or at least every CSR row of its own tile;
    only the rank's tile is sketched and solved.  Results stay sharded on the device (`beta_own`, in tile
    order) until `gather_outputs` assembles float64 arrays in input order on every rank."""

    def __init__(self, csr, coords_dev, tables, n_types: int, group=None, slice_rows: Optional[Tuple[int, int]] = None):
        """`csr` is either the full (replicated) matrix -- every rank then sketches the rows of its own tile -- or, with
        slice_rows = (a, b), only input rows [a, b): the rank sketches THOSE rows and the fused kernel writes every
        H row into the buffers of the rank that owns its tile position (peer memory), so that each rank uploads 1/R of
        the counts.  The slices of all ranks must cover all rows exactly once."""
        import torch
        import torch.distributed as dist
        from . import _native, pipeline
        if int(n_types) > pipeline.MAX_TYPES:
            raise ValueError(f"the tiled multi-GPU path supports at most {pipeline.MAX_TYPES} cell types (got {int(n_types)}); "
                             "more types run on one GPU (pipeline.DevicePath, csrc/wide.cu)")
        self.torch, self.dist, self.pl = torch, dist, pipeline
        self.lib, self.check = _native.lib, _native.check
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.csr, self.coords, self.tables, self.K = csr, coords_dev, tables, int(n_types)
        self.slice_rows = slice_rows
        self.Kp = _native.padded_types(self.K)
        self.dev = csr.indices.device
        self.gene_bucket = torch.from_numpy(tables.gene_bucket).to(self.dev)
        self.gene_weight = torch.from_numpy(tables.gene_weight).to(self.dev)
        self.d_dev = 4 * ((tables.d + 3) // 4)                     # see pipeline.DevicePath
        xst = np.zeros((self.d_dev, self.Kp), dtype=np.float32)
        xst[: tables.d, : self.K] = tables.X_sketch.T
        self.x_sketch_t = torch.from_numpy(xst).to(self.dev)
        self.gram32 = np.ascontiguousarray(tables.gram, dtype=np.float32)
        self.state = torch.zeros(16, dtype=torch.int32, device=self.dev)
        self.graph = None
        self.plan: Optional[TilePlan] = None
        import os
        self.mode = os.environ.get("FDB_TILED_MODE", "peer")         # peer (NVLink peer memory) | nccl | torch
        if os.environ.get("FDB_TILED_TORCH"):
            self.mode = "torch"
        self.comm = self._native_comm() if self.mode == "nccl" else None

    @property
    def mode_description(self) -> str:
        return {"peer": "direct NVLink peer-memory row pushes + flag/max-norm hand-shake (no NCCL on the data path)",
                "nccl": "ncclSend/ncclRecv + MAX all-reduce issued from the native loop",
                "torch": "torch.distributed isend/irecv + all_reduce"}[self.mode]

    def _native_comm(self):
        """NCCL communicator owned by libfdb200 (the solve loop issues its collectives from C).  The unique id
        travels over the existing torch.distributed group; the communicator is created once per group and
        cached (ncclCommInitRank costs seconds).  FDB_TILED_TORCH=1 keeps everything in torch."""
        import os
        if os.environ.get("FDB_TILED_TORCH"):
            return None
        key = id(self.group) if self.group is not None else None
        if key in _COMM_CACHE:
            return _COMM_CACHE[key]
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            self.check(self.lib.fdb_comm_unique_id(buf), "comm_unique_id")
        box = [buf.raw]
        self.dist.broadcast_object_list(box, src=self.dist.get_global_rank(self.group, 0) if self.group else 0,
                                        group=self.group)
        comm = C.c_void_p()
        self.check(self.lib.fdb_comm_init(self.rank, self.world, box[0], C.byref(comm)), "comm_init")
        _COMM_CACHE[key] = comm
        return comm

    def close(self):
        """Kept for symmetry; the communicator is cached per group and released by `release_communicators`."""
        self.comm = None

    def _stream(self):
        return self.pl._stream(self.torch)

    def stage_graph(self, method="knn", k=6, radius=None):
        pl, t = self.pl, self.torch
        self.graph = pl.build_graph(self.coords, method, k, radius)       # replicated, deterministic
        n = int(self.graph.order.numel())
        self.bounds = tile_bounds(n, self.world)
        if self.mode == "peer":
            self.plan = plan_tile_native(self.graph.indptr, self.graph.indices, self.graph.nnz, self.bounds, self.rank)
        else:
            self.plan = plan_tile(self.graph.indptr, self.graph.indices[: max(self.graph.nnz, 1)], self.bounds, self.rank)
        p = self.plan
        if p.indices.numel() == 0:
            p.indices = t.zeros(1, dtype=t.int32, device=self.dev)
        self.h = t.empty((max(p.n_own, 1), self.Kp), dtype=t.float32, device=self.dev)
        self.ysq = t.empty(max(p.n_own, 1), dtype=t.float32, device=self.dev)
        if self.slice_rows is not None and self.mode != "peer":
            raise RuntimeError("row-sliced inputs need the peer-memory mode (H rows are scattered over NVLink)")
        if self.mode == "peer":
            try:
                self._setup_peer()
                return self.graph
            except Exception as exc:          # no P2P mapping on this box: fall back to the NCCL loop
                if self.slice_rows is not None:
                    raise
                import warnings
                warnings.warn(f"peer-memory halo exchange unavailable ({exc!r}); using NCCL send/recv")
                self.mode = "nccl"
                self.plan = p = plan_tile(self.graph.indptr, self.graph.indices[: max(self.graph.nnz, 1)], self.bounds,
                                          self.rank)
                if self.comm is None:
                    self.comm = self._native_comm()
        self.beta_a = t.empty((max(p.n_total, 1), self.Kp), dtype=t.float32, device=self.dev)
        self.beta_b = t.empty((max(p.n_total, 1), self.Kp), dtype=t.float32, device=self.dev)
        self.send_bufs = [t.empty((rows.numel(), self.Kp), dtype=t.float32, device=self.dev) for _, rows in p.send]
        return self.graph

    def _setup_peer(self):
        """Symmetric buffers (mapped by every peer), cached per (size, group).  Layout in floats:
        [beta_a cap x Kp][beta_b cap x Kp][comm][H own_max x Kp][ysq own_max]."""
        import torch.distributed._symmetric_memory as symm_mem
        t, p, dist = self.torch, self.plan, self.dist
        cap_rows = p.cap_rows
        own_max = max(b[1] - b[0] for b in self.bounds)
        comm_floats = int(self.lib.fdb_peer_comm_floats())
        off_h = 2 * cap_rows * self.Kp + comm_floats
        off_ysq = off_h + own_max * self.Kp
        total = off_ysq + ((own_max + 63) // 64) * 64
        key = (total, id(self.group))
        if key not in _SYMM_CACHE:
            buf = symm_mem.empty(total, dtype=t.float32, device=self.dev)
            buf.zero_()
            grp = self.group if self.group is not None else dist.group.WORLD
            hdl = symm_mem.rendezvous(buf, grp)
            t.cuda.synchronize()
            dist.barrier(group=self.group)                       # every comm block is zero before anyone signals
            _SYMM_CACHE[key] = (buf, hdl, [int(x) for x in hdl.buffer_ptrs])
        self.symm_buf, self.symm_hdl, self.peer_ptrs = _SYMM_CACHE[key]
        self.cap_rows = cap_rows
        self.beta_a = self.symm_buf[: cap_rows * self.Kp].view(cap_rows, self.Kp)
        self.beta_b = self.symm_buf[cap_rows * self.Kp: 2 * cap_rows * self.Kp].view(cap_rows, self.Kp)
        self.h = self.symm_buf[off_h: off_h + max(p.n_own, 1) * self.Kp].view(max(p.n_own, 1), self.Kp)
        self.ysq = self.symm_buf[off_ysq: off_ysq + max(p.n_own, 1)]
        self.peer_h = [ptr + 4 * off_h for ptr in self.peer_ptrs]
        self.peer_ysq = [ptr + 4 * off_ysq for ptr in self.peer_ptrs]

    def stage_sketch(self):
        c, tb, p = self.csr, self.tables, self.plan
        if self.slice_rows is not None:
            a, b = self.slice_rows
            if b > a:
                R = self.world
                hb = (C.c_int32 * (R + 1))(*([x[0] for x in self.bounds] + [int(self.graph.order.numel())]))
                ph = (C.c_void_p * R)(*self.peer_h)
                py = (C.c_void_p * R)(*self.peer_ysq)
                self._row_map = self.graph.rank[a:b].contiguous()
                self.check(self.lib.fdb_sketch_contract_scatter_csr(
                    self.pl._ptr(c.indptr), int(c.indptr.dtype == self.torch.int64), self.pl._ptr(c.indices),
                    self.pl._ptr(c.data), b - a, c.shape[1], self.pl._ptr(self.gene_bucket), self.pl._ptr(self.gene_weight),
                    self.d_dev, self.pl._ptr(self.x_sketch_t), self.K, self.pl._ptr(self._row_map), int(len(tb.bucket)),
                    int(tb.linear), R, hb, ph, py, self._stream()), "sketch_contract_scatter_csr")
            self._row_ids = self.graph.order[p.lo:p.hi].contiguous()
            return
        if p.n_own == 0:
            return
        row_ids = self.graph.order[p.lo:p.hi].contiguous()
        self._row_ids = row_ids
        fn = self.lib.fdb_sketch_linear_contract_csr if tb.linear else self.lib.fdb_sketch_contract_csr
        self.check(fn(
            self.pl._ptr(c.indptr), int(c.indptr.dtype == self.torch.int64), self.pl._ptr(c.indices),
            self.pl._ptr(c.data), p.n_own, c.shape[1], self.pl._ptr(self.gene_bucket), self.pl._ptr(self.gene_weight),
            self.d_dev, self.pl._ptr(self.x_sketch_t), self.K, self.pl._ptr(None), self.pl._ptr(row_ids), int(len(tb.bucket)),
            self.pl._ptr(self.h), self.pl._ptr(self.ysq), self._stream()), "sketch_contract_csr")

    def lambda_auto(self, alpha=0.005) -> float:
        n = int(self.graph.order.numel())
        return float(alpha * float(np.mean(np.diag(self.tables.gram))) / max(self.graph.nnz / max(n, 1), 1.0))

    def rho_scaled(self, rho) -> float:
        return float(rho) * float(np.mean(np.diag(self.tables.gram)))

    def _pack(self, beta, rows, out):
        self.check(self.lib.fdb_rows_gather(self.pl._ptr(beta), self.pl._ptr(rows), rows.numel(), self.Kp,
                                            self.pl._ptr(out), self._stream()), "rows_gather")
        return out

    def _exchange(self, beta):
        dist, p = self.dist, self.plan
        ops = []
        for peer, first, count in p.recv:
            ops.append(dist.P2POp(dist.irecv, beta[p.n_own + first: p.n_own + first + count], peer, self.group))
        for (peer, rows), buf in zip(p.send, self.send_bufs):
            ops.append(dist.P2POp(dist.isend, self._pack(beta, rows, buf), peer, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def stage_solve(self, lam, rho_scaled, max_iter, tol):
        """Jacobi sweeps with halo exchange; no host synchronisation inside the loop."""
        t, p, pl = self.torch, self.plan, self.pl
        st = self._stream()
        gram = self.gram32.ctypes.data_as(C.c_void_p)
        nnz_local = int(p.indices.numel())
        self.sweep_plan = pl.build_sweep_plan(p.indptr, p.indices, p.n_own, nnz_local, self.K) if (p.n_own and max_iter) else None
        if self.mode == "peer":
            global _SEQ_BASE
            bases = (C.c_void_p * self.world)(*self.peer_ptrs)
            seq = _SEQ_BASE
            _SEQ_BASE += int(max_iter) + 8
            self.check(self.lib.fdb_bcd_solve_peer(
                pl._ptr(self.h), gram, bases, self.rank, self.world, self.cap_rows, pl._ptr(p.indptr), pl._ptr(p.indices),
                p.n_own, p.n_total, self.K, float(lam), float(rho_scaled), int(max_iter), float(tol), pl._ptr(self.state),
                pl._ptr(p.push_ptr), pl._ptr(p.push_ent), pl._ptr(p.patch_order), pl._ptr(p.n_boundary), seq,
                pl._ptr(self.sweep_plan), st),
                "bcd_solve_peer")
            return
        if self.comm is not None:
            i32, i64, vp = C.c_int32, C.c_int64, C.c_void_p
            nr, ns = len(p.recv), len(p.send)
            recv_peer = (i32 * max(nr, 1))(*[r[0] for r in p.recv])
            recv_first = (i64 * max(nr, 1))(*[r[1] for r in p.recv])
            recv_count = (i64 * max(nr, 1))(*[r[2] for r in p.recv])
            send_peer = (i32 * max(ns, 1))(*[q for q, _ in p.send])
            send_rows = (vp * max(ns, 1))(*[rows.data_ptr() for _, rows in p.send])
            send_count = (i64 * max(ns, 1))(*[rows.numel() for _, rows in p.send])
            send_buf = (vp * max(ns, 1))(*[b.data_ptr() for b in self.send_bufs])
            self.check(self.lib.fdb_bcd_solve_tiled(
                pl._ptr(self.h), gram, pl._ptr(self.beta_a), pl._ptr(self.beta_b), pl._ptr(p.indptr), pl._ptr(p.indices),
                p.n_own, p.n_total, self.K, float(lam), float(rho_scaled), int(max_iter), float(tol),
                pl._ptr(self.state), nr, recv_peer, recv_first, recv_count, ns, send_peer, send_rows, send_count,
                send_buf, self.comm, pl._ptr(self.sweep_plan), st), "bcd_solve_tiled")
            return
        self.check(self.lib.fdb_bcd_init(pl._ptr(self.beta_a), p.n_total, self.K, pl._ptr(self.state), st), "bcd_init")
        self.check(self.lib.fdb_bcd_init(pl._ptr(self.beta_b), p.n_total, self.K, pl._ptr(None), st), "bcd_init")
        norms = self.state[:2].view(t.float32)          # max-norm bit patterns of non-negative floats order like floats
        cur, nxt = self.beta_a, self.beta_b
        for _ in range(max_iter):
            if p.n_own:
                self.check(self.lib.fdb_bcd_sweep(pl._ptr(self.h), gram, pl._ptr(cur), pl._ptr(nxt), pl._ptr(p.indptr),
                                                  pl._ptr(p.indices), p.n_own, self.K, float(lam), float(rho_scaled),
                                                  float(tol), 0, pl._ptr(self.state), pl._ptr(self.sweep_plan), st),
                           "bcd_sweep")
            self._exchange(nxt)
            self.dist.all_reduce(norms, op=self.dist.ReduceOp.MAX, group=self.group)
            self.check(self.lib.fdb_bcd_finalize(pl._ptr(self.state), float(tol), st), "bcd_finalize")
            cur, nxt = nxt, cur

    def read_state(self):
        st = self.state.cpu()
        if int(st[4]) == 2:
            raise RuntimeError("peer-memory halo exchange timed out waiting for another rank")
        return int(st[3]), bool(int(st[4])), float(st[5:6].view(self.torch.float32)[0])

    def current_beta(self, n_iter):
        return self.beta_a if n_iter % 2 == 0 else self.beta_b

    def objective(self, beta_dev, lam, rho_scaled) -> float:
        t, p, pl = self.torch, self.plan, self.pl
        out = t.zeros(5, dtype=t.float64, device=self.dev)
        if p.n_own:
            self.check(self.lib.fdb_objective_terms(pl._ptr(beta_dev), pl._ptr(self.h), pl._ptr(self.ysq),
                                                    self.gram32.ctypes.data_as(C.c_void_p), pl._ptr(p.indptr),
                                                    pl._ptr(p.indices), p.n_own, self.K, pl._ptr(out), self._stream()),
                       "objective_terms")
        self.dist.all_reduce(out, op=self.dist.ReduceOp.SUM, group=self.group)
        cross, quad, lap, l1, yty = out.cpu().tolist()
        return 0.5 * (yty - 2.0 * cross + quad) + 0.5 * lam * lap + rho_scaled * l1

    def finish_sharded(self, beta_dev):
        """float64 beta / proportions of this rank's rows, compact and in TILE order: (n_own_max x 2K), beta | proportions."""
        t, p, pl = self.torch, self.plan, self.pl
        own_max = max(b[1] - b[0] for b in self.bounds)
        if not hasattr(self, "_own64") or self._own64.shape[0] != own_max:
            self._own64 = t.zeros((2, own_max, self.K), dtype=t.float64, device=self.dev)
        if p.n_own:
            self.check(self.lib.fdb_finish(pl._ptr(beta_dev), pl._ptr(None), p.n_own, self.K, pl._ptr(self._own64[0]),
                                           pl._ptr(self._own64[1]), self._stream()), "finish")
        return self._own64[0, : p.n_own], self._own64[1, : p.n_own]

    def gather_outputs(self):
        """All-gather of the ranks' own rows (tile order), then one un-permute to input order on every rank."""
        t = self.torch
        n = int(self.graph.order.numel())
        own_max = self._own64.shape[1]
        allr = t.empty((self.world, 2, own_max, self.K), dtype=t.float64, device=self.dev)
        self.dist.all_gather_into_tensor(allr, self._own64, group=self.group)
        if not hasattr(self, "_b64") or self._b64.shape[0] != n:
            self._b64 = t.empty((n, self.K), dtype=t.float64, device=self.dev)
            self._p64 = t.empty((n, self.K), dtype=t.float64, device=self.dev)
        order = self.graph.order.long()
        for q, (lo, hi) in enumerate(self.bounds):
            if hi > lo:
                self._b64.index_copy_(0, order[lo:hi], allr[q, 0, : hi - lo])
                self._p64.index_copy_(0, order[lo:hi], allr[q, 1, : hi - lo])
        return self._b64, self._p64

    def run_resident(self, *, method="knn", k=6, radius=None, lam="auto", rho=0.01, max_iter=100, tol=1e-4,
                     events=None, gather=False):
        t = self.torch

        def mark(name, fn):
            if events is None:
                return fn()
            a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            events[name] = (a, b)
            return out

        mark("graph", lambda: self.stage_graph(method, k, radius))
        mark("sketch", self.stage_sketch)
        lam_used = self.lambda_auto() if (isinstance(lam, str) and lam == "auto") else float(lam)
        rho_s = self.rho_scaled(rho)
        mark("solve", lambda: self.stage_solve(lam_used, rho_s, max_iter, tol))
        n_iter, conv, rel = self.read_state()
        beta_dev = self.current_beta(n_iter)
        obj = mark("objective", lambda: self.objective(beta_dev, lam_used, rho_s))
        b64, p64 = mark("finish", lambda: self.finish_sharded(beta_dev))
        if gather:
            b64, p64 = self.gather_outputs()
        info = dict(converged=conv, n_iterations=n_iter, final_objective=obj, objectives=[],
                    final_change=rel if max_iter else 0.0)
        return b64, p64, info, lam_used


def _host_row_slice(Y, a: int, b: int, dev):
    """Rows [a, b) of the host counts as a DeviceCSR (rebased row pointers) + the bytes that crossed PCIe."""
    import torch
    from scipy import sparse
    from . import pipeline
    if isinstance(Y, pipeline.HostCSR):
        e0, e1 = int(Y.indptr[a]), int(Y.indptr[b])
        ptr = Y.indptr[a:b + 1].to(dev, non_blocking=True)
        csr = pipeline.DeviceCSR(ptr - e0, Y.indices[e0:e1].to(dev, non_blocking=True),
                                 Y.data[e0:e1].to(dev, non_blocking=True), (b - a, Y.shape[1]))
    else:
        if not sparse.issparse(Y):
            Y = sparse.csr_matrix(np.asarray(Y)[a:b])
            a, b = 0, Y.shape[0]
        csr = pipeline.csr_to_device(Y.tocsr()[a:b])
    nbytes = sum(int(x.numel() * x.element_size()) for x in (csr.indptr, csr.indices, csr.data))
    return csr, nbytes


def deconvolve_path_tiled(Y, X, coords, gene_idx, leverage, *, sketch_dim=512, lambda_spatial="auto",
                          rho_sparsity=0.01, spatial_method="knn", k_neighbors=6, radius=None, max_iter=100,
                          tol=1e-4, random_state=0, pinned_out=False, group=None, preprocess="log_cpm",
                          y_col_mean=None, download="all"):
    """Multi-GPU counterpart of pipeline.deconvolve_path: every rank passes the same HOST inputs but uploads only its
    1/R slice of the rows; the fused sketch kernel delivers each H row to the rank that owns its spatial tile (peer
    memory), the ranks solve their tiles with per-sweep halo pushes, and the float64 outputs are all-gathered (own rows)
    so that every rank returns the full arrays.  download="rank0": only rank 0 copies the result to the host (the other
    ranks return beta = proportions = None and keep nothing but the info dict) -- R downloads of the same 16 K bytes per
    spot through shared PCIe switches are what dominates the end-to-end time of a multi-GPU call otherwise."""
    import torch
    import torch.distributed as dist
    from . import pipeline
    if download not in ("all", "rank0"):
        raise ValueError(f"download must be 'all' or 'rank0', got {download!r}")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if preprocess == "pearson" and y_col_mean is None:
        raise ValueError("preprocess='pearson' on several GPUs needs y_col_mean (per-gene means of all spots)")
    tables = pipeline.build_tables(X, gene_idx, leverage, sketch_dim, random_state, Y.shape[1], preprocess, y_col_mean)
    dev = torch.device("cuda", torch.cuda.current_device())
    n = Y.shape[0]
    a, b = (n * rank) // world, (n * (rank + 1)) // world
    c = coords if torch.is_tensor(coords) else torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float64))
    coords_dev = c.to(dev, non_blocking=True)
    mode = __import__("os").environ.get("FDB_TILED_MODE", "peer")
    if mode == "peer":
        csr, h2d = _host_row_slice(Y, a, b, dev)
        path = TiledPath(csr, coords_dev, tables, np.asarray(X).shape[0], group, slice_rows=(a, b))
    else:
        csr = pipeline.csr_to_device(Y)
        h2d = sum(int(x.numel() * x.element_size()) for x in (csr.indptr, csr.indices, csr.data))
        path = TiledPath(csr, coords_dev, tables, np.asarray(X).shape[0], group)
    h2d += int(coords_dev.numel() * coords_dev.element_size())
    b64, p64, info, lam = path.run_resident(method=spatial_method, k=k_neighbors, radius=radius, lam=lambda_spatial,
                                            rho=rho_sparsity, max_iter=max_iter, tol=tol, gather=True)
    path.close()
    if download == "rank0" and rank != 0:
        torch.cuda.current_stream().synchronize()
        res = pipeline.SolveResult(None, None, info, lam, path.graph, tables)
        res.h2d_bytes, res.d2h_bytes = h2d, 0
        return res
    if pinned_out:
        hb = torch.empty(b64.shape, dtype=torch.float64, pin_memory=True)
        hp = torch.empty(p64.shape, dtype=torch.float64, pin_memory=True)
        hb.copy_(b64, non_blocking=True)
        hp.copy_(p64, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        beta, prop = hb.numpy(), hp.numpy()
    else:
        beta, prop = b64.cpu().numpy(), p64.cpu().numpy()
    res = pipeline.SolveResult(beta, prop, info, lam, path.graph, tables)
    res.h2d_bytes = h2d
    res.d2h_bytes = int(beta.nbytes + prop.nbytes)
    return res

"""ctypes binding of libfdb200.so (the C ABI declared in include/fdb200.h).

There is NO CPU fallback: importing this module without the built library, or
calling into it without a CUDA device, raises.  Build with
``python -m flashdeconv_b200.build`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDB200_LIB", os.path.join(_HERE, "libfdb200.so"))

# every symbol include/fdb200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
SIGNATURES = {
    "fdb_abi_version": (C.c_int, []),
    "fdb_last_error": (C.c_char_p, []),
    "fdb_padded_types": (C.c_int, [C.c_int]),
    "fdb_launch_count": (C.c_longlong, []),
    "fdb_sketch_logcpm_csr": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp]),
    "fdb_sketch_project_csr": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp]),
    "fdb_contract": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "fdb_sketch_contract_csr": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _i32,
                                          _vp, _vp, _i32, _vp, _vp, _vp]),
    "fdb_sketch_linear_contract_csr": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _i32,
                                          _vp, _vp, _i32, _vp, _vp, _vp]),
    "fdb_sketch_contract_scatter_csr": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _i32,
                                                  _i32, _i32, _vp, _vp, _vp, _vp]),
    "fdb_gene_sums_csr": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "fdb_graph_workspace_bytes": (_i64, [_i64, _i32]),
    "fdb_graph_build": (C.c_int, [_vp, _i64, _i32, _i32, _f64, _vp, _vp, _vp, _vp, _i64,
                                  C.POINTER(_i64), C.POINTER(_f64), _vp, _i64, _vp]),
    "fdb_graph_build_nd": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _f64, _vp, _vp, _vp, _vp, _i64,
                                     C.POINTER(_i64), C.POINTER(_f64), _vp, _i64, _vp]),
    "fdb_graph_to_input_order": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp]),
    "fdb_bcd_init": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "fdb_bcd_sweep": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _f32, _i32, _vp, _vp, _vp]),
    "fdb_bcd_plan_bytes": (_i64, [_i64, _i64, _i32]),
    "fdb_bcd_plan_build": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _i64, _vp]),
    "fdb_bcd_finalize": (C.c_int, [_vp, _f32, _vp]),
    "fdb_bcd_solve": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _i32, _f32, _vp, _vp, _vp]),
    "fdb_objective_terms": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "fdb_bcd_solve_wide": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _i32, _f32, _vp, _vp]),
    "fdb_bcd_sweep_wide": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _f32, _i32, _vp, _vp]),
    "fdb_objective_terms_wide": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "fdb_finish": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "fdb_gene_moments_csr": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "fdb_dominant_type": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "fdb_group_sums_csr": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "fdb_rows_gather": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "fdb_peer_comm_floats": (_i64, []),
    "fdb_bcd_solve_peer": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _i64, _i64, _i32, _f32, _f32, _i32, _f32,
                                     _vp, _vp, _vp, _vp, _vp, C.c_uint32, _vp, _vp]),
    "fdb_tile_plan_workspace_bytes": (_i64, [_i64, _i64, _i32]),
    "fdb_tile_plan_counts": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _i64, _vp, _i64, _vp, _vp, _vp]),
    "fdb_tile_plan_build": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp]),
    "fdb_comm_unique_id": (C.c_int, [C.c_char_p]),
    "fdb_comm_init": (C.c_int, [_i32, _i32, C.c_char_p, C.POINTER(_vp)]),
    "fdb_comm_destroy": (C.c_int, [_vp]),
    "fdb_bcd_solve_tiled": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _f32, _f32, _i32, _f32, _vp,
                                      _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}


class NativeLibraryError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing. flashdeconv_b200 has no CPU fallback: build the CUDA "
            "library first with `python -m flashdeconv_b200.build` (needs nvcc).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.fdb_last_error().decode("utf-8", "replace")
        kind = {-1: "invalid argument", -2: "CUDA failure", -3: "workspace too small",
                -4: "unsupported"}.get(rc, f"status {rc}")
        if rc in (-1, -4):
            raise ValueError(f"libfdb200 {what}: {kind}: {msg}")
        raise RuntimeError(f"libfdb200 {what}: {kind}: {msg}")


def padded_types(k: int) -> int:
    return int(lib.fdb_padded_types(int(k)))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("flashdeconv_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch


def bind_host_to_gpu(device_index: int = 0):
    """Restricts this process to the CPU cores local to a GPU (NVML's CPU affinity of the device), so that pinned host
    buffers allocated afterwards live on the GPU's NUMA node and host<->device copies do not cross the socket
    interconnect (observed: 55 vs 70 GB/s on the same box type).  Returns the previous affinity set (pass it to
    ``os.sched_setaffinity(0, prev)`` to undo), or None when nothing was changed (no NVML, no such mask, not Linux)."""
    if os.environ.get("FDB_NO_NUMA_BIND"):
        return None
    try:
        import pynvml
        torch = require_cuda()
        props = torch.cuda.get_device_properties(device_index)
        pynvml.nvmlInit()
        try:
            bus = "%08x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        prev = os.sched_getaffinity(0)
        want = local & prev
        if not want or want == prev:
            return None
        os.sched_setaffinity(0, want)
        return prev
    except Exception:
        return None

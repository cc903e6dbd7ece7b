"""Python binding of the igd_b200 C-ABI (include/igd_b200.h) via ctypes.

Mirrors the reference's handle-based binding surface -- ``igd_py.open / search_1 / search_n /
get_nFiles`` (``/root/reference/src_py/igd_py.pyx:21-44``) and R's ``search_nr``
(``/root/reference/IGDr/src/igd_search.c:340-355``) -- on top of ``libigd_b200.so``.  Every search runs in
the CUDA library; if the library or a GPU is missing the calls raise, they never fall back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libigd_b200.so")
CLI_PATH = os.path.join(_HERE, "bin", "igd")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "igd_b200.h")


class IGDError(RuntimeError):
    pass


class _HitList(C.Structure):
    _fields_ = [("total", C.c_int64), ("n_queries", C.c_int64), ("query_off", C.POINTER(C.c_int64)),
                ("start", C.POINTER(C.c_int32)), ("end", C.POINTER(C.c_int32)), ("file", C.POINTER(C.c_int32))]


class _Queries(C.Structure):
    _fields_ = [("n", C.c_int64), ("name_id", C.POINTER(C.c_int32)), ("qs", C.POINTER(C.c_int32)),
                ("qe", C.POINTER(C.c_int32)), ("n_names", C.c_int32), ("names", C.POINTER(C.c_char_p))]


class _Overlaps(C.Structure):          # igd_b200_overlaps = overlaps_t, src/igd_base.h:116-119
    _fields_ = [("nn", C.c_int32), ("mm", C.c_int32), ("olist", C.c_void_p)]


# igd_b200_overlap = overlap_t, src/igd_base.h:109-114
OVERLAP_DTYPE = np.dtype([("idx_t", np.int32), ("idx_g", np.int32), ("idx_f", np.int32), ("sm", np.float32)])

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

_lib = None


def load_library():
    """dlopen libigd_b200.so (built by ``__graft_entry__.build()`` / ``make -C igd_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IGDError(f"{LIB_PATH} is missing: build it with `make -C igd_b200/csrc` "
                       "(igd_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.igd_b200_open.restype = vp
    L.igd_b200_open.argtypes = [C.c_char_p, C.c_int]
    L.igd_b200_close.argtypes = [vp]
    L.igd_b200_last_error.restype = C.c_char_p
    L.igd_b200_device_count.restype = C.c_int
    for f in ("nfiles", "nctg", "nbp", "gtype"):
        getattr(L, "igd_b200_" + f).restype = C.c_int32
        getattr(L, "igd_b200_" + f).argtypes = [vp]
    for f in ("ntiles", "nrecords", "device_bytes", "launch_count"):
        getattr(L, "igd_b200_" + f).restype = C.c_int64
        getattr(L, "igd_b200_" + f).argtypes = [vp]
    L.igd_b200_ntile.restype = C.c_int32
    L.igd_b200_ntile.argtypes = [vp, C.c_int32]
    L.igd_b200_ncnt.restype = C.c_int32
    L.igd_b200_ncnt.argtypes = [vp, C.c_int32, C.c_int32]
    L.igd_b200_file_nr.restype = C.c_int32
    L.igd_b200_file_nr.argtypes = [vp, C.c_int32]
    L.igd_b200_file_name.restype = C.c_char_p
    L.igd_b200_file_name.argtypes = [vp, C.c_int32]
    L.igd_b200_ctg_name.restype = C.c_char_p
    L.igd_b200_ctg_name.argtypes = [vp, C.c_int32]
    L.igd_b200_ctg_id.restype = C.c_int32
    L.igd_b200_ctg_id.argtypes = [vp, C.c_char_p]
    L.get_overlaps_n.restype = C.c_int64
    L.get_overlaps_n.argtypes = [vp, C.c_int64, C.POINTER(C.c_char_p), _i32p, _i32p, C.c_int32, _i64p]
    L.get_overlaps_n32.restype = C.c_int64
    L.get_overlaps_n32.argtypes = [vp, C.c_int64, C.POINTER(C.c_char_p), _i32p, _i32p, _i32p]
    L.igd_b200_bin_data.restype = C.c_int32
    L.igd_b200_bin_data.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp]
    L.create_iGD.restype = C.c_int
    L.create_iGD.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    L.igd_b200_count.restype = C.c_int64
    L.igd_b200_count.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_int32, _i64p]
    L.igd_b200_count_submit.restype = C.c_int
    L.igd_b200_count_submit.argtypes = [vp, C.c_int, C.c_int64, vp, vp, vp, C.c_int32]
    L.igd_b200_count_wait.restype = C.c_int64
    L.igd_b200_count_wait.argtypes = [vp, C.c_int, _i64p]
    L.igd_b200_count_device.restype = C.c_int
    L.igd_b200_count_device.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_int32, vp, vp]
    L.igd_b200_count_device_finish.restype = C.c_int
    L.igd_b200_count_device_finish.argtypes = [vp]
    L.igd_b200_count_device_tiles.restype = C.c_int
    L.igd_b200_count_device_tiles.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_int32, vp, vp, C.c_int32, C.c_int32]
    L.igd_b200_shard_cost.restype = C.c_int
    L.igd_b200_shard_cost.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp]
    L.igd_b200_shard_cuts.restype = C.c_int
    L.igd_b200_shard_cuts.argtypes = [vp, vp, C.c_int32, vp, vp]
    L.igd_b200_shard_partition.restype = C.c_int
    L.igd_b200_shard_partition.argtypes = [vp, C.c_int64, vp, vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.igd_b200_full.restype = C.c_int
    L.igd_b200_full.argtypes = [vp, C.c_int64, _i32p, _i32p, _i32p, C.POINTER(_HitList)]
    L.igd_b200_hitlist_free.argtypes = [C.POINTER(_HitList)]
    L.igd_b200_full_device.restype = C.c_int64
    L.igd_b200_full_device.argtypes = [vp, C.c_int64, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp)]
    L.igd_b200_full_restore.restype = C.c_int
    L.igd_b200_full_restore.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_int64, vp, vp, vp]
    L.igd_b200_hitsmap.restype = C.c_int64
    L.igd_b200_hitsmap.argtypes = [vp, C.c_int32, _u32p]
    L.igd_b200_hitsmap_tiles.restype = C.c_int64
    L.igd_b200_hitsmap_tiles.argtypes = [vp, C.c_int32, C.c_int64, C.c_int64, _u32p]
    L.igd_b200_hitsmap_tiles_device.restype = C.c_int
    L.igd_b200_hitsmap_tiles_device.argtypes = [vp, C.c_int32, C.c_int64, C.c_int64, vp, vp]
    L.igd_b200_read_bed.restype = C.POINTER(_Queries)
    L.igd_b200_read_bed.argtypes = [C.c_char_p]
    L.igd_b200_queries_free.argtypes = [C.POINTER(_Queries)]
    L.igd_b200_resolve.argtypes = [vp, C.POINTER(_Queries), _i32p]
    L.search.restype = C.c_int64
    L.search.argtypes = [vp, C.c_char_p, C.c_int32, _i64p]
    L.igd_b200_seqpare.restype = C.c_int
    L.igd_b200_seqpare.argtypes = [vp, C.c_char_p, _f64p]
    L.igd_b200_seq_overlaps.restype = C.c_int64
    L.igd_b200_seq_overlaps.argtypes = [vp, C.c_int64, _i32p, _i32p, _i32p, C.POINTER(C.POINTER(C.c_int64)),
                                        C.POINTER(vp)]
    L.seq_overlaps.restype = None
    L.seq_overlaps.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(_Overlaps)]
    L.igd_b200_query_cost.restype = C.c_int
    L.igd_b200_query_cost.argtypes = [vp, C.c_int64, _i32p, _i32p, _i32p, _i64p]
    L.igd_b200_last_search_kernel_ms.restype = C.c_float
    L.igd_b200_last_search_kernel_ms.argtypes = [vp]
    L.igd_b200_last_timing.argtypes = [vp, C.POINTER(C.c_float * 3)]
    L.igd_b200_last_plan.argtypes = [vp, C.POINTER(C.c_int64 * 4)]
    L.igd_b200_stage_timing.argtypes = [vp, C.c_int]
    L.igd_b200_last_stage_ms.restype = C.c_int
    L.igd_b200_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float * 4)]
    L.igd_b200_dispatcher_create.restype = vp
    L.igd_b200_dispatcher_create.argtypes = [vp, C.c_int, C.c_int, C.c_int64]
    L.igd_b200_dispatcher_handle.argtypes = [vp, vp]
    L.igd_b200_dispatcher_connect.argtypes = [vp, vp]
    L.igd_b200_dispatch.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.igd_b200_dispatch_async.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                          C.POINTER(vp)]
    L.igd_b200_dispatcher_status.argtypes = [vp]
    L.igd_b200_dispatcher_destroy.argtypes = [vp]
    L.igd_b200_count_device_shaped.argtypes = [vp, C.c_int64, vp, vp, vp, vp, C.c_int32, vp, vp]
    L.igd_b200_sharded_create.restype = vp
    L.igd_b200_sharded_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int64]
    L.igd_b200_sharded_handles.argtypes = [vp, vp]
    L.igd_b200_sharded_connect.argtypes = [vp, vp]
    L.igd_b200_sharded_count.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_int32, vp, vp]
    L.igd_b200_sharded_status.argtypes = [vp]
    L.igd_b200_sharded_destroy.argtypes = [vp]
    L.igd_b200_reducer_create.restype = vp
    L.igd_b200_reducer_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int32]
    L.igd_b200_reducer_handle.argtypes = [vp, vp]
    L.igd_b200_reducer_connect.argtypes = [vp, vp]
    L.igd_b200_reducer_allreduce.argtypes = [vp, vp, vp]
    L.igd_b200_reducer_allreduce_into.argtypes = [vp, vp, vp, vp]
    L.igd_b200_reducer_status.restype = C.c_int
    L.igd_b200_reducer_status.argtypes = [vp]
    L.igd_b200_reducer_destroy.argtypes = [vp]
    L.igd_b200_use.argtypes = [vp]
    L.igd_b200_current.restype = vp
    L.get_overlaps.restype = C.c_int32
    L.get_overlaps.argtypes = [C.c_char_p, C.c_int32, C.c_int32, _i64p]
    L.get_overlaps_v.restype = C.c_int32
    L.get_overlaps_v.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, _i64p]
    L.getOverlaps.restype = C.c_int64
    L.getOverlaps.argtypes = [C.c_char_p, _i64p]
    L.getOverlaps_v.restype = C.c_int64
    L.getOverlaps_v.argtypes = [C.c_char_p, _i64p, C.c_int32]
    L.get_overlaps0.restype = C.c_int32
    L.get_overlaps0.argtypes = [C.c_char_p, C.c_int32, C.c_int32, _i64p]
    L.getOverlaps0.restype = C.c_int64
    L.getOverlaps0.argtypes = [C.c_char_p, _i64p]
    L.getMap.restype = C.c_int64
    L.getMap_v.restype = C.c_int64
    L.seqOverlaps.argtypes = [C.c_char_p, _f64p]
    _lib = L
    return L


def last_error() -> str:
    return load_library().igd_b200_last_error().decode(errors="replace")


def create(iPath: str, oPath: str, igdName: str, bin_size: int = 16384) -> bool:
    """igd_py.create(iPath, oPath, igdName, bin_size) (src_py/igd_py.pyx:33-34): build <oPath>/<igdName>.igd from
    the BED(.gz) files under iPath -- host C, no GPU needed.  False if the database exists already."""
    return load_library().create_iGD(os.fsencode(iPath), os.fsencode(oPath), igdName.encode(), int(bin_size)) == 0


class device_view:
    """A typed window on device memory owned by the library (``__cuda_array_interface__``): ``torch.as_tensor(
    device_view(ptr, n, "<i4"), device="cuda")`` is a tensor over the same bytes, no copy."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": (int(n),), "typestr": typestr,
                                         "version": 2, "strides": None}


class HitsReducer:
    """The exchange step of the multi-GPU count path over NVLink peer memory (igd_b200_reducer_*, igd_dev_p2p.cu):
    ``allreduce`` turns each rank's device ``hits[]`` into the sum over the ranks with one small kernel per rank.
    The 64-byte IPC handles are exchanged with ``torch.distributed`` (whatever backend the group has)."""

    def __init__(self, device: int, nfiles: int, group=None):
        import torch
        import torch.distributed as dist
        self._L = load_library()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        on = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else torch.device("cpu")

        def everybody(ok: bool) -> bool:
            f = torch.tensor([1 if ok else 0], dtype=torch.int32, device=on)
            dist.all_reduce(f, op=dist.ReduceOp.MIN, group=group)
            return bool(f.item())

        # Every rank walks through the SAME collectives whether or not its own step worked, and all of them raise
        # together: a rank that failed alone would otherwise leave the others waiting in the next collective.
        err = ""
        mine = (C.c_uint8 * 64)()
        self._r = self._L.igd_b200_reducer_create(device, self.rank, self.world, nfiles)
        if not self._r or self._L.igd_b200_reducer_handle(self._r, mine):
            err = last_error() or "igd_b200_reducer_create failed"
        t = torch.tensor(list(mine), dtype=torch.uint8, device=on)
        allh = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(allh, t, group=group)
        good = everybody(not err)
        if good:
            blob = bytes(torch.cat(allh).cpu().tolist())
            if self._L.igd_b200_reducer_connect(self._r, C.c_char_p(blob)):
                err = last_error() or "igd_b200_reducer_connect failed"
            good = everybody(not err)           # also: every mailbox is mapped everywhere before the first push
        if not good:
            self.close()
            raise IGDError(err or "the peer-memory exchange could not be set up on another rank")

    def allreduce(self, d_hits_ptr: int, stream: int = 0) -> None:
        if self._L.igd_b200_reducer_allreduce(self._r, d_hits_ptr, stream):
            raise IGDError(last_error())

    def allreduce_into(self, d_in_ptr: int, d_out_ptr: int, stream: int = 0) -> None:
        """d_out = sum over the ranks of d_in; d_in is left zeroed (ready for the next batch's counts)."""
        if self._L.igd_b200_reducer_allreduce_into(self._r, d_in_ptr, d_out_ptr, stream):
            raise IGDError(last_error())

    def status(self) -> int:
        """0 = every exchange so far met all its peers (waits for the device)."""
        return int(self._L.igd_b200_reducer_status(self._r))

    def close(self) -> None:
        if getattr(self, "_r", None):
            self._L.igd_b200_reducer_destroy(self._r)
            self._r = None


def _exchange_ipc_handles(mine_bytes, group, on):
    """all ranks' 64-byte CUDA IPC handles, rank-major, exchanged with torch.distributed"""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(mine_bytes), dtype=torch.uint8, device=on)
    allh = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(allh, t, group=group)
    return bytes(torch.cat(allh).cpu().tolist())


class QueryDispatcher:
    """The query exchange of a multi-GPU search over NVLink peer memory (igd_b200_dispatcher_*, igd_dev_shard.cu): the
    partition by destination rank fused with its all-to-all -- every query is stored straight into the receive buffer
    of the rank that will search it.  ``cap`` = queries a receive buffer holds (the whole query set is always enough)."""

    def __init__(self, igd: "IGD", device: int, cap: int, group=None):
        import torch
        import torch.distributed as dist
        self._L = load_library()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        on = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else torch.device("cpu")

        def everybody(ok: bool) -> bool:
            f = torch.tensor([1 if ok else 0], dtype=torch.int32, device=on)
            dist.all_reduce(f, op=dist.ReduceOp.MIN, group=group)
            return bool(f.item())

        err = ""
        mine = (C.c_uint8 * 64)()
        self._x = self._L.igd_b200_dispatcher_create(igd._need(), self.rank, self.world, int(cap))
        if not self._x or self._L.igd_b200_dispatcher_handle(self._x, mine):
            err = last_error() or "igd_b200_dispatcher_create failed"
        blob = _exchange_ipc_handles(mine, group, on)
        good = everybody(not err)
        if good and self.world > 1:
            if self._L.igd_b200_dispatcher_connect(self._x, C.c_char_p(blob)):
                err = last_error() or "igd_b200_dispatcher_connect failed"
            good = everybody(not err)           # also: every block is mapped everywhere before the first store
        if not good:
            self.close()
            raise IGDError(err or "the peer-memory query exchange could not be set up on another rank")

    def dispatch(self, n: int, d_ctg: int, d_qs: int, d_qe: int, d_cuts: int, stream: int = 0):
        """-> (m, ctg_ptr, qs_ptr, qe_ptr, tile_lo, tile_hi): this rank's shard, device pointers into its receive buffer"""
        o = [C.c_void_p() for _ in range(3)]
        m, lo, hi = C.c_int64(), C.c_int32(), C.c_int32()
        if self._L.igd_b200_dispatch(self._x, int(n), d_ctg, d_qs, d_qe, d_cuts, stream, C.byref(o[0]), C.byref(o[1]),
                                     C.byref(o[2]), C.byref(m), C.byref(lo), C.byref(hi)):
            raise IGDError(last_error())
        return int(m.value), o[0].value, o[1].value, o[2].value, int(lo.value), int(hi.value)

    def dispatch_async(self, n: int, d_ctg: int, d_qs: int, d_qe: int, d_cuts: int, stream: int = 0):
        """-> (ctg_ptr, qs_ptr, qe_ptr, shape_ptr): nothing is read back; shape_ptr -> device {queries, tile_lo, tile_hi}
        for IGD.count_device_shaped on the same stream"""
        o = [C.c_void_p() for _ in range(4)]
        if self._L.igd_b200_dispatch_async(self._x, int(n), d_ctg, d_qs, d_qe, d_cuts, stream, C.byref(o[0]),
                                           C.byref(o[1]), C.byref(o[2]), C.byref(o[3])):
            raise IGDError(last_error())
        return o[0].value, o[1].value, o[2].value, o[3].value

    def status(self) -> None:
        """waits for the device; raises if a dispatch did not complete (a peer missing, a receive buffer too small)"""
        if self._L.igd_b200_dispatcher_status(self._x):
            raise IGDError(last_error())

    def close(self) -> None:
        if getattr(self, "_x", None):
            self._L.igd_b200_dispatcher_destroy(self._x)
            self._x = None


class ShardedCounter:
    """One multi-GPU count step as one C call (igd_b200_sharded_*): per-tile cost, its sum over the ranks, cuts, the
    query exchange, the search of the shard and the sum of hits[] -- the library's own kernels over NVLink peer memory,
    no host round trip in between.  ``cap`` >= queries of the whole set.  torch.distributed only carries the 192-byte
    IPC handles at set-up."""

    def __init__(self, igd: "IGD", device: int, cap: int, group=None):
        import torch
        import torch.distributed as dist
        self._L = load_library()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        on = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else torch.device("cpu")

        def everybody(ok: bool) -> bool:
            f = torch.tensor([1 if ok else 0], dtype=torch.int32, device=on)
            dist.all_reduce(f, op=dist.ReduceOp.MIN, group=group)
            return bool(f.item())

        err = ""
        mine = (C.c_uint8 * 192)()
        self._s = self._L.igd_b200_sharded_create(igd._need(), device, self.rank, self.world, int(cap))
        if not self._s or self._L.igd_b200_sharded_handles(self._s, mine):
            err = last_error() or "igd_b200_sharded_create failed"
        blob = _exchange_ipc_handles(mine, group, on)
        good = everybody(not err)
        if good:
            if self._L.igd_b200_sharded_connect(self._s, C.c_char_p(blob)):
                err = last_error() or "igd_b200_sharded_connect failed"
            good = everybody(not err)
        if not good:
            self.close()
            raise IGDError(err or "the sharded count step could not be set up on another rank")

    def count(self, n: int, d_ctg: int, d_qs: int, d_qe: int, v: int, d_hits: int, stream: int = 0) -> None:
        """d_hits[nFiles] (device uint64) = hits[] of all ranks' queries (overwritten); enqueued on ``stream``"""
        if self._L.igd_b200_sharded_count(self._s, int(n), d_ctg, d_qs, d_qe, int(v), d_hits, stream):
            raise IGDError(last_error())

    def status(self) -> None:
        if self._L.igd_b200_sharded_status(self._s):
            raise IGDError(last_error())

    def close(self) -> None:
        if getattr(self, "_s", None):
            self._L.igd_b200_sharded_destroy(self._s)
            self._s = None


class QuerySet:
    """Accepted lines of a BED(.gz) query file (parse_bed's rule, src/igd_base.c:53-72)."""

    def __init__(self, path: str):
        L = load_library()
        p = L.igd_b200_read_bed(os.fsencode(path))
        if not p:
            raise FileNotFoundError(path)
        try:
            q = p.contents
            self.n = int(q.n)
            mk = lambda ptr: (np.ctypeslib.as_array(ptr, (self.n,)).copy() if self.n else np.zeros(0, np.int32))
            self.name_id, self.qs, self.qe = mk(q.name_id), mk(q.qs), mk(q.qe)
            self.names = [q.names[i].decode() for i in range(q.n_names)]
        finally:
            L.igd_b200_queries_free(p)


class IGD:
    """An .igd database resident in the HBM of one GPU."""

    def __init__(self, igd_path: str | None = None, device: int = 0):
        self._L = load_library()
        self._h = None
        if igd_path is not None:
            self.open(igd_path, device)

    # -- igd_py.open (src_py/igd_py.pyx:27-30) ---------------------------------------------------------
    def open(self, igd_path: str, device: int = 0) -> "IGD":
        self.close()
        h = self._L.igd_b200_open(os.fsencode(igd_path), device)
        if not h:
            raise IGDError(last_error())
        self._h = h
        L = self._L
        self.path = igd_path
        self.device = device
        self.nfiles = L.igd_b200_nfiles(h)
        self.nctg = L.igd_b200_nctg(h)
        self.nbp = L.igd_b200_nbp(h)
        self.gtype = L.igd_b200_gtype(h)
        self.ntiles = L.igd_b200_ntiles(h)
        self.nrecords = L.igd_b200_nrecords(h)
        self.ctg_names = [L.igd_b200_ctg_name(h, c).decode(errors="replace") for c in range(self.nctg)]
        self.file_names = [L.igd_b200_file_name(h, i).decode() for i in range(self.nfiles)]
        self.file_nr = np.array([L.igd_b200_file_nr(h, i) for i in range(self.nfiles)], np.int32)
        self._ctg = {n: i for i, n in reversed(list(enumerate(self.ctg_names)))}
        return self

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.igd_b200_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _need(self):
        if not self._h:
            raise IGDError("database is not open")
        return self._h

    # -- metadata ----------------------------------------------------------------------------------
    def get_nFiles(self) -> int:                                  # igd_py.get_nFiles
        return self.nfiles

    def ctg_id(self, name: str) -> int:
        return self._ctg.get(name, -1)

    def ntile(self, ctg: int) -> int:
        return self._L.igd_b200_ntile(self._need(), ctg)

    def ncnt(self, ctg: int, tile: int) -> int:
        """iGD_t.nCnt[ctg][tile]: records of one tile."""
        return self._L.igd_b200_ncnt(self._need(), ctg, tile)

    def device_bytes(self) -> int:
        return self._L.igd_b200_device_bytes(self._need())

    def launch_count(self) -> int:
        return self._L.igd_b200_launch_count(self._need())

    def last_search_kernel_ms(self) -> float:
        return float(self._L.igd_b200_last_search_kernel_ms(self._need()))

    def last_plan(self):
        """Plan of the last settled count batch and the bytes its search kernel had to read."""
        t = (C.c_int64 * 4)()
        self._L.igd_b200_last_plan(self._need(), C.byref(t))
        return {"segments": t[0], "items": t[1], "records": t[2], "own_bytes": t[3]}

    def stage_timing(self, on: bool = True) -> None:
        self._L.igd_b200_stage_timing(self._need(), int(on))

    def last_stage_ms(self):
        t = (C.c_float * 4)()
        if self._L.igd_b200_last_stage_ms(self._need(), C.byref(t)) < 0:
            return None
        return {"seg_count": t[0], "plan": t[1], "seg_scatter": t[2], "search": t[3]}

    def last_timing(self):
        t = (C.c_float * 3)()
        self._L.igd_b200_last_timing(self._need(), C.byref(t))
        return {"h2d_ms": t[0], "kernel_ms": t[1], "d2h_ms": t[2]}

    def resolve_names(self, names) -> np.ndarray:
        return np.array([self.ctg_id(n) for n in names], np.int32)

    def resolve(self, q: QuerySet) -> np.ndarray:
        m = self.resolve_names(q.names) if q.names else np.zeros(0, np.int32)
        return m[q.name_id] if q.n else np.zeros(0, np.int32)

    # -- count search ------------------------------------------------------------------------------
    @staticmethod
    def _a(x):
        return np.ascontiguousarray(x, dtype=np.int32)

    def count(self, ctg, qs, qe, v: int = 0, hits: np.ndarray | None = None):
        """Batched count search on contig ids.  Returns (hits, overlaps found by this call)."""
        ctg, qs, qe = self._a(ctg), self._a(qs), self._a(qe)
        if hits is None:
            hits = np.zeros(self.nfiles, np.int64)
        r = self._L.igd_b200_count(self._need(), len(ctg), ctg.ctypes.data, qs.ctypes.data, qe.ctypes.data,
                                   int(v), hits)
        if r < 0:
            raise IGDError(last_error())
        return hits, int(r)

    def count_ptr(self, n: int, ctg_ptr: int, qs_ptr: int, qe_ptr: int, v: int, hits: np.ndarray) -> int:
        """Same, on raw host pointers (e.g. pinned torch tensors)."""
        r = self._L.igd_b200_count(self._need(), n, ctg_ptr, qs_ptr, qe_ptr, int(v), hits)
        if r < 0:
            raise IGDError(last_error())
        return int(r)

    def count_submit(self, slot: int, n: int, ctg_ptr: int, qs_ptr: int, qe_ptr: int, v: int = 0) -> None:
        """Pipelined count: enqueue H2D + search of one batch in slot 0 / 1 (raw host pointers, pinned for full PCIe
        speed; they must stay untouched until count_wait(slot) returns)."""
        if self._L.igd_b200_count_submit(self._need(), int(slot), int(n), ctg_ptr, qs_ptr, qe_ptr, int(v)) != 0:
            raise IGDError(last_error())

    def count_wait(self, slot: int, hits: np.ndarray) -> int:
        """Result of the batch submitted in `slot` (hits[] incremented); waits come in submission order."""
        r = self._L.igd_b200_count_wait(self._need(), int(slot), hits)
        if r < 0:
            raise IGDError(last_error())
        return int(r)

    def get_overlaps_n(self, chrm: list[str], qs, qe, v: int = 0, hits: np.ndarray | None = None):
        """search_nr(igdr, n, chrm, qs, qe) (IGDr/R/IGDr.R:54-57)."""
        qs, qe = self._a(qs), self._a(qe)
        if hits is None:
            hits = np.zeros(self.nfiles, np.int64)
        arr = (C.c_char_p * len(chrm))(*[c.encode() for c in chrm])
        r = self._L.get_overlaps_n(self._need(), len(chrm), arr, qs, qe, int(v), hits)
        if r < 0:
            raise IGDError(last_error())
        return hits, int(r)

    def search_nr(self, chrm: list[str], qs, qe) -> np.ndarray:
        """R's search_nr(igdr, n, chrm, qs, qe) (IGDr/R/IGDr.R:54-57): a fresh 32-bit hits vector."""
        qs, qe = self._a(qs), self._a(qe)
        hits = np.zeros(self.nfiles, np.int32)
        arr = (C.c_char_p * len(chrm))(*[c.encode() for c in chrm])
        if self._L.get_overlaps_n32(self._need(), len(chrm), arr, qs, qe, hits) < 0:
            raise IGDError(last_error())
        return hits

    def bin_data(self, ctg: int, bin: int):
        """R's get_binData(igdr, ichr, bin) (IGDr/src/igd_search.c:357-393), 0-based here: (idx, start, end, value)
        of one tile, read back from the HBM image."""
        n = self._L.igd_b200_bin_data(self._need(), ctg, bin, 0, None, None, None, None)
        if n < 0:
            raise IGDError(last_error())
        cols = [np.zeros(n, np.int32) for _ in range(4)]
        if n and self._L.igd_b200_bin_data(self._need(), ctg, bin, n, *[c.ctypes.data for c in cols]) < 0:
            raise IGDError(last_error())
        return tuple(cols)

    def search_1(self, chrm: str, qs: int, qe: int, hits: np.ndarray | None = None, v: int = 0):
        """igd_py.search_1 (src_py/igd_py.pyx:39-41)."""
        return self.get_overlaps_n([chrm], [qs], [qe], v, hits)

    def search_n(self, qfile: str, hits: np.ndarray | None = None, v: int = 0):
        """igd_py.search_n (src_py/igd_py.pyx:42-44): whole query file."""
        if hits is None:
            hits = np.zeros(self.nfiles, np.int64)
        r = self._L.search(self._need(), os.fsencode(qfile), int(v), hits)
        if r < 0:
            raise IGDError(last_error())
        return hits, int(r)

    def count_device(self, n: int, d_ctg: int, d_qs: int, d_qe: int, v: int, d_hits: int, stream: int = 0) -> None:
        """Device-pointer form: enqueue on a CUDA stream; d_hits is uint64[nfiles], incremented.
        Call :meth:`count_device_finish` before consuming d_hits."""
        if self._L.igd_b200_count_device(self._need(), n, d_ctg, d_qs, d_qe, int(v), d_hits, stream):
            raise IGDError(last_error())

    def count_device_tiles(self, n: int, d_ctg: int, d_qs: int, d_qe: int, v: int, d_hits: int, stream: int,
                           tile_lo: int, tile_hi: int) -> None:
        """count_device for a shard: no query touches a tile outside [tile_lo, tile_hi)."""
        if self._L.igd_b200_count_device_tiles(self._need(), n, d_ctg, d_qs, d_qe, int(v), d_hits, stream,
                                               int(tile_lo), int(tile_hi)):
            raise IGDError(last_error())

    # -- the shard plan of a multi-GPU search, on the device (device pointers; enqueued on `stream`) --------
    def shard_cost(self, n: int, d_ctg: int, d_qs: int, d_qe: int, d_tile_cost: int, stream: int = 0) -> None:
        if self._L.igd_b200_shard_cost(self._need(), n, d_ctg, d_qs, d_qe, d_tile_cost, stream):
            raise IGDError(last_error())

    def shard_cuts(self, d_tile_cost: int, n_parts: int, d_cuts: int, stream: int = 0) -> None:
        if self._L.igd_b200_shard_cuts(self._need(), d_tile_cost, int(n_parts), d_cuts, stream):
            raise IGDError(last_error())

    def shard_partition(self, n: int, d_ctg: int, d_qs: int, d_qe: int, d_cuts: int, n_parts: int, o_ctg: int, o_qs: int,
                        o_qe: int, d_part_off: int, d_reach: int = 0, stream: int = 0) -> None:
        if self._L.igd_b200_shard_partition(self._need(), n, d_ctg, d_qs, d_qe, d_cuts, int(n_parts), o_ctg, o_qs, o_qe,
                                            d_part_off, d_reach, stream):
            raise IGDError(last_error())

    def count_device_shaped(self, n_max: int, d_shape: int, d_ctg: int, d_qs: int, d_qe: int, v: int, d_hits: int,
                            stream: int = 0) -> None:
        """count_device_tiles for a batch whose shape {queries, tile_lo, tile_hi} is in device memory (QueryDispatcher)"""
        if self._L.igd_b200_count_device_shaped(self._need(), int(n_max), d_shape, d_ctg, d_qs, d_qe, int(v), d_hits, stream):
            raise IGDError(last_error())

    def count_device_finish(self) -> None:
        if self._L.igd_b200_count_device_finish(self._need()):
            raise IGDError(last_error())

    def query_cost(self, ctg, qs, qe) -> np.ndarray:
        ctg, qs, qe = self._a(ctg), self._a(qs), self._a(qe)
        cost = np.zeros(len(ctg), np.int64)
        self._L.igd_b200_query_cost(self._need(), len(ctg), ctg, qs, qe, cost)
        return cost

    # -- -f ------------------------------------------------------------------------------------------
    def full(self, ctg, qs, qe, copy: bool = True):
        """-> (query_off[n+1], start, end, file) in the reference's per-query order.  With ``copy=False`` the
        arrays are views of the handle's pinned result arena: valid until the next ``full`` on this handle."""
        ctg, qs, qe = self._a(ctg), self._a(qs), self._a(qe)
        hl = _HitList()
        if self._L.igd_b200_full(self._need(), len(ctg), ctg, qs, qe, C.byref(hl)):
            raise IGDError(last_error())
        try:
            n, t = len(ctg), int(hl.total)
            fin = (lambda a: a.copy()) if copy else (lambda a: a)
            qoff = fin(np.ctypeslib.as_array(hl.query_off, (n + 1,)))
            mk = lambda p: fin(np.ctypeslib.as_array(p, (t,))) if t else np.zeros(0, np.int32)
            return qoff, mk(hl.start), mk(hl.end), mk(hl.file)
        finally:
            self._L.igd_b200_hitlist_free(C.byref(hl))

    def full_device(self, n: int, d_ctg: int, d_qs: int, d_qe: int, stream: int = 0):
        """-f, device to device: raw device pointers in, ``(total, d_query_off, d_lists)`` out -- ``d_query_off``
        is int64[n + 1], ``d_lists`` int32 ``start[total] | end[total] | file[total]``; both are owned by the
        handle until the next ``full_device``.  :func:`device_view` wraps them for torch without a copy."""
        qoff, lst = C.c_void_p(), C.c_void_p()
        t = self._L.igd_b200_full_device(self._need(), n, d_ctg, d_qs, d_qe, stream, C.byref(qoff), C.byref(lst))
        if t < 0:
            raise IGDError(last_error())
        return int(t), qoff.value or 0, lst.value or 0

    def full_restore(self, n: int, d_order: int, d_cnt: int, d_lists_sorted: int, total: int, d_query_off: int,
                     d_lists_out: int, stream: int = 0) -> None:
        """multi-GPU -f, gathering rank: shard-ordered lists -> the caller's query order (device pointers)."""
        if self._L.igd_b200_full_restore(self._need(), n, d_order, d_cnt, d_lists_sorted, total, d_query_off,
                                         d_lists_out, stream):
            raise IGDError(last_error())

    # -- -m / -s -----------------------------------------------------------------------------------------
    def hitsmap(self, v: int = 0, tiles: tuple[int, int] | None = None) -> np.ndarray:
        """-m; ``tiles=(lo, hi)`` restricts the self-join to that range of global tile indices (one rank's
        share of a multi-GPU run: the maps of disjoint ranges add up to the whole map)."""
        hm = np.zeros((self.nfiles, self.nfiles), np.uint32)
        if tiles is None:
            r = self._L.igd_b200_hitsmap(self._need(), int(v), hm.reshape(-1))
        else:
            r = self._L.igd_b200_hitsmap_tiles(self._need(), int(v), int(tiles[0]), int(tiles[1]), hm.reshape(-1))
        if r < 0:
            raise IGDError(last_error())
        return hm

    def hitsmap_tiles_device(self, v: int, tile_lo: int, tile_hi: int, d_map: int, stream: int = 0) -> None:
        """-m over a tile range into a uint32[nFiles*nFiles] matrix in device memory (incremented), enqueued."""
        if self._L.igd_b200_hitsmap_tiles_device(self._need(), int(v), int(tile_lo), int(tile_hi), d_map, stream):
            raise IGDError(last_error())

    def seq_overlaps(self, ctg, qs, qe):
        """seq_overlaps for a batch: (query_off[n+1], structured array of overlap_t records in scan order)"""
        ctg, qs, qe = self._a(ctg), self._a(qs), self._a(qe)
        qoff, lst = C.POINTER(C.c_int64)(), C.c_void_p()
        t = self._L.igd_b200_seq_overlaps(self._need(), len(ctg), ctg, qs, qe, C.byref(qoff), C.byref(lst))
        if t < 0:
            raise IGDError(last_error())
        off = np.ctypeslib.as_array(qoff, (len(ctg) + 1,)).copy()
        rec = np.frombuffer(C.string_at(lst, 16 * t), OVERLAP_DTYPE).copy() if t else np.zeros(0, OVERLAP_DTYPE)
        return off, rec

    def seqpare(self, qfile: str) -> np.ndarray:
        sm = np.zeros(self.nfiles, np.float64)
        if self._L.igd_b200_seqpare(self._need(), os.fsencode(qfile), sm):
            raise IGDError(last_error())
        return sm

    # -- reference-named global-state entry points ---------------------------------------------------------
    def use(self) -> None:
        """Make this handle the one the reference-named symbols (get_overlaps, getOverlaps, ...) act on."""
        self._L.igd_b200_use(self._need())

"""TEST INFRASTRUCTURE ONLY: ctypes front-ends for

* ``liboracle.so`` -- the CPU restatement (oracle/igd_oracle.c), and
* ``_ref/libigdref.so`` / ``_ref/igd`` -- the unmodified reference compiled in place by oracle/Makefile.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may import
this module.  The product package ``igd_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_LIB = os.path.join(REF_DIR, "libigdref.so")
REF_BIN = os.path.join(REF_DIR, "igd")


def build(ref: bool = True) -> None:
    """Compile the restatement (and the reference, if /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "all" if ref else os.path.join(HERE, "liboracle.so")],
                   check=True, capture_output=True)


def have_ref() -> bool:
    return os.path.exists(REF_LIB) and os.path.exists(REF_BIN)


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in
                ("queries", "visited", "gated", "probes", "scanned", "hits", "bytes", "tile_bytes")]

    def asdict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class _Queries(C.Structure):
    _fields_ = [("n", C.c_int64), ("name_id", C.POINTER(C.c_int32)), ("qs", C.POINTER(C.c_int32)),
                ("qe", C.POINTER(C.c_int32)), ("n_names", C.c_int32), ("names", C.POINTER(C.c_char_p))]


_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

# overlap_t, src/igd_base.h:109-114
OVERLAP_DTYPE = np.dtype([("idx_t", np.int32), ("idx_g", np.int32), ("idx_f", np.int32), ("sm", np.float32)])


class _Overlaps(C.Structure):          # overlaps_t, src/igd_base.h:116-119
    _fields_ = [("nn", C.c_int32), ("mm", C.c_int32), ("olist", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build(ref=False)
        L = C.CDLL(LIB)
        L.orc_open.restype = C.c_void_p
        L.orc_open.argtypes = [C.c_char_p]
        L.orc_close.argtypes = [C.c_void_p]
        for f in ("orc_nfiles", "orc_nctg", "orc_nbp", "orc_gtype"):
            getattr(L, f).restype = C.c_int32
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_nrecords.restype = C.c_int64
        L.orc_nrecords.argtypes = [C.c_void_p]
        L.orc_ntile.restype = C.c_int32
        L.orc_ntile.argtypes = [C.c_void_p, C.c_int32]
        L.orc_file_nr.restype = C.c_int32
        L.orc_file_nr.argtypes = [C.c_void_p, C.c_int32]
        L.orc_file_name.restype = C.c_char_p
        L.orc_file_name.argtypes = [C.c_void_p, C.c_int32]
        L.orc_ctg_name.restype = C.c_char_p
        L.orc_ctg_name.argtypes = [C.c_void_p, C.c_int32]
        L.orc_ctg_id.restype = C.c_int32
        L.orc_ctg_id.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_read_bed.restype = C.POINTER(_Queries)
        L.orc_read_bed.argtypes = [C.c_char_p]
        L.orc_free_queries.argtypes = [C.POINTER(_Queries)]
        L.orc_resolve.argtypes = [C.c_void_p, C.POINTER(_Queries), _i32p]
        L.orc_count.restype = C.c_int64
        L.orc_count.argtypes = [C.c_void_p, C.c_int64, _i32p, _i32p, _i32p, C.c_int32, _i64p, C.POINTER(Stats)]
        L.orc_full.restype = C.c_int64
        L.orc_full.argtypes = [C.c_void_p, C.c_int64, _i32p, _i32p, _i32p, C.c_int64,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_seqpare.argtypes = [C.c_void_p, C.POINTER(_Queries), _f64p]
        L.orc_seq_overlaps.restype = C.c_int64
        L.orc_seq_overlaps.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]
        L.orc_hitsmap.restype = C.c_int64
        L.orc_hitsmap.argtypes = [C.c_void_p, C.c_int32, _u32p]
        L.orc_hitsmap_tiles.restype = C.c_int64
        L.orc_hitsmap_tiles.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, _u32p]
        _lib = L
    return _lib


class Queries:
    """Accepted query lines of a BED(.gz) file, by the reference's parse_bed rule."""

    def __init__(self, path: str):
        self._p = lib().orc_read_bed(path.encode())
        if not self._p:
            raise FileNotFoundError(path)
        q = self._p.contents
        self.n = int(q.n)
        self.name_id = np.ctypeslib.as_array(q.name_id, (self.n,)).copy() if self.n else np.zeros(0, np.int32)
        self.qs = np.ctypeslib.as_array(q.qs, (self.n,)).copy() if self.n else np.zeros(0, np.int32)
        self.qe = np.ctypeslib.as_array(q.qe, (self.n,)).copy() if self.n else np.zeros(0, np.int32)
        self.names = [q.names[i].decode() for i in range(q.n_names)]

    def __del__(self):
        if getattr(self, "_p", None):
            lib().orc_free_queries(self._p)
            self._p = None


class OracleDB:
    def __init__(self, igd_path: str):
        self.L = lib()
        self.h = self.L.orc_open(igd_path.encode())
        if not self.h:
            raise OSError(f"oracle cannot open {igd_path}")
        self.nfiles = self.L.orc_nfiles(self.h)
        self.nctg = self.L.orc_nctg(self.h)
        self.nbp = self.L.orc_nbp(self.h)
        self.nrecords = self.L.orc_nrecords(self.h)
        self.gtype = self.L.orc_gtype(self.h)
        self.ctg_names = [self.L.orc_ctg_name(self.h, c).decode(errors="replace") for c in range(self.nctg)]
        self.ntile = [self.L.orc_ntile(self.h, c) for c in range(self.nctg)]
        self.file_names = [self.L.orc_file_name(self.h, i).decode() for i in range(self.nfiles)]
        self.file_nr = np.array([self.L.orc_file_nr(self.h, i) for i in range(self.nfiles)], np.int32)

    def close(self):
        if self.h:
            self.L.orc_close(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def ctg_id(self, name: str) -> int:
        return self.L.orc_ctg_id(self.h, name.encode())

    def resolve(self, q: Queries) -> np.ndarray:
        out = np.empty(q.n, np.int32)
        if q.n:
            self.L.orc_resolve(self.h, q._p, out)
        return out

    def resolve_names(self, names: list[str]) -> np.ndarray:
        return np.array([self.ctg_id(n) for n in names], np.int32)

    @staticmethod
    def _a(x):
        return np.ascontiguousarray(x, dtype=np.int32)

    def count(self, ctg, qs, qe, v: int = 0, hits: np.ndarray | None = None, stats: bool = False):
        ctg, qs, qe = self._a(ctg), self._a(qs), self._a(qe)
        if hits is None:
            hits = np.zeros(self.nfiles, np.int64)
        st = Stats()
        ret = self.L.orc_count(self.h, len(ctg), ctg, qs, qe, v, hits, C.byref(st) if stats else None)
        return (hits, ret, st.asdict()) if stats else (hits, ret)

    def full(self, ctg, qs, qe):
        """-> (per_query counts, q, start, end, fidx) in the reference's order."""
        ctg, qs, qe = self._a(ctg), self._a(qs), self._a(qe)
        n = len(ctg)
        per = np.zeros(n, np.int64)
        total = self.L.orc_full(self.h, n, ctg, qs, qe, 0, None, None, None, None, per.ctypes.data)
        out = [np.empty(total, np.int32) for _ in range(4)]
        self.L.orc_full(self.h, n, ctg, qs, qe, total, *[o.ctypes.data for o in out], None)
        return (per, *out)

    def count_file(self, qfile: str, v: int = 0):
        """getOverlaps / getOverlaps_v on a query file: (hits, the reference's return value)."""
        q = Queries(qfile)
        hits, ret = self.count(self.resolve(q), q.qs, q.qe, v)[:2]
        return hits, ret

    def seqpare(self, q: Queries) -> np.ndarray:
        sm = np.zeros(self.nfiles, np.float64)
        self.L.orc_seqpare(self.h, q._p, sm)
        return sm

    def seq_overlaps(self, ctg: int, qs: int, qe: int) -> np.ndarray:
        """seq_overlaps of one query: structured array (idx_t, idx_g, idx_f, sm) in scan order"""
        n = self.L.orc_seq_overlaps(self.h, ctg, qs, qe, 0, None)
        out = np.zeros(n, OVERLAP_DTYPE)
        if n:
            self.L.orc_seq_overlaps(self.h, ctg, qs, qe, n, out.ctypes.data)
        return out

    def hitsmap(self, v: int = 0, tiles=None) -> np.ndarray:
        hm = np.zeros((self.nfiles, self.nfiles), np.uint32)
        if tiles is None:
            self.L.orc_hitsmap(self.h, v, hm.reshape(-1))
        else:
            self.L.orc_hitsmap_tiles(self.h, v, int(tiles[0]), int(tiles[1]), hm.reshape(-1))
        return hm


# ---------------------------------------------------------------------------------------------------
# the unmodified reference
# ---------------------------------------------------------------------------------------------------
class RefLib:
    """oracle/_ref/libigdref.so: reference objects + oracle/ref_driver.c (process-global state: one
    database open at a time, exactly like the reference)."""

    def __init__(self, igd_path: str):
        stem = os.path.basename(igd_path)[:-4]
        assert len(igd_path) < 100 and len(stem) < 60, "reference aborts on long paths (SURVEY §0.7)"
        self.L = C.CDLL(REF_LIB)
        self.L.ref_open.argtypes = [C.c_char_p]
        self.nfiles = self.L.ref_open(igd_path.encode())
        if self.nfiles < 0:
            raise OSError(igd_path)
        self.L.getOverlaps.restype = C.c_int64
        self.L.getOverlaps.argtypes = [C.c_char_p, _i64p]
        self.L.getOverlaps_v.restype = C.c_int64
        self.L.getOverlaps_v.argtypes = [C.c_char_p, _i64p, C.c_int32]
        self.L.get_overlaps.restype = C.c_int32
        self.L.get_overlaps.argtypes = [C.c_char_p, C.c_int32, C.c_int32, _i64p]
        self.L.get_overlaps_v.restype = C.c_int32
        self.L.get_overlaps_v.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, _i64p]
        self.L.seqOverlaps.argtypes = [C.c_char_p, _f64p]
        self.L.ref_getmap_dense.restype = C.c_int64
        self.L.ref_getmap_dense.argtypes = [_u32p, C.c_int32]
        self.L.getOverlaps0.restype = C.c_int64
        self.L.getOverlaps0.argtypes = [C.c_char_p, _i64p]
        self.L.get_overlaps0.restype = C.c_int32
        self.L.get_overlaps0.argtypes = [C.c_char_p, C.c_int32, C.c_int32, _i64p]
        self.gtype = self.L.ref_gtype()

    def close(self):
        self.L.ref_close()

    def count_file(self, qfile: str, v: int = 0):
        hits = np.zeros(self.nfiles, np.int64)
        if self.gtype == 0:         # igd_search's dispatch, src/igd_search.c:1024-1025: -v is not looked at
            return hits, self.L.getOverlaps0(qfile.encode(), hits)
        ret = self.L.getOverlaps_v(qfile.encode(), hits, v) if v > 0 else self.L.getOverlaps(qfile.encode(), hits)
        return hits, ret

    def count_one(self, chrm: str, qs: int, qe: int, v: int = 0, hits=None):
        if hits is None:
            hits = np.zeros(self.nfiles, np.int64)
        b = C.create_string_buffer(chrm.encode())
        if self.gtype == 0:         # src/igd_search.c:1042-1043
            return hits, self.L.get_overlaps0(b, qs, qe, hits)
        ret = (self.L.get_overlaps_v(b, qs, qe, v, hits) if v > 0 else self.L.get_overlaps(b, qs, qe, hits))
        return hits, ret

    def seqpare_file(self, qfile: str) -> np.ndarray:
        sm = np.zeros(self.nfiles, np.float64)
        self.L.seqOverlaps(qfile.encode(), sm)
        return sm

    def seq_overlaps(self, chrm: str, qs: int, qe: int) -> np.ndarray:
        """the reference's own seq_overlaps(chrm, qs, qe, olp) on an empty list (the list is malloc'ed by libc)"""
        libc = C.CDLL(None)
        libc.malloc.restype = C.c_void_p
        libc.free.argtypes = [C.c_void_p]
        o = _Overlaps(0, 16, libc.malloc(16 * 16))
        self.L.seq_overlaps.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(_Overlaps)]
        self.L.seq_overlaps.restype = None
        self.L.seq_overlaps(C.create_string_buffer(chrm.encode()), qs, qe, C.byref(o))
        out = np.frombuffer(C.string_at(o.olist, 16 * o.nn), OVERLAP_DTYPE).copy() if o.nn else np.zeros(0, OVERLAP_DTYPE)
        libc.free(o.olist)
        return out

    def hitsmap(self, v: int = 0) -> np.ndarray:
        hm = np.zeros((self.nfiles, self.nfiles), np.uint32)
        # getMap prints a progress counter to stdout (src/igd_search.c:783-784); silence fd 1
        import sys
        sys.stdout.flush()
        save = os.dup(1)
        null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(null, 1)
        try:
            self.L.ref_getmap_dense(hm.reshape(-1), v)
            C.CDLL(None).fflush(None)
        finally:
            os.dup2(save, 1)
            os.close(null)
            os.close(save)
        return hm


def ref_cli(args: list[str], cwd: str | None = None) -> str:
    """stdout of the reference binary: ``igd search ...``"""
    r = subprocess.run([REF_BIN] + args, capture_output=True, cwd=cwd)
    return r.stdout.decode(errors="replace")


def ref_create(bed_dir: str, out_dir: str, name: str, b: int = 14, gtype: int = 1) -> str:
    """the reference's own `igd create` (gtype 0: `-s 0`, gdata0_t records)"""
    os.makedirs(out_dir, exist_ok=True)
    if not bed_dir.endswith("/"):
        bed_dir += "/"
    if not out_dir.endswith("/"):
        out_dir += "/"
    subprocess.run([REF_BIN, "create", bed_dir, out_dir, name, "-b", str(b)] + (["-s", "0"] if gtype == 0 else []),
                   check=True, capture_output=True)
    return os.path.join(out_dir, name + ".igd")


def parse_full_stdout(text: str):
    """reference -f stdout -> list of (chrm, qs, qe, [(start, end, file), ...]) and the Total."""
    out, total = [], None
    for line in text.splitlines():
        if line.startswith("Query "):
            body = line[len("Query "):].rstrip().rstrip(":")
            chrm, qs, qe = [x.strip() for x in body.split(",")]
            out.append((chrm, int(qs), int(qe), []))
        elif line.startswith("Total overlaps:"):
            total = int(line.split(":")[1])
        elif line.strip():
            k, s, e, f = [x.strip() for x in line.split("\t")]
            out[-1][3].append((int(s), int(e), f))
    return out, total

# cython: language_level=3
"""igd_py -- the reference's Python module (src_py/igd_py.pyx:21-44), same class and method names, over the
C-ABI of libigd_b200.so: ``igd_py().create / open / get_nFiles / search_1 / search_n``.  A script written for the
reference (src_py/igd_test.py) runs unchanged; the search runs on the GPU, ``create`` in host C.

Differences, both additions: ``igd_py(device=k)`` selects the CUDA device, and failures raise ``RuntimeError`` with
the library's message where the reference prints and carries on.
"""
from libc.stdint cimport int32_t, int64_t


cdef extern from "igd_b200.h":
    ctypedef struct igd_b200_t:
        pass
    igd_b200_t *igd_b200_open(const char *igd_path, int device)
    void igd_b200_close(igd_b200_t *h)
    const char *igd_b200_last_error()
    int32_t igd_b200_nfiles(const igd_b200_t *h)
    int64_t get_overlaps_n(igd_b200_t *h, int64_t n, const char *const *chrm, const int32_t *qs,
                           const int32_t *qe, int32_t v, int64_t *hits)
    int64_t search(igd_b200_t *h, const char *qFile, int32_t v, int64_t *hits)
    int create_iGD(const char *iPath, const char *oPath, const char *igdName, int tile_size_bp)


cdef class igd_py:
    cdef igd_b200_t *h
    cdef int device

    def __cinit__(self, device=0):
        self.h = NULL
        self.device = device

    def __dealloc__(self):                                  # close_iGD, src_py/igd_py.pyx:27-28
        if self.h != NULL:
            igd_b200_close(self.h)
            self.h = NULL

    def get_nFiles(self):                                   # :30-31
        return igd_b200_nfiles(self.h) if self.h != NULL else 0

    def create(self, iPath, oPath, igdName, bin_size):      # :33-34 -> create_iGD (src_py/igd_create.c:19)
        create_iGD(str.encode(iPath), str.encode(oPath), str.encode(igdName), int(bin_size))

    def open(self, igdFile):                                # :36-37 -> open_iGD
        if self.h != NULL:
            igd_b200_close(self.h)
        self.h = igd_b200_open(str.encode(igdFile), self.device)
        if self.h == NULL:
            raise RuntimeError(igd_b200_last_error().decode(errors="replace"))

    def search_1(self, chrm, int32_t qs, int32_t qe, int64_t[::1] hits not None):      # :39-40 -> get_overlaps
        if self.h == NULL:
            raise RuntimeError("no database is open")
        if hits.shape[0] < igd_b200_nfiles(self.h):
            raise ValueError("hits is shorter than get_nFiles()")
        cdef bytes b = str.encode(chrm)
        cdef const char *c = b
        if get_overlaps_n(self.h, 1, &c, &qs, &qe, 0, &hits[0]) < 0:
            raise RuntimeError(igd_b200_last_error().decode(errors="replace"))

    def search_n(self, qFile, int64_t[::1] hits not None):                             # :42-44 -> getOverlaps
        """hits[] is incremented; returns the sum of hits[] afterwards, as src_py/igd_search.c:104-128 does."""
        if self.h == NULL:
            raise RuntimeError("no database is open")
        cdef int32_t nf = igd_b200_nfiles(self.h)
        if hits.shape[0] < nf:
            raise ValueError("hits is shorter than get_nFiles()")
        if search(self.h, str.encode(qFile), 0, &hits[0]) < 0:
            raise RuntimeError(igd_b200_last_error().decode(errors="replace"))
        cdef int64_t total = 0
        cdef int32_t i
        for i in range(nf):
            total += hits[i]
        return total

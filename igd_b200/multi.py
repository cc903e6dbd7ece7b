"""Multi-GPU driver of the search path: one process per GPU, the database replicated, the WORK sharded
(SURVEY.md §8(e)).  ``torch.distributed`` is the plumbing (NCCL over NVLink on the GPU box, gloo in CPU tests).

* count / -v : queries sorted by (contig, start) and cut into contiguous ranges balanced by the prefix sum of
               estimated scan bytes; per-rank ``hits[]``; ONE exchange of nFiles int64 -- with the CUDA engine and
               a CUDA communicator the library's own kernel over NVLink peer memory (``api.HitsReducer``,
               igd_dev_p2p.cu) queued right behind the search, otherwise an all-reduce of the group's backend.
* -f         : the same shards; every rank sends its overlaps to ``dst`` (sizes by all_gather, payload by
               point-to-point send/recv), which restores the ORIGINAL query order -- the result is identical to
               a single-GPU ``full()`` on the unsorted input.  With the CUDA engine and a CUDA communicator the
               lists never touch the host on the way: ``igd_b200_full_device`` leaves them in HBM, NCCL sends
               them from there, and the order is restored on ``dst``'s GPU.
* -m         : tiles are independent (getMap, src/igd_search.c:780-782): contiguous tile ranges balanced by the
               self-join cost estimate sum(nCnt^2); per-rank uint32 maps; one allreduce (two's-complement int32
               addition wraps exactly like the reference's uint32 counters).
* -s         : the greedy matching couples all queries of a (chrom, file): replicas only, not sharded.

The class is written against an ENGINE -- any object with ``nfiles``, ``count(ctg, qs, qe, v)``,
``full(ctg, qs, qe)`` and ``hitsmap(v, tiles=(lo, hi))`` -- which in the product is :class:`igd_b200.api.IGD`
(the CUDA library; there is no CPU engine in this package).  Host-side planning lives in ``workloads``.
"""
from __future__ import annotations

import numpy as np

from . import workloads


def tile_bounds(hdr: dict, n_parts: int) -> list[int]:
    """Cut the global tile range into ``n_parts`` contiguous ranges of equal estimated -m cost: a record of a
    tile with n records scans O(n) partners, so a tile costs about n^2 (+ n for the pass over its records)."""
    n = hdr["ncnt"].astype(np.float64)
    cum = np.cumsum(n * n + n)
    total = float(cum[-1]) if len(cum) else 0.0
    b = [0]
    for p in range(1, n_parts):
        b.append(int(np.searchsorted(cum, total * p / n_parts, side="left")) if total > 0 else 0)
    b.append(len(n))
    return [int(x) for x in np.maximum.accumulate(b)]


class DeviceSharder:
    """K8 on the device (igd_dev_shard.cu through the C-ABI): every rank holds ANY 1/N of the query set in HBM;
    ``shard`` leaves it holding the queries whose first bin lies in its contiguous tile range, the ranges balanced
    by the prefix sum of estimated scan bytes.  Collectives (NCCL over NVLink): one all-reduce of the per-tile cost
    (nTiles + 1 words), one all-gather of the per-destination counts and reach, and ONE all-to-all of the queries
    themselves, as (ctg, qs, qe) rows."""

    def __init__(self, engine, dev, group=None, peer_cap: int = 0):
        """peer_cap > 0: the queries travel through the library's own exchange over NVLink peer memory
        (api.QueryDispatcher, receive buffers of peer_cap queries -- the size of the whole query set is always
        enough) instead of an NCCL all-to-all."""
        import torch
        import torch.distributed as dist
        self.e, self.dev, self.group, self.dist, self.torch = engine, torch.device(dev), group, dist, torch
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.T = int(engine.ntiles)
        self.cost = torch.zeros(self.T + 1, dtype=torch.int64, device=self.dev)
        self.cuts = torch.zeros(self.world + 1, dtype=torch.int32, device=self.dev)
        self.off = torch.zeros(self.world + 1, dtype=torch.int64, device=self.dev)
        self.reach = torch.zeros(self.world, dtype=torch.int32, device=self.dev)
        self.buf = None
        self.marks = None           # diagnostics: set to [] to collect (label, CUDA event) at the phase boundaries
        self.peer = None
        if peer_cap > 0:
            from .api import QueryDispatcher
            self.peer = QueryDispatcher(engine, self.dev.index or 0, peer_cap, group)

    def close(self):
        if self.peer is not None:
            self.peer.close()
            self.peer = None

    def _mark(self, label):
        if self.marks is not None:
            ev = self.torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((label, ev))

    def shard_async(self, d_ctg, d_qs, d_qe):
        """The same with the host out of the loop (peer mode only): -> (ctg_ptr, qs_ptr, qe_ptr, shape_ptr), raw device
        pointers -- the shard lands in this rank's receive buffer and its shape {queries, tile_lo, tile_hi} in device
        memory, for ``engine.count_device_shaped`` on the same stream; nothing is read back."""
        n = int(d_ctg.shape[0])
        st = self.torch.cuda.current_stream(self.dev).cuda_stream
        self.cost.zero_()
        self.e.shard_cost(n, d_ctg.data_ptr(), d_qs.data_ptr(), d_qe.data_ptr(), self.cost.data_ptr(), st)
        self.dist.all_reduce(self.cost, group=self.group)
        self.e.shard_cuts(self.cost.data_ptr(), self.world, self.cuts.data_ptr(), st)
        return self.peer.dispatch_async(n, d_ctg.data_ptr(), d_qs.data_ptr(), d_qe.data_ptr(), self.cuts.data_ptr(), st)

    def shard(self, d_ctg, d_qs, d_qe):
        """torch int32 device tensors (this rank's slice, any order) -> (ctg, qs, qe, tile_lo, tile_hi) of its shard.
        Runs on the current torch stream."""
        torch, dist, W = self.torch, self.dist, self.world
        n = int(d_ctg.shape[0])
        st = torch.cuda.current_stream(self.dev).cuda_stream
        self._mark("start")
        self.cost.zero_()
        self.e.shard_cost(n, d_ctg.data_ptr(), d_qs.data_ptr(), d_qe.data_ptr(), self.cost.data_ptr(), st)
        self._mark("cost")
        dist.all_reduce(self.cost, group=self.group)
        self._mark("cost_allreduce")
        self.e.shard_cuts(self.cost.data_ptr(), W, self.cuts.data_ptr(), st)
        if self.peer is not None:
            # partition + all-to-all in the library's kernels: the shard arrives in this rank's receive buffer
            from .api import device_view
            m, p_c, p_s, p_e, lo, hi = self.peer.dispatch(n, d_ctg.data_ptr(), d_qs.data_ptr(), d_qe.data_ptr(),
                                                          self.cuts.data_ptr(), st)
            self._mark("dispatch")
            if m == 0:
                z = torch.empty(0, dtype=torch.int32, device=self.dev)
                return z, z, z, lo, hi
            return tuple(torch.as_tensor(device_view(p, m, "<i4"), device=self.dev) for p in (p_c, p_s, p_e)) + (lo, hi)
        if self.buf is None or self.buf.shape[1] < n:
            self.buf = torch.empty((3, max(n, 1)), dtype=torch.int32, device=self.dev)
        o = self.buf
        self.e.shard_partition(n, d_ctg.data_ptr(), d_qs.data_ptr(), d_qe.data_ptr(), self.cuts.data_ptr(), W,
                               o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), self.off.data_ptr(),
                               self.reach.data_ptr(), st)
        self._mark("cuts_partition")
        # what every rank must know about every other one -- how many queries go where, how far each destination's
        # queries reach -- in ONE all-gather (with this rank's copy of the cuts riding along), then the one host sync
        mine = torch.cat([self.off[1:] - self.off[:-1], self.reach.long(), self.cuts.long()])
        meta = torch.empty((W, mine.numel()), dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(meta, mine, group=self.group)
        self._mark("meta_gather")
        meta = meta.cpu()
        send_l, recv_l = meta[self.rank, :W].tolist(), meta[:, self.rank].tolist()
        reach = meta[:, W:2 * W].max(dim=0).values.tolist()
        cuts = meta[self.rank, 2 * W:].tolist()
        m = int(sum(recv_l))
        # the queries themselves: one all-to-all of (ctg, qs, qe) rows
        rows = torch.stack((o[0, :n], o[1, :n], o[2, :n]), dim=1)
        got = torch.empty((max(m, 1), 3), dtype=torch.int32, device=self.dev)
        dist.all_to_all_single(got[:m], rows, recv_l, send_l, group=self.group)
        self._mark("all_to_all")
        out = got[:m].t().contiguous()
        self._mark("transpose")
        if m == 0:
            out = torch.empty((3, 1), dtype=torch.int32, device=self.dev)
        lo = int(cuts[self.rank])
        return out[0, :m], out[1, :m], out[2, :m], lo, max(int(reach[self.rank]), lo)


class ShardedSearch:
    def __init__(self, engine, hdr: dict, group=None, comm_device="cpu"):
        import torch.distributed as dist
        self.e, self.hdr, self.group, self.dev = engine, hdr, group, comm_device
        self.dist = dist
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    # ---- helpers ---------------------------------------------------------------------------------
    def _t(self, a):
        import torch
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def _plan(self, ctg, qs, qe):
        ctg, qs, qe = (np.ascontiguousarray(x, np.int32) for x in (ctg, qs, qe))
        order = np.lexsort((qs, ctg))               # stable: ties keep input order, identically on every rank
        cs, ss, es = ctg[order], qs[order], qe[order]
        cost = workloads.query_cost(self.hdr, cs, ss, es)
        cum = np.cumsum(cost)
        total = int(cum[-1]) if len(cum) else 0
        b = [0] + [int(np.searchsorted(cum, total * p / self.world)) for p in range(1, self.world)] + [len(cs)]
        return order, cs, ss, es, b

    # ---- count / -v ------------------------------------------------------------------------------
    def count(self, ctg, qs, qe, v: int = 0) -> np.ndarray:
        """hits[nFiles] of the whole query set, on every rank."""
        _, cs, ss, es, b = self._plan(ctg, qs, qe)
        lo, hi = b[self.rank], b[self.rank + 1]
        if str(self.dev) != "cpu" and hasattr(self.e, "count_device"):
            return self._count_on_device(cs[lo:hi], ss[lo:hi], es[lo:hi], v)
        hits, _ = self.e.count(cs[lo:hi], ss[lo:hi], es[lo:hi], v)
        t = self._t(hits.astype(np.int64))
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def _count_on_device(self, cs, ss, es, v):
        """shard -> HBM, search, exchange by the peer-memory kernel on the same stream, one D2H of hits[]"""
        import torch

        from .api import HitsReducer
        if getattr(self, "_reducer", None) is None:
            self._reducer = HitsReducer(torch.device(self.dev).index or 0, self.e.nfiles, self.group)
            self._stream = torch.cuda.Stream(device=self.dev)
        # one real stream carries the copies, the search and the exchange (a NULL stream would mean the library's
        # own for the search and the legacy default for the exchange: no order between them)
        with torch.cuda.stream(self._stream):
            d_c, d_s, d_e = self._t(cs), self._t(ss), self._t(es)
            d_hits = torch.zeros(self.e.nfiles, dtype=torch.int64, device=self.dev)
            stream = self._stream.cuda_stream
            if len(cs):
                self.e.count_device(len(cs), d_c.data_ptr(), d_s.data_ptr(), d_e.data_ptr(), v, d_hits.data_ptr(), stream)
                self.e.count_device_finish()
            self._reducer.allreduce(d_hits.data_ptr(), stream)
            out = d_hits.cpu()
        self._stream.synchronize()
        return out.numpy()

    # ---- -f --------------------------------------------------------------------------------------
    def full(self, ctg, qs, qe, dst: int = 0):
        """On ``dst``: (query_off[n+1], start, end, file) in the input's query order, each query's overlaps in
        the reference's order -- what a single GPU returns.  None elsewhere."""
        import torch
        dist = self.dist
        if str(self.dev) != "cpu" and hasattr(self.e, "full_device"):
            return self._full_on_device(ctg, qs, qe, dst)
        order, cs, ss, es, b = self._plan(ctg, qs, qe)
        lo, hi = b[self.rank], b[self.rank + 1]
        qoff, s, e, f = self.e.full(cs[lo:hi], ss[lo:hi], es[lo:hi])
        mine = torch.tensor([int(qoff[-1])], dtype=torch.int64, device=self.dev)
        sizes = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(sizes, mine, group=self.group)
        sizes = [int(x.item()) for x in sizes]
        cnt = np.diff(qoff).astype(np.int64)
        payload = np.stack([s, e, f]).astype(np.int32) if len(s) else np.zeros((3, 0), np.int32)
        if self.rank != dst:
            dist.send(self._t(cnt), dst, group=self.group)
            if sizes[self.rank]:
                dist.send(self._t(payload), dst, group=self.group)
            return None
        cnts, pays = [], []
        for r in range(self.world):
            if r == dst:
                cnts.append(cnt), pays.append(payload)
                continue
            c = torch.empty(b[r + 1] - b[r], dtype=torch.int64, device=self.dev)
            dist.recv(c, r, group=self.group)
            cnts.append(c.cpu().numpy())
            p = torch.empty((3, sizes[r]), dtype=torch.int32, device=self.dev)
            if sizes[r]:
                dist.recv(p, r, group=self.group)
            pays.append(p.cpu().numpy())
        cnt_sorted = np.concatenate(cnts) if cnts else np.zeros(0, np.int64)
        pay = np.concatenate(pays, axis=1)
        # sorted position of every original query, then a gather of its run of overlaps
        n = len(order)
        pos = np.empty(n, np.int64)
        pos[order] = np.arange(n)
        soff = np.concatenate([[0], np.cumsum(cnt_sorted)])
        cnt_orig = cnt_sorted[pos]
        qoff_o = np.concatenate([[0], np.cumsum(cnt_orig)]).astype(np.int64)
        total = int(qoff_o[-1])
        src = np.repeat(soff[pos], cnt_orig) + (np.arange(total) - np.repeat(qoff_o[:-1], cnt_orig))
        return qoff_o, pay[0][src], pay[1][src], pay[2][src]

    def _pinned(self, name, n, dtype):
        """a page-locked host buffer kept between calls (the lists come back at PCIe speed, like a single GPU's arena)"""
        import torch
        bufs = self.__dict__.setdefault("_pin", {})
        t = bufs.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(int(n * 1.25), 1024), dtype=dtype).pin_memory()
            bufs[name] = t
        return t[:n]

    def _full_on_device(self, ctg, qs, qe, dst):
        """-f with everything after the upload on the devices: the shard plan (per-tile cost of the whole set by the
        library's kernels, cuts, each query's destination = the range its first tile falls in), the search
        (igd_b200_full_device: lists left in HBM), the gather over NCCL, the caller's order restored by
        igd_b200_full_restore, and ONE D2H of the result into page-locked buffers kept by this object (the arrays
        returned are views of them, valid until the next call)."""
        import torch

        from .api import device_view
        dist, dev, W, e, hdr = self.dist, self.dev, self.world, self.e, self.hdr
        n = len(ctg)
        st = torch.cuda.current_stream(dev).cuda_stream
        d = [self._t(np.ascontiguousarray(x, np.int32)) for x in (ctg, qs, qe)]
        T = int(e.ntiles)
        if getattr(self, "_tb", None) is None:
            self._tb = torch.from_numpy(np.ascontiguousarray(hdr["tbase"], np.int64)).to(dev)
            self._nt = torch.from_numpy(np.ascontiguousarray(hdr["ntile"], np.int64)).to(dev)
        cost = torch.zeros(T + 1, dtype=torch.int64, device=dev)
        cuts = torch.zeros(W + 1, dtype=torch.int32, device=dev)
        e.shard_cost(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), cost.data_ptr(), st)
        e.shard_cuts(cost.data_ptr(), W, cuts.data_ptr(), st)
        c = d[0].long()
        ok = (c >= 0) & (c < hdr["nctg"])
        c0 = torch.where(ok, c, torch.zeros_like(c))
        n1 = torch.div(d[1].long(), hdr["nbp"], rounding_mode="floor")
        ok &= (d[1] >= 0) & (n1 < self._nt[c0])
        tile = torch.where(ok, self._tb[c0] + n1, torch.full_like(c, T))
        dest = torch.where(ok, torch.bucketize(tile, cuts[1:-1].long().contiguous(), right=True),
                           torch.full_like(c, W - 1))
        perm = torch.argsort(dest, stable=True)         # concatenated shard order -> index in the caller's arrays
        off = [0] + torch.cumsum(torch.bincount(dest, minlength=W), 0).tolist()       # (host sync)
        mine = perm[off[self.rank]:off[self.rank + 1]]
        m = int(mine.numel())
        sc, ss, se = (x[mine].contiguous() for x in d)
        total, p_qoff, p_lists = e.full_device(m, sc.data_ptr(), ss.data_ptr(), se.data_ptr(), st)
        qoff = torch.as_tensor(device_view(p_qoff, m + 1, "<i8"), device=dev)
        cnt = (qoff[1:] - qoff[:-1]).contiguous()
        pay = (torch.as_tensor(device_view(p_lists, 3 * total, "<i4"), device=dev).view(3, total) if total
               else torch.zeros((3, 0), dtype=torch.int32, device=dev))
        tot_t = torch.tensor([total], dtype=torch.int64, device=dev)
        sizes = torch.empty(W, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, tot_t, group=self.group)
        sizes = sizes.tolist()
        if self.rank != dst:
            ops = [dist.P2POp(dist.isend, cnt, dst, group=self.group)] if m else []
            if total:
                ops.append(dist.P2POp(dist.isend, pay.contiguous(), dst, group=self.group))
            for w in (dist.batch_isend_irecv(ops) if ops else []):
                w.wait()
            torch.cuda.current_stream(dev).synchronize()    # the handle's list buffers are reused by the next call
            return None
        cnts, pays, ops = [], [], []
        for r in range(W):
            if r == dst:
                cnts.append(cnt), pays.append(pay)
                continue
            cr = torch.empty(off[r + 1] - off[r], dtype=torch.int64, device=dev)
            pr = torch.empty((3, sizes[r]), dtype=torch.int32, device=dev)
            if cr.numel():
                ops.append(dist.P2POp(dist.irecv, cr, r, group=self.group))
            if sizes[r]:
                ops.append(dist.P2POp(dist.irecv, pr, r, group=self.group))
            cnts.append(cr), pays.append(pr)
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        cnt_sorted, pay_all = torch.cat(cnts).contiguous(), torch.cat(pays, dim=1).contiguous()
        tot = int(pay_all.shape[1])
        # the caller's query order, restored by the library's kernels (igd_b200_full_restore)
        qoff_o = torch.empty(n + 1, dtype=torch.int64, device=dev)
        out = torch.empty((3, max(tot, 1)), dtype=torch.int32, device=dev)
        e.full_restore(n, perm.data_ptr(), cnt_sorted.data_ptr(), pay_all.data_ptr(), tot, qoff_o.data_ptr(),
                       out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        h_qoff, h_out = self._pinned("qoff", n + 1, torch.int64), self._pinned("lists", 3 * max(tot, 1), torch.int32)
        h_qoff.copy_(qoff_o, non_blocking=True)
        h_out.copy_(out.view(-1), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        h = h_out.numpy().reshape(3, max(tot, 1))
        return (h_qoff.numpy(), h[0, :tot], h[1, :tot], h[2, :tot])

    # ---- -m --------------------------------------------------------------------------------------
    def hitsmap(self, v: int = 0) -> np.ndarray:
        """uint32[nFiles, nFiles] of the whole database, on every rank."""
        b = tile_bounds(self.hdr, self.world)
        if str(self.dev) != "cpu" and hasattr(self.e, "hitsmap_tiles_device"):
            # the rank's matrix stays in HBM: self-join of its tiles, all-reduce (NCCL over NVLink), one D2H
            import torch
            nf = self.e.nfiles
            t = torch.zeros(nf * nf, dtype=torch.int32, device=self.dev)
            self.e.hitsmap_tiles_device(v, b[self.rank], b[self.rank + 1], t.data_ptr(),
                                        torch.cuda.current_stream(self.dev).cuda_stream)
            self.dist.all_reduce(t, group=self.group)
            return t.cpu().numpy().view(np.uint32).reshape(nf, nf)
        hm = self.e.hitsmap(v, tiles=(b[self.rank], b[self.rank + 1]))
        t = self._t(hm.view(np.int32))
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy().view(np.uint32)

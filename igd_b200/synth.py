"""Deterministic synthetic inputs for the IGD search path.

Two producers of a database:

* :func:`write_bed_files` -- BED text files, to be indexed by the *reference* ``igd create``
  (``/root/reference/src/igd_create.c:25-121``); used for oracle-scale cases and golden fixtures.
* :func:`write_igd_direct` -- writes ``<stem>.igd`` + ``<stem>_index.tsv`` straight from arrays in the
  on-disk layout of ``igd_save`` (``/root/reference/src/igd_base.c:396-461``) for the 1e8 / 1e9 interval
  cases where a text round trip is pointless.  It is cross-checked against ``igd create`` in
  ``tests/test_synth.py`` (identical modulo the order of equal starts inside a tile, which the
  reference's unstable radix sort leaves unspecified).

Everything here is host-side numpy; nothing is on the search hot path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

# hg38 primary contigs; the reference embeds no genome.
HG38 = [
    ("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555),
    ("chr5", 181538259), ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636),
    ("chr9", 138394717), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
    ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345),
    ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
    ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415),
]
HG38_NAMES = [n for n, _ in HG38]
HG38_SIZES = np.array([s for _, s in HG38], dtype=np.int64)

SEED_BASE = 20261017
QUERY_FILE_INDEX = 999999


def seed_for(config: int, file_index: int) -> int:
    return SEED_BASE + 1000 * config + file_index


@dataclass
class Intervals:
    """A bag of intervals on the hg38 contig list (ctg is an index into ``names``)."""
    ctg: np.ndarray     # int32
    start: np.ndarray   # int32
    end: np.ndarray     # int32
    value: np.ndarray   # int32 (BED score column)

    def __len__(self) -> int:
        return int(self.ctg.shape[0])


def _lengths(rng: np.random.Generator, n: int, law: tuple) -> np.ndarray:
    kind = law[0]
    if kind == "lognormal":
        _, mu, sigma = law
        ln = np.floor(rng.lognormal(mu, sigma, n)).astype(np.int64)
        return np.clip(ln, 1, 1_000_000)
    if kind == "fixed":
        return np.full(n, int(law[1]), dtype=np.int64)
    raise ValueError(law)


def gen_intervals(seed: int, n: int, law: tuple, sizes: np.ndarray = HG38_SIZES) -> Intervals:
    """chrom ∝ size; start ~ U[0, size-len); score ~ U{0..1000}."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    p = sizes / sizes.sum()
    ctg = rng.choice(len(sizes), size=n, p=p).astype(np.int32)
    ln = _lengths(rng, n, law)
    room = np.maximum(sizes[ctg] - ln, 1)
    start = np.floor(rng.random(n) * room).astype(np.int64)
    end = np.minimum(start + ln, sizes[ctg])
    value = rng.integers(0, 1001, size=n, dtype=np.int32)
    return Intervals(ctg, start.astype(np.int32), end.astype(np.int32), value)


def gen_hotspot_queries(seed: int, n: int, length: int, n_hot: int = 64, frac_hot: float = 0.8,
                        sd: float = 2e6, sizes: np.ndarray = HG38_SIZES) -> Intervals:
    """cfg 5: long queries, 80 % drawn around hotspot centres (start ~ N(centre, sd) clipped)."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    p = sizes / sizes.sum()
    hot_ctg = rng.choice(len(sizes), size=n_hot, p=p)
    hot_pos = np.floor(rng.random(n_hot) * sizes[hot_ctg]).astype(np.int64)
    is_hot = rng.random(n) < frac_hot
    which = rng.integers(0, n_hot, size=n)
    ctg = np.where(is_hot, hot_ctg[which], rng.choice(len(sizes), size=n, p=p)).astype(np.int32)
    room = np.maximum(sizes[ctg] - length, 1)
    uni = np.floor(rng.random(n) * room).astype(np.int64)
    nor = np.floor(hot_pos[which] + rng.normal(0.0, sd, n)).astype(np.int64)
    start = np.where(is_hot, np.clip(nor, 0, room - 1), uni)
    end = np.minimum(start + length, sizes[ctg])
    value = np.zeros(n, dtype=np.int32)
    return Intervals(ctg, start.astype(np.int32), end.astype(np.int32), value)


def write_bed(path: str, iv: Intervals, names=HG38_NAMES) -> None:
    """5-column BED: chrom, start, end, '.', score."""
    nm = np.array(names, dtype=object)
    with open(path, "w") as f:
        # chunked to keep the temporary list bounded
        for a in range(0, len(iv), 1 << 20):
            b = min(len(iv), a + (1 << 20))
            rows = zip(nm[iv.ctg[a:b]], iv.start[a:b].tolist(), iv.end[a:b].tolist(), iv.value[a:b].tolist())
            f.write("".join(f"{c}\t{s}\t{e}\t.\t{v}\n" for c, s, e, v in rows))


def write_bed_files(outdir: str, config: int, n_files: int, n_per_file: int, law: tuple) -> list[Intervals]:
    """f%05d.bed so that glob order (igd_create.c:33-39) == file index."""
    os.makedirs(outdir, exist_ok=True)
    out = []
    for i in range(n_files):
        iv = gen_intervals(seed_for(config, i), n_per_file, law)
        write_bed(os.path.join(outdir, f"f{i:05d}.bed"), iv)
        out.append(iv)
    return out


# ---------------------------------------------------------------------------------------------
# direct .igd writer
# ---------------------------------------------------------------------------------------------

def _expand_sort(crank, start, end, nbp, m, use_cuda=None):
    """Replicate every interval into each bin it touches (igd_add, igd_base.c:162-172) and order the
    copies by (contig rank, bin, start), stable in (file, line) order.  Returns (source index of every
    record in file order, nCnt per tile, nTile per contig, total tiles).  The heavy repeat/sort runs on
    the GPU through torch when one is present (input generation only -- not the search path)."""
    if use_cuda is None:
        try:
            import torch
            use_cuda = torch.cuda.is_available() and start.shape[0] > 2_000_000
        except Exception:
            use_cuda = False
    if use_cuda:
        import torch
        dev = torch.device("cuda")
        t_rank = torch.from_numpy(np.ascontiguousarray(crank)).to(dev)
        t_s = torch.from_numpy(start).to(dev).long()
        t_e = torch.from_numpy(end).to(dev).long()
        n1 = t_s // nbp
        n2 = (t_e - 1) // nbp
        ntile = torch.zeros(m, dtype=torch.int64, device=dev)
        ntile.scatter_reduce_(0, t_rank, n2 + 1, reduce="amax", include_self=True)
        tile_base = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(ntile, 0)])
        span = n2 - n1 + 1
        src = torch.repeat_interleave(torch.arange(t_s.shape[0], device=dev), span)
        first = torch.cumsum(span, 0) - span
        within = torch.arange(src.shape[0], device=dev) - first[src]
        gtile = tile_base[t_rank[src]] + n1[src] + within
        del within, first
        key = (gtile << 32) | t_s[src]
        del gtile
        key, order = torch.sort(key, stable=True)
        src = src[order]
        del order
        total_tiles = int(tile_base[-1].item())
        ncnt = torch.bincount(key >> 32, minlength=total_tiles).to(torch.int32)
        out = (src.cpu().numpy(), ncnt.cpu().numpy(), ntile.to(torch.int32).cpu().numpy(), total_tiles)
        del key, src
        torch.cuda.empty_cache()
        return out
    n1 = start.astype(np.int64) // nbp
    n2 = (end.astype(np.int64) - 1) // nbp
    ntile = np.zeros(m, dtype=np.int32)
    np.maximum.at(ntile, crank, (n2 + 1).astype(np.int32))
    tile_base = np.concatenate([[0], np.cumsum(ntile.astype(np.int64))])
    span = n2 - n1 + 1
    rep = np.repeat(np.arange(len(start), dtype=np.int64), span)
    within = np.arange(rep.shape[0], dtype=np.int64) - np.repeat(np.cumsum(span) - span, span)
    gtile = tile_base[crank[rep]] + n1[rep] + within
    del within
    key = (gtile << 32) | start[rep].astype(np.int64)
    order = np.argsort(key, kind="stable")
    del key
    rep = rep[order]
    gtile = gtile[order]
    total_tiles = int(tile_base[-1])
    ncnt = np.bincount(gtile, minlength=total_tiles).astype(np.int32)
    return rep, ncnt, ntile, total_tiles


def write_igd_direct(stem: str, files: list[Intervals] | None = None, *, nbp: int = 16384,
                     names=HG38_NAMES, file_names: list[str] | None = None,
                     gen=None, n_files: int | None = None, ctg_order: list[int] | None = None,
                     use_cuda: bool | None = None, gtype: int = 1) -> dict:
    """Write ``stem + '.igd'`` and ``stem + '_index.tsv'``.  ``gtype`` 1 writes gdata_t records (16 B), 0 the
    value-less gdata0_t records of ``igd create -s 0`` (12 B, /root/reference/src/igd_base.h:48-52).

    ``files`` is a list of per-file interval bags, or pass ``gen(i) -> Intervals`` with ``n_files``.
    Layout follows /root/reference/src/igd_base.c:408-458 (header, nTile[], nCnt[][], 40-byte names,
    tiles contig-major, each tile sorted by start) and /root/reference/src/igd_create.c:96-111 (tsv).
    Contig order defaults to first appearance over (file, line) like igd_add (igd_base.c:127-136).
    Intervals with start >= end are dropped from tiles but still counted in the tsv (igd_create.c:70).
    Returns summary statistics.
    """
    if files is None:
        assert gen is not None and n_files is not None
        files = [gen(i) for i in range(n_files)]
    n_files = len(files)
    if file_names is None:
        file_names = [f"f{i:05d}.bed" for i in range(n_files)]

    nr = np.array([len(f) for f in files], dtype=np.int64)
    ctg = np.concatenate([f.ctg for f in files]).astype(np.int32)
    start = np.concatenate([f.start for f in files]).astype(np.int32)
    end = np.concatenate([f.end for f in files]).astype(np.int32)
    value = np.concatenate([f.value for f in files]).astype(np.int32)
    fidx = np.repeat(np.arange(n_files, dtype=np.int32), nr)
    avg = np.array([float((f.end.astype(np.int64) - f.start.astype(np.int64)).sum()) / max(len(f), 1)
                    for f in files])
    del files

    keep = start < end                       # igd_add: if(s >= e) return  (igd_base.c:120)
    ctg, start, end, value, fidx = ctg[keep], start[keep], end[keep], value[keep], fidx[keep]

    if ctg_order is None:
        _, first = np.unique(ctg, return_index=True)
        ctg_order = ctg[np.sort(first)].tolist()
    rank = np.full(len(names), -1, dtype=np.int64)
    rank[np.array(ctg_order, dtype=np.int64)] = np.arange(len(ctg_order))
    m = len(ctg_order)

    rep, ncnt, ntile, total_tiles = _expand_sort(rank[ctg], start, end, nbp, m, use_cuda)

    assert gtype in (0, 1)
    W = 4 if gtype == 1 else 3
    rec = np.empty((rep.shape[0], W), dtype=np.int32)   # gdata_t {idx,start,end,value} igd_base.h:41-46
    rec[:, 0] = fidx[rep]
    rec[:, 1] = start[rep]
    rec[:, 2] = end[rep]
    if gtype == 1:
        rec[:, 3] = value[rep]

    with open(stem + ".igd", "wb") as f:
        np.array([nbp, gtype, m], dtype=np.int32).tofile(f)
        ntile.tofile(f)
        ncnt.tofile(f)
        for c in ctg_order:
            f.write(names[c].encode().ljust(40, b"\0"))
        rec.tofile(f)
    with open(stem + "_index.tsv", "w") as f:
        f.write("Index\tFile\tNumber of regions\tAvg size\n")
        for i in range(n_files):
            f.write("%i\t%s\t%i\t%f\n" % (i, file_names[i], nr[i], avg[i]))
    return {"n_files": n_files, "n_intervals": int(nr.sum()), "n_records": int(rec.shape[0]),
            "n_tiles": total_tiles, "n_ctg": m, "bytes": 12 + 44 * m + 4 * total_tiles + 4 * W * int(rec.shape[0])}


# ---------------------------------------------------------------------------------------------
# the same writer for the 1e8 / 1e9-interval benchmark databases: flat arrays in, contig by contig
# ---------------------------------------------------------------------------------------------

def write_igd_flat(stem: str, ctg: np.ndarray, start: np.ndarray, end: np.ndarray, value: np.ndarray,
                   per_file: int, *, nbp: int = 16384, names=HG38_NAMES, device: str | None = None) -> dict:
    """``write_igd_direct`` for inputs too large to shuffle with numpy: the intervals of all files as flat arrays in
    (file, line) order, ``per_file`` of them per file.  Same bytes as ``write_igd_direct`` (checked in
    tests/test_synth_flat.py): contig order by first appearance, tiles contig-major, each tile ordered by
    start, stable in (file, line).  Works contig by contig with torch on the GPU when there is one (input
    generation only -- not the search path), so that 1e9 intervals need a few GB at a time."""
    import torch
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    n = int(ctg.shape[0])
    n_files = n // per_file
    assert n_files * per_file == n
    t_ctg = torch.from_numpy(np.ascontiguousarray(ctg)).to(dev)
    t_s = torch.from_numpy(np.ascontiguousarray(start)).to(dev)
    t_e = torch.from_numpy(np.ascontiguousarray(end)).to(dev)
    t_v = torch.from_numpy(np.ascontiguousarray(value)).to(dev)
    # contig order = first appearance over (file, line) among the intervals that are kept (igd_add drops s >= e)
    keep = t_s < t_e
    first = torch.full((len(names),), n, dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, t_ctg.long()[keep], torch.arange(n, device=dev)[keep], reduce="amin", include_self=True)
    first = first.cpu().numpy()
    ctg_order = [int(c) for c in np.argsort(first, kind="stable") if first[c] < n]
    m = len(ctg_order)
    ntile, ncnt, recs = [], [], []
    for c in ctg_order:
        sel = torch.nonzero((t_ctg == c) & keep).squeeze(1)               # ascending = (file, line) order
        s = t_s[sel].long()
        e = t_e[sel].long()
        n1 = torch.div(s, nbp, rounding_mode="floor")
        n2 = torch.div(e - 1, nbp, rounding_mode="floor")
        nt = int(n2.max().item()) + 1
        span = n2 - n1 + 1
        rep = torch.repeat_interleave(torch.arange(sel.shape[0], device=dev), span)
        firstpos = torch.cumsum(span, 0) - span
        tile = n1[rep] + (torch.arange(rep.shape[0], device=dev) - firstpos[rep])
        del firstpos, span, n1, n2
        key = (tile << 32) | s[rep]
        key, order = torch.sort(key, stable=True)
        src = sel[rep[order]]
        del rep, order
        ncnt.append(torch.bincount(key >> 32, minlength=nt).to(torch.int32).cpu().numpy())
        del key, tile
        rec = torch.stack([torch.div(src, per_file, rounding_mode="floor").int(), t_s[src].int(), t_e[src].int(),
                           t_v[src].int()], dim=1).contiguous()
        recs.append(rec.cpu().numpy())
        ntile.append(nt)
        del rec, src, sel, s, e
    with open(stem + ".igd", "wb") as f:
        np.array([nbp, 1, m], dtype=np.int32).tofile(f)
        np.array(ntile, dtype=np.int32).tofile(f)
        for x in ncnt:
            x.tofile(f)
        for c in ctg_order:
            f.write(names[c].encode().ljust(40, b"\0"))
        for r in recs:
            r.tofile(f)
    # _index.tsv: all lines counted, average size over all of them (igd_create.c:70,96-111)
    ln = (t_e.long() - t_s.long()).reshape(n_files, per_file).sum(1).double().cpu().numpy() / max(per_file, 1)
    with open(stem + "_index.tsv", "w") as f:
        f.write("Index\tFile\tNumber of regions\tAvg size\n")
        for i in range(n_files):
            f.write("%i\t%s\t%i\t%f\n" % (i, f"f{i:05d}.bed", per_file, ln[i]))
    n_rec = int(sum(r.shape[0] for r in recs))
    n_tiles = int(sum(ntile))
    if dev.type == "cuda":
        del t_ctg, t_s, t_e, t_v, keep
        torch.cuda.empty_cache()
    return {"n_files": n_files, "n_intervals": n, "n_records": n_rec, "n_tiles": n_tiles, "n_ctg": m,
            "bytes": 12 + 44 * m + 4 * n_tiles + 16 * n_rec}

"""BASELINE.json's benchmark configurations as synthetic inputs.

Host-side setup only: builds (and caches under a work directory) the ``.igd`` database of a named
configuration with :func:`igd_b200.synth.write_igd_direct` and generates its query set.
"""
from __future__ import annotations

import json
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np

from . import synth

CONFIGS = {
    # 10 files x 10k; the reference's own CPU-runnable case
    "cfg1": dict(config=1, n_files=10, per_file=10_000, db_law=("lognormal", 6.5, 1.2),
                 n_queries=10_000, q_law=("lognormal", 6.5, 1.2)),
    # Roadmap-scale: 1,905 files x 50k (~1e8 intervals), 1M x 1 kb queries, count mode, 1 GPU
    "cfg2": dict(config=2, n_files=1905, per_file=50_000, db_law=("lognormal", 6.0, 1.0),
                 n_queries=1_000_000, q_law=("fixed", 1000)),
    # 1e9 intervals (10k files), 10M queries, query-sharded over 1/2/4/8 GPUs
    "cfg3": dict(config=3, n_files=10_000, per_file=100_000, db_law=("lognormal", 6.0, 1.0),
                 n_queries=10_000_000, q_law=("fixed", 1000)),
    # cfg3's per-tile statistics (5.5k records and ~53 queries per tile, 10k files) on a genome 1/10 the size:
    # what profiling runs use, since the full cfg3 database takes minutes to generate
    "cfg3s": dict(config=3, n_files=10_000, per_file=10_000, db_law=("lognormal", 6.0, 1.0),
                  n_queries=1_000_000, q_law=("fixed", 1000), genome_scale=0.1),
    # a 1/10 cut of cfg2 for smoke runs and CPU-side tests
    "mini": dict(config=2, n_files=200, per_file=50_000, db_law=("lognormal", 6.0, 1.0),
                 n_queries=100_000, q_law=("fixed", 1000)),
}


def genome_sizes(name: str) -> np.ndarray:
    scale = CONFIGS[name].get("genome_scale", 1.0)
    return synth.HG38_SIZES if scale == 1.0 else np.maximum((synth.HG38_SIZES * scale).astype(np.int64), 1)


def _gen_file(args):
    config, i, per_file, law, sizes = args
    return synth.gen_intervals(synth.seed_for(config, i), per_file, law, sizes)


def _gen_into(args):
    """worker of the flat build: generate files [i0, i1) straight into the shared arrays (no pickling of results)"""
    config, i0, i1, per_file, law, sizes, paths = args
    arrs = [np.load(p, mmap_mode="r+") for p in paths]
    for i in range(i0, i1):
        iv = synth.gen_intervals(synth.seed_for(config, i), per_file, law, sizes)
        a, b = i * per_file, (i + 1) * per_file
        arrs[0][a:b], arrs[1][a:b], arrs[2][a:b], arrs[3][a:b] = iv.ctg, iv.start, iv.end, iv.value
    for x in arrs:
        x.flush()
    return i1 - i0


FLAT_BUILD_MIN = 20_000_000     # intervals from which the database is built from flat shared arrays (write_igd_flat)


def build_db(name: str, workdir: str, procs: int | None = None) -> str:
    """Create (once) ``<workdir>/<name>.igd`` + ``_index.tsv``; returns the .igd path."""
    cfg = CONFIGS[name]
    os.makedirs(workdir, exist_ok=True)
    stem = os.path.join(workdir, name)
    marker = stem + ".json"
    want = json.dumps({k: cfg.get(k) for k in ("config", "n_files", "per_file", "db_law", "genome_scale")},
                      sort_keys=True)
    if os.path.exists(marker) and os.path.exists(stem + ".igd") and open(marker).read() == want:
        return stem + ".igd"
    sizes = genome_sizes(name)
    procs = procs or min(os.cpu_count() or 1, 32)
    n_files, per_file = cfg["n_files"], cfg["per_file"]
    if n_files * per_file >= FLAT_BUILD_MIN:
        n = n_files * per_file
        paths = [stem + f".gen{k}.npy" for k in range(4)]
        for p, dt in zip(paths, (np.int8, np.int32, np.int32, np.int16)):
            np.lib.format.open_memmap(p, mode="w+", dtype=dt, shape=(n,)).flush()
        step = max(1, n_files // (procs * 8))
        jobs = [(cfg["config"], a, min(a + step, n_files), per_file, cfg["db_law"], sizes, paths)
                for a in range(0, n_files, step)]
        with ProcessPoolExecutor(procs) as ex:
            list(ex.map(_gen_into, jobs))
        arrs = [np.load(p, mmap_mode="r") for p in paths]
        stats = synth.write_igd_flat(stem, arrs[0], arrs[1], arrs[2], arrs[3], per_file)
        del arrs
        for p in paths:
            os.unlink(p)
    else:
        jobs = [(cfg["config"], i, per_file, cfg["db_law"], sizes) for i in range(n_files)]
        if procs > 1 and len(jobs) >= 64:
            with ProcessPoolExecutor(procs) as ex:
                files = list(ex.map(_gen_file, jobs, chunksize=max(1, len(jobs) // (procs * 4))))
        else:
            files = [_gen_file(j) for j in jobs]
        stats = synth.write_igd_direct(stem, files)
    with open(stem + ".stats.json", "w") as f:
        json.dump(stats, f)
    with open(marker, "w") as f:
        f.write(want)
    return stem + ".igd"


def queries(name: str, n: int | None = None) -> synth.Intervals:
    cfg = CONFIGS[name]
    return synth.gen_intervals(synth.seed_for(cfg["config"], synth.QUERY_FILE_INDEX),
                               n or cfg["n_queries"], cfg["q_law"], genome_sizes(name))


def read_igd_header(path: str) -> dict:
    """nbp, gType, nCtg, nTile[], nCnt[] (flattened contig-major) and contig names of an .igd file
    (layout: /root/reference/src/igd_base.c:408-422; read like get_igdinfo, :269-323).  Host-only."""
    with open(path, "rb") as f:
        nbp, gtype, m = np.fromfile(f, np.int32, 3)
        ntile = np.fromfile(f, np.int32, m)
        ncnt = np.fromfile(f, np.int32, int(ntile.sum()))
        names = [f.read(40).split(b"\0")[0].decode(errors="replace") for _ in range(m)]
    tbase = np.concatenate([[0], np.cumsum(ntile.astype(np.int64))])
    return {"nbp": int(nbp), "gtype": int(gtype), "nctg": int(m), "ntile": ntile, "ncnt": ncnt, "tbase": tbase,
            "names": names}


def query_cost(hdr: dict, ctg: np.ndarray, qs: np.ndarray, qe: np.ndarray) -> np.ndarray:
    """Estimated scan bytes per query: 16 + sum over the bins it touches of (8 + 4 * nCnt) -- the same figure
    as the library's igd_b200_query_cost, computed from the header alone."""
    nbp = hdr["nbp"]
    ctg = np.asarray(ctg, np.int64)
    ok = (ctg >= 0) & (ctg < hdr["nctg"])
    c = np.where(ok, ctg, 0)
    n1 = np.asarray(qs, np.int64) // nbp
    n1 = np.where(np.asarray(qs) < 0, -(-np.asarray(qs, np.int64) // nbp), n1)      # C truncating division
    n2 = (np.asarray(qe, np.int64) - 1) // nbp
    n2 = np.where(np.asarray(qe, np.int64) - 1 < 0, -(-(np.asarray(qe, np.int64) - 1) // nbp), n2)
    mT = hdr["ntile"][c].astype(np.int64) - 1
    ok &= (n1 >= 0) & (n1 <= mT)
    n2 = np.maximum(np.minimum(n2, mT), n1)
    # prefix sums of (8 + 4 nCnt) over the flattened tile table make any bin range O(1)
    w = np.concatenate([[0], np.cumsum(8 + 4 * hdr["ncnt"].astype(np.int64))])
    lo = hdr["tbase"][c] + np.where(ok, n1, 0)
    hi = hdr["tbase"][c] + np.where(ok, n2, 0) + 1
    return np.where(ok, 16 + w[hi] - w[lo], 16).astype(np.int64)


def sort_and_shard(hdr: dict, ctg: np.ndarray, qs: np.ndarray, qe: np.ndarray, n_parts: int):
    """Sort queries by (contig id, start) and cut them into ``n_parts`` contiguous ranges balanced by the
    prefix sum of estimated per-query scan bytes.  Host-only; every rank computes the same
    bounds from the same inputs.  Returns sorted arrays + bounds (len n_parts + 1)."""
    order = np.lexsort((qs, ctg))
    ctg, qs, qe = ctg[order], qs[order], qe[order]
    cost = query_cost(hdr, ctg, qs, qe)
    cum = np.cumsum(cost)
    total = int(cum[-1]) if len(cum) else 0
    bounds = [0]
    for p in range(1, n_parts):
        bounds.append(int(np.searchsorted(cum, total * p / n_parts)))
    bounds.append(len(ctg))
    return ctg, qs, qe, bounds

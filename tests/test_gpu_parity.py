"""GPU parity: the CUDA path (through the C-ABI) against the oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

from igd_b200 import synth
from tests.conftest import SMALL_NAMES, SMALL_SIZES, make_queries

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def api():
    from igd_b200 import api
    return api


def _queries(kind):
    if kind == "short":
        return make_queries(101, 20000, ("lognormal", 6.0, 1.0))
    if kind == "fixed1k":
        return make_queries(102, 20000, ("fixed", 1000))
    if kind == "long":
        return make_queries(103, 3000, ("fixed", 150000))
    raise KeyError(kind)


@pytest.mark.parametrize("kind", ["short", "fixed1k", "long"])
@pytest.mark.parametrize("v", [0, 1, 500, 1000])
def test_count_dense(dense_db, O, api, kind, v):
    q = _queries(kind)
    odb = O.OracleDB(dense_db)
    g = api.IGD(dense_db)
    want, want_ret = odb.count(q.ctg, q.start, q.end, v)
    got, got_n = g.count(q.ctg, q.start, q.end, v)
    assert np.array_equal(got, want)
    assert got_n == int(want.sum())
    g.close()


@pytest.mark.parametrize("v", [0, 500])
def test_count_sparse_cfg1(sparse_db, O, api, v):
    law = ("lognormal", 6.5, 1.2)
    q = synth.gen_intervals(synth.seed_for(1, synth.QUERY_FILE_INDEX), 10000, law)
    odb = O.OracleDB(sparse_db)
    g = api.IGD(sparse_db)
    want, _ = odb.count(q.ctg, q.start, q.end, v)
    got, _ = g.count(q.ctg, q.start, q.end, v)
    assert want.sum() > 0
    assert np.array_equal(got, want)
    g.close()


def test_hits_accumulate(dense_db, O, api):
    q = _queries("short")
    odb = O.OracleDB(dense_db)
    g = api.IGD(dense_db)
    seed = np.arange(g.nfiles, dtype=np.int64) * 7 + 3
    want, _ = odb.count(q.ctg, q.start, q.end, 0, hits=seed.copy())
    got, _ = g.count(q.ctg, q.start, q.end, 0, hits=seed.copy())
    got, _ = g.count(q.ctg[:0], q.start[:0], q.end[:0], 0, hits=got)     # empty batch is a no-op
    assert np.array_equal(got, want)
    g.close()


@pytest.mark.parametrize("kind", ["short", "long"])
def test_full_dense(dense_db, O, api, kind):
    q = _queries(kind)
    n = 2000 if kind == "long" else 20000
    odb = O.OracleDB(dense_db)
    g = api.IGD(dense_db)
    per, oq, os_, oe, of = odb.full(q.ctg[:n], q.start[:n], q.end[:n])
    qoff, s, e, f = g.full(q.ctg[:n], q.start[:n], q.end[:n])
    assert np.array_equal(np.diff(qoff), per)
    # exact reference order, not just the same multiset
    assert np.array_equal(s, os_) and np.array_equal(e, oe) and np.array_equal(f, of)
    g.close()


@pytest.mark.parametrize("v", [0, 500])
def test_hitsmap_dense(dense_db, O, api, v):
    odb = O.OracleDB(dense_db)
    g = api.IGD(dense_db)
    assert np.array_equal(g.hitsmap(v), odb.hitsmap(v))
    g.close()


def _write_bed(path, q, names):
    synth.write_bed(path, q, names)
    return path


@pytest.mark.parametrize("kind", ["short", "long"])
def test_seqpare_dense(dense_db, O, api, kind, workdir):
    """-s: sm[] identical to the oracle's doubles (the oracle is pinned bit-exact against the reference)."""
    q = _queries(kind)
    n = 1500 if kind == "long" else 8000
    q = synth.Intervals(q.ctg[:n], q.start[:n], q.end[:n], q.value[:n])
    path = _write_bed(os.path.join(workdir, f"seq_{kind}.bed"), q, SMALL_NAMES)
    odb = O.OracleDB(dense_db)
    g = api.IGD(dense_db)
    want = odb.seqpare(O.Queries(path))
    got = g.seqpare(path)
    assert want.max() > 0
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=0)       # north_star's tolerance
    assert np.array_equal(got, want)                               # and in fact the same doubles
    g.close()


def test_query_cost_c_equals_host_planner(dense_db, api):
    """igd_b200_query_cost (C, from the open handle) == workloads.query_cost (numpy, from the header alone)."""
    from igd_b200 import workloads
    q = _queries("long")
    g = api.IGD(dense_db)
    hdr = workloads.read_igd_header(dense_db)
    assert np.array_equal(g.query_cost(q.ctg, q.start, q.end), workloads.query_cost(hdr, q.ctg, q.start, q.end))
    g.close()


def test_count_device_pointers_and_growth(dense_db, O, api):
    """igd_b200_count_device + _finish on device buffers, batches of growing size (the scratch buffers are
    sized by the first batch; a larger one must be detected, re-run and still give the reference's hits)."""
    import torch
    odb = O.OracleDB(dense_db)
    g = api.IGD(dense_db)
    dev = torch.device("cuda", 0)
    for n, kind in ((50, "short"), (20000, "short"), (3000, "long")):
        q = _queries(kind)
        ctg, qs, qe = (torch.from_numpy(np.ascontiguousarray(x[:n])).to(dev) for x in (q.ctg, q.start, q.end))
        hits = torch.zeros(g.nfiles, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        g.count_device(n, ctg.data_ptr(), qs.data_ptr(), qe.data_ptr(), 0, hits.data_ptr(), 0)
        g.count_device_finish()
        torch.cuda.synchronize()
        want, _ = odb.count(q.ctg[:n], q.start[:n], q.end[:n], 0)
        assert np.array_equal(hits.cpu().numpy(), want), (n, kind)
    g.close()


def test_degenerate_and_extreme_queries(dense_db, O, api):
    """qs == qe, qs > qe, negative starts, INT_MAX ends, unknown contigs, queries beyond the last tile."""
    I = np.iinfo(np.int32)
    ctg = np.array([0, 0, 0, 0, 1, 2, -1, 7, 0, 1, 0, 0], np.int32)
    qs = np.array([5000, 6000, -5, I.min, 0, 199999, 5, 5, 999_999, 16384, 0, 16383], np.int32)
    qe = np.array([5000, 5500, 300, I.max, I.max, 200001, 50, 50, 2_000_000, 16384, 1, 16385], np.int32)
    odb = O.OracleDB(dense_db)
    g = api.IGD(dense_db)
    for v in (0, 1, 700):
        want, _ = odb.count(ctg, qs, qe, v)
        got, _ = g.count(ctg, qs, qe, v)
        assert np.array_equal(got, want), v
    g.close()


def _abnormal_mix(n=30000, frac=0.15, seed=5):
    """A query batch in which `frac` of the queries have qs >= qe or qs < 0 (parse_bed accepts them: only
    en > 0 is required, src/igd_base.c:69) -- the packed layout hands exactly these to the int32 pipeline."""
    q = make_queries(seed, n, ("lognormal", 6.5, 1.3))
    rng = np.random.default_rng(seed)
    qs, qe = q.start.copy(), q.end.copy()
    pick = rng.random(n) < frac
    kind = rng.integers(0, 4, n)
    swap = pick & (kind == 0)
    qs[swap], qe[swap] = q.end[swap], np.maximum(q.start[swap], 1)          # qs > qe
    eq = pick & (kind == 1)
    qe[eq] = np.maximum(qs[eq], 1)                                          # qs == qe (or 0, 1)
    qs[eq] = qe[eq]
    neg = pick & (kind == 2)
    qs[neg] = -rng.integers(1, 40000, int(neg.sum()))                       # negative start
    far = pick & (kind == 3)
    qs[far] = qe[far] + rng.integers(0, 3000, int(far.sum()))               # a record must contain [qe, qs]
    return q.ctg, qs.astype(np.int32), qe.astype(np.int32)


@pytest.mark.parametrize("v", [0, 400])
def test_count_abnormal_mix(dense_db, O, api, v):
    ctg, qs, qe = _abnormal_mix()
    odb = O.OracleDB(dense_db)
    g = api.IGD(dense_db)
    want, _ = odb.count(ctg, qs, qe, v)
    got, n = g.count(ctg, qs, qe, v)
    assert want.sum() > 0 and np.array_equal(got, want) and n == want.sum()
    # the same batch again: the second (int32) pipeline must leave the scratch state clean
    got2, _ = g.count(ctg, qs, qe, v)
    assert np.array_equal(got2, want)
    g.close()


@pytest.mark.parametrize("v", [0, 500])
def test_count_int32_layout(dense_db, O, api, v, monkeypatch):
    """The int32 record layout (what databases with nbp > 16384 or more than 65536 files use) gives the same
    hits[] as the packed 16-bit layout and the oracle."""
    monkeypatch.setenv("IGD_B200_NO_PACKED", "1")
    g = api.IGD(dense_db)
    odb = O.OracleDB(dense_db)
    for ctg, qs, qe in ((lambda q: (q.ctg, q.start, q.end))(_queries("fixed1k")), _abnormal_mix(8000)):
        want, _ = odb.count(ctg, qs, qe, v)
        got, _ = g.count(ctg, qs, qe, v)
        assert np.array_equal(got, want)
    assert g.last_plan()["own_bytes"] > 0
    g.close()


@pytest.mark.parametrize("env", ["IGD_B200_NO_PRESCALE", "IGD_B200_BULK"])
@pytest.mark.parametrize("v", [0, 500])
def test_count_search_kernel_variants(dense_db, O, api, v, env, monkeypatch):
    """The other instantiations of the search kernel: file indices kept unscaled in the count-mode copy (what a database
    of 16 383 .. 65 535 files uses) and the bulk-copy (TMA) staging of the entries -- same hits[] as the default."""
    monkeypatch.setenv(env, "1")
    g = api.IGD(dense_db)
    odb = O.OracleDB(dense_db)
    for ctg, qs, qe in ((lambda q: (q.ctg, q.start, q.end))(_queries("short")), _abnormal_mix(8000)):
        want, _ = odb.count(ctg, qs, qe, v)
        got, _ = g.count(ctg, qs, qe, v)
        assert np.array_equal(got, want)
    g.close()


def test_count_more_files_than_a_shared_histogram_holds(workdir, O, api):
    """50 000 files: the per-CTA histogram (200 KB) leaves no room for the warps' rings, so the search kernel adds
    to hits[] in global memory, and the file indices are too large to be kept pre-scaled."""
    stem = os.path.join(workdir, "manyfiles")
    rng = np.random.default_rng(97)
    n_files, per = 50000, 12
    ctg = rng.choice(3, size=n_files * per, p=SMALL_SIZES / SMALL_SIZES.sum()).astype(np.int32)
    ln = np.clip(np.exp(rng.normal(6.0, 1.2, n_files * per)), 1, 50000).astype(np.int64)
    start = np.floor(rng.random(n_files * per) * np.maximum(SMALL_SIZES[ctg] - ln, 1)).astype(np.int64)
    end = np.minimum(start + ln, SMALL_SIZES[ctg])
    val = rng.integers(0, 1001, size=n_files * per, dtype=np.int32)
    files = [synth.Intervals(ctg[i * per:(i + 1) * per], start[i * per:(i + 1) * per].astype(np.int32),
                             end[i * per:(i + 1) * per].astype(np.int32), val[i * per:(i + 1) * per]) for i in range(n_files)]
    synth.write_igd_direct(stem, files, names=SMALL_NAMES)
    odb = O.OracleDB(stem + ".igd")
    g = api.IGD(stem + ".igd")
    assert g.nfiles == n_files
    q = _queries("short")
    for v in (0, 500):
        want, _ = odb.count(q.ctg, q.start, q.end, v)
        got, _ = g.count(q.ctg, q.start, q.end, v)
        assert want.sum() > 0 and np.array_equal(got, want), v
    g.close()


@pytest.mark.parametrize("b", [9, 12, 15])
def test_count_other_bin_sizes(workdir, O, api, b):
    """igd create -b: nbp = 2^b.  b <= 14 runs on the packed layout, b = 15 on int32."""
    stem = os.path.join(workdir, f"bins{b}")
    files = [synth.gen_intervals(4100 + i, 2500, ("lognormal", 6.5, 1.4), SMALL_SIZES) for i in range(11)]
    synth.write_igd_direct(stem, files, names=SMALL_NAMES, nbp=1 << b)
    odb = O.OracleDB(stem + ".igd")
    g = api.IGD(stem + ".igd")
    assert g.nbp == 1 << b
    ctg, qs, qe = _abnormal_mix(15000, 0.05, seed=b)
    for v in (0, 300):
        want, _ = odb.count(ctg, qs, qe, v)
        got, _ = g.count(ctg, qs, qe, v)
        assert want.sum() > 0 and np.array_equal(got, want), (b, v)
    per, _, s, e, f = odb.full(ctg[:3000], qs[:3000], qe[:3000])
    qoff, gs, ge, gf = g.full(ctg[:3000], qs[:3000], qe[:3000])
    assert np.array_equal(np.diff(qoff), per) and np.array_equal(gs, s) and np.array_equal(ge, e) and np.array_equal(gf, f)
    g.close()


def test_count_many_tiles(workdir, O, api):
    """-b 12 on hg38: 754k tiles, more than one plan CTA per 512 tiles can be resident at once -- the chained
    scan of k_plan must cap its grid (a grid larger than the resident capacity would wait for itself)."""
    stem = os.path.join(workdir, "manytiles")
    law = ("lognormal", 7.0, 1.5)
    files = [synth.gen_intervals(synth.seed_for(1, 50 + i), 20000, law) for i in range(10)]
    synth.write_igd_direct(stem, files, nbp=4096)
    q = synth.gen_intervals(synth.seed_for(1, 999), 200000, ("lognormal", 8.0, 1.5))
    odb, g = O.OracleDB(stem + ".igd"), api.IGD(stem + ".igd")
    assert g.ntiles > 700000
    for v in (0, 500):
        want, _ = odb.count(q.ctg, q.start, q.end, v)
        got, _ = g.count(q.ctg, q.start, q.end, v)
        assert want.sum() > 0 and np.array_equal(got, want)
    g.close()


@pytest.mark.parametrize("n_files", [50_000, 70_000])
def test_count_many_files(workdir, O, api, n_files):
    """More files than a per-CTA shared-memory histogram holds: 50 000 runs the packed kernel with global
    atomics, 70 000 (> 65 536: file indices no longer fit 16 bits) the int32 kernel with global atomics."""
    stem = os.path.join(workdir, f"manyfiles{n_files}")
    rng = np.random.default_rng(n_files)
    files = []
    for i in range(n_files):
        n = 12
        ctg = rng.integers(0, 3, n).astype(np.int32)
        start = (rng.random(n) * (SMALL_SIZES[ctg] - 5000)).astype(np.int32)
        files.append(synth.Intervals(ctg, start, (start + rng.integers(1, 4000, n)).astype(np.int32),
                                     rng.integers(0, 1000, n).astype(np.int32)))
    synth.write_igd_direct(stem, files, names=SMALL_NAMES)
    q = make_queries(9, 3000, ("lognormal", 7.0, 1.2))
    odb, g = O.OracleDB(stem + ".igd"), api.IGD(stem + ".igd")
    assert g.nfiles == n_files
    for v in (0, 600):
        want, _ = odb.count(q.ctg, q.start, q.end, v)
        got, _ = g.count(q.ctg, q.start, q.end, v)
        assert want.sum() > 0 and np.array_equal(got, want)
    g.close()


def test_count_v_thresholds_and_negative_values(workdir, O, api):
    """-v on the packed path keeps value clamped to [0, 32767] in 16 bits: thresholds inside 1 .. 32767 take it,
    thresholds outside (and get_overlaps_v's v <= 0 through the reference-named entry point) take the int32
    arrays; values below 0 and above 32767 are in the database."""
    stem = os.path.join(workdir, "vals")
    rng = np.random.default_rng(77)
    files = []
    for i in range(12):
        iv = synth.gen_intervals(6000 + i, 3000, ("lognormal", 6.5, 1.2), SMALL_SIZES)
        val = rng.choice(np.array([-500, -1, 0, 1, 499, 500, 501, 32766, 32767, 32768, 40000, 2_000_000_000], np.int32),
                         len(iv))
        files.append(synth.Intervals(iv.ctg, iv.start, iv.end, val.astype(np.int32)))
    synth.write_igd_direct(stem, files, names=SMALL_NAMES)
    q = make_queries(8, 8000, ("lognormal", 7.0, 1.3))
    odb, g = O.OracleDB(stem + ".igd"), api.IGD(stem + ".igd")
    for v in (1, 2, 500, 501, 32766, 32767, 32768, 40001, 2_000_000_000):
        want, _ = odb.count(q.ctg, q.start, q.end, v)
        got, n = g.count(q.ctg, q.start, q.end, v)
        assert np.array_equal(got, want), v
        assert n == want.sum()
    assert odb.count(q.ctg, q.start, q.end, 32767)[0].sum() > 0
    g.close()


def test_full_device_matches_full(dense_db, api):
    """igd_b200_full_device (queries read from HBM, lists left in HBM) = igd_b200_full, viewed through torch
    without a copy; the handle keeps the buffers until the next call"""
    import torch
    q = make_queries(31, 6000, ("lognormal", 7.5, 1.5))
    with api.IGD(dense_db) as g:
        qoff, s, e, f = g.full(q.ctg, q.start, q.end)
        dev = torch.device("cuda", 0)
        d = [torch.from_numpy(np.ascontiguousarray(x, np.int32)).to(dev) for x in (q.ctg, q.start, q.end)]
        n = len(q.ctg)
        total, p_off, p_lists = g.full_device(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr())
        assert total == len(s) > 0
        t_off = torch.as_tensor(api.device_view(p_off, n + 1, "<i8"), device=dev)
        t_l = torch.as_tensor(api.device_view(p_lists, 3 * total, "<i4"), device=dev).view(3, total)
        assert np.array_equal(t_off.cpu().numpy(), qoff)
        got = t_l.cpu().numpy()
        assert np.array_equal(got[0], s) and np.array_equal(got[1], e) and np.array_equal(got[2], f)
        # an empty batch, and a batch without overlaps
        total, p_off, p_lists = g.full_device(0, 0, 0, 0)
        assert total == 0
        far = torch.full((4,), -1, dtype=torch.int32, device=dev)          # unknown contigs
        total, p_off, _ = g.full_device(4, far.data_ptr(), d[1].data_ptr(), d[2].data_ptr())
        assert total == 0
        assert torch.as_tensor(api.device_view(p_off, 5, "<i8"), device=dev).cpu().tolist() == [0] * 5


def test_list_results_survive_a_file_search_on_a_small_database(workdir, O, api):
    """-f, then search() of a query FILE (which grows the handle's pinned staging buffer on a database smaller than
    a file chunk), then -f again: the pinned result arena of -f is a different buffer and must stay intact (a stray
    cudaFreeHost of the arena in the staging-buffer growth path was found and removed in round 1)"""
    stem = os.path.join(workdir, "tiny_arena")
    files = [synth.gen_intervals(88000 + i, 300, ("lognormal", 6.0, 1.2), SMALL_SIZES) for i in range(10)]
    synth.write_igd_direct(stem, files, names=SMALL_NAMES)
    q = make_queries(12, 4000, ("lognormal", 7.0, 1.5))
    qbed = os.path.join(workdir, "tiny_arena_q.bed")
    synth.write_bed(qbed, q, SMALL_NAMES)
    odb = O.OracleDB(stem + ".igd")
    per, _, s, e, f = odb.full(q.ctg, q.start, q.end)
    want, _ = odb.count(q.ctg, q.start, q.end)
    with api.IGD(stem + ".igd") as g:
        for _ in range(3):
            qoff, gs, ge, gf = g.full(q.ctg, q.start, q.end, copy=False)      # views of the arena
            hits, _ = g.search_n(qbed)
            assert np.array_equal(hits, want)
            assert np.array_equal(np.diff(qoff), per) and np.array_equal(gs, s) and np.array_equal(ge, e) and np.array_equal(gf, f)
            off, rec = g.seq_overlaps(q.ctg[:50], q.start[:50], q.end[:50])
            assert off[-1] == len(rec)


def test_open_rejects_file_index_outside_tsv(workdir, api):
    """a record whose idx is outside [0, nFiles) would make the reference write outside hits[]; the open fails
    (the check runs on the device while the payload streams in)"""
    stem = os.path.join(workdir, "bad_idx")
    files = [synth.gen_intervals(91000 + i, 200, ("lognormal", 6.0, 1.2), SMALL_SIZES) for i in range(10)]
    synth.write_igd_direct(stem, files, names=SMALL_NAMES)
    raw = np.fromfile(stem + ".igd", np.int32)
    m = raw[2]
    hdr = 3 + m + int(raw[3:3 + m].sum()) + 10 * m
    for bad in (10, -1):
        r = raw.copy()
        r[hdr + 4 * 77] = bad
        p = os.path.join(workdir, f"bad_idx_{bad & 0xff}.igd")
        r.tofile(p)
        import shutil
        shutil.copy(stem + "_index.tsv", p[:-4] + "_index.tsv")
        with pytest.raises(api.IGDError, match="file index outside"):
            api.IGD(p)
    api.IGD(stem + ".igd").close()

"""The handle-flavoured binding surface of the reference's R and Cython packages over the C-ABI:
igd_py.create / open / search_1 / search_n / get_nFiles (src_py/igd_py.pyx:21-44), R's search_nr with its 32-bit
counters and get_binData (IGDr/src/igd_search.c:340-393)."""
import os

import numpy as np
import pytest

from igd_b200 import api, synth
from oracle import oracle as O
from tests.conftest import SMALL_SIZES
from tests.golden_util import GoldenCase

NAMES = ["chr1", "chr2", "chr3"]


@pytest.fixture(scope="module")
def built(workdir):
    """a database built through the Python binding (igd_py.create), -b 13"""
    bed = os.path.join(workdir, "bind_bed")
    os.makedirs(bed)
    files = [synth.gen_intervals(61000 + i, 1500, ("lognormal", 6.5, 1.3), SMALL_SIZES) for i in range(10)]
    for i, iv in enumerate(files):
        synth.write_bed(os.path.join(bed, f"f{i:05d}.bed"), iv, NAMES)
    out = os.path.join(workdir, "bind_out")
    assert api.create(bed, out, "db", 8192) is True
    assert api.create(bed, out, "db", 8192) is False        # exists: nothing is written
    return os.path.join(out, "db.igd"), bed


def test_python_create_matches_reference_create(built, workdir):
    igd, bed = built
    db = O.OracleDB(igd)
    assert db.nbp == 8192 and db.nfiles == 10 and db.file_nr.tolist() == [1500] * 10
    if O.have_ref():
        from tests.test_create import canonical_igd
        ref = O.ref_create(bed, os.path.join(workdir, "bind_ref"), "db", b=13)
        assert canonical_igd(ref) == canonical_igd(igd)
        assert open(ref[:-4] + "_index.tsv").read() == open(igd[:-4] + "_index.tsv").read()


@pytest.mark.gpu
def test_search_nr_32bit_and_search_1(built):
    igd, _ = built
    db = O.OracleDB(igd)
    q = synth.gen_intervals(5, 3000, ("lognormal", 7.0, 1.5), SMALL_SIZES)
    chrm = [NAMES[c] for c in q.ctg] + ["chrNope"]
    qs, qe = np.append(q.start, 5), np.append(q.end, 500)
    want, _ = db.count(db.resolve_names(chrm), qs, qe)
    with api.IGD(igd) as g:
        assert g.get_nFiles() == 10
        hits = g.search_nr(chrm, qs, qe)
        assert hits.dtype == np.int32 and np.array_equal(hits, want)
        h1, _ = g.search_1(chrm[0], int(qs[0]), int(qe[0]))
        assert np.array_equal(h1, db.count(db.resolve_names(chrm[:1]), qs[:1], qe[:1])[0])
        # counters are incremented, and wrap like the reference's 32-bit ints
        L = api.load_library()
        import ctypes as C
        pre = np.full(10, 2**31 - 1, np.int32)
        arr = (C.c_char_p * 1)(chrm[0].encode())
        L.get_overlaps_n32(g._h, 1, arr, qs[:1].astype(np.int32), qe[:1].astype(np.int32), pre)
        assert np.array_equal(pre, (np.full(10, 2**31 - 1, np.int64) + h1).astype(np.uint32).astype(np.int32))


@pytest.mark.gpu
def test_bin_data_is_the_file(built):
    """get_binData: a tile read back from the HBM image = the tile's bytes in the .igd file"""
    igd, _ = built
    raw = np.fromfile(igd, np.int32)
    nbp, gtype, m = raw[:3]
    ntile = raw[3:3 + m]
    ncnt = raw[3 + m:3 + m + ntile.sum()]
    rec = raw[3 + m + ntile.sum() + 10 * m:].reshape(-1, 4)
    off = np.concatenate([[0], np.cumsum(ncnt)])
    with api.IGD(igd) as g:
        t = 0
        for c in range(m):
            for b in range(ntile[c]):
                if (t % 7 == 0 or ncnt[t] == ncnt.max()):
                    idx, s, e, v = g.bin_data(c, b)
                    assert len(idx) == ncnt[t]
                    assert np.array_equal(np.stack([idx, s, e, v], 1), rec[off[t]:off[t + 1]])
                t += 1
        assert len(g.bin_data(0, int(ntile[0]))[0]) == 0 and len(g.bin_data(m, 0)[0]) == 0


@pytest.mark.gpu
def test_create_then_search_end_to_end(workdir):
    """`igd create` of this repo, then `igd search` of this repo, against the reference's committed stdout: the
    golden `dense` case rebuilt from BED files extracted from its own database"""
    import subprocess
    case = GoldenCase("dense", workdir)
    raw = np.fromfile(case.igd, np.int32)
    m = raw[2]
    ntile = raw[3:3 + m]
    T = int(ntile.sum())
    names = [bytes(raw[3 + m + T + 10 * c:3 + m + T + 10 * (c + 1)].tobytes()).split(b"\0")[0].decode() for c in range(m)]
    ncnt = raw[3 + m:3 + m + T]
    rec = raw[3 + m + T + 10 * m:].reshape(-1, 4)
    ctg_of = np.repeat(np.repeat(np.arange(m), ntile), ncnt)
    bin_of = np.repeat(np.concatenate([np.arange(k) for k in ntile]), ncnt)
    first = (rec[:, 1] // 16384) == bin_of                  # each interval once: the copy in its first tile
    bed = os.path.join(workdir, "e2e_bed")
    os.makedirs(bed)
    for f in range(10):
        sel = first & (rec[:, 0] == f)
        with open(os.path.join(bed, f"f{f:05d}.bed"), "w") as fh:
            for c, r in zip(ctg_of[sel], rec[sel]):
                fh.write(f"{names[c]}\t{r[1]}\t{r[2]}\t.\t{r[3]}\n")
    out = os.path.join(workdir, "e2e_out") + "/"
    r = subprocess.run([api.CLI_PATH, "create", bed, out, "db"], capture_output=True)
    assert r.returncode == 0
    for key in ("q", "q_v500"):
        r = subprocess.run([api.CLI_PATH, "search", out + "db.igd"] + case.cli_args(key), capture_output=True)
        assert case.normalise(r.stdout.decode()) == case.stdout(key), key

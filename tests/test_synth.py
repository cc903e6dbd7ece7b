"""The direct .igd writer (igd_b200/synth.py) against the reference's own `igd create` (oracle/_ref/igd, compiled
from /root/reference by oracle/Makefile).  Needs the reference build, i.e. runs in the build container only; the
GPU box skips it (the fixtures under tests/golden/ were made by `igd create` itself)."""
import os

import numpy as np
import pytest

from igd_b200 import synth, workloads
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="needs oracle/_ref (the reference build)")


def _tiles(path):
    hdr = workloads.read_igd_header(path)
    off = 12 + 44 * hdr["nctg"] + 4 * int(hdr["ntile"].sum())
    rec = np.fromfile(path, np.int32, offset=off).reshape(-1, 4)
    return hdr, rec


@pytest.mark.parametrize("law", [("lognormal", 6.5, 1.3), ("fixed", 40000)])
def test_direct_writer_equals_igd_create(tmp_path, law):
    sizes = np.array([400_000, 250_000, 120_000], dtype=np.int64)
    names = ["chr1", "chr2", "chr3"]
    bed = tmp_path / "bed"
    bed.mkdir()
    files = []
    for i in range(10):
        iv = synth.gen_intervals(900 + i, 1500, law, sizes)
        synth.write_bed(str(bed / f"f{i:05d}.bed"), iv, names)
        files.append(iv)
    ref = O.ref_create(str(bed), str(tmp_path / "ref"), "db")
    synth.write_igd_direct(str(tmp_path / "mine"), files, names=names)
    mine = str(tmp_path / "mine.igd")
    h1, r1 = _tiles(ref)
    h2, r2 = _tiles(mine)
    assert h1["nbp"] == h2["nbp"] and h1["names"] == h2["names"]
    assert np.array_equal(h1["ntile"], h2["ntile"]) and np.array_equal(h1["ncnt"], h2["ncnt"])
    assert os.path.getsize(ref) == os.path.getsize(mine)
    # identical tile by tile up to the order of equal starts (the reference's radix sort is unstable)
    a = 0
    for n in h1["ncnt"]:
        t1, t2 = r1[a:a + n], r2[a:a + n]
        assert np.array_equal(t1[:, 1], t2[:, 1])                       # starts, sorted
        k1 = np.lexsort((t1[:, 3], t1[:, 0], t1[:, 2], t1[:, 1]))
        k2 = np.lexsort((t2[:, 3], t2[:, 0], t2[:, 2], t2[:, 1]))
        assert np.array_equal(t1[k1], t2[k2])
        a += n
    # and the tsv: index, name, number of regions
    l1 = [x.split("\t")[:3] for x in open(ref[:-4] + "_index.tsv").read().splitlines()]
    l2 = [x.split("\t")[:3] for x in open(mine[:-4] + "_index.tsv").read().splitlines()]
    assert l1 == l2

"""The oracle restatement against the UNMODIFIED reference compiled in place (oracle/_ref/libigdref.so), on
fresh seeded inputs beyond the committed fixtures.  Build container only (needs /root/reference at build time)."""
import os

import numpy as np
import pytest

from igd_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="needs oracle/_ref (the reference build)")

SIZES = np.array([300_000, 150_000], dtype=np.int64)
NAMES = ["chr1", "chr2"]


@pytest.fixture(scope="module")
def db(tmp_path_factory):
    d = tmp_path_factory.mktemp("ovr")
    bed = d / "bed"
    bed.mkdir()
    for i in range(12):
        synth.write_bed(str(bed / f"f{i:05d}.bed"), synth.gen_intervals(5100 + i, 1200, ("lognormal", 6.8, 1.4), SIZES),
                        NAMES)
    return O.ref_create(str(bed), str(d / "o"), "db"), d


@pytest.mark.parametrize("law", [("lognormal", 6.0, 1.0), ("fixed", 1000), ("fixed", 60000)])
def test_counts_v_seqpare_hitsmap(db, law):
    igd, d = db
    q = synth.gen_intervals(sum(map(ord, law[0])) + int(law[-1]) % 977, 1500, law, SIZES)
    qf = str(d / f"q_{law[0]}_{law[-1]}.bed")
    synth.write_bed(qf, q, NAMES)
    ref = O.RefLib(igd)
    odb = O.OracleDB(igd)
    oq = O.Queries(qf)
    ctg = odb.resolve(oq)
    for v in (0, 1, 500):
        rh, rr = ref.count_file(qf, v)
        oh, orr = odb.count(ctg, oq.qs, oq.qe, v)
        assert np.array_equal(rh, oh) and rr == orr, v
    assert np.array_equal(ref.seqpare_file(qf), odb.seqpare(oq))
    ref.close()


def test_hitsmap(db):
    igd, _ = db
    ref = O.RefLib(igd)
    odb = O.OracleDB(igd)
    for v in (0, 300):
        assert np.array_equal(ref.hitsmap(v), odb.hitsmap(v))
    ref.close()

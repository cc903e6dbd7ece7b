"""Query-file ingest against tests/golden/nasty (made by tests/golden/make_nasty.py with the unmodified reference):
a BED file of every kind of line parse_bed has an opinion about -- signs, blanks, junk after the digits, long ->
int32 truncation and saturation, CRLF, NUL bytes, 39/40-char contigs, unknown contigs, a 70 kB line, no final
newline -- searched on the `dense` golden database."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.golden_util import GOLDEN, GoldenCase


@pytest.fixture(scope="module")
def nasty(workdir):
    case = GoldenCase("dense", workdir)
    exp = json.load(open(os.path.join(GOLDEN, "nasty", "expected.json")))
    gz = os.path.join(GOLDEN, "nasty", "q.bed.gz")
    plain = os.path.join(workdir, "nasty.bed")
    with gzip.open(gz, "rb") as g, open(plain, "wb") as f:
        f.write(g.read())
    return case, exp, gz, plain


def _key(v):
    return ("hits", "ret") if v == 0 else (f"hits_v{v}", f"ret_v{v}")


@pytest.mark.parametrize("v", [0, 1, 500])
def test_oracle_ingest_matches_reference(nasty, v):
    case, exp, gz, plain = nasty
    db = O.OracleDB(case.igd)
    for path in (plain, gz):
        hits, ret = db.count_file(path, v)
        assert np.array_equal(hits, np.array(exp[_key(v)[0]], np.int64))
        assert ret == exp[_key(v)[1]]
    db.close()


def test_host_reader_matches_oracle_reader(nasty):
    """igd_b200_read_bed (what -f and -s use) accepts the same lines with the same values."""
    from igd_b200 import api
    _, _, gz, plain = nasty
    for path in (plain, gz):
        mine, ref = api.QuerySet(path), O.Queries(path)
        assert mine.n == ref.n and mine.names == ref.names
        assert np.array_equal(mine.qs, ref.qs) and np.array_equal(mine.qe, ref.qe)
        assert np.array_equal(mine.name_id, ref.name_id)


@pytest.mark.gpu
@pytest.mark.parametrize("v", [0, 1, 500])
def test_device_ingest_matches_reference(nasty, v):
    """search() parses the text ON THE GPU (k_parse_bed); hits[] must be the reference's."""
    from igd_b200 import api
    case, exp, gz, plain = nasty
    g = api.IGD(case.igd)
    want = np.array(exp[_key(v)[0]], np.int64)
    for path in (plain, gz):
        hits, n = g.search_n(path, v=v)
        assert np.array_equal(hits, want), path
        assert n == want.sum()
    # the reference-named entry points take the same route
    g.use()
    L = api.load_library()
    hits = np.zeros(g.nfiles, np.int64)
    ret = L.getOverlaps_v(os.fsencode(plain), hits, v) if v > 0 else L.getOverlaps(os.fsencode(plain), hits)
    assert np.array_equal(hits, want) and ret == exp[_key(v)[1]]
    g.close()


@pytest.mark.gpu
def test_device_ingest_many_chunks(nasty, workdir):
    """A file larger than the 16 MiB staging chunk: lines straddle chunk boundaries; compared with the array API
    on the host-parsed queries of the same file."""
    from igd_b200 import api
    case, _, _, plain = nasty
    text = open(plain, "rb").read()
    if not text.endswith(b"\n"):
        text += b"\n"
    big = os.path.join(workdir, "nasty_big.bed")
    with open(big, "wb") as f:
        for _ in range(int(40e6 // len(text)) + 1):
            f.write(text)
    g = api.IGD(case.igd)
    q = api.QuerySet(big)
    want, _ = g.count(g.resolve(q), q.qs, q.qe, 0)
    hits, _ = g.search_n(big)
    assert want.sum() > 0 and np.array_equal(hits, want)
    g.close()

"""The arithmetic identity behind the count-mode copy (igd_dev_impl.cuh: pk_rel, pk_rel_end;
igd_dev_count.cu: k_count_rank), restated in numpy and checked against the reference's int32 predicate on random
tiles.  Per record r of a tile, with F = the tile's first-bin segments and L = its later-bin segments:

  #{F: start_r < qe and end_r > qs} + [start_r >= bd] #{L: start_r < qe}  =  H(start'_r) - K(end'_r)
  H(s) = #{F u L: qe' > s} for s >= 1, H(0) = |F|;     K(e) = #{F: qs' >= e}

evaluated as bit 15 of (x - key) with START key S = start' + 1, END key E = end', x = QE = qe' + 0x8000 resp.
x = QS + 0x7fff = qs' + 0x8000, x' = clamp(x - bd + 1, 0, nbp + 1); no 16-bit half may borrow; a segment past the
LAST key of a sorted group is past all of them (base count), one not past the FIRST is past none.
This pins the MATH; that the kernel implements it is what the -m gpu parity tests check."""
import numpy as np
import pytest


def pk_rel(x, bd, nbp):
    return np.where(x < bd, 0, np.minimum(x - bd, nbp) + 1)


def pk_rel_end(x, bd, nbp):
    return np.where(x <= bd, 1, np.minimum(x - bd, nbp) + 1)


@pytest.mark.parametrize("nbp", [512, 4096, 16384])
def test_rank_form_equals_int32_predicate(nbp):
    rng = np.random.default_rng(nbp)
    for _ in range(150):
        j = int(rng.integers(0, 60))
        bd = j * nbp
        n, m = 300, 200
        rs = np.sort(rng.integers(max(0, bd - 3 * nbp), bd + nbp, n))       # start < bd + nbp, start order
        re = np.maximum(rs + 1 + rng.integers(0, 3 * nbp, n), bd + 1)        # end > bd, end > start
        first = rng.random(m) < 0.5 if bd > 0 else np.ones(m, bool)
        qs = np.where(first, rng.integers(bd, bd + nbp, m), rng.integers(0, max(bd, 1), m))
        qe = np.where(first, qs + 1 + rng.integers(0, 3 * nbp, m), bd + 1 + rng.integers(0, 3 * nbp, m))
        S = (pk_rel(rs, bd, nbp) + 1).astype(np.int64)
        E = pk_rel_end(re, bd, nbp).astype(np.int64)
        QS = (pk_rel(qs, bd, nbp) + 1).astype(np.int64)
        QE = (pk_rel_end(qe, bd, nbp) + 0x8000).astype(np.int64)
        assert S.max() <= 0x8000 and E.max() < 0x7fff and QS.max() <= 0x8000 and QE.max() < 0x10000
        assert (QS[~first] == 1).all()
        xs, xe = QE, QS + 0x7fff
        # keys as the kernel sees them, pads included
        Sp, Ep = np.append(S, 0x8000), np.append(E, 0x7fff)
        A = xs[:, None] - Sp[None, :]
        B = xe[:, None] - Ep[None, :]
        KL = np.where(Sp <= 1, Sp + 0x7fff, Sp)                             # S == 1 halves read 0x8000 for L
        AL = xs[:, None] - KL[None, :]
        for d in (A, B, AL):
            assert d.min() >= 0 and d.max() < 0x10000                        # a half never borrows or carries
        fa, fb, fl = (A >> 15 & 1), (B >> 15 & 1), (AL >> 15 & 1)
        assert not fa[:, -1].any() and not fb[:, -1].any() and not fl[:, -1].any()      # pads count nothing
        H = fa[first].sum(0) + fl[~first].sum(0)
        K = fb[first].sum(0)
        got = (H - K)[:-1]
        hit = (rs[None, :] < qe[:, None]) & (re[None, :] > qs[:, None])
        want = hit[first].sum(0) + (hit[~first] & (rs[None, :] >= bd)).sum(0)
        assert np.array_equal(got, want)
        # the window rule on a sorted group of keys
        for keys, x in ((np.sort(S), xs), (np.sort(E), xe)):
            g = keys[40:296]
            lo_t, hi_t = g[0] + 0x8000, g[-1] + 0x8000
            f = (x[:, None] - g[None, :]) >> 15 & 1
            assert f[x >= hi_t].all() and not f[x < lo_t].any()

"""The arithmetic identity behind the packed count-mode layout (igd_dev_impl.cuh: pk_rel, pk_rel_end; igd_dev_count.cu:
k_count_packed), restated in numpy and checked against the reference's int32 predicate on random tiles:

  start_r < qe  and  end_r > qs   <=>   bit 15 of (QE - S)  and  bit 15 of (E - QS)
  later bins additionally start_r >= bd   <=>   bit 15 of (S + 0x7fff - QS)   (QS = 1 there)

with S = start' + 1, E = end' + 0x8000, QS = qs' + 1, QE = qe' + 0x8000, x' = clamp(x - bd + 1, 0, nbp + 1); no
16-bit half may borrow, and the group prune (QE >= min S + 0x8000, QS <= max E - 0x8000) must never drop a hit.
This pins the MATH; that the kernel implements it is what the -m gpu parity tests check."""
import numpy as np
import pytest


def pk_rel(x, bd, nbp):
    return np.where(x < bd, 0, np.minimum(x - bd, nbp) + 1)


def pk_rel_end(x, bd, nbp):
    return np.where(x <= bd, 1, np.minimum(x - bd, nbp) + 1)


@pytest.mark.parametrize("nbp", [512, 4096, 16384])
def test_packed_compare_equals_int32_predicate(nbp):
    rng = np.random.default_rng(nbp)
    for _ in range(150):
        j = int(rng.integers(0, 60))
        bd = j * nbp
        n, m = 300, 200
        rs = rng.integers(max(0, bd - 3 * nbp), bd + nbp, n)                 # start < bd + nbp
        re = np.maximum(rs + 1 + rng.integers(0, 3 * nbp, n), bd + 1)        # end > bd, end > start
        first = rng.random(m) < 0.5 if bd > 0 else np.ones(m, bool)
        qs = np.where(first, rng.integers(bd, bd + nbp, m), rng.integers(0, max(bd, 1), m))
        qe = np.where(first, qs + 1 + rng.integers(0, 3 * nbp, m), bd + 1 + rng.integers(0, 3 * nbp, m))
        S = (pk_rel(rs, bd, nbp) + 1).astype(np.int64)
        E = (pk_rel_end(re, bd, nbp) + 0x8000).astype(np.int64)
        QS = (pk_rel(qs, bd, nbp) + 1).astype(np.int64)
        QE = (pk_rel_end(qe, bd, nbp) + 0x8000).astype(np.int64)
        assert S.max() <= 0x8000 and QS.max() <= 0x8000 and E.max() < 0x10000 and QE.max() < 0x10000
        A = QE[:, None] - S[None, :]
        B = E[None, :] - QS[:, None]
        X = (S + 0x7fff)[None, :] - QS[:, None]
        for d in (A, B, X):
            assert d.min() >= 0 and d.max() < 0x10000                        # a half never borrows or carries
        hit_first = ((A & B) >> 15 & 1).astype(bool)
        hit_later = ((A & X) >> 15 & 1).astype(bool)
        want = (rs[None, :] < qe[:, None]) & (re[None, :] > qs[:, None])
        want_later = want & (rs[None, :] >= bd)
        assert (QS[~first] == 1).all()
        assert np.array_equal(np.where(first[:, None], hit_first, hit_later),
                              np.where(first[:, None], want, want_later))
        keep = (QE >= S.min() + 0x8000) & (QS <= E.max() - 0x8000)
        assert not hit_first[~keep].any() and not hit_later[~keep].any()
        # pads and filtered records (S = 0x8000) match nothing
        assert not (((QE - 0x8000) >> 15) & 1).any()

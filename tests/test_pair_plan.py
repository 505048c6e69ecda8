"""CPU: the host-side plan of the fused pass's pair pipeline (whole-tile waves + the tail split) through the C ABI test hook
`mil_pair_plan_item` -- no device work.  For many bag sizes: every (tile, pipeline stage) is covered exactly once, a split tile has exactly
one owner and one helper with adjacent stage ranges, the helper's item is its FIRST and the owner's its LAST (the ordering the kernel's
barrier phases and the exchange latency rely on), partial indices are distinct and fit the workspace's exchange area."""
import ctypes

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L():
    import mhimk  # noqa: F401
    from mhimk import _lib
    return _lib.lib()


def plan(L, N, D=1024, prec=0):                      # MIL_PREC_BF16X3 = 0, FP16 = 1, BF16 = 2, FP16X3 = 3
    out = (ctypes.c_int64 * 9)()
    L.mil_pair_plan_item(N, D, prec, 0, -1, out)
    pairs, S, full, rem = int(out[5]), int(out[6]), int(out[7]), int(out[8])
    items = []
    for p in range(pairs):
        n = L.mil_pair_plan_item(N, D, prec, p, -1, out)
        row = []
        for i in range(n):
            assert L.mil_pair_plan_item(N, D, prec, p, i, out) == n
            row.append(tuple(int(out[j]) for j in range(5)))
        items.append(row)
    return pairs, S, full, rem, items


SIZES = [1, 100, 128, 129, 1000, 19 * 128, 37 * 128, 37 * 128 + 1, 38 * 128, 74 * 128, 74 * 128 + 1, 10000, 75 * 128 + 5, 111 * 128, 112 * 128,
         148 * 128, 25000, 50000, 200000]


@pytest.mark.parametrize("D,prec", [(1024, 0), (1536, 3), (96, 0), (32, 0), (1024, 1), (96, 2)])
def test_plan_covers_every_stage_once(L, D, prec):
    ksub = 2 if (prec in (1, 2) and D % 64 == 0) else 1       # 64-wide stages in the single-product modes when D allows
    kst = D // 32 // ksub
    for N in SIZES:
        pairs, S, full, rem, items = plan(L, N, D, prec)
        T = (N + 127) // 128
        assert 1 <= pairs <= 74 and S in (1, 2)
        cover = np.zeros((T, kst), dtype=np.int32)
        owners, helpers, pidx_seen = {}, {}, set()
        for p, row in enumerate(items):
            assert len(row) >= 1, (N, p)
            for i, (tile, kb, ke, kind, pidx) in enumerate(row):
                assert 0 <= tile < T and 0 <= kb < ke <= kst
                cover[tile, kb:ke] += 1
                if kind == 0:
                    assert (kb, ke) == (0, kst)
                elif kind == 1:
                    assert i == len(row) - 1 and tile not in owners          # the owner's split item is its last one
                    owners[tile] = (p, kb, ke, pidx)
                else:
                    assert kind == 2 and i == 0 and tile not in helpers      # the helper's is its first
                    helpers[tile] = (p, kb, ke, pidx)
                    assert pidx not in pidx_seen and 0 <= pidx < 74
                    pidx_seen.add(pidx)
        assert (cover == 1).all(), (N, D)
        assert owners.keys() == helpers.keys()
        for t, (po, kbo, keo, pio) in owners.items():
            ph, kbh, keh, pih = helpers[t]
            assert ph == po + 1 and kbo == 0 and keo == kbh and keh == kst and pio == pih and min(keo - kbo, keh - kbh) >= 4
        if S == 1:
            assert not owners
        else:
            assert len(owners) == rem and 2 * rem <= 74 and kst >= 8


def test_plan_of_the_headline_and_of_a_one_wave_bag(L):
    pairs, S, full, rem, items = plan(L, 50000)
    assert (pairs, S, full, rem) == (74, 2, 5, 21)                  # 391 tiles: 5 waves + 21 tiles shared by 42 pairs
    assert sorted(len(r) for r in items) == [5] * 32 + [6] * 42
    pairs, S, full, rem, items = plan(L, 1000)
    assert (pairs, S, full, rem) == (16, 2, 0, 8) and all(len(r) == 1 for r in items)
    pairs, S, full, rem, items = plan(L, 38 * 128)                 # more than half of the pairs: no split
    assert (pairs, S) == (38, 1)

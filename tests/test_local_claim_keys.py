"""CPU checks of two bit-level arguments the CUDA kernels rely on (csrc/pd_dense.cuh); no GPU, no library.

1. k_claim_punish settles reproduction claims on tile-local cells with an atomic max on
   key = (priority & ~7) | j   (j = 7 - claimed direction, unique among the rivals of a cell),
   which orders like the exact (priority, j) of claim_word unless two rivals agree in the top 29 bits of
   their priorities.  The kernel notices that from the atomic's RETURN value -- `(old ^ key) >> 3 == 0` -- and then
   sends the whole tile through the claim plane.  Claim: whenever the key's winner differs from the exact winner,
   some poster has raised the flag, whatever the posting order.
2. k_play_travel builds a cell's play word branch-free from the flag / strategy plane words; for an occupied cell
   it must equal the spelled-out layout, and an empty cell must have bit 31 clear and compare below every
   threshold an occupied cell uses.
"""
import itertools
import random

import numpy as np


def _post_all(order, prio, j, tie_shift=3):
    """The shared-memory table entry of ONE cell: returns (final max key, flag raised)."""
    cur, flag = 0, False
    for k in order:
        key = (prio[k] & ~7) | j[k]
        old = cur                      # atomicMax returns the previous value
        cur = max(cur, key)
        if ((old ^ key) >> tie_shift) == 0:
            flag = True
    return cur, flag


def test_tie_flag_is_complete_exhaustive_small():
    """All priorities of 5 bits (top part = 2 bits), 2 and 3 rivals, every posting order."""
    for n in (2, 3):
        for js in itertools.combinations(range(8), n):
            for prio in itertools.product(range(32), repeat=n):
                exact = max(range(n), key=lambda k: (prio[k], js[k]))
                for order in itertools.permutations(range(n)):
                    cur, flag = _post_all(order, prio, js)
                    by_key = max(range(n), key=lambda k: (prio[k] & ~7) | js[k])
                    assert cur == ((prio[by_key] & ~7) | js[by_key])
                    assert flag or by_key == exact, (prio, js, order)
        if n == 3:
            break


def test_tie_flag_is_complete_random_32bit():
    rng = random.Random(7)
    for _ in range(20000):
        n = rng.randint(1, 8)
        js = rng.sample(range(8), n)
        # priorities that often agree in their top bits
        base = rng.getrandbits(32)
        prio = [(base & ~0xFF) | rng.getrandbits(8) if rng.random() < 0.7 else rng.getrandbits(32) for _ in range(n)]
        order = list(range(n))
        rng.shuffle(order)
        cur, flag = _post_all(order, prio, js)
        exact = max(range(n), key=lambda k: (prio[k], js[k]))
        by_key = max(range(n), key=lambda k: (prio[k] & ~7) | js[k])
        assert flag or by_key == exact


def test_wider_tie_test_only_adds_false_alarms():
    """debug_tie_shift > 3 (the parity suite's way into the detour) flags a superset of the default's cases."""
    rng = random.Random(11)
    for _ in range(5000):
        n = rng.randint(1, 6)
        js = rng.sample(range(8), n)
        prio = [rng.getrandbits(32) for _ in range(n)]
        order = list(range(n))
        rng.shuffle(order)
        _, f3 = _post_all(order, prio, js, 3)
        _, f24 = _post_all(order, prio, js, 24)
        assert f24 or not f3


# ---- the play word (pd_dense.cuh, k_play_travel) ----------------------------------------------
R_OCC, R_DFLAG, R_LOW = 0x80000000, 0x800, 0x7FFF
TF_OCC, TF_D, TF_TRAIT = 0x80, 0x78, 0x03


def _word_spelled_out(t, sb, w):
    if not (t & TF_OCC):
        return 0
    return R_OCC | ((w >> 16) << 15) | (R_DFLAG if (t & TF_D) else 0) | (sb << 3) | ((t & TF_TRAIT) << 1)


def _words_branch_free(T, S, ww):
    out = []
    for k in range(4):
        r = ((T if k == 3 else (T << (24 - 8 * k))) & R_OCC) | ((ww[k] >> 1) & 0x7FFF8000)
        r |= ((S << 3) if k == 0 else (S >> (8 * k - 3))) & 0x7F8
        r |= ((T << 1) if k == 0 else (T >> (8 * k - 1))) & 0x6
        if (T >> (8 * k)) & TF_D:
            r |= R_DFLAG
        out.append(r & 0xFFFFFFFF)
    return out


def test_branch_free_play_words():
    rng = np.random.default_rng(3)
    for _ in range(20000):
        tb = [int(rng.integers(0, 4)) | (TF_OCC if rng.random() < 0.5 else 0) | (int(rng.integers(0, 16)) << 3 if rng.random() < 0.1 else 0)
              for _ in range(4)]
        tb = [t if (t & TF_OCC) else 0 for t in tb]          # canonical: an empty cell's flag byte is 0
        sbs = [int(rng.integers(0, 256)) for _ in range(4)]  # (an empty cell may hold a stale strategy byte)
        ww = [int(rng.integers(0, 1 << 32)) for _ in range(4)]
        T = sum(t << (8 * k) for k, t in enumerate(tb))
        S = sum(s << (8 * k) for k, s in enumerate(sbs))
        got = _words_branch_free(T, S, ww)
        for k in range(4):
            if tb[k] & TF_OCC:
                assert got[k] == _word_spelled_out(tb[k], sbs[k], ww[k])
            else:
                assert not (got[k] & R_OCC) and not (got[k] & R_DFLAG) and not (got[k] & 0x6)
    # an empty cell's word never "challenges": it is below both thresholds of any occupied cell
    for _ in range(2000):
        me = R_OCC | (int(rng.integers(0, 1 << 16)) << 15) | int(rng.integers(0, 1 << 11))
        tA, tB = me | R_LOW, ((me & ~R_LOW) - 1) & 0xFFFFFFFF
        empty = int(rng.integers(0, 1 << 31))
        assert not empty > tA and not empty > tB

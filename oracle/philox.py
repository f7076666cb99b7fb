"""Philox4x32-10 counter-based RNG, vectorised over NumPy arrays.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): this file is the checker's
copy of the generator.  The product path computes the same words on the GPU in
``flamegpu2-prisoners-dilemma-abm_b200/csrc/philox.cuh``.

Algorithm: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11),
Philox-4x32 with 10 rounds; multipliers 0xD2511F53 / 0xCD9E8D57 and Weyl key
increments 0x9E3779B9 / 0xBB67AE85 (the same constants cuRAND ships in
``curand_philox4x32_x.h``).  The reference model draws from FLAMEGPU2's
per-thread cuRAND streams (``FLAMEGPU->random.uniform<float>()``, e.g.
/root/reference/src/model.py:364, :611-612, :806-810); those streams depend on
FLAMEGPU2's agent-list order and cannot be reproduced, so every draw site is
re-keyed as ``(seed; cell, step, stream)`` -- see DESIGN.md "RNG keying".
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
SH32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Return the four 32-bit output words for counters (c0..c3) under key (k0,k1).

    Counters may be scalars or equally-shaped arrays; outputs are uint32 arrays.
    """
    c0 = np.asarray(c0).astype(np.uint64) & MASK32
    c1 = np.asarray(c1).astype(np.uint64) & MASK32
    c2 = np.asarray(c2).astype(np.uint64) & MASK32
    c3 = np.asarray(c3).astype(np.uint64) & MASK32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0  # 32x32 -> 64, exact in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> SH32, p0 & MASK32
        hi1, lo1 = p1 >> SH32, p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def philox_scalar(ctr, key):
    """Pure-Python scalar version (used by the known-answer tests and the literal oracle)."""
    c = [int(v) & 0xFFFFFFFF for v in ctr]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k1, p0 & 0xFFFFFFFF]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c)

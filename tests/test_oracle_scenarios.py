"""Hand-derived micro-scenarios pinning the oracle's semantics: the
reference ships no tests, so each case below states the expected outcome from the reference's
text (/root/reference/src/model.py line numbers in the docstrings) and recomputes the Philox
words it depends on independently with the scalar generator."""
import numpy as np
import pytest

from oracle.philox import philox_scalar, philox4x32_10
from oracle.pd_oracle import (Oracle, Params, DX, DY, S_ROLL, S_NOISE0, S_TRAVEL, ROLL_SHIFT, S_REPRO, S_CULL, S_ENERGY,
                              normal_from_words, prob_threshold, NORMAL_SCALE)


def world(W, agents, **kw):
    """agents: list of (x, y, energy, trait, [s0,s1,s2,s3])."""
    occ = np.zeros((W, W), np.uint8)
    en = np.zeros((W, W), np.float32)
    tr = np.zeros((W, W), np.uint8)
    st = np.zeros((W, W, 4), np.uint8)
    for x, y, e, t, s in agents:
        occ[y, x], en[y, x], tr[y, x], st[y, x] = 1, e, t, s
    kw.setdefault("seed", 12345)
    o = Oracle(Params(width=W, **kw))
    o.load_planes(occ, en, tr, st)
    return o


def words(o, cell, stream, step=None):
    step = o.step_index if step is None else step
    return philox_scalar((cell & 0xFFFFFFFF, cell >> 32, step, stream), (o.k0, o.k1))


def play_word(o, cell):
    """The play-phase word of a cell: one S_ROLL call per four cells in a row (counter = cell >> 2, word cell & 3):
    top 16 bits = pair-ordering roll, low 16 bits = the RANDOM coins."""
    return words(o, cell >> 2, S_ROLL)[cell & 3]


# ----------------------------------------------------------------------------- RNG pins
def test_philox_known_answers():
    """Random123 known-answer vectors (also the constants of curand_philox4x32_x.h)."""
    assert philox_scalar((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert philox_scalar((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    pi = philox_scalar((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0))
    assert pi == (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_philox_vector_equals_scalar():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 2 ** 32, size=(64, 4), dtype=np.uint64)
    out = philox4x32_10(c[:, 0], c[:, 1], c[:, 2], c[:, 3], 0xDEADBEEF, 0x12345678)
    for i in range(64):
        exp = philox_scalar(tuple(int(v) for v in c[i]), (0xDEADBEEF, 0x12345678))
        assert tuple(int(o[i]) for o in out) == exp


def test_probability_thresholds():
    assert prob_threshold(0.0) == 0
    assert prob_threshold(1.0) == 2 ** 32
    assert prob_threshold(0.5) == 2 ** 31
    assert prob_threshold(0.05) == int(np.floor(float(np.float32(0.05)) * 2 ** 32))


def test_normal_stand_in_moments():
    """Binomial(96,1/2)+dither: mean 0, variance 1 by construction; check on 2^18 draws."""
    i = np.arange(1 << 18, dtype=np.uint64)
    z = normal_from_words(*philox4x32_10(i, 0, 7, 13, 1, 2)).astype(np.float64)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs(((z ** 4).mean()) - 3.0) < 0.1  # kurtosis of a normal
    assert abs(NORMAL_SCALE * 2 ** 32 * np.sqrt(24 + 1 / 12) - 1.0) < 1e-12


# ----------------------------------------------------------------------------- geometry
def test_direction_algebra():
    """Offsets model.py:302-307, opposite = 7 - i, wrap by modulo (:311-312), bucket x + y*W (:339)."""
    assert list(DX) == [-1, -1, -1, 0, 0, 1, 1, 1] and list(DY) == [-1, 0, 1, -1, 1, -1, 0, 1]
    for i in range(8):
        assert DX[i] == -DX[7 - i] and DY[i] == -DY[7 - i]
    o = world(16, [(0, 0, 50.0, 0, [0] * 4), (15, 15, 50.0, 0, [0] * 4)])
    idx = np.array([0])
    assert o._nb_bucket(idx, 0)[0] == 15 + 15 * 16          # (-1,-1) wraps to (15,15)
    assert o._nb_bucket(idx, 7)[0] == 1 + 1 * 16
    assert o._nb_bucket(np.array([1]), 7)[0] == 0           # (15,15)+(1,1) wraps to (0,0)


def test_probe_order_is_the_triangular_permutation():
    """(l + i) % 8 accumulated: s, s+1, s+3, s+6, s+2, s+7, s+5, s+4 (model.py:826-831)."""
    o = world(16, [(5, 5, 50.0, 0, [0] * 4)])
    for s in range(8):
        seen = []
        for _ in range(8):
            o.nb_list[:] = 0
            for d in seen:
                o.nb_list[0, d] = 1
            found, l, seq = o._probe(np.array([0]), np.array([s]), np.array([0]))
            assert found[0] and seq[0] == len(seen) + 1
            seen.append(int(l[0]))
        assert seen == [(s + k) % 8 for k in (0, 1, 3, 6, 2, 7, 5, 4)]
        o.nb_list[:] = 1
        found, _, _ = o._probe(np.array([0]), np.array([s]), np.array([0]))
        assert not found[0]


# ----------------------------------------------------------------------------- play
STRATS = {0: "C", 1: "D", 2: "C"}  # COOP, DEFECT, TFT(==COOP, B-3)


@pytest.mark.parametrize("sa", [0, 1, 2])
@pytest.mark.parametrize("sb", [0, 1, 2])
def test_two_agent_play_table(sa, sb):
    """Only the RESPONDER (lower roll) is paid (play_resolve is never scheduled, model.py:1974-1978);
    payoffs model.py:677-689; the challenger's energy and games_played stay untouched."""
    o = world(16, [(4, 4, 50.0, 0, [sa] * 4), (5, 4, 60.0, 1, [sb] * 4)])
    o.search_for_neighbours()
    o.get_game_list()
    roll_a, roll_b = play_word(o, 4 + 4 * 16) >> ROLL_SHIFT, play_word(o, 5 + 4 * 16) >> ROLL_SHIFT
    assert roll_a != roll_b
    o.pdgame_model()
    a_is_responder = roll_a < roll_b
    my, their = (sa, sb) if a_is_responder else (sb, sa)
    pay = {("C", "C"): 3.0, ("D", "D"): 0.0, ("C", "D"): -1.0, ("D", "C"): 5.0}[(STRATS[my], STRATS[their])]
    ea, eb = float(o.energy[0]), float(o.energy[1])
    if a_is_responder:
        assert (ea, eb) == (50.0 + pay, 60.0)
    else:
        assert (ea, eb) == (50.0, 60.0 + pay)
    assert o.stats["games"] == 1


def test_random_strategy_and_noise_use_the_responders_words():
    """RANDOM cooperates iff its coin bit is set (model.py:631, :664): bits 2c (challenger) and 2c+1 (responder)
    of the RESPONDER's play word; noise flips a choice iff its word < noise * 2^32 (:638-640,
    :672-674): words of the responder's call S_NOISE0 + c//2, (x, y) for even c and (z, w) for odd c, where
    c = direction from challenger to responder."""
    for noise in (0.3, 0.0):
        o = world(16, [(4, 4, 50.0, 2, [3] * 4), (5, 5, 60.0, 1, [3] * 4)], env_noise=noise)
        o.search_for_neighbours()
        o.get_game_list()
        ca, cb = 4 + 4 * 16, 5 + 5 * 16
        a_resp = (play_word(o, ca) >> ROLL_SHIFT) < (play_word(o, cb) >> ROLL_SHIFT)
        resp_cell = ca if a_resp else cb
        c = 0 if a_resp else 7      # challenger b at (5,5) reaches a at (4,4) in direction 0; a reaches b in 7
        coins = play_word(o, resp_cell) & 0xFFFF
        w = words(o, resp_cell, S_NOISE0 + c // 2)
        my_roll, ch_roll = (w[2], w[3]) if c & 1 else (w[0], w[1])
        thr = prob_threshold(noise)
        i_coop = (((coins >> (2 * c + 1)) & 1) == 1) ^ (my_roll < thr)
        ch_coop = (((coins >> (2 * c)) & 1) == 1) ^ (ch_roll < thr)
        pay = 3.0 if (i_coop and ch_coop) else 0.0 if (not i_coop and not ch_coop) else -1.0 if i_coop else 5.0
        o.pdgame_model()
        e = float(o.energy[0 if a_resp else 1])
        assert e == (50.0 if a_resp else 60.0) + pay


def test_roll_tie_goes_to_the_neighbour_in_a_direction_ge_4():
    """The pair-ordering roll has 16 bits; on a tie the agent that sees the other in a direction < 4
    challenges (replaces the id tie-break of model.py:413): exactly one challenger per pair, from both sides."""
    for dirn in range(8):
        o = world(16, [(8, 8, 50.0, 0, [0] * 4), (8 + int(DX[dirn]), 8 + int(DY[dirn]), 50.0, 0, [0] * 4)])
        o.search_for_neighbours()
        o.roll[:] = 7            # force the tie
        o.msg_roll[o.bucket] = o.roll
        o.get_game_list()
        ia = int(np.flatnonzero((o.x == 8) & (o.y == 8))[0])
        ib = 1 - ia
        assert int(o.act[ia, dirn]) == (1 if dirn < 4 else 0)
        assert int(o.act[ib, 7 - dirn]) == (0 if dirn < 4 else 1)


def test_play_death_and_energy_clamp():
    """energy <= 0 after a game -> dead immediately (model.py:695-697); else min(energy, max) (:698-701)."""
    o = world(16, [(4, 4, 0.5, 0, [0] * 4), (5, 4, 149.0, 0, [1] * 4)])
    o.search_for_neighbours()
    o.get_game_list()
    coop_responds = (play_word(o, 4 + 64) >> ROLL_SHIFT) < (play_word(o, 5 + 64) >> ROLL_SHIFT)
    o.pdgame_model()
    if coop_responds:
        assert o.n == 1 and o.stats["play_deaths"] == 1 and float(o.energy[0]) == 149.0
    else:
        assert o.n == 2 and float(o.energy[1]) == 150.0   # 149 + 5 clamped to MAX_ENERGY


def test_mover_rule_pure_challenger_moves():
    """An agent that never completed a game as responder goes to MOVEMENT_UNRESOLVED iff
    act[7]==1 or act[0]==0 (model.py:486-496, :547-556, :719-730; SURVEY B-2)."""
    from oracle.pd_oracle import MOVEMENT_UNRESOLVED, READY, READY_TO_CHALLENGE
    for dirn in range(8):
        o = world(16, [(8, 8, 50.0, 0, [0] * 4), (8 + int(DX[dirn]), 8 + int(DY[dirn]), 50.0, 0, [0] * 4)])
        o.search_for_neighbours()
        o.get_game_list()
        ia = int(np.flatnonzero((o.x == 8) & (o.y == 8))[0])   # the list is in cell order
        ib = 1 - ia
        a_chal = bool(o.act[ia, dirn] == 1)
        o.pdgame_model()
        sa, sb = int(o.status[ia]), int(o.status[ib])
        # The terminal rule is only evaluated in sub-step 7, and only by agents that challenge in
        # direction 7 or respond from direction 0 there; everybody else leaves the submodel in
        # READY_TO_CHALLENGE (:486-496) whatever it played before.
        if a_chal:   # a challenges b at sub-step dirn; b responds from its direction 7-dirn
            assert sa == (MOVEMENT_UNRESOLVED if dirn == 7 else READY_TO_CHALLENGE)   # games_played == 0
            assert sb == (READY if dirn == 7 else READY_TO_CHALLENGE)                 # games_played == 1
        else:        # b challenges a at sub-step 7-dirn; a responds from its direction dirn
            assert sb == (MOVEMENT_UNRESOLVED if dirn == 0 else READY_TO_CHALLENGE)
            assert sa == (READY if dirn == 0 else READY_TO_CHALLENGE)
        assert o.stats["games"] == 1


# ----------------------------------------------------------------------------- travel
def test_isolated_agent_moves_one_cell_and_pays():
    """No neighbours -> move (model.py:429-433): energy - travel_cost (:800), start direction
    from the draw (:810), first free slot = the start; then cost of living (:1357-1368)."""
    o = world(16, [(8, 8, 50.0, 3, [1, 2, 3, 0])])
    cell = 8 + 8 * 16
    s = words(o, cell, S_TRAVEL)[1] >> 29
    st = o.step()
    assert st["movers"] == 1 and st["moved"] == 1
    assert (int(o.x[0]), int(o.y[0])) == (8 + int(DX[s]), 8 + int(DY[s]))
    assert float(o.energy[0]) == 50.0 - 0.5 - 1.0
    assert list(o.strat[0]) == [1, 2, 3, 0] and int(o.trait[0]) == 3


def test_travel_death_and_living_death():
    o = world(16, [(3, 3, 0.4, 0, [0] * 4), (10, 10, 1.2, 0, [0] * 4)])
    st = o.step()
    assert st["travel_deaths"] == 1     # 0.4 - 0.5 <= 0 (model.py:801-804)
    assert st["living_deaths"] == 1     # 1.2 - 0.5 = 0.7; 0.7 - 1.0 <= 0 (:1364-1366)
    assert o.n == 0


def test_two_movers_one_target_highest_roll_wins():
    """Conflict on a target: max (roll, tag) wins (model.py:929); the loser retries with the same
    roll and no further cost (:940-943)."""
    found = False
    for seed in range(400):
        o = world(16, [(4, 4, 50.0, 0, [0] * 4), (6, 4, 50.0, 0, [0] * 4)], seed=seed)
        ca, cb = 4 + 64, 6 + 64
        wa, wb = words(o, ca, S_TRAVEL), words(o, cb, S_TRAVEL)
        ta = (4 + int(DX[wa[1] >> 29]), 4 + int(DY[wa[1] >> 29]))
        tb = (6 + int(DX[wb[1] >> 29]), 4 + int(DY[wb[1] >> 29]))
        if ta != tb:
            continue
        found = True
        st = o.step()
        a_wins = (wa[0], ca) > (wb[0], cb)
        pos = {(int(x), int(y)) for x, y in zip(o.x, o.y)}
        assert ta in pos
        winner_start = (4, 4) if a_wins else (6, 4)
        assert winner_start not in pos       # the winner left its cell
        assert st["movers"] == 2 and st["move_rounds"] >= 2
        assert np.all(o.energy == np.float32(50.0 - 0.5 - 1.0))   # the retry costs nothing extra
        break
    assert found


# ----------------------------------------------------------------------------- reproduction
def test_reproduction_threshold_cost_child_and_newborn_exemption():
    """energy >= 100 -> attempt (model.py:975-981); parent pays 50 (:1211); child: parent's trait
    (:1224-1226), N(50,10) clamped to [5,150] (:1228-1250), strategies copied when mutation_rate=0;
    newborn pays no living cost in its birth step (:1343)."""
    o = world(16, [(8, 8, 130.0, 2, [0, 1, 2, 3])])
    st = o.step()
    assert st["births"] == 1 and o.n == 2
    parent = int(np.argmax(o.energy))
    child = 1 - parent
    assert float(o.energy[parent]) == 130.0 - 0.5 - 50.0 - 1.0
    pcell = int(o.bucket[parent])
    z = normal_from_words(*[np.array([w], dtype=np.uint32) for w in words(o, pcell, S_ENERGY, step=0)])[0]
    exp = np.float32(np.float32(z * np.float32(10.0)) + np.float32(50.0))
    exp = min(max(exp, np.float32(5.0)), np.float32(150.0))
    assert o.energy[child] == exp
    assert int(o.trait[child]) == 2
    # kin/other mode: the child's "other" slots are all the parent's FIRST other slot (:1296-1323)
    assert list(o.strat[child]) == [0, 0, 2, 0]
    d = words(o, pcell, S_REPRO, step=0)[1] >> 29
    assert (int(o.x[child]), int(o.y[child])) == ((int(o.x[parent]) + int(DX[d])) % 16,
                                                  (int(o.y[parent]) + int(DY[d])) % 16)


def test_no_reproduction_below_threshold_or_when_surrounded():
    from oracle.pd_oracle import REPRODUCTION_IMPOSSIBLE
    o = world(16, [(8, 8, 99.0, 0, [0] * 4)])
    assert o.step()["births"] == 0
    ring = [(8 + int(DX[d]), 8 + int(DY[d]), 20.0, 0, [0] * 4) for d in range(8)]
    o = world(16, [(8, 8, 140.0, 0, [0] * 4)] + ring)
    o.search_for_neighbours()
    o.get_game_list()
    o.neighbourhood_model()
    assert int(o.status[np.flatnonzero((o.x == 8) & (o.y == 8))[0]]) == REPRODUCTION_IMPOSSIBLE   # :1028


def test_mutation_modes():
    """pure: every child mutates when rate > 0 (model.py:1262-1267, B-7); per-trait: each slot
    independently (:1273-1286); kin/other without mutation copies (:1287-1326)."""
    o = world(16, [(8, 8, 130.0, 1, [2, 2, 2, 2])], strategy_pure=1, mutation_rate=1e-9)
    o.step()
    child = int(np.argmin(o.energy))
    cs = list(o.strat[child])
    assert len(set(cs)) == 1 and cs[0] != 2
    o = world(16, [(8, 8, 130.0, 1, [0, 1, 2, 3])], strategy_per_trait=1, mutation_rate=1.0)
    o.step()
    child = int(np.argmin(o.energy))
    assert all(int(c) != p for c, p in zip(o.strat[child], [0, 1, 2, 3]))
    o = world(16, [(8, 8, 130.0, 1, [0, 1, 0, 0])], mutation_rate=0.0)
    o.step()
    assert list(o.strat[int(np.argmin(o.energy))]) == [0, 1, 0, 0]


def test_carrying_capacity_priority_cull():
    """olds + births > max_agents -> only the max_agents - olds newborns with the highest
    (cull word, cell) keys survive (replaces the list-index cull model.py:1343/:1354, B-4)."""
    agents = [(2 + 4 * i, 2 + 4 * j, 130.0, 0, [0] * 4) for i in range(3) for j in range(3)]
    o = world(16, agents, max_agents=12)
    st = o.step()
    assert st["births"] == 9 and st["culled"] == 6 and o.n == 12
    o2 = world(16, agents, max_agents=1000)   # same world without the cap: all 9 children live
    o2.step()
    kids = np.flatnonzero(o2.id > 9)
    assert kids.size == 9
    keys = sorted(((words(o2, int(o2.bucket[k]), S_CULL, step=0)[0], int(o2.bucket[k])) for k in kids), reverse=True)
    survivors = {c for _, c in keys[:3]}
    got = {int(b) for b, i in zip(o.bucket, o.id) if i > 9}
    assert got == survivors


def test_counts_bins():
    """population_strat_count[4*my + other] (model.py:1413-1419, :1501-1511)."""
    o = world(16, [(1, 1, 50.0, 0, [3, 1, 1, 1]), (5, 5, 50.0, 2, [0, 0, 2, 0]), (9, 9, 50.0, 1, [1, 3, 1, 1])])
    c = o.strategy_counts()
    assert c[4 * 3 + 1] == 2 and c[4 * 2 + 0] == 1 and c.sum() == 3


# ----------------------------------------------------------------------------- invariants
@pytest.mark.parametrize("kw", [dict(), dict(env_noise=0.1, mutation_rate=0.05, strategy_per_trait=1),
                                dict(strategy_pure=1, cost_of_living=2.0)])
def test_step_invariants(kw):
    """<= 1 agent per cell; N' = N + births - deaths - culled; 0 < energy <= max; cap respected."""
    o = Oracle(Params(width=32, seed=5, **kw))
    o.init_population()
    assert o.n == int(32 * 32 * 0.16)
    for _ in range(80):
        n0 = o.n
        st = o.step()
        o.planes()  # asserts one agent per cell
        assert o.n == n0 - st["play_deaths"] - st["travel_deaths"] + st["births"] - st["culled"] - st["living_deaths"]
        assert o.n <= o.p.max_agents
        assert np.all(o.energy > 0) and np.all(o.energy <= np.float32(150.0))
        assert o.population_strat_count.sum() == o.n

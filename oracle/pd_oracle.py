"""NumPy transcription of the reference's agent functions (agent-LIST based).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED (no golden
vectors exist upstream); this file restates /root/reference/src/model.py
function by function so that the CUDA grid-plane implementation (which shares no
code and no data layout with it) can be checked bit-exactly.

Structure mirrors the reference: a structure-of-arrays *agent list* (like
FLAMEGPU2's), bucket "messages" as dense per-cell arrays, the four submodels
with their host exit conditions and 8-iteration limits, agent death = list
compaction, agent birth = list append.  Every method cites the reference lines
it follows.  Three deliberate restatements (DESIGN.md "Semantics"):

* all random draws are Philox4x32-10 words keyed (seed; cell, step, stream)
  instead of per-thread cuRAND (model.py:364, :611-612, :631, :664, :806-810,
  :1076-1079, :1236, :1263-1314);
* ONE Philox call serves the play phase of four cells in a row (counter = cell >> 2,
  the cell's word is word cell & 3): its top 16 bits are the pair-ordering roll
  (model.py:364; a tie is won by the neighbour in a direction >= 4 -- the reference:
  a float and the agent id, :413), its low 16 bits the RANDOM-strategy coins of the
  games the agent RESPONDS to (:631, :664; bits 2c and 2c+1 for sub-step c), so a
  game without noise needs no random call of its own; the noise rolls (:611-612)
  come from four calls per responder (streams 1..4, two sub-steps per call);
* travel / reproduction claims are ordered by a 32-bit word of a per-cell call and
  ties are broken by where the claimant stands as seen from the contested cell
  (7 - claimed direction, unique among the rivals of one cell; the reference
  uses the agent id, model.py:929, :1192);
* the list-index cull (model.py:1343, :1354) is replaced by "the surplus
  newborns with the lowest Philox cull priority die" (SURVEY Appendix B-4).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

from .philox import philox4x32_10

# --- status values, model.py:196-208 -----------------------------------------
READY = 1
READY_TO_CHALLENGE = 2
READY_TO_RESPOND = 8
MOVEMENT_UNRESOLVED = 64
MOVING = 128
ATTEMPTING_REPRODUCTION = 512
REPRODUCTION_IMPOSSIBLE = 1024
REPRODUCTION_COMPLETE = 2048
NEW_AGENT = 4096

# --- strategies, model.py:121-124 --------------------------------------------
COOP, DEFECT, TFT, RANDOM = 0, 1, 2, 3

# --- Moore offsets, model.py:302-307; opposite direction of i is 7-i ----------
DX = np.array([-1, -1, -1, 0, 0, 1, 1, 1], dtype=np.int64)
DY = np.array([-1, 0, 1, -1, 1, -1, 0, 1], dtype=np.int64)
NDIR = 8  # SPACES_WITHIN_RADIUS, model.py:255

# --- Philox stream ids (4th counter word); step index is the 3rd --------------
S_ROLL = 0      # counter = cell >> 2, word = cell & 3: top 16 bits pair-ordering roll (:364), bits 2c /
                # 2c+1: the challenger's / my RANDOM coin (:631 / :664) of the game I respond to in
                # sub-step c
S_NOISE0 = 1    # +(c>>1), keyed by the RESPONDER's cell: even c: w0 my noise roll (:611) w1 the
                # challenger's (:612); odd c: w2 mine, w3 the challenger's
ROLL_SHIFT = 16  # the pair-ordering roll is the top half of the cell's word, the coins the low half
S_TRAVEL = 14   # w0 travel priority (:806) w1>>29 travel start direction (:810)
S_REPRO = 9     # w0 reproduction claim priority (:1076) w1>>29 start dir (:1079)
S_MUT = 10      # w0..w3 mutation rolls (:1263, :1278, :1291-1292), keyed by the PARENT's cell
S_PICK = 11     # w0..w3 replacement-strategy picks (:1265, :1281, :1303, :1314)
S_CULL = 12     # w0 newborn cull priority, keyed by the CHILD's cell (replaces :1343/:1354)
S_ENERGY = 13   # w0..w3 -> child energy normal deviate (:1236), keyed by the PARENT's cell
INIT_STEP = 0xFFFFFFFF  # separate namespace for init_fn draws (:1438-1498)
I_PLACE = 0     # w0 placement priority (:1438-1450) w1>>30 trait (:1467)
I_STRAT = 1     # w0..w3 strategy draws (:1473-1497)
I_ENERGY = 2    # w0..w3 -> initial energy normal deviate (:1460-1465)

# 1 / (2^32 * sqrt(24 + 1/12)): scale of the integer-exact normal deviate below
NORMAL_SCALE = float.fromhex("0x1.a15286234fcddp-35")


def _normal_scale_check() -> None:
    import math
    exact = 1.0 / (4294967296.0 * math.sqrt(24.0 + 1.0 / 12.0))
    assert abs(NORMAL_SCALE - exact) < 1e-22, (NORMAL_SCALE.hex(), exact.hex())


def normal_from_words(w0, w1, w2, w3) -> np.ndarray:
    """Integer-exact N(0,1) stand-in: Binomial(96,1/2) + uniform dither, as float32.

    z = ((popc(w0)+popc(w1)+popc(w2) - 48) * 2^32 + (w3 - 2^31)) * NORMAL_SCALE,
    evaluated in float64 (the integer is < 2^39, exact) and rounded once to
    float32.  Replaces ``FLAMEGPU->random.normal<float>()`` (model.py:1236) and
    ``random.normalvariate`` (model.py:1460-1462) with something NumPy and CUDA
    compute bit-identically.
    """
    pc = (np.bitwise_count(w0).astype(np.int64) + np.bitwise_count(w1).astype(np.int64)
          + np.bitwise_count(w2).astype(np.int64) - 48)
    zi = pc * (1 << 32) + (w3.astype(np.int64) - (1 << 31))
    return (zi.astype(np.float64) * NORMAL_SCALE).astype(np.float32)


def prob_threshold(p) -> int:
    """P(event) = p  <=>  (u32 word) < threshold; p is first rounded to float32 like an
    environment property (model.py:1683, :1738)."""
    v = float(np.float32(p)) * 4294967296.0
    if not v > 0.0:
        return 0
    return int(min(4294967296.0, np.floor(v)))


def pick3(word) -> np.ndarray:
    """Uniform pick in {0,1,2} from a u32 word (multiply-shift)."""
    return ((word.astype(np.uint64) * np.uint64(3)) >> np.uint64(32)).astype(np.uint8)


@dataclass
class Params:
    """Mirror of the reference's configuration block (model.py:39-169) and
    environment properties (model.py:1678-1743)."""
    width: int = 512                       # ENV_MAX, model.py:211
    seed: int = 12345                      # RANDOM_SEED, model.py:39
    init_agent_count: int = -1             # INIT_AGENT_COUNT, model.py:45 (-1: 16 % of cells)
    max_agents: int = -1                   # AGENT_HARD_LIMIT, model.py:49 (-1: 50 % of cells)
    payoff_cc: float = 3.0                 # model.py:93
    payoff_dc: float = 5.0                 # model.py:95
    payoff_cd: float = -1.0                # model.py:97
    payoff_dd: float = 0.0                 # model.py:99
    env_noise: float = 0.0                 # model.py:117
    cost_of_living: float = 1.0            # model.py:78
    travel_cost: float = 0.5               # model.py:106
    reproduce_min_energy: float = 100.0    # model.py:81
    reproduce_cost: float = 50.0           # model.py:83
    reproduction_inheritence: float = 0.0  # model.py:88
    max_children_per_step: int = 1         # model.py:90
    max_energy: float = 150.0              # model.py:109
    init_energy_mu: float = 50.0           # model.py:111
    init_energy_sigma: float = 10.0        # model.py:112
    init_energy_min: float = 5.0           # model.py:115
    mutation_rate: float = 0.0             # model.py:164
    strategy_pure: int = 0                 # model.py:158
    strategy_per_trait: int = 0            # model.py:161
    strategy_weights: List[float] = field(default_factory=lambda: [0.25, 0.25, 0.25, 0.25])  # :127-149
    # ORACLE-ONLY switch: "priority" = the cull the product implements (SURVEY B-4); "list_index" = the
    # reference's own rule (model.py:1343, :1354) under the FLAMEGPU2 list semantics of SURVEY Appendix D-2
    # [FG2-ext]; tests/test_cull_variants.py compares the two statistically
    cull: str = "priority"
    # ORACLE-ONLY switch (SURVEY B-8): count agents by the reference's agent_strategy_id VARIABLE instead of by
    # their strategies.  god_multiply sets it for the children of pure and kin/other runs (model.py:1271, :1326) but
    # not for per-trait runs (:1273-1286), whose children keep the default 0 and are all counted in bin 0 -- this is
    # what a per-trait log of the reference shows, so a comparison against one needs it.  Reporting only: the
    # dynamics do not depend on it.  The CUDA path derives the bin (it has no per-agent id state to carry).
    legacy_counts: bool = False

    def __post_init__(self):
        w = self.width
        if w < 4 or (w & (w - 1)):
            raise ValueError("width must be a power of two >= 4 (model.py:311-312 wraps by unsigned modulo)")
        cells = w * w
        if self.init_agent_count < 0:
            self.init_agent_count = int(cells * 0.16)   # model.py:45
        if self.max_agents < 0:
            self.max_agents = int(cells * 0.5)          # model.py:49


def strategy_cum_thresholds(weights) -> np.ndarray:
    """Three u64 thresholds t1<=t2<=t3; strategy = #(word >= t_k).  model.py:218-224, :1475."""
    w = np.asarray(weights, dtype=np.float32).astype(np.float64)  # properties are float32
    wsum = 0.0
    for v in w:
        wsum += float(v)
    cum, acc = [], 0.0
    for v in w:
        acc += float(v)
        cum.append(acc / wsum)
    return np.array([min(4294967296.0, np.floor(c * 4294967296.0)) for c in cum[:3]], dtype=np.uint64)


class Oracle:
    """One simulation: ``init_population()`` then ``step()`` repeatedly."""

    def __init__(self, params: Params):
        _normal_scale_check()
        self.p = params
        self.W = params.width
        self.cells = self.W * self.W
        self.k0 = params.seed & 0xFFFFFFFF
        self.k1 = (params.seed >> 32) & 0xFFFFFFFF
        self.step_index = 0
        self.next_id = 1
        f32 = np.float32
        self.payoff_cc, self.payoff_cd = f32(params.payoff_cc), f32(params.payoff_cd)
        self.payoff_dc, self.payoff_dd = f32(params.payoff_dc), f32(params.payoff_dd)
        self.noise_thr = prob_threshold(params.env_noise)
        self.mut_thr = prob_threshold(params.mutation_rate)
        self.strat_thr = strategy_cum_thresholds(params.strategy_weights)
        # agent list (structure of arrays) -- model.py:1630-1657
        self.id = np.zeros(0, np.int64)
        self.x = np.zeros(0, np.int64)
        self.y = np.zeros(0, np.int64)
        self.energy = np.zeros(0, np.float32)
        self.status = np.zeros(0, np.int64)
        self.trait = np.zeros(0, np.uint8)
        self.strat = np.zeros((0, 4), np.uint8)
        self.nb_list = np.zeros((0, NDIR), np.int64)       # neighbour_list (0 = ID_NOT_SET)
        self.act = np.full((0, NDIR), -1, np.int8)          # my_actions
        self.roll = np.zeros(0, np.uint32)                  # die_roll
        self.roll_cell = np.zeros(0, np.int64)              # tie-break tag: cell at draw time
        self.coins = np.zeros(0, np.uint32)                 # RANDOM coins of the games I respond to
        self.sid_bin = np.zeros(0, np.int64)                # count bin of agent_strategy_id (legacy_counts)
        self.bucket = np.zeros(0, np.int64)                 # my_bucket
        self.stats: Dict[str, int] = {}
        self.population_strat_count = np.zeros(16, np.int64)

    # ------------------------------------------------------------------ helpers
    def _rng(self, cell, stream, step=None):
        step = self.step_index if step is None else step
        cell = np.asarray(cell, dtype=np.uint64)
        return philox4x32_10(cell & np.uint64(0xFFFFFFFF), cell >> np.uint64(32), step, stream,
                             self.k0, self.k1)

    _ARRS = ("id", "x", "y", "energy", "status", "trait", "strat", "nb_list", "act", "roll",
             "roll_cell", "bucket", "coins", "sid_bin")

    def _compact(self, alive: np.ndarray, extra: Dict[str, np.ndarray] | None = None):
        """Remove dead agents from every per-agent array (FLAMEGPU2 death compaction)."""
        for n in self._ARRS:
            setattr(self, n, getattr(self, n)[alive])
        if extra:
            for k in list(extra):
                extra[k] = extra[k][alive]

    def _partition(self, passing: np.ndarray, sub: Dict[str, np.ndarray] | None = None, remap_chal: bool = False):
        """cull == "list_index" only.  [FG2-ext, SURVEY Appendix D-2]: a function condition stably
        partitions the agent list -- failing agents first, passing agents after them -- and the order is
        not restored; deaths compact stably (_compact) and newborns append at the end.  Nothing but the
        list-index cull (model.py:1343, :1354) ever looks at the order."""
        if self.p.cull != "list_index":
            return
        order = np.concatenate([np.flatnonzero(~passing), np.flatnonzero(passing)])
        for n in self._ARRS:
            setattr(self, n, getattr(self, n)[order])
        if sub:
            for k in list(sub):
                sub[k] = sub[k][order]
        if remap_chal:   # this layer's challenge messages hold list indices (challenger of a bucket)
            cb = self.chal_bucket
            inv = np.empty_like(order)
            inv[order] = np.arange(order.size)
            self.chal_bucket = np.where(cb >= 0, inv[np.maximum(cb, 0)], -1)

    def _nb_bucket(self, idx: np.ndarray, d) -> np.ndarray:
        """pos_from_moore_seq + pos_to_bucket_id, model.py:294-342."""
        W = self.W
        nx = (self.x[idx] + DX[d]) % W
        ny = (self.y[idx] + DY[d]) % W
        return nx + ny * W

    @property
    def n(self) -> int:
        return int(self.id.shape[0])

    # ------------------------------------------------------------------ init_fn
    def play_word(self, cell) -> np.ndarray:
        """The play-phase word of a cell: one S_ROLL call per group of four cells (counter = cell >> 2)."""
        cell = np.asarray(cell, dtype=np.uint64)
        w = np.stack(self._rng(cell >> np.uint64(2), S_ROLL), axis=-1)
        k = (cell & np.uint64(3)).astype(np.int64)
        return np.take_along_axis(w, k[..., None], axis=-1)[..., 0] if w.ndim > 1 else w[int(k)]

    def init_population(self):
        """init_fn.run, model.py:1423-1524 -- Philox-keyed, exact count (B-10)."""
        p, W, cells = self.p, self.W, self.cells
        n = p.init_agent_count
        allc = np.arange(cells, dtype=np.uint64)
        w0, w1, _, _ = self._rng(allc, I_PLACE, INIT_STEP)
        # n cells with the smallest (priority, cell) key are occupied (model.py:1438-1450)
        key = (w0.astype(np.uint64) << np.uint64(32)) | allc
        chosen = np.sort(np.argpartition(key, n - 1)[:n]) if 0 < n < cells else (
            np.arange(cells) if n >= cells else np.zeros(0, np.int64))
        chosen = chosen.astype(np.int64)
        trait = (w1[chosen] >> np.uint32(30)).astype(np.uint8)                  # :1467
        e0, e1, e2, e3 = self._rng(chosen, I_ENERGY, INIT_STEP)
        z = normal_from_words(e0, e1, e2, e3)
        energy = (z * np.float32(p.init_energy_sigma) + np.float32(p.init_energy_mu)).astype(np.float32)
        energy = np.maximum(energy, np.float32(p.init_energy_min))             # :1460-1462
        if p.max_energy > 0.0:
            energy = np.minimum(energy, np.float32(p.max_energy))              # :1463-1464
        s0, s1, s2, s3 = self._rng(chosen, I_STRAT, INIT_STEP)
        draw = lambda w: (w.astype(np.uint64)[:, None] >= self.strat_thr[None, :]).sum(axis=1).astype(np.uint8)
        strat = np.zeros((n, 4), np.uint8)
        if p.strategy_pure:                                                    # :1473-1477
            strat[:] = draw(s0)[:, None]
        elif p.strategy_per_trait:                                             # :1478-1482
            for i, s in enumerate((s0, s1, s2, s3)):
                strat[:, i] = draw(s)
        else:                                                                  # :1483-1497
            my, other = draw(s0), draw(s1)
            strat[:] = other[:, None]
            strat[np.arange(n), trait] = my
        self.id = np.arange(1, n + 1, dtype=np.int64)
        self.next_id = n + 1
        self.x = chosen % W
        self.y = chosen // W
        self.energy = energy
        self.status = np.full(n, READY, np.int64)
        self.trait = trait
        self.strat = strat
        self.nb_list = np.zeros((n, NDIR), np.int64)
        self.act = np.full((n, NDIR), -1, np.int8)
        self.roll = np.zeros(n, np.uint32)
        self.roll_cell = np.zeros(n, np.int64)
        self.coins = np.zeros(n, np.uint32)
        self.sid_bin = self._derived_bin(strat, trait)                         # agent_strategy_id, :1501-1511
        self.bucket = self.x + self.y * W
        self.step_index = 0
        self.population_strat_count = self.strategy_counts()                   # :1517-1522

    def load_planes(self, occ, energy, trait, strat):
        """Build the agent list from grid planes (row-major [y, x])."""
        W = self.W
        cell = np.flatnonzero(np.asarray(occ).reshape(-1))
        n = cell.shape[0]
        self.id = np.arange(1, n + 1, dtype=np.int64)
        self.next_id = n + 1
        self.x, self.y = cell % W, cell // W
        self.energy = np.asarray(energy, np.float32).reshape(-1)[cell].copy()
        self.trait = np.asarray(trait, np.uint8).reshape(-1)[cell].copy()
        self.strat = np.asarray(strat, np.uint8).reshape(-1, 4)[cell].copy()
        self.status = np.full(n, READY, np.int64)
        self.nb_list = np.zeros((n, NDIR), np.int64)
        self.act = np.full((n, NDIR), -1, np.int8)
        self.roll = np.zeros(n, np.uint32)
        self.roll_cell = np.zeros(n, np.int64)
        self.coins = np.zeros(n, np.uint32)
        self.sid_bin = self._derived_bin(self.strat, self.trait)
        self.bucket = self.x + self.y * W
        self.population_strat_count = self.strategy_counts()

    # --------------------------------------------------- main layers 1-2 (K1, K2)
    def search_for_neighbours(self):
        """model.py:352-372: bucket id, per-step die roll, post (id, roll) at my bucket."""
        self.bucket = self.x + self.y * self.W                                  # :361
        word = self.play_word(self.bucket)
        self.roll = word >> np.uint32(ROLL_SHIFT)                               # :364
        self.roll_cell = self.bucket.copy()
        self.coins = word & np.uint32(0xFFFF)
        self.msg_id = np.zeros(self.cells, np.int64)                           # player_search_msg
        self.msg_roll = np.zeros(self.cells, np.uint32)
        self.msg_id[self.bucket] = self.id
        self.msg_roll[self.bucket] = self.roll

    def get_game_list(self):
        """model.py:373-440: neighbour scan, challenge/respond role per neighbour."""
        n = self.n
        idx = np.arange(n)
        num_neighbours = np.zeros(n, np.int64)
        for i in range(NDIR):                                                   # :393
            nb = self._nb_bucket(idx, i)                                        # :394-395
            nid = self.msg_id[nb]                                               # :403
            nroll = self.msg_roll[nb]                                           # :408
            present = nid != 0
            num_neighbours += present
            # I challenge iff my roll is higher -- :413; a tie goes to whoever sees the other in a
            # direction < 4 (see module doc)
            higher = (self.roll > nroll) | ((self.roll == nroll) & (i < 4))
            self.act[:, i] = np.where(present, np.where(higher, 1, 0), -1)      # :415-418, :399
            self.nb_list[:, i] = np.where(present, nid, 0)                      # :423
        self.status = np.where(num_neighbours == 0, MOVEMENT_UNRESOLVED, READY_TO_CHALLENGE)  # :429-436

    # ------------------------------------------------------ pdgame submodel (L3)
    def pdgame_model(self):
        """Submodel wiring model.py:1900-1978; exit_play_fn model.py:1527-1556."""
        n = self.n
        # submodel-local variables restart from their defaults (model.py:1690-1696)
        sub = {
            "challenge_sequence": np.zeros(n, np.int64),
            "games_played": np.zeros(n, np.int64),
        }
        iterations = 0
        games = deaths = 0
        while True:
            self.chal_from = {}
            self._play_challenge(sub)
            g, d = self._play_response(sub)
            games += g
            deaths += d
            iterations += 1                                                     # :1536
            if iterations < NDIR and np.any((self.status == READY_TO_CHALLENGE)
                                            | (self.status == READY_TO_RESPOND)):  # :1538-1553
                continue
            break
        self.stats["games"] = games
        self.stats["play_deaths"] = deaths

    def _play_challenge(self, sub):
        """play_challenge, model.py:450-566 (condition :442-448)."""
        self._partition(self.status == READY_TO_CHALLENGE, sub)
        sel = np.flatnonzero(self.status == READY_TO_CHALLENGE)
        # message list of this layer: target bucket -> challenger agent index
        self.chal_bucket = np.full(self.cells, -1, np.int64)
        if sel.size == 0:
            return
        c = sub["challenge_sequence"][sel]                                      # :459
        # (:462-475 "sequence >= 8" never fires: the submodel stops after 8 iterations)
        r = NDIR - c - 1                                                        # :477
        a_c = self.act[sel, c]                                                  # :480
        a_r = self.act[sel, r]                                                  # :481
        ch = a_c == 1                                                           # :482
        rs = a_r == 0                                                           # :483
        # challengers post a message keyed by the neighbour's bucket -- :511-537
        chs = sel[ch]
        if chs.size:
            tgt = self._nb_bucket(chs, c[ch])                                   # :518-520
            self.chal_bucket[tgt] = chs
        self.chal_dir = c                                                       # (sub-step, per agent)
        c1 = c + 1
        sub["challenge_sequence"][sel] = c1                                     # :491, :499, :539
        st = self.status[sel]
        st = np.where(rs, READY_TO_RESPOND, st)                                 # :500, :542
        # challenged but no response pending: next sub-step or terminal rule -- :544-557
        term = ch & ~rs & (c1 >= NDIR)
        st = np.where(term, np.where(sub["games_played"][sel] < 1, MOVEMENT_UNRESOLVED, READY), st)
        # (!ch && !rs) stays READY_TO_CHALLENGE even at c == 7 -- :486-496
        self.status[sel] = st

    def _play_response(self, sub):
        """play_response, model.py:575-734 (condition :569-574)."""
        p = self.p
        self._partition(self.status == READY_TO_RESPOND, sub, remap_chal=True)
        sel = np.flatnonzero(self.status == READY_TO_RESPOND)
        if sel.size == 0:
            return 0, 0
        chal = self.chal_bucket[self.bucket[sel]]                               # :585
        has = chal >= 0
        # responder_id == my_id check -- :586-587 (always true for a live occupant)
        c_sub = sub["challenge_sequence"][sel] - 1                              # the sub-step being played
        if has.any():
            cc = np.where(has, chal, 0)
            d_resp = NDIR - 1 - c_sub
            assert np.all(self.nb_list[cc, c_sub][has] == self.id[sel][has])
            assert np.all(self.nb_list[sel, d_resp][has] == self.id[cc][has])
        me = sel[has]
        ch = chal[has]
        games = int(me.size)
        dead = np.zeros(self.n, bool)
        if games:
            c_g = c_sub[has]
            my_strategy = self.strat[me, self.trait[ch]]                        # :599
            ch_strategy = self.strat[ch, self.trait[me]]                        # :600
            coins = self.coins[me] >> (2 * c_g).astype(np.uint32)               # :631, :664
            # COOP -> C, DEFECT -> D, TFT -> C (memory never matches: it is submodel-local and
            # each pair plays once per step, :621-629, :646-661), RANDOM -> coin (:630-635)
            ch_coop = np.where(ch_strategy == RANDOM, (coins & np.uint32(1)) == 1, ch_strategy != DEFECT)
            i_coop = np.where(my_strategy == RANDOM, (coins & np.uint32(2)) == 2, my_strategy != DEFECT)
            if self.noise_thr:
                n0, n1, n2, n3 = self._rng(self.bucket[me], S_NOISE0 + (c_g >> 1))   # :611-612
                odd = (c_g & 1) == 1
                my_roll, ch_roll = np.where(odd, n2, n0), np.where(odd, n3, n1)
                ch_coop = ch_coop ^ (ch_roll.astype(np.uint64) < np.uint64(self.noise_thr))   # :638-640
                i_coop = i_coop ^ (my_roll.astype(np.uint64) < np.uint64(self.noise_thr))     # :672-674
            pay = np.where(i_coop & ch_coop, self.payoff_cc,
                           np.where(~i_coop & ~ch_coop, self.payoff_dd,
                                    np.where(i_coop & ~ch_coop, self.payoff_cd, self.payoff_dc)))  # :677-689
            e = (self.energy[me] + pay.astype(np.float32)).astype(np.float32)
            died = e <= np.float32(0)                                           # :695-697
            e = np.minimum(e, np.float32(p.max_energy))                         # :698-701
            self.energy[me] = np.where(died, self.energy[me], e)               # :702
            sub["games_played"][me[~died]] += 1                                 # :710
            dead[me[died]] = True
        # next sub-step or terminal rule -- :715-731
        cs = sub["challenge_sequence"][sel]
        st = np.where(cs < NDIR, READY_TO_CHALLENGE,
                      np.where(sub["games_played"][sel] < 1, MOVEMENT_UNRESOLVED, READY))
        self.status[sel] = st
        nd = int(dead.sum())
        if nd:
            self._compact(~dead, sub)
        return games, nd

    # ---------------------------------------------------- movement submodel (L4)
    def movement_model(self):
        """Submodel wiring model.py:1980-2030; exit_move_fn model.py:1559-1576."""
        n = self.n
        sub = {  # model.py:1708-1711
            "last_move_attempt": np.full(n, NDIR + 1, np.int64),
            "request_bucket": np.zeros(n, np.int64),
            "move_sequence": np.zeros(n, np.int64),
        }
        iterations = 0
        self.stats.update(movers=0, moved=0, travel_deaths=0, move_rounds=0)
        while True:
            self._move_request(sub, first=(iterations == 0))
            self._move_response(sub)
            iterations += 1                                                     # :1568
            self.stats["move_rounds"] = iterations
            if iterations < NDIR and np.any(self.status == MOVEMENT_UNRESOLVED):   # :1570-1573
                continue
            break

    @staticmethod
    def _claim_key(roll, direction) -> np.ndarray:
        """Contest key of a claim (:929, :1192): the highest roll wins; equal rolls (2^-32 per pair of rivals) are
        decided by where the claimant stands as seen from the contested cell, j = 7 - claimed direction, which is
        unique among the rivals of one cell and takes the place of the reference's id comparison.  +1 keeps 0 for
        'no claim'."""
        j = (7 - np.asarray(direction)).astype(np.uint64)
        return ((np.asarray(roll).astype(np.uint64) << np.uint64(3)) | j) + np.uint64(1)

    def _probe(self, sel, start, seq):
        """Triangular probe over neighbour_list, model.py:826-839 / :1088-1101.

        Returns (found mask, direction found, new start, new sequence count)."""
        m = sel.size
        l = start.copy()
        seq = seq.copy()
        found = np.zeros(m, bool)
        active = np.ones(m, bool)
        for i in range(NDIR):
            seq = np.where(active, seq + 1, seq)                                # :827
            active &= ~(seq > NDIR)                                             # :828-830
            l = np.where(active, (l + i) % NDIR, l)                             # :831
            free = active & (self.nb_list[sel, l] == 0)                         # :832
            found |= free
            active &= ~free                                                     # :834-838
        return found, l, seq

    def _move_request(self, sub, first):
        """move_request, model.py:784-872 (condition :777-782)."""
        p = self.p
        self._partition(self.status == MOVEMENT_UNRESOLVED, sub)
        sel = np.flatnonzero(self.status == MOVEMENT_UNRESOLVED)
        self.claim_key = np.zeros(self.cells, np.uint64)                        # agent_move_request_msg
        if sel.size == 0:
            return
        lma = sub["last_move_attempt"][sel]
        fresh = lma > NDIR                                                      # :794
        dead = np.zeros(self.n, bool)
        if fresh.any():
            f = sel[fresh]
            if first:
                self.stats["movers"] += int(f.size)
            e = (self.energy[f] - np.float32(p.travel_cost)).astype(np.float32)   # :800
            dies = e <= np.float32(0)                                           # :801-804
            dead[f[dies]] = True
            self.energy[f] = np.where(dies, self.energy[f], e)                  # :805
            w0, w1, _, _ = self._rng(self.bucket[f], S_TRAVEL)
            self.roll[f] = w0                                                   # :806-807
            self.roll_cell[f] = self.bucket[f]
            lma = lma.copy()
            lma[fresh] = (w1 >> np.uint32(29)).astype(np.int64)                 # :810
        live = ~dead[sel]
        found, l, seq = self._probe(sel, lma, sub["move_sequence"][sel])
        found &= live
        nf = sel[live & ~found]
        self.status[nf] = READY                                                 # :842-846
        g = sel[found]
        lg = l[found]
        sub["last_move_attempt"][g] = (lg + 1) % NDIR                           # :848, :859
        sub["move_sequence"][g] = seq[found]                                    # :849
        self.status[g] = MOVING                                                 # :857
        rb = self._nb_bucket(g, lg)                                             # :861
        sub["request_bucket"][g] = rb                                           # :862
        # one message per requester keyed by the requested bucket (:864-868); the reader keeps
        # the maximum (roll, tag) so we fold the bucket's messages into that maximum here
        key = self._claim_key(self.roll[g], lg)
        np.maximum.at(self.claim_key, rb, key)
        nd = int(dead.sum())
        self.stats["travel_deaths"] += nd
        if nd:
            self._compact(~dead, sub)

    def _move_response(self, sub):
        """move_response, model.py:881-960 (condition :874-879)."""
        self._partition(self.status == MOVING, sub)
        sel = np.flatnonzero(self.status == MOVING)
        if sel.size == 0:
            return
        rb = sub["request_bucket"][sel]
        for i in range(NDIR):                                                   # :907
            nb = self._nb_bucket(sel, i)                                        # :908-909
            claimed = self.claim_key[nb] != 0
            # any claim on a neighbouring cell marks that slot occupied -- :922-925, :934
            self.nb_list[sel, i] = np.where(claimed, -1, self.nb_list[sel, i])
        mine = self._claim_key(self.roll[sel], (sub["last_move_attempt"][sel] - 1) % NDIR)
        won = self.claim_key[rb] == mine                                        # :929-935, :940
        self.status[sel[~won]] = MOVEMENT_UNRESOLVED                            # :941
        w = sel[won]
        self.x[w] = rb[won] % self.W                                            # :946-947
        self.y[w] = rb[won] // self.W
        self.bucket[w] = rb[won]                                                # :955
        self.status[w] = READY                                                  # :956
        self.stats["moved"] += int(w.size)

    # ----------------------------------------------- neighbourhood submodel (L5)
    def neighbourhood_model(self):
        """neighbourhood_broadcast :963-972, neighbourhood_update :975-1031; one iteration (:1591-1596)."""
        occ = np.zeros(self.cells, np.int64)
        occ[self.bucket] = self.id                                              # :968-969
        self._partition(self.energy >= np.float32(self.p.reproduce_min_energy))
        sel = np.flatnonzero(self.energy >= np.float32(self.p.reproduce_min_energy))   # :978-979
        if sel.size == 0:
            return
        num = np.zeros(sel.size, np.int64)
        for i in range(NDIR):                                                   # :1001
            nid = occ[self._nb_bucket(sel, i)]
            num += nid != 0                                                     # :1013-1016
            self.nb_list[sel, i] = nid                                          # :1020
        self.status[sel] = np.where(num < NDIR, ATTEMPTING_REPRODUCTION, REPRODUCTION_IMPOSSIBLE)  # :1024-1028

    # ----------------------------------------------------------- god submodel (L6)
    def god_model(self):
        """Submodel wiring model.py:2088-2137; exit_god_fn model.py:1607-1627."""
        n = self.n
        sub = {  # model.py:1719-1723
            "request_bucket": np.zeros(n, np.int64),
            "last_reproduction_attempt": np.full(n, NDIR + 1, np.int64),
            "reproduce_sequence": np.zeros(n, np.int64),
            "agents_spawned": np.zeros(n, np.int64),
        }
        overpopulated = 0                                                       # :1682 (submodel-local copy)
        iterations = 0
        self.stats.update(births=0, repro_rounds=0, repro_claims=0)
        while True:
            self._god_go_forth(sub, overpopulated)
            self._god_multiply(sub, overpopulated)
            overpopulated = 1 if self.n > self.p.max_agents else 0              # :1397-1402, :1616
            iterations += 1                                                     # :1617
            self.stats["repro_rounds"] = iterations
            if iterations < NDIR and np.any(self.status == ATTEMPTING_REPRODUCTION) \
                    and overpopulated < 1:                                      # :1618-1625
                continue
            break

    def _god_go_forth(self, sub, overpopulated):
        """god_go_forth, model.py:1042-1134 (condition :1033-1039)."""
        self.repro_key = np.zeros(self.cells, np.uint64)                        # god_go_forth_msg
        if overpopulated >= 1:
            return
        self._partition(self.status == ATTEMPTING_REPRODUCTION, sub)
        sel = np.flatnonzero(self.status == ATTEMPTING_REPRODUCTION)
        if sel.size == 0:
            return
        done = sub["agents_spawned"][sel] >= self.p.max_children_per_step       # :1054-1058
        self.status[sel[done]] = REPRODUCTION_COMPLETE
        sel = sel[~done]
        exhausted = sub["reproduce_sequence"][sel] >= NDIR                      # :1067-1071
        self.status[sel[exhausted]] = REPRODUCTION_IMPOSSIBLE
        sel = sel[~exhausted]
        if sel.size == 0:
            return
        lra = sub["last_reproduction_attempt"][sel].copy()
        fresh = lra > NDIR                                                      # :1075
        if fresh.any():
            f = sel[fresh]
            w0, w1, _, _ = self._rng(self.bucket[f], S_REPRO)
            self.roll[f] = w0                                                   # :1076-1077
            self.roll_cell[f] = self.bucket[f]
            lra[fresh] = (w1 >> np.uint32(29)).astype(np.int64)                 # :1079
        found, l, seq = self._probe(sel, lra, sub["reproduce_sequence"][sel])
        self.status[sel[~found]] = REPRODUCTION_IMPOSSIBLE                      # :1104-1108
        g = sel[found]
        lg = l[found]
        sub["last_reproduction_attempt"][g] = (lg + 1) % NDIR                   # :1110, :1121
        sub["reproduce_sequence"][g] = seq[found]                               # :1111
        self.status[g] = ATTEMPTING_REPRODUCTION                                # :1119
        rb = self._nb_bucket(g, lg)                                             # :1123
        sub["request_bucket"][g] = rb                                           # :1124
        key = self._claim_key(self.roll[g], lg)
        np.maximum.at(self.repro_key, rb, key)                                  # :1126-1130
        self.stats["repro_claims"] += int(g.size)

    def _mutate(self, parent_strat, words_roll, words_pick):
        """`while (child == my) child = uniform(0,3)` (model.py:1264-1266 etc.) restated as
        (my + 1 + pick3) & 3: uniform over the three other strategies, one draw."""
        return ((parent_strat.astype(np.int64) + 1 + pick3(words_pick).astype(np.int64)) & 3).astype(np.uint8)

    def _god_multiply(self, sub, overpopulated):
        """god_multiply, model.py:1144-1335 (condition :1135-1141) + agent output :2129."""
        p = self.p
        if overpopulated >= 1:
            return
        self._partition(self.status == ATTEMPTING_REPRODUCTION, sub)
        sel = np.flatnonzero(self.status == ATTEMPTING_REPRODUCTION)
        if sel.size == 0:
            return
        rb = sub["request_bucket"][sel]
        for i in range(NDIR):                                                   # :1170
            nb = self._nb_bucket(sel, i)
            claimed = self.repro_key[nb] != 0
            self.nb_list[sel, i] = np.where(claimed, -1, self.nb_list[sel, i])  # :1184-1187, :1197
        mine = self._claim_key(self.roll[sel], (sub["last_reproduction_attempt"][sel] - 1) % NDIR)
        won = self.repro_key[rb] == mine                                        # :1192-1198, :1203
        w = sel[won]
        if w.size == 0:
            return
        tgt = rb[won]
        e = (self.energy[w] - np.float32(p.reproduce_cost)).astype(np.float32)   # :1211
        self.energy[w] = e                                                      # :1212
        self.status[w] = REPRODUCTION_COMPLETE                                  # :1213, :1331
        sub["agents_spawned"][w] += 1                                           # :1329-1330
        my_trait = self.trait[w]                                                # :1224-1225
        inh = np.float32(p.reproduction_inheritence)
        if inh <= 0.0 or inh > 1.0:                                             # :1231
            z = normal_from_words(*self._rng(self.bucket[w], S_ENERGY))         # :1236
            child_e = (z * np.float32(p.init_energy_sigma)).astype(np.float32)  # :1237
            child_e = (child_e + np.float32(p.init_energy_mu)).astype(np.float32)   # :1238
        else:
            child_e = (inh * e).astype(np.float32)                              # :1242
        lo, hi = np.float32(p.init_energy_min), np.float32(p.max_energy)
        child_e = np.where(child_e < lo, lo, np.where(child_e > hi, hi, child_e)).astype(np.float32)  # :1245-1249
        ps = self.strat[w]
        m0, m1, m2, m3 = self._rng(self.bucket[w], S_MUT)
        k0, k1, k2, k3 = self._rng(self.bucket[w], S_PICK)
        thr = np.uint64(self.mut_thr)
        rate_pos = np.float32(p.mutation_rate) > 0.0
        cs = ps.copy()
        if p.strategy_pure == 1:                                                # :1259-1271
            s = ps[:, 0]
            if rate_pos:   # the roll is drawn but never compared: every child mutates (B-7)
                s = self._mutate(s, m0, k0)
            cs[:] = s[:, None]
        elif p.strategy_per_trait == 1:                                         # :1273-1286
            for i, (m, k) in enumerate(((m0, k0), (m1, k1), (m2, k2), (m3, k3))):
                mut = rate_pos & (m.astype(np.uint64) < thr)
                cs[:, i] = np.where(mut, self._mutate(ps[:, i], m, k), ps[:, i])
        else:                                                                   # :1287-1326
            rows = np.arange(w.size)
            first_other = np.where(my_trait == 0, 1, 0)
            s_my = ps[rows, my_trait]
            s_ot = ps[rows, first_other]
            mut_my = rate_pos & (m0.astype(np.uint64) < thr)                    # :1291, :1301
            mut_ot = rate_pos & (m1.astype(np.uint64) < thr)                    # :1292, :1312
            c_my = np.where(mut_my, self._mutate(s_my, m0, k0), s_my)
            c_ot = np.where(mut_ot, self._mutate(s_ot, m1, k1), s_ot)
            cs[:] = c_ot[:, None]
            cs[rows, my_trait] = c_my
        nb = w.size
        newid = np.arange(self.next_id, self.next_id + nb, dtype=np.int64)
        self.next_id += nb
        # append the newborns -- agent output, model.py:1216-1328
        self.id = np.concatenate([self.id, newid])
        self.x = np.concatenate([self.x, tgt % self.W])
        self.y = np.concatenate([self.y, tgt // self.W])
        self.energy = np.concatenate([self.energy, child_e])
        self.status = np.concatenate([self.status, np.full(nb, NEW_AGENT, np.int64)])   # :1328
        self.trait = np.concatenate([self.trait, my_trait])
        self.strat = np.concatenate([self.strat, cs])
        self.nb_list = np.concatenate([self.nb_list, np.zeros((nb, NDIR), np.int64)])
        self.act = np.concatenate([self.act, np.full((nb, NDIR), -1, np.int8)])
        self.roll = np.concatenate([self.roll, np.zeros(nb, np.uint32)])
        self.roll_cell = np.concatenate([self.roll_cell, np.zeros(nb, np.int64)])
        self.coins = np.concatenate([self.coins, np.zeros(nb, np.uint32)])
        # agent_strategy_id of the child: set in pure (:1271) and kin/other (:1326) mode, NOT in per-trait mode
        # (:1273-1286), where it keeps the default 0 = bin 0 (B-8)
        child_bin = (np.zeros(nb, np.int64) if (p.strategy_per_trait == 1 and p.strategy_pure != 1)
                     else self._derived_bin(cs, my_trait))
        self.sid_bin = np.concatenate([self.sid_bin, child_bin])
        self.bucket = np.concatenate([self.bucket, tgt])                        # :1252
        for k, dflt in (("request_bucket", 0), ("last_reproduction_attempt", NDIR + 1),
                        ("reproduce_sequence", 0), ("agents_spawned", 0)):
            sub[k] = np.concatenate([sub[k], np.full(nb, dflt, np.int64)])
        self.stats["births"] += nb

    # ---------------------------------------------------------- main layer 7 (K12)
    def environmental_punishment(self):
        """environmental_punishment, model.py:1339-1371, with the priority cull (B-4)."""
        p = self.p
        if p.cull == "list_index":
            return self._environmental_punishment_list_index()
        newborn = self.status == NEW_AGENT
        n_old = int((~newborn).sum())
        n_new = int(newborn.sum())
        alive = np.ones(self.n, bool)
        culled = 0
        if n_old + n_new > p.max_agents:                                        # :1343, :1354
            keep = max(0, p.max_agents - n_old)
            nb_idx = np.flatnonzero(newborn)
            w0, _, _, _ = self._rng(self.bucket[nb_idx], S_CULL)
            key = (w0.astype(np.uint64) << np.uint64(32)) | self.bucket[nb_idx].astype(np.uint64)
            order = np.argsort(key)[::-1]                                       # highest priority first
            alive[nb_idx[order[keep:]]] = False
            culled = n_new - keep
        old = np.flatnonzero(~newborn)
        e = np.minimum(self.energy[old], np.float32(p.max_energy))              # :1360-1362
        e = (e - np.float32(p.cost_of_living)).astype(np.float32)               # :1363
        dies = e <= np.float32(0)                                               # :1364-1366
        self.energy[old] = np.where(dies, self.energy[old], e)                  # :1367
        self.status[old] = READY                                                # :1368
        alive[old[dies]] = False
        self.stats["culled"] = culled
        self.stats["living_deaths"] = int(dies.sum())
        self._compact(alive)

    def _environmental_punishment_list_index(self):
        """environmental_punishment with the reference's OWN cull (model.py:1339-1371) under the list
        semantics of SURVEY Appendix D-2 [FG2-ext]: the condition (:1343) sees the index in the whole
        list, the function (:1354) the index within its working set (the agents that passed)."""
        p = self.p
        idx = np.arange(self.n)
        passing = (self.status != NEW_AGENT) | (idx >= p.max_agents)            # :1343
        self._partition(passing)
        ws = np.arange(self.n - int(passing.sum()), self.n)                     # the working set, in list order
        ws_index = np.arange(ws.size)
        cull = ws_index >= p.max_agents                                         # :1354-1356
        e = np.minimum(self.energy[ws], np.float32(p.max_energy))              # :1360-1362
        e = (e - np.float32(p.cost_of_living)).astype(np.float32)               # :1363
        dies = ~cull & (e <= np.float32(0))                                     # :1364-1366
        keep = ~cull & ~dies
        self.energy[ws[keep]] = e[keep]                                         # :1367
        self.status[ws[keep]] = READY                                           # :1368
        alive = np.ones(self.n, bool)
        alive[ws[cull | dies]] = False
        self.stats["culled"] = int(cull.sum())
        self.stats["living_deaths"] = int(dies.sum())
        self._compact(alive)

    # ------------------------------------------------------------------ step_fn
    @staticmethod
    def _derived_bin(strat, trait) -> np.ndarray:
        """4*my + other of the strategies an agent holds (:1501-1511)."""
        rows = np.arange(trait.shape[0])
        my = strat[rows, trait].astype(np.int64)
        other = strat[rows, np.where(trait == 0, 1, 0)].astype(np.int64)       # :1501-1508
        return 4 * my + other

    def strategy_counts(self) -> np.ndarray:
        """step_fn, model.py:1405-1419: bin k = 4*my + other, derived from strategies+trait (B-8) -- or, with
        legacy_counts, the bin of the agent_strategy_id variable as the reference maintains it."""
        if self.n == 0:
            return np.zeros(16, np.int64)
        bins = self.sid_bin if self.p.legacy_counts else self._derived_bin(self.strat, self.trait)
        return np.bincount(bins, minlength=16).astype(np.int64)

    # ------------------------------------------------------------------ one step
    def step(self) -> Dict[str, int]:
        """Main layers 1-7, model.py:2140-2163, then step_fn (:1405-1419)."""
        self.stats = {"agents_start": self.n}
        self.search_for_neighbours()        # L1
        self.get_game_list()                # L2
        self.pdgame_model()                 # L3
        self.movement_model()               # L4
        self.neighbourhood_model()          # L5
        self.god_model()                    # L6
        self.environmental_punishment()     # L7
        self.population_strat_count = self.strategy_counts()
        self.stats["agents_end"] = self.n
        self.step_index += 1
        return dict(self.stats)

    # ------------------------------------------------------------------ planes
    def planes(self) -> Dict[str, np.ndarray]:
        """Agent list -> dense grid planes [y, x] for comparison with the CUDA layout."""
        W = self.W
        b = self.x + self.y * W
        assert np.unique(b).size == b.size, "two agents in one cell"
        occ = np.zeros(self.cells, np.uint8)
        energy = np.zeros(self.cells, np.float32)
        trait = np.zeros(self.cells, np.uint8)
        strat = np.zeros((self.cells, 4), np.uint8)
        occ[b] = 1
        energy[b] = self.energy
        trait[b] = self.trait
        strat[b] = self.strat
        return {"occ": occ.reshape(W, W), "energy": energy.reshape(W, W),
                "trait": trait.reshape(W, W), "strat": strat.reshape(W, W, 4)}

"""Literal per-agent transcription of the reference's agent functions -- SMALL CASES ONLY.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Where oracle/pd_oracle.py vectorises each
agent function over a NumPy agent list, this file keeps the reference's own shape: one Python
dict per agent, one Python list of message dicts per bucket, the reference's statements in the
reference's order (/root/reference/src/model.py line numbers in the comments).  It is slow
(pure-Python loops) and exists to pin the vectorised oracle on small worlds
(tests/test_oracle_literal.py).  The three restatements are the same as in pd_oracle.py:
Philox-keyed draws, cell-index tie-breaks, priority cull.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np

from .philox import philox_scalar
from .pd_oracle import (Params, prob_threshold, NORMAL_SCALE, S_ROLL, S_NOISE0, S_TRAVEL, ROLL_SHIFT, S_REPRO, S_MUT, S_PICK,
                        S_CULL, S_ENERGY, READY, READY_TO_CHALLENGE, READY_TO_RESPOND,
                        MOVEMENT_UNRESOLVED, MOVING, ATTEMPTING_REPRODUCTION, REPRODUCTION_IMPOSSIBLE,
                        REPRODUCTION_COMPLETE, NEW_AGENT)

X_OFF = [-1, -1, -1, 0, 0, 1, 1, 1]   # model.py:302-304
Y_OFF = [-1, 0, 1, -1, 1, -1, 0, 1]   # model.py:305-307
f32 = np.float32


def _normal(w):
    pc = bin(w[0]).count("1") + bin(w[1]).count("1") + bin(w[2]).count("1") - 48
    zi = pc * (1 << 32) + (w[3] - (1 << 31))
    return f32(float(zi) * NORMAL_SCALE)


class LiteralModel:
    def __init__(self, params: Params):
        self.p = params
        self.W = params.width
        self.key = (params.seed & 0xFFFFFFFF, (params.seed >> 32) & 0xFFFFFFFF)
        self.step_index = 0
        self.agents = []          # list of dicts, in FLAMEGPU2 list order
        self.next_id = 1
        self.noise_thr = prob_threshold(params.env_noise)
        self.mut_thr = prob_threshold(params.mutation_rate)

    # ---------------------------------------------------------------- helpers
    def rng(self, cell, stream):
        return philox_scalar((cell & 0xFFFFFFFF, cell >> 32, self.step_index, stream), self.key)

    def pos_from_moore_seq(self, x, y, i):          # model.py:294-332
        return (x + X_OFF[i]) % self.W, (y + Y_OFF[i]) % self.W

    def pos_to_bucket_id(self, x, y):               # model.py:334-342
        return x + y * self.W

    def load(self, oracle):
        """Copy the agent list of a pd_oracle.Oracle (same order, same ids)."""
        self.agents = []
        for k in range(oracle.n):
            self.agents.append(dict(id=int(oracle.id[k]), x=int(oracle.x[k]), y=int(oracle.y[k]),
                                    energy=f32(oracle.energy[k]), status=READY, trait=int(oracle.trait[k]),
                                    strat=[int(v) for v in oracle.strat[k]], nb=[0] * 8, act=[-1] * 8,
                                    roll=0, roll_cell=0, bucket=0))
        self.next_id = oracle.next_id
        self.step_index = oracle.step_index

    # ---------------------------------------------------------------- main L1/L2
    def search_and_game_list(self):
        msgs = {}
        for a in self.agents:                                       # search_for_neighbours :352-372
            a["bucket"] = self.pos_to_bucket_id(a["x"], a["y"])
            w = self.rng(a["bucket"] >> 2, S_ROLL)[a["bucket"] & 3]   # one call per four cells in a row
            a["roll"] = w >> ROLL_SHIFT                             # :364 (top 16 bits)
            a["coins"] = w & 0xFFFF                                 # the coins of :631 / :664
            a["roll_cell"] = a["bucket"]
            msgs[a["bucket"]] = (a["id"], a["roll"])
        for a in self.agents:                                       # get_game_list :373-440
            n = 0
            for i in range(8):
                nx, ny = self.pos_from_moore_seq(a["x"], a["y"], i)
                nb = self.pos_to_bucket_id(nx, ny)
                nid, action = 0, -1
                if nb in msgs:
                    nid, nroll = msgs[nb]
                    n += 1
                    if a["roll"] > nroll or (nroll == a["roll"] and i < 4):   # :413 (tie: direction class)
                        action = 1
                    else:
                        action = 0
                a["nb"][i] = nid
                a["act"][i] = action
            a["status"] = MOVEMENT_UNRESOLVED if n == 0 else READY_TO_CHALLENGE     # :429-436

    # ---------------------------------------------------------------- pdgame submodel
    def pdgame(self):
        p = self.p
        for a in self.agents:                                       # submodel-local defaults :1690-1696
            a["cs"] = 0
            a["games"] = 0
        iterations = 0
        while True:
            bucket_msgs = defaultdict(list)
            for a in self.agents:                                   # play_challenge :450-566
                if a["status"] != READY_TO_CHALLENGE:
                    continue
                cs = a["cs"]
                rs_seq = 8 - cs - 1                                 # :477
                a["rs_seq"] = rs_seq
                my_challenge = a["act"][cs] == 1                    # :482
                my_response = a["act"][rs_seq] == 0                 # :483
                if not my_challenge and not my_response:            # :486-496
                    a["cs"] = cs + 1
                    a["status"] = READY_TO_CHALLENGE
                    continue
                elif not my_challenge and my_response:              # :497-507
                    a["cs"] = cs + 1
                    a["status"] = READY_TO_RESPOND
                    continue
                if my_challenge:                                    # :511-537
                    nx, ny = self.pos_from_moore_seq(a["x"], a["y"], cs)
                    bucket_msgs[self.pos_to_bucket_id(nx, ny)].append(
                        dict(challenger_id=a["id"], responder_id=a["nb"][cs], strat=list(a["strat"]),
                             trait=a["trait"]))
                a["cs"] = cs + 1                                    # :539
                if my_response:                                     # :541-542
                    a["status"] = READY_TO_RESPOND
                elif a["cs"] < 8:                                   # :544-545
                    a["status"] = READY_TO_CHALLENGE
                else:                                               # :547-556
                    a["status"] = MOVEMENT_UNRESOLVED if a["games"] < 1 else READY
            survivors = []
            for a in self.agents:                                   # play_response :575-734
                if a["status"] != READY_TO_RESPOND:
                    survivors.append(a)
                    continue
                dead = False
                for m in bucket_msgs.get(a["bucket"], []):
                    if m["responder_id"] != a["id"]:                # :586-587
                        continue
                    my_strategy = a["strat"][m["trait"]]            # :599
                    ch_strategy = m["strat"][a["trait"]]            # :600
                    c = a["cs"] - 1                                 # the sub-step being played
                    coin_ch = (a["coins"] >> (2 * c)) & 1
                    coin_me = (a["coins"] >> (2 * c + 1)) & 1
                    my_roll = ch_roll = 0xFFFFFFFF
                    if self.noise_thr:                              # :611-612
                        w = self.rng(a["bucket"], S_NOISE0 + (c >> 1))
                        my_roll, ch_roll = (w[2], w[3]) if c & 1 else (w[0], w[1])
                    if ch_strategy == 0:
                        ch_coop = True
                    elif ch_strategy == 1:
                        ch_coop = False
                    elif ch_strategy == 2:
                        ch_coop = True                              # :621-629 memory never matches
                    else:
                        ch_coop = coin_ch == 1                      # :631
                    if ch_roll < self.noise_thr:                    # :638-640
                        ch_coop = not ch_coop
                    if my_strategy == 0:
                        i_coop = True
                    elif my_strategy == 1:
                        i_coop = False
                    elif my_strategy == 2:
                        i_coop = True                               # :646-661
                    else:
                        i_coop = coin_me == 1                       # :664
                    if my_roll < self.noise_thr:                    # :672-674
                        i_coop = not i_coop
                    e = a["energy"]
                    if i_coop and ch_coop:                          # :677-689
                        e = f32(e + f32(p.payoff_cc))
                    elif not i_coop and not ch_coop:
                        e = f32(e + f32(p.payoff_dd))
                    elif i_coop and not ch_coop:
                        e = f32(e + f32(p.payoff_cd))
                    else:
                        e = f32(e + f32(p.payoff_dc))
                    if e <= 0:                                      # :695-697
                        dead = True
                        break
                    if e > f32(p.max_energy):                       # :698-701
                        e = f32(p.max_energy)
                    a["energy"] = e                                 # :702
                    a["games"] += 1                                 # :710
                    break
                if dead:
                    continue
                if a["cs"] < 8:                                     # :715-717
                    a["status"] = READY_TO_CHALLENGE
                else:                                               # :718-730
                    a["status"] = MOVEMENT_UNRESOLVED if a["games"] < 1 else READY
                survivors.append(a)
            self.agents = survivors
            iterations += 1                                         # exit_play_fn :1534-1556
            if iterations < 8 and any(a["status"] in (READY_TO_CHALLENGE, READY_TO_RESPOND) for a in self.agents):
                continue
            break

    # ---------------------------------------------------------------- shared probe
    @staticmethod
    def _probe(nb, last_attempt, sequence):
        """:826-839 / :1088-1101; returns (found_dir or None, last_attempt, sequence)."""
        found = None
        for i in range(8):
            sequence += 1
            if sequence > 8:
                break
            last_attempt = (last_attempt + i) % 8
            if nb[last_attempt] == 0:
                found = last_attempt
                break
        return found, last_attempt, sequence

    # ---------------------------------------------------------------- movement submodel
    def movement(self):
        p = self.p
        for a in self.agents:                                       # :1708-1711
            a["lma"] = 9
            a["mseq"] = 0
            a["req"] = 0
        iterations = 0
        while True:
            msgs = defaultdict(list)
            survivors = []
            for a in self.agents:                                   # move_request :784-872
                if a["status"] != MOVEMENT_UNRESOLVED:
                    survivors.append(a)
                    continue
                if a["lma"] > 8:                                    # :794
                    e = f32(a["energy"] - f32(p.travel_cost))       # :800
                    if e <= 0.0:                                    # :801-804
                        continue
                    a["energy"] = e
                    w = self.rng(a["bucket"], S_TRAVEL)
                    a["roll"] = w[0]                                # :806
                    a["roll_cell"] = a["bucket"]
                    a["lma"] = w[1] >> 29                           # :810
                found, lma, mseq = self._probe(a["nb"], a["lma"], a["mseq"])
                survivors.append(a)
                if found is None:                                   # :842-846
                    a["status"] = READY
                    continue
                a["lma"] = (lma + 1) % 8                            # :848
                a["mseq"] = mseq                                    # :849
                a["status"] = MOVING                                # :857
                nx, ny = self.pos_from_moore_seq(a["x"], a["y"], found)
                a["req"] = self.pos_to_bucket_id(nx, ny)            # :861-862
                msgs[a["req"]].append(dict(id=a["id"], roll=a["roll"], tag=7 - found, x=nx, y=ny))
            self.agents = survivors
            for a in self.agents:                                   # move_response :881-960
                if a["status"] != MOVING:
                    continue
                best = None
                for i in range(8):                                  # :907
                    nx, ny = self.pos_from_moore_seq(a["x"], a["y"], i)
                    nb = self.pos_to_bucket_id(nx, ny)
                    for m in msgs.get(nb, []):
                        if nb != a["req"]:                          # :922-925
                            a["nb"][i] = m["id"]
                            break
                        if best is None or (m["roll"], m["tag"]) > (best["roll"], best["tag"]):   # :929
                            best = m
                            a["nb"][i] = m["id"]                    # :934
                if best is None or best["id"] != a["id"]:           # :940-943
                    a["status"] = MOVEMENT_UNRESOLVED
                    continue
                a["x"], a["y"] = best["x"], best["y"]               # :946-947
                a["bucket"] = a["req"]                              # :955
                a["status"] = READY                                 # :956
            iterations += 1                                         # exit_move_fn :1566-1576
            if iterations < 8 and any(a["status"] == MOVEMENT_UNRESOLVED for a in self.agents):
                continue
            break

    # ---------------------------------------------------------------- neighbourhood submodel
    def neighbourhood(self):
        occ = {a["bucket"]: a["id"] for a in self.agents}           # :963-972
        for a in self.agents:
            if not a["energy"] >= f32(self.p.reproduce_min_energy):  # :975-981
                continue
            n = 0
            for i in range(8):                                      # :1001-1021
                nx, ny = self.pos_from_moore_seq(a["x"], a["y"], i)
                nid = occ.get(self.pos_to_bucket_id(nx, ny), 0)
                if nid:
                    n += 1
                a["nb"][i] = nid
            a["status"] = ATTEMPTING_REPRODUCTION if n < 8 else REPRODUCTION_IMPOSSIBLE   # :1024-1028

    # ---------------------------------------------------------------- god submodel
    def _mut(self, s, pick_word):
        return (s + 1 + ((pick_word * 3) >> 32)) & 3

    def god(self):
        p = self.p
        for a in self.agents:                                       # :1719-1723
            a["lra"] = 9
            a["rseq"] = 0
            a["spawned"] = 0
            a["req"] = 0
        overpopulated = 0
        iterations = 0
        while True:
            msgs = defaultdict(list)
            active = [a for a in self.agents if overpopulated < 1 and a["status"] == ATTEMPTING_REPRODUCTION]
            for a in active:                                        # god_go_forth :1042-1134
                if a["spawned"] >= p.max_children_per_step:         # :1054-1058
                    a["status"] = REPRODUCTION_COMPLETE
                    continue
                if a["rseq"] >= 8:                                  # :1067-1071
                    a["status"] = REPRODUCTION_IMPOSSIBLE
                    continue
                if a["lra"] > 8:                                    # :1075-1080
                    w = self.rng(a["bucket"], S_REPRO)
                    a["roll"] = w[0]
                    a["roll_cell"] = a["bucket"]
                    a["lra"] = w[1] >> 29
                found, lra, rseq = self._probe(a["nb"], a["lra"], a["rseq"])
                if found is None:                                   # :1104-1108
                    a["status"] = REPRODUCTION_IMPOSSIBLE
                    continue
                a["lra"] = (lra + 1) % 8                            # :1110
                a["rseq"] = rseq                                    # :1111
                nx, ny = self.pos_from_moore_seq(a["x"], a["y"], found)
                a["req"] = self.pos_to_bucket_id(nx, ny)
                msgs[a["req"]].append(dict(id=a["id"], roll=a["roll"], tag=7 - found, x=nx, y=ny))
            newborn = []
            for a in self.agents:                                   # god_multiply :1144-1335
                if not (overpopulated < 1 and a["status"] == ATTEMPTING_REPRODUCTION):
                    continue
                best = None
                for i in range(8):
                    nx, ny = self.pos_from_moore_seq(a["x"], a["y"], i)
                    nb = self.pos_to_bucket_id(nx, ny)
                    for m in msgs.get(nb, []):
                        if nb != a["req"]:                          # :1184-1188
                            a["nb"][i] = m["id"]
                            break
                        if best is None or (m["roll"], m["tag"]) > (best["roll"], best["tag"]):   # :1192
                            best = m
                            a["nb"][i] = m["id"]
                if best is None or best["id"] != a["id"]:           # :1203-1205
                    continue
                e = f32(a["energy"] - f32(p.reproduce_cost))        # :1211
                a["energy"] = e
                inh = f32(p.reproduction_inheritence)
                if inh <= 0.0 or inh > 1.0:                         # :1231-1238
                    ce = f32(_normal(self.rng(a["bucket"], S_ENERGY)) * f32(p.init_energy_sigma))
                    ce = f32(ce + f32(p.init_energy_mu))
                else:
                    ce = f32(inh * e)                               # :1242
                if ce < f32(p.init_energy_min):                     # :1245-1249
                    ce = f32(p.init_energy_min)
                elif ce > f32(p.max_energy):
                    ce = f32(p.max_energy)
                m_w = self.rng(a["bucket"], S_MUT)
                k_w = self.rng(a["bucket"], S_PICK)
                rate_pos = f32(p.mutation_rate) > 0.0
                ps = a["strat"]
                cs = list(ps)
                if p.strategy_pure == 1:                            # :1259-1271
                    s = ps[0]
                    if rate_pos:
                        s = self._mut(s, k_w[0])
                    cs = [s] * 4
                elif p.strategy_per_trait == 1:                     # :1273-1286
                    for i in range(4):
                        if rate_pos and m_w[i] < self.mut_thr:
                            cs[i] = self._mut(ps[i], k_w[i])
                else:                                               # :1287-1326
                    c_my = c_ot = None
                    for i in range(4):
                        if i == a["trait"]:
                            if c_my is None:
                                c_my = ps[i]
                                if rate_pos and m_w[0] < self.mut_thr:
                                    c_my = self._mut(ps[i], k_w[0])
                            cs[i] = c_my
                        else:
                            if c_ot is None:
                                c_ot = ps[i]
                                if rate_pos and m_w[1] < self.mut_thr:
                                    c_ot = self._mut(ps[i], k_w[1])
                            cs[i] = c_ot
                newborn.append(dict(id=self.next_id, x=best["x"], y=best["y"], energy=ce, status=NEW_AGENT,
                                    trait=a["trait"], strat=cs, nb=[0] * 8, act=[-1] * 8, roll=0, roll_cell=0,
                                    bucket=a["req"], lra=9, rseq=0, spawned=0, req=0))
                self.next_id += 1
                a["spawned"] += 1                                   # :1329-1330
                a["status"] = REPRODUCTION_COMPLETE                 # :1331
            self.agents.extend(newborn)
            overpopulated = 1 if len(self.agents) > p.max_agents else 0     # exit_god_fn :1614-1627
            iterations += 1
            if iterations < 8 and overpopulated < 1 and any(
                    a["status"] == ATTEMPTING_REPRODUCTION for a in self.agents):
                continue
            break

    # ---------------------------------------------------------------- main L7
    def punishment(self):
        p = self.p
        olds = [a for a in self.agents if a["status"] != NEW_AGENT]
        news = [a for a in self.agents if a["status"] == NEW_AGENT]
        if len(olds) + len(news) > p.max_agents:                    # priority cull (B-4)
            keep = max(0, p.max_agents - len(olds))
            news.sort(key=lambda a: (self.rng(a["bucket"], S_CULL)[0], a["bucket"]), reverse=True)
            news = news[:keep]
        out = []
        for a in olds:                                              # :1357-1368
            e = a["energy"]
            if e > f32(p.max_energy):
                e = f32(p.max_energy)
            e = f32(e - f32(p.cost_of_living))
            if e <= 0:
                continue
            a["energy"] = e
            a["status"] = READY
            out.append(a)
        self.agents = out + news

    def step(self):
        self.search_and_game_list()
        self.pdgame()
        self.movement()
        self.neighbourhood()
        self.god()
        self.punishment()
        self.step_index += 1

    def planes(self):
        W = self.W
        occ = np.zeros((W, W), np.uint8)
        energy = np.zeros((W, W), np.float32)
        trait = np.zeros((W, W), np.uint8)
        strat = np.zeros((W, W, 4), np.uint8)
        for a in self.agents:
            assert occ[a["y"], a["x"]] == 0, "two agents in one cell"
            occ[a["y"], a["x"]] = 1
            energy[a["y"], a["x"]] = a["energy"]
            trait[a["y"], a["x"]] = a["trait"]
            strat[a["y"], a["x"]] = a["strat"]
        return {"occ": occ, "energy": energy, "trait": trait, "strat": strat}

"""Decomposition invariance: one world stepped as 2 or 4 row slabs (ghost zones,
halo exchange, global gate and cull through the slab API) must stay bit-identical to the same world
stepped by a single handle.  LocalComm keeps all slabs on one GPU, so this needs only one device;
the NCCL transport is exercised by tests/test_decomp_nccl.py on a multi-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _require_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


CASES = {
    "defaults": (dict(width=256, seed=31), 75),
    "per_trait_noise_mut": (dict(width=256, seed=32, strategy_per_trait=1, env_noise=0.05, mutation_rate=0.01), 75),
    "harsh": (dict(width=256, seed=33, payoff_cd=-30.0, payoff_dd=-5.0, cost_of_living=3.0, travel_cost=4.0,
                   init_agent_count=30000), 25),
    "dense_conflicts": (dict(width=256, seed=34, init_agent_count=45000, max_agents=52000, reproduce_min_energy=40.0,
                             reproduce_cost=10.0, payoff_cc=10.0), 25),
    "dying_parents": (dict(width=256, seed=36, reproduce_min_energy=1.5, reproduce_cost=0.5, cost_of_living=2.0,
                           init_energy_mu=4.0, init_energy_sigma=2.0, init_energy_min=1.0, init_agent_count=24000), 30),
    "select_fallback": (dict(width=256, seed=37, debug_select_cap=1, init_agent_count=30000, max_agents=31000,
                             reproduce_min_energy=40.0, reproduce_cost=10.0), 30),
    "claim_detour": (dict(width=256, seed=38, debug_tie_shift=26, init_agent_count=30000, max_agents=40000,
                          reproduce_min_energy=40.0, reproduce_cost=10.0), 30),
    "inherit": (dict(width=256, seed=35, reproduction_inheritence=0.5, env_noise=0.2, mutation_rate=0.2), 70),
}


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("name", sorted(CASES))
def test_slabs_equal_whole_world(name, world):
    from pdabm_b200 import Simulation, SimConfig, DecomposedWorld, LocalComm
    kw, steps = CASES[name]
    ref = Simulation(SimConfig(**kw))
    dec = DecomposedWorld(SimConfig(**kw), LocalComm(world))
    try:
        ref.init_population()
        dec.init_population()
        for s in range(steps + 1):
            if s:
                ref.step(1)
                dec.step(1)
            a, b = ref.planes(), dec.planes()
            for k in ("occ", "trait", "strat", "energy", "tf_raw"):
                if not np.array_equal(a[k], b[k]):
                    bad = np.argwhere(a[k] != b[k])
                    raise AssertionError(f"{name} world={world} step {s}: plane {k} differs at {len(bad)} cells, "
                                         f"first {bad[:5].tolist()}")
            ca, na = ref.counts()
            cb, nb = dec.counts()
            assert na == nb and np.array_equal(ca, cb), f"{name} world={world} step {s}: counts differ"
    finally:
        ref.close()
        dec.close()


def test_non_square_world_slabs():
    """width x (2*width) world (used by the weak-scaling bench): whole handle vs 2 and 4 slabs."""
    from pdabm_b200 import Simulation, SimConfig, DecomposedWorld, LocalComm
    kw = dict(width=128, height=512, seed=8, strategy_per_trait=1, env_noise=0.05, mutation_rate=0.01)
    ref = Simulation(SimConfig(**kw))
    decs = [DecomposedWorld(SimConfig(**kw), LocalComm(w)) for w in (2, 4)]
    try:
        ref.init_population()
        for d in decs:
            d.init_population()
        for s in range(70):
            ref.step(1)
            a = ref.planes()
            for d in decs:
                d.step(1)
                b = d.planes()
                for k in ("occ", "trait", "strat", "energy"):
                    assert np.array_equal(a[k], b[k]), f"step {s}: plane {k} differs ({d.comm.world} slabs)"
                assert ref.counts()[1] == d.counts()[1]
    finally:
        ref.close()
        for d in decs:
            d.close()


def test_full_size_2p28_slabs_equal_whole_world():
    """BASELINE configs[2] at its full size (16384^2 = 2^28 cells, per-trait strategies), far beyond what
    the oracle can follow: the size-independent properties are (a) the decomposed world (4 slabs, halo
    exchange, global gate and cull) stays bit-identical to the whole world through growth, saturation and
    the first culls, (b) the population bookkeeping closes every step, (c) one agent per cell, energies in
    (0, max_energy], counts sum to the population, cap respected."""
    from pdabm_b200 import Simulation, SimConfig, DecomposedWorld, LocalComm
    kw = dict(width=16384, seed=12345, strategy_per_trait=1, collect_stats=True)
    ref = Simulation(SimConfig(**kw))
    dec = DecomposedWorld(SimConfig(**kw), LocalComm(4))
    try:
        ref.init_population()
        dec.init_population()
        culled = 0
        for s in range(66):
            ref.step(1)
            dec.step(1)
            st = ref.stats()
            ca, na = ref.counts()
            cb, nb = dec.counts()
            assert na == nb and np.array_equal(ca, cb), f"step {s}: counts differ"
            assert st["agents_end"] == na == ca.sum()
            assert st["agents_end"] == (st["agents_start"] - st["play_deaths"] - st["travel_deaths"]
                                        + st["births"] - st["culled"] - st["living_deaths"])
            assert na <= ref.cfg.max_agents
            culled += st["culled"]
        assert culled > 0, "the run must reach the carrying capacity"
        a, b = ref.planes(), dec.planes()
        for k in ("occ", "trait", "strat", "energy"):
            assert np.array_equal(a[k], b[k]), f"plane {k} differs"
        occ = a["occ"].astype(bool)
        assert int(occ.sum()) == na
        e = a["energy"][occ]
        assert e.min() > 0 and e.max() <= ref.cfg.max_energy
    finally:
        ref.close()
        dec.close()


def test_halo_strips_equal_plane_copies():
    """pd_halo_pack / pd_halo_unpack (the peer-memory transport's two kernels) fill the ghost rows exactly like
    LocalComm's plane-to-plane copies: pack every slab's boundary rows into strips, unpack the neighbours' strips."""
    import ctypes as C
    import torch
    from pdabm_b200 import SimConfig, DecomposedWorld, LocalComm
    from pdabm_b200.dist import ring_neighbours
    cfg = dict(width=256, seed=41, strategy_per_trait=1)
    a = DecomposedWorld(SimConfig(**cfg), LocalComm(4))
    b = DecomposedWorld(SimConfig(**cfg), LocalComm(4))
    try:
        a.init_population()
        b.init_population()
        a.step(3)
        b.step(3)
        a.comm.exchange(a.slabs)          # reference: tensor copies
        strips = []
        for s in b.slabs:                 # pack ...
            nb = C.c_uint64()
            s.sim._check(s.sim.lib.pd_halo_bytes(s.sim._h, C.byref(nb)))
            assert nb.value == 64 * 256 * 6
            box = torch.zeros(2 * nb.value, dtype=torch.uint8, device=s.sim.device)
            s.sim._check(s.sim.lib.pd_halo_pack(s.sim._h, box.data_ptr(), box.data_ptr() + nb.value, s.sim.stream_ptr))
            strips.append((box, nb.value))
        for s in b.slabs:                 # ... and unpack the neighbours' strips
            up, down = ring_neighbours(s.rank, 4)
            n = strips[0][1]
            s.sim._check(s.sim.lib.pd_halo_unpack(s.sim._h, strips[up][0].data_ptr() + n, strips[down][0].data_ptr(),
                                                  s.sim.stream_ptr))
        torch.cuda.synchronize()
        for sa, sb in zip(a.slabs, b.slabs):
            for name in ("energy", "strat", "tf"):
                assert torch.equal(getattr(sa.sim, name), getattr(sb.sim, name)), (sa.rank, name)
        a.step(5)                         # and the worlds stay in step afterwards
        b.step(5)
        pa, pb = a.planes(), b.planes()
        assert all(np.array_equal(pa[k], pb[k]) for k in pa)
    finally:
        a.close()
        b.close()


def test_speculative_finish_needs_confirmation():
    """PD_PART_CAND / PD_PART_APPLY with arg = 1 leave the host-side step counter alone until PD_PART_DONE; a
    non-slab handle rejects the slab API."""
    from pdabm_b200 import Simulation, SimConfig, PdabmError
    from pdabm_b200._lib import PD_PART_PRE
    sim = Simulation(SimConfig(width=64, seed=3))
    try:
        sim.init_population()
        with pytest.raises(PdabmError):
            sim._check(sim.lib.pd_step_part(sim._h, PD_PART_PRE, 0, sim.stream_ptr, 0))
    finally:
        sim.close()

"""GPU parity: the CUDA pipeline (through the C ABI) against the NumPy oracle, step by step.

Integer state (occupancy, trait, strategies, counts, event counters) must be bit-exact;
energies must agree within 1e-5 relative (north star) -- in practice they are bit-exact too
and the test reports which.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module", autouse=True)
def _require_gpu():
    if not _have_gpu():
        pytest.skip("no CUDA device")


def _pair(**kw):
    from pdabm_b200 import Simulation, SimConfig
    from oracle.pd_oracle import Oracle, Params
    sim = Simulation(SimConfig(collect_stats=True, **kw))
    okw = {k: v for k, v in kw.items() if k not in ("sparse_ctas", "debug_select_cap", "debug_tie_shift")}
    orc = Oracle(Params(**okw))
    return sim, orc


def _compare(sim, orc, tag):
    got, ref = sim.planes(), orc.planes()
    for k in ("occ", "trait", "strat"):
        if not np.array_equal(got[k], ref[k]):
            bad = np.argwhere(got[k] != ref[k])
            raise AssertionError(f"{tag}: plane {k} differs at {len(bad)} cells, first {bad[:5].tolist()}")
    assert np.all((got["tf_raw"] & 0x7C) == 0), f"{tag}: transient tf bits left set"
    e_exact = np.array_equal(got["energy"], ref["energy"])
    if not e_exact:
        np.testing.assert_allclose(got["energy"], ref["energy"], rtol=1e-5, atol=0, err_msg=tag)
    counts, n = sim.counts()
    assert n == orc.n, f"{tag}: agent count {n} != {orc.n}"
    assert np.array_equal(counts, orc.population_strat_count), f"{tag}: strategy counts differ"
    return e_exact


def _run(steps, **kw):
    sim, orc = _pair(**kw)
    try:
        sim.init_population()
        orc.init_population()
        exact = _compare(sim, orc, "init")
        for s in range(steps):
            sim.step(1)
            st = orc.step()
            exact &= _compare(sim, orc, f"step {s}")
            gs = sim.stats()
            for k_gpu, k_or in (("play_deaths", "play_deaths"), ("movers", "movers"), ("moved", "moved"),
                                ("travel_deaths", "travel_deaths"), ("births", "births"),
                                ("culled", "culled"), ("living_deaths", "living_deaths"),
                                ("agents_start", "agents_start"), ("agents_end", "agents_end")):
                assert gs[k_gpu] == st[k_or], f"step {s}: stat {k_gpu} gpu {gs[k_gpu]} != oracle {st[k_or]}"
        assert exact, "energies agree only within tolerance, expected bit-exact"
    finally:
        sim.close()


def test_philox_kat():
    """Random123 known-answer vectors through the DEVICE Philox (csrc/philox.cuh)."""
    from pdabm_b200 import _lib
    lib = _lib.load()
    vec = np.array([
        [0, 0, 0, 0, 0, 0],
        [0xFFFFFFFF] * 6,
        [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0],
    ], dtype=np.uint32)
    out = np.zeros((3, 4), dtype=np.uint32)
    rc = lib.pd_debug_philox(vec.ctypes.data, out.ctypes.data, 3)
    assert rc == 0
    exp = np.array([
        [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8],
        [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD],
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1],
    ], dtype=np.uint32)
    assert np.array_equal(out, exp)


@pytest.mark.parametrize("width", [16, 32, 64, 128, 256])
def test_default_kin_other(width):
    """Reference defaults (model.py:39-169): kin/other strategies, no noise, no mutation."""
    _run(40 if width <= 64 else 25, width=width, seed=12345 + width)


def test_default_long_saturation():
    """Run past the carrying capacity (cap reached near step 60) so the cull is exercised."""
    _run(90, width=64, seed=99)


@pytest.mark.parametrize("mode", ["pure", "per_trait", "kin_other"])
def test_modes_with_noise_and_mutation(mode):
    kw = dict(width=128, seed=4242, env_noise=0.05, mutation_rate=0.01)
    if mode == "pure":
        kw["strategy_pure"] = 1
    elif mode == "per_trait":
        kw["strategy_per_trait"] = 1
    _run(70, **kw)


def test_inheritance_mode():
    _run(70, width=64, seed=5, reproduction_inheritence=0.5, mutation_rate=0.2, env_noise=0.2)


def test_harsh_world_many_deaths():
    """Large negative payoffs and costs: most agents are at risk, play deaths chain."""
    _run(30, width=64, seed=11, payoff_cd=-30.0, payoff_dd=-5.0, cost_of_living=3.0, travel_cost=4.0,
         init_agent_count=2500)


def test_parents_that_die_of_the_living_cost():
    """Reproduction threshold below the cost of living: claim losers that the living cost would kill must stay
    live parents through the retry rounds (deferred list), winners pay reproduction cost first, then die."""
    _run(40, width=64, seed=77, reproduce_min_energy=1.5, reproduce_cost=0.5, cost_of_living=2.0,
         init_energy_mu=4.0, init_energy_sigma=2.0, init_energy_min=1.0, payoff_cc=3.0, payoff_dd=0.0,
         init_agent_count=1500, max_agents=3000)
    _run(25, width=64, seed=78, reproduce_min_energy=2.0, reproduce_cost=1.0, cost_of_living=2.5,
         init_energy_mu=5.0, init_energy_sigma=3.0, init_energy_min=1.0, env_noise=0.1, mutation_rate=0.05,
         strategy_per_trait=1, init_agent_count=1200, max_agents=4096)


def test_cull_select_fallback_path():
    """The cull's k-th-key select normally finishes in shared memory; with the stage capped at 1 key it must
    take the 8-bit passes over global memory whenever the threshold bin holds more -- same survivors."""
    _run(75, width=128, seed=41, debug_select_cap=1)
    _run(40, width=128, seed=42, debug_select_cap=1, sparse_ctas=8, init_agent_count=7000, max_agents=7200,
         reproduce_min_energy=40.0, reproduce_cost=10.0)


def test_local_claim_detour_path():
    """Reproduction claims on cells only one tile can reach are settled in shared memory by an atomic max of
    (priority top 29 bits | position); if two rivals agree in those bits the tile's claims take the exact route through
    the claim plane.  With the near-tie test widened to the top 4 / 8 bits almost every / every few tiles take that
    route -- the same winners either way."""
    _run(75, width=128, seed=43, debug_tie_shift=28)
    _run(75, width=256, seed=44, debug_tie_shift=24, strategy_per_trait=1, env_noise=0.05, mutation_rate=0.01)
    _run(25, width=128, seed=45, debug_tie_shift=30, init_agent_count=7000, max_agents=12000,
         reproduce_min_energy=40.0, reproduce_cost=10.0, payoff_cc=10.0)


def test_dense_world_conflicts():
    """Dense start: many travel/reproduction conflicts, retry rounds, full neighbourhoods."""
    _run(30, width=64, seed=3, init_agent_count=3000, max_agents=3600, reproduce_min_energy=40.0,
         reproduce_cost=10.0, payoff_cc=10.0)


def test_multi_cta_sparse_path():
    """Force the cooperative multi-CTA path of the persistent sparse kernels on a small world."""
    _run(70, width=128, seed=77, env_noise=0.1, mutation_rate=0.05, sparse_ctas=8,
         payoff_cd=-10.0, reproduce_min_energy=60.0)


@pytest.mark.parametrize("mode", ["kin_other", "pure"])
def test_config1_512_full_run(mode):
    """BASELINE config 1 at its own size and length: 512x512, model.py defaults, all 100 steps -- through the
    growth phase, the cap (near step 60) and 40 steps of cull -- in the literal default mode (kin/other,
    model.py:158-161) and in the "global strategy" (pure) variant BASELINE.json names."""
    _run(100, width=512, seed=12345, **({"strategy_pure": 1} if mode == "pure" else {}))


def test_config2_1024_own_size_past_the_cull():
    """BASELINE config 2 at its own size: 1024x1024, noise 0.05, mutation 0.01, kin/other; 72 steps so that the
    carrying capacity is reached and the cull (model.py:1339-1371) runs for several steps."""
    _run(72, width=1024, seed=12345, env_noise=0.05, mutation_rate=0.01)


def test_saturated_4096_per_trait_spot_check():
    """A size the 128x16-tile kernels had only been checked at through invariants: 4096x4096 (32 x 256 CTAs per
    sweep), per-trait mode, saturated.  The CUDA path runs 80 steps from init, its state is loaded into the
    oracle, and both then take 3 steps: every plane, the counts and the event counters bit-exact."""
    from pdabm_b200 import Simulation, SimConfig
    from oracle.pd_oracle import Oracle, Params
    kw = dict(width=4096, seed=2025, strategy_per_trait=1)
    sim = Simulation(SimConfig(collect_stats=True, **kw))
    try:
        sim.init_population()
        sim.step(80)
        pl = sim.planes()
        _, n = sim.counts()
        assert n > 0.95 * sim.cfg.max_agents
        orc = Oracle(Params(**kw))
        orc.load_planes(pl["occ"], pl["energy"], pl["trait"], pl["strat"].reshape(-1, 4))
        orc.step_index = sim.step_index
        exact = _compare(sim, orc, "loaded")
        for s in range(3):
            sim.step(1)
            st = orc.step()
            exact &= _compare(sim, orc, f"step {s}")
            gs = sim.stats()
            for k in ("play_deaths", "movers", "moved", "travel_deaths", "births", "culled", "living_deaths"):
                assert gs[k] == st[k], f"step {s}: stat {k} gpu {gs[k]} != oracle {st[k]}"
            assert st["culled"] > 0
        assert exact
    finally:
        sim.close()


def test_size_independent_invariants_large():
    """At a size the oracle cannot follow (4096^2) check the domain's invariants: population
    bookkeeping N' = N + births - culled - deaths, <= 1 agent per cell by construction,
    0 < energy <= max_energy, counts sum to the population, cap respected."""
    from pdabm_b200 import Simulation, SimConfig
    cfg = SimConfig(width=4096, seed=1, collect_stats=True, env_noise=0.05, mutation_rate=0.01)
    sim = Simulation(cfg)
    try:
        sim.init_population()
        counts, n = sim.counts()
        assert n == cfg.init_agent_count and counts.sum() == n
        for s in range(70):
            sim.step(1)
            st = sim.stats()
            counts, n = sim.counts()
            assert st["agents_end"] == n == counts.sum()
            assert st["agents_end"] == (st["agents_start"] - st["play_deaths"] - st["travel_deaths"]
                                        + st["births"] - st["culled"] - st["living_deaths"])
            assert n <= cfg.max_agents
        pl = sim.planes()
        occ = pl["occ"].astype(bool)
        assert occ.sum() == n
        e = pl["energy"][occ]
        assert e.min() > 0 and e.max() <= cfg.max_energy
        assert n == cfg.max_agents or n > 0.95 * cfg.max_agents  # saturated by step 70
    finally:
        sim.close()


def test_determinism_same_seed():
    from pdabm_b200 import Simulation, SimConfig
    outs = []
    for ctas in (0, 4):
        sim = Simulation(SimConfig(width=256, seed=321, env_noise=0.02, sparse_ctas=ctas))
        sim.init_population()
        sim.step(40)
        pl = sim.planes()
        outs.append((pl["tf_raw"].copy(), pl["energy"].copy(), pl["strat"].copy()))
        sim.close()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


def test_ensemble_driver_and_json_logs(tmp_path):
    """CUDAEnsemble-style driver (model.py:1801-1838): independent runs on concurrent streams, one JSON
    step log per run; each log must equal a stand-alone run of the same seed."""
    import json
    from pdabm_b200 import model, SimConfig
    base = SimConfig(width=64, seed=500)
    runs = model.configure_runplan(base, multi_run_count=2, multi_run_steps=12)[:6]
    paths = model.run_ensemble(runs, out_directory=str(tmp_path), devices=[0], streams_per_device=3)
    assert len(paths) == 6
    for run, path in zip(runs, paths):
        doc = json.loads(open(path).read().splitlines()[0])
        assert doc["config"]["random_seed"] == run["config"].seed
        assert doc["config"]["environment"]["cost_of_living"] == run["config"].cost_of_living
        assert [s["step_index"] for s in doc["steps"]] == list(range(13))
        solo = model.simulate(run["config"], 12)
        assert doc["steps"] == solo.doc["steps"]


def test_state_export(tmp_path):
    from pdabm_b200 import Simulation, SimConfig
    sim = Simulation(SimConfig(width=64, seed=3))
    sim.init_population()
    sim.step(3)
    sim.export_state(str(tmp_path / "w.npz"))
    sim.export_state(str(tmp_path / "w.ppm"))
    z = np.load(tmp_path / "w.npz")
    assert z["occ"].sum() == sim.counts()[1] and z["strat"].shape == (64, 64, 4)
    assert open(tmp_path / "w.ppm", "rb").read(2) == b"P6"
    sim.close()


# ----------------------------------------------------------------------------- edge cases
def test_empty_world():
    _run(3, width=32, seed=1, init_agent_count=0)


def test_single_agent_world():
    _run(30, width=32, seed=2, init_agent_count=1)


def test_full_world_no_free_cell():
    """Every cell occupied: nobody can move or reproduce until deaths open cells."""
    _run(25, width=32, seed=3, init_agent_count=32 * 32, max_agents=32 * 32, cost_of_living=4.0)


def test_extinction():
    """A cost of living that kills everybody: the population reaches 0 and stays there (exit condition, :1579-1588)."""
    sim, orc = _pair(width=32, seed=4, cost_of_living=60.0)
    try:
        sim.init_population()
        orc.init_population()
        for s in range(4):
            sim.step(1)
            orc.step()
            _compare(sim, orc, f"step {s}")
        assert sim.counts()[1] == 0
    finally:
        sim.close()


def test_no_children_allowed_and_non_negative_payoffs():
    """max_children_per_step = 0 (:1054) and a payoff matrix without negative entries (nobody is ever at risk)."""
    _run(40, width=64, seed=5, max_children_per_step=0, payoff_cd=0.0)
    _run(70, width=64, seed=6, payoff_cd=0.5, payoff_dd=0.25)


def test_cap_below_initial_population():
    """max_agents below the initial population: every newborn is culled, olds are never culled."""
    _run(40, width=64, seed=7, init_agent_count=1500, max_agents=1000, reproduce_min_energy=45.0, reproduce_cost=5.0)


def test_skewed_strategy_weights_and_energy_parameters():
    _run(50, width=64, seed=8, strategy_weights=[0.7, 0.1, 0.1, 0.1], init_energy_mu=80.0, init_energy_sigma=30.0,
         init_energy_min=1.0, max_energy=120.0, reproduce_min_energy=90.0, strategy_per_trait=1, mutation_rate=0.5)


def test_multi_step_call_equals_single_steps():
    from pdabm_b200 import Simulation, SimConfig
    a = Simulation(SimConfig(width=128, seed=9, env_noise=0.05))
    b = Simulation(SimConfig(width=128, seed=9, env_noise=0.05))
    try:
        a.init_population()
        b.init_population()
        a.step(37)
        for _ in range(37):
            b.step(1, want_counts=False)
        b.step(0)
        pa, pb = a.planes(), b.planes()
        for k in ("tf_raw", "energy", "strat"):
            assert np.array_equal(pa[k], pb[k])
        assert a.counts()[1] == b.counts()[1] and a.step_index == b.step_index == 37
    finally:
        a.close()
        b.close()


def test_host_buffer_step_roundtrip():
    """pd_step_host (the e2e path): host planes in, host planes out == device-resident stepping."""
    import torch
    from pdabm_b200 import Simulation, SimConfig
    a = Simulation(SimConfig(width=128, seed=10, mutation_rate=0.02))
    b = Simulation(SimConfig(width=128, seed=10, mutation_rate=0.02))
    try:
        a.init_population()
        b.init_population()
        h_e, h_s, h_t = a.energy.cpu().pin_memory(), a.strat.cpu().pin_memory(), a.tf.cpu().pin_memory()
        for s in range(8):
            a.step(1)
            counts_b, n_b = b.step_host(h_e, h_s, h_t, 1, h_e, h_s, h_t)
            assert torch.equal(h_t, a.tf.cpu()) and torch.equal(h_e, a.energy.cpu())
            ca, na = a.counts()
            assert na == n_b and np.array_equal(ca, counts_b)
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("width", [128, 1024])
def test_graph_replay_equals_direct_launches_and_oracle(width):
    """On an explicit stream a small world replays one captured CUDA graph per step (the step number lives
    on the device).  The replayed run must equal the directly launched one bit for bit -- with and without
    counts, across multi-step calls -- and, at the small size, the oracle."""
    import torch
    from pdabm_b200 import Simulation, SimConfig
    from oracle.pd_oracle import Oracle, Params
    kw = dict(width=width, seed=99, env_noise=0.05, mutation_rate=0.01, strategy_per_trait=1)
    st = torch.cuda.Stream()
    g = Simulation(SimConfig(use_graphs=1, **kw), stream=st)
    d = Simulation(SimConfig(use_graphs=-1, **kw), stream=st)
    orc = Oracle(Params(**kw)) if width <= 128 else None
    try:
        g.init_population()
        d.init_population()
        if orc:
            orc.init_population()
        for rnd, (k, wc) in enumerate([(1, True), (1, False), (3, True), (2, False), (1, True), (60, True), (5, True)]):
            g.step(k, want_counts=wc)
            d.step(k, want_counts=wc)
            a, b = g.planes(), d.planes()
            for name in ("occ", "trait", "strat", "energy", "tf_raw"):
                assert np.array_equal(a[name], b[name]), f"round {rnd}: plane {name} differs (graph vs direct)"
            if wc:
                ca, na = g.counts()
                cb, nb = d.counts()
                assert na == nb and np.array_equal(ca, cb)
            if orc:
                for _ in range(k):
                    orc.step()
                ref = orc.planes()
                for name in ("occ", "trait", "strat", "energy"):
                    assert np.array_equal(a[name], ref[name]), f"round {rnd}: plane {name} differs from the oracle"
                if wc:
                    assert na == orc.n and np.array_equal(ca, orc.population_strat_count)
        assert g.agent_steps() == d.agent_steps()
        assert g.graph_replays() == 73 and d.graph_replays() == 0  # every step of g came from a graph
    finally:
        g.close()
        d.close()


def test_device_step_log_equals_per_step_readback():
    """pd_set_step_log: counts of every counted step stay on the device and are read back together; they must
    equal what pd_read_counts returns step by step, in order, with the step index and the agent count."""
    import torch
    from pdabm_b200 import Simulation, SimConfig, PdabmError
    kw = dict(width=128, seed=5, env_noise=0.05, mutation_rate=0.02)
    a = Simulation(SimConfig(**kw), stream=torch.cuda.Stream())
    b = Simulation(SimConfig(**kw))
    try:
        a.init_population()
        b.init_population()
        a.enable_step_log(64)
        a.step(10, want_counts="every")
        a.step(5, want_counts=True)
        a.step(3, want_counts=False)
        a.step(1, want_counts=True)
        rows = a.read_step_log()
        assert [int(r[0]) for r in rows] == list(range(1, 11)) + [15, 19]
        expect = {}
        for s in range(1, 20):
            b.step(1)
            expect[s] = b.counts()
        for r in rows:
            counts, n = expect[int(r[0])]
            assert int(r[1]) == n and np.array_equal(r[2:], counts), f"step {int(r[0])}"
        assert a.read_step_log().shape == (0, 18)  # emptied by the read
        a.enable_step_log(4)
        a.step(6, want_counts="every")
        with pytest.raises(PdabmError):
            a.read_step_log()
    finally:
        a.close()
        b.close()


def test_checkpoint_resume_is_bit_identical(tmp_path):
    """Planes + step index are the whole state of a run: a fresh handle that loads them continues exactly like
    the uninterrupted run (the step index is a Philox key word and lives on the device)."""
    import torch
    from pdabm_b200 import Simulation, SimConfig, PdabmError
    kw = dict(width=128, seed=2024, env_noise=0.05, mutation_rate=0.02, strategy_per_trait=1)
    a = Simulation(SimConfig(**kw))
    b = Simulation(SimConfig(**kw))
    c = Simulation(SimConfig(**kw), stream=torch.cuda.Stream())  # resumes through the graph-replay path
    d = Simulation(SimConfig(**{**kw, "seed": 1}))
    try:
        a.init_population()
        a.step(70)
        b.init_population()
        b.step(41)
        assert b.step_index == 41
        path = str(tmp_path / "ckpt.npz")
        b.save_checkpoint(path)
        c.load_checkpoint(path)
        assert c.step_index == 41
        c.step(29)
        pa, pc = a.planes(), c.planes()
        for k in ("occ", "trait", "strat", "energy"):
            assert np.array_equal(pa[k], pc[k]), f"plane {k} differs after resume"
        ca, na = a.counts()
        cc, nc = c.counts()
        assert na == nc and np.array_equal(ca, cc)
        with pytest.raises(PdabmError):
            d.load_checkpoint(path)
    finally:
        for s in (a, b, c, d):
            s.close()

"""Host-side checks of the round-2 additions: the restated end-of-generation bookkeeping
(oracle/ga_es.generation_end, the checker of cev_generation_end_f64), the two selection orders,
and engine checkpoint / resume through the checker backend."""
import os
import sys
import types

import numpy as np
import torch

from oracle import ga_es

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROLES = ("agent_0", "agent_1", "adversary_0")


def _np_mean_like_the_kernel(a):
    """The summation order generation_end_kernel uses (NumPy's pairwise sum for n < 128)."""
    n = len(a)
    if n < 8:
        res = 0.0
        for x in a:
            res += x
    else:
        r = list(a[:8])
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] += a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += a[i]
            i += 1
    return res / n


def test_kernel_mean_order_is_numpys():
    rng = np.random.default_rng(0)
    for n in range(1, 41):
        for _ in range(50):
            a = list(rng.normal(-20, 7, size=n))
            assert _np_mean_like_the_kernel(a) == np.mean(a), n


def test_generation_end_follows_the_reference_loop_tail():
    """Literal transcription of genetic_algorithm.py:301-345 / evolutionary_strategy.py:272-354 on Python
    lists vs the array-state restatement."""
    rng = np.random.default_rng(3)
    cap = 40
    gs = ga_es.generation_state([0.005, 0.05, 0.07], cap)
    sig = [0.005, 0.05, 0.07]
    hist = [[], [], []]
    best, stale, stopped = [-np.inf] * 3, [0] * 3, None
    smax, smin, min_delta, patience = 0.2, 0.001, 0.1, 5
    trend = np.concatenate([np.linspace(-30, -5, 14), np.linspace(-5, -25, 13), np.full(13, -25.0)])
    for gen in range(36):
        out = rng.normal(trend[gen], 2.0, size=(10, 4))
        out[:, 1] = rng.normal(0, 0.3, size=10)
        # evaluate_current_weights: running sums of play_game's triple over 10 games, / 10
        tot = [0.0, 0.0, 0.0]
        for g in range(10):
            sg, lg, sa = out[g, 0], out[g, 1], out[g, 2]
            tot[0] += sg - lg          # agent_0's slot: good reward up to cycle 24 (Appendix B)
            tot[1] += sa
            tot[2] += sg - lg
        ev = [t / 10 for t in tot]
        for r in range(3):
            hist[r].append(ev[r])
        sig = list(ga_es.adaptive_sigma(sig[0], sig[1], sig[2], hist[0], hist[1], hist[2], gen, smin, smax))
        if stopped is None:
            for r in range(3):
                if ev[r] > best[r] + min_delta:
                    best[r], stale[r] = ev[r], 0
                else:
                    stale[r] += 1
            for r in range(3):
                if stale[r] >= patience:
                    stopped = (r, gen)
                    break
        ga_es.generation_end(out, gs, cap, agent_step_limit=400, reference_compat=True, adaptive=True,
                             sigma_max=smax, sigma_min=smin, early_stopping=True, min_delta=min_delta,
                             patience=patience)
        assert list(gs[ga_es.GS_SIGMA:ga_es.GS_SIGMA + 3]) == sig
        assert list(gs[ga_es.GS_LAST_EVAL:ga_es.GS_LAST_EVAL + 3]) == ev
        assert int(gs[ga_es.GS_GEN]) == gen + 1
        if stopped is not None:
            assert int(gs[ga_es.GS_STOP]) == 1 + stopped[0] and int(gs[ga_es.GS_STOP_GEN]) == stopped[1]
    assert stopped is not None
    h = gs[ga_es.GS_HIST:ga_es.GS_HIST + 3 * cap].reshape(cap, 3)
    assert np.array_equal(h[:36].T, np.asarray(hist))


def test_selection_orders():
    rng = np.random.default_rng(1)
    for n in (3, 8, 16):
        for _ in range(50):
            f = rng.integers(0, 3, size=n).astype(np.float64)
            # the reference's expression with NumPy's stable (insertion / merge) sort; the DEFAULT kind is
            # x86-simd-sort on AVX2 / AVX-512 hosts since NumPy 1.25 and leaves ties in unspecified order
            assert np.array_equal(ga_es.select_topk(f, n, order=1), np.argsort(f, kind="stable")[::-1])
            assert np.array_equal(ga_es.select_topk(f, n, order=0), np.argsort(-f, kind="stable"))
    assert list(ga_es.select_topk([1.0, 3.0, 3.0, 2.0, 3.0], 3, order=1)) == [4, 2, 1]
    assert list(ga_es.select_topk([1.0, np.nan, 3.0], 3, order=1)) == [1, 2, 0]               # NaN first (Appendix C #18)


def _args(algorithm, P):
    return types.SimpleNamespace(
        algorithm=algorithm, generations=4, population=P, hof_size=2, game="simple_adversary_v3",
        mutation_power_agent_0=0.05, mutation_power_agent_1=0.04, mutation_power_adversary=0.03,
        learning_rate=0.1, max_timesteps_per_episode=400, max_evaluation_steps=400, elites_number=2,
        adaptive=True, max_mutation_power=0.5, min_mutation_power=0.001, fitness_sharing=True,
        early_stopping=False, patience=300, min_delta=0.1, debug=False, precision="float32", save=False,
        envs_per_member=1, reference_compat=True, init_states="reference", seed=99, plots=False,
        record_history=False)


def _engine(algorithm, P=4):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_backend
    from coevonet_b200 import engine, layout
    from coevonet_b200.utils import mpe_spec
    from oracle import weights
    args = _args(algorithm, P)
    rows = {}
    for i, r in enumerate(ROLES):
        in_dim = layout.OBS_DIM[r]
        full = np.zeros((P + 3, layout.fc_pitch(in_dim)), dtype=np.float32)
        full[:, :layout.fc_dim(in_dim)] = weights.make_fc_rows(P + 3, in_dim, 300 + i, ln_jitter=0.02)
        rows[r] = torch.from_numpy(full)
    env = mpe_spec.DeviceMPEEnv()
    if algorithm == "ES":
        return engine.ESEngine(args, "cpu", {r: rows[r][P + 2].clone() for r in ROLES}, env=env,
                               kernels=oracle_backend)
    return engine.GAEngine(args, "cpu", {r: rows[r][:P].clone() for r in ROLES},
                           {r: rows[r][P:P + 2].clone() for r in ROLES},
                           {r: rows[r][P - 1].clone() for r in ROLES}, env=env, kernels=oracle_backend)


def test_resume_equals_uninterrupted_run_host_logic(tmp_path):
    for algorithm in ("ES", "GA"):
        full = _engine(algorithm)
        for _ in range(4):
            full.step(sync=False)
        first = _engine(algorithm)
        for _ in range(2):
            first.step(sync=False)
        path = tmp_path / f"{algorithm}.pt"
        torch.save(first.state_dict(), path)
        second = _engine(algorithm)
        second.load_state_dict(torch.load(path, weights_only=False))
        for _ in range(2):
            second.step(sync=False)
        assert torch.equal(second.gstate, full.gstate), algorithm
        if algorithm == "ES":
            for r in ROLES:
                assert torch.equal(second.theta[r], full.theta[r])
        else:
            for r in ROLES:
                assert torch.equal(second.pop[r], full.pop[r]) and torch.equal(second.hof[r], full.hof[r])
        hs = second.host_state()
        assert hs["generations"] == 4 and len(hs["sigma_history"]["agent_1"]) == 5


def test_crossover_restatement_properties():
    """The crossover extension (the reference has none, README.md:47 vs genetic_algorithm.py:32-48):
    rate 0 is the reference's clone + mutate; at rate 1 every child parameter comes from one of two
    DIFFERENT elites, roughly half from each, and child 0 stays the unmutated best."""
    from oracle import philox
    rng = np.random.default_rng(2)
    E, D, n = 3, 1000, 9
    elites = rng.normal(size=(E, D)).astype(np.float32)
    z = np.zeros((n, D), dtype=np.float32)
    members = np.arange(n)
    plain = ga_es.ga_repopulate_crossover(elites, 0.05, z, 0.0, 11, 1, 4, members)
    for c in range(1, n):
        assert np.array_equal(plain[c], elites[(c - 1) % E])
    crossed = ga_es.ga_repopulate_crossover(elites, 0.05, z, 1.0, 11, 1, 4, members)
    assert np.array_equal(crossed[0], elites[0])
    for c in range(1, n):
        a = (c - 1) % E
        from_a = crossed[c] == elites[a]
        others = [b for b in range(E) if b != a and np.array_equal(crossed[c][~from_a], elites[b][~from_a])]
        assert len(others) == 1, "the complement must come from exactly one other elite"
        assert 0.4 < from_a.mean() < 0.6
    half = ga_es.ga_repopulate_crossover(elites, 0.05, np.zeros((400, D), dtype=np.float32), 0.5, 11, 1, 4, np.arange(400))
    n_crossed = sum(not np.array_equal(half[c], elites[(c - 1) % E]) for c in range(1, 400))
    assert 150 < n_crossed < 250

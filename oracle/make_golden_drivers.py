"""Driver-level golden runs: the UNMODIFIED reference ``genetic_algorithm_train``
and ``evolution_strategy_train`` for a few generations, with their unseeded
noise sources replaced by the counter-based Philox noise the device uses.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``); build container only:

    python -m oracle.make_golden_drivers

What is injected (nothing in /root/reference is edited):
* ``torch.manual_seed`` before the run -> founders reproducible;
* ``torch.normal`` (GA, agent.py:28) / ``np.random.normal`` (ES, agent.py:52)
  return ``sigma * z`` with z = Philox(seed, kind, role, gen, member, param);
* observers around ``np.argsort``, ``diversity_penalty``,
  ``evaluate_current_weights``, ``compute_weight_update`` record what the
  reference computed.
Outputs: tests/golden/ga_run.npz, tests/golden/es_run.npz.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import layout, philox, stubs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 1870300
ROLES = ("agent_0", "agent_1", "adversary_0")


def run_ga(ref, P=6, hof=2, elites=2, gens=3, torch_seed=123):
    ga = ref.genetic_algorithm
    args = stubs.RefArgs(algorithm="GA", population=P, hof_size=hof, elites_number=elites, generations=gens,
                         mutation_power_agent_0=0.005, mutation_power_agent_1=0.05,
                         mutation_power_adversary=0.05, adaptive=True, fitness_sharing=True, save=False)
    rec = dict(fitness=[], diversity=[], evals=[], sigma=[])
    ctx = dict(role=None, gen=-1, child=1, off=0)
    dims = {r: layout.fc_dim(layout.OBS_DIM[r]) for r in ROLES}

    orig_normal, orig_argsort = torch.normal, np.argsort
    orig_me, orig_dp, orig_ev = ga.mutate_elites, ga.diversity_penalty, ga.evaluate_current_weights

    def fake_normal(mean, std, size):
        n = int(np.prod(size))
        D = dims[ctx["role"]]
        z = philox.normals(SEED, philox.KIND_GA, philox.ROLE_ID[ctx["role"]], ctx["gen"], [ctx["child"]], D)[0]
        out = (np.float32(std) * z[ctx["off"]:ctx["off"] + n]).astype(np.float32).reshape(tuple(size))
        ctx["off"] += n
        if ctx["off"] >= D:
            ctx["off"] = 0
            ctx["child"] += 1
        return torch.from_numpy(out)

    def me(env, elites_, args_, role):
        if role == "agent_0":
            ctx["gen"] += 1
        ctx.update(role=role, child=1, off=0)
        return orig_me(env, elites_, args_, role)

    def argsort(a, *k, **kw):
        rec["fitness"].append(np.asarray(a, dtype=np.float64).copy())
        return orig_argsort(a, *k, **kw)

    div_seen = []

    def dp(*a, **kw):
        v = orig_dp(*a, **kw)
        div_seen.append(float(v))
        return v

    def ev(*a, **kw):
        v = orig_ev(*a, **kw)
        rec["evals"].append(np.asarray(v, dtype=np.float64))
        rec["sigma"].append([args.mutation_power_agent_0, args.mutation_power_agent_1,
                             args.mutation_power_adversary])
        # one diversity value per role per generation (all P calls return the same, Appendix C #3)
        per_role = np.asarray(div_seen).reshape(3, -1)
        assert np.all(per_role == per_role[:, :1])
        rec["diversity"].append(per_role[:, 0].copy())
        div_seen.clear()
        return v

    torch.normal, np.argsort = fake_normal, argsort
    ga.mutate_elites, ga.diversity_penalty, ga.evaluate_current_weights = me, dp, ev
    try:
        torch.manual_seed(torch_seed)
        env = ref.utils_game_logic_functions.initialize_env(args)
        ga.genetic_algorithm_train(env, env.agents[0], args, "/tmp")
    finally:
        torch.normal, np.argsort = orig_normal, orig_argsort
        ga.mutate_elites, ga.diversity_penalty, ga.evaluate_current_weights = orig_me, orig_dp, orig_ev
    fit = np.asarray(rec["fitness"]).reshape(gens, 3, P)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ga_run.npz"),
                        source="genetic_algorithm.py:51 genetic_algorithm_train, Philox noise injected",
                        P=P, hof=hof, elites=elites, gens=gens, torch_seed=torch_seed, seed=SEED,
                        sigma0=np.array([0.005, 0.05, 0.05]), max_sigma=args.max_mutation_power,
                        min_sigma=args.min_mutation_power, fitness=fit, diversity=np.asarray(rec["diversity"]),
                        evals=np.asarray(rec["evals"]), sigma_used=np.asarray(rec["sigma"]),
                        final_sigma=np.array([args.mutation_power_agent_0, args.mutation_power_agent_1,
                                              args.mutation_power_adversary]))
    print("[ga_run] fitness gen0 agent_0:", fit[0, 0])
    print("[ga_run] evals:", np.asarray(rec["evals"]))


def run_es(ref, P=8, gens=3, torch_seed=321):
    es = ref.evolutionary_strategy
    args = stubs.RefArgs(algorithm="ES", population=P, hof_size=1, generations=gens,
                         mutation_power_agent_0=0.05, mutation_power_agent_1=0.05,
                         mutation_power_adversary=0.05, learning_rate=0.1, adaptive=True,
                         fitness_sharing=True, save=False, max_mutation_power=0.5, min_mutation_power=0.001)
    rec = dict(rewards=[], updates=[], diversity=[], evals=[])
    ctx = dict(role=None, gen=0, member={r: 0 for r in ROLES})
    pidx = {r: layout.fc_perturbable_index(layout.OBS_DIM[r]) for r in ROLES}
    dims = {r: layout.fc_dim(layout.OBS_DIM[r]) for r in ROLES}
    orig_np_normal = np.random.normal
    orig_mw, orig_cwu, orig_ev = es.mutate_weights, es.compute_weight_update, es.evaluate_current_weights

    def fake_np_normal(loc=0.0, scale=1.0, size=None):
        r = ctx["role"]
        z = philox.normals(SEED, philox.KIND_ES, philox.ROLE_ID[r], ctx["gen"], [ctx["member"][r]], dims[r])[0]
        return (np.float32(scale) * z[pidx[r]]).astype(np.float32).astype(np.float64)

    def mw(env, a0, a1, adv, args_, role, step, *logs):
        ctx.update(role=role, gen=step)
        out = orig_mw(env, a0, a1, adv, args_, role, step, *logs)
        ctx["member"][role] += 1
        return out

    def cwu(noises, rewards, args_, role, **kw):
        upd, div = orig_cwu(noises, rewards, args_, role, **kw)
        rec["rewards"].append(np.asarray(rewards, dtype=np.float64))
        rec["updates"].append(upd.copy())
        rec["diversity"].append(float(div))
        return upd, div

    def ev(*a, **kw):
        v = orig_ev(*a, **kw)
        rec["evals"].append(np.asarray(v, dtype=np.float64))
        ctx["member"] = {r: 0 for r in ROLES}
        return v

    np.random.normal = fake_np_normal
    es.mutate_weights, es.compute_weight_update, es.evaluate_current_weights = mw, cwu, ev
    try:
        torch.manual_seed(torch_seed)
        env = ref.utils_game_logic_functions.initialize_env(args)
        a0, a1, adv = es.evolution_strategy_train(env, args, "/tmp")
    finally:
        np.random.normal = orig_np_normal
        es.mutate_weights, es.compute_weight_update, es.evaluate_current_weights = orig_mw, orig_cwu, orig_ev
    upd_norm = np.array([[np.linalg.norm(rec["updates"][g * 3 + i]) for i in range(3)] for g in range(gens)])
    final = [m.model.get_perturbable_weights() for m in (a0, a1, adv)]
    np.savez_compressed(os.path.join(GOLDEN_DIR, "es_run.npz"),
                        source="evolutionary_strategy.py:151 evolution_strategy_train, Philox noise injected",
                        P=P, gens=gens, torch_seed=torch_seed, seed=SEED, lr=0.1, sigma0=0.05,
                        max_sigma=0.5, min_sigma=0.001,
                        rewards=np.asarray(rec["rewards"]).reshape(gens, 3, P),
                        diversity=np.asarray(rec["diversity"]).reshape(gens, 3),
                        evals=np.asarray(rec["evals"]), update_norm=upd_norm,
                        update_head=np.asarray([u[:256] for u in rec["updates"]]).reshape(gens, 3, 256),
                        final_head=np.asarray([f[:256] for f in final]),
                        final_norm=np.array([np.linalg.norm(f) for f in final]))
    print("[es_run] rewards gen0 agent_0:", np.asarray(rec["rewards"][0]))
    print("[es_run] evals:", np.asarray(rec["evals"]))


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = stubs.import_reference()
    torch.set_num_threads(1)
    run_ga(ref)
    run_es(ref)
    return 0


if __name__ == "__main__":
    sys.exit(main())

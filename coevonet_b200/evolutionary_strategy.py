"""Co-ES driver with the reference's entry points (``evolutionary_strategy.py``).

``evolution_strategy_train(env, args, output_dir) -> (agent_0, agent_1, adversary)``
keeps the reference's signature, side effects and -- under
``args.reference_compat`` -- its quirks.  Perturbation, rollouts and the
fitness-weighted update run on the device through ``engine.ESEngine``; noise is
regenerated from Philox counters instead of being stored per member.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from tqdm import tqdm

from . import engine as _engine
from . import layout, ops
from .utils.game_logic_functions import create_agent, play_game
from .utils.utils_pth_and_plots import plot_experiment_metrics, plot_weights_logging, save_model

ROLES = layout.ROLES


def get_numpy_dtype(precision):
    if precision == "float16":
        return np.float16
    return np.float32


def evaluate_current_weights(agent_0, agent_1, adversary, env, args):
    """Mean reward triple of 10 evaluation games (evolutionary_strategy.py:22-59)."""
    tot = np.zeros(3)
    for _ in range(_engine.N_EVAL_GAMES):
        tot += play_game(env=env, player1=agent_0.model, player2=agent_1.model,
                         adversary=adversary.model, args=args, eval=True)
    return tuple(tot / _engine.N_EVAL_GAMES)


def mutate_weights(env, agent_0, agent_1, adversary, args, role, step, weights_logging_agent_0,
                   weights_logging_agent_1, weights_logging_adversary):
    """One perturbed member of ``role`` and its episode against the other two BASE
    agents (evolutionary_strategy.py:63-116): returns (own-seat reward, noise
    fp32[D_pert] already scaled by sigma, mutated flat weights)."""
    np_dtype = get_numpy_dtype(args.precision)
    base = {"agent_0": agent_0, "agent_1": agent_1, "adversary_0": adversary}[role]
    mutant = base.clone(env, args, role=role)
    noise = mutant.mutate_ES(args, role=role, step=step, weights_logging_agent_0=weights_logging_agent_0,
                             weights_logging_agent_1=weights_logging_agent_1,
                             weights_logging_adversary=weights_logging_adversary).astype(np_dtype)
    weights = mutant.model.get_weights_ES()
    players = {"agent_0": agent_0, "agent_1": agent_1, "adversary_0": adversary}
    players[role] = mutant
    r = play_game(env=env, player1=players["agent_0"].model, player2=players["agent_1"].model,
                  adversary=players["adversary_0"].model, args=args)
    return r[ROLES.index(role)], noise, weights


def compute_weight_update(noises, rewards, args, role, individual_weights=None, population_weights=None):
    """delta = lr / (n * sigma) * noises^T fitness (evolutionary_strategy.py:120-148) for
    caller-supplied noise arrays.  The contraction is a [D x P] GEMV over host
    data handed in by the caller; it runs on the device (torch matmul is library
    plumbing here -- the population path uses the K6 kernel, which regenerates
    the noise instead of reading it)."""
    from .utils.game_logic_functions import _device, diversity_penalty
    dev = _device()
    noises = torch.from_numpy(np.asarray(noises, dtype=np.float32)).to(dev)
    fit = torch.from_numpy(np.asarray(rewards, dtype=np.float32)).to(dev)
    diversity = None
    if args.fitness_sharing:
        diversity = diversity_penalty(individual_weights=individual_weights,
                                      population_weights=population_weights, args=args)
        fit = fit / (1 + diversity)
    sigma = {"agent_0": args.mutation_power_agent_0, "agent_1": args.mutation_power_agent_1,
             "adversary_0": args.mutation_power_adversary}[role]
    upd = (args.learning_rate / (noises.shape[0] * sigma)) * (noises.T @ fit)
    return upd.cpu().numpy().astype(np.float32), diversity


def evolution_strategy_train(env, args, output_dir):
    """Reference signature (evolutionary_strategy.py:151)."""
    files = {"agent_0": os.path.join(output_dir, "agent_0.pth"), "agent_1": os.path.join(output_dir, "agent_1.pth"),
             "adversary_0": os.path.join(output_dir, "adversary.pth")}
    plots = {"agent_0": os.path.join(output_dir, "results_plot_file_agent_0.png"),
             "agent_1": os.path.join(output_dir, "results_plot_file_agent_1.png"),
             "adversary_0": os.path.join(output_dir, "results_plot_file_adversary.png")}
    weights_plot = os.path.join(output_dir, "weights_results_plot_file.png")

    from .utils.game_logic_functions import _device
    dev = _device()
    comm = _engine.Comm()
    agents = {r: create_agent(env, args, role=r) for r in ROLES}
    for name, r in (("agent_0", "agent_0"), ("agent_1", "agent_1"), ("adversary", "adversary_0")):
        print(f"\nNumber of parameters for {name} network: {sum(p.numel() for p in agents[r].model.parameters())}")
    theta = {r: agents[r].model.flat_row() for r in ROLES}
    eng = _engine.ESEngine(args, dev, theta, env=env, comm=comm)

    rewards = {r: [] for r in ROLES}
    wlog = {r: [] for r in ROLES}
    diversity = {r: [] if args.fitness_sharing else None for r in ROLES}
    fitness = {r: [] if args.fitness_sharing else None for r in ROLES}
    sigma_hist = {r: [eng.sigma(r)] if args.adaptive else None for r in ROLES}
    best = {r: -float("inf") for r in ROLES}
    stale = {r: 0 for r in ROLES}

    for gen in tqdm(range(args.generations), desc="Training Generations"):
        ev = dict(zip(ROLES, eng.step()))
        for r in ROLES:
            rewards[r].append(ev[r])
            if args.fitness_sharing:
                diversity[r].append(eng.diversity[r])
                fitness[r].append(ev[r] / (1 + eng.diversity[r]))
        if getattr(args, "log_weight_stats", True):
            # per-generation statistics of the base weights (the reference logs them per member,
            # MPE/mpe_agent.py:30-50; one sample per generation keeps the plot and drops P syncs)
            for r in ROLES:
                pidx = torch.from_numpy(layout.fc_perturbable_index(layout.OBS_DIM[r])).to(dev)
                w = eng.theta[r][pidx]
                wlog[r].append({"step": gen, "mean": float(w.mean()), "min": float(w.min()),
                                "max": float(w.max()), "std": float(w.std())})
        if args.adaptive:
            _engine.adapt_sigma(args, rewards["agent_0"], rewards["agent_1"], rewards["adversary_0"], gen)
            for r in ROLES:
                sigma_hist[r].append(eng.sigma(r))
        if args.early_stopping:
            stop = None
            for r in ROLES:
                if ev[r] > best[r] + args.min_delta:
                    best[r], stale[r] = ev[r], 0
                else:
                    stale[r] += 1
            for r, name in (("agent_0", "agent_0"), ("agent_1", "agent_1"), ("adversary_0", "adversary")):
                if stale[r] >= args.patience:
                    print(f"Early stopping triggered at generation {gen} for {name}. Best reward: {best[r]}")
                    stop = r
                    break
            if stop is not None:
                break
        if (args.save or gen == args.generations - 1) and comm.rank == 0:
            for r in ROLES:
                agents[r].model.load_flat_row(eng.theta[r])
                if args.save:
                    save_model(agents[r], files[r])
        if comm.rank == 0 and getattr(args, "plots", True):
            for r in ROLES:
                plot_experiment_metrics(rewards=rewards[r], mutation_power_history=sigma_hist[r],
                                        fitness=fitness[r], diversity=diversity[r], file_path=plots[r], args=args)
            plot_weights_logging(weights_plot, wlog["agent_0"], wlog["agent_1"], wlog["adversary_0"])
    for r in ROLES:
        agents[r].model.load_flat_row(eng.theta[r])
    args._es_engine = eng
    return agents["agent_0"], agents["agent_1"], agents["adversary_0"]

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
from . import layout
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
    caller-supplied noise arrays: the K6 kernel in its members form with theta = 0, so the
    ``noises`` rows themselves are the sigma*z terms (any row length: layout-agnostic path)."""
    from . import ops
    from .utils.game_logic_functions import _device, diversity_penalty
    dev = _device()
    noises = np.ascontiguousarray(np.asarray(noises, dtype=np.float32))
    n, D = noises.shape
    pitch = (D + 31) // 32 * 32
    rows = torch.zeros((n, pitch), dtype=torch.float32, device=dev)
    rows[:, :D] = torch.from_numpy(noises).to(dev)
    fit = torch.from_numpy(np.asarray(rewards, dtype=np.float32)).to(dev)
    diversity = None
    if args.fitness_sharing:
        diversity = diversity_penalty(individual_weights=individual_weights,
                                      population_weights=population_weights, args=args)
        fit = fit / (1 + diversity)
    sigma = {"agent_0": args.mutation_power_agent_0, "agent_1": args.mutation_power_agent_1,
             "adversary_0": args.mutation_power_adversary}[role]
    zero = torch.zeros(pitch, dtype=torch.float32, device=dev)
    upd = ops.es_update_members(fit.to(torch.float64).contiguous(), rows, zero, 0, sigma, args.learning_rate, n)
    return upd[:D].cpu().numpy().astype(np.float32), diversity


def checkpoint_path(output_dir):
    """Engine checkpoint written beside the reference-format ``.pth`` files (true resume, N2)."""
    rank = _engine.Comm().rank
    return os.path.join(output_dir, f"engine_state_rank{rank}.pt")


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
    # every rank builds the base agents from rank 0's generator state (unseeded runs included)
    _engine.sync_torch_rng(comm, dev)
    agents = {r: create_agent(env, args, role=r) for r in ROLES}
    if comm.rank == 0:
        for name, r in (("agent_0", "agent_0"), ("agent_1", "agent_1"), ("adversary", "adversary_0")):
            print(f"\nNumber of parameters for {name} network: "
                  f"{sum(p.numel() for p in agents[r].model.parameters())}")
    theta = {r: agents[r].model.flat_row() for r in ROLES}
    eng = _engine.ESEngine(args, dev, theta, env=env, comm=comm)
    start = 0
    if getattr(args, "resume", False) and os.path.isfile(checkpoint_path(output_dir)):
        eng.load_state_dict(torch.load(checkpoint_path(output_dir), weights_only=False))
        start = eng.gen
        if comm.rank == 0:
            print(f"Resuming from generation {start} ({checkpoint_path(output_dir)})")

    wlog = {r: [] for r in ROLES}
    diversity = {r: [] if args.fitness_sharing else None for r in ROLES}
    fitness = {r: [] if args.fitness_sharing else None for r in ROLES}
    # the per-generation host view (rewards, sigma history) is read back only when something on the host
    # consumes it: plots, checkpoints, early stopping, the last generation
    want_plots = comm.rank == 0 and getattr(args, "plots", True)
    pidx = {r: torch.from_numpy(layout.fc_perturbable_index(layout.OBS_DIM[r])).to(dev) for r in ROLES}

    for gen in tqdm(range(start, args.generations), desc="Training Generations", initial=start,
                    total=args.generations):
        eng.step(sync=False)
        last = gen == args.generations - 1
        if getattr(args, "log_weight_stats", True):
            # statistics of the base weights, kept on the device until plotted (the reference logs them per
            # perturbed member, MPE/mpe_agent.py:30-50: args.log_member_weight_stats restores that, K5 + one
            # reduction kernel per role, in eng.weight_stats)
            for r in ROLES:
                w = eng.theta[r][pidx[r]]
                wlog[r].append((gen, torch.stack([w.mean(), w.min(), w.max(), w.std()])))
        if args.fitness_sharing:
            for r in ROLES:
                diversity[r].append(eng.diversity[r])
        stop = eng.should_stop() if args.early_stopping else 0
        if stop:
            hs = eng.host_state()
            name = {"agent_0": "agent_0", "agent_1": "agent_1", "adversary_0": "adversary"}[ROLES[stop - 1]]
            print(f"Early stopping triggered at generation {gen} for {name}. "
                  f"Best reward: {hs['best'][ROLES[stop - 1]]}")
            break
        if (args.save or last) and comm.rank == 0:
            for r in ROLES:
                agents[r].model.load_flat_row(eng.theta[r])
                if args.save:
                    save_model(agents[r], files[r])
        if args.save:
            torch.save(eng.state_dict(), checkpoint_path(output_dir))
        if want_plots:
            hs = eng.host_state()
            div_host = {r: [float(d) for d in diversity[r]] if args.fitness_sharing else None for r in ROLES}
            for r in ROLES:
                rew = list(hs["rewards"][r])
                fit = [x / (1 + d) for x, d in zip(rew[start:], div_host[r])] if args.fitness_sharing else None
                plot_experiment_metrics(rewards=rew,
                                        mutation_power_history=list(hs["sigma_history"][r]) if args.adaptive else None,
                                        fitness=fit, diversity=div_host[r], file_path=plots[r], args=args)
            logs = {r: [{"step": g, "mean": float(t[0]), "min": float(t[1]), "max": float(t[2]), "std": float(t[3])}
                        for g, t in wlog[r]] for r in ROLES}
            plot_weights_logging(weights_plot, logs["agent_0"], logs["agent_1"], logs["adversary_0"])
    eng.check_status()
    eng.write_back_args()                 # the reference leaves the adapted sigmas in `args`
    for r in ROLES:
        agents[r].model.load_flat_row(eng.theta[r])
    args._es_engine = eng
    return agents["agent_0"], agents["agent_1"], agents["adversary_0"]

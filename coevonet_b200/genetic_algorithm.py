"""Co-GA driver with the reference's entry points (``genetic_algorithm.py``).

``genetic_algorithm_train(env, agent, args, output_dir)`` keeps the reference's
signature, hyper-parameter bag, per-generation side effects (6 ``.pth`` files,
plots) and -- under ``args.reference_compat`` (default) -- its
behaviour-defining quirks.  The three evaluation loops,
selection, HoF update and re-population run on the device through
``engine.GAEngine``; nothing in the loop touches a per-agent Python object.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from tqdm import tqdm

from . import engine as _engine
from . import layout
from .utils.game_logic_functions import create_agent, play_game
from .utils.utils_pth_and_plots import plot_experiment_metrics, save_model

ROLES = layout.ROLES


def evaluate_current_weights(agent_0, agent_1, adversary, env, args):
    """Mean reward triple of 10 evaluation games (genetic_algorithm.py:12-29)."""
    tot = np.zeros(3)
    for _ in range(_engine.N_EVAL_GAMES):
        tot += play_game(env=env, player1=agent_0.model, player2=agent_1.model,
                         adversary=adversary.model, args=args, eval=True)
    return tuple(tot / _engine.N_EVAL_GAMES)


def mutate_elites(env, elites, args, role):
    """``population - 1`` mutated clones of the elites, elite ``i % elites_number``
    for child ``i`` (genetic_algorithm.py:32-48), produced by ONE K3 launch."""
    from .utils.game_logic_functions import _device
    dev = _device()
    sigma = {"agent_0": args.mutation_power_agent_0, "agent_1": args.mutation_power_agent_1}.get(
        role, args.mutation_power_adversary)
    in_dim = layout.OBS_DIM[role]
    rows = layout.pack_models([e.model for e in elites[:args.elites_number]], in_dim, dev)
    n = args.population - 1
    gen = getattr(args, "_mutate_calls", 0)
    args._mutate_calls = gen + 1
    from . import ops
    children = ops.ga_repopulate(rows, layout.fc_dim(in_dim), sigma, getattr(args, "seed", 1870300),
                                 role, 0x80000000 + gen, 1, n).cpu()
    out = []
    for i in range(n):
        child = create_agent(env, args, role)
        child.model.load_flat_row(children[i])
        out.append(child)
    return out


def _rows_to_agents(rows, env, args, role):
    agents = []
    for row in rows.cpu():
        a = create_agent(env, args, role)
        a.model.load_flat_row(row)
        agents.append(a)
    return agents


def build_initial_state(env, args):
    """Initial HoF, (discarded) elite placeholders and populations created in the
    reference's order (genetic_algorithm.py:63-68,110-117) so that, under the same
    ``torch.manual_seed``, every founder has the reference's weights."""
    hof = {}
    hof["agent_1"] = [create_agent(env, args, "agent_1") for _ in range(args.hof_size)]
    hof["agent_0"] = [create_agent(env, args, "agent_0") for _ in range(args.hof_size)]
    hof["adversary_0"] = [create_agent(env, args, "adversary_0") for _ in range(args.hof_size)]
    # elites_* placeholders: created (they consume the init RNG) and overwritten later
    for role in ("agent_1", "agent_0", "adversary_0"):
        for _ in range(args.hof_size):
            create_agent(env, args, role)
    pop = {r: [] for r in ROLES}
    for _ in tqdm(range(args.population), desc=f"Creating initial population (n = {args.population})", leave=False):
        for role in ROLES:
            pop[role].append(create_agent(env, args, role))
    return hof, pop


def genetic_algorithm_train(env, agent, args, output_dir):
    """Reference signature (genetic_algorithm.py:51); ``agent`` is unused there too."""
    files = {r: (os.path.join(output_dir, f"hall_of_fame_{n}.pth"), os.path.join(output_dir, f"elite_weights_{n}.pth"))
             for r, n in (("agent_0", "agent_0"), ("agent_1", "agent_1"), ("adversary_0", "adversary"))}
    plots = {"agent_0": os.path.join(output_dir, "results_agent_0._plot.png"),
             "agent_1": os.path.join(output_dir, "results_agent_1_plot.png"),
             "adversary_0": os.path.join(output_dir, "results_adversary_plot.png")}

    from .utils.game_logic_functions import _device
    dev = _device()
    comm = _engine.Comm()
    shard = _engine.Shard(args.population, comm.rank, comm.world)
    sl = slice(shard.row0, shard.row0 + shard.n_local)
    if getattr(args, "device_init", False):
        # founders drawn on the device (same distribution as PyTorch's default init, Philox
        # stream, global member ids): no 3*P nn.Module constructions on the host
        from . import ops
        seed, P, H = getattr(args, "seed", 1870300), args.population, args.hof_size
        pop_rows = {r: ops.fc_init(layout.OBS_DIM[r], seed, r, shard.row0, shard.n_local, dev) for r in ROLES}
        hof_rows = {r: ops.fc_init(layout.OBS_DIM[r], seed, r, P, H, dev) for r in ROLES}
        # Appendix C #3: diversity is measured against the LAST founder created for each role
        founder = {r: ops.fc_init(layout.OBS_DIM[r], seed, r, P - 1, 1, dev)[0] for r in ROLES}
        if comm.rank == 0:
            for name, role in (("agent_0", "agent_1"), ("agent_1", "agent_0"), ("adversary", "adversary_0")):
                print(f"\nNumber of parameters for {name} network: {layout.fc_dim(layout.OBS_DIM[role])}")
    else:
        # every rank draws the founders from rank 0's generator state (unseeded runs included), so the
        # shards are the row blocks of ONE population and the replicated rows agree
        _engine.sync_torch_rng(comm, dev)
        hof_agents, pop_agents = build_initial_state(env, args)
        if comm.rank == 0:
            for name, role in (("agent_0", "agent_1"), ("agent_1", "agent_0"), ("adversary", "adversary_0")):
                n_par = sum(p.numel() for p in hof_agents[role][0].model.parameters())
                print(f"\nNumber of parameters for {name} network: {n_par}")
        pop_rows = {r: layout.pack_models([a.model for a in pop_agents[r][sl]], layout.OBS_DIM[r]) for r in ROLES}
        hof_rows = {r: layout.pack_models([a.model for a in hof_agents[r]], layout.OBS_DIM[r]) for r in ROLES}
        # Appendix C #3: diversity is measured against the LAST founder created for each role
        founder = {r: layout.pack_models([pop_agents[r][-1].model], layout.OBS_DIM[r])[0] for r in ROLES}
        del pop_agents
    eng = _engine.GAEngine(args, dev, pop_rows, hof_rows, founder, env=env, comm=comm)
    ckpt = os.path.join(output_dir, f"engine_state_rank{comm.rank}.pt")
    start = 0
    if getattr(args, "resume", False) and os.path.isfile(ckpt):
        eng.load_state_dict(torch.load(ckpt, weights_only=False))
        start = eng.gen
        if comm.rank == 0:
            print(f"Resuming from generation {start} ({ckpt})")

    diversity = {r: [] if args.fitness_sharing else None for r in ROLES}
    want_plots = comm.rank == 0 and getattr(args, "plots", True)

    for gen in tqdm(range(start, args.generations), desc="Generations", initial=start, total=args.generations):
        eng.step(sync=False)
        if args.save and comm.rank == 0:
            for r in ROLES:
                save_model(_rows_to_agents(eng.hof[r], env, args, r), files[r][0])
                save_model(_rows_to_agents(eng.elites[r], env, args, r), files[r][1])
        if args.save:
            torch.save(eng.state_dict(), ckpt)
        if args.fitness_sharing:
            for r in ROLES:
                diversity[r].append(eng.diversity[r])
        if want_plots:
            hs = eng.host_state()
            for r in ROLES:
                rew = list(hs["rewards"][r])
                div = [float(d) for d in diversity[r]] if args.fitness_sharing else None
                fit = [x / (1 + d) for x, d in zip(rew[start:], div)] if args.fitness_sharing else None
                plot_experiment_metrics(rewards=rew,
                                        mutation_power_history=list(hs["sigma_history"][r]) if args.adaptive else None,
                                        fitness=fit, diversity=div, file_path=plots[r], args=args)
    eng.check_status()
    eng.write_back_args()          # the reference leaves the adapted sigmas in `args`
    args._ga_engine = eng          # handle for callers that want the final device state
    return None

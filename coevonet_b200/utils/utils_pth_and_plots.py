"""Persistence / plotting helpers with the reference's names
(``utils/utils_pth_and_plots.py``).  Checkpoints keep the reference's on-disk
convention: ``torch.save`` of agent objects (lists of ``MPEAgent`` for GA HoF /
elites, single agents for ES).  Plots need matplotlib and are skipped without it."""
from __future__ import annotations

import os

import torch


def save_model(obj, file_path):
    torch.save(obj, file_path)


def _load(path):
    return torch.load(path, weights_only=False)


def load_agent_for_testing(args, env=None):
    if args.algorithm == "GA":
        paths = (args.GA_hof_to_test_agent_0, args.GA_hof_to_test_agent_1, args.GA_hof_to_test_adversary)
        for name, p in zip(("agent_0", "agent_1", "adversary_0"), paths):
            if p is None:
                raise ValueError(f"Error: Model file for {name} not specified. Please specify the agent to test")
            if not os.path.exists(p):
                raise ValueError(f"Error: Model file {p} not found.")
        print("Loading Agents for testing...")
        return tuple(_load(p)[-1] for p in paths)
    if args.algorithm == "ES":
        paths = (args.ES_model_to_test_agent_0, args.ES_model_to_test_agent_1, args.ES_model_to_test_adversary_0)
        for name, p in zip(("agent_0", "agent_1", "adversary_0"), paths):
            if p is None:
                raise ValueError(f"Error: Model file for {name} not specified. Please specify the agent to test")
            if not os.path.exists(p):
                raise ValueError(f"Error: Model file {p} not found.")
        print("Loading Agents for testing...")
        return tuple(_load(p) for p in paths)
    raise ValueError(f"unknown algorithm {args.algorithm}")


def create_output_dir(args):
    """Same directory naming as ``utils/utils_pth_and_plots.py:83-96``."""
    dir_name = (f"{args.algorithm}_models/gens{args.generations}_pop{args.population}_hof{args.hof_size}"
                f"_game{args.game}_tslimit{args.max_timesteps_per_episode}"
                f"_fitness-sharing{args.fitness_sharing}_adaptive{args.adaptive}")
    if args.adaptive:
        dir_name += f"max_mutation{args.max_mutation_power}_min_mutation{args.min_mutation_power}"
    if args.algorithm == "ES":
        dir_name += f"_lr{args.learning_rate}"
    os.makedirs(dir_name, exist_ok=True)
    return dir_name


def _plt():
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        return plt
    except Exception:
        return None


def plot_weights_logging(file_path, weights_logging_agent_0, weights_logging_agent_1, weights_logging_adversary):
    plt = _plt()
    if plt is None:
        return
    for data, title, suffix in ((weights_logging_agent_0, "Agent 0", "agent_0"),
                                (weights_logging_agent_1, "Agent 1", "agent_1"),
                                (weights_logging_adversary, "Adversary", "adversary")):
        if not data:
            continue
        steps = [d["step"] for d in data]
        plt.figure(figsize=(10, 6))
        for key, style in (("mean", "-"), ("min", "--"), ("max", ":")):
            plt.plot(steps, [d[key] for d in data], linestyle=style, label=key)
        plt.title(f"Weight Statistics Over Generations ({title})")
        plt.xlabel("Step")
        plt.ylabel("Weight Value")
        plt.legend()
        plt.savefig(f"{file_path}_{suffix}.png")
        plt.close()


def plot_experiment_metrics(rewards=None, mutation_power_history=None, fitness=None, diversity=None,
                            file_path="experiment_metrics.png", args=None):
    plt = _plt()
    if plt is None:
        return
    series = [(n, s) for n, s in (("Evaluation reward", rewards), ("Mutation power", mutation_power_history),
                                  ("Fitness", fitness), ("Diversity", diversity)) if s]
    if not series:
        return
    fig, axes = plt.subplots(len(series), 1, figsize=(10, 4 * len(series)), squeeze=False)
    for ax, (name, s) in zip(axes[:, 0], series):
        ax.plot(range(len(s)), s)
        ax.set_title(name)
        ax.set_xlabel("Generation")
        ax.grid(True, linestyle="--", alpha=0.7)
    fig.tight_layout()
    fig.savefig(file_path)
    plt.close(fig)

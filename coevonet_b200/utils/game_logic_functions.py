"""Rollout / game-logic surface of the reference (``utils/game_logic_functions.py``)
on the B200 kernels: same names, argument order and return conventions."""
from __future__ import annotations

import numpy as np
import torch

from .. import layout, ops
from . import mpe_spec


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("coevonet_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def diversity_penalty(individual_weights, population_weights, args, sigma=None):
    """Fitness-sharing score (``utils/game_logic_functions.py:12-37``):
    d_i = ||p_i - w||, sigma = mean(d) unless given, sum(max(0, 1 - d_i/sigma)).

    Inputs are flat PERTURBABLE weight vectors like the reference's
    ``get_weights_ES()`` output; distances run on the K7 kernel."""
    dev = _device()
    ref = np.asarray(individual_weights, dtype=np.float32)
    in_dim = {len(layout.fc_perturbable_index(8)): 8, len(layout.fc_perturbable_index(10)): 10}.get(ref.size)
    if in_dim is None:
        raise ValueError("diversity_penalty expects flat FCNetwork perturbable weights")
    pidx = torch.from_numpy(layout.fc_perturbable_index(in_dim))
    pitch = layout.fc_pitch(in_dim)
    pop = torch.zeros((len(population_weights), pitch), dtype=torch.float32)
    pop[:, pidx] = torch.from_numpy(np.stack([np.asarray(p, dtype=np.float32) for p in population_weights]))
    row = torch.zeros(pitch, dtype=torch.float32)
    row[pidx] = torch.from_numpy(ref)
    dist = ops.diversity_dist(pop.to(dev), row.to(dev), in_dim)
    if sigma is None:
        sigma = dist.mean()
    if getattr(args, "debug", False):
        print("Distance Range:", float(dist.min()), float(dist.max()), "Sigma:", float(sigma))
    return float(torch.clamp(1 - dist / sigma, min=0).sum())


def initialize_env(args):
    """Env factory (``utils/game_logic_functions.py:41-55``): returns the device
    env handle, seeded with the reference's constant 1870300."""
    if args.game != "simple_adversary_v3":
        # no ROMs here and the reference's own Atari loop never runs: the
        # wrapper chain of utils/game_logic_functions.py:48-53 over the synthetic emulator (csrc/atari_synth.cu)
        env = DeviceAtariEnv(args.game)
        env.reset(seed=mpe_spec.ENV_SEED)
        return env
    env = mpe_spec.DeviceMPEEnv(render_mode="human" if getattr(args, "render", False) else None)
    env.reset(seed=mpe_spec.ENV_SEED)
    return env


class _Space:
    def __init__(self, shape=None, n=None):
        self.shape, self.n = shape, n


class DeviceAtariEnv:
    """Env handle of the synthetic Atari-like game: the attributes the reference reads from its wrapped
    PettingZoo env (agents, observation_space(a).shape = (84, 84, 6), action_space(a).n = 6 for pong_v3 /
    18 for boxing_v2) and the episode counter that keys the synthetic emulator."""

    N_ACTIONS = {"pong_v3": 6, "boxing_v2": 18}

    def __init__(self, game):
        self.game = game
        self.possible_agents = ["first_0", "second_0"]
        self.agents = list(self.possible_agents)
        self.seed_value = mpe_spec.ENV_SEED
        self.episode = -1

    def observation_space(self, agent):
        return _Space(shape=(84, 84, 6))

    def action_space(self, agent):
        return _Space(n=self.N_ACTIONS[self.game])

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.seed_value, self.episode = int(seed), -1
        self.episode += 1

    def close(self):
        pass


def play_atari(env, player1, player2, args, eval=False):
    """The reference's Atari episode loop (``utils/game_logic_functions.py:84-119``) with its three
    defects repaired (argument count, ``forward`` signature, ``len(actions)``)
    on the synthetic emulator: returns (rewards['first_0'], rewards['second_0'])."""
    from ..atari_rollout import atari_rollout
    dev = _device()
    limit = args.max_evaluation_steps if eval else args.max_timesteps_per_episode
    if limit is None:
        raise ValueError("the synthetic Atari game never terminates: an agent-step limit is required")
    n_act = env.action_space(env.agents[0]).n
    r1, r2 = atari_rollout(0, player1.flat_row(dev).unsqueeze(0), player2.flat_row(dev), n_act, 1, limit,
                           env.seed_value, ep0=env.episode,
                           reference_compat=getattr(args, "reference_compat", True))
    return float(r1), float(r2)


def create_agent(env, args, role=None):
    from ..MPE.mpe_agent import MPEAgent
    if args.game == "simple_adversary_v3":
        return MPEAgent(env, args, role)
    if args.game in ("pong_v3", "boxing_v2"):
        from ..Atari.atari_agent import AtariAgent
        return AtariAgent(env, args)
    raise ValueError(f"Unsupported game type: {args.game}")


def preprocess_observation(obs, args):
    obs = torch.from_numpy(np.asarray(obs)).to(torch.float32)
    if args.game != "simple_adversary_v3":
        obs = obs.permute(2, 0, 1).unsqueeze(0)
    return obs


def play_game(env, player1, player2, adversary=None, args=None, eval=False):
    """One episode (``utils/game_logic_functions.py:215-228``): ``player1`` in
    agent_0's seat, ``player2`` in agent_1's, ``adversary`` in adversary_0's.
    Returns the reference's triple ``(agent_0, agent_1, adversary_0)`` with its
    rotated attribution unless ``args.reference_compat``
    is False."""
    env.reset()
    if args.game != "simple_adversary_v3":
        return play_atari(env, player1, player2, args, eval)       # utils/game_logic_functions.py:224-227
    if adversary is None:
        raise ValueError("adversary not specified")
    dev = _device()
    for m in (player1, player2, adversary):
        if not hasattr(m, "flat_row"):
            raise TypeError("play_game runs on the device kernels and needs FCNetwork players "
                            f"(got {type(m).__name__})")
    limit = args.max_evaluation_steps if eval else args.max_timesteps_per_episode
    init = torch.from_numpy(env.take_pending()).reshape(1, 1, 1, mpe_spec.INIT_STATE_DIM).to(dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    out = ops.mpe_rollout("agent_0", player1.flat_row(dev).unsqueeze(0), adversary.flat_row(dev).unsqueeze(0),
                          player2.flat_row(dev).unsqueeze(0), init, n_cycles=ops.cycles_for_limit(limit),
                          pos_first=getattr(args, "integrate_pos_first", True), status=status)
    ops.raise_on_status(status)
    s0, s1, sadv = ops.reward_slots(out.reshape(4), agent_step_limit=limit,
                                    reference_compat=getattr(args, "reference_compat", True))
    return float(s0), float(s1), float(sadv)

"""``AtariAgent`` with the reference's interface (``Atari/atari_agent.py``)."""
from __future__ import annotations

import torch.optim as optim

from ..agent import Agent
from .deepqn import DeepQN


class AtariAgent(Agent):
    def __init__(self, env, args):
        self.input_channels = env.observation_space(env.agents[0]).shape[-1]
        self.n_actions = env.action_space(env.agents[0]).n
        self.model = DeepQN(self.input_channels, self.n_actions, args.precision)
        self.optimizer = optim.Adam(self.model.parameters(), lr=0.0001)
        super().__init__(self.model, self.optimizer, args)

    def clone(self, env, args, role=None):
        clone = AtariAgent(env, args)
        clone.model.load_state_dict(self.model.state_dict())
        return clone

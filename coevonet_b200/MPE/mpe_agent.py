"""``MPEAgent`` with the reference's interface (``MPE/mpe_agent.py``)."""
from __future__ import annotations

import torch.optim as optim

from ..agent import Agent
from .fcnetwork import FCNetwork


class MPEAgent(Agent):
    def __init__(self, env, args, role):
        self.input_channels = env.observation_space(role).shape[-1]
        self.n_actions = env.action_space(role).n
        self.model = FCNetwork(self.input_channels, self.n_actions, args.precision)
        # the reference builds an (unused) Adam per agent (MPE/mpe_agent.py:20); kept so that
        # pickled checkpoints have the same attributes
        self.optimizer = optim.Adam(self.model.parameters(), lr=0.0001)
        super().__init__(self.model, self.optimizer, args)

    def clone(self, env, args, role):
        clone = MPEAgent(env, args, role)
        clone.model.load_state_dict(self.model.state_dict())
        return clone

    def log_weight_statistics(self, step, weights_logging_agent_0=None, weights_logging_agent_1=None,
                              weights_logging_adversary=None, role=None):
        """mean / std / min / max of the perturbable weights (``MPE/mpe_agent.py:30-50``)."""
        w = self.model.get_perturbable_weights()
        entry = {"step": step, "mean": w.mean(), "min": w.min(), "max": w.max(), "std": w.std()}
        target = {"agent_0": weights_logging_agent_0, "agent_1": weights_logging_agent_1,
                  "adversary_0": weights_logging_adversary}.get(role)
        if target is not None:
            target.append(entry)

"""Base ``Agent`` with the reference's interface (``agent.py``).

Single-agent mutation helpers are kept for API compatibility; the population
path mutates whole ``[P, pitch]`` tensors with the K3/K5 kernels instead.  Both
use the same counter-based Philox noise (the reference draws from unseeded
global generators and is irreproducible).
"""
from __future__ import annotations

import itertools

import numpy as np
import torch

from . import layout, ops

_counter = itertools.count(1)


class Agent:
    #: run seed of the Philox noise streams (the constant the reference seeds its env with)
    noise_seed = 1870300

    def __init__(self, model, optimizer, args):
        self.model = model
        self.optimizer = optimizer
        self.precision = args.precision

    def apply_precision(self, model, precision):
        if self.precision == "float32":
            self.model = self.model.float()
        else:
            raise ValueError(f"Unsupported precision: {self.precision}")

    def _device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("Agent mutation runs on the CUDA kernels; no GPU is visible")
        return torch.device("cuda", torch.cuda.current_device())

    def mutate(self, noise_std):
        """GA mutation of EVERY parameter tensor, LayerNorm included
        (``agent.py:25-29``), through the K3 kernel."""
        dev = self._device()
        in_dim = self.model.input_channels
        row = self.model.flat_row(dev).unsqueeze(0)
        member = next(_counter)          # member id 0 is "unmutated" in K3
        child = ops.ga_repopulate(row, layout.fc_dim(in_dim), noise_std, self.noise_seed,
                                  0, 0xFFFFFFFF, member, 1)
        self.model.load_flat_row(child[0])

    def mutate_ES(self, args, role, step, weights_logging_agent_0, weights_logging_agent_1,
                  weights_logging_adversary):
        """ES mutation of the Linear parameters only (``agent.py:31-70``) through the
        K5 kernel; returns the applied noise (already scaled by sigma) as the
        reference does."""
        sigma = {"agent_0": args.mutation_power_agent_0, "agent_1": args.mutation_power_agent_1,
                 "adversary_0": args.mutation_power_adversary}[role]
        dev = self._device()
        in_dim = self.model.input_channels
        pitch = layout.fc_pitch(in_dim)
        noise = torch.zeros((1, pitch), dtype=torch.float32, device=dev)
        rows = ops.es_perturb(self.model.flat_row(dev), in_dim, sigma, self.noise_seed, role,
                              0xFFFFFFFF, next(_counter), 1, noise_out=noise)
        self.model.load_flat_row(rows[0])
        pidx = layout.fc_perturbable_index(in_dim)
        scaled = (np.float32(sigma) * noise[0].cpu().numpy()[pidx]).astype(np.float32)
        self.log_weight_statistics(step=step, weights_logging_agent_0=weights_logging_agent_0,
                                   weights_logging_agent_1=weights_logging_agent_1,
                                   weights_logging_adversary=weights_logging_adversary, role=role)
        return scaled.astype(np.float64)

    def set_weights(self, weights):
        self.model.load_state_dict(weights)

    def get_weights(self):
        return self.model.state_dict()

    def clone(self, args):
        raise NotImplementedError("The clone method should be implemented by the specific agent type.")

"""``FCNetwork`` with the reference's interface (``MPE/fcnetwork.py``).

The module is the host-side container of one policy's weights: same layers,
same registration order (so ``parameters()`` / ``state_dict`` / default
initialisation under ``torch.manual_seed`` are identical to the reference) and
the same flat-weight get/set helpers.  Its arithmetic does not run here:
``forward`` / ``determine_action`` pack the weights into a flat row and call the
CUDA kernel ``cev_fc_forward_f32``; the population hot path never goes through
this class at all (it works on ``[P, pitch]`` device tensors).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import layout, ops


class FCNetwork(nn.Module):
    """obs -> 512 -> LN -> ReLU -> 256 -> LN -> ReLU -> n_actions
    (reference ``MPE/fcnetwork.py:9-22``)."""

    def __init__(self, input_channels, n_actions, precision):
        super().__init__()
        if precision == "float16":
            # the reference keeps LayerNorm in fp32 under --precision float16, which makes
            # its own forward fail on dtype mismatch; not supported here
            raise ValueError("precision float16 is not supported by the B200 path (fp32 only)")
        if n_actions != layout.NACT or input_channels not in (8, 10):
            raise ValueError("FCNetwork kernels are built for simple_adversary_v3 (8/10 inputs, 5 actions)")
        self.dtype = torch.float32
        self.input_channels = input_channels
        self.layers = []
        self.fc1 = nn.Linear(input_channels, layout.H1)
        self.ln1 = nn.LayerNorm(layout.H1)
        self.layers.append(self.fc1)
        self.fc2 = nn.Linear(layout.H1, layout.H2)
        self.ln2 = nn.LayerNorm(layout.H2)
        self.layers.append(self.fc2)
        self.output = nn.Linear(layout.H2, n_actions)
        self.layers.append(self.output)

    # ---- flat row view ---------------------------------------------------
    def flat_row(self, device=None):
        """float32[pitch] row in the device layout (``coevonet_b200.layout``)."""
        row = layout.pack_state_dict(self.state_dict(), self.input_channels)
        return row.to(device) if device is not None else row

    def load_flat_row(self, row):
        self.load_state_dict(layout.unpack_to_state_dict(row, self.input_channels))

    # ---- forward / action (device kernel) ---------------------------------
    def forward(self, x, args=None):
        """Logits for one observation (1-D) or a batch (2-D); raises ``ValueError``
        on non-finite inputs/activations like ``MPE/fcnetwork.py:39-65``."""
        x = torch.as_tensor(x, dtype=torch.float32)
        single = x.dim() == 1
        obs = x.reshape(-1, self.input_channels).contiguous()
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if dev is None:
            raise RuntimeError("FCNetwork.forward runs on the CUDA kernel; no GPU is visible "
                               "(coevonet_b200 has no CPU fallback)")
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        logits, _ = ops.fc_forward(self.flat_row(dev).unsqueeze(0), self.input_channels,
                                   obs.to(dev), status=status)
        ops.raise_on_status(status)
        logits = logits.cpu()
        return logits[0] if single else logits

    def determine_action(self, inputs, args=None):
        """Arg-max with the reference's strict ``>`` scan (lowest index on ties,
        ``MPE/fcnetwork.py:73-90``)."""
        actions = self.forward(inputs, args)
        best, pos = -float("inf"), -1
        for i in range(len(actions)):
            if actions[i] > best:
                pos, best = i, actions[i]
        if pos == -1:
            raise ValueError(f"ERROR: current_best_position = {pos} after checking for best action, "
                             f"the action probabilities are {actions}")
        return pos

    # ---- state_dict style (GA) ------------------------------------------------
    def get_weights(self, layers=None):
        sd = self.state_dict()
        if layers is None:
            return {k: v.clone() for k, v in sd.items()}
        return {k: v.clone() for k, v in sd.items() if any(k.startswith(name) for name in layers)}

    def set_weights(self, new_weights, layers=None):
        current = self.state_dict()
        target = {k: current[k] for k in new_weights.keys()} if layers is not None else current
        for key in target:
            if key not in new_weights:
                raise ValueError(f"Missing key in new_weights: {key}")
            if new_weights[key].shape != target[key].shape:
                raise ValueError(f"Shape mismatch for key '{key}': expected {target[key].shape}, "
                                 f"got {new_weights[key].shape}")
        self.load_state_dict({k: new_weights[k] for k in target}, strict=False)

    # ---- flat style (ES) ---------------------------------------------------------
    def get_perturbable_layers(self):
        return [m for name, m in self.named_modules() if name and not isinstance(m, nn.LayerNorm)]

    def get_weights_ES(self, layers=None):
        layers = layers if layers else self.layers
        parts = []
        for layer in layers:
            parts.append(layer.weight.detach().cpu().numpy().reshape(-1))
            if layer.bias is not None:
                parts.append(layer.bias.detach().cpu().numpy().reshape(-1))
        return np.concatenate(parts)

    def get_perturbable_weights(self):
        return self.get_weights_ES(self.get_perturbable_layers())

    def set_weights_ES(self, flat_weights, args=None, layers=None):
        layers = self.get_perturbable_layers() if layers is None else layers
        flat = np.asarray(flat_weights)
        i = 0
        for layer in layers:
            n = layer.weight.numel()
            layer.weight.data.copy_(torch.tensor(flat[i:i + n].reshape(tuple(layer.weight.shape)),
                                                 dtype=torch.float32))
            i += n
            if layer.bias is not None:
                n = layer.bias.numel()
                layer.bias.data.copy_(torch.tensor(flat[i:i + n], dtype=torch.float32))
                i += n

    def set_perturbable_weights(self, weights_to_set, args=None):
        self.set_weights_ES(weights_to_set, args, self.get_perturbable_layers())

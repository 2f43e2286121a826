"""``DeepQN`` with the reference's interface (``Atari/deepqn.py``): the host-side
weight container of the Nature-DQN policy; its forward runs on the K2 kernel
(``cev_deepqn_forward``), per-frame train-mode BatchNorm included (SURVEY.md
Appendix C #12)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import layout, ops


class DeepQN(nn.Module):
    def __init__(self, input_channels, n_actions, precision):
        super().__init__()
        if precision == "float16":
            raise ValueError("precision float16 is not supported by the B200 path (fp32 rows)")
        self.dtype = torch.float32
        self.input_channels, self.n_actions = input_channels, n_actions
        self.layers = []
        self.conv1 = nn.Conv2d(input_channels, 32, kernel_size=8, stride=4)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=4, stride=2)
        self.conv3 = nn.Conv2d(64, 64, kernel_size=3, stride=1)
        self.fc1 = nn.Linear(64 * 7 * 7, 512)
        self.output = nn.Linear(512, n_actions)
        self.vbn1 = nn.BatchNorm2d(32)
        self.vbn2 = nn.BatchNorm2d(64)
        self.vbn3 = nn.BatchNorm2d(64)
        self.layers += [self.conv1, self.conv2, self.conv3, self.fc1, self.output, self.vbn1, self.vbn2, self.vbn3]

    def flat_row(self, device=None):
        row = layout.pack_dqn_state_dict(self.state_dict(), self.input_channels, self.n_actions)
        return row.to(device) if device is not None else row

    def forward(self, x, args=None):
        """x: [B, C, 84, 84] (0..255) -> logits [B, A]; every frame is normalised on
        its own like the reference's batch-1 calls."""
        if not torch.cuda.is_available():
            raise RuntimeError("DeepQN.forward runs on the CUDA kernel; no GPU is visible")
        dev = torch.device("cuda", torch.cuda.current_device())
        x = torch.as_tensor(x)
        frames = x.round().clamp(0, 255).to(torch.uint8).reshape(1, -1, self.input_channels, 84, 84)
        logits, _ = ops.deepqn_forward(self.flat_row(dev).unsqueeze(0), frames.to(dev).contiguous(),
                                       self.input_channels, self.n_actions)
        return logits[0].cpu()

    def determine_action(self, inputs, args=None):
        actions = self.forward(inputs, args)[0]
        best, pos = -float("inf"), -1
        for i in range(len(actions)):
            if actions[i] > best:
                pos, best = i, actions[i]
        return pos

    def get_weights(self, layers=None):
        sd = self.state_dict()
        if layers is None:
            return {k: v.clone() for k, v in sd.items()}
        return {k: v.clone() for k, v in sd.items() if any(k.startswith(n) for n in layers)}

    def set_weights(self, new_weights, layers=None):
        self.load_state_dict(new_weights, strict=layers is None)

    def get_perturbable_layers(self):
        return [m for name, m in self.named_modules() if name and not isinstance(m, nn.BatchNorm2d)]

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
            layer.weight.data.copy_(torch.tensor(flat[i:i + n].reshape(tuple(layer.weight.shape)), dtype=torch.float32))
            i += n
            if layer.bias is not None:
                n = layer.bias.numel()
                layer.bias.data.copy_(torch.tensor(flat[i:i + n], dtype=torch.float32))
                i += n

    def set_perturbable_weights(self, weights_to_set, args=None):
        self.set_weights_ES(weights_to_set, args, self.get_perturbable_layers())

"""Max abs error of the device normals vs the oracle's (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from coevonet_b200 import layout, ops
from oracle import philox
P, in_dim = 64, 10
pitch = layout.fc_pitch(in_dim); D = layout.fc_dim(in_dim)
theta = torch.zeros(pitch, device="cuda")
noise = torch.empty((P, pitch), device="cuda")
ops.es_perturb(theta, in_dim, 1.0, 1870300, "agent_0", 2, 0, P, noise_out=noise)
z = philox.normals(1870300, philox.KIND_ES, philox.ROLE_ID["agent_0"], 2, np.arange(P), D)
zd = noise.cpu().numpy()[:, :D]
mask = zd != 0
err = np.abs(zd - z)[mask]
print("max abs err", err.max(), "mean", err.mean(), "p99.99", np.quantile(err, 0.9999), "n", err.size)
i = np.argmax(np.abs(zd - z) * mask); print("worst at", np.unravel_index(i, zd.shape), zd.flat[i], z.flat[i])

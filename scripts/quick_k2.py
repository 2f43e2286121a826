"""Quick K2 (DeepQN forward) timing sweep (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from coevonet_b200 import layout, ops
c_in, n_act = 4, 6
pitch = layout.dqn_pitch(c_in, n_act)
for P, B in ((592, 1), (592, 4), (592, 12), (2048, 1)):
    members = (torch.rand((P, pitch), device="cuda") - 0.5) * 0.05
    frames = ops.random_frames(1, (P, B, c_in, 84, 84), "cuda")
    for _ in range(2): ops.deepqn_forward(members, frames, c_in, n_act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ops.deepqn_forward(members, frames, c_in, n_act)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    gb = P * layout.dqn_dim(c_in, n_act) * 4 / 1e9 + P * B * c_in * 7056 / 1e9
    print(f"P={P} B={B}: {ms:.3f} ms  {P*B/ms*1e3:.0f} forwards/s  {P*B*18.69e6/ms/1e9:.2f} TFLOP/s  "
          f"{gb/ms*1e3:.0f} GB/s algorithmic")

#!/usr/bin/env python
"""BASELINE configs 4 and 5 (DeepQN, synthetic 84x84x4 frames): not the driver's bench, a measurement aid.

  config 4  pong_v3 Co-GA policy forward: P members x B frames through K2 (conv + fc on tcgen05)
  config 5  boxing_v2 Co-ES step at population 4096 sharded over the ranks: K5 perturb (DeepQN rows) -> K2
            forward -> synthetic fitness (the reference's Atari rollout is dead code, SURVEY.md Appendix C
            #9-11) -> fitness all-gather -> K6 from the members -> all-reduce of delta (6.77 MB) -> apply

    python scripts/bench_dqn.py                                  # 1 GPU
    torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_dqn.py
One JSON line per config on rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from coevonet_b200 import layout, ops

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)

def timed(fn, n, warm=3, collective=True):
    """Device time per call; with ``collective`` every rank must call it (barriers + max over ranks)."""
    sync = collective and world > 1
    for _ in range(warm): fn()
    if sync: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record()
    if sync: dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    if sync: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

# ---- config 4: pong forward (C=4, A=6) ------------------------------------------------------------
if rank == 0:
    c_in, n_act = 4, 6
    pitch, D = layout.dqn_pitch(c_in, n_act), layout.dqn_dim(c_in, n_act)
    for P, B in ((2048, 1), (2048, 4)):
        members = (torch.rand((P, pitch), device=dev) - 0.5) * 0.05
        frames = ops.random_frames(1, (P, B, c_in, 84, 84), dev)
        ms = timed(lambda: ops.deepqn_forward(members, frames, c_in, n_act), 5, collective=False)   # rank 0 only
        gb = (P * D * 4 + P * B * c_in * 7056) / 1e9
        print(json.dumps({"config": "4: pong_v3 DeepQN forward, synthetic frames", "members": P, "frames_per_member": B,
                          "ms": ms, "forwards_per_s": P * B / ms * 1e3, "algorithmic_GBps": gb / ms * 1e3,
                          "algorithmic_TFLOPs": P * B * 18.69e6 / ms / 1e9}), flush=True)
        del members, frames
if world > 1: dist.barrier()

# ---- config 5: boxing ES step (C=4, A=18), population 4096 over the ranks ---------------------------
c_in, n_act, P, B, sigma, lr, seed = 4, 18, 4096, 1, 0.05, 0.1, 1870300
pitch = layout.dqn_pitch(c_in, n_act)
n_local = P // world; row0 = rank * n_local
theta = (torch.rand(pitch, device=dev) - 0.5) * 0.05
if world > 1: dist.broadcast(theta, 0)
members = torch.empty((n_local, pitch), dtype=torch.float32, device=dev)
frames = ops.random_frames(7 + rank, (n_local, B, c_in, 84, 84), dev)
gen = [0]
parts = {}
def step():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    ops.es_perturb_dqn(theta, c_in, n_act, sigma, seed, "agent_0", gen[0], row0, n_local, out=members)
    ev[1].record()
    logits, _ = ops.deepqn_forward(members, frames, c_in, n_act)
    fit_local = logits.max(dim=2).values.mean(dim=1).to(torch.float64).contiguous()      # synthetic fitness
    ev[2].record()
    if world > 1:
        fit_all = torch.empty(P, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(fit_all, fit_local)
    delta = ops.es_update_members(fit_local, members, theta, 0, sigma, lr, P)
    ev[3].record()
    if world > 1: dist.all_reduce(delta)
    ops.axpy(1.0, delta, theta)
    ev[4].record()
    gen[0] += 1
    parts["ev"] = ev
ms = timed(step, 5)
torch.cuda.synchronize()
ev = parts["ev"]
names = ("perturb_ms", "forward_ms", "gather_update_ms", "allreduce_apply_ms")
split = {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}
if rank == 0:
    print(json.dumps({"config": "5: boxing_v2 Co-ES DeepQN step, synthetic frames", "population": P, "n_gpus": world,
                      "members_per_gpu": n_local, "frames_per_member": B, "ms_per_step": ms,
                      "members_per_s": P / ms * 1e3, "delta_bytes": pitch * 4, **split}), flush=True)
if world > 1: dist.destroy_process_group()

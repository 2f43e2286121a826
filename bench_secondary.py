"""Secondary measurements carried on bench.py's JSON line (``secondary``): the BASELINE configs the
headline metric is not quoted on, and the per-kernel rooflines of the population kernels.

  kernels   K3 / K5 / K6 / K7 (+ per-member weight statistics) at 1024 and 8192 rows: CUDA-event device
            time per launch, algorithmic bytes (4 B per parameter written or read) against the
            measured HBM figure
  config3   Co-GA with Hall-of-Fame, 8192 members per GPU (65,536 over 8 GPUs), hof 3: ms per generation
  config4   pong_v3 DeepQN policy forward (K2) on synthetic 84x84x4 frames at B = 1 and B = 4 frames per member
  config5   boxing_v2 Co-ES DeepQN step at population 4096 over the ranks: perturb / forward / update split

Everything here runs the product kernels through ``coevonet_b200.ops`` / ``engine``; nothing imports the oracle.
"""
from __future__ import annotations

import types

import torch
import torch.distributed as dist

ROLES = ("agent_0", "agent_1", "adversary_0")


def _timed(fn, n, warm, dev, world=1, collective=False):
    """Mean device ms per call (CUDA events on the current stream; max over ranks when collective)."""
    sync = collective and world > 1
    for _ in range(warm):
        fn()
    if sync:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    if sync:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    if sync:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def kernel_rooflines(dev, hbm_gbs, peak_src, sizes=(1024, 8192)):
    """K3 / K5 / K6 / K7 device time per launch and achieved GB/s of ALGORITHMIC bytes vs the HBM peak."""
    from coevonet_b200 import layout, ops
    in_dim, seed = 10, 1870300
    D, pitch = layout.fc_dim(in_dim), layout.fc_pitch(in_dim)
    theta = ops.fc_init(in_dim, seed, "agent_0", 0, 1, dev)[0].contiguous()
    out = []
    for P in sizes:
        rows = torch.empty((P, pitch), dtype=torch.float32, device=dev)
        elites = ops.fc_init(in_dim, seed, "agent_0", 1, 5, dev)
        fit = torch.linspace(-30, -5, P, dtype=torch.float64, device=dev)
        row_bytes = P * pitch * 4.0
        gen = [0]

        def k5():
            gen[0] += 1
            ops.es_perturb(theta, in_dim, 0.05, seed, "agent_0", gen[0], 0, P, out=rows)

        def k3():
            gen[0] += 1
            ops.ga_repopulate(elites, D, 0.05, seed, "agent_0", gen[0], 0, P, out=rows)

        cases = (
            ("K5 es_perturb_kernel", k5, row_bytes, "hbm-write", "4 B per parameter written (theta read from L2)"),
            ("K3 ga_repopulate_kernel", k3, row_bytes, "hbm-write", "4 B per parameter written (5 elite rows read from L2)"),
            ("K6 es_update_members_partial_kernel + finish",
             lambda: ops.es_update_members(fit, rows, theta, in_dim, 0.05, 0.1, P), row_bytes, "hbm-read",
             "4 B per parameter per member read"),
            ("K6 es_update_partial_kernel + finish (noise regenerated)",
             lambda: ops.es_update(fit, in_dim, 0.05, 0.1, P, seed, "agent_0", 1, 0), 8.0 * pitch, "alu",
             "8 B per parameter of theta / delta traffic; P normals per parameter of Philox + Box-Muller"),
            ("K7 diversity_dist_kernel", lambda: ops.diversity_dist(rows, theta, in_dim), row_bytes, "hbm-read",
             "4 B per parameter per member read"),
            ("weight_stats_kernel", lambda: ops.weight_stats(rows, in_dim), row_bytes, "hbm-read",
             "4 B per parameter per member read"),
        )
        for name, fn, nbytes, bound, unit in cases:
            ms = _timed(fn, 10, 3, dev)
            gbs = nbytes / (ms * 1e-3) / 1e9
            rec = {"kernel": name, "rows": P, "us_per_launch": ms * 1e3, "bound": bound,
                   "algorithmic_bytes_per_launch": nbytes, "bytes_def": unit}
            if bound == "alu":
                rec["normals_per_s"] = P * (D - 1536) / (ms * 1e-3)
            else:
                rec.update(achieved=gbs, peak=hbm_gbs, unit="GB/s", frac=gbs / hbm_gbs,
                           peak_source=f"MEASURED_PEAKS.json hbm_gbs ({peak_src})")
            out.append(rec)
        del rows
    return out


def config3_ga(dev, comm, P_per_gpu=8192, gens=3):
    """Co-GA with Hall-of-Fame at 8192 members per GPU (BASELINE configs[2]: 65,536 over 8 GPUs)."""
    from coevonet_b200 import engine, layout, ops
    P = P_per_gpu * comm.world
    args = types.SimpleNamespace(
        algorithm="GA", generations=gens, population=P, hof_size=3, game="simple_adversary_v3",
        mutation_power_agent_0=0.05, mutation_power_agent_1=0.05, mutation_power_adversary=0.05,
        learning_rate=0.1, max_timesteps_per_episode=400, max_evaluation_steps=400, elites_number=5,
        adaptive=True, max_mutation_power=0.5, min_mutation_power=0.001, fitness_sharing=True,
        early_stopping=False, patience=300, min_delta=0.1, debug=False, precision="float32",
        save=False, envs_per_member=1, reference_compat=True, init_states="device",
        seed=1870300, plots=False, record_history=False)
    shard = engine.Shard(P, comm.rank, comm.world)
    pop = {r: ops.fc_init(layout.OBS_DIM[r], 7, r, shard.row0, shard.n_local, dev) for r in ROLES}
    hof = {r: ops.fc_init(layout.OBS_DIM[r], 7, r, P, 3, dev) for r in ROLES}
    founder = {r: ops.fc_init(layout.OBS_DIM[r], 7, r, P - 1, 1, dev)[0] for r in ROLES}
    eng = engine.GAEngine(args, dev, pop, hof, founder, comm=comm)
    ms = _timed(lambda: eng.step(sync=False), gens, 1, dev, comm.world, collective=True)
    eng.check_status()
    games = 3 * P                                     # one game per member and role reaches the fitness
    return {"config": "BASELINE configs[2]: Co-GA + Hall-of-Fame, simple_adversary_v3", "population": P,
            "members_per_gpu": P_per_gpu, "n_gpus": comm.world, "hof_size": 3, "elites": 5,
            "ms_per_generation": ms, "generations_per_hour": 3.6e6 / ms,
            "world_steps_per_s": (games + engine.N_EVAL_GAMES) * 25 / (ms * 1e-3),
            "games_per_generation": games,
            "note": "reference_compat: of the hof_size games per member the reference overwrites all but the "
                    "last (genetic_algorithm.py:140), so one game per member and role is simulated; "
                    "k1 variant " + str(eng.variant)}


def config4_dqn_forward(dev, hbm_gbs, tf32_tflops, P=592):
    """pong_v3 DeepQN forward (K2), C = 4, A = 6, synthetic frames."""
    from coevonet_b200 import layout, ops
    c_in, n_act = 4, 6
    pitch, D = layout.dqn_pitch(c_in, n_act), layout.dqn_dim(c_in, n_act)
    members = (torch.rand((P, pitch), device=dev) - 0.5) * 0.05
    out = []
    for B in (1, 4):
        frames = ops.random_frames(1, (P, B, c_in, 84, 84), dev)
        ms = _timed(lambda: ops.deepqn_forward(members, frames, c_in, n_act), 8, 3, dev)
        nbytes = P * D * 4.0 + P * B * c_in * 7056.0
        flop = P * B * 18692096.0
        out.append({"config": "BASELINE configs[3]: pong_v3 DeepQN policy forward (K2), synthetic 84x84x4 frames",
                    "members": P, "frames_per_member": B, "ms": ms, "forwards_per_s": P * B / (ms * 1e-3),
                    "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_gbs,
                                 "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_gbs,
                                 "algorithmic_bytes": nbytes,
                                 "bytes_def": "member row (6.75 MB) + B frames of 28 KB, once per forward"},
                    "tensor": {"algorithmic_tflops": flop / (ms * 1e-3) / 1e12,
                               "issued_tf32_tflops": 3 * flop / (ms * 1e-3) / 1e12, "tf32_peak_tflops": tf32_tflops,
                               "note": "every contraction is 3xTF32 (three MMAs per product); HBM bound below "
                                       "B ~ 75 (SURVEY.md 8d)"}})
        del frames
    return out


def config5_dqn_es(dev, comm, P=4096, steps=5):
    """boxing_v2 Co-ES DeepQN step (C = 4, A = 18), population 4096 over the ranks."""
    from coevonet_b200 import layout, ops
    c_in, n_act, B, sigma, lr, seed = 4, 18, 1, 0.05, 0.1, 1870300
    world, rank = comm.world, comm.rank
    pitch = layout.dqn_pitch(c_in, n_act)
    n_local = P // world
    row0 = rank * n_local
    theta = (torch.rand(pitch, device=dev) - 0.5) * 0.05
    comm.broadcast0(theta)
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
        # the reference's Atari rollout is dead code: synthetic fitness
        fit_local = logits.max(dim=2).values.mean(dim=1).to(torch.float64).contiguous()
        ev[2].record()
        if world > 1:
            fit_all = torch.empty(P, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(fit_all, fit_local)
        delta = ops.es_update_members(fit_local, members, theta, 0, sigma, lr, P)
        ev[3].record()
        if world > 1:
            dist.all_reduce(delta)
        ops.axpy(1.0, delta, theta)
        ev[4].record()
        gen[0] += 1
        parts["ev"] = ev

    ms = _timed(step, steps, 2, dev, world, collective=True)
    ev = parts["ev"]
    names = ("perturb_ms", "forward_ms", "gather_update_ms", "allreduce_apply_ms")
    split = {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}
    row_gb = n_local * pitch * 4.0 / 1e9
    return {"config": "BASELINE configs[4]: boxing_v2 Co-ES DeepQN step, synthetic frames", "population": P,
            "n_gpus": world, "members_per_gpu": n_local, "frames_per_member": B, "ms_per_step": ms,
            "members_per_s": P / (ms * 1e-3), "delta_bytes": pitch * 4, **split,
            "perturb_GBps_written": row_gb / (split["perturb_ms"] * 1e-3),
            "update_GBps_read": row_gb / (split["gather_update_ms"] * 1e-3)}

"""N > 1 on real GPUs (NCCL): a generation sharded over two ranks against the same generation on one GPU
(evolutionary_strategy.py:236-265, genetic_algorithm.py:125-290).  Needs two GPUs: skipped on
the single-GPU box, run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROLES = ("agent_0", "agent_1", "adversary_0")


def _args(algorithm, P, E):
    return types.SimpleNamespace(
        algorithm=algorithm, generations=2, population=P, hof_size=2, game="simple_adversary_v3",
        mutation_power_agent_0=0.05, mutation_power_agent_1=0.05, mutation_power_adversary=0.05,
        learning_rate=0.1, max_timesteps_per_episode=400, max_evaluation_steps=400, elites_number=3,
        adaptive=True, max_mutation_power=0.5, min_mutation_power=0.001, fitness_sharing=True,
        early_stopping=False, patience=300, min_delta=0.1, debug=False, precision="float32", save=False,
        envs_per_member=E, reference_compat=True, init_states="device", seed=31337, plots=False,
        record_history=False)


def _worker(rank, world, algorithm, P, E, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from coevonet_b200 import engine, layout, ops
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    comm = engine.Comm()
    args = _args(algorithm, P, E)
    shard = engine.Shard(P, comm.rank, comm.world)
    if algorithm == "ES":
        torch.manual_seed(100 + rank)          # DIFFERENT generators per rank: the engine takes rank 0's rows
        theta = {r: ops.fc_init(layout.OBS_DIM[r], 5 + rank, r, 0, 1, dev)[0].cpu() for r in ROLES}
        if world == 1:
            theta = {r: ops.fc_init(layout.OBS_DIM[r], 5, r, 0, 1, dev)[0].cpu() for r in ROLES}
        eng = engine.ESEngine(args, dev, theta, comm=comm)
        eng.evaluate()
        eng.update()
        out = dict(rewards=np.stack([comm.all_gather_rows(eng.rewards[r], eng.shard).cpu().numpy() for r in ROLES]),
                   delta=eng.delta_cat.cpu().numpy(), variant=eng.variant)
    else:
        pop = {r: ops.fc_init(layout.OBS_DIM[r], 9, r, shard.row0, shard.n_local, dev) for r in ROLES}
        hof = {r: ops.fc_init(layout.OBS_DIM[r], 9, r, P, 2, dev) for r in ROLES}
        founder = {r: ops.fc_init(layout.OBS_DIM[r], 9, r, P - 1, 1, dev)[0] for r in ROLES}
        eng = engine.GAEngine(args, dev, pop, hof, founder, comm=comm)
        eng.step(sync=False)
        eng.step(sync=False)
        eng.check_status()
        pop_all = np.stack([comm.all_gather_rows(eng.pop[r][:, :2048].contiguous(), eng.shard).cpu().numpy()
                            for r in ROLES])
        out = dict(fitness=np.stack([eng.fitness[r].cpu().numpy() for r in ROLES]),
                   elite_ids=np.stack([eng.elite_ids[r].cpu().numpy() for r in ROLES]),
                   hof=np.stack([eng.hof[r].cpu().numpy()[:, :2048] for r in ROLES]), pop=pop_all,
                   gstate=eng.gstate.cpu().numpy(), variant=eng.variant)
    if rank == 0:
        q.put(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _launch(world, algorithm, P, E, port):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, algorithm, P, E, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_es_generation_sharded_over_two_gpus_equals_one_gpu():
    P, E = 384, 16                                 # 6,144 episodes per role: the lockstep (tcgen05) kernels
    one = _launch(1, "ES", P, E, 29701)
    two = _launch(2, "ES", P, E, 29702)
    assert one["variant"] == two["variant"] == 3
    assert np.array_equal(one["rewards"], two["rewards"])       # Philox by global member id, same K1 variant
    scale = np.abs(one["delta"]).max()
    assert np.abs(one["delta"] - two["delta"]).max() <= 1e-5 * scale    # all-reduce summation order only


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ga_generations_sharded_over_two_gpus_equal_one_gpu():
    P, E = 1200, 2
    one = _launch(1, "GA", P, E, 29703)
    two = _launch(2, "GA", P, E, 29704)
    assert one["variant"] == two["variant"]
    for k in ("fitness", "elite_ids", "hof", "pop", "gstate"):           # no floating-point reduction on this path
        assert np.array_equal(one[k], two[k]), k

"""World-size-2 gloo tests of the sharding / collective logic.

The engines run with the TEST-ONLY oracle backend (tests/oracle_backend.py) so
that the N > 1 host path -- row-block partition, fitness all-gather, elite
all-reduce (= HoF broadcast), ES delta all-reduce, replicated selection -- is
exercised on CPU.  A sharded run must reproduce the single-process run.
"""
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROLES = ("agent_0", "agent_1", "adversary_0")


def _args(algorithm, P, fitness_sharing=True):
    return types.SimpleNamespace(
        algorithm=algorithm, generations=2, population=P, hof_size=2, game="simple_adversary_v3",
        mutation_power_agent_0=0.05, mutation_power_agent_1=0.05, mutation_power_adversary=0.05,
        learning_rate=0.1, max_timesteps_per_episode=400, max_evaluation_steps=400, elites_number=2,
        adaptive=False, max_mutation_power=0.5, min_mutation_power=0.001, fitness_sharing=fitness_sharing,
        early_stopping=False, patience=300, min_delta=0.1, debug=False, precision="float32", save=False,
        envs_per_member=1, reference_compat=True, init_states="device", seed=99, plots=False,
        record_history=True)


def _founders(P):
    sys.path.insert(0, ROOT)
    from coevonet_b200 import layout
    from oracle import weights
    pop, hof, founder, theta = {}, {}, {}, {}
    for i, r in enumerate(ROLES):
        in_dim = layout.OBS_DIM[r]
        rows = np.zeros((P + 3, layout.fc_pitch(in_dim)), dtype=np.float32)
        rows[:, :layout.fc_dim(in_dim)] = weights.make_fc_rows(P + 3, in_dim, 700 + i, ln_jitter=0.02)
        pop[r] = torch.from_numpy(rows[:P])
        hof[r] = torch.from_numpy(rows[P:P + 2])
        founder[r] = torch.from_numpy(rows[P - 1].copy())
        theta[r] = torch.from_numpy(rows[P + 2].copy())
    return pop, hof, founder, theta


def _run(rank, world, algorithm, P, port, q, fitness_sharing=True):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_backend
    from coevonet_b200 import engine
    comm = engine.Comm()
    args = _args(algorithm, P, fitness_sharing)
    pop, hof, founder, theta = _founders(P)
    shard = engine.Shard(P, comm.rank, comm.world)
    sl = slice(shard.row0, shard.row0 + shard.n_local)
    if algorithm == "GA":
        eng = engine.GAEngine(args, "cpu", {r: pop[r][sl].clone() for r in ROLES}, hof, founder,
                              kernels=oracle_backend, comm=comm)
    else:
        eng = engine.ESEngine(args, "cpu", theta, kernels=oracle_backend, comm=comm)
    evs = [eng.step() for _ in range(2)]
    out = dict(evs=np.asarray(evs), world=world, rank=rank, row0=shard.row0, n_local=shard.n_local)
    if algorithm == "GA":
        out["fitness"] = np.stack([eng.history[g]["fitness"][r] for g in range(2) for r in ROLES])
        out["elite_ids"] = np.stack([eng.history[g]["elite_ids"][r] for g in range(2) for r in ROLES])
        out["hof"] = np.stack([eng.hof[r].numpy()[:, :1000] for r in ROLES])
        out["pop_head"] = np.stack([eng.pop[r].numpy()[:, :1000] for r in ROLES])
    else:
        out["rewards"] = np.stack([eng.history[g]["rewards"][r] for g in range(2) for r in ROLES])
        out["theta"] = np.stack([eng.theta[r].numpy()[:5000] for r in ROLES])
        out["fitness"] = np.stack([eng.fitness[r].numpy() for r in ROLES])
    q.put(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _launch(world, algorithm, P, port, fitness_sharing=True):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_run, args=(r, world, algorithm, P, port, q, fitness_sharing)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(outs, key=lambda o: o["rank"])


def test_shard_partition():
    sys.path.insert(0, ROOT)
    from coevonet_b200.engine import Shard
    for P, W in ((7, 2), (1024, 8), (65536, 8), (5, 8)):
        shards = [Shard(P, r, W) for r in range(W)]
        assert sum(s.n_local for s in shards) == P
        assert shards[0].row0 == 0
        for a, b in zip(shards, shards[1:]):
            assert b.row0 == a.row0 + a.n_local
        for row in (0, P // 2, P - 1):
            r = shards[0].owner(row)
            assert shards[r].row0 <= row < shards[r].row0 + shards[r].n_local


@pytest.mark.timeout(900)
def test_ga_two_ranks_match_one_rank():
    P = 7                                     # uneven shards: 4 + 3
    single = _launch(1, "GA", P, 29611)[0]
    two = _launch(2, "GA", P, 29612)
    for o in two:
        np.testing.assert_allclose(o["fitness"], single["fitness"], rtol=1e-12, atol=0)   # all-gather
        assert np.array_equal(o["elite_ids"], single["elite_ids"])                       # replicated K4
        assert np.array_equal(o["hof"], single["hof"])                                   # elite all-reduce
        np.testing.assert_allclose(o["evs"], single["evs"], rtol=1e-12)
        sl = slice(o["row0"], o["row0"] + o["n_local"])
        assert np.array_equal(o["pop_head"], single["pop_head"][:, sl])                  # global member ids


@pytest.mark.timeout(900)
def test_es_two_ranks_match_one_rank():
    P = 6
    single = _launch(1, "ES", P, 29613)[0]
    two = _launch(2, "ES", P, 29614)
    for o in two:
        np.testing.assert_allclose(o["rewards"], single["rewards"], rtol=1e-12, atol=0)
        # delta all-reduce: partial sums differ from the single-rank sum only in fp32 rounding.  The engine
        # reads sigma*z_i back as members[i] - theta, so a 1-ulp difference of theta after generation 1
        # re-rounds the noise of generation 2 (|theta| has grown to ~5 here: ulp/2 = 2.4e-7 per term, times
        # |fitness| ~ 40 and lr/(n sigma) = 0.33 over 6 members): 1e-5 absolute on a delta of magnitude ~5
        np.testing.assert_allclose(o["theta"], single["theta"], rtol=0, atol=3e-5)
    assert np.array_equal(two[0]["theta"], two[1]["theta"])                              # replicas agree


@pytest.mark.timeout(900)
def test_es_two_ranks_without_fitness_sharing_gather_under_the_update():
    """Without fitness sharing the update needs the local fitness only: the all-gather is started before and finished
    after the three K6 calls (Comm.all_gather_rows_start).  Uneven shards (3 + 2); the gathered fitness is the
    record every rank keeps."""
    P = 5
    single = _launch(1, "ES", P, 29621, fitness_sharing=False)[0]
    two = _launch(2, "ES", P, 29622, fitness_sharing=False)
    for o in two:
        np.testing.assert_allclose(o["rewards"], single["rewards"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(o["fitness"], single["fitness"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(o["theta"], single["theta"], rtol=0, atol=3e-5)
    assert np.array_equal(two[0]["theta"], two[1]["theta"])
    assert np.array_equal(two[0]["fitness"], two[1]["fitness"])


def _run_unseeded(rank, world, port, q):
    """Two ranks whose torch generators DIFFER (what an unseeded torchrun launch gives): the base agents
    must still come out identical -- rank 0's -- on both (evolutionary_strategy.py:222-233)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_backend
    from coevonet_b200 import engine, layout
    from coevonet_b200.MPE.fcnetwork import FCNetwork
    comm = engine.Comm()
    torch.manual_seed(1000 + 17 * rank)
    expect_first = FCNetwork(10, 5, "float32").flat_row() if rank == 0 else None
    torch.manual_seed(1000 + 17 * rank)
    # (a) the drivers' path: generator state of rank 0 on every rank, then the reference's constructors
    engine.sync_torch_rng(comm)
    synced = {r: FCNetwork(layout.OBS_DIM[r], 5, "float32").flat_row() for r in ROLES}
    # (b) the engine's own guard: rows that differ per rank are replaced by rank 0's
    torch.manual_seed(5 + rank)
    different = {r: FCNetwork(layout.OBS_DIM[r], 5, "float32").flat_row() for r in ROLES}
    args = _args("ES", 4)
    eng = engine.ESEngine(args, "cpu", different, kernels=oracle_backend, comm=comm)
    ev = eng.step()
    q.put(dict(rank=rank, synced=np.stack([synced[r].numpy()[:3000] for r in ROLES]),
               theta=np.stack([eng.theta[r].numpy()[:3000] for r in ROLES]), ev=np.asarray(ev),
               first=None if expect_first is None else expect_first.numpy()[:3000]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_unseeded_ranks_start_from_rank0_state():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_run_unseeded, args=(r, 2, 29617, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(2)], key=lambda o: o["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(outs[0]["synced"], outs[1]["synced"])          # same founders on both ranks
    assert np.array_equal(outs[0]["synced"][0], outs[0]["first"])         # and they are what rank 0 alone draws
    assert np.array_equal(outs[0]["theta"], outs[1]["theta"])            # engine guard: replicas agree after a step
    assert np.array_equal(outs[0]["ev"], outs[1]["ev"])


def test_empty_shard_contributes_nothing():
    """population < world size: ranks with no rows must run -- host-side shapes."""
    sys.path.insert(0, ROOT)
    from coevonet_b200.engine import Shard
    s = [Shard(3, r, 8) for r in range(8)]
    assert [x.n_local for x in s] == [1, 1, 1, 0, 0, 0, 0, 0]
    assert all(x.row0 == 3 for x in s[3:])

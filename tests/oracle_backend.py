"""Checker backend for the engines: the call signatures of ``coevonet_b200.ops``
implemented with the CPU oracle on torch CPU tensors.

TEST-ONLY.  It lets the world-size-2 gloo tests exercise the sharding /
collective logic of ``coevonet_b200.engine`` in a container without a GPU.  The
product never constructs this object (``engine.default_kernels`` is
``coevonet_b200.ops``, which refuses CPU tensors).
"""
import numpy as np
import torch

from coevonet_b200 import layout, ops as _ops
from oracle import ga_es, layout as olayout, philox, rollout as orollout

cycles_for_limit = _ops.cycles_for_limit        # host arithmetic, device agnostic
reward_slots = _ops.reward_slots
diversity_from_dist = _ops.diversity_from_dist
MAX_CYCLES = 25


def raise_on_status(status):
    if int(status.item()) != 0:
        raise ValueError("\n\t Warning: output contains inf or NaN")


def generation_state(sigmas, hist_capacity, device):
    return torch.from_numpy(ga_es.generation_state([float(x) for x in sigmas], int(hist_capacity)))


def generation_end(eval_out, gstate, hist_capacity, **kw):
    ga_es.generation_end(eval_out.numpy(), gstate.numpy(), int(hist_capacity), **kw)   # in place (shared memory)
    return gstate


def weight_stats(rows, in_dim, *, out=None):
    return torch.from_numpy(ga_es.weight_stats(rows.numpy(), olayout.fc_perturbable_index(in_dim)).astype(np.float32))


def mpe_rollout(member_role, members, opp_a, opp_b, init, *, n_cycles=MAX_CYCLES, pos_first=True,
                init_shared=False, variant=0, out=None, status=None):
    seat = layout.SEAT_OF[member_role] if isinstance(member_role, str) else int(member_role)
    others = [s for s in range(3) if s != seat]
    seats = layout.SEATS
    nets = {seats[seat]: members.numpy(), seats[others[0]]: opp_a.numpy(), seats[others[1]]: opp_b.numpy()}
    P, K = members.shape[0], opp_a.shape[0]
    E = init.shape[1] if init_shared else init.shape[2]
    idx = np.zeros((P * K * E, 3), dtype=np.int64)
    flat = np.zeros((P * K * E, 11))
    e = 0
    ini = init.numpy()
    for m in range(P):
        for k in range(K):
            for i in range(E):
                for s in range(3):
                    idx[e, s] = m if s == seat else k
                flat[e] = ini[k, i] if init_shared else ini[m, k, i]
                e += 1
    res = orollout.rollout(nets, idx, flat, n_cycles=n_cycles, pos_first=pos_first)
    o = np.stack([res["sum_good"], res["last_good"], res["sum_adv"], res["min_gap"].astype(np.float64)], axis=1)
    t = torch.from_numpy(o.reshape(P, K, E, 4))
    if out is not None:
        out.copy_(t)
        return out
    return t


def diversity_dist(pop, ref, in_dim, *, out=None):
    pidx = olayout.fc_perturbable_index(in_dim)
    d = np.array([np.linalg.norm(r[pidx] - ref.numpy()[pidx]) for r in pop.numpy()], dtype=np.float32)
    return torch.from_numpy(d)


def select_topk(fitness, k, order=0):
    return torch.from_numpy(ga_es.select_topk(fitness.numpy(), k, order))


def gather_rows(src, idx, *, row0=0, n_local=None, out=None):
    pitch = src.shape[-1]
    res = torch.zeros((idx.shape[0], pitch), dtype=torch.float32) if out is None else out
    for i, g in enumerate(idx.tolist()):
        s = g - row0
        if n_local is None or 0 <= s < n_local:
            res[i] = src[s]
        else:
            res[i] = 0
    return res


def ga_repopulate(elites, dim, sigma, seed, role, gen, row0, n_rows, *, out=None, noise_out=None,
                  crossover_rate=0.0):
    role_id = layout.ROLE_ID[role] if isinstance(role, str) else int(role)
    sigma = float(sigma)
    E, pitch = elites.shape
    res = torch.zeros((n_rows, pitch), dtype=torch.float32) if out is None else out
    el = elites.numpy()
    z = philox.normals(seed, philox.KIND_GA, role_id, gen, np.arange(row0, row0 + n_rows), dim)
    if crossover_rate > 0:
        rows = ga_es.ga_repopulate_crossover(el[:, :dim], sigma, z, crossover_rate, seed, role_id, gen,
                                             np.arange(row0, row0 + n_rows))
        res[:, :dim] = torch.from_numpy(rows)
        res[:, dim:] = 0
        return res
    for r in range(n_rows):
        c = row0 + r
        if c == 0:
            row = el[0, :dim].copy()
        else:
            row = el[(c - 1) % E, :dim] + (np.float32(sigma) * z[r]).astype(np.float32)
        res[r, :dim] = torch.from_numpy(row.astype(np.float32))
        res[r, dim:] = 0
    return res


def es_perturb(theta, in_dim, sigma, seed, role, gen, row0, n_rows, *, out=None, noise_out=None):
    role_id = layout.ROLE_ID[role] if isinstance(role, str) else int(role)
    sigma = float(sigma)
    D = layout.fc_dim(in_dim)
    pidx = olayout.fc_perturbable_index(in_dim)
    z = philox.normals(seed, philox.KIND_ES, role_id, gen, np.arange(row0, row0 + n_rows), D)
    rows, _ = ga_es.es_perturb(theta.numpy()[:D], sigma, z, pidx)
    res = torch.zeros((n_rows, layout.fc_pitch(in_dim)), dtype=torch.float32) if out is None else out
    res[:, :D] = torch.from_numpy(rows)
    res[:, D:] = 0
    return res


def es_update(fitness, in_dim, sigma, lr, n_total, seed, role, gen, row0, *, out=None):
    role_id = layout.ROLE_ID[role] if isinstance(role, str) else int(role)
    sigma = float(sigma)
    D = layout.fc_dim(in_dim)
    pidx = olayout.fc_perturbable_index(in_dim)
    n = fitness.shape[0]
    z = philox.normals(seed, philox.KIND_ES, role_id, gen, np.arange(row0, row0 + n), D)
    noise = (np.float32(sigma) * z[:, pidx]).astype(np.float32)
    f = fitness.numpy().astype(np.float32)
    coef = np.float32(lr / (n_total * sigma))
    delta = np.zeros(layout.fc_pitch(in_dim), dtype=np.float32)
    delta[pidx] = coef * (noise.T @ f)
    res = torch.from_numpy(delta)
    if out is not None:
        out.copy_(res)
        return out
    return res


def es_update_members(fitness, members, theta, in_dim, sigma, lr, n_total, *, out=None):
    """K6 from the materialised members: sigma*z_i taken as members[i] - theta (the engine's default)."""
    sigma = float(sigma)
    pitch = members.shape[1]
    noise = (members.numpy() - theta.numpy()[:pitch]).astype(np.float32)
    f = fitness.numpy().astype(np.float32)
    coef = np.float32(lr / (n_total * sigma))
    res = torch.from_numpy((coef * (noise.T @ f)).astype(np.float32))
    if out is not None:
        out.copy_(res)
        return out
    return res


def axpy(a, x, y):
    y.add_(x, alpha=a)
    return y


def init_states(seed, stream_id, n, device, rec0=0):
    return torch.from_numpy(philox.init_states(seed, stream_id, n, rec0=rec0))

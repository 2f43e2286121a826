"""K3-K7 + RNG parity: CUDA population kernels (through the C ABI) vs the oracle.

Integer work (Philox words, selection indices, gathers) must be bit-exact;
fp32 arithmetic given identical noise is bit-exact where the op order is fixed
(mutation, perturbation) and within a stated tolerance where a reduction order
differs (ES update, distances).
"""
import numpy as np
import pytest
import torch

from oracle import ga_es, layout as olayout, philox, weights

pytestmark = pytest.mark.gpu

SEED = 1870300


def _pad(rows, in_dim):
    from coevonet_b200 import layout
    out = np.zeros((rows.shape[0], layout.fc_pitch(in_dim)), dtype=np.float32)
    out[:, :rows.shape[1]] = rows
    return torch.from_numpy(out).cuda()


def test_philox_words_bit_exact():
    from coevonet_b200 import ops
    got = ops.philox_words(SEED, philox.KIND_GA, 2, 7, 5, 3, 100, "cuda").cpu().numpy().view(np.uint32)
    want = philox.words(SEED, philox.KIND_GA, 2, 7, np.arange(5, 8), 100)
    assert np.array_equal(got, want)
    got = ops.philox_words(2 ** 40 + 17, philox.KIND_ES, 0, 0, 0, 1, 4, "cuda").cpu().numpy().view(np.uint32)
    assert np.array_equal(got, philox.words(2 ** 40 + 17, philox.KIND_ES, 0, 0, [0], 4))


def test_init_states_bit_exact():
    from coevonet_b200 import ops
    got = ops.init_states(SEED, 3, 1000, "cuda").cpu().numpy()
    want = philox.init_states(SEED, 3, 1000)
    assert np.array_equal(got, want)
    assert set(np.unique(got[:, 0])) == {0.0, 1.0}
    assert np.all(np.abs(got[:, 1:]) < 1.0)
    part = ops.init_states(SEED, 3, 100, "cuda", rec0=400).cpu().numpy()
    assert np.array_equal(part, want[400:500])


@pytest.mark.parametrize("role,in_dim", [("agent_0", 10), ("adversary_0", 8)])
def test_ga_repopulate(role, in_dim):
    from coevonet_b200 import layout, ops
    D = layout.fc_dim(in_dim)
    E, P, gen, sigma = 3, 11, 4, 0.05
    elites = weights.make_fc_rows(E, in_dim, 41, ln_jitter=0.02)
    dev_el = _pad(elites, in_dim)
    noise = torch.zeros((P, layout.fc_pitch(in_dim)), dtype=torch.float32, device="cuda")
    out = ops.ga_repopulate(dev_el, D, sigma, SEED, role, gen, 0, P, noise_out=noise)
    z_dev = noise.cpu().numpy()[:, :D]
    z = philox.normals(SEED, philox.KIND_GA, philox.ROLE_ID[role], gen, np.arange(P), D)
    np.testing.assert_allclose(z_dev[1:], z[1:], rtol=0, atol=3e-6)      # libm vs CUDA ulps
    assert abs(z_dev[1:].mean()) < 5e-3 and abs(z_dev[1:].std() - 1) < 5e-3
    # arithmetic is bit-exact given the device noise: pop = elites stacked as a population
    want = ga_es.ga_repopulate(np.concatenate([elites, np.zeros((P - E, D), np.float32)]),
                               np.arange(E), sigma, z_dev)
    got = out.cpu().numpy()
    assert np.array_equal(got[:, :D], want)
    assert np.all(got[:, D:] == 0)
    # sharded generation (rows 4..10 on "another rank") equals the slice
    part = ops.ga_repopulate(dev_el, D, sigma, SEED, role, gen, 4, P - 4)
    assert torch.equal(part, out[4:])


@pytest.mark.parametrize("role,in_dim", [("agent_1", 10), ("adversary_0", 8)])
def test_es_perturb_and_update(role, in_dim):
    from coevonet_b200 import layout, ops
    D = layout.fc_dim(in_dim)
    pidx = olayout.fc_perturbable_index(in_dim)
    P, gen, sigma, lr = 96, 2, 0.05, 0.1
    theta = weights.make_fc_rows(1, in_dim, 43, ln_jitter=0.02)
    dev_theta = _pad(theta, in_dim)[0]
    noise = torch.zeros((P, layout.fc_pitch(in_dim)), dtype=torch.float32, device="cuda")
    rows = ops.es_perturb(dev_theta, in_dim, sigma, SEED, role, gen, 0, P, noise_out=noise)
    z_dev = noise.cpu().numpy()[:, :D]
    z = philox.normals(SEED, philox.KIND_ES, philox.ROLE_ID[role], gen, np.arange(P), D)
    np.testing.assert_allclose(z_dev[:, pidx], z[:, pidx], rtol=0, atol=3e-6)
    mask = np.ones(D, bool)
    mask[pidx] = False
    assert np.all(z_dev[:, mask] == 0)                               # LayerNorm untouched
    want_rows, want_noise = ga_es.es_perturb(theta[0], sigma, z_dev, pidx)
    assert np.array_equal(rows.cpu().numpy()[:, :D], want_rows)
    # K6: regenerated-noise update vs the reference formula on the exported noise
    fitness = np.linspace(-3, 2, P)
    delta = ops.es_update(torch.from_numpy(fitness).cuda(), in_dim, sigma, lr, P, SEED, role, gen, 0)
    want = ga_es.es_update(want_noise.astype(np.float64), fitness, lr, sigma)
    got = delta.cpu().numpy()
    scale = np.abs(want).max()
    np.testing.assert_allclose(got[pidx], want, rtol=0, atol=2e-5 * scale)
    assert np.all(got[:D][mask] == 0)
    # sharded partial sums add up (what the all-reduce does)
    d0 = ops.es_update(torch.from_numpy(fitness[:40]).cuda(), in_dim, sigma, lr, P, SEED, role, gen, 0)
    d1 = ops.es_update(torch.from_numpy(fitness[40:]).cuda(), in_dim, sigma, lr, P, SEED, role, gen, 40)
    np.testing.assert_allclose((d0 + d1).cpu().numpy()[pidx], want, rtol=0, atol=2e-5 * scale)
    # K7 distances: ||theta_i - theta|| = sigma * ||z_i||
    dist = ops.diversity_dist(rows, dev_theta, in_dim).cpu().numpy()
    div_want, d_want = ga_es.diversity_penalty(theta[0][pidx], want_rows[:, pidx])
    np.testing.assert_allclose(dist, d_want, rtol=2e-6)
    div = float(ops.diversity_from_dist(torch.from_numpy(dist).cuda()))
    assert abs(div - div_want) <= 1e-4 * max(1.0, abs(div_want)) + 1e-5


def test_es_update_matches_reference_golden(golden):
    """compute_weight_update outputs produced by the reference itself."""
    from coevonet_b200 import layout, ops
    g = golden("mutation")
    P, sigma, lr, gen = int(g["es_P"]), float(g["es_sigma"]), float(g["lr"]), int(g["es_gen"])
    pidx = olayout.fc_perturbable_index(10)
    delta = ops.es_update(torch.from_numpy(g["es_rewards"]).cuda(), 10, sigma, lr, P, int(g["seed"]),
                          "agent_0", gen, 0).cpu().numpy()
    want = g["es_update"]
    np.testing.assert_allclose(delta[pidx], want, rtol=0, atol=2e-5 * np.abs(want).max())
    # GA child of the golden parent: device Philox + device arithmetic vs reference Agent.mutate
    parent = weights.make_fc_rows(1, 10, int(g["parent_seed"]), ln_jitter=float(g["ln_jitter"]))
    D = layout.fc_dim(10)
    member = int(g["ga_member"])
    child = ops.ga_repopulate(_pad(parent, 10), D, float(g["ga_sigma"]), int(g["seed"]), "agent_0",
                              int(g["ga_gen"]), member, 1).cpu().numpy()[0, :D]
    np.testing.assert_allclose(child[:64], g["ga_child_head"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(child[-64:], g["ga_child_tail"], rtol=0, atol=1e-6)


def test_select_topk_bit_exact(golden):
    from coevonet_b200 import ops
    g = golden("selection")
    f = torch.from_numpy(g["fitness"]).cuda()
    assert np.array_equal(ops.select_topk(f, 7).cpu().numpy(), g["top7"])
    ft = torch.from_numpy(g["fitness_ties"]).cuda()
    assert np.array_equal(ops.select_topk(ft, 7).cpu().numpy(), g["top7_ties"])
    rng = np.random.Generator(np.random.PCG64(9))
    for P, k in ((1, 1), (5, 5), (20, 5), (1025, 33), (65536, 5), (70001, 64)):
        v = np.round(rng.standard_normal(P), 2)          # plenty of ties
        got = ops.select_topk(torch.from_numpy(v).cuda(), k).cpu().numpy()
        assert np.array_equal(got, ga_es.select_topk(v, k)), (P, k)


@pytest.mark.parametrize("role,in_dim", [("agent_0", 10), ("adversary_0", 8)])
def test_fc_init_bit_exact(role, in_dim):
    from coevonet_b200 import layout, ops
    D = layout.fc_dim(in_dim)
    got = ops.fc_init(in_dim, SEED, role, 5, 4, "cuda").cpu().numpy()
    want = philox.fc_init_rows(SEED, philox.ROLE_ID[role], np.arange(5, 9), in_dim)
    assert np.array_equal(got[:, :D], want)
    assert np.all(got[:, D:] == 0)
    segs, _ = olayout.fc_segments(in_dim)
    off = {n: o for n, o, _, _ in segs}
    w2 = got[:, off["fc2.weight"]:off["fc2.bias"]]
    assert abs(w2.std() - (1 / np.sqrt(512)) / np.sqrt(3)) < 2e-4 and np.abs(w2).max() <= 1 / np.sqrt(512)


def test_gather_rows_and_axpy():
    from coevonet_b200 import ops
    src = torch.randn((9, 64), device="cuda")
    idx = torch.tensor([8, 0, 3, 3], dtype=torch.int64, device="cuda")
    assert torch.equal(ops.gather_rows(src, idx), src[idx])
    x = torch.randn(1000, device="cuda")
    y = torch.randn(1000, device="cuda")
    want = (y.cpu().numpy() + (np.float32(0.5) * x.cpu().numpy()).astype(np.float32)).astype(np.float32)
    ops.axpy(0.5, x, y)
    assert np.array_equal(y.cpu().numpy(), want)


def test_fp32_peak_probe_runs():
    from coevonet_b200 import ops
    t0 = ops.fp32_peak("cuda:0", 0)
    t1 = ops.fp32_peak("cuda:0", 1)
    assert 5 < t0 < 200 and 5 < t1 < 200
    print(f"fp32 peak: scalar FFMA {t0:.1f} TFLOP/s, FFMA2 {t1:.1f} TFLOP/s")


def test_es_update_from_members_matches_regenerated_update():
    """K6 read back from the materialised members (members - theta = sigma*z up to one rounding) vs K6
    regenerated from the Philox key: same delta to 1e-6 of its scale, exact zeros on LayerNorm rows."""
    from coevonet_b200 import layout, ops
    from oracle import weights
    P, in_dim, sigma, lr, seed = 300, 10, 0.05, 0.1, 77
    pitch = layout.fc_pitch(in_dim)
    theta = torch.zeros(pitch, device="cuda")
    theta[:layout.fc_dim(in_dim)] = torch.from_numpy(weights.make_fc_rows(1, in_dim, 5)[0]).cuda()
    members = ops.es_perturb(theta, in_dim, sigma, seed, "agent_1", 3, 0, P)
    fit = torch.randn(P, dtype=torch.float64, device="cuda")
    want = ops.es_update(fit, in_dim, sigma, lr, P, seed, "agent_1", 3, 0)
    got = ops.es_update_members(fit, members, theta, in_dim, sigma, lr, P)
    scale = float(want.abs().max())
    assert scale > 0
    assert float((got - want).abs().max()) <= 1e-6 * scale + 1e-9
    assert torch.equal(got == 0, want == 0)            # LayerNorm entries and padding stay exactly zero

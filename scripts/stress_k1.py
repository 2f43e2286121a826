"""Large-P stress of cev_mpe_rollout_f32 for a given library build (raw ctypes)."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coevonet_b200 import layout
from coevonet_b200._lib import RolloutCfg
libname, P, K, E, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "coevonet_b200", "csrc", libname))
lib.cev_last_error.restype = ctypes.c_char_p
h = ctypes.c_void_p(); assert lib.cev_create(0, ctypes.byref(h)) == 0
def rows(n, in_dim, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = (torch.rand((n, layout.fc_pitch(in_dim)), device="cuda", generator=g) - 0.5) * 0.2
    segs, total = layout.fc_segments(in_dim)
    for name, off, shape, _ in segs:
        n_el = 1
        for s_ in shape: n_el *= s_
        if name.startswith("ln") and name.endswith("weight"): w[:, off:off + n_el] = 1.0
        if name.startswith("ln") and name.endswith("bias"): w[:, off:off + n_el] = 0.0
    w[:, total:] = 0
    return w
pop, adv, a1 = rows(P, 10, 1), rows(K, 8, 2), rows(K, 10, 3)
g = torch.Generator(device="cuda").manual_seed(9)
init = (torch.rand((P, K, E, 11), device="cuda", generator=g, dtype=torch.float64) * 2 - 1)
init[..., 0] = (init[..., 0] > 0).double()
out = torch.empty((P, K, E, 4), dtype=torch.float64, device="cuda")
cfg = RolloutCfg(25, 1, 2, 0)
vp = ctypes.c_void_p
for r in range(reps):
    t0 = time.time()
    rc = lib.cev_mpe_rollout_f32(h, 1, vp(pop.data_ptr()), P, ctypes.c_int64(pop.stride(0)), vp(adv.data_ptr()),
                                 ctypes.c_int64(adv.stride(0)), vp(a1.data_ptr()), ctypes.c_int64(a1.stride(0)), K,
                                 vp(init.data_ptr()), 0, E, ctypes.byref(cfg), vp(out.data_ptr()), None, None)
    try:
        torch.cuda.synchronize()
        print(libname, "rep", r, "rc", rc, "ok", round(time.time() - t0, 3), float(out[..., 0].mean()), flush=True)
    except Exception as e:
        print(libname, "rep", r, "FAIL", round(time.time() - t0, 3), str(e)[:60], flush=True); sys.exit(1)

"""Phase timing of the cluster rollout kernel (needs the -DCEV_PROFILE build)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coevonet_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libcoevonet_b200_prof.so")
import numpy as np, torch
from coevonet_b200 import layout, ops
lib = _lib.load()
lib.cev_debug_profile.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int]
theta = {"agent_0": ops.fc_init(10, 1, "agent_0", 0, 1, "cuda"), "agent_1": ops.fc_init(10, 2, "agent_1", 0, 1, "cuda"),
         "adversary_0": ops.fc_init(8, 3, "adversary_0", 0, 1, "cuda")}
names = ["tail(argmax/physics/obs)", "w1 wait", "layer1", "fc2 member", "fc2 streamed", "part reduce",
         "LN2 stats+dsmem", "cluster barrier 1", "normalise+logits+dsmem", "cluster barrier 2"]
FLAGS = [int(x) for x in sys.argv[1:]] or [0]
for flags, P, E in [(f, P, E) for f in FLAGS for (P, E) in ((264, 16), (264, 4))]:
    lib.cev_debug_set_flags(flags)
    members = ops.es_perturb(theta["agent_0"][0], 10, 0.05, 1, "agent_0", 0, 0, P)
    init = ops.init_states(1, 0, P * E, "cuda").reshape(P, 1, E, 11)
    ops.mpe_rollout("agent_0", members, theta["adversary_0"], theta["agent_1"], init, variant=2)
    torch.cuda.synchronize()
    buf = (ctypes.c_double * 16)()
    lib.cev_debug_profile(buf, 16, 1)
    ops.mpe_rollout("agent_0", members, theta["adversary_0"], theta["agent_1"], init, variant=2)
    torch.cuda.synchronize()
    lib.cev_debug_profile(buf, 16, 1)
    v = np.array(list(buf))[:10]
    ncyc = 8 * 25        # tiles per CTA x cycles
    print(f"flags={flags} P={P} E={E}: total {v.sum()/ncyc:.0f} clk per cycle")
    for n, x in zip(names, v):
        print(f"   {n:28s} {x/ncyc:9.0f} clk/cycle  {100*x/v.sum():5.1f}%")

import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
from coevonet_b200 import _lib
if os.environ.get("CEV_PROF"):
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libcoevonet_b200_prof.so")
import torch
from coevonet_b200 import layout, ops
def run(P, K, E, rows_from=0):
    pop = ops.fc_init(10, 1, "agent_0", 0, P, "cuda")
    adv = ops.fc_init(8, 1, "adversary_0", P, K, "cuda")
    a1 = ops.fc_init(10, 1, "agent_1", P, K, "cuda")
    init = ops.init_states(1, 0, P * K * E, "cuda").reshape(P, K, E, 11)
    torch.cuda.synchronize()
    t0 = time.time()
    try:
        out = ops.mpe_rollout("agent_0", pop[rows_from:], adv, a1, init[rows_from:].contiguous())
        torch.cuda.synchronize(); print("rollout ok", P, K, E, rows_from, round(time.time() - t0, 3), flush=True)
        if os.environ.get("CEV_PROF"):
            lib = _lib.load(); buf = (ctypes.c_double * 16)()
            lib.cev_debug_profile.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int]
            lib.cev_debug_profile(buf, 16, 1); print("prof tail:", [hex(int(x)) for x in list(buf)[13:]], flush=True)
    except Exception as e:
        print("FAIL", P, K, E, rows_from, round(time.time() - t0, 3), str(e)[:80], flush=True); sys.exit(1)
import itertools
case = sys.argv[1]
if case == "a": run(65536, 1, 1)
if case == "b": run(49152, 3, 1)
if case == "c": run(65536, 3, 1, rows_from=40000)
if case == "d": run(40000, 3, 4)

"""Where a Co-GA generation spends its time at BASELINE config-3 scale per GPU (8192 members, hof 3)."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coevonet_b200 import engine, layout, ops

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
init_mode = sys.argv[2] if len(sys.argv) > 2 else "device"
dev = torch.device("cuda", 0)
args = types.SimpleNamespace(
    algorithm="GA", generations=1, population=P, hof_size=3, game="simple_adversary_v3",
    mutation_power_agent_0=0.05, mutation_power_agent_1=0.05, mutation_power_adversary=0.05,
    learning_rate=0.1, max_timesteps_per_episode=400, max_evaluation_steps=400, elites_number=5,
    adaptive=True, max_mutation_power=0.5, min_mutation_power=0.001, fitness_sharing=True,
    early_stopping=False, patience=300, min_delta=0.1, debug=False, precision="float32",
    save=False, envs_per_member=1, reference_compat=True, init_states=init_mode,
    seed=1870300, plots=False, record_history=False)
pop = {r: ops.fc_init(layout.OBS_DIM[r], 7, r, 0, P, dev) for r in engine.ROLES}
hof = {r: ops.fc_init(layout.OBS_DIM[r], 8, r, P, 3, dev) for r in engine.ROLES}
founder = {r: pop[r][P - 1].clone() for r in engine.ROLES}
eng = engine.GAEngine(args, dev, pop, hof, founder)

def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"  {name:28s} {1e3 * (time.perf_counter() - t0):8.2f} ms", flush=True)
    return r

for g in range(3):
    print(f"generation {g} (P={P}, init={init_mode})")
    t0 = time.perf_counter()
    timed("evaluate (3 roles, 3 streams)", eng.evaluate)
    timed("select_and_repopulate (3 roles)", eng.select_and_repopulate)
    best = {r: eng.elites[r][0] for r in engine.ROLES}
    timed("eval games + generation_end", lambda: eng._finish_generation(best["agent_0"], best["agent_1"], best["adversary_0"]))
    eng.gen += 1
    print(f"  total {1e3 * (time.perf_counter() - t0):.1f} ms")

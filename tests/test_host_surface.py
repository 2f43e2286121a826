"""Host-side mirror of the reference surface that needs no GPU: CLI flags, the Args
bag, output-dir naming, checkpoint format (lists of MPEAgent / single agents pickled
with torch.save, utils/utils_pth_and_plots.py:8-96), the initial-state stream and the
reward-slot mapping."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from coevonet_b200 import main as cli
from coevonet_b200 import ops
from coevonet_b200.utils import mpe_spec
from coevonet_b200.utils.utils_pth_and_plots import create_output_dir, load_agent_for_testing, save_model


def _args(argv):
    return cli.Args(cli.parse_arguments(argv))


def test_reference_flags_and_train_script_quirk():
    # train_GA.sh passes --initial_mutation_power_agent_0 three times (Appendix C #14)
    a = _args(["--algorithm=GA", "--train", "--game=simple_adversary_v3", "--population=20", "--hof_size=3",
               "--elites_number=5", "--initial_mutation_power_agent_0=0.005",
               "--initial_mutation_power_agent_0=0.005", "--initial_mutation_power_agent_0=0.005",
               "--adaptive", "--max_mutation_power=0.7", "--min_mutation_power=0.0001", "--fitness_sharing",
               "--max_timesteps_per_episode=400", "--max_evaluation_steps=400"])
    assert a.mutation_power_agent_0 == 0.005 and a.mutation_power_agent_1 == 0.05
    assert a.reference_compat and a.envs_per_member == 1 and a.init_states == "reference"
    assert a.average_window == 50 and not a.device_init
    big = _args(["--population=65536", "--game=simple_adversary_v3"])
    assert big.device_init and big.init_states == "device"
    with pytest.raises(ValueError):
        _args(["--average_window=500", "--generations=100"])


def test_output_dir_naming(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    a = _args(["--algorithm=ES", "--game=simple_adversary_v3", "--generations=500", "--population=20",
               "--hof_size=1", "--max_timesteps_per_episode=400", "--fitness_sharing", "--adaptive",
               "--max_mutation_power=0.5", "--min_mutation_power=0.001", "--learning_rate=0.1"])
    d = create_output_dir(a)
    assert d == ("ES_models/gens500_pop20_hof1_gamesimple_adversary_v3_tslimit400_fitness-sharingTrue_"
                 "adaptiveTruemax_mutation0.5_min_mutation0.001_lr0.1")
    assert os.path.isdir(d)


def test_checkpoint_roundtrip_in_reference_format(tmp_path):
    from coevonet_b200.MPE.mpe_agent import MPEAgent
    args = types.SimpleNamespace(precision="float32", game="simple_adversary_v3")
    env = mpe_spec.DeviceMPEEnv()
    torch.manual_seed(5)
    hof = {r: [MPEAgent(env, args, r) for _ in range(2)] for r in ("agent_0", "agent_1", "adversary_0")}
    paths = {r: str(tmp_path / f"hall_of_fame_{r}.pth") for r in hof}
    for r in hof:
        save_model(hof[r], paths[r])
    targs = types.SimpleNamespace(algorithm="GA", GA_hof_to_test_agent_0=paths["agent_0"],
                                  GA_hof_to_test_agent_1=paths["agent_1"],
                                  GA_hof_to_test_adversary=paths["adversary_0"])
    a0, a1, adv = load_agent_for_testing(targs, env)
    assert torch.equal(a0.model.flat_row(), hof["agent_0"][-1].model.flat_row())     # newest HoF entry
    assert adv.model.input_channels == 8 and a1.model.input_channels == 10
    # ES: single agents
    save_model(hof["agent_0"][0], str(tmp_path / "agent_0.pth"))
    targs = types.SimpleNamespace(algorithm="ES", ES_model_to_test_agent_0=str(tmp_path / "agent_0.pth"),
                                  ES_model_to_test_agent_1=str(tmp_path / "agent_0.pth"),
                                  ES_model_to_test_adversary_0=paths["adversary_0"])
    b0, _, _ = load_agent_for_testing(targs, env)
    assert torch.equal(b0.model.flat_row(), hof["agent_0"][0].model.flat_row())
    with pytest.raises(ValueError):
        load_agent_for_testing(types.SimpleNamespace(algorithm="GA", GA_hof_to_test_agent_0=None,
                                                     GA_hof_to_test_agent_1=None, GA_hof_to_test_adversary=None))


def test_init_state_stream_matches_oracle_env():
    from oracle import mpe_env
    assert np.array_equal(mpe_spec.InitStateStream().draw(64), mpe_env.draw_initial_states(64))
    env = mpe_spec.DeviceMPEEnv()
    env.reset(seed=mpe_spec.ENV_SEED)
    first = env.take_pending()
    assert np.array_equal(first, mpe_env.draw_initial_states(1)[0])
    assert env.observation_space("adversary_0").shape == (8,) and env.action_space("agent_1").n == 5
    assert env.agents == ["adversary_0", "agent_0", "agent_1"]


def test_reward_slot_mapping_matches_oracle():
    from oracle import rollout as orollout
    rng = np.random.default_rng(0)
    out = torch.from_numpy(rng.standard_normal((7, 4)))
    res = dict(sum_good=out[:, 0].numpy(), last_good=out[:, 1].numpy(), sum_adv=out[:, 2].numpy())
    for limit in (None, 0, 1, 2, 3, 4, 5, 30, 73, 74, 75, 400):
        want = orollout.compat_slots(res, limit)
        got = ops.reward_slots(out, agent_step_limit=limit)
        for w, g in zip(want, got):
            np.testing.assert_allclose(g.numpy(), w, rtol=0, atol=1e-15)
        assert ops.cycles_for_limit(limit) == orollout.cycles_for_limit(limit)
    t0, t1, t2 = ops.reward_slots(out, reference_compat=False)
    assert torch.equal(t0, out[:, 0]) and torch.equal(t1, out[:, 0]) and torch.equal(t2, out[:, 2])

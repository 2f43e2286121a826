"""Flat-row layouts: product vs oracle vs a torch module built like the reference."""
import numpy as np
import pytest
import torch

from coevonet_b200 import layout
from oracle import layout as olayout


def test_dims_match_survey():
    assert layout.fc_dim(10) == 139781 and layout.fc_dim(8) == 138757
    assert len(layout.fc_perturbable_index(10)) == 138245
    assert len(layout.fc_perturbable_index(8)) == 137221
    assert layout.fc_pitch(10) % 32 == 0 and layout.fc_pitch(8) % 32 == 0
    assert layout.dqn_dim(4, 6) == 1687526 and layout.dqn_dim(6, 18) == 1697778
    for in_dim in (8, 10):
        a, ta = layout.fc_segments(in_dim)
        b, tb = olayout.fc_segments(in_dim)
        assert a == b and ta == tb
        assert np.array_equal(layout.fc_perturbable_index(in_dim), olayout.fc_perturbable_index(in_dim))
    assert layout.dqn_segments(4, 18) == olayout.dqn_segments(4, 18)


def test_sub_tensors_are_16_byte_aligned():
    for in_dim in (8, 10):
        segs, _ = layout.fc_segments(in_dim)
        offs = {n: o for n, o, _, _ in segs}
        assert offs["fc2.weight"] % 32 == 0          # cp.async / bulk-copy source
        assert offs["fc1.weight"] == 0


def test_pack_unpack_roundtrip_with_module():
    from coevonet_b200.MPE.fcnetwork import FCNetwork
    torch.manual_seed(0)
    net = FCNetwork(10, 5, "float32")
    row = layout.pack_state_dict(net.state_dict(), 10)
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    assert torch.equal(row[:layout.fc_dim(10)], flat)               # parameters() order
    sd = layout.unpack_to_state_dict(row, 10)
    for k, v in net.state_dict().items():
        assert torch.equal(sd[k], v)
    # ES view == get_perturbable_weights
    pert = net.get_perturbable_weights()
    assert np.array_equal(row.numpy()[layout.fc_perturbable_index(10)], pert)


@pytest.mark.reference
def test_layout_matches_reference_modules():
    from oracle import stubs
    ref = stubs.import_reference()
    torch.manual_seed(1)
    net = ref.MPE_fcnetwork.FCNetwork(8, 5, "float32")
    names = [n for n, _ in net.named_parameters()]
    assert names == [s[0] for s in layout.fc_segments(8)[0]]
    row = layout.pack_state_dict(net.state_dict(), 8).numpy()
    assert np.array_equal(row[layout.fc_perturbable_index(8)], net.get_perturbable_weights())
    dq = ref.Atari_deepqn.DeepQN(4, 6, "float32")
    assert [n for n, _ in dq.named_parameters()] == [s[0] for s in layout.dqn_segments(4, 6)[0]]

"""The C-ABI library loads and exports every symbol include/coevonet_b200.h
declares (no compute calls: this runs without a GPU)."""
import os
import re

import pytest

from coevonet_b200 import _lib, layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "coevonet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cev_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 20
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_geometry_helpers_agree_with_python_layout():
    lib = _lib.load()
    for in_dim in (8, 10):
        assert lib.cev_fc_dim(in_dim) == layout.fc_dim(in_dim)
        assert lib.cev_fc_pitch(in_dim) == layout.fc_pitch(in_dim)
    for c, a in ((4, 6), (4, 18), (6, 6), (6, 18)):
        assert lib.cev_dqn_dim(c, a) == layout.dqn_dim(c, a)
        assert lib.cev_dqn_pitch(c, a) == layout.dqn_pitch(c, a)
    assert lib.cev_version() >= 100


def test_ops_refuse_cpu_tensors():
    import torch
    from coevonet_b200 import ops
    with pytest.raises(_lib.CevError):
        ops.select_topk(torch.zeros(4, dtype=torch.float64), 2)

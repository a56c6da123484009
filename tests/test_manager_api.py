"""Host-side behaviour of the manager API, mirrored from the reference's contract (CPU, no kernels)."""

import pytest
import torch

from constraints_as_terminations_b200 import ConstraintManager, ConstraintsManager, ConstraintTerm, ConstraintTermCfg
from constraints_as_terminations_b200 import constraints, curriculums
from constraints_as_terminations_b200 import synthetic_env as se
from constraints_as_terminations_b200._isaaclab_compat import SceneEntityCfg


def _env(n=8):
    return se.SyntheticSolo12Env(n, device="cpu", pool=1)


def test_aliases():
    assert ConstraintsManager is ConstraintManager
    assert ConstraintTerm is ConstraintTermCfg


def test_term_parsing_and_order():
    mgr = ConstraintManager(se.solo12_constraints_cfg(), _env())
    assert mgr.active_terms == list(se.solo12_constraints_cfg().keys())
    assert len(mgr.active_terms) == 13
    # SceneEntityCfg resolution as Isaac Lab does it: all joints -> slice(None), subsets -> id lists
    assert mgr.get_term_cfg("joint_torque").params["asset_cfg"].joint_ids == slice(None)
    assert mgr.get_term_cfg("front_hfe_position").params["asset_cfg"].joint_ids == [1, 4]
    assert mgr.get_term_cfg("hip_position").params["asset_cfg"].joint_ids == [0, 3, 6, 9]
    assert mgr.get_term_cfg("contact").params["asset_cfg"].body_ids == [0, 2, 6, 10, 14]
    assert mgr.get_term_cfg("air_time").params["asset_cfg"].body_ids == [4, 8, 12, 16]
    text = str(mgr)
    assert "contains 13 active terms" in text and "joint_torque" in text and "Max p" in text


def test_none_terms_skipped_and_type_errors():
    cfg = se.solo12_constraints_cfg()
    cfg["contact"] = None
    assert "contact" not in ConstraintManager(cfg, _env()).active_terms
    with pytest.raises(TypeError, match="is not ConstraintTermCfg"):
        ConstraintManager({"bad": 3.0}, _env())
    bad = ConstraintTermCfg(func=constraints.upsidedown, max_p="high", params={"limit": 0.0, "asset_cfg": SceneEntityCfg("robot")})
    with pytest.raises(TypeError, match="must be float or int"):
        ConstraintManager({"bad": bad}, _env())


def test_get_set_term_cfg_and_curriculum():
    env = _env()
    mgr = ConstraintManager(se.solo12_constraints_cfg(), env)
    env.constraint_manager = mgr
    with pytest.raises(ValueError, match="not found"):
        mgr.get_term_cfg("nope")
    with pytest.raises(ValueError, match="not found"):
        mgr.set_term_cfg("nope", mgr.get_term_cfg("contact"))
    env.common_step_counter = 0
    assert curriculums.modify_constraint_p(env, None, "joint_torque", 24000, 0.25) == pytest.approx(0.05)
    assert mgr.get_term_cfg("joint_torque").max_p == pytest.approx(0.05)
    env.common_step_counter = 24000
    assert curriculums.modify_constraint_p(env, None, "joint_torque", 24000, 0.25) == pytest.approx(0.25)
    env.common_step_counter = 12000
    assert curriculums.modify_constraint_p(env, None, "joint_torque", 24000, 0.25) == pytest.approx(1 / 12)


def test_cfg_object_form_accepted():
    from constraints_as_terminations_b200._isaaclab_compat import configclass

    @configclass
    class ConstraintsCfg:
        upsidedown = ConstraintTerm(func=constraints.upsidedown, max_p=1.0, params={"limit": 0.0, "asset_cfg": SceneEntityCfg("robot")})
        nothing = None

    mgr = ConstraintManager(ConstraintsCfg(), _env())
    assert mgr.active_terms == ["upsidedown"]


def test_empty_manager_returns_empty_tensor():
    mgr = ConstraintManager({}, _env())
    out = mgr.compute()
    assert isinstance(out, torch.Tensor) and out.numel() == 0
    assert mgr.reset() == {}


def test_lazy_stats_dict_behaves_like_the_reference_dict():
    """`fused_reset_stats()` hands out a dict that splits the packed statistics vector on first access."""
    import torch

    from constraints_as_terminations_b200.constraint_manager import _LazyStats

    keys = ["Episode_Constraint_violation/a", "Episode_Constraint_probability/a"]
    d = _LazyStats(keys, torch.tensor([25.0, 0.5]))
    assert dict.__len__(d) == 0  # nothing materialised yet
    assert "Episode_Constraint_violation/a" in d and len(d) == 2 and list(d) == keys
    assert float(d[keys[0]]) == 25.0 and float(d.get(keys[1])) == 0.5 and d[keys[0]].ndim == 0
    merged = {}
    merged.update(_LazyStats(keys, torch.tensor([1.0, 2.0])))  # extras["log"].update(info) as in CaTEnv._reset_idx
    assert list(merged) == keys and float(merged[keys[1]]) == 2.0
    d[keys[0]] = d[keys[0]].unsqueeze(0)  # the trainer's logging loop rewrites entries in place (ppo.py:241-244)
    assert d[keys[0]].shape == (1,)

"""Drop-in proof against the reference's own config objects.

The `ConstraintsCfg` and `CurriculumCfg` class bodies are cut, verbatim, out of the reference's task file
(S12/cat_flat_env_cfg.py:259-355,384-) with `ast` -- the file itself cannot be imported (Isaac Lab / Isaac Sim are not
installed) -- and executed with the names they use bound to THIS package's objects (`ConstraintTerm`, `constraints`,
`curriculums`, `SceneEntityCfg`, `configclass`), exactly what swapping the import lines of the task file amounts to.

CPU: the unmodified cfg constructs, passes Isaac Lab's static signature check (the shim restates it), resolves its
regex joint / body selectors, equals the constants restated in synthetic_env.solo12_constraints_cfg, and the PPO
hyper-parameters of S12/agents/clean_rl_ppo_cfg.py equal solo12_flat_ppo_cfg.  GPU: the unmodified cfg objects drive this
repo's ConstraintManager + modify_constraint_p through the golden fixture of the real reference, bit for bit.
"""

import ast
import inspect
import os
import types

import pytest
import torch

from constraints_as_terminations_b200 import ConstraintManager, ConstraintTermCfg, constraints, curriculums, solo12_flat_ppo_cfg
from constraints_as_terminations_b200 import synthetic_env as se
from constraints_as_terminations_b200._isaaclab_compat import SceneEntityCfg, configclass
from oracle import ref_loader

S12 = "exts/cat_envs/cat_envs/tasks/locomotion/velocity/config/solo12"
needs_reference = pytest.mark.skipif(not ref_loader.reference_available(), reason="no reference tree (neither /root/reference nor oracle/_ref archive)")


def _cut(path, *names):
    """Source segments of the named top-level classes / assignments of a python file."""
    src = open(path).read()
    out = []
    for node in ast.parse(src).body:
        target = node.name if isinstance(node, ast.ClassDef) else (node.targets[0].id if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) else None)
        if target in names:
            seg = ast.get_source_segment(src, node)
            if isinstance(node, ast.ClassDef) and node.decorator_list:
                seg = "@configclass\n" + seg
            out.append(seg)
    assert len(out) == len(names), f"{names} not all found in {path}"
    return "\n\n".join(out)


def reference_cfg_objects():
    path = os.path.join(ref_loader.REFERENCE_ROOT, S12, "cat_flat_env_cfg.py")
    code = _cut(path, "ConstraintsCfg", "MAX_CURRICULUM_ITERATIONS", "CurriculumCfg")

    class CurrTerm:  # isaaclab.managers.CurriculumTermCfg: func + params is all the curriculum manager reads
        def __init__(self, func, params):
            self.func, self.params = func, params

    ns = {"configclass": configclass, "ConstraintTerm": ConstraintTermCfg, "constraints": constraints, "curriculums": curriculums,
          "SceneEntityCfg": SceneEntityCfg, "CurrTerm": CurrTerm}  # fmt: skip
    exec(compile(code, path, "exec"), ns)
    return ns["ConstraintsCfg"](), ns["CurriculumCfg"]()


@needs_reference
def test_reference_constraints_cfg_equals_the_restated_constants():
    cfg, cur = reference_cfg_objects()
    ours = se.solo12_constraints_cfg()
    got = {k: v for k, v in cfg.__dict__.items() if v is not None}
    assert list(got) == list(ours)
    for name, term in got.items():
        want = ours[name]
        assert isinstance(term, ConstraintTermCfg) and term.func is want.func and term.max_p == want.max_p, name
        assert set(term.params) == set(want.params), name
        for k, v in term.params.items():
            w = want.params[k]
            if isinstance(v, SceneEntityCfg):
                assert (v.name, v.joint_names, v.body_names) == (w.name, w.joint_names, w.body_names), (name, k)
            else:
                assert v == w, (name, k)
    cur_terms = {k: v for k, v in cur.__dict__.items() if not k.startswith("_")}
    assert tuple(cur_terms) == se.SOLO12_CURRICULUM_TERMS
    for name, t in cur_terms.items():
        assert t.func is curriculums.modify_constraint_p
        assert t.params == {"term_name": name, "num_steps": 24 * 1000, "init_max_p": 0.25}


@needs_reference
def test_reference_ppo_cfg_equals_the_restated_constants():
    path = os.path.join(ref_loader.REFERENCE_ROOT, S12, "agents/clean_rl_ppo_cfg.py")
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef))
    values = {}
    for node in cls.body:
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            values[node.targets[0].id] = ast.literal_eval(node.value)
        elif isinstance(node, ast.AnnAssign) and node.value is not None:
            values[node.target.id] = ast.literal_eval(node.value)
    ours = solo12_flat_ppo_cfg()
    assert len(values) >= 10
    for k, v in values.items():
        assert getattr(ours, k) == v, k


def test_builtin_terms_pass_the_isaaclab_signature_check():
    """Isaac Lab compares inspect.signature(func) with the keys of params; the wrapped built-in terms must
    show the reference's own signatures (env, limit, asset_cfg, ...), and a (env, **params) function must be refused."""
    from constraints_as_terminations_b200 import constraints as C

    ref_sigs = {
        "joint_position": ["env", "limit", "asset_cfg"], "joint_position_when_moving_forward": ["env", "limit", "velocity_deadzone", "asset_cfg"],
        "joint_torque": ["env", "limit", "asset_cfg"], "joint_velocity": ["env", "limit", "asset_cfg"], "joint_acceleration": ["env", "limit", "asset_cfg"],
        "upsidedown": ["env", "limit", "asset_cfg"], "contact": ["env", "asset_cfg"], "base_orientation": ["env", "limit", "asset_cfg"],
        "air_time": ["env", "limit", "velocity_deadzone", "asset_cfg"], "n_foot_contact": ["env", "number_of_desired_feet", "min_command_value", "asset_cfg"],
        "joint_range": ["env", "limit", "asset_cfg"], "action_rate": ["env", "limit", "asset_cfg"], "foot_contact_force": ["env", "limit", "asset_cfg"],
        "min_base_height": ["env", "limit", "asset_cfg"], "no_move": ["env", "velocity_deadzone", "joint_vel_limit", "asset_cfg"],
    }  # fmt: skip  (reference constraints.py:23-235)
    for name, params in ref_sigs.items():
        assert list(inspect.signature(getattr(C, name)).parameters) == params, name
    env = se.SyntheticSolo12Env(8, device="cpu", seed=0, pool=1)
    mgr = types.SimpleNamespace(_env=env)
    from constraints_as_terminations_b200._isaaclab_compat import ManagerBase

    for name, term in se.solo12_constraints_cfg(stress=True).items():
        ManagerBase._resolve_common_term_cfg(mgr, name, term, min_argc=1)  # raises on a mismatch

    def sloppy(env, **params):
        return torch.zeros(env.num_envs)

    with pytest.raises(ValueError, match="expects mandatory parameters"):
        ManagerBase._resolve_common_term_cfg(mgr, "sloppy", ConstraintTermCfg(func=sloppy, params={"limit": 1.0}, max_p=1.0), min_argc=1)


@needs_reference
@pytest.mark.gpu
def test_unmodified_reference_cfg_drives_the_cuda_manager_through_the_golden(golden_dir):
    from tests.helpers import replay_cat_golden

    gold = torch.load(os.path.join(golden_dir, "cat_solo12.pt"), weights_only=False)
    cfg, cur = reference_cfg_objects()
    dev = "cuda:0"
    env = se.SyntheticSolo12Env(gold["num_envs"], device=dev, seed=gold["seed"], pool=1, adversarial=True)
    mgr = ConstraintManager(cfg, env)  # a configclass instance, as CaTEnv.load_managers passes it (cat_env.py:38-40)
    env.constraint_manager = mgr
    assert mgr.active_terms == gold["names"]
    cur_terms = [t for k, t in cur.__dict__.items() if not k.startswith("_")]

    def step_fn(step, state, rec):
        reward, dones = mgr.compute_step(state["raw_reward"], rec["reset_buf"])
        return {"cstr_prob": mgr._cstr_prob_buf.clone(), "running_max": mgr.cat.get_running_maxes().squeeze(0).clone(),
                "reward": reward.clone(), "dones": dones.clone()}  # fmt: skip

    def set_max_p(recorded):
        # the curriculum of the task file, run the way Isaac Lab's CurriculumManager does: func(env, env_ids, **params)
        for t in cur_terms:
            t.func(env, None, **t.params)
        assert [float(mgr.get_term_cfg(n).max_p) for n in mgr.active_terms] == recorded

    replay_cat_golden(gold, env, step_fn, mgr.reset, set_max_p, exact=True, device=dev)

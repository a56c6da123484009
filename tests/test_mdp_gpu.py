"""Env-side MDP terms: the fused velocity-command update and random-push kernels against the CPU
restatement of the reference (oracle/mdp_oracle.py, U/mdp/commands.py:39-93, U/mdp/events.py:59-96).

The reference draws from torch's global generator, which nothing else can reproduce, so the random numbers are made
explicit: (a) the same uniforms are fed to both sides -> every output bit-exact (the heading yaw rate to 1e-6: fmod vs
torch.remainder); (b) the kernels draw from the device Philox stream and the oracle is fed the numbers the numpy Philox
restatement produces for the same (seed, offset) -> same bit-exact outputs, i.e. the device draws ARE that stream."""

import types

import numpy as np
import pytest
import torch

from constraints_as_terminations_b200 import mdp, ops
from constraints_as_terminations_b200.mdp.commands import make_command_cfg
from oracle import mdp_oracle, philox_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RANGES = {"lin_vel_x": (-0.3, 1.0), "lin_vel_y": (-0.7, 0.7), "ang_vel_z": (-0.78, 0.78), "heading": (-3.14, 3.14)}
PHYS_DT, EP_S = 0.005, 10.0  # cat_flat_env_cfg.py:479-484


def _state(n, seed):
    g = torch.Generator().manual_seed(seed)
    cmd = torch.stack([torch.empty(n).uniform_(*RANGES[k], generator=g) for k in ("lin_vel_x", "lin_vel_y", "ang_vel_z")], dim=1)
    cmd[::9] *= 0.05  # inside the dead zone
    cmd[1::13] = 0.0
    heading_target = torch.empty(n).uniform_(-3.14, 3.14, generator=g)
    heading_w = torch.empty(n).uniform_(-3.14, 3.14, generator=g)
    is_heading = torch.rand(n, generator=g) < 0.6
    is_standing = torch.rand(n, generator=g) < 0.1
    return cmd, heading_target, heading_w, is_heading, is_standing


def _kcfg(heading_command):
    return make_command_cfg(types.SimpleNamespace(**RANGES), 0.1, heading_command, 0.5, 0.7, 0.02, PHYS_DT, EP_S)


def _oracle(state, u, heading_command):
    return mdp_oracle.update_command(*state, u, ranges=RANGES, deadzone=0.1, heading_command=heading_command, stiffness=0.5,
                                     rel_heading=0.7, rel_standing=0.02, physics_dt=PHYS_DT, max_episode_length_s=EP_S)  # fmt: skip


def _check(got, want, heading_command):
    cmd, ht, ih, istand, res = got
    w_cmd, w_ht, w_ih, w_is, w_res = want
    assert torch.equal(res.cpu(), w_res) and torch.equal(istand.cpu(), w_is)
    assert torch.equal(cmd[:, :2].cpu(), w_cmd[:, :2])
    if heading_command:  # the yaw rate of heading envs went through wrap_to_pi: fmodf vs torch.remainder
        assert torch.equal(ih.cpu(), w_ih) and torch.equal(ht.cpu(), w_ht)
        torch.testing.assert_close(cmd[:, 2].cpu(), w_cmd[:, 2], rtol=0, atol=1e-6)
    else:
        assert torch.equal(cmd[:, 2].cpu(), w_cmd[:, 2])


@pytest.mark.parametrize("heading_command", [False, True])
@pytest.mark.parametrize("n", [1, 1000, 4096])
def test_command_update_with_shared_uniforms(n, heading_command):
    state = _state(n, seed=n)
    g = torch.Generator().manual_seed(n + 1)
    u = torch.rand(n, 8, generator=g)
    u[::5, 0] = 0.0  # force resamples (p >= 5e-4 always)
    u[2::7, 7] = 0.0  # force yaw flips
    dev = [t.to(DEV).contiguous() for t in state]
    res = mdp.update_velocity_command(_kcfg(heading_command), dev[0], dev[1], dev[2], dev[3], dev[4], uniforms=u.to(DEV))
    _check((dev[0], dev[1], dev[3], dev[4], res), _oracle(state, u, heading_command), heading_command)
    assert int(res.sum()) >= n // 5


def test_command_update_and_push_draw_the_documented_philox_stream():
    n, seed, off = 3000, 4242, 96
    state = _state(n, seed=3)
    dev = [t.to(DEV).contiguous() for t in state]
    rng = ops.make_rng_state(seed, DEV, offset=off)
    res = mdp.update_velocity_command(_kcfg(True), dev[0], dev[1], dev[2], dev[3], dev[4], rng_state=rng)
    u = torch.from_numpy(philox_oracle.uniform(seed, philox_oracle.STREAM_UNIFORM, off, 8 * n).reshape(n, 8).copy())
    _check((dev[0], dev[1], dev[3], dev[4], res), _oracle(state, u, True), True)
    assert rng.cpu().tolist() == [seed, off + 8 * n]
    # random push: p = dt / (2 T) = 2.5e-4 -> use a larger N so that some envs are pushed
    n = 200000
    g = torch.Generator().manual_seed(9)
    vel = torch.randn(n, 6, generator=g)
    vr = {"x": (-0.5, 0.5), "y": (-0.5, 0.5), "yaw": (-0.3, 0.3)}
    rng = ops.make_rng_state(seed, DEV, offset=7)
    pushed, new_vel = mdp.select_pushes(vel.to(DEV), PHYS_DT / (EP_S * 2), vr, rng_state=rng)
    u = torch.from_numpy(philox_oracle.uniform(seed, philox_oracle.STREAM_UNIFORM, 7, 8 * n).reshape(n, 8)[:, :7].copy())
    w_pushed, w_vel = mdp_oracle.select_pushes(vel, u, physics_dt=PHYS_DT, max_episode_length_s=EP_S, velocity_range=vr)
    assert torch.equal(pushed.cpu(), w_pushed) and torch.equal(new_vel.cpu(), w_vel)
    assert 10 < int(pushed.sum()) < 150  # ~50 expected
    assert torch.equal(new_vel.cpu()[~w_pushed], vel[~w_pushed])  # unpushed envs keep their velocity


def test_push_event_term_writes_all_envs_without_an_index_list():
    n = 5000
    written = {}
    asset = types.SimpleNamespace(data=types.SimpleNamespace(root_vel_w=torch.randn(n, 6, device=DEV)),
                                  write_root_velocity_to_sim=lambda v, env_ids=None: written.update(v=v, ids=env_ids))  # fmt: skip
    env = types.SimpleNamespace(scene={"robot": asset}, physics_dt=0.5, max_episode_length_s=1.0, device=torch.device(DEV), cfg=types.SimpleNamespace(seed=3))
    pushed = mdp.push_by_setting_velocity_with_random_envs(env, None, {"x": (1.0, 2.0)})
    assert written["ids"] is None and written["v"].shape == (n, 6)
    frac = float(pushed.float().mean())
    assert 0.2 < frac < 0.3  # p = 0.5 / 2 = 0.25
    v = written["v"]
    assert bool(((v[pushed][:, 0] >= 1.0) & (v[pushed][:, 0] <= 2.0)).all()) and float(v[pushed][:, 1:].abs().max()) == 0.0
    assert torch.equal(v[~pushed], asset.data.root_vel_w[~pushed])


def test_observation_assembly_matches_the_oracle():
    """The 45-d Solo12 policy observation (cat_flat_env_cfg.py:137-172) in one launch: bit-exact against the CPU
    restatement with shared uniforms, and with the uniforms the Philox restatement yields for the device draws."""
    from constraints_as_terminations_b200 import synthetic_env as se

    n = 3000
    g = torch.Generator().manual_seed(1)
    ang, cmd, grav = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    jp, jv, act = torch.randn(n, 12, generator=g), 8 * torch.randn(n, 12, generator=g), torch.randn(n, 12, generator=g)
    joint_ids = [0, 1, 2, 3, 4, 5, 9, 10, 11, 6, 7, 8]  # FL, FR, HR, HL in an (FL, FR, HL, HR) articulation
    robot = types.SimpleNamespace(data=types.SimpleNamespace(root_ang_vel_b=ang.to(DEV), projected_gravity_b=grav.to(DEV), joint_pos=jp.to(DEV), joint_vel=jv.to(DEV)))
    env = types.SimpleNamespace(scene={"robot": robot}, command_manager=types.SimpleNamespace(get_command=lambda name: cmd.to(DEV)),
                                action_manager=types.SimpleNamespace(_action=act.to(DEV)))  # fmt: skip
    terms = mdp.solo12_policy_terms(joint_ids)
    asm = mdp.ObservationAssembler(terms, DEV, seed=77)
    spec = [{"ids": t.ids, "noise": t.noise, "scale": t.scale} for t in terms]
    sources = [ang, cmd, grav, jp, jv, act]
    u = torch.rand(n, se.OBS_DIM, generator=g)
    got = asm.assemble(env, uniforms=u.to(DEV))
    assert got.shape == (n, se.OBS_DIM)
    assert torch.equal(got.cpu(), mdp_oracle.assemble_obs(sources, spec, u))
    got = asm.assemble(env)  # device Philox, offset 0
    u = torch.from_numpy(philox_oracle.uniform(77, philox_oracle.STREAM_UNIFORM, 0, n * se.OBS_DIM).reshape(n, se.OBS_DIM).copy())
    assert torch.equal(got.cpu(), mdp_oracle.assemble_obs(sources, spec, u))
    assert asm.rng_state.cpu().tolist() == [77, n * se.OBS_DIM]

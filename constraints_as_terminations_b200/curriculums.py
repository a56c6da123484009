"""Curriculum term that ramps a constraint's `max_p`.

Same contract as the reference (`exts/cat_envs/cat_envs/tasks/utils/cat/curriculums.py:21-41`):
the expected time-to-termination is interpolated linearly from 20 steps to
`1 / init_max_p` steps over `num_steps` common env steps, and the result is
written back through `ConstraintManager.get_term_cfg / set_term_cfg`.  The
manager picks the new value up on its next `compute()` (it re-reads every
term's `max_p` into its device-side parameter table when it changes).
"""

from __future__ import annotations

from collections.abc import Sequence


def modify_constraint_p(env, env_ids: Sequence[int], term_name: str, num_steps: int, init_max_p: float):
    progress = min(env.common_step_counter / num_steps, 1.0)
    steps_at_start = 20
    steps_at_end = 1 / init_max_p
    new_max_p = 1 / (steps_at_start + progress * (steps_at_end - steps_at_start))

    manager = env.constraint_manager
    term_cfg = manager.get_term_cfg(term_name)
    term_cfg.max_p = new_max_p
    manager.set_term_cfg(term_name, term_cfg)
    return new_max_p

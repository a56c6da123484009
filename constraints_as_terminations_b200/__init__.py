"""B200-native hot path of Gepetto/constraints-as-terminations behind the reference's manager API.

Public names mirror the reference (`exts/cat_envs/cat_envs/tasks/utils/{cat,cleanrl}`):

    ConstraintTermCfg / ConstraintTerm      per-term cfg              (cat/manager_constraint_cfg.py)
    ConstraintManager / ConstraintsManager  manager + `CaT` engine    (cat/constraint_manager.py)
    CaTEnv                                  Isaac Lab env subclass    (cat/cat_env.py; needs Isaac Lab)
    constraints                             the 15 term functions     (cat/constraints.py)
    curriculums.modify_constraint_p         max_p curriculum          (cat/curriculums.py)
    RunningMeanStd, Agent, PPO              trainer                   (cleanrl/ppo.py)

The arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of `include/catb200.h`
(`lib/libcatb200.so`, built by `__graft_entry__.build()`); python only owns tensors and orders calls.
"""

from . import constraints, curriculums
from .cat_env import CaTEnv
from .constraint_manager import CaT, ConstraintManager, ConstraintsManager
from .manager_constraint_cfg import ConstraintTerm, ConstraintTermCfg
from .ppo import PPO, Agent, PPOTrainer, RunningMeanStd
from .rl_cfg import CleanRlPpoActorCriticCfg, solo12_flat_ppo_cfg

__all__ = [
    "CaT",
    "CaTEnv",
    "ConstraintManager",
    "ConstraintsManager",
    "ConstraintTerm",
    "ConstraintTermCfg",
    "constraints",
    "curriculums",
    "PPO",
    "Agent",
    "PPOTrainer",
    "RunningMeanStd",
    "CleanRlPpoActorCriticCfg",
    "solo12_flat_ppo_cfg",
]

"""skrl front-end pieces of the reference (`exts/cat_envs/cat_envs/tasks/utils/skrl/`) on the libcatb200 kernels."""

from .ppo import compute_gae

__all__ = ["compute_gae"]

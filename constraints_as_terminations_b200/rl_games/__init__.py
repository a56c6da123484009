"""rl_games front-end pieces of the reference (`exts/cat_envs/cat_envs/tasks/utils/rl_games/`) on the libcatb200 kernels."""

from .cat_common import CaTDiscountMixin, discount_values

__all__ = ["CaTDiscountMixin", "discount_values"]

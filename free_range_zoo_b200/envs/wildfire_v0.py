"""wildfire_v0 entry points (mirrors free_range_zoo/envs/wildfire_v0.py)."""
from free_range_zoo_b200.envs.wildfire.env.wildfire import env, parallel_env, raw_env

__all__ = ['raw_env', 'env', 'parallel_env']

"""rideshare_v0 entry points (mirrors free_range_zoo/envs/rideshare_v0.py)."""
from free_range_zoo_b200.envs.rideshare.env.rideshare import env, parallel_env, raw_env

__all__ = ['raw_env', 'env', 'parallel_env']

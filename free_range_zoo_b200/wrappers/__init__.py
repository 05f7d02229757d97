from free_range_zoo_b200.wrappers.action_task import action_mapping_wrapper_v0  # noqa: F401

from gym.envs.classic_control import rendering  # noqa: F401

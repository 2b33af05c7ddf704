"""Minimal stand-in for the (unpinned, not installed) `gym` dependency of the reference.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  It exists so that the unmodified
reference environment under /root/reference/src/gym-snake can be imported in the build
container to (a) generate the golden vectors in tests/golden/ and (b) validate the
oracle restatement.  Nothing in the product package imports it.

Only the names the reference touches are provided (SURVEY.md section 8c lists them).
"""
from gym.core import Env, Wrapper, ObservationWrapper  # noqa: F401
from gym import spaces, error, utils, core  # noqa: F401
from gym.envs.registration import make, register, registry  # noqa: F401

__version__ = "0.0-shim"

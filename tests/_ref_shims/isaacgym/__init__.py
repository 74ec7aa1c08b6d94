"""Fake `isaacgym` backend (test infrastructure).

Isaac Gym Preview 3 is closed source and absent from this image.  This module
answers the ~30 gymapi calls the reference makes while building a vectorised
env (legged_robot.py:1171-1300, 939-971) from the robot URDF, and hands out
plain torch tensors for the four simulator state tensors so a test can fill
them with synthetic state.  Every simulation call is a no-op.  See SURVEY.md
Appendix B for the surface.
"""
from . import gymapi, gymtorch, gymutil, torch_utils, terrain_utils  # noqa: F401

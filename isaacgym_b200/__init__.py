"""B200-native humanoid ping-pong task hot path (drop-in for the VecTask methods of
mjmj531/isaacgym's `tasks/humanoid_pingpong*.py`).  See DESIGN.md."""
from .config import CONFIGS, TaskConfig, get_config  # noqa: F401

__version__ = "0.1.0"

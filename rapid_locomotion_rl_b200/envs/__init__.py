from .legged_robot import LeggedRobot  # noqa: F401
from .velocity_tracking import VelocityTrackingEasyEnv  # noqa: F401
from .history_wrapper import HistoryWrapper  # noqa: F401
from .curriculum import RewardThresholdCurriculum  # noqa: F401

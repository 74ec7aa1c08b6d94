from .rollout_storage import RolloutStorage  # noqa: F401
from .actor_critic import AC_Args, ActorCritic  # noqa: F401
from .ppo import PPO, PPO_Args  # noqa: F401
from .runner import Runner, RunnerArgs, export_policy  # noqa: F401

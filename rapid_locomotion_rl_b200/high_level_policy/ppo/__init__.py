"""Drop-in for high_level_policy/ppo (actor_critic.py, ppo.py, rollout_storage.py, __init__.py): the classes of
`rapid_locomotion_rl_b200.ppo` specialised the way the reference's copy differs from `mini_gym_learn/ppo`:

* `AC_Args.activation = 'tanh'` (actor_critic.py:15) - the hidden-layer epilogues of the chain kernels become
  tanh / (1 - y^2) (RL_CHAIN_EPI_BIAS_TANH / RL_CHAIN_EPI_DTANH);
* `USE_LATENT` (high_level_policy/__init__.py:5, False as shipped): no encoder, no adaptation module, `actor_body` /
  `critic_body` take the observations only (actor_critic.py:39-84, :146-150, :170-192), `PPO.update` skips the adaptation
  step and reports 0 for its loss (ppo.py:157-179), the state_dict holds the bodies and `std` only;
* `process_env_step` stores no command bins (ppo.py:81, rollout_storage.py:70);
* `RunnerArgs.num_steps_per_env = 200` (__init__.py:49).
"""
from .. import USE_LATENT
from ... import ppo as _base
from ...ppo import PPO_Args, RolloutStorage  # noqa: F401  (identical in the reference's copy)


class AC_Args(_base.AC_Args):
    activation = "tanh"


class RunnerArgs(_base.RunnerArgs):
    num_steps_per_env = 200


class ActorCritic(_base.ActorCritic):
    ac_args = AC_Args
    use_latent = USE_LATENT


class PPO(_base.PPO):
    pass


class Runner(_base.Runner):
    runner_args = RunnerArgs
    actor_critic_class = ActorCritic
    ppo_class = PPO

"""The reference's second learner package, `high_level_policy/` (high_level_policy/__init__.py:1-5): the same PPO stack as
`mini_gym_learn.ppo` with tanh networks, no latent (USE_LATENT = False: the bodies see the observations only, there is no
environment-factor encoder and no adaptation module to train), no command-bin bookkeeping and 200 steps per env and
iteration.  `rapid_locomotion_rl_b200.high_level_policy.ppo` mirrors `high_level_policy.ppo` on the same fused kernels.
"""
USE_LATENT = False

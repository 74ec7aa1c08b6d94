"""B200-native hot path of rapid-locomotion-rl: fused env step, reset / curriculum, GAE and PPO
kernels (hand-written sm_100a CUDA behind include/rl_b200.h) under the reference's
mini_gym / mini_gym_learn Python API.  See DESIGN.md."""
__version__ = "0.1.0"

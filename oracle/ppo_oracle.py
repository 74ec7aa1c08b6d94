"""ORACLE (test infrastructure, never shipped on a product path).

CPU restatement in torch of the learner side of dhruvmetha/rapid-locomotion-rl:
RolloutStorage.compute_returns (mini_gym_learn/ppo/rollout_storage.py:76-90), the ActorCritic
forward (mini_gym_learn/ppo/actor_critic.py:23-173) and one PPO.update minibatch step
(mini_gym_learn/ppo/ppo.py:94-178).  Pinned by tests/test_oracle_vs_golden.py against
tests/golden/learner.npz, which tests/golden/make_golden.py produced from the reference itself.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import it.
"""
import torch


def compute_returns(rewards, values, dones, last_values, gamma, lam):
    """rollout_storage.py:76-90.  rewards/values [T,N,1] float, dones [T,N,1] uint8, last_values [N,1].
    Returns (returns, normalised advantages), both [T,N,1]."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    adv = 0
    for t in reversed(range(T)):
        nxt = last_values if t == T - 1 else values[t + 1]
        alive = 1.0 - dones[t].float()
        delta = rewards[t] + alive * gamma * nxt - values[t]
        adv = delta + alive * gamma * lam * adv
        returns[t] = adv + values[t]
    a = returns - values
    a = (a - a.mean()) / (a.std() + 1e-8)
    return returns, a

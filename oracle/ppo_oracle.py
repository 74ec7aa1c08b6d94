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


# ----------------------------------------------------------------------------------------------
# ActorCritic + PPO.update (mini_gym_learn/ppo/actor_critic.py:23-173, ppo.py:94-178)
# ----------------------------------------------------------------------------------------------
import math

import torch.nn.functional as F

# parameter order of ActorCritic.parameters() in the reference (registration order, :55-108)
PARAM_ORDER = (["std"] + ["env_factor_encoder.%d.%s" % (i, k) for i in (0, 2, 4) for k in ("weight", "bias")]
               + ["adaptation_module.%d.%s" % (i, k) for i in (0, 2, 4) for k in ("weight", "bias")]
               + ["actor_body.%d.%s" % (i, k) for i in (0, 2, 4, 6) for k in ("weight", "bias")]
               + ["critic_body.%d.%s" % (i, k) for i in (0, 2, 4, 6) for k in ("weight", "bias")])


def mlp(p, prefix, idxs, x):
    """nn.Sequential(Linear, ELU, ..., Linear): ELU between layers, none on the output."""
    for n, i in enumerate(idxs):
        x = F.linear(x, p["%s.%d.weight" % (prefix, i)], p["%s.%d.bias" % (prefix, i)])
        if n < len(idxs) - 1:
            x = F.elu(x)
    return x


def actor_mean(p, obs, priv):
    latent = mlp(p, "env_factor_encoder", (0, 2, 4), priv)
    return mlp(p, "actor_body", (0, 2, 4, 6), torch.cat((obs, latent), dim=-1))


def critic_value(p, obs, priv):
    latent = mlp(p, "env_factor_encoder", (0, 2, 4), priv)
    return mlp(p, "critic_body", (0, 2, 4, 6), torch.cat((obs, latent), dim=-1))


def normal_log_prob(value, loc, scale):
    var = scale ** 2
    return -((value - loc) ** 2) / (2 * var) - scale.log() - math.log(math.sqrt(2 * math.pi))


def minibatch_losses(p, mb, clip=0.2, value_coef=1.0, entropy_coef=0.01, clipped_value=True):
    """ppo.py:102-144 for one minibatch `mb` (dict of tensors).  Returns (loss, surrogate, value_loss, kl_mean)."""
    mu = actor_mean(p, mb["obs"], mb["priv"])
    sigma = mu * 0.0 + p["std"]
    logp = normal_log_prob(mb["actions"], mu, sigma).sum(dim=-1)
    value = critic_value(p, mb["obs"], mb["priv"])
    entropy = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma)).sum(dim=-1)
    with torch.no_grad():
        kl = torch.sum(torch.log(sigma / mb["old_sigma"] + 1.e-5)
                       + (torch.square(mb["old_sigma"]) + torch.square(mb["old_mu"] - mu)) / (2.0 * torch.square(sigma)) - 0.5,
                       axis=-1)
        kl_mean = torch.mean(kl)
    ratio = torch.exp(logp - torch.squeeze(mb["old_logp"]))
    adv = torch.squeeze(mb["advantages"])
    surrogate = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - clip, 1.0 + clip)).mean()
    if clipped_value:
        v_clip = mb["values"] + (value - mb["values"]).clamp(-clip, clip)
        value_loss = torch.max((value - mb["returns"]).pow(2), (v_clip - mb["returns"]).pow(2)).mean()
    else:
        value_loss = (mb["returns"] - value).pow(2).mean()
    loss = surrogate + value_coef * value_loss - entropy_coef * entropy.mean()
    return loss, surrogate, value_loss, kl_mean


def adaptation_loss(p, mb):
    """ppo.py:157-164."""
    pred = mlp(p, "adaptation_module", (0, 2, 4), mb["hist"])
    with torch.no_grad():
        target = mlp(p, "env_factor_encoder", (0, 2, 4), mb["priv"])
    return F.mse_loss(pred, target)


class PPOOracle:
    """PPO.update of the reference on explicit tensors: two torch Adam optimisers over ALL parameters
    (ppo.py:44-46), clip_grad_norm_ over all, KL-adaptive learning rate, one permutation for all epochs."""

    def __init__(self, state_dict, lr=1e-3, adapt_lr=1e-3):
        self.p = {k: state_dict[k].clone().float().requires_grad_(True) for k in PARAM_ORDER}
        plist = [self.p[k] for k in PARAM_ORDER]
        self.opt = torch.optim.Adam(plist, lr=lr)
        self.opt_adapt = torch.optim.Adam(plist, lr=adapt_lr)
        self.lr = lr

    def step(self, mb, desired_kl=0.01, max_grad_norm=1.0, **kw):
        plist = [self.p[k] for k in PARAM_ORDER]
        loss, surr, vloss, kl = minibatch_losses(self.p, mb, **kw)
        if kl > desired_kl * 2.0:
            self.lr = max(1e-5, self.lr / 1.5)
        elif kl < desired_kl / 2.0 and kl > 0.0:
            self.lr = min(1e-2, self.lr * 1.5)
        for g in self.opt.param_groups:
            g["lr"] = self.lr
        self.opt.zero_grad()
        loss.backward()
        self.last_grads = {k: (None if self.p[k].grad is None else self.p[k].grad.clone()) for k in PARAM_ORDER}
        torch.nn.utils.clip_grad_norm_(plist, max_grad_norm)
        self.opt.step()
        al = adaptation_loss(self.p, mb)
        self.opt_adapt.zero_grad()
        al.backward()
        self.last_adapt_grads = {k: (None if self.p[k].grad is None else self.p[k].grad.clone()) for k in PARAM_ORDER}
        self.opt_adapt.step()
        return surr.item(), vloss.item(), al.item(), kl.item()

    def update(self, storage, perm, num_mini_batches=4, num_epochs=5, **kw):
        """storage: dict of flattened [T*N, .] tensors; perm: the minibatch permutation."""
        mbs = perm.numel() // num_mini_batches
        acc = [0.0, 0.0, 0.0]
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                idx = perm[i * mbs:(i + 1) * mbs]
                mb = {k: v[idx] for k, v in storage.items()}
                s, v, a, _ = self.step(mb, **kw)
                acc[0] += v; acc[1] += s; acc[2] += a
        n = num_epochs * num_mini_batches
        return acc[0] / n, acc[1] / n, acc[2] / n

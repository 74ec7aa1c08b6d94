"""Times the learner passes on both paths (fused chains vs one GEMM per layer) with CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapid_locomotion_rl_b200.ppo import PPO, ActorCritic  # noqa: E402
from rapid_locomotion_rl_b200.ppo import chain  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3     # us


for B in [int(x) for x in os.environ.get('BS', '24000,196608').split(',')]:
    n, T = B * 4 // 24, 24
    for use_chain in ((True,) if os.environ.get('CHAIN_ONLY') else (False, True)):
        torch.manual_seed(0)
        ac = ActorCritic(42, 18, 630, 12, device="cuda:0")
        ac.use_chain = use_chain
        ppo = PPO(ac, device="cuda:0")
        ppo.init_storage(n, T, [42], [18], [630], [12])
        st = ppo.storage
        st.observations.normal_(); st.privileged_observations.uniform_(-1, 1); st.observation_histories.normal_()
        st.actions.normal_(); st.values.normal_(); st.returns.normal_(); st.advantages.normal_()
        st.actions_log_prob.fill_(-17.0); st.mu.normal_(); st.sigma.fill_(1.0)
        idx = torch.randperm(n * T, device="cuda")[:B]
        ppo.minibatch_step(idx)
        w = ac._ws
        res = {}
        res["teacher_fwd"] = timeit(lambda: ac.forward_teacher(B, save=True))
        if use_chain:
            res["teacher_fwd_nosave"] = timeit(lambda: ac._chain(("teacher", False, True, True), lambda T: chain.teacher_forward(T, save=False)).run(B))
        res["adapt_fwd"] = timeit(lambda: ac.forward_adaptation(B, save=True))
        if use_chain:
            res["trunk_bwd"] = timeit(lambda: ac._chain(("trunk_backward",), chain.trunk_backward).run(B))
            res["adapt_bwd"] = timeit(lambda: ac._chain(("adaptation_backward",), chain.adaptation_backward_program).run(B))
        res["minibatch_step"] = timeit(lambda: ppo.minibatch_step(idx), 5)
        fl = {"teacher_fwd": 2 * (39680 + 196096 + 194688) * B, "adapt_fwd": 2 * 170048 * B,
              "trunk_bwd": 2 * (39680 + 196096 + 194688) * B, "teacher_fwd_nosave": 2 * (39680 + 196096 + 194688) * B, "adapt_bwd": 2 * 170048 * B, "minibatch_step": 3.603e6 * B}
        print("B=%d chain=%s " % (B, use_chain) + "  ".join("%s %.0f us (%.0f TF/s)" % (k, v, fl[k] / v / 1e6) for k, v in res.items()))

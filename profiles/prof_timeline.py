"""Kernel timeline of PPO.update (graph replays with the side branch) from CUPTI via torch.profiler: start / duration /
stream of every kernel of two consecutive minibatch steps in the middle of an update.  PPO_ENVS (default 4000)."""
import json
import os
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapid_locomotion_rl_b200.ppo import PPO, ActorCritic  # noqa: E402

n, T = int(os.environ.get("PPO_ENVS", 4000)), 24
torch.manual_seed(0)
ac = ActorCritic(42, 18, 630, 12, device="cuda:0")
ppo = PPO(ac, device="cuda:0")
ppo.init_storage(n, T, [42], [18], [630], [12])
st = ppo.storage


def fill():
    st.observations.normal_(); st.privileged_observations.uniform_(-1, 1); st.observation_histories.normal_()
    st.actions.normal_(); st.values.normal_(); st.returns.normal_(); st.advantages.normal_()
    st.actions_log_prob.fill_(-17.0); st.mu.normal_(); st.sigma.fill_(1.0); st.step = T


for _ in range(3):
    fill(); ppo.update()
torch.cuda.synchronize()
fill()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ppo.update()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
print("kernels in the update: %d, span %.1f us" % (len(ev), ev[-1]["ts"] + ev[-1]["dur"] - t0))
# the gather kernel opens a minibatch step on the main path: print steps 8 and 9
starts = [i for i, e in enumerate(ev) if "ppo_gather" in e["name"] and "history" not in e["name"]]
print("minibatch steps found:", len(starts))
a, b = starts[8], starts[10]
base = ev[a]["ts"]
for e in ev[a:b]:
    print("%9.1f %8.1f  s%-3s %s" % (e["ts"] - base, e["dur"], e["args"].get("stream", "?"), e["name"][:70]))
print("step period: %.1f us" % ((ev[starts[10]]["ts"] - ev[starts[8]]["ts"]) / 2))

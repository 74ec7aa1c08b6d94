"""Per-op timeline of one tile of a chain program (CTA 0), from the kernel's clock64 trace."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapid_locomotion_rl_b200.ppo import ActorCritic  # noqa: E402
from rapid_locomotion_rl_b200.ppo import chain  # noqa: E402

B = int(os.environ.get("B", 196608))
which = os.environ.get("PROG", "teacher")
it = int(os.environ.get("IT", 2))
torch.manual_seed(0)
ac = ActorCritic(42, 18, 630, 12, device="cuda:0")
w = ac.workspace(B, backward=True)
for k in ("Xp", "Xac", "Xh", "dmean", "dvalue", "dpred", "H1", "H2", "Y1", "A2", "A3", "C2", "C3", "D1", "D2"):
    w[k].copy_(torch.randn(w[k].shape, device="cuda") * 0.5)
builders = {"teacher": lambda T: chain.teacher_forward(T, save=True), "trunk_backward": chain.trunk_backward,
            "adaptation": lambda T: chain.adaptation_forward_program(T, save=True), "adaptation_backward": chain.adaptation_backward_program}
prog = ac._chain((which, "trace"), builders[which])
prog.run(B); prog.run(B)
prog.trace(it)
prog.run(B)
tr = prog.read_trace()
ld, mm, ep = tr["load"], tr["mma"], tr["epi"]
print("program %s: %d loads %d mmas %d epis; tile iteration %d of CTA 0" % (which, len(ld), len(mm), len(ep), it))
allv = [x for x in ld if x is not None] + [x for t in mm for x in t if x is not None] + [x for t in ep for x in t if x is not None]
print("span %d cycles" % (max(allv) - min(allv)))
print("LOAD issue times:", " ".join(str(x) for x in ld))
print("MMA  (waited, committed):")
for i, t in enumerate(mm):
    o = prog.mmas[i]
    print("  %3d n=%3d col=%3d k=%d acc=%d waits=%s  %s" % (i, o["n"], o["tmem_col"], o["k_steps"], o["accumulate"],
          [prog.bar_name[x.bar] for x in o["waits"]], " ".join("%7s" % x for x in t)))
print("EPI  (start, acc ready, regs, math, end):")
ordered = [o for k in range(4) for o in prog.epis if o["worker"] == k]
for i, t in enumerate(ep):
    o = ordered[i]
    print("  %3d w%d mode=%d ncols=%2d col=%3d store=%s  %s   [wait %d ld %d math %d write %d]" % (
        i, o["worker"], o["mode"], o["ncols"], o["tmem_col"], o["store_tensor"] != 255, " ".join("%7d" % x for x in t),
        t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3]))

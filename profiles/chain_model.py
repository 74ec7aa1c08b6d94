"""Timing model of a chain program (development tool, CPU only): replays the LOAD / MMA / EPI op lists of
rapid_locomotion_rl_b200/ppo/chain.py with the per-op costs measured by `rl_chain_trace` on a B200
(profiles/r01_chain_trace_*.txt, profiles/r01_chain_ncu.md) and the same synchronisation rules as the kernel
(mbarrier phases, in-order LOAD role, in-order tensor pipe, per-worker bulk-store groups).

What it is for: the surprises of round 1 were all SCHEDULE effects - a load stuck behind another ring's blocked
load, a layer boundary exposing an epilogue's full latency, the MMA warp's fixed ~850 cycles per op - and each
cost a GPU run to see.  The model shows them on the CPU: per-tile period, per-role busy / waiting time and the
barrier every role waited on most.  It does NOT predict absolute kernel times to better than ~10 % (the costs
are averages; issue-slot contention between warps is a two-level constant); `python profiles/chain_model.py`
prints model vs measured for the programs that have traces.

Calibration (six measured points, 196608 rows, grid search over six of the costs): teacher forward -10 %, trunk
backward -10 %, adaptation forward -2 %, adaptation backward -12 %, and the two programs this round replaced - the
64-column-chunk backward (544 us measured) +5 %, the single-ring adaptation forward (177 us) -6 %.  It ranks the
old and new programs correctly but overstates the backward's gain (1.30x modelled, 1.14x measured): treat a
modelled gain below ~15 % as noise.

Costs are in SM cycles (1965 MHz)."""
import heapq
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from rapid_locomotion_rl_b200.ppo import chain  # noqa: E402

COST = dict(
    load_issue=100, load_loop=420,                 # LOAD role: TMA issue after the wait; fetch of the next op
    tma_l2_base=700, tma_l2_bpc=16.0,              # weights (L2 resident): latency = base + bytes / bpc
    tma_hbm_base=1500, tma_hbm_bpc=16.0,           # tile rows (activations / inputs from HBM)
    poll=60,                                       # a satisfied mbarrier wait is seen this much later
    mma_wait=250, mma_issue=130, mma_commit=110, mma_loop=370, mma_contend=700,
    kstep=lambda n: 32 + n / 4.0,                  # tensor pipe per K16 step: shared-memory operand fetch + N / 4
    commit_lat=50,
    epi_loop=270, epi_ld32=120, epi_ld64=185,
    elu_alone=1100, elu_both=1450, delu=1300, bias=280, plain=130, f32_col=146,
    write_full=1250, write_ready=500, write_partial=3050, store_read=600,
)


class _Bar:
    def __init__(self, count):
        self.count, self.arrivals, self.tx_pending, self.done, self.waiters = count, [], 0, [], []


class Model:
    def __init__(self, prog, tiles=4, cost=None):
        prog.finalize()
        self.p, self.tiles, self.c = prog, tiles, dict(COST, **(cost or {}))
        self.bars = [_Bar(c) for c in prog.bar_count]
        self.heap, self.seq = [], 0
        self.pipe_free = 0.0
        self.busy = {}
        self.waited = {}
        self.tile_end = [0.0] * tiles
        self.workers = sorted({o["worker"] for o in prog.epis})
        self.math_until = [0.0] * (max(self.workers) + 1 if self.workers else 1)

    # ---- event machinery ----
    def at(self, t, fn):
        self.seq += 1
        heapq.heappush(self.heap, (t, self.seq, fn))

    def arrive(self, bar, t, tx_done=False):
        b = self.bars[bar]
        if tx_done:
            b.tx_pending -= 1
        else:
            b.arrivals.append(t)
        self._check(bar, t)

    def expect(self, bar):
        self.bars[bar].tx_pending += 1

    def _check(self, bar, t):
        b = self.bars[bar]
        k = len(b.done)
        if len(b.arrivals) >= (k + 1) * b.count and b.tx_pending == 0:
            b.done.append(t)
            for (need, resume) in [w for w in b.waiters if w[0] <= len(b.done)]:
                b.waiters.remove((need, resume))
                resume(t)

    def wait(self, role, w, it, now, then):
        """Calls then(t) when barrier phase (w.need + it * phases per tile) is complete, t >= now."""
        if w is None:
            return then(now)
        need = w.need + it * self.p.bar_phases[w.bar]
        b = self.bars[w.bar]
        name = self.p.bar_name[w.bar]

        def resume(t):
            t2 = max(now, t + self.c["poll"])
            self.waited.setdefault(role, {}).setdefault(name, 0.0)
            self.waited[role][name] += t2 - now
            self.at(t2, lambda: then(t2))
        if need == 0 or len(b.done) >= need:
            return then(now) if need == 0 else resume(b.done[need - 1] - self.c["poll"] if b.done[need - 1] <= now else b.done[need - 1])
        b.waiters.append((need, resume))

    def add_busy(self, role, dt):
        self.busy[role] = self.busy.get(role, 0.0) + dt

    # ---- roles ----
    def run(self):
        for k in (0, 1):           # two issuing warps per role, each with its own in-order list
            self._load(0, 0, 0.0, k)
            self._mma(0, 0, 0.0, k)
        for wk in self.workers:
            self._epi(wk, 0, 0, 0.0, {"issued": [], })
        while self.heap:
            t, _, fn = heapq.heappop(self.heap)
            fn()
        return self

    def _load(self, it, i, now, k=0):
        p, c = self.p, self.c
        ops = [o for o in p.loads if o.get("issuer", 0) == k]
        if it >= self.tiles or not ops:
            return
        if i >= len(ops):
            return self._load(it + 1, 0, now, k)
        o = ops[i]

        def go(t):
            self.expect(o["full_bar"])
            self.arrive(o["full_bar"], t + c["load_issue"])
            hbm = bool(o["tile_rows"])
            lat = (c["tma_hbm_base"] + o["bytes"] / c["tma_hbm_bpc"]) if hbm else (c["tma_l2_base"] + o["bytes"] / c["tma_l2_bpc"])
            self.at(t + c["load_issue"] + lat, lambda: self.arrive(o["full_bar"], t + c["load_issue"] + lat, tx_done=True))
            self.add_busy("load%d" % k, c["load_issue"] + c["load_loop"])
            self.at(t + c["load_issue"] + c["load_loop"], lambda: self._load(it, i + 1, t + c["load_issue"] + c["load_loop"], k))
        self.wait("load%d" % k, o["wait"], it, now, go)

    def _mma(self, it, i, now, k=0):
        p, c = self.p, self.c
        ops = [o for o in p.mmas if o.get("issuer", 0) == k]
        if it >= self.tiles or not ops:
            return
        if i >= len(ops):
            return self._mma(it + 1, 0, now, k)
        o = ops[i]
        waits = list(o["waits"])
        kk = k
        t0 = now + c["mma_wait"]

        def after(k, t):
            if k < len(waits):
                return self.wait("mma%d" % kk, waits[k], it, t, lambda t2: after(k + 1, t2))
            stall = c["mma_contend"] if sum(m > t for m in self.math_until) >= 2 else 0.0   # >= 2 workers in their math phase
            t_iss = t + c["mma_issue"] + stall
            start = max(t_iss, self.pipe_free)
            self.pipe_free = start + o["k_steps"] * c["kstep"](o["n"])
            done = self.pipe_free + c["commit_lat"]
            t_end = t_iss + c["mma_commit"]
            for b in o["commits"]:
                self.at(max(done, t_end), (lambda b=b, d=max(done, t_end): self.arrive(b, d)))
            self.add_busy("mma%d" % kk, c["mma_wait"] + c["mma_issue"] + stall + c["mma_commit"] + c["mma_loop"])
            self.at(t_end + c["mma_loop"], lambda: self._mma(it, i + 1, t_end + c["mma_loop"], kk))
        after(0, t0)

    def _epi(self, wk, it, i, now, sg):
        p, c = self.p, self.c
        ops = [o for o in p.epis if o["worker"] == wk]
        if it >= self.tiles or not ops:
            return
        if i >= len(ops):
            self.tile_end[it] = max(self.tile_end[it], now)           # a tile ends when its last epilogue op does
            return self._epi(wk, it + 1, 0, now, sg)
        o = ops[i]
        role = "epi%d" % wk
        nc, mode = o["ncols"], o["mode"]

        def acc_ready(t):
            t_ld = t + (c["epi_ld64"] if nc > 32 else c["epi_ld32"])
            if o["arrive_acc_free"] != chain.NONE:
                for _ in range(4):
                    self.arrive(o["arrive_acc_free"], t_ld)
            if mode == chain.EPI_DELU:
                return self.wait(role, o["wait_aux"], it, t_ld, lambda t2: math_done(t2 + c["delu"], t_ld))
            if mode == chain.EPI_BIAS_ELU:
                others = sum(m > t_ld for k, m in enumerate(self.math_until) if k != wk)
                dur = c["elu_alone"] + (c["elu_both"] - c["elu_alone"]) * others       # MUFU pipe shared by the workers
            elif mode == chain.EPI_BIAS:
                dur = c["bias"]
            elif mode == chain.EPI_PLAIN:
                dur = c["plain"]
            else:
                dur = c["f32_col"] * max(2, nc)
            math_done(t_ld + dur, t_ld)

        def math_done(t, t_start):
            self.math_until[wk] = t
            if mode == chain.EPI_BIAS_F32:
                self.add_busy(role, t - t_start)
                return self.at(t + c["epi_loop"], lambda: self._epi(wk, it, i + 1, t + c["epi_loop"], sg))
            # box reuse: a bulk store of this worker may still read the unit
            pend = o["store_wait_pending"]
            t2 = t
            if pend >= 0 and len(sg["issued"]) > pend:
                t2 = max(t, sg["issued"][-(pend + 1)] if pend < len(sg["issued"]) else t)
            self.wait(role, o["wait_dst"], it, t2, lambda t3: write(t3, t_start))

        def write(t, t_start):
            full = o["dst_col0"] == 0 and nc == 64
            dur = c["write_full"] if full else c["write_partial"]
            t_ready = t + (c["write_ready"] if full else dur - 300)
            if o["arrive_dst_ready"] != chain.NONE:
                for _ in range(4):
                    self.at(t_ready, (lambda: self.arrive(o["arrive_dst_ready"], t_ready)))
            if o["release_aux"] != chain.NONE:
                self.at(t_ready + 100, lambda: self.arrive(o["release_aux"], t_ready + 100))
            t_end = t + dur
            if o["store_tensor"] != chain.NONE:
                sg["issued"].append(t_end + c["store_read"])          # time at which the store has read the box
                if o["release_after_store"] != chain.NONE:
                    t_end += c["store_read"]
                    self.at(t_end, lambda: self.arrive(o["release_after_store"], t_end))
            self.add_busy(role, t_end - t_start)
            self.at(t_end + c["epi_loop"], lambda: self._epi(wk, it, i + 1, t_end + c["epi_loop"], sg))
        self.wait(role, o["wait_acc"], it, now, acc_ready)

    # ---- report ----
    def report(self):
        ends = self.tile_end
        period = (ends[-1] - ends[0]) / (len(ends) - 1) if len(ends) > 1 else ends[0]
        out = {"period": period, "first_tile": ends[0], "busy_per_tile": {k: v / self.tiles for k, v in self.busy.items()}}
        out["top_waits"] = {r: sorted(((v / self.tiles, k) for k, v in w.items()), reverse=True)[:3] for r, w in self.waited.items()}
        return out


def _cpu_tensors(rows=512):
    import chainkit as ck
    return ck.make_tensors(rows, 0)


MEASURED = {   # cycles per tile in steady state = kernel time x 1.965 GHz / (1536 tiles / 148 CTAs), 196608 rows
    "teacher_forward": 405e-6 * 1.965e9 / (1536 / 148.0),
    "trunk_backward": 476e-6 * 1.965e9 / (1536 / 148.0),
    "adaptation_forward": 164e-6 * 1.965e9 / (1536 / 148.0),
    "adaptation_backward": 82e-6 * 1.965e9 / (1536 / 148.0),
}


def main():
    T = _cpu_tensors()
    progs = [chain.teacher_forward_program(T), chain.trunk_backward_program(T), chain.adaptation_forward_program(T),
             chain.adaptation_backward_program(T)]
    for prog in progs:
        r = Model(prog, tiles=6).run().report()
        meas = MEASURED.get(prog.name)
        print("%-20s %3d loads %3d mmas %3d epis | model %6.0f cycles / tile (first tile %6.0f) | measured %6.0f | busy %s" % (
            prog.name, len(prog.loads), len(prog.mmas), len(prog.epis), r["period"], r["first_tile"], meas or 0,
            {k: int(v) for k, v in sorted(r["busy_per_tile"].items())}))
        for role, ws in sorted(r["top_waits"].items()):
            print("    %-5s waits most on: %s" % (role, ", ".join("%s %.0f" % (k, v) for v, k in ws)))


if __name__ == "__main__":
    main()

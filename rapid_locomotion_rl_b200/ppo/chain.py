"""Host side of the fused MLP chains (csrc/chain.cu, include/rl_b200.h "Fused MLP chains").

A chain program is three op lists (LOAD / MMA / EPI) that the three warp roles of one persistent CTA
execute in order for every 128-row tile.  This module
  * builds the programs for the learner's networks (mini_gym_learn/ppo/actor_critic.py:38-100 forward,
    and the dgrad half of the autograd backward behind ppo.py:146-168) with a small resource-tracking
    builder (shared-memory units, ring stages, tensor-memory regions, mbarrier phase bookkeeping), and
  * EMULATES a program on the CPU: the three roles and the asynchronous agents (TMA loads / stores, the
    tensor-core pipe) run under a randomised scheduler with the exact mbarrier semantics of the device
    (parity waits, arrival counts, transaction bytes), checking deadlock freedom, phase-parity
    soundness and shared-memory hazards, and computing the numerical result in bf16 / fp32.
The emulator is how a schedule is proved before it ever reaches the GPU; the product path only uses
`ChainProgram.compile()` + `run()`.
"""
import ctypes as C
import random

import torch

from .. import _lib

UNIT = 16384
NONE = 255
EPI_BIAS_ELU, EPI_BIAS, EPI_BIAS_F32, EPI_DELU, EPI_PLAIN = range(5)
# The programs name the hidden-layer epilogues EPI_BIAS_ELU / EPI_DELU; a program whose `activation` is "tanh" (the
# reference's high_level_policy learner, high_level_policy/ppo/actor_critic.py:15) is packed with these kernel modes instead
EPI_BIAS_TANH, EPI_DTANH = 5, 6
ACTIVATIONS = ("elu", "tanh")


class _Wait:
    __slots__ = ("bar", "need")

    def __init__(self, bar, need):
        self.bar, self.need = bar, need

    def key(self):
        return (self.bar, self.need)


class StageUse:
    """One landing of a TMA box in a ring stage.  `full` is None until the first MMA that reads the box names its issuer
    (every issuing warp has its own "full" barrier per stage: a barrier must have a single waiting agent)."""
    __slots__ = ("stage", "off", "full", "empty_bar", "released", "ring", "load_op")

    def __init__(self, stage, off, full, empty_bar, ring=None, load_op=None):
        self.stage, self.off, self.full, self.empty_bar, self.released = stage, off, full, empty_bar, False
        self.ring, self.load_op = ring, load_op


class BoxUse:
    """One content of an activation unit (written by the epilogue, or loaded by TMA for inputs).  A pool box's `ready`
    is None until its first reading MMA names the issuer whose "ready" barrier the producing epilogue op arrives on."""
    __slots__ = ("unit", "off", "ready", "free_bar", "released", "has_reader", "pool_index", "producer_op")

    def __init__(self, unit, off, ready, free_bar, pool_index=None, producer_op=None):
        self.unit, self.off, self.ready, self.free_bar = unit, off, ready, free_bar
        self.released, self.has_reader = False, True
        self.pool_index, self.producer_op = pool_index, producer_op


class AccUse:
    __slots__ = ("region", "col", "full", "free_bar", "wait_free", "started", "closed", "last_op", "alias_free")

    def __init__(self, region, col, free_bar, wait_free):
        self.region, self.col, self.free_bar, self.wait_free = region, col, free_bar, wait_free
        self.full, self.started, self.closed = None, False, False
        self.last_op = {}               # epilogue worker -> its last op on this accumulator use
        self.alias_free = []            # free waits of overlapping regions (ChainProgram(alias_waits=True))


class ChainProgram:
    """Builder + container of one chain program."""

    def __init__(self, n_pool, n_stages, n_inputs, regions, name="chain", region_worker=None, stage_units=1, rings=None,
                 n_workers=2, alias_waits=False, region_issuer=None, n_loaders=1):
        """Shared-memory units: [inputs | pool | ring stages]; `regions`: {name: (first tmem column, width)};
        `region_worker`: regions whose accumulator uses are read by ONE epilogue op -> the worker that owns
        them (a waiter must see every phase of a barrier, so such a region cannot change hands); the other
        regions are read by all workers, chunk c by worker c % n_workers.
        `n_workers`: epilogue warp groups.  The kernel (csrc/chain.cu) has two; programs for three can be built,
        emulated and timed (profiles/chain_model.py) - rl_chain_create rejects them until the kernel grows a third.
        `rings`: [(name, n_stages, units per stage)] - several independent TMA rings (default: one ring "main" of
        n_stages x stage_units); the LOAD role still issues every load in program order."""
        self.name = name
        self.n_inputs, self.n_pool, self.n_workers = n_inputs, n_pool, n_workers
        # alias_waits: regions may share tensor-memory columns; the first MMA of a use then also waits until the latest
        # use of every overlapping region has been read by the epilogue (acc(..., implied=) names the ones another
        # wait of the same op already implies)
        self.alias_waits = alias_waits
        import os
        self.stagger_ns = int(os.environ.get("RL_CHAIN_STAGGER_NS", "0"))
        # Two issuing warps per role (csrc/chain.cu): MMAs are dealt by accumulator region (`region_issuer`: region -> 0 / 1;
        # the MMAs into one accumulator must come from one thread), loads by ring stage (stage s -> loader s % n_loaders,
        # so every stage keeps a single waiting loader).  Nothing orders the two MMA issuers against each other except the
        # mbarriers, so programs whose regions share columns need alias_waits.
        self.region_issuer = dict(region_issuer or {})
        self.n_loaders = n_loaders
        assert n_loaders in (1, 2) and all(v in (0, 1) for v in self.region_issuer.values())
        rings = rings or [("main", n_stages, stage_units)]
        self.n_stages, self.stage_units = rings[0][1], rings[0][2]      # of the default (first) ring
        self.tensors = []                # (torch tensor 2-D view, box_rows)
        self.bar_count = []
        self.bar_phases = []             # completions per tile
        self.bar_name = []
        self.loads, self.mmas, self.epis = [], [], []
        self.outputs = [None] * _lib.DEFINES["RL_CHAIN_MAX_OUTPUTS"]
        self.params = None
        self.regions = dict(regions)
        # resources
        self.pool_units = [n_inputs + i for i in range(n_pool)]
        # a barrier must have ONE waiting agent that sees every phase (waits name a phase only by its parity):
        # each stage has one "full" barrier per consumer role and the producer signals the one of the role
        # that will read this particular landing
        self.rings = {}
        unit = n_inputs + n_pool
        for rname, ns, su in rings:
            tag = "" if rname == "main" else rname + "."
            self.rings[rname] = dict(n=ns, su=su, unit0=[unit + i * su for i in range(ns)], pos=0, uses=[0] * ns, tag=tag,
                                     empty=[self._bar(1, "%sstage%d.empty" % (tag, i)) for i in range(ns)],
                                     full={"mma": [self._bar(1, "%sstage%d.full.mma" % (tag, i)) for i in range(ns)]})
            unit += ns * su
        self.default_ring = rings[0][0]
        self.n_units = unit
        assert self.n_units <= _lib.DEFINES["RL_CHAIN_MAX_UNITS"]
        self.pool_ready = [{0: self._bar(4, "pool%d.ready" % i)} for i in range(n_pool)]     # per MMA issuer (1: on demand)
        self.pool_free = [self._bar(1, "pool%d.free" % i) for i in range(n_pool)]
        self.pool_pos = 0
        self.pool_last = [None] * n_pool           # last BoxUse per pool unit
        self.acc_full = {r: self._bar(1, "acc_%s.full" % r) for r in regions}
        self.acc_free = {r: self._bar(4, "acc_%s.free" % r) for r in regions}
        self.acc_uses = {r: 0 for r in regions}
        self.acc_open = {r: None for r in regions}
        self.input_full, self.input_free, self.input_loads = {}, {}, {}
        self.waited = {"load": set(), "mma0": set(), "mma1": set()}
        self.waited.update({"epi%d" % k: set() for k in range(n_workers)})
        # two epilogue workers (4 warps each) take the EPI ops alternately; a pool unit always belongs to the
        # worker of its parity, so the TMA-store bookkeeping (one bulk group per store, committed by the
        # worker's first thread) stays within one thread
        self.region_worker = dict(region_worker or {"C0": 0, "C1": 1})
        self.n_stores = [0] * n_workers
        self.unit_last_store = {}        # unit -> store index within the tile (of the unit's worker)
        self.acc_participants = {r: set() for r in regions}
        self._finalized = False

    # ---------------------------------------------------------------------------------------------
    def _bar(self, count, name):
        self.bar_count.append(count)
        self.bar_phases.append(0)
        self.bar_name.append(name)
        assert len(self.bar_count) <= _lib.DEFINES["RL_CHAIN_MAX_BARRIERS"]
        return len(self.bar_count) - 1

    def _signal(self, bar):
        """One more phase completion of `bar` per tile; returns its ordinal (1-based)."""
        self.bar_phases[bar] += 1
        return self.bar_phases[bar]

    def tensor(self, t, box_rows):
        assert t.dim() == 2 and t.dtype == torch.bfloat16 and t.stride(1) == 1
        assert (t.stride(0) * 2) % 16 == 0 and t.data_ptr() % 16 == 0, "TMA operand alignment"
        assert 8 <= box_rows <= 256 and box_rows * 128 <= UNIT * max(r["su"] for r in self.rings.values())
        self.tensors.append((t, box_rows))
        assert len(self.tensors) <= _lib.DEFINES["RL_CHAIN_MAX_TENSORS"]
        return len(self.tensors) - 1

    def output(self, idx, t):
        assert t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1
        self.outputs[idx] = t
        return idx

    def _w(self, role, wait):
        """Drop waits this role has already performed (same barrier, same completion)."""
        if wait is None:
            return None
        if wait.key() in self.waited[role]:
            return None
        self.waited[role].add(wait.key())
        return wait

    # ---- inputs: TMA-loaded [128 x 64] boxes in dedicated units ------------------------------------
    def load_input(self, slot, tensor, col0, extra_free_arrivals=0):
        """Loads the tile's rows of `tensor` (64 columns from col0) into input unit `slot`."""
        assert slot < self.n_inputs and slot not in self.input_loads
        full = self._bar(1, "in%d.full" % slot)
        free = self._bar(1 + extra_free_arrivals, "in%d.free" % slot)
        self.input_full[slot], self.input_free[slot] = full, free
        ordn = self._signal(full)
        self.loads.append(dict(wait=_Wait(free, 0), full_bar=full, tensor=tensor, smem_off=slot * UNIT, col0=col0, row0=0,
                               bytes=UNIT, tile_rows=1, issuer=0))
        use = BoxUse(slot, slot * UNIT, _Wait(full, ordn), free)
        self.input_loads[slot] = use
        return use

    # ---- ring stages ------------------------------------------------------------------------------
    def load_stage(self, tensor, col0, row0, tile_rows=False, consumer="mma", ring=None):
        """One TMA box into the next stage of `ring`; `consumer`: the role that will wait for it ("mma", "epi0", "epi1")."""
        t, box_rows = self.tensors[tensor]
        R = self.rings[ring or self.default_ring]
        assert box_rows * 128 <= UNIT * R["su"], "box does not fit a stage of this ring"
        s = R["pos"] % R["n"]
        R["pos"] += 1
        wait = _Wait(R["empty"][s], R["uses"][s])
        R["uses"][s] += 1
        off = R["unit0"][s] * UNIT
        op = dict(wait=wait, full_bar=None, tensor=tensor, smem_off=off, col0=col0, row0=row0,
                  bytes=box_rows * 128, tile_rows=int(tile_rows), issuer=s % self.n_loaders)
        self.loads.append(op)
        use = StageUse(s, off, None, R["empty"][s], ring=R, load_op=op)
        if consumer != "mma":
            self._resolve_stage(use, consumer)
        return use

    def _resolve_stage(self, use, consumer):
        """Names the "full" barrier of a landing once its consumer is known ("mma" / "mma1" / "epiK")."""
        if use.full is not None:
            assert consumer in self.bar_name[use.full.bar].split(".")[-1:], "stage box read by two different consumers"
            return
        R, s = use.ring, use.stage
        if consumer not in R["full"]:
            R["full"][consumer] = [self._bar(1, "%sstage%d.full.%s" % (R["tag"], i, consumer)) for i in range(R["n"])]
        full = R["full"][consumer][s]
        use.load_op["full_bar"] = full
        use.full = _Wait(full, self._signal(full))

    def worker_for(self, acc, col):
        """Epilogue worker that reads accumulator columns [col, col + 64) of this use."""
        return self.region_worker[acc.region] if acc.region in self.region_worker else (col // 64) % self.n_workers

    # ---- tensor-memory accumulators -----------------------------------------------------------------
    def acc(self, region, implied=(), carry=()):
        assert self.acc_open[region] is None or self.acc_open[region].closed, "accumulator %s still open" % region
        col, width = self.regions[region]
        a = AccUse(region, col, self.acc_free[region], _Wait(self.acc_free[region], self.acc_uses[region]))
        if self.alias_waits:
            for r2, (c2, w2) in self.regions.items():
                if r2 != region and r2 not in implied and c2 < col + width and col < c2 + w2:
                    assert self.acc_open[r2] is None or self.acc_open[r2].closed, "%s opened over the open accumulator %s" % (region, r2)
                    # a region not used yet in this tile was last used by the PREVIOUS tile: its reads are normally
                    # implied by that tile's program order; `carry` names the ones that are not (need 0 = the previous
                    # tile's last completion)
                    if self.acc_uses[r2] > 0 or r2 in carry:
                        a.alias_free.append(_Wait(self.acc_free[r2], self.acc_uses[r2]))
        self.acc_uses[region] += 1
        self.acc_open[region] = a
        return a

    # ---- MMA ------------------------------------------------------------------------------------------
    def mma(self, a, b, n, acc, col_off=0, k_steps=4, accumulate=True, acc_last=False, a_release=False, b_release=True):
        """acc[:, col_off : col_off + n] (+)= A box * B box^T over k_steps K16 steps."""
        waits, commits = [], []
        issuer = self.region_issuer.get(acc.region, 0)
        for x in (a, b):
            if isinstance(x, StageUse):
                self._resolve_stage(x, "mma" if issuer == 0 else "mma%d" % issuer)
            elif x.ready is None:
                self._resolve_box(x, issuer)
        for w in [a.full if isinstance(a, StageUse) else a.ready, b.full] + ([] if acc.started else [acc.wait_free] + acc.alias_free):
            w = self._w("mma%d" % issuer, w)
            if w is not None:
                waits.append(w)
        acc.started = True
        if b_release:
            assert not b.released
            b.released = True
            commits.append(b.empty_bar)
            self._signal(b.empty_bar)
        if a_release:
            assert not a.released
            a.released = True
            bar = a.empty_bar if isinstance(a, StageUse) else a.free_bar
            commits.append(bar)
            self._signal(bar)
        if acc_last:
            acc.full = _Wait(self.acc_full[acc.region], self._signal(self.acc_full[acc.region]))
            commits.append(self.acc_full[acc.region])
        width = self.regions[acc.region][1]
        assert col_off + n <= width and n % 16 == 0 and 16 <= n <= 256
        assert len(waits) <= 4 and len(commits) <= 3, (len(waits), len(commits))
        self.mmas.append(dict(a_off=a.off, b_off=b.off, n=n, tmem_col=acc.col + col_off, k_steps=k_steps,
                              accumulate=int(accumulate), waits=waits, commits=commits, issuer=issuer))

    def _resolve_box(self, box, issuer):
        """The producing epilogue op arrives on the "ready" barrier of the issuer that reads the box."""
        i = box.pool_index
        if issuer not in self.pool_ready[i]:
            self.pool_ready[i][issuer] = self._bar(4, "pool%d.ready.mma%d" % (i, issuer))
        bar = self.pool_ready[i][issuer]
        box.producer_op["arrive_dst_ready"] = bar
        box.ready = _Wait(bar, self._signal(bar))

    # ---- epilogue ---------------------------------------------------------------------------------------
    def _epi_common(self, acc, col, ncols, mode, bias_off, last):
        assert acc.full is not None, "accumulator has no closing MMA yet"
        w = self.worker_for(acc, col)
        assert acc.region not in self.region_worker or not acc.last_op, "region %s is single-op" % acc.region
        op = dict(worker=w, wait_acc=self._w("epi%d" % w, acc.full), wait_dst=None, wait_aux=None, arrive_acc_free=NONE,
                  arrive_dst_ready=NONE, release_aux=NONE, mode=mode, ncols=ncols, dst_col0=0, tmem_col=acc.col + col,
                  store_tensor=NONE, store_wait_pending=-1, release_after_store=NONE, out_id=NONE, out_ld=0,
                  bias_off=bias_off, dst_off=0, aux_off=0, store_col0=0, delay_ns=0)
        # de-phase the workers: the first op of worker w on an accumulator use several workers read starts w x stagger later
        if self.stagger_ns and acc.region not in self.region_worker and w not in acc.last_op and w > 0:
            op["delay_ns"] = w * self.stagger_ns
        acc.last_op[w] = op
        if last:
            # every participating worker arrives (4 warps each) after ITS last read of this accumulator use
            for o in acc.last_op.values():
                o["arrive_acc_free"] = acc.free_bar
            self.acc_participants[acc.region].add(len(acc.last_op))
            self._signal(acc.free_bar)
            acc.closed = True
        return op

    def epi_box(self, acc, col, mode, bias_off=0, ncols=64, aux=None, store=None, last=False, has_reader=True):
        """Accumulator columns [col, col+ncols) -> bf16 box in the next free pool unit of the worker's parity."""
        op = self._epi_common(acc, col, ncols, mode, bias_off, last)
        w = op["worker"]
        # next unit, in round-robin order, whose content already has its releasing reader in the program
        for probe in range(self.n_pool):
            i = (self.pool_pos + probe) % self.n_pool
            prev = self.pool_last[i]
            if i % self.n_workers == w and (prev is None or prev.released or not prev.has_reader):
                break
        else:
            raise AssertionError("%s: every pool unit of worker %d holds a live box" % (self.name, w))
        self.pool_pos = i + 1
        unit = self.pool_units[i]
        # frees emitted so far for this unit == completions the writer must have seen
        op["wait_dst"] = self._w("epi%d" % w, _Wait(self.pool_free[i], self.bar_phases[self.pool_free[i]]))
        op["arrive_dst_ready"] = None          # named by the first MMA that reads the box (_resolve_box)
        op["dst_off"] = unit * UNIT
        # the unit may still be read by a TMA store this worker issued earlier (this tile or the previous one)
        op["_store_dep"] = (self.unit_last_store.get(unit), self.n_stores[w])
        op["_unit"] = unit
        if mode == EPI_DELU:
            assert aux is not None and not aux.released
            assert "epi%d" % w in self.bar_name[aux.full.bar], "aux box was loaded for another consumer"
            op["wait_aux"] = self._w("epi%d" % w, aux.full)
            op["aux_off"] = aux.off
            op["release_aux"] = aux.empty_bar
            aux.released = True
            self._signal(aux.empty_bar)
        if store is not None:
            op["store_tensor"], op["store_col0"] = store
            self.unit_last_store[unit] = self.n_stores[w]
            self.n_stores[w] += 1
        use = BoxUse(unit, unit * UNIT, None, self.pool_free[i], pool_index=i, producer_op=op)
        use.has_reader = has_reader
        self.pool_last[i] = use
        self.epis.append(op)
        return use

    def epi_merge(self, acc, col, ncols, mode, bias_off, box, dst_col0, ready_bar_count4, store=None, release_after_store=NONE,
                  last=True):
        """Accumulator columns -> `ncols` columns at dst_col0 of an already loaded input box (the encoder latent
        next to the observations).  Waits for the box's TMA load, signals `ready_bar_count4`."""
        op = self._epi_common(acc, col, ncols, mode, bias_off, last)
        w = op["worker"]
        op["wait_dst"] = self._w("epi%d" % w, box.ready)
        op["dst_off"], op["dst_col0"] = box.off, dst_col0
        ordn = self._signal(ready_bar_count4)
        op["arrive_dst_ready"] = ready_bar_count4
        if store is not None:
            op["store_tensor"], op["store_col0"] = store
            self.n_stores[w] += 1
            op["release_after_store"] = release_after_store
        self.epis.append(op)
        merged = BoxUse(box.unit, box.off, _Wait(ready_bar_count4, ordn), box.free_bar)
        return merged

    def epi_out(self, acc, col, ncols, bias_off, out_id, last=True):
        op = self._epi_common(acc, col, ncols, EPI_BIAS_F32, bias_off, last)
        op["out_id"], op["out_ld"] = out_id, self.outputs[out_id].stride(0)
        self.epis.append(op)

    # ---------------------------------------------------------------------------------------------
    def finalize(self):
        if self._finalized:
            return self
        for rname, R in self.rings.items():
            for s in range(R["n"]):
                assert self.bar_phases[R["empty"][s]] == R["uses"][s], "ring %s stage %d: uses and releases differ" % (rname, s)
        for slot, use in self.input_loads.items():
            assert self.bar_phases[self.input_free[slot]] == 1, "input %d is never released (or more than once)" % slot
        for r, parts in self.acc_participants.items():
            assert len(parts) <= 1, "region %s is read by %s epilogue workers in different uses" % (r, sorted(parts))
            if parts:
                self.bar_count[self.acc_free[r]] = 4 * next(iter(parts))
        for op in self.epis:
            dep = op.pop("_store_dep", None)
            unit = op.pop("_unit", None)
            if dep is None:
                continue
            k, m = dep                                # the unit's last store / stores this worker issued so far
            if k is not None:
                pend = m - 1 - k
            elif unit in self.unit_last_store:        # last stored in the previous tile
                pend = m - 1 - (self.unit_last_store[unit] - self.n_stores[op["worker"]])
            else:
                pend = None
            op["store_wait_pending"] = -1 if pend is None else max(0, min(7, pend))
        for op in self.epis:
            if op["arrive_dst_ready"] is None:        # a box no MMA reads (stored only)
                op["arrive_dst_ready"] = NONE
        for o in self.loads:
            assert o["full_bar"] is not None, "a stage box nobody reads"
        # a parity wait names a phase only modulo 2: every barrier an MMA issuer waits on must have that single waiter
        seen = {}
        for o in self.mmas:
            for w in o["waits"]:
                assert seen.setdefault(w.bar, o["issuer"]) == o["issuer"], \
                    "%s: barrier %s is waited on by both MMA issuers" % (self.name, self.bar_name[w.bar])
        # the kernel takes each role's list grouped by issuer (program order kept within an issuer)
        self.loads.sort(key=lambda o: o["issuer"])
        self.mmas.sort(key=lambda o: o["issuer"])
        self._finalized = True
        return self

    def _spec(self, w):
        if w is None:
            return NONE
        flip = self.bar_phases[w.bar] & 1
        return w.bar | (((w.need - 1) & 1) << 8) | (flip << 9)

    def pack(self):
        """ctypes arrays + RlChainDesc (keeps references alive on self)."""
        self.finalize()
        L = (_lib.RlChainLoadOp * max(1, len(self.loads)))()
        for i, o in enumerate(self.loads):
            x = L[i]
            x.wait, x.full_bar, x.tensor, x.smem_off = self._spec(o["wait"]), o["full_bar"], o["tensor"], o["smem_off"]
            x.col0, x.row0, x.expect_bytes, x.tile_rows, x.issuer = o["col0"], o["row0"], o["bytes"], o["tile_rows"], o["issuer"]
        M = (_lib.RlChainMmaOp * max(1, len(self.mmas)))()
        for i, o in enumerate(self.mmas):
            x = M[i]
            x.a_off, x.b_off, x.n, x.tmem_col, x.k_steps, x.accumulate = o["a_off"], o["b_off"], o["n"], o["tmem_col"], o["k_steps"], o["accumulate"]
            x.issuer = o["issuer"]
            ws = [self._spec(w) for w in o["waits"]] + [NONE] * 4
            x.wait0, x.wait1, x.wait2, x.wait3 = ws[:4]
            cs = list(o["commits"]) + [NONE] * 3
            x.commit0, x.commit1, x.commit2 = cs[:3]
        E = (_lib.RlChainEpiOp * max(1, len(self.epis)))()
        for i, o in enumerate(self.epis):
            x = E[i]
            x.wait_acc, x.wait_dst, x.wait_aux = self._spec(o["wait_acc"]), self._spec(o["wait_dst"]), self._spec(o["wait_aux"])
            for k in ("worker", "arrive_acc_free", "arrive_dst_ready", "release_aux", "mode", "ncols", "dst_col0", "tmem_col", "store_tensor",
                      "store_wait_pending", "release_after_store", "out_id", "out_ld", "bias_off", "dst_off", "aux_off", "store_col0",
                      "delay_ns"):
                setattr(x, k, o[k])
            if getattr(self, "activation", "elu") == "tanh":
                x.mode = {EPI_BIAS_ELU: EPI_BIAS_TANH, EPI_DELU: EPI_DTANH}.get(o["mode"], o["mode"])
        d = _lib.RlChainDesc()
        for i, (t, box_rows) in enumerate(self.tensors):
            T = d.tensors[i]
            T.base, T.rows, T.cols, T.ld, T.box_rows = t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), box_rows
        d.n_tensors, d.n_units, d.n_barriers = len(self.tensors), self.n_units, len(self.bar_count)
        d.n_loads, d.n_mmas, d.n_epis = len(self.loads), len(self.mmas), len(self.epis)
        for i, c in enumerate(self.bar_count):
            d.barrier_count[i] = c
        d.loads_host, d.mmas_host, d.epis_host = C.cast(L, C.c_void_p), C.cast(M, C.c_void_p), C.cast(E, C.c_void_p)
        d.params = self.params.data_ptr() if self.params is not None else None
        for i, t in enumerate(self.outputs):
            d.outputs[i] = t.data_ptr() if t is not None else None
        self._packed = (L, M, E, d)
        return d

    # ---- product path -----------------------------------------------------------------------------------
    def compile(self):
        lib = _lib.lib()
        d = self.pack()
        h = C.c_void_p()
        _lib.check(lib.rl_chain_create(C.byref(d), C.byref(h)))
        self._handle, self._libref = h, lib
        return self

    def run(self, rows, stream=None, tiles=None):
        """All row tiles of the `rows`-row batch, or only tiles [tiles[0], tiles[1]) (128 rows each)."""
        st = _lib.current_stream() if stream is None else stream
        if tiles is None:
            _lib.check(self._libref.rl_chain_run(self._handle, int(rows), st))
        else:
            _lib.check(self._libref.rl_chain_run_tiles(self._handle, int(rows), int(tiles[0]), int(tiles[1]), st))

    def set_ppo_loss(self, loss):
        """loss: an `_lib.RlChainPpoLoss` (the value-output epilogue of the launches that follow also evaluates the PPO
        loss of its rows: include/rl_b200.h) or None (off)."""
        self._loss_ref = loss
        _lib.check(self._libref.rl_chain_set_ppo_loss(self._handle, None if loss is None else C.byref(loss)))

    def trace(self, tile_iteration):
        _lib.check(self._libref.rl_chain_trace(self._handle, int(tile_iteration)))

    def read_trace(self):
        """{"load": [t], "mma": [(t_waited, t_committed)], "epi": [(start, acc ready, regs, math, end)]} in
        SM clock cycles relative to the first stamp (profiling aid)."""
        n = len(self.loads) + 8 * len(self.mmas) + 5 * len(self.epis)
        buf = (C.c_uint64 * n)()
        got = self._libref.rl_chain_read_trace(self._handle, buf, n)
        if got < 0:
            _lib.check(int(got))
        v = list(buf)
        t0 = min(x for x in v if x)
        v = [x - t0 if x else None for x in v]
        nl, nm = len(self.loads), len(self.mmas)
        return {"load": v[:nl], "mma": [tuple(v[nl + 8 * i: nl + 8 * i + 8]) for i in range(nm)],
                "epi": [tuple(v[nl + 8 * nm + 5 * i: nl + 8 * nm + 5 * i + 5]) for i in range(len(self.epis))]}

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and getattr(self, "_libref", None) is not None:
            try:
                self._libref.rl_chain_destroy(h)
            except Exception:
                pass


# =====================================================================================================
# Emulator
# =====================================================================================================
class ChainHazard(AssertionError):
    pass


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


class Emulator:
    """Executes a ChainProgram on CPU tensors with the device's synchronisation semantics."""

    def __init__(self, prog, rows, n_ctas=1, seed=0):
        prog.finalize()
        self.p, self.rows, self.rng = prog, rows, random.Random(seed)
        self.num_tiles = (rows + 127) // 128
        self.n_ctas = n_ctas

    # ---- mbarrier -------------------------------------------------------------------------------------
    class Bar:
        def __init__(self, count):
            self.count, self.pending, self.tx, self.n = count, count, 0, 0

        def _check(self):
            if self.pending == 0 and self.tx == 0:
                self.n += 1
                self.pending = self.count

        def arrive(self):
            if self.pending <= 0:
                raise ChainHazard("mbarrier over-arrival")
            self.pending -= 1
            self._check()

        def arrive_expect(self, nbytes):
            self.tx += nbytes
            self.arrive()

        def complete_tx(self, nbytes):
            self.tx -= nbytes
            self._check()

    def run(self):
        for cta in range(min(self.n_ctas, self.num_tiles)):
            self._run_cta(cta)

    def _run_cta(self, cta):
        p = self.p
        tiles = list(range(cta, self.num_tiles, self.n_ctas))
        bars = [Emulator.Bar(c) for c in p.bar_count]
        flat_smem = torch.zeros(128 * p.n_units, 64)       # unit u = rows [128 u, 128 u + 128)
        units = [flat_smem[128 * u:128 * u + 128] for u in range(p.n_units)]
        nun = lambda off, rows: range(off // UNIT, (off + rows * 128 - 1) // UNIT + 1)
        unit_readers = [0] * p.n_units       # issued, not yet executed async reads (MMA operands, TMA stores)
        unit_loading = [0] * p.n_units       # TMA loads in flight into the unit
        tmem = torch.zeros(128, 512)
        tmem_unread = torch.zeros(512, dtype=torch.bool)   # written by an MMA, not yet loaded by an epilogue op
        mma_queues = [[], []]                # per issuing thread: issued MMAs / commits, executed in that thread's order
        loads_inflight = []
        NW = p.n_workers
        stores_inflight = [[] for _ in range(NW)]           # per epilogue worker: bulk groups complete in order
        store_groups = [{"issued": 0, "read": 0} for _ in range(NW)]
        state = {"done": 0}

        def spec_wait(w, it, role):
            """Generator step: block until the hardware parity test passes; checks phase soundness."""
            if w is None:
                return
            need = w.need + it * p.bar_phases[w.bar]
            parity = (need - 1) & 1
            b = bars[w.bar]
            while True:
                if b.n > need:
                    raise ChainHazard("%s: barrier %s overran: %d completions, waiter needs %d (tile it %d)" %
                                      (role, p.bar_name[w.bar], b.n, need, it))
                if (b.n & 1) != parity:
                    if b.n != need:
                        raise ChainHazard("%s: parity alias on %s" % (role, p.bar_name[w.bar]))
                    return
                if b.n != need - 1:
                    raise ChainHazard("%s: barrier %s is %d phases behind" % (role, p.bar_name[w.bar], need - b.n))
                yield

        def touches(off, nbytes):
            return range(off // UNIT, (off + nbytes - 1) // UNIT + 1)

        # ---------------- roles as generators ----------------
        def load_role(k):
            for it, tile in enumerate(tiles):
                m0 = tile * 128
                for o in p.loads:
                    if o["issuer"] != k:
                        continue
                    yield from spec_wait(o["wait"], it, "load")
                    for u in nun(o["smem_off"], o["bytes"] // 128):
                        if unit_readers[u] or unit_loading[u]:
                            raise ChainHazard("TMA load into unit %d while it is still read / loaded (%s)" % (u, p.name))
                        unit_loading[u] += 1
                    bars[o["full_bar"]].arrive_expect(o["bytes"])
                    loads_inflight.append((o, m0))
                    yield
            state["done"] += 1

        def land(o, m0):
            t, box_rows = p.tensors[o["tensor"]]
            u = o["smem_off"] // UNIT
            r0 = o["row0"] + (m0 if o["tile_rows"] else 0)
            box = torch.zeros(box_rows, 64)
            rr = max(0, min(box_rows, t.shape[0] - r0))
            cc = max(0, min(64, t.shape[1] - o["col0"]))
            if rr > 0 and cc > 0:
                box[:rr, :cc] = t[r0:r0 + rr, o["col0"]:o["col0"] + cc].float()
            for uu in nun(o["smem_off"], box_rows):
                if unit_readers[uu]:
                    raise ChainHazard("TMA load landed in unit %d under a reader" % uu)
                unit_loading[uu] -= 1
            flat_smem[128 * u:128 * u + box_rows] = box
            bars[o["full_bar"]].complete_tx(o["bytes"])

        def mma_role(k):
            for it, tile in enumerate(tiles):
                for o in p.mmas:
                    if o["issuer"] != k:
                        continue
                    for w in o["waits"]:
                        yield from spec_wait(w, it, "mma")
                    for u in [o["a_off"] // UNIT] + list(nun(o["b_off"], o["n"])):
                        if unit_loading[u]:
                            raise ChainHazard("MMA reads unit %d while a TMA load is in flight" % u)
                        unit_readers[u] += 1
                    # (one in-order queue per issuing thread; the tensor pipe interleaves the queues arbitrarily)
                    mma_queues[k].append(("mma", o))
                    for c in o["commits"]:
                        mma_queues[k].append(("commit", c))
                    yield
            state["done"] += 1

        def exec_mma(o):
            ua, ub = o["a_off"] // UNIT, o["b_off"] // UNIT
            n, c0 = o["n"], o["tmem_col"]
            k = 16 * o["k_steps"]
            if not o["accumulate"] and bool(tmem_unread[c0:c0 + n].any()):
                # accumulator names may share columns (trunk_backward): a new accumulation must never start on
                # columns whose previous content the epilogue has not loaded yet
                raise ChainHazard("%s: MMA (tmem col %d, n %d) overwrites accumulator columns the epilogue has not read" %
                                  (p.name, c0, n))
            tmem_unread[c0:c0 + n] = True
            A, B = units[ua][:, :k], flat_smem[128 * ub:128 * ub + n, :k]
            d = A @ B.t()
            if o["accumulate"]:
                tmem[:, c0:c0 + n] += d
            else:
                tmem[:, c0:c0 + n] = d
            for u in [ua] + list(nun(o["b_off"], n)):
                unit_readers[u] -= 1

        def epi_role(wk):
            sg = store_groups[wk]
            for it, tile in enumerate(tiles):
                m0 = tile * 128
                for o in p.epis:
                    if o["worker"] != wk:
                        continue
                    yield from spec_wait(o["wait_acc"], it, "epi")
                    nc = o["ncols"]
                    f = tmem[:, o["tmem_col"]:o["tmem_col"] + nc].clone()
                    tmem_unread[o["tmem_col"]:o["tmem_col"] + (64 if nc > 32 else 32)] = False     # tcgen05.ld width
                    if o["arrive_acc_free"] != NONE:
                        for _ in range(4):
                            bars[o["arrive_acc_free"]].arrive()
                    mode = o["mode"]
                    if mode in (EPI_BIAS_ELU, EPI_BIAS, EPI_BIAS_F32):
                        f = f + p.params[o["bias_off"]:o["bias_off"] + nc].float()
                        if mode == EPI_BIAS_ELU:
                            f = torch.tanh(f) if getattr(p, "activation", "elu") == "tanh" else torch.where(f > 0, f, torch.exp(f) - 1)
                    elif mode == EPI_DELU:
                        yield from spec_wait(o["wait_aux"], it, "epi")
                        ua = o["aux_off"] // UNIT
                        if unit_loading[ua]:
                            raise ChainHazard("epilogue reads aux unit %d while a TMA load is in flight" % ua)
                        y = units[ua][:, :nc]
                        f = f * ((1 - y * y) if getattr(p, "activation", "elu") == "tanh" else torch.where(y > 0, torch.ones_like(y), y + 1))
                    if mode == EPI_BIAS_F32:
                        out = p.outputs[o["out_id"]]
                        rr = max(0, min(128, self.rows - m0))
                        out[m0:m0 + rr, :nc] = f[:rr]
                        continue
                    if o["store_wait_pending"] >= 0:
                        while sg["issued"] - sg["read"] > o["store_wait_pending"]:
                            yield
                    yield from spec_wait(o["wait_dst"], it, "epi")
                    u = o["dst_off"] // UNIT
                    if unit_readers[u] or unit_loading[u]:
                        raise ChainHazard("epilogue overwrites unit %d under a reader (%s, epi op tmem_col %d)" %
                                          (u, p.name, o["tmem_col"]))
                    units[u][:, o["dst_col0"]:o["dst_col0"] + nc] = _bf16(f)
                    if o["arrive_dst_ready"] != NONE:
                        for _ in range(4):
                            bars[o["arrive_dst_ready"]].arrive()
                    if o["release_aux"] != NONE:
                        bars[o["release_aux"]].arrive()
                    if o["store_tensor"] != NONE:
                        unit_readers[u] += 1
                        sg["issued"] += 1
                        stores_inflight[wk].append((o, m0, u))
                        if o["release_after_store"] != NONE:
                            while sg["issued"] != sg["read"]:
                                yield
                            bars[o["release_after_store"]].arrive()
                    yield
            state["done"] += 1

        def do_store(o, m0, u):
            t, _ = p.tensors[o["store_tensor"]]
            rr = max(0, min(128, t.shape[0] - m0))
            cc = max(0, min(64, t.shape[1] - o["store_col0"]))
            if rr > 0 and cc > 0:
                t[m0:m0 + rr, o["store_col0"]:o["store_col0"] + cc] = units[u][:rr, :cc].to(t.dtype)
            unit_readers[u] -= 1
            store_groups[o["worker"]]["read"] += 1

        roles = [load_role(0), load_role(1), mma_role(0), mma_role(1)] + [epi_role(k) for k in range(NW)]
        NR = 4 + NW
        alive = [True] * NR
        blocked_rounds = 0
        # adversarial speeds: every agent (3 roles, TMA loads, tensor pipe, TMA stores) gets its own firing
        # probability per round, re-drawn now and then, so slow-consumer / slow-producer races are exercised
        speeds = [1.0] * (NR + 2 + NW)
        rounds = 0
        while True:
            progressed = False
            if rounds % 400 == 0:
                speeds = [self.rng.choice((0.03, 0.3, 1.0)) for _ in range(NR + 2 + NW)]
            rounds += 1
            order = list(range(NR + 2 + NW))
            self.rng.shuffle(order)
            for a in order:
                if self.rng.random() > speeds[a]:
                    continue
                if a < NR:
                    if not alive[a]:
                        continue
                    before = (len(loads_inflight), len(mma_queues[0]) + len(mma_queues[1]), tuple(len(x) for x in stores_inflight),
                              tuple(b.n for b in bars), tuple(b.pending for b in bars))
                    try:
                        next(roles[a])
                    except StopIteration:
                        alive[a] = False
                        progressed = True
                        continue
                    after = (len(loads_inflight), len(mma_queues[0]) + len(mma_queues[1]), tuple(len(x) for x in stores_inflight),
                             tuple(b.n for b in bars), tuple(b.pending for b in bars))
                    progressed |= before != after
                elif a == NR and loads_inflight:
                    i = self.rng.randrange(len(loads_inflight))      # TMA completes out of order
                    land(*loads_inflight.pop(i))
                    progressed = True
                elif a == NR + 1 and (mma_queues[0] or mma_queues[1]):
                    qs = [q for q in mma_queues if q]
                    kind, x = self.rng.choice(qs).pop(0)              # in order per issuing thread, any interleaving of the two
                    if kind == "mma":
                        exec_mma(x)
                    else:
                        bars[x].arrive()
                    progressed = True
                elif a >= NR + 2 and stores_inflight[a - NR - 2]:
                    do_store(*stores_inflight[a - NR - 2].pop(0))     # bulk groups complete in order
                    progressed = True
            inflight = loads_inflight or mma_queues[0] or mma_queues[1] or any(stores_inflight)
            if not any(alive) and not inflight:
                break
            if progressed:
                blocked_rounds = 0
            else:
                blocked_rounds += 1
                if blocked_rounds > 5000 and not inflight:
                    raise ChainHazard("deadlock in %s: roles alive %s, barrier completions %s" %
                                      (p.name, alive, {p.bar_name[i]: b.n for i, b in enumerate(bars)}))


# =====================================================================================================
# Programs for the learner's networks
# =====================================================================================================
REGIONS = {"C0": (0, 64), "C1": (64, 64), "BIG": (128, 256), "MID": (384, 128)}


def _dense(p, a_boxes, w_tensor, n_out, acc, k_last_steps=4, release_a=True, w_row0=0):
    """acc[:, :n_out] = sum_j A box j * W[w_row0 : w_row0 + n_out, 64 j : 64 j + 64]^T in halves of <= 128 output
    columns (one ring stage per weight box).  Closes the accumulator."""
    nk = len(a_boxes)
    wmax = 128 * p.stage_units
    halves = [(h, min(wmax, n_out - h)) for h in range(0, n_out, wmax)]
    for j, a in enumerate(a_boxes):
        for hi, (h0, hn) in enumerate(halves):
            st = p.load_stage(w_tensor, col0=64 * j, row0=w_row0 + h0)
            p.mma(a, st, n=hn, acc=acc, col_off=h0, k_steps=k_last_steps if j == nk - 1 else 4, accumulate=j > 0,
                  acc_last=(j == nk - 1 and hi == len(halves) - 1), a_release=release_a and hi == len(halves) - 1)


def _boxes(p, acc, width, mode, bias_off, store_tensor=None, store_col0=0, aux_tensor=None, aux_col0=0, has_reader=True):
    """Epilogue of a `width`-column accumulator into width/64 pool boxes (the last may be narrower)."""
    out = []
    n = (width + 63) // 64
    for c in range(n):
        nc = min(64, width - 64 * c)
        aux = None
        if aux_tensor is not None:
            aux = p.load_stage(aux_tensor, col0=aux_col0 + 64 * c, row0=0, tile_rows=True, consumer="epi%d" % p.worker_for(acc, 64 * c))
        out.append(p.epi_box(acc, 64 * c, mode, bias_off=bias_off + 64 * c, ncols=nc, aux=aux,
                             store=None if store_tensor is None else (store_tensor, store_col0 + 64 * c),
                             last=(c == n - 1), has_reader=has_reader))
    return out


def _round16(n):
    return (n + 15) // 16 * 16


REGIONS3 = {"C0": (0, 64), "C1": (64, 64), "C2": (128, 64), "BIG": (192, 256), "MID": (0, 128)}


def teacher_forward_program(T, save=True, trunk=True, want_mean=True, want_value=True, n_stages=3, stage_units=2, n_workers=2,
                            lookahead=None, two_issuers=False):
    """encoder(priv) -> latent merged into the [obs | latent] box -> actor mean / critic value
    (actor_critic.py:124-173 `act` / `evaluate` on one batch; ppo.py:102-107 inside the update).
    T: tensors + parameter offsets (see ActorCritic._chain_tensors).  save: also store every hidden
    activation (the backward needs them).  trunk=False: only refresh the latent slot of Xac.

    n_workers=3 (study for the next kernel revision; not runnable on the two-worker kernel): three 64-column chunk
    accumulators, one per epilogue worker, and the 128-column layer accumulator MID shares the columns of C0 and C1 -
    safe by program order like SB / D2 in the backward (MID is written only after the MMAs that consumed every
    chunk box of the stream, and a stream starts only after the MMAs that consumed MID's boxes)."""
    nw = n_workers
    # two MMA issuers: the second-layer accumulator (8 of a network's 22 ops, all waiting for epilogue boxes) gets its own
    # warp; with two workers the third-layer accumulator joins it (with three it shares columns with the chunk
    # accumulators and stays with them: the overlap is safe by ONE issuer's program order only)
    issuers = ({"BIG": 1, "MID": 1} if nw == 2 else {"BIG": 1}) if two_issuers else {}
    p = ChainProgram(n_pool=6, n_stages=n_stages, n_inputs=2, regions=REGIONS if nw == 2 else REGIONS3, name="teacher_forward",
                     stage_units=stage_units, n_workers=nw, region_worker={"C%d" % k: k for k in range(nw)},
                     region_issuer=issuers)
    p.params = T["params"]
    wbox = lambda w: min(128 * stage_units, _round16(w.shape[0]))      # weight box rows
    tXp, tXac = p.tensor(T["Xp"], 128), p.tensor(T["Xac"], 128)
    tWe1, tWe2 = p.tensor(T["We1"], wbox(T["We1"])), p.tensor(T["We2"], wbox(T["We2"]))
    lat = T["We3"].shape[0]
    tWe3 = p.tensor(T["We3"], _round16(lat))
    st = lambda name: p.tensor(T[name], 128) if save else None
    tH1, tH2 = st("H1"), st("H2")
    xp = p.load_input(0, tXp, 0)
    xac = p.load_input(1, tXac, 0, extra_free_arrivals=1)
    # ---- encoder ----
    kp = T["Xp"].shape[1]
    acc = p.acc("BIG")
    _dense(p, [xp], tWe1, T["We1"].shape[0], acc, k_last_steps=(kp + 15) // 16)
    h1 = _boxes(p, acc, T["We1"].shape[0], EPI_BIAS_ELU, T["b_e1"], tH1)
    acc = p.acc("MID")
    _dense(p, h1, tWe2, T["We2"].shape[0], acc)
    h2 = _boxes(p, acc, T["We2"].shape[0], EPI_BIAS_ELU, T["b_e2"], tH2)
    acc = p.acc("C0")
    _dense(p, h2, tWe3, _round16(lat), acc)
    merged_ready = p._bar(4, "xac.merged")
    xacm = p.epi_merge(acc, 0, lat, EPI_BIAS, T["b_e3"], xac, T["num_obs"], merged_ready, store=(tXac, 0),
                       release_after_store=xac.free_bar)
    if not trunk:
        # nobody multiplies the merged box: release it from the tensor-pipe side with a 1-step dummy?  No -
        # simply give the free barrier its second arrival from the epilogue store alone.
        p.bar_count[xac.free_bar] = 1
        p._signal(xac.free_bar)
        return p.finalize()
    # ---- actor / critic: first layer streamed in 64-column chunks straight into the second layer ----
    tWcat = p.tensor(T["Wcat"], 64)
    tY1 = st("Y1")
    H = T["Wcat"].shape[0] // 2
    nets = []
    if want_mean:
        nets.append(("a", 0, T["Wa2"], T["Wa3"], T["Wa4"], "A2", "A3", T["b_a2"], T["b_a3"], T["b_a4"], 0))
    if want_value:
        nets.append(("c", H, T["Wc2"], T["Wc3"], T["Wc4"], "C2", "C3", T["b_c2"], T["b_c3"], T["b_c4"], 1))
    for ni, (tag, off, W2, W3, W4, n2, n3, b2, b3, b4, out_id) in enumerate(nets):
        tW2, tW3 = p.tensor(W2, wbox(W2)), p.tensor(W3, wbox(W3))
        n_out = W4.shape[0]
        tW4 = p.tensor(W4, _round16(n_out))
        t2, t3 = st(n2), st(n3)
        nch = H // 64
        acc2 = p.acc("BIG")
        chunk_box = [None] * nch

        def l1(j):
            a1 = p.acc("C%d" % (j % nw))
            s = p.load_stage(tWcat, col0=0, row0=off + 64 * j)
            p.mma(xacm, s, n=64, acc=a1, k_steps=4, accumulate=False, acc_last=True,
                  a_release=(ni == len(nets) - 1 and j == nch - 1))
            chunk_box[j] = p.epi_box(a1, 0, EPI_BIAS_ELU, bias_off=T["b_cat"] + off + 64 * j,
                                     store=None if tY1 is None else (tY1, off + 64 * j), last=True)
            if p.stagger_ns and 0 < j < nw:          # the stream's first chunk of worker j: de-phase it
                p.epis[-1]["delay_ns"] = j * p.stagger_ns

        def l2(j):
            wmax = 128 * p.stage_units
            for h in range(0, W2.shape[0], wmax):
                s = p.load_stage(tW2, col0=64 * j, row0=h)
                p.mma(chunk_box[j], s, n=min(wmax, W2.shape[0] - h), acc=acc2, col_off=h, k_steps=4, accumulate=j > 0,
                      acc_last=(j == nch - 1 and h + wmax >= W2.shape[0]), a_release=(h + wmax >= W2.shape[0]))
        # software pipeline: the first-layer chunk `la` ahead is issued BEFORE the second-layer MMA that waits for chunk
        # j's box, so a worker finds its next accumulator full the moment it has written a box (la = number of chunk
        # accumulators: every worker always has one chunk in flight)
        la = nw if lookahead is None else lookahead
        for j in range(min(la, nch)):
            l1(j)
        for j in range(nch):
            if j + la < nch:
                l1(j + la)
            l2(j)
        a2 = _boxes(p, acc2, W2.shape[0], EPI_BIAS_ELU, b2, t2)
        acc3 = p.acc("MID")
        _dense(p, a2, tW3, W3.shape[0], acc3)
        a3 = _boxes(p, acc3, W3.shape[0], EPI_BIAS_ELU, b3, t3)
        # (three workers: MID overlaps C0 / C1, and the head's first MMA waits for the first MID box only - its
        # accumulator must lie outside MID)
        acc4 = p.acc("C%d" % (ni % 2) if nw == 2 else "C2")
        _dense(p, a3, tW4, _round16(n_out), acc4)
        p.output(out_id, T["mean"] if tag == "a" else T["value"])
        p.epi_out(acc4, 0, n_out, b4, out_id)
    return p.finalize()


# Four epilogue workers (16 warps): the wide layers are produced 256 columns at a time - ONE N = 256 MMA op feeds one
# 64-column chunk to each worker - so the tensor-issuing warp runs far fewer ops (its ~800 cycles per op were what the
# two-worker stream waited for) and the SM's four schedulers always have a worker in its MUFU-bound ELU phase while
# another one writes / stores its box.  Tensor-memory map: S [0, 256) (first-layer staging: two uses per network),
# BIG [256, 512) (second-layer accumulator); the narrow accumulators share S's columns: MID [0, 128) (third layer /
# encoder hidden 2), O0 / O1 / O2 [128.., 32 each) (mean, value, latent).  alias_waits=True: the first MMA into a region
# waits until the epilogue has read the latest use of every region it overlaps.
REGIONS4 = {"S": (0, 256), "BIG": (256, 256), "MID": (0, 128), "O0": (128, 32), "O1": (160, 32), "O2": (192, 32)}


def teacher_forward_program4(T, save=True, trunk=True, want_mean=True, want_value=True, n_stages=4):
    """teacher_forward_program for the four-worker kernel (same tensors, same results)."""
    p = ChainProgram(n_pool=4, n_stages=n_stages, n_inputs=2, regions=REGIONS4, name="teacher_forward4", stage_units=2,
                     n_workers=4, region_worker={"O0": 2, "O1": 3, "O2": 2}, alias_waits=True)
    p.params = T["params"]
    wbox = lambda w: min(256, _round16(w.shape[0]))
    tXp, tXac = p.tensor(T["Xp"], 128), p.tensor(T["Xac"], 128)
    tWe1, tWe2 = p.tensor(T["We1"], wbox(T["We1"])), p.tensor(T["We2"], wbox(T["We2"]))
    lat = T["We3"].shape[0]
    tWe3 = p.tensor(T["We3"], _round16(lat))
    st = lambda name: p.tensor(T[name], 128) if save else None
    tH1, tH2 = st("H1"), st("H2")
    assert T["We1"].shape[0] <= 256 and T["We2"].shape[0] <= 128 and _round16(lat) <= 32
    xp = p.load_input(0, tXp, 0)
    xac = p.load_input(1, tXac, 0, extra_free_arrivals=1)
    # ---- encoder ----
    kp = T["Xp"].shape[1]
    # (the previous tile's last accumulator inside S: its output epilogue may still have to read it)
    acc = p.acc("S", carry=(("O2",) if not trunk else ("O1",) if want_value else ("O0",)))
    _dense(p, [xp], tWe1, T["We1"].shape[0], acc, k_last_steps=(kp + 15) // 16)
    h1 = _boxes(p, acc, T["We1"].shape[0], EPI_BIAS_ELU, T["b_e1"], tH1)
    acc = p.acc("MID")
    _dense(p, h1, tWe2, T["We2"].shape[0], acc)
    h2 = _boxes(p, acc, T["We2"].shape[0], EPI_BIAS_ELU, T["b_e2"], tH2)
    acc = p.acc("O2")
    _dense(p, h2, tWe3, _round16(lat), acc)
    merged_ready = p._bar(4, "xac.merged")
    xacm = p.epi_merge(acc, 0, lat, EPI_BIAS, T["b_e3"], xac, T["num_obs"], merged_ready, store=(tXac, 0),
                       release_after_store=xac.free_bar)
    if not trunk:
        p.bar_count[xac.free_bar] = 1
        p._signal(xac.free_bar)
        return p.finalize()
    # ---- actor / critic ----
    tWcat = p.tensor(T["Wcat"], 256)
    tY1 = st("Y1")
    H = T["Wcat"].shape[0] // 2
    assert H % 256 == 0
    nets = []
    if want_mean:
        nets.append(("a", 0, T["Wa2"], T["Wa3"], T["Wa4"], "A2", "A3", T["b_a2"], T["b_a3"], T["b_a4"], 0, "O0"))
    if want_value:
        nets.append(("c", H, T["Wc2"], T["Wc3"], T["Wc4"], "C2", "C3", T["b_c2"], T["b_c3"], T["b_c4"], 1, "O1"))
    for ni, (tag, off, W2, W3, W4, n2, n3, b2, b3, b4, out_id, oreg) in enumerate(nets):
        assert W2.shape[0] <= 256 and W3.shape[0] <= 128
        tW2, tW3 = p.tensor(W2, wbox(W2)), p.tensor(W3, wbox(W3))
        n_out = W4.shape[0]
        tW4 = p.tensor(W4, _round16(n_out))
        t2, t3 = st(n2), st(n3)
        nsc = H // 256
        acc2 = p.acc("BIG")
        boxes = [None] * nsc
        accs = [None] * nsc
        last_net = ni == len(nets) - 1

        def l1(k):
            # the latent's read of O2 is implied by the merged box the op waits for
            a1 = p.acc("S", implied=("O2",))
            s_ = p.load_stage(tWcat, col0=0, row0=off + 256 * k)
            p.mma(xacm, s_, n=256, acc=a1, k_steps=4, accumulate=False, acc_last=True, a_release=(last_net and k == nsc - 1))
            accs[k] = a1

        def l1_epi(k):
            boxes[k] = _boxes(p, accs[k], 256, EPI_BIAS_ELU, T["b_cat"] + off + 256 * k, tY1, store_col0=off + 256 * k)

        def l2(k):
            for b, box in enumerate(boxes[k]):
                j = 4 * k + b
                s_ = p.load_stage(tW2, col0=64 * j, row0=0)
                p.mma(box, s_, n=W2.shape[0], acc=acc2, k_steps=4, accumulate=j > 0, acc_last=(j == H // 64 - 1), a_release=True)
        # software pipeline: super-chunk k + 1's MMA is issued before the second-layer MMAs that wait for the boxes of k
        l1(0)
        l1_epi(0)
        for k in range(1, nsc):
            l1(k)
            l2(k - 1)
            l1_epi(k)
        l2(nsc - 1)
        a2 = _boxes(p, acc2, W2.shape[0], EPI_BIAS_ELU, b2, t2)
        acc3 = p.acc("MID")
        _dense(p, a2, tW3, W3.shape[0], acc3)
        a3 = _boxes(p, acc3, W3.shape[0], EPI_BIAS_ELU, b3, t3)
        acc4 = p.acc(oreg)
        _dense(p, a3, tW4, _round16(n_out), acc4)
        p.output(out_id, T["mean"] if tag == "a" else T["value"])
        p.epi_out(acc4, 0, n_out, b4, out_id)
    return p.finalize()


def trunk_backward_program(T, n_stages=8, two_issuers=False):
    """dgrad half of the backward of actor + critic + encoder (what autograd does behind ppo.py:146-148 for
    the layer inputs): from d(loss)/d(mean) and d(loss)/d(value) down to the encoder's first hidden layer,
    multiplying by ELU' of the saved activations, storing every layer-output gradient for the wgrad GEMMs.
    The two first-layer gradient streams (actor, critic) also accumulate d(latent).

    The wide first-layer gradient (2 x 512 columns) is produced in 128-column super-chunks: four N = 128 MMAs
    each (one per 64-wide k-block of the second layer's gradient) - the MMA warp pays ~900 cycles per op
    whatever its size, so half as many ops as 64-column chunks is what shortens the tile.

    Tensor-memory map (columns): D2 [0, 256) | SA [256, 384) | LAT [384, 416); SB [0, 128) lies INSIDE D2.
    The builder tracks each name separately; the overlap is safe because of the MMA warp's program order:
      * D2 is written again only after the last use of SB has been read - the MMAs that consumed SB's boxes
        (the d(latent) accumulation) waited for those boxes, i.e. for the epilogue's reads of SB, and come
        earlier in the program;
      * SB is first written by super-chunk 1, after super-chunk 0's four MMAs have waited for all four boxes
        of the 256-wide gradient, i.e. for every read of D2.
    The emulator's numeric check under adversarial scheduling covers exactly this (an early overwrite of an
    unread accumulator shows up as a wrong result)."""
    regions = {"D2": (0, 256), "SB": (0, 128), "SA": (256, 128), "LAT": (384, 32)}
    # two issuers per role: the d(latent) accumulation (16 of the 63 MMA ops, each waiting for an epilogue box) gets its
    # own MMA warp - D2 / SB / SA share columns and stay on one issuer - and the 99 loads alternate between two LOAD
    # warps by ring stage (the sensitivity study put half of this program's time on the LOAD warp's critical path)
    p = ChainProgram(n_pool=6, n_stages=n_stages, n_inputs=0, regions=regions, name="trunk_backward",
                     region_worker={"LAT": 0}, region_issuer={"LAT": 1} if two_issuers else None,
                     n_loaders=2 if two_issuers else 1, alias_waits=two_issuers)
    # (alias_waits with two issuers: the program-order argument below ran through the d(latent) MMAs, which now belong
    # to the other issuer - the first MMA into D2 / SB waits explicitly until the other has been read)
    H = T["Wcat_t"].shape[1] // 2
    num_obs = T["num_obs"]
    assert H % 128 == 0
    tY1, tdY1 = p.tensor(T["Y1"], 128), p.tensor(T["dY1"], 128)
    tWcat_t = p.tensor(T["Wcat_t"], 32)
    acc_lat = p.acc("LAT")
    nets = [("a", T["dmean"], 0, "Wa4t", "Wa3t", "Wa2t", "A3", "A2", "dA3", "dA2"),
            ("c", T["dvalue"], H, "Wc4t", "Wc3t", "Wc2t", "C3", "C2", "dC3", "dC2")]
    nsc = H // 128
    lat_ops = {"n": 0, "total": 2 * (H // 64)}
    for ni, (tag, d_out, off, w4, w3, w2, s3, s2, g3, g2) in enumerate(nets):
        tW4, tW3, tW2 = p.tensor(T[w4], 128), p.tensor(T[w3], 128), p.tensor(T[w2], 128)
        tS3, tS2, tG3, tG2 = p.tensor(T[s3], 128), p.tensor(T[s2], 128), p.tensor(T[g3], 128), p.tensor(T[g2], 128)
        n3, n2 = T[w4].shape[0], T[w3].shape[0]
        assert n3 <= 128 and n2 <= 256 and n2 % 64 == 0
        d_in = p.load_stage(p.tensor(d_out, 128), col0=0, row0=0, tile_rows=True)
        acc = p.acc("SA")
        _dense(p, [d_in], tW4, n3, acc, k_last_steps=(d_out.shape[1] + 15) // 16)
        d3 = _boxes(p, acc, n3, EPI_DELU, 0, tG3, aux_tensor=tS3)
        acc = p.acc("D2", carry=("SB",) if ni == 0 else ())       # (the previous tile's last super-chunk)
        _dense(p, d3, tW3, n2, acc)
        d2 = _boxes(p, acc, n2, EPI_DELU, 0, tG2, aux_tensor=tS2)
        accs, boxes = [None] * nsc, [None] * nsc

        def sc_mma(k):
            a1 = p.acc("SA" if k % 2 == 0 else "SB")
            for j, a in enumerate(d2):
                s = p.load_stage(tW2, col0=64 * j, row0=128 * k)
                p.mma(a, s, n=128, acc=a1, k_steps=4, accumulate=j > 0, acc_last=(j == len(d2) - 1), a_release=(k == nsc - 1))
            accs[k] = a1

        def sc_epi(k):
            boxes[k] = _boxes(p, accs[k], 128, EPI_DELU, 0, tdY1, store_col0=off + 128 * k, aux_tensor=tY1, aux_col0=off + 128 * k)

        def lat(k):
            for b, box in enumerate(boxes[k]):
                s = p.load_stage(tWcat_t, col0=off + 128 * k + 64 * b, row0=num_obs)
                lat_ops["n"] += 1
                p.mma(box, s, n=32, acc=acc_lat, k_steps=4, accumulate=lat_ops["n"] > 1,
                      acc_last=(lat_ops["n"] == lat_ops["total"]), a_release=True)
        # software pipeline: the MMAs of super-chunk k run while the epilogue workers finish super-chunk k - 1;
        # the d(latent) MMAs of k - 1 come next (they release the two pool units super-chunk k's boxes go to)
        sc_mma(0)
        sc_epi(0)
        for k in range(1, nsc):
            sc_mma(k)
            lat(k - 1)
            sc_epi(k)
        lat(nsc - 1)
    # ---- encoder ----
    tdLat = p.tensor(T["dLat"], 128)
    dlat = _boxes(p, acc_lat, 32, EPI_PLAIN, 0, tdLat)
    tWe3t, tWe2t = p.tensor(T["We3t"], 128), p.tensor(T["We2t"], 128)
    tH2, tH1, tdH2, tdH1 = p.tensor(T["H2"], 128), p.tensor(T["H1"], 128), p.tensor(T["dH2"], 128), p.tensor(T["dH1"], 128)
    assert T["We3t"].shape[0] <= 128 and T["We2t"].shape[0] <= 256
    acc = p.acc("SA")
    _dense(p, dlat, tWe3t, T["We3t"].shape[0], acc, k_last_steps=2)
    dh2 = _boxes(p, acc, T["We3t"].shape[0], EPI_DELU, 0, tdH2, aux_tensor=tH2)
    acc = p.acc("D2")
    _dense(p, dh2, tWe2t, T["We2t"].shape[0], acc)
    _boxes(p, acc, T["We2t"].shape[0], EPI_DELU, 0, tdH1, aux_tensor=tH1, has_reader=False)
    return p.finalize()


def adaptation_forward_program(T, save=True, a_depth=4):
    """adaptation_module(obs_history) (actor_critic.py:158-162; ppo.py:157): the 630-wide input streams
    through its own ring next to the first layer's weights, the two small layers stay on chip.

    Two rings: the input boxes come from HBM (~2 us latency, 158 KB per tile) and need `a_depth` boxes in flight;
    the weight boxes come from L2 and two 32 KB stages are enough.  The LOAD role issues in program order, so the
    input loads are emitted `a_depth` k-blocks ahead of the weight loads - a weight load waiting for its stage
    never holds back an input load that could already be in flight."""
    n_w = (14 - 6 - a_depth) // 2
    p = ChainProgram(n_pool=6, n_stages=0, n_inputs=0, regions=REGIONS, name="adaptation_forward",
                     rings=[("w", n_w, 2), ("x", a_depth, 1)])
    p.params = T["params"]
    tXh, tWd1 = p.tensor(T["Xh"], 128), p.tensor(T["Wd1"], min(256, _round16(T["Wd1"].shape[0])))
    n1, n2, n3 = T["Wd1"].shape[0], T["Wd2"].shape[0], T["Wd3"].shape[0]
    assert n1 <= 256
    tWd2, tWd3 = p.tensor(T["Wd2"], _round16(n2)), p.tensor(T["Wd3"], _round16(n3))
    st = lambda name: p.tensor(T[name], 128) if save else None
    tD1, tD2 = st("D1"), st("D2")
    K = T["Xh"].shape[1]
    nk = (K + 63) // 64
    acc = p.acc("BIG")
    xs = {}

    def load_x(j):
        if j < nk:
            xs[j] = p.load_stage(tXh, col0=64 * j, row0=0, tile_rows=True, ring="x")
    ws = {}

    def load_w(j):
        if j < nk:
            ws[j] = p.load_stage(tWd1, col0=64 * j, row0=0, ring="w")
    for j in range(a_depth):
        load_x(j)
    for j in range(n_w):
        load_w(j)
    for j in range(nk):
        p.mma(xs.pop(j), ws.pop(j), n=n1, acc=acc, k_steps=min(4, (K - 64 * j + 15) // 16), accumulate=j > 0,
              acc_last=(j == nk - 1), a_release=True)
        load_x(j + a_depth)          # both wait for MMA j (it frees their stages): neither holds the other back
        load_w(j + n_w)
    d1 = _boxes(p, acc, n1, EPI_BIAS_ELU, T["b_d1"], tD1)
    acc = p.acc("C0")
    _dense(p, d1, tWd2, _round16(n2), acc)
    d2 = _boxes(p, acc, n2, EPI_BIAS_ELU, T["b_d2"], tD2)
    acc = p.acc("C1")
    _dense(p, d2, tWd3, _round16(n3), acc, k_last_steps=(n2 + 15) // 16)
    p.output(0, T["pred"])
    p.epi_out(acc, 0, n3, T["b_d3"], 0)
    return p.finalize()


def adaptation_backward_program(T):
    """dgrad of the adaptation-module regression (ppo.py:164-166): d(pred) -> d(D2) -> d(D1)."""
    p = ChainProgram(n_pool=6, n_stages=6, n_inputs=1, regions=REGIONS, name="adaptation_backward")
    dp = p.load_input(0, p.tensor(T["dpred"], 128), 0)
    n2, n1 = T["Wd3t"].shape[0], T["Wd2t"].shape[0]
    tW3, tW2 = p.tensor(T["Wd3t"], _round16(n2)), p.tensor(T["Wd2t"], 128)
    tD2, tD1, tdD2, tdD1 = p.tensor(T["D2"], 128), p.tensor(T["D1"], 128), p.tensor(T["dD2"], 128), p.tensor(T["dD1"], 128)
    acc = p.acc("C0")
    _dense(p, [dp], tW3, _round16(n2), acc, k_last_steps=(T["dpred"].shape[1] + 15) // 16)
    dd2 = _boxes(p, acc, n2, EPI_DELU, 0, tdD2, aux_tensor=tD2)
    acc = p.acc("BIG")
    _dense(p, dd2, tW2, n1, acc, k_last_steps=(n2 + 15) // 16)
    _boxes(p, acc, n1, EPI_DELU, 0, tdD1, aux_tensor=tD1, has_reader=False)
    return p.finalize()


# =====================================================================================================
# Which programs the learner runs
# =====================================================================================================
def workers():
    """Epilogue workers of the teacher-forward chain program: RL_CHAIN_WORKERS = 2 | 3 | 4.  Measured at 196608 rows
    (profiles/r02_chain_ncu.md): 2 workers 438 us (423 with the first-layer chunks issued two ahead), 3 workers with
    three chunks ahead 411 us, 4 workers (teacher_forward_program4; 96 registers per thread) 549 us."""
    import os
    n = int(os.environ.get("RL_CHAIN_WORKERS", "3"))
    assert n in (2, 3, 4), "RL_CHAIN_WORKERS must be 2, 3 or 4"
    return n


def issuers():
    """MMA / LOAD issuing warps the learner's programs use: RL_CHAIN_ISSUERS = 1 | 2."""
    import os
    n = int(os.environ.get("RL_CHAIN_ISSUERS", "2"))
    assert n in (1, 2)
    return n


def teacher_forward(T, **kw):
    import os
    if workers() == 4:
        return teacher_forward_program4(T, **kw)
    la = os.environ.get("RL_CHAIN_LOOKAHEAD")
    return teacher_forward_program(T, n_workers=workers(), lookahead=int(la) if la else None, two_issuers=issuers() == 2, **kw)


def trunk_backward(T):
    return trunk_backward_program(T, two_issuers=issuers() == 2)

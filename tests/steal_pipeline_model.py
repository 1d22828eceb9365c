"""Discrete model of the intra-CTA role choreography of conv_steal_kernel (defer_b200/csrc/conv_umma.cu).

Every role of the kernel (claim/TMA producer, MMA issuer, store/publish thread, 8 epilogue warps) is a generator that
performs the SAME sequence of mbarrier waits / arrives, with the same parities and arrival counts, as the CUDA code;
asynchronous completions (TMA landing, tcgen05.commit, bulk-store completion) are separate events.  A random scheduler
interleaves them.  The model detects
  * deadlock (no runnable role / event while work remains),
  * a resource overwritten while still in use: operand-ring stage, TMEM accumulator, residual tile, staging tile,
    descriptor FIFO slot,
  * an mbarrier over-arrival (more arrivals than the barrier's count in one phase).
mbarrier semantics: `count` arrivals complete the current phase and flip the phase bit; wait(parity) succeeds iff the
phase bit differs from `parity` (the phase with that parity has completed) - exactly try_wait.parity."""
from __future__ import annotations

import random

FIFO = 2          # STEAL_FIFO of the kernel (the tests also run 4)
EPI_WARPS = 8


class MBar:
    def __init__(self, count, name):
        self.count, self.pending, self.bit, self.name = count, count, 0, name

    def arrive(self):
        assert self.pending > 0, f"over-arrival on {self.name}"
        self.pending -= 1
        if self.pending == 0:
            self.pending = self.count
            self.bit ^= 1

    def passed(self, parity):
        return self.bit != parity


class Violation(Exception):
    pass


class CTA:
    """State of one CTA processing `tiles` = [(k_blocks, has_res, direct), ...] in claim order."""

    def __init__(self, tiles, stages, rng):
        self.tiles, self.S, self.rng = tiles, stages, rng
        B = lambda c, n: MBar(c, n)
        self.full = [B(1, f"full{s}") for s in range(stages)]
        self.empty = [B(1, f"empty{s}") for s in range(stages)]
        self.tfull = [B(1, f"tfull{b}") for b in range(2)]
        self.tempty = [B(EPI_WARPS, f"tempty{b}") for b in range(2)]
        self.rfull = [B(1, f"rfull{b}") for b in range(2)]
        self.rfree = [B(EPI_WARPS, f"rfree{b}") for b in range(2)]
        self.ofull = [B(EPI_WARPS, f"ofull{b}") for b in range(2)]
        self.ofree = [B(1, f"ofree{b}") for b in range(2)]
        self.qfull = [B(1, f"qfull{q}") for q in range(FIFO)]
        self.qempty = [B(2 + EPI_WARPS, f"qempty{q}") for q in range(FIFO)]
        # resource ownership bookkeeping (what the barriers are supposed to protect)
        self.ring_owner = [None] * stages        # None = free, ('tma', it, kb) in flight / landed, until MMA done
        self.acc_readers = [0, 0]                # epilogue warps that still have to read accumulator b
        self.acc_busy = [False, False]           # MMA writing accumulator b
        self.rbuf_readers = [0, 0]
        self.obuf_state = ["free", "free"]       # free -> writing (warps) -> storing -> free
        self.obuf_writers = [0, 0]
        self.fifo_readers = [0] * FIFO
        self.events = []                         # pending asynchronous completions: callables
        self.published = 0
        self.done_tiles = 0

    def later(self, fn):
        self.events.append(fn)

    # ------------------------------------------------------------------ roles (generators yield while blocked)
    def producer(self):
        stage, phase, rit = 0, 0, 0
        for it, (kbs, has_res, direct) in enumerate(self.tiles):
            q, u = it % FIFO, it // FIFO
            while not self.qempty[q].passed((u & 1) ^ 1):      # slot first, then the claim (kernel order)
                yield
            yield                                              # claim attempt(s)
            if self.fifo_readers[q] != 0:
                raise Violation(f"FIFO slot {q} overwritten with {self.fifo_readers[q]} readers left (tile {it})")
            self.fifo_readers[q] = 2 + EPI_WARPS
            self.qfull[q].arrive()
            if has_res and not direct:
                rb, ur = rit & 1, rit >> 1
                rit += 1
                while not self.rfree[rb].passed((ur & 1) ^ 1):
                    yield
                if self.rbuf_readers[rb] != 0:
                    raise Violation(f"residual tile {rb} overwritten while {self.rbuf_readers[rb]} warps still read it")
                self.rbuf_readers[rb] = EPI_WARPS

                def land(rb=rb):
                    self.rfull[rb].arrive()
                self.later(land)
            for kb in range(kbs):
                while not self.empty[stage].passed(phase ^ 1):
                    yield
                if self.ring_owner[stage] is not None:
                    raise Violation(f"ring stage {stage} overwritten while owned by {self.ring_owner[stage]}")
                self.ring_owner[stage] = ("tma", it, kb)

                def land(stage=stage):
                    self.full[stage].arrive()
                self.later(land)
                stage += 1
                if stage == self.S:
                    stage, phase = 0, phase ^ 1
                yield
        it = len(self.tiles)
        q, u = it % FIFO, it // FIFO
        while not self.qempty[q].passed((u & 1) ^ 1):
            yield
        self.fifo_readers[q] = 2 + EPI_WARPS
        self.qfull[q].arrive()

    def _read_desc(self, it):
        q, u = it % FIFO, it // FIFO
        while not self.qfull[q].passed(u & 1):
            yield False
        if self.fifo_readers[q] <= 0:
            raise Violation(f"FIFO slot {q} read twice")
        self.fifo_readers[q] -= 1
        self.qempty[q].arrive()
        yield True

    def mma(self):
        stage, phase = 0, 0
        it = 0
        while True:
            for ok in self._read_desc(it):
                if not ok:
                    yield
            if it == len(self.tiles):
                return
            kbs = self.tiles[it][0]
            buf, aphase = it & 1, (it >> 1) & 1
            while not self.tempty[buf].passed(aphase ^ 1):
                yield
            if self.acc_readers[buf] != 0:
                raise Violation(f"accumulator {buf} overwritten with {self.acc_readers[buf]} warps still to read (tile {it})")
            self.acc_busy[buf] = True
            for kb in range(kbs):
                while not self.full[stage].passed(phase):
                    yield
                if self.ring_owner[stage] != ("tma", it, kb):
                    raise Violation(f"MMA of tile {it} kb {kb} found ring stage {stage} = {self.ring_owner[stage]}")

                def consumed(stage=stage):
                    self.ring_owner[stage] = None
                    self.empty[stage].arrive()
                self.later(consumed)
                stage += 1
                if stage == self.S:
                    stage, phase = 0, phase ^ 1
                yield

            def acc_done(buf=buf):
                self.acc_busy[buf] = False
                self.acc_readers[buf] = EPI_WARPS
                self.tfull[buf].arrive()
            self.later(acc_done)
            it += 1

    def store(self):
        it = 0
        while True:
            for ok in self._read_desc(it):
                if not ok:
                    yield
            if it == len(self.tiles):
                return
            direct = self.tiles[it][2]
            buf, ophase = it & 1, (it >> 1) & 1
            while not self.ofull[buf].passed(ophase):
                yield
            if not direct:
                if self.obuf_state[buf] != "written":
                    raise Violation(f"store of tile {it} found staging {buf} in state {self.obuf_state[buf]}")
                self.obuf_state[buf] = "storing"
                yield                                      # bulk store in flight, wait_group 0
                self.obuf_state[buf] = "free"
            self.ofree[buf].arrive()
            self.done_tiles += 1
            it += 1

    def epilogue(self, w):
        rit = 0
        it = 0
        while True:
            for ok in self._read_desc(it):
                if not ok:
                    yield
            if it == len(self.tiles):
                return
            _, has_res, direct = self.tiles[it]
            buf, aphase = it & 1, (it >> 1) & 1
            if has_res and not direct:
                rb, ur = rit & 1, rit >> 1
                rit += 1
                while not self.rfull[rb].passed(ur & 1):
                    yield
                if self.rbuf_readers[rb] <= 0:
                    raise Violation(f"warp {w}: residual tile {rb} read without a pending load")
                self.rbuf_readers[rb] -= 1
                self.rfree[rb].arrive()
            while not self.tfull[buf].passed(aphase):
                yield
            if self.acc_busy[buf] or self.acc_readers[buf] <= 0:
                raise Violation(f"warp {w}: accumulator {buf} read while busy / not ready (tile {it})")
            self.acc_readers[buf] -= 1
            self.tempty[buf].arrive()
            yield                                          # epilogue math
            if not direct:
                uo = it >> 1
                while not self.ofree[buf].passed((uo & 1) ^ 1):
                    yield
                if self.obuf_state[buf] not in ("free", "writing"):
                    raise Violation(f"warp {w}: staging {buf} written while {self.obuf_state[buf]} (tile {it})")
                self.obuf_state[buf] = "writing"
                self.obuf_writers[buf] += 1
                if self.obuf_writers[buf] == EPI_WARPS:
                    self.obuf_writers[buf] = 0
                    self.obuf_state[buf] = "written"
            self.ofull[buf].arrive()
            it += 1


def simulate(tiles, stages, seed, max_steps=2_000_000):
    rng = random.Random(seed)
    cta = CTA(tiles, stages, rng)
    roles = {"producer": cta.producer(), "mma": cta.mma(), "store": cta.store()}
    for w in range(EPI_WARPS):
        roles[f"epi{w}"] = cta.epilogue(w)
    idle_rounds = 0
    for _ in range(max_steps):
        if not roles and not cta.events:
            break
        progressed = False
        names = list(roles)
        rng.shuffle(names)
        # a random subset of roles takes one step each; asynchronous events fire at random
        for n in names:
            if rng.random() < 0.7:
                snapshot = _state(cta)
                try:
                    next(roles[n])
                except StopIteration:
                    del roles[n]
                    progressed = True
                    continue
                if _state(cta) != snapshot:
                    progressed = True
        if cta.events and rng.random() < 0.8:
            cta.events.pop(rng.randrange(len(cta.events)))()
            progressed = True
        idle_rounds = 0 if progressed else idle_rounds + 1
        if idle_rounds > 2000:
            raise Violation(f"deadlock: roles left {sorted(roles)}, events {len(cta.events)}, tiles stored {cta.done_tiles}/{len(tiles)}")
    else:
        raise Violation("step budget exhausted")
    if cta.done_tiles != len(tiles):
        raise Violation(f"only {cta.done_tiles} of {len(tiles)} tiles stored")
    return cta


def _state(c):
    bars = c.full + c.empty + c.tfull + c.tempty + c.rfull + c.rfree + c.ofull + c.ofree + c.qfull + c.qempty
    return (tuple((b.pending, b.bit) for b in bars), tuple(c.ring_owner), tuple(c.acc_readers), tuple(c.rbuf_readers),
            tuple(c.obuf_state), tuple(c.obuf_writers), tuple(c.fifo_readers), len(c.events), c.done_tiles)

"""CPU simulation of the mbarrier protocols of csrc/attention_tc.cu (the experimental tcgen05 attention).

The kernels have not run on hardware yet; what CAN be checked here is the synchronisation design: the warp roles
(TMA producer, MMA issuer, element-wise warps), the asynchronous engines (TMA; the tensor pipe, which completes
tcgen05.mma / tcgen05.commit in issue order) and every mbarrier with its arrival count and the phase parity each
wait uses are restated below as cooperating generators and run under a random scheduler.  Every buffer (shared
memory tiles, TMEM column ranges) is a `Resource` that asserts, at the moment an access EXECUTES, that a read sees
the item it expects and that a write does not clobber data somebody still has to read.  A wrong parity, a missing
barrier or a wrong arrival count shows up as a hazard assertion or as a deadlock, for any interleaving tried.
The restatement mirrors the kernels statement by statement (same barrier names)."""
import random

import pytest


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier's count in one phase"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):          # mbarrier.try_wait.parity: has the phase with this parity completed?
        return (self.phase & 1) != parity


class Resource:
    """A buffer written by `writers` parties and then read by `readers` parties, per item."""
    def __init__(self, name, writers, readers):
        self.name, self.nw, self.nr = name, writers, readers
        self.tag, self.w_left, self.r_left = None, 0, 0

    def write(self, tag):
        if self.w_left == 0:         # first writer of a new item
            assert self.r_left == 0, f"{self.name}: item {tag} overwrites item {self.tag} with {self.r_left} reads pending"
            self.tag, self.w_left, self.r_left = tag, self.nw, self.nr
        assert self.tag == tag, f"{self.name}: writers of items {self.tag} and {tag} interleave"
        self.w_left -= 1

    def read(self, tag):
        assert self.tag == tag and self.w_left == 0, f"{self.name}: read for item {tag} sees item {self.tag} ({self.w_left} writes pending)"
        assert self.r_left > 0, f"{self.name}: more reads of item {tag} than declared"
        self.r_left -= 1


class Engine:
    """In-order asynchronous engine (tensor pipe / TMA): queued closures execute later, one at a time."""
    def __init__(self):
        self.q = []

    def push(self, fn):
        self.q.append(fn)

    def actor(self, rng):
        while True:
            if self.q and rng.random() < 0.5:
                self.q.pop(0)()
            yield None


def wait(bar, parity):
    while not bar.done(parity):
        yield ("blocked", bar)
    yield None


def run(actors, engines, rng, max_ticks=2_000_000):
    live = list(actors)
    eng = [e.actor(rng) for e in engines]
    for _ in range(max_ticks):
        if not live:
            for e in engines:                     # drain what is still queued (commits after the last item)
                while e.q:
                    e.q.pop(0)()
            return
        for e in eng:
            next(e)
        a = rng.choice(live)
        try:
            next(a)
        except StopIteration:
            live.remove(a)
    raise AssertionError("deadlock or livelock: actors still blocked after max_ticks")


def active_warps(S, t):
    rows = S - 128 * t
    return 0 if rows <= 0 else (4 if rows >= 128 else (rows + 31) // 32)


# ------------------------------------------------------------------------------------------- forward
def forward_protocol(S, items, rng):
    nt = 2 if S > 128 else 1
    na = [active_warps(S, t) for t in range(2)]
    full = [MBar(2), MBar(2)]
    empty = [MBar(1), MBar(1)]
    sfull = [MBar(1), MBar(1)]
    pfull = [MBar(max(na[t], 1)) for t in range(2)]
    ofull = [MBar(1), MBar(1)]
    oempty = [MBar(max(na[t], 1)) for t in range(2)]
    tma, pipe = Engine(), Engine()
    qkv = [Resource(f"QKV[{b}]", 1, 2 * nt) for b in range(2)]
    mbias = [Resource(f"mbias[{b}]", 1, sum(na[:nt])) for b in range(2)]
    s_t = [Resource(f"S[{t}]", 1, na[t]) for t in range(2)]
    p_t = [Resource(f"P[{t}]", na[t], 1) for t in range(2)]
    o_t = [Resource(f"O[{t}]", 1, na[t]) for t in range(2)]
    done = []

    def producer():
        for it in range(items):
            buf, ph = it & 1, (it >> 1) & 1
            yield from wait(empty[buf], ph ^ 1)
            tma.push(lambda buf=buf, it=it: (qkv[buf].write(it), full[buf].arrive()))   # expect_tx + complete_tx
            mbias[buf].write(it)
            yield None
            full[buf].arrive()
            yield None

    def mma():
        for it in range(items):
            buf, ph, itp = it & 1, (it >> 1) & 1, it & 1
            yield from wait(full[buf], ph)
            for t in range(nt):
                pipe.push(lambda buf=buf, t=t, it=it: (qkv[buf].read(it), s_t[t].write(it)))
                pipe.push(lambda t=t: sfull[t].arrive())
                yield None
            for t in range(nt):
                yield from wait(pfull[t], itp)
                yield from wait(oempty[t], itp ^ 1)
                pipe.push(lambda buf=buf, t=t, it=it: (p_t[t].read(it), qkv[buf].read(it), o_t[t].write(it)))
                pipe.push(lambda t=t: ofull[t].arrive())
                yield None
            pipe.push(lambda buf=buf: empty[buf].arrive())

    def softmax(t, q):
        for it in range(items):
            buf, ph, itp = it & 1, (it >> 1) & 1, it & 1
            yield from wait(full[buf], ph)
            yield from wait(sfull[t], itp)
            mbias[buf].read(it)
            s_t[t].read(it)
            yield None
            p_t[t].write(it)
            yield None
            pfull[t].arrive()
            yield from wait(ofull[t], itp)
            o_t[t].read(it)
            yield None
            oempty[t].arrive()
            done.append((it, t, q))

    actors = [producer(), mma()] + [softmax(t, q) for t in range(nt) for q in range(na[t])]
    run(actors, [tma, pipe], rng)
    assert len(done) == items * sum(na[:nt])


# ------------------------------------------------------------------------------------------- backward
def backward_protocol(S, items, rng, nsplit=2):
    nu = 2 if S > 128 else 1
    na = [nsplit * active_warps(S, u) for u in range(2)]     # nsplit column parts per active lane quarter
    ld_full, ld_empty = MBar(1), MBar(1)
    sd_full = [MBar(1), MBar(1)]
    pds_full = [MBar(max(na[u], 1)) for u in range(2)]
    kv_full = [MBar(1), MBar(1)]
    s_empty = [MBar(max(na[u], 1)) for u in range(2)]
    dq_full, dq_empty = MBar(1), MBar(na[0])
    tma, pipe = Engine(), Engine()
    tiles = Resource("Q/K/V/dO", 1, 3 * nu)                   # read by the S/dP, the dV/dK and the dQ MMAs of each key tile
    sdp = Resource("S^T,dP^T", 1, 0)                          # readers set per key tile below
    pds = Resource("Pd^T,dS^T", 0, 2)                         # dV/dK MMAs and dQ MMAs
    dvk = Resource("dV,dK", 1, 0)
    dq = Resource("dQ", nu, na[0])
    vecs = [Resource(f"lse2/delta[{b}]", 4 * nsplit, sum(na[:nu])) for b in range(2)]
    group = {"count": 0, "gen": 0}                            # bar.sync 1, 128 * nsplit among the element-wise warps
    done = []

    def producer():
        for it in range(items):
            yield from wait(ld_empty, (it & 1) ^ 1)
            tma.push(lambda it=it: (tiles.write(it), ld_full.arrive()))
            yield None

    def mma():
        for it in range(items):
            itp = it & 1
            yield from wait(ld_full, itp)
            for u in range(nu):
                if u == 0:
                    yield from wait(s_empty[nu - 1], itp ^ 1)
                else:
                    yield from wait(s_empty[u - 1], itp)

                def s_op(it=it, u=u):
                    tiles.read(it)
                    sdp.nr = na[u]
                    sdp.write((it, u))
                    # the dV / dK accumulators alias these TMEM columns: they must have been read out
                    assert dvk.r_left == 0, "S^T / dP^T overwrite dV / dK that are still being read"
                pipe.push(s_op)
                pipe.push(lambda u=u: sd_full[u].arrive())
                yield None
                yield from wait(pds_full[u], itp)
                if u == 0:
                    yield from wait(dq_empty, itp ^ 1)

                def kv_op(it=it, u=u):
                    pds.read((it, u)); tiles.read(it)
                    assert sdp.r_left == 0, "dV / dK overwrite S^T / dP^T columns that are still being read"
                    dvk.nr = na[u]
                    dvk.write((it, u))

                def dq_op(it=it, u=u):
                    pds.read((it, u)); tiles.read(it)
                    dq.write(it)
                pipe.push(kv_op)
                pipe.push(dq_op)
                pipe.push(lambda u=u: kv_full[u].arrive())
                yield None
            pipe.push(ld_empty.arrive)
            pipe.push(dq_full.arrive)

    def elementwise(q, half):
        for it in range(items):
            itp = it & 1
            vecs[it & 1].write(it)                            # this warp's share of lse2 / delta
            yield None
            gen = group["gen"]                                # bar.sync 1, 256
            group["count"] += 1
            if group["count"] == 4 * nsplit:
                group["count"] = 0
                group["gen"] += 1
            while group["gen"] == gen:
                yield None
            for u in range(nu):
                if u * 128 + q * 32 >= S:
                    continue
                yield from wait(sd_full[u], itp)
                vecs[it & 1].read(it)
                sdp.read((it, u))
                yield None
                pds.nw = na[u]
                pds.write((it, u))
                yield None
                pds_full[u].arrive()
                yield from wait(kv_full[u], itp)
                dvk.read((it, u))
                yield None
                s_empty[u].arrive()
            if q * 32 < S:
                yield from wait(dq_full, itp)
                dq.read(it)
                yield None
                dq_empty.arrive()
            done.append((it, q, half))

    actors = [producer(), mma()] + [elementwise(q, h) for q in range(4) for h in range(nsplit)]
    run(actors, [tma, pipe], rng)
    assert len(done) == items * 4 * nsplit


@pytest.mark.parametrize("S", [160, 145, 129, 128, 100, 97, 96, 76, 65, 33, 16, 1])
@pytest.mark.parametrize("seed", range(6))
def test_forward_protocol(S, seed):
    forward_protocol(S, items=7, rng=random.Random(1000 * S + seed))


@pytest.mark.parametrize("S", [160, 145, 129, 128, 100, 97, 96, 76, 65, 33, 16, 1])
@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("nsplit", [2, 4])
def test_backward_protocol(S, seed, nsplit):
    backward_protocol(S, items=7, rng=random.Random(1000 * S + seed), nsplit=nsplit)


def test_simulator_catches_a_missing_wait():
    """Sanity of the method: drop one wait from a copy of the forward protocol and the hazard is reported."""
    full, sfull, pfull = MBar(1), MBar(1), MBar(1)
    pipe = Engine()
    s = Resource("S", 1, 1)

    def mma():
        for it in range(4):
            # BUG under test: no wait on pfull (S columns still being read by the softmax warp)
            pipe.push(lambda it=it: s.write(it))
            pipe.push(sfull.arrive)
            yield None

    def softmax():
        for it in range(4):
            yield from wait(sfull, it & 1)
            yield None
            s.read(it)
            pfull.arrive()

    with pytest.raises(AssertionError):
        for seed in range(50):
            full.phase = 0
            run([mma(), softmax()], [pipe], random.Random(seed), max_ticks=10_000)

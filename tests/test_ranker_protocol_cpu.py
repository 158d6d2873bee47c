"""The barrier protocol of the fused ranker kernel (csrc/ranker_fused.cu: compute groups, TMA loader with two chunk rings,
forward and gradient MMA issuers), replayed on the CPU under random schedules.  mbarrier waits test a phase PARITY, so a
waiter that falls two phases behind (or looks two ahead) aliases: this model reproduces that rule and caught the
hand-over of the shared dZ tiles going wrong before the per-group `dzmine` barrier existed."""
import random

import pytest


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):          # mbarrier.try_wait.parity: true iff the current phase's parity differs
        return (self.phase & 1) != parity


def run(G, NSF, NSD, NDZ, NCH, nsl, seed, own_barrier=True, max_steps=400000):
    rnd = random.Random(seed)
    full_f, empty_f = [Bar(1) for _ in range(NSF)], [Bar(1) for _ in range(NSF)]
    full_d, empty_d = [Bar(1) for _ in range(NSD)], [Bar(1) for _ in range(NSD)]
    zfull, zfree, dzmine = [Bar(1) for _ in range(G)], [Bar(1) for _ in range(G)], [Bar(1) for _ in range(G)]
    dzfull, dzfree = [Bar(1) for _ in range(NDZ)], [Bar(1) for _ in range(NDZ)]
    pend, order = [], []

    def wait(b, par):
        while not b.done(par):
            yield

    def group(g):
        it = 0
        for k in range(g, nsl, G):
            yield from wait(zfull[g], it & 1)
            for _ in range(rnd.randint(1, 30)):                   # the loss
                yield
            tile, tuse = k % NDZ, k // NDZ
            if own_barrier and it >= 1:
                yield from wait(dzmine[g], (it - 1) & 1)
            if tuse >= 1:
                yield from wait(dzfree[tile], (tuse - 1) & 1)
            dzfull[tile].arrive()
            zfree[g].arrive()
            it += 1

    def loader():
        total = nsl * NCH
        fn = fk = fc = fslot = fuse = 0
        dn = dk = dc = dslot = duse = 0
        while fn < total or dn < total:
            if fn < total and (fuse == 0 or empty_f[fslot].done((fuse - 1) & 1)):
                pend.append([rnd.randint(1, 10), full_f[fslot]])
                fn, fc, fslot = fn + 1, fc + 1, fslot + 1
                if fc == NCH:
                    fc, fk = 0, fk + 1
                if fslot == NSF:
                    fslot, fuse = 0, fuse + 1
            if dn < total and (fk > dk or fn == total) and (duse == 0 or empty_d[dslot].done((duse - 1) & 1)):
                pend.append([rnd.randint(1, 5), full_d[dslot]])
                dn, dc, dslot = dn + 1, dc + 1, dslot + 1
                if dc == NCH:
                    dc, dk = 0, dk + 1
                if dslot == NSD:
                    dslot, duse = 0, duse + 1
            yield

    def f_issuer():
        slot = use = 0
        for k in range(nsl):
            g, it = k % G, k // G
            if it >= 1:
                yield from wait(zfree[g], (it - 1) & 1)
            for _ in range(NCH):
                yield from wait(full_f[slot], use & 1)
                empty_f[slot].arrive()
                slot += 1
                if slot == NSF:
                    slot, use = 0, use + 1
            zfull[g].arrive()

    def d_issuer():
        slot = use = 0
        for k in range(nsl):
            tile, tuse = k % NDZ, k // NDZ
            yield from wait(dzfull[tile], tuse & 1)
            for _ in range(NCH):
                yield from wait(full_d[slot], use & 1)
                empty_d[slot].arrive()
                slot += 1
                if slot == NSD:
                    slot, use = 0, use + 1
            order.append(k)
            dzfree[tile].arrive()
            dzmine[k % G].arrive()

    procs = [group(g) for g in range(G)] + [loader(), f_issuer(), d_issuer()]
    alive = [True] * len(procs)
    for _ in range(max_steps):
        for e in pend:
            e[0] -= 1
        for e in [e for e in pend if e[0] <= 0]:
            e[1].arrive()
        pend[:] = [e for e in pend if e[0] > 0]
        i = rnd.randrange(len(procs))
        if alive[i]:
            try:
                next(procs[i])
            except StopIteration:
                alive[i] = False
        if not any(alive):
            return order
    return None


@pytest.mark.parametrize("G,NCH", [(4, 5), (4, 1), (3, 5), (2, 3)])
def test_protocol_terminates_and_accumulates_in_slate_order(G, NCH):
    for nsl in (1, 2, 3, 5, 8, 28):
        for seed in range(25):
            order = run(G, 5, 4, 2, NCH, nsl, seed)
            assert order is not None, f"deadlock: G={G} chunks={NCH} slates={nsl} seed={seed}"
            assert order == list(range(nsl))                      # gradient jobs in slate order: bitwise reproducible sums


def test_protocol_without_the_own_slate_barrier_deadlocks():
    """The model is sharp enough to see the bug it was written for: a group two phases ahead of the shared tile barrier."""
    bad = [run(4, 5, 4, 2, 5, 28, seed, own_barrier=False, max_steps=100000) for seed in range(10)]
    assert any(o is None for o in bad)

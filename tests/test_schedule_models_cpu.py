"""CPU models of two pieces of device logic whose correctness must not depend on what the hardware happens to do:

* the one-wave slate hand-out of lambda_kernel (csrc/lambda_loss.cu): the k-th CTA to start on SM s takes entry
  k * nsm + (s or nsm - 1 - s) of the longest-first list, guarded by claim flags, and steals from the short end when that entry
  does not exist or is gone.  Whatever the placement of CTAs on SMs and whatever the interleaving, every slate is processed exactly
  once and no CTA takes two;
* the register bitonic network of csrc/sort_common.cuh (block_bitonic_packed_regs): thread t holds entries t E .. t E + E - 1, a
  compare-exchange at distance j is local (j < E), a lane exchange (j < 32 E) or a shared-memory exchange; the model runs the same
  index arithmetic on numpy arrays and must sort descending for every (threads, E) the kernels instantiate.
"""
import random

import numpy as np
import pytest


def _hand_out(B, nsm, placement, order_of_events, rng):
    """placement[c] = smid of CTA c; order_of_events: CTA indices in the order their pick runs (atomics are sequentially
    consistent per address, so any interleaving of whole picks and of steal steps is some sequential order of the atomics)."""
    arrive = [0] * 1024
    claimed = [0] * B
    cursor = 0
    got = {}
    pending = []                                         # CTAs in the middle of stealing: (cta,)
    for c in order_of_events:
        s = placement[c] % nsm
        k = arrive[s & 1023]
        arrive[s & 1023] += 1
        pos = k * nsm + (nsm - 1 - s if k & 1 else s)
        if pos < B and claimed[pos] == 0:
            claimed[pos] = 1
            got[c] = pos
        else:
            pending.append(c)
        # let a random subset of the stealing CTAs advance one step each (interleaved with later arrivals)
        for d in list(pending):
            if rng.random() < 0.5:
                continue
            q = cursor
            cursor += 1
            if q >= B:
                pending.remove(d)
                got[d] = None
                continue
            at = B - 1 - q
            if claimed[at] == 0:
                claimed[at] = 1
                got[d] = at
                pending.remove(d)
    while pending:                                       # drain
        d = pending[0]
        q = cursor
        cursor += 1
        if q >= B:
            got[d] = None
            pending.pop(0)
            continue
        at = B - 1 - q
        if claimed[at] == 0:
            claimed[at] = 1
            got[d] = at
            pending.pop(0)
    return got, claimed


@pytest.mark.parametrize("B,nsm", [(1024, 148), (1036, 148), (296, 148), (600, 148), (7, 3), (1000, 132)])
def test_one_wave_hand_out_is_exactly_once_for_any_placement(B, nsm):
    rng = random.Random(B * 1000 + nsm)
    for trial in range(30):
        kind = trial % 5
        if kind == 0:                                    # the expected placement: round robin over the SMs
            placement = [c % nsm for c in range(B)]
        elif kind == 1:                                  # everything the hardware could do: random SMs
            placement = [rng.randrange(nsm) for _ in range(B)]
        elif kind == 2:                                  # sparse smids beyond the SM count (collisions after the modulo)
            placement = [rng.randrange(4 * nsm) for _ in range(B)]
        elif kind == 3:                                  # everything on a few SMs: most CTAs must steal
            placement = [rng.randrange(3) for _ in range(B)]
        else:                                            # GPC-wise fill: consecutive CTAs on the same SM
            placement = [(c // 7) % nsm for c in range(B)]
        events = list(range(B))
        rng.shuffle(events)
        got, claimed = _hand_out(B, nsm, placement, events, rng)
        taken = [v for v in got.values() if v is not None]
        assert len(got) == B, "every CTA finishes its pick"
        assert sorted(taken) == list(range(B)), (B, nsm, kind, "every entry exactly once")
        assert all(claimed)


def test_round_robin_placement_balances_the_pair_work():
    """With the placement the schedule is built for, the per-SM totals of lim^2 agree to a few percent (random slates: +-25 %)."""
    rng = np.random.default_rng(3)
    B, nsm = 1024, 148
    lim = np.sort(rng.integers(128, 257, B))[::-1].astype(np.float64)          # the longest-first list
    work = lim ** 2
    per_sm = np.zeros(nsm)
    for s in range(nsm):
        for k in range(7):
            pos = k * nsm + (nsm - 1 - s if k & 1 else s)
            if pos < B:
                per_sm[s] += work[pos]
    full = per_sm[per_sm > 0]
    assert full.max() / full.mean() < 1.06
    shuffled = rng.permutation(work)
    rand_sm = np.array([shuffled[s::nsm].sum() for s in range(nsm)])
    assert rand_sm.max() / rand_sm.mean() > 1.15


def _bitonic_regs_model(keys, threads, E):
    """block_bitonic_packed_regs on a numpy array: v[t, e] = entry t * E + e."""
    LEN = threads * E
    v = np.asarray(keys, dtype=np.uint64).reshape(threads, E).copy()
    t = np.arange(threads)
    lane = t & 31
    i0 = t * E
    k = 2
    while k <= LEN:
        j = k >> 1
        while j > 0:
            if j < E:
                for e in range(E):
                    if (e & j) == 0 and (e | j) < E:
                        up = ((i0 + e) & k) == 0
                        a, c = v[:, e].copy(), v[:, e | j].copy()
                        sw = np.where(up, c > a, c < a)
                        v[:, e] = np.where(sw, c, a)
                        v[:, e | j] = np.where(sw, a, c)
            elif j < 32 * E:
                m = j // E
                take_max = ((lane & m) == 0) == ((i0 & k) == 0)
                partner = t ^ m                               # same warp: t ^ m keeps the upper bits of t
                assert np.all((partner >> 5) == (t >> 5))
                for e in range(E):
                    o = v[partner, e]
                    a = v[:, e]
                    v[:, e] = np.where((o > a) == take_max, o, a)
            else:
                flat = v.reshape(-1).copy()                   # through shared memory
                take_max = ((i0 & j) == 0) == ((i0 & k) == 0)
                for e in range(E):
                    o = flat[(i0 + e) ^ j]
                    a = v[:, e]
                    v[:, e] = np.where((o > a) == take_max, o, a)
            j >>= 1
        k <<= 1
    return v.reshape(-1)


@pytest.mark.parametrize("threads,E", [(128, 1), (128, 2), (128, 4), (128, 8), (64, 2), (32, 4)])
def test_register_bitonic_network_sorts_descending(threads, E):
    rng = np.random.default_rng(threads * 10 + E)
    for trial in range(4):
        LEN = threads * E
        if trial == 0:
            hi = rng.integers(0, 2 ** 32, LEN, dtype=np.uint64)
        elif trial == 1:
            hi = np.sort(rng.integers(0, 2 ** 32, LEN, dtype=np.uint64))               # ascending input
        elif trial == 2:
            hi = rng.integers(0, 4, LEN, dtype=np.uint64)                              # heavy key ties: the index bits decide
        else:
            hi = np.zeros(LEN, dtype=np.uint64)
        lo = (~np.arange(LEN, dtype=np.uint64)) & np.uint64(0xffffffff)                # ~index: ascending index first
        keys = (hi << np.uint64(32)) | lo
        out = _bitonic_regs_model(keys, threads, E)
        assert np.array_equal(out, np.sort(keys)[::-1])


# ---------------------------------------------------------------------------------------------------------------------
# csrc/peer_allreduce.cu: two slot parities are enough
# ---------------------------------------------------------------------------------------------------------------------
def _peer_allreduce_model(world, calls, ctas, rng, nbuf=2):
    """Every (rank, CTA) is a little state machine: PUSH (stores into slot [q & 1][rank] of every peer, one peer per micro-step, in
    a random order), FLAG (release-stores q into flag [q & 1][rank][cta] of every peer, after all of its pushes), WAIT (until its
    own flags [q & 1][*][cta] carry q), GATHER (reads the slots, one per micro-step), DONE.  A rank starts call q + 1 only when all
    of its CTAs are DONE with call q (stream order).  The scheduler picks a random runnable machine each step; a read that sees a
    value pushed for another call is the bug the two parities must exclude."""
    slot = [[[[None] * ctas for _ in range(world)] for _ in range(2)] for _ in range(world)]   # [owner][par][from][cta] -> call id
    par = (lambda q: q & 1) if nbuf == 2 else (lambda q: 0)
    flag = [[[[0] * ctas for _ in range(world)] for _ in range(2)] for _ in range(world)]
    call = [1] * world                                                 # the call a rank is executing
    state = {}                                                         # (rank, cta) -> [phase, todo list]

    def start(r):
        for b in range(ctas):
            peers = list(range(world))
            rng.shuffle(peers)
            state[(r, b)] = ["push", peers]

    for r in range(world):
        start(r)
    finished = 0
    steps = 0
    while finished < world:
        steps += 1
        assert steps < 10 ** 6, "deadlock"
        runnable = []
        for (r, b), (ph, todo) in state.items():
            q = call[r]
            if ph == "wait":
                if all(flag[r][par(q)][s][b] == q for s in range(world)):
                    runnable.append((r, b))
            elif ph != "done":
                runnable.append((r, b))
        assert runnable, "deadlock"
        r, b = rng.choice(runnable)
        ph, todo = state[(r, b)]
        q = call[r]
        if ph == "push":
            peer = todo.pop()
            slot[peer][par(q)][r][b] = q
            if not todo:
                state[(r, b)] = ["flag", list(range(world))]
        elif ph == "flag":
            peer = todo.pop()
            flag[peer][par(q)][r][b] = q
            if not todo:
                state[(r, b)] = ["wait", None]
        elif ph == "wait":
            state[(r, b)] = ["gather", list(range(world))]
        elif ph == "gather":
            s = todo.pop()
            assert slot[r][par(q)][s][b] == q, (r, b, s, q, slot[r][par(q)][s][b])
            if not todo:
                state[(r, b)] = ["done", None]
                if all(state[(r, c)][0] == "done" for c in range(ctas)):
                    if call[r] == calls:
                        finished += 1
                    else:
                        call[r] += 1
                        start(r)
    return steps


@pytest.mark.parametrize("world,ctas", [(2, 1), (2, 3), (3, 2), (8, 2)])
def test_peer_allreduce_two_parities_never_mix_calls(world, ctas):
    rng = random.Random(world * 100 + ctas)
    for trial in range(20):
        _peer_allreduce_model(world, calls=7, ctas=ctas, rng=rng)


def test_peer_allreduce_model_catches_a_single_parity():
    """The same model with ONE slot per sender fails for some interleaving (a rank pushes call q + 1 over values its peer has not
    gathered yet) -- the test above is not vacuous."""
    rng = random.Random(5)
    caught = 0
    for trial in range(50):
        try:
            _peer_allreduce_model(2, calls=5, ctas=1, rng=rng, nbuf=1)
        except AssertionError:
            caught += 1
    assert caught > 0

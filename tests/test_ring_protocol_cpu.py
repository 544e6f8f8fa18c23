"""Host model of the tile-ring hand-off of fdm_b200/csrc/xform_ring.cuh (two consumer groups, three tile buffers,
mbarrier parity waits), run over random schedules.

An mbarrier parity wait `try_wait.parity(P)` succeeds as soon as the phase of parity P is complete, i.e. whenever the
number of completed phases c satisfies (c & 1) != P -- it cannot tell phase j from phase j - 2.  With only the `full`
barriers, the consumer of use j of a buffer can reach its wait while use j - 1 (armed by itself, consumed by the OTHER
group) has not even landed: c = j - 1, the test passes on the stale phase and the group reads a tile that is not there
(the launch failure seen on the GPU when a concurrent kernel saturated DRAM).  With the second barrier `armed` -- arrived
on by whoever arms the buffer -- the consumer first waits for the arming of its own use; the previous arming of that
buffer was done by the consumer's own group, so it is never more than one phase ahead there, and once use j is armed
`full` is in phase j or j + 1.  The model checks exactly that: the fixed protocol never consumes a tile that has not
landed, in any schedule; the old protocol does in some."""
import random

NBUF, NG = 3, 2


class Barrier:
    def __init__(self):
        self.completed = 0          # phases completed so far

    def passes(self, parity):
        return (self.completed & 1) != parity


def run(ntiles, rng, with_armed, slow_load_prob):
    """Returns the list of (tile, what) protocol violations of one random schedule."""
    full = [Barrier() for _ in range(NBUF)]
    armed = [Barrier() for _ in range(NBUF)]
    content = [None] * NBUF             # tile whose data sits in the buffer (None while a load is in flight)
    in_flight = []                      # [buffer, tile, remaining delay]
    pos = [g for g in range(NG)]        # next tile index of each group
    state = ["wait"] * NG               # wait -> compute -> wait ...
    busy = [0] * NG
    bad = []

    def arm(tile):
        s = tile % NBUF
        content[s] = None
        delay = rng.randint(20, 60) if rng.random() < slow_load_prob else rng.randint(0, 3)
        in_flight.append([s, tile, delay])
        armed[s].completed += 1

    for t in range(min(NBUF, ntiles)):
        arm(t)
    steps = 0
    while any(p < ntiles for p in pos) and steps < 20000:
        steps += 1
        for ld in in_flight:            # loads make progress independently of the groups
            ld[2] -= 1
        for ld in [x for x in in_flight if x[2] <= 0]:
            in_flight.remove(ld)
            content[ld[0]] = ld[1]
            full[ld[0]].completed += 1
        g = rng.randrange(NG)           # one group gets the next scheduling slot
        i = pos[g]
        if i >= ntiles:
            continue
        s, parity = i % NBUF, (i // NBUF) & 1
        if state[g] == "wait":
            if with_armed and not armed[s].passes(parity):
                continue
            if not full[s].passes(parity):
                continue
            if content[s] != i:
                bad.append((i, f"group {g} passed the wait with buffer content {content[s]}"))
                content[s] = i          # (keep the run going)
            state[g] = "compute"
            busy[g] = rng.randint(1, 6)
        else:
            busy[g] -= 1
            if busy[g] <= 0:
                if i + NBUF < ntiles:
                    arm(i + NBUF)       # re-arm this buffer for the tile the other group will consume
                pos[g] += NG
                state[g] = "wait"
    if steps >= 20000:
        bad.append((-1, "the schedule did not terminate (the barriers went out of step)"))
    return bad


def test_fixed_protocol_never_consumes_an_unlanded_tile():
    rng = random.Random(1)
    for trial in range(300):
        bad = run(ntiles=rng.randint(1, 40), rng=rng, with_armed=True, slow_load_prob=rng.choice([0.0, 0.2, 0.6]))
        assert not bad, (trial, bad[:3])


def test_old_protocol_fails_under_slow_loads():
    rng = random.Random(2)
    failures = sum(bool(run(ntiles=30, rng=rng, with_armed=False, slow_load_prob=0.5)) for _ in range(200))
    assert failures > 0          # the stale-parity pass-through is reachable without the armed barriers


def test_old_protocol_is_fine_when_loads_are_fast():
    """... which is why it went unnoticed: with loads that land within a tile time nobody gets two phases ahead."""
    rng = random.Random(3)
    for _ in range(100):
        assert not run(ntiles=30, rng=rng, with_armed=False, slow_load_prob=0.0)

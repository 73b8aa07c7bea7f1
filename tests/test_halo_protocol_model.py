"""The hand-shake of the NVLink ghost-value exchange (saena_b200/csrc/halo_sync.cuh) as a state machine, run under
random schedules on the CPU.  What the device code relies on is restated rule by rule -- monotonic `arrived` /
`consumed` counters per (sender, receiver) pair, two landing buffers selected by the parity of the application count,
the pack role waiting for `consumed >= k - 1`, the receiving role waiting for `arrived >= k + 1`, per-rank epochs
advanced by the last CTA of a role -- and checked for the two properties the kernels need:

  safety    a receiver never reads a landing buffer that holds anything but the values of the application it is in
            (no overwrite before `consumed`, no read before `arrived`), whatever the interleaving, with ranks running
            up to two applications apart and with the fused form (one launch: pack CTAs and waiting CTAs concurrently)
            and the separate launches (pack kernel complete before the wait kernel starts) mixed freely
  progress  every schedule that keeps stepping some enabled action finishes all applications (no deadlock), provided a
            rank's launches run one after the other -- and the variant in which a rank's NEXT launch may occupy the
            device while this launch's pack role has not run (round 1's separate-launch form without the join of the
            comm stream, simplified to one operator) is shown to deadlock: the 4-GPU hang of round 1's driver run.

Pure Python, no device: this is a check of the protocol's rules, not of the CUDA code (tests/multigpu_check.py runs
that on real GPUs, fault injection included)."""
import random

import pytest


class Rank:
    def __init__(self, r, peers, n_apps, fused):
        self.r, self.peers, self.n_apps, self.fused = r, peers, n_apps, fused
        self.epoch_send = 0            # applications the pack role has completed
        self.epoch_recv = 0            # applications the receiving role has completed
        self.arrived = {p: 0 for p in peers}    # in MY memory, written by sender p
        self.consumed = {p: 0 for p in peers}   # in MY memory, written by receiver p
        self.landing = {p: [None, None] for p in peers}   # two buffers per sender
        # per-application progress of the current launch
        self.packed = False
        self.received = False


def enabled_actions(ranks, join_comm_stream=True, chip_full_of_waiters=False):
    """actions that may run now: ('pack', r) / ('recv', r).  One application per rank at a time (stream order); the
    fused form may run its two roles in either order, the separate launches pack first when the comm stream is joined."""
    out = []
    for R in ranks:
        if R.epoch_send < R.n_apps and not R.packed:
            ks = R.epoch_send
            # pack role of application ks: buffer ks & 1 of every receiver was last read in application ks - 2
            if all(R.consumed[p] >= ks - 1 for p in R.peers):
                blocked = False
                if chip_full_of_waiters and not R.received and R.epoch_recv == ks:
                    # the hazard of round 1: this rank's waiting CTAs already hold every SM slot and spin, so its own
                    # pack kernel (another stream, never joined) cannot be placed
                    blocked = getattr(R, "waiters_resident", False)
                if not blocked:
                    out.append(("pack", R.r))
        if R.epoch_recv < R.n_apps and not R.received:
            kr = R.epoch_recv
            must_pack_first = (not R.fused) and join_comm_stream
            if must_pack_first and not (R.packed or R.epoch_send > kr):
                continue
            if chip_full_of_waiters:
                R.waiters_resident = True      # the launch is on the device and spins
            if all(R.arrived[p] >= kr + 1 for p in R.peers):
                out.append(("recv", R.r))
    return out


def step(ranks, action):
    kind, r = action
    R = ranks[r]
    if kind == "pack":
        k = R.epoch_send
        for p in R.peers:
            P = ranks[p]
            # safety: the buffer I overwrite must not be in use -- the receiver is past application k - 2
            assert k < 2 or P.epoch_recv >= k - 1, f"rank {r} overwrites buffer {k & 1} of rank {p} too early"
            P.landing[r][k & 1] = (r, k)
            P.arrived[r] = k + 1
        R.packed = True
        R.epoch_send = k + 1
    else:
        k = R.epoch_recv
        for p in R.peers:
            assert R.landing[p][k & 1] == (p, k), f"rank {r} application {k}: buffer holds {R.landing[p][k & 1]}"
        for p in R.peers:
            ranks[p].consumed[r] = k + 1
        R.received = True
        R.epoch_recv = k + 1
        if hasattr(R, "waiters_resident"):
            R.waiters_resident = False
    # an application is over when both roles are done: the next launch of this rank may start
    if R.packed and R.received:
        R.packed = R.received = False


def run(n_ranks, n_apps, seed, forms, **kw):
    rng = random.Random(seed)
    ranks = []
    for r in range(n_ranks):
        peers = [p for p in (r - 1, r + 1) if 0 <= p < n_ranks]        # slab partition: two neighbours
        ranks.append(Rank(r, peers, n_apps, forms[r]))
    steps = 0
    while any(R.epoch_recv < n_apps or R.epoch_send < n_apps for R in ranks):
        acts = enabled_actions(ranks, **kw)
        if not acts:
            return False, steps, ranks
        step(ranks, rng.choice(acts))
        steps += 1
    return True, steps, ranks


@pytest.mark.parametrize("n_ranks", [2, 3, 4, 8])
def test_every_schedule_is_safe_and_finishes_with_any_mix_of_forms(n_ranks):
    for seed in range(200):
        rng = random.Random(1000 + seed)
        forms = [rng.random() < 0.5 for _ in range(n_ranks)]          # True: fused kernel, False: separate launches
        done, steps, ranks = run(n_ranks, 12, seed, forms)
        assert done, (n_ranks, seed, forms, [(R.epoch_send, R.epoch_recv) for R in ranks])
        assert all(R.epoch_send == R.epoch_recv == 12 for R in ranks)


def test_ranks_may_run_two_applications_apart_and_no_further():
    """the sender of a pair may be at most two applications ahead of the receiver: what the two buffers allow"""
    worst = 0
    for seed in range(300):
        rng = random.Random(seed)
        ranks = [Rank(0, [1], 10, True), Rank(1, [0], 10, True)]
        while any(R.epoch_recv < 10 or R.epoch_send < 10 for R in ranks):
            acts = enabled_actions(ranks)
            assert acts
            # bias the schedule towards rank 0 so that it runs ahead as far as the rules let it
            pref = [a for a in acts if a[1] == 0]
            step(ranks, rng.choice(pref if pref and rng.random() < 0.9 else acts))
            worst = max(worst, ranks[0].epoch_send - ranks[1].epoch_recv)
    assert worst == 2


def test_round_1_hazard_deadlocks_without_the_join_and_not_with_it():
    """separate launches whose pack kernel is not joined + a chip-full of spinning waiters of the same rank: some
    schedule stalls for ever (the driver's 4-GPU hang); with the join (pack complete before the wait starts) none does"""
    stalled = 0
    for seed in range(200):
        done, _, _ = run(4, 6, seed, [False] * 4, join_comm_stream=False, chip_full_of_waiters=True)
        stalled += not done
    assert stalled > 0
    for seed in range(200):
        done, _, _ = run(4, 6, seed, [False] * 4, join_comm_stream=True, chip_full_of_waiters=True)
        assert done

"""CPU model check of the flagged one-shot exchange of csrc/dbw_comm.cu (all_reduce_small_kernel): the protocol -- not the CUDA
code -- restated as per-(rank, element) state machines over shared words and run under thousands of random and adversarial
interleavings.  It checks what the kernel's header claims: every rank ends every exchange with the same, correct sums; a word
is never overwritten before its reader has consumed it; nobody waits forever -- and that the DEFERRED "done reading" stamp
(wait_previous_exchange + the last block's st.release of flagsB) is what guarantees it: with the wait removed the same checker
finds the overwrite.  The GPU tests (tests/test_multigpu.py) run the real kernel; this pins the reasoning behind it."""
import random

import pytest


class Overwritten(Exception):
    pass


def run_exchanges(world, n, exchanges, seed, wait_for_stamps=True, bias=None):
    """each rank runs `exchanges` flagged one-shot all-reduces of n values; a rank's exchange e is n element threads
    (wait -> push element i to every rank -> read element i from every slot -> arrive); the rank stamps `done e` into every
    peer's flag word when all its threads arrived, and only then starts exchange e + 1 (kernel boundary on its stream)."""
    rng = random.Random(seed)
    ll = [[[(0.0, 0) for _ in range(n)] for _ in range(world)] for _ in range(world)]      # ll[rank][slot][i] = (value, epoch)
    flags_b = [[0] * world for _ in range(world)]                                          # flags_b[rank][peer]
    value = lambda r, e, i: float((r + 1) * 1000 + e * 10 + i)
    results = [[None] * exchanges for _ in range(world)]

    def element_thread(r, e, i, out):
        if wait_for_stamps:
            for p in range(world):
                while flags_b[r][p] < e - 1:
                    yield 'spin'
        for q in range(world):
            ll[(r + q) % world][r][i] = (value(r, e, i), e)
            yield 'step'
        acc = 0.0
        for q in range(world):
            while True:
                v, ep = ll[r][q][i]
                if ep > e:
                    raise Overwritten(f'rank {r} exchange {e} element {i}: slot {q} already holds epoch {ep}')
                if ep == e:
                    break
                yield 'spin'
            acc += v
            yield 'step'
        out[i] = acc

    epoch = [1] * world
    outs = [[None] * n for _ in range(world)]
    threads = [[element_thread(r, 1, i, outs[r]) for i in range(n)] for r in range(world)]
    alive = [[True] * n for _ in range(world)]
    idle_rounds = 0
    while any(epoch[r] <= exchanges for r in range(world)):
        runnable = [(r, i) for r in range(world) if epoch[r] <= exchanges for i in range(n) if alive[r][i]]
        if bias is not None and rng.random() < 0.9:                      # adversarial: one rank gets almost all the turns
            favoured = [(r, i) for r, i in runnable if r == bias]
            runnable = favoured or runnable
        r, i = rng.choice(runnable)
        try:
            progressed = next(threads[r][i]) == 'step'
        except StopIteration:
            alive[r][i], progressed = False, True
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 20000 * world * n, 'every thread is spinning: deadlock'
        if not any(alive[r]):                                            # the rank's last thread arrived: stamp, next exchange
            e = epoch[r]
            results[r][e - 1] = list(outs[r])
            for p in range(world):
                flags_b[p][r] = e
            epoch[r] = e + 1
            if epoch[r] <= exchanges:
                outs[r] = [None] * n
                threads[r] = [element_thread(r, epoch[r], i2, outs[r]) for i2 in range(n)]
                alive[r] = [True] * n
    expect = [[sum(value(r, e, i) for r in range(world)) for i in range(n)] for e in range(1, exchanges + 1)]
    for r in range(world):
        assert results[r] == expect, (r, results[r], expect)


@pytest.mark.parametrize('world', [2, 3, 8])
def test_flagged_one_shot_exchange_is_correct_under_any_interleaving(world):
    for seed in range(150 if world < 8 else 25):
        run_exchanges(world, n=3, exchanges=4, seed=seed)
        run_exchanges(world, n=2, exchanges=4, seed=seed, bias=seed % world)          # one rank runs far ahead of the others


def test_without_the_deferred_stamp_a_fast_rank_overwrites_unread_words():
    """negative control: drop wait_previous_exchange and the checker finds the hazard the stamp exists for"""
    found = 0
    for seed in range(200):
        try:
            run_exchanges(2, n=2, exchanges=4, seed=seed, wait_for_stamps=False, bias=seed % 2)
        except (Overwritten, AssertionError):
            found += 1
    assert found > 0

"""Exhaustive interleaving check of the grid-level cross-GPU barrier protocol (ops/csrc/tfy_common.cuh:
tfy_grid_epoch / tfy_grid_entry / tfy_grid_exit / tfy_grid_finish) on a small model.

compute-sanitizer's racecheck does not model system-scope flag protocols between GPUs, and functional tests only
see the interleavings the hardware happens to produce.  This test enumerates EVERY interleaving of R ranks x C CTAs
running K back-to-back launches of a fused-step-shaped kernel (as a CUDA graph replays them) and checks:

  P1  a CTA only reads a peer's gradients of launch k after that peer has STARTED launch k (its backward is done);
  P2  a rank's launch k only completes after every rank's stores of launch k were issued (the all-gather landed);
  P3  a peer's stores of launch k never hit a rank that is still reading in launch k-1 (replay safety);
  P4  no deadlock: every reachable state can still make progress until all launches are complete;
  P5  the per-slot epoch words end at K on every rank (monotonic pads + local epochs survive replays).

Model of one CTA (program order; "atomic" = one indivisible step, as `red.release.sys`, `atom.acq_rel` and
`ld.acquire.sys` are):
  0 e := epoch + 1                          (tfy_grid_epoch, read at kernel start)
  1 CTA 0 only: entry_pad[p][me] += 1       one atomic step per rank p, self included   (tfy_grid_signal)
  2 wait entry_pad[me][p] >= e for all p    (tfy_grid_poll)
  3 read the peers' gradients of this launch                                       -> P1, P3 checked here
  4 store the results into every replica                                           (marks stores[me][k])
  5 arrive on the local exit counter; the LAST arriver signals exit_pad[p][me] += 1 for all p, then waits for
    exit_pad[me][p] >= e for all p; the other CTAs leave at once, or poll as well when `all_wait` (the variant used
    when the kernel goes on to clear gradients the peers were reading)             (tfy_grid_exit)
  6 arrive on the local finish counter; the LAST arriver sets epoch := e and clears both counters (tfy_grid_finish)
A rank starts launch k+1 when all its CTAs finished launch k (stream order).  `mutation` removes one ingredient of
the protocol so that the checker is shown to catch the corresponding bug.
"""
from collections import deque

import pytest


def explore(R: int, C: int, K: int, mutation: str = "", all_wait: bool = False):
    """Breadth-first search over all interleavings; returns (n_states, violations, deadlocks, final_epochs)."""
    # rank state : (launch k in 1..K+1, epoch word, exit counter, finish counter, ctas)
    # CTA state  : (pc, e, sub)   pc 0..7 (7 = waiting for its siblings), sub = position inside a signal loop
    # pads       : entry[dst][src], exit_[dst][src]  monotonic counters in dst's memory
    fresh = tuple((0, 0, 0) for _ in range(C))
    zero = tuple(tuple(0 for _ in range(R)) for _ in range(R))
    start = (tuple((1, 0, 0, 0, fresh) for _ in range(R)), zero, zero)

    def bump(pads, dst, src):
        row = pads[dst]
        return pads[:dst] + (row[:src] + (row[src] + 1,) + row[src + 1:],) + pads[dst + 1:]

    def in_launch(rank, k):           # the rank's launch k has been issued (its earlier work is complete)
        return rank[0] >= k

    def stores_issued(rank, k):       # every CTA of the rank is past its stores of launch k
        return rank[0] > k or (rank[0] == k and all(cta[0] >= 5 for cta in rank[4]))

    violations = set()
    deadlocks = 0
    final_epochs = None
    seen = {start}
    todo = deque([start])
    while todo:
        ranks, entry, exit_ = todo.popleft()
        succ = []
        if all(rank[0] > K for rank in ranks):
            final_epochs = tuple(rank[1] for rank in ranks)
            continue
        for r in range(R):
            k, epoch, exit_cnt, fin_cnt, ctas = ranks[r]
            if k > K:
                continue
            for c in range(C):
                pc, e, sub = ctas[c]

                def nxt(new_cta, new_entry=entry, new_exit=exit_, epoch_=epoch, exit_cnt_=exit_cnt, fin_cnt_=fin_cnt):
                    rank = (k, epoch_, exit_cnt_, fin_cnt_, ctas[:c] + (new_cta,) + ctas[c + 1:])
                    succ.append((ranks[:r] + (rank,) + ranks[r + 1:], new_entry, new_exit))

                if pc == 0:                                   # e = epoch + 1
                    nxt((1, epoch + 1, 0))
                elif pc == 1:                                 # entry signals (CTA 0 only), one rank per step
                    if c != 0 or mutation == "no_entry_signal":
                        nxt((2, e, 0))
                    else:
                        nxt((1, e, sub + 1) if sub + 1 < R else (2, e, 0), new_entry=bump(entry, sub, r))
                elif pc == 2:                                 # entry poll
                    if mutation == "no_entry_poll" or all(entry[r][p] >= e for p in range(R)):
                        nxt((3, e, 0))
                elif pc == 3:                                 # read the peers' gradients of launch k
                    if not all(in_launch(ranks[p], k) for p in range(R)):
                        violations.add("P1: gradients read before a peer issued the launch")
                    nxt((4, e, 0))
                elif pc == 4:                                 # store the results into every replica
                    if not all(in_launch(ranks[p], k) for p in range(R)):
                        violations.add("P3: stores hit a rank that has not finished the previous launch")
                    nxt((5, e, 0))
                elif pc == 5:                                 # exit barrier
                    if sub == 0:                              # arrive on the local counter
                        if exit_cnt != C - 1:                  # not last: leave at once, or (all_wait) poll as well
                            nxt((5, e, R + 1) if all_wait else (6, e, 0), exit_cnt_=exit_cnt + 1)
                        elif mutation == "no_exit_signal":
                            nxt((5, e, R + 1), exit_cnt_=exit_cnt + 1)
                        else:
                            nxt((5, e, 1), exit_cnt_=exit_cnt + 1)
                    elif sub <= R:                            # last arriver: signal rank sub-1
                        nxt((5, e, sub + 1), new_exit=bump(exit_, sub - 1, r))
                    elif mutation == "no_exit_poll" or all(exit_[r][p] >= e for p in range(R)):
                        nxt((6, e, 0))                        # ... and wait for everybody's exit signal
                elif pc == 6:                                 # finish
                    if fin_cnt != C - 1:
                        nxt((7, e, 0), fin_cnt_=fin_cnt + 1)
                    else:
                        # the rank's launch k is complete (all CTAs exited): P2, then the next launch is issued
                        if not all(stores_issued(ranks[p], k) for p in range(R) if p != r):
                            violations.add("P2: launch completed before a peer's stores were issued")
                        rank = (k + 1, e, 0, 0, fresh)
                        succ.append((ranks[:r] + (rank,) + ranks[r + 1:], entry, exit_))
                # pc == 7: this CTA is done; the launch ends when the last sibling finishes
        if not succ:
            deadlocks += 1
            continue
        for st in succ:
            if st not in seen:
                seen.add(st)
                todo.append(st)
    return len(seen), violations, deadlocks, final_epochs


@pytest.mark.parametrize("all_wait", [False, True])
@pytest.mark.parametrize("R,C,K", [(2, 2, 2), (3, 1, 2), (2, 2, 3), (4, 1, 1), (3, 2, 1)])
def test_protocol_holds_under_every_interleaving(R, C, K, all_wait):
    n, violations, deadlocks, final_epochs = explore(R, C, K, all_wait=all_wait)
    assert n > 100                                   # the search really branched
    assert violations == set(), violations
    assert deadlocks == 0
    assert final_epochs == tuple(K for _ in range(R))


@pytest.mark.parametrize("mutation,expected", [
    ("no_entry_poll", "P1"),          # reading without waiting for the peers' "started" signals
    ("no_exit_poll", "P2"),           # completing without waiting for the peers' "stores landed" signals
])
def test_checker_catches_a_broken_protocol(mutation, expected):
    _, violations, _, _ = explore(2, 2, 2, mutation)
    assert any(v.startswith(expected) for v in violations), (mutation, violations)


def test_missing_signal_deadlocks():
    _, _, deadlocks, _ = explore(2, 1, 1, "no_entry_signal")
    assert deadlocks > 0

"""Exhaustive interleaving check of the IVF-PQ scan's tile ring (csrc/ivfpq.cu, `ivfpq_scan_query_kernel`).

The kernel streams tiles through a ring of B shared-memory buffers: one producer thread issues asynchronous fills
(TMA bulk copies that complete a `full[b]` mbarrier phase when they LAND, in any order), two consumer groups take
alternate tiles, wait for `full[b]` with `try_wait.parity(use & 1)`, read the buffer and arrive on `empty[b]`; the
producer waits for `empty[b]` of the previous use before refilling.  `try_wait.parity(p)` succeeds when the
barrier's current phase parity differs from p - it cannot tell "phase u completed" from "phase u-1 not yet
completed".  With an ODD ring depth a buffer alternates between the groups, so a group can wait for use u of a
buffer without ever having waited for use u-1: the protocol the kernel shipped with until round 2 ("old") lets a
consumer read a buffer whose fill has not landed (the fault of profiles/r02_ivfpq_ring_fault_old_kernel.log).

This test explores EVERY interleaving of producer, fills and consumers on a small ring for three protocols:
  old      - as shipped in round 1: must exhibit the stale read (the model reproduces the bug)
  counter  - the fix in the tree: a consumer first waits until the producer has ISSUED its tile
  all_pass - the first fix (every group passes every tile's barriers; correct, 8 % slower on the GPU)
The fixed protocols must be free of stale reads, overwrites of unread buffers and deadlocks.
"""
from collections import deque

import pytest


def explore(protocol: str, B: int, T: int, groups: int = 2):
    """Returns (violations, deadlocks, states).  A state is
       (next tile to issue, in-flight fills, full phase counts, empty (arrivals, phases) per buffer,
        buffer contents, per consumer group: (next tile, stage))."""
    need = groups if protocol == "all_pass" else 1          # arrivals that complete an empty phase
    first = tuple(0 if protocol == "all_pass" else g for g in range(groups))
    start = (0, frozenset(), (0,) * B, ((0, 0),) * B, (-1,) * B, tuple((first[g], 0) for g in range(groups)))
    seen, todo = {start}, deque([start])
    violations, deadlocks = [], []

    def parity_ok(completed, p):            # mbarrier.try_wait.parity
        return (completed & 1) != p

    while todo:
        st = todo.popleft()
        issue, flight, full, empty, content, cons = st
        nxt = []
        # producer: issue the next tile once the previous use of its buffer was released
        if issue < T:
            b, use = issue % B, issue // B
            if use == 0 or parity_ok(empty[b][1], (use - 1) & 1):
                nxt.append((issue + 1, flight | {issue}, full, empty, content, cons))
        # any in-flight fill may land now (completion order is not issue order)
        for g in flight:
            b = g % B
            full2 = full[:b] + (full[b] + 1,) + full[b + 1:]
            nxt.append((issue, flight - {g}, full2, empty, content[:b] + (g,) + content[b + 1:], cons))
        # consumers
        for c, (g, stage) in enumerate(cons):
            if g >= T:
                continue
            b, use = g % B, g // B
            mine = g % groups == c
            if protocol == "counter" and issue <= g:
                continue                                   # spins on the issued-tile counter
            if not parity_ok(full[b], use & 1):
                continue                                   # blocked on full[b]
            if mine and content[b] != g:
                violations.append((protocol, "stale or overwritten buffer", g, st))
                continue
            arr, ph = empty[b]
            arr += 1
            if arr == need:
                arr, ph = 0, ph + 1
            empty2 = empty[:b] + ((arr, ph),) + empty[b + 1:]
            step = 1 if protocol == "all_pass" else groups
            nxt.append((issue, flight, full, empty2, content, cons[:c] + ((g + step, 0),) + cons[c + 1:]))
        done = issue >= T and not flight and all(g >= T for g, _ in cons)
        if not nxt and not done:
            deadlocks.append((protocol, st))
        for s in nxt:
            if s not in seen:
                seen.add(s)
                todo.append(s)
    return violations, deadlocks, len(seen)


@pytest.mark.parametrize("B,T", [(3, 9), (5, 12)])
def test_old_ring_protocol_has_the_stale_read_with_an_odd_depth(B, T):
    violations, _, _ = explore("old", B, T)
    assert violations, "the model should reproduce the round-1 fault"
    # the earliest stale read is the first re-use of a buffer: tile B (buffer 0, group B % 2 = 1) let through while
    # tile 0 (buffer 0, group 0) has not landed
    assert min(v[2] for v in violations) == B


@pytest.mark.parametrize("B,T", [(4, 12), (2, 8)])
def test_old_ring_protocol_is_fine_with_an_even_depth(B, T):
    """m = 16 (4 tiles) never had the hazard: a buffer always belongs to the same group."""
    violations, deadlocks, _ = explore("old", B, T)
    assert not violations and not deadlocks


@pytest.mark.parametrize("protocol", ["counter", "all_pass"])
@pytest.mark.parametrize("B,T", [(3, 9), (5, 12), (4, 10), (7, 16)])
def test_fixed_ring_protocols_are_safe_in_every_interleaving(protocol, B, T):
    violations, deadlocks, states = explore(protocol, B, T)
    assert not violations, violations[:1]
    assert not deadlocks, deadlocks[:1]
    assert states > T

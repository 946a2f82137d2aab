"""Exhaustive interleaving check of the buffer hand-over protocol of the ring-buffered row transforms
(k2_fft16_ring / k4_fft16_ring, csrc/k2_fft.cu) - host-side model, no GPU.

The protocol: a CTA owns rows 0 .. n-1; two groups take them alternately (group g: rows g, g+2, ...); row j
lives in buffer j mod 3.  Rows 0-2 are requested up front.  A group waits for its row on an mbarrier with a phase
PARITY (an mbarrier parity wait succeeds when the barrier's current phase parity differs from the awaited one,
i.e. it cannot tell "phase n has completed" from "phase n - 1 has not"), transforms the row in place, and then
requests row j + 3 into the same buffer, which the OTHER group will consume.  Requested loads complete after
an arbitrary delay, in any order.

Two barrier assignments are modelled: ONE barrier per buffer (index j mod 3, parity (j // 3) & 1 - the first
version of the kernel), which the search shows to be unsafe when a load is slow enough (a group can sail through
the wait for row j while the copy of row j - 3 is still in flight), and TWO per buffer used alternately (index
(j mod 3) * 2 + ((j // 3) & 1), parity (j // 6) & 1 - ring_bar / ring_parity in the kernel), which is safe for
every interleaving: the previous use of that barrier is row j - 6, which the waiting group consumed itself.

Checked over every interleaving of the two groups and the load completions:
  * a load is never requested into, and never lands in, a buffer a group is working in;
  * a group only ever starts on a buffer that holds exactly its row (the parity wait is never fooled by an
    older or newer phase);
  * at most one load is in flight per buffer (its mbarrier is armed once per phase);
  * no deadlock: every row is transformed exactly once."""
import collections

NBUF = 3


def bar_single(j):
    return j % NBUF, (j // NBUF) & 1


def bar_double(j):
    return (j % NBUF) * 2 + ((j // NBUF) & 1), (j // (2 * NBUF)) & 1


def explore(n, bar_of=bar_double, nbar=2 * NBUF):
    # state: (pos[2], working[2], content[3], inflight[3], phase[3])
    #   pos[g]      : next row index of group g (g, g+2, ...)
    #   working[g]  : row the group is transforming, or -1
    #   content[b]  : row whose data the buffer holds (-1 = none)
    #   inflight[b] : row being loaded into the buffer, or -1
    #   phase[i]    : completed phases of mbarrier i
    init_inflight = tuple(j if j < n else -1 for j in range(NBUF))
    start = ((0, 1), (-1, -1), (-1,) * NBUF, init_inflight, (0,) * nbar)
    seen, todo, finals = {start}, collections.deque([start]), 0
    while todo:
        pos, working, content, inflight, phase = st = todo.popleft()
        succ = []
        # a load completes
        for b in range(NBUF):
            if inflight[b] >= 0:
                assert b not in [w % NBUF for w in working if w >= 0], f"load lands in a buffer in use: {st}"
                c, f, p = list(content), list(inflight), list(phase)
                i, _ = bar_of(inflight[b])
                c[b], f[b], p[i] = inflight[b], -1, phase[i] + 1
                succ.append((pos, working, tuple(c), tuple(f), tuple(p)))
        for g in range(2):
            if working[g] < 0 and pos[g] < n:
                j = pos[g]
                b = j % NBUF
                i, awaited = bar_of(j)
                # mbarrier.try_wait.parity succeeds when the parity of the barrier's current (incomplete) phase,
                # i.e. of the number of completed phases, differs from the awaited parity
                if phase[i] % 2 != awaited:
                    assert content[b] == j, f"group {g} starts row {j} on a buffer holding row {content[b]}: {st}"
                    assert inflight[b] < 0, f"group {g} starts on a buffer with a load in flight: {st}"
                    w = list(working)
                    w[g] = j
                    succ.append((pos, tuple(w), content, inflight, phase))
            elif working[g] >= 0:
                j = working[g]
                b = j % NBUF
                w, p, f = list(working), list(pos), list(inflight)
                w[g], p[g] = -1, j + 2
                if j + NBUF < n:
                    assert inflight[b] < 0, f"second load requested into buffer {b}: {st}"
                    assert all(x < 0 or x % NBUF != b for k, x in enumerate(working) if k != g), \
                        f"load requested into a buffer the other group works in: {st}"
                    f[b] = j + NBUF
                succ.append((tuple(p), tuple(w), content, tuple(f), phase))
        if not succ:
            assert pos[0] >= n and pos[1] >= n and working == (-1, -1) and all(x < 0 for x in inflight), f"deadlock: {st}"
            finals += 1
        for s in succ:
            if s not in seen:
                seen.add(s)
                todo.append(s)
    return len(seen), finals


def test_ring_hand_over_is_safe_for_every_interleaving():
    for n in (1, 2, 3, 4, 5, 6, 7, 10, 13, 14):
        states, finals = explore(n)
        assert finals >= 1 and states > n


def test_one_barrier_per_buffer_is_not_safe():
    """The search is sharp enough to find the flaw of the first version: with one barrier per buffer a group
    can start on a buffer that does not hold its row (needs a load slower than a whole row transform)."""
    import pytest
    with pytest.raises(AssertionError, match="on a buffer holding row"):
        explore(8, bar_of=bar_single, nbar=NBUF)

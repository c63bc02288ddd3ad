"""Model check (CPU, pure Python) of the shared-memory ring of the dense Cholesky's deep update
(apex_solver_b200/csrc/explicit.cu: chol_syrk128_kernel): three buffers, slice s lives in buffer s % 3, every warp copies its own
part of a slice asynchronously (cp.async), the slice barrier is SPLIT - a warp ARRIVES in the middle of slice s (step 4 of 8) after
waiting for its own copies of slice s+1, and WAITS at the last step of slice s for everybody's arrival; behind the wait it issues its
copies of slice s+2 (into the buffer slice s-1 used) and reads its first fragments of slice s+1. Fragments of step st+1 are read during
step st (one step ahead).
Checked under random warp interleavings and random landing times of the asynchronous copies: every fragment read sees ALL warps'
parts of the slice it expects (landed, not yet overwritten). Negative controls: issuing the copies of slice s+2 BEFORE the wait, or
reading slice s+1 before the wait, is caught."""
import random

import pytest

STEPS = 8


def run(W, S, rng, issue_before_wait=False, read_before_wait=False, max_steps=500000):
    buf = [[None] * W for _ in range(3)]          # buf[b][w] = slice whose part w currently sits in buffer b
    pending = []                                  # (warp, slice) copies in flight
    arrived = [0] * (S + 1)                       # arrivals per phase (= slice index)

    def issue(w, sl):
        if sl < S:
            pending.append((w, sl))

    def land(i):
        w, sl = pending.pop(i)
        buf[sl % 3][w] = sl

    def land_own(w, upto=None):                   # cp.async.wait_group: this warp's copies (of slices <= upto) have landed
        for i in range(len(pending) - 1, -1, -1):
            if pending[i][0] == w and (upto is None or pending[i][1] <= upto):
                land(i)

    def check(sl):
        b = buf[sl % 3]
        return all(x == sl for x in b)

    for w in range(W):                            # prologue: stage(0), stage(1); wait_group 1; __syncthreads
        issue(w, 0); issue(w, 1)
    for w in range(W):
        land_own(w, upto=0)
    pc = [(0, 0, "top")] * W                      # (slice, step, micro state)
    for _ in range(max_steps):
        if pending and rng.random() < 0.3:
            land(rng.randrange(len(pending)))     # an asynchronous copy lands at a random moment
            continue
        live = [w for w in range(W) if pc[w][0] < S]
        if not live:
            return "ok", None
        w = rng.choice(live)
        s, st, ms = pc[w]
        if ms == "top":
            if st == STEPS // 2:
                land_own(w)                       # wait_group 0
                arrived[s] += 1                   # mbarrier.arrive
            if st + 1 < STEPS:
                if not check(s):                  # fragments of step st+1 from slice s
                    return "bad read (current slice)", (w, s, st, buf)
                pc[w] = (s, st + 1, "top")
            else:
                if issue_before_wait:
                    issue(w, s + 2)
                if read_before_wait and s + 1 < S and not check(s + 1):
                    return "bad read (next slice, before the wait)", (w, s, buf)
                pc[w] = (s, st, "wait")
        else:                                     # last step of slice s: wait for everybody's arrival
            if arrived[s] < W:
                continue
            if not issue_before_wait:
                issue(w, s + 2)
            if s + 1 < S and not check(s + 1):    # first fragments of slice s+1
                return "bad read (next slice)", (w, s, buf)
            pc[w] = (s + 1, 0, "top")
    return "stuck", pc


@pytest.mark.parametrize("W,S", [(2, 6), (4, 8), (16, 5)])
def test_split_slice_barrier_ring_is_safe(W, S):
    rng = random.Random(99 + W)
    for _ in range(200):
        kind, detail = run(W, S, rng)
        assert kind == "ok", (kind, detail)


def test_ring_negative_controls():
    rng = random.Random(5)
    assert any(run(4, 8, rng, issue_before_wait=True)[0] != "ok" for _ in range(300)), "overwriting slice s-1's buffer before everybody left it must be caught"
    assert any(run(4, 8, rng, read_before_wait=True)[0] != "ok" for _ in range(300)), "reading slice s+1 before everybody's copies landed must be caught"

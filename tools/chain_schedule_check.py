#!/usr/bin/env python
"""Offline check of the split-chain Cholesky schedule (gpk_fit_begin, option chainsplit): the launches, event records
and stream waits of the C++ loop are transcribed into a task graph (stream order + events = happens-before), and the
script verifies that
  * every two launches that touch the same 128 x 128 tile with at least one write are ordered, and
  * every tile receives its panels in increasing order (bit-identical to the plain look-ahead schedule).
Runs without a GPU:  python tools/chain_schedule_check.py [nb ...]"""
import itertools
import sys


def build(nb):
    tasks = []          # (name, stream, reads, writes, k)
    edges = set()       # (a, b): a happens before b
    last_on = {}        # stream -> last task index
    events = {}         # event name -> task index it was recorded after

    def launch(name, stream, reads, writes, k):
        idx = len(tasks)
        tasks.append((name, stream, set(reads), set(writes), k))
        if stream in last_on:
            edges.add((last_on[stream], idx))
        for w in pending_waits.pop(stream, []):
            edges.add((w, idx))
        last_on[stream] = idx
        return idx

    pending_waits = {}

    def record(ev, stream):
        # an event completes when everything before it on the stream has completed (incl. what the stream waited for)
        events[ev] = ("rec", last_on.get(stream), list(pending_waits.get(stream, [])))

    def wait(stream, ev):
        kind, t, extra = events[ev]
        for x in ([t] if t is not None else []) + extra:
            pending_waits.setdefault(stream, []).append(x)

    K = lambda i, j: ("K", i, j)
    Pt = lambda k: ("P", k, k)
    haveX, havePU, haveRA = [0] * nb, [0] * nb, [0] * nb
    kbuild = launch("kbuild", "C", [], [K(i, j) for i in range(nb + 1) for j in range(min(i, nb - 1) + 1)], -1)
    record("fork", "C"); wait("P", "fork"); wait("R", "fork")
    for k in range(nb):
        launch("diag(%d)" % k, "C", [K(k, k)], [K(k, k), Pt(k)], k)
        record("D%d" % k, "C")
        rows_below = list(range(k + 1, nb + 1))                      # block rows of the panel (nb = rhs row)
        hasX = k + 1 < nb
        if hasX:
            if k >= 1 and havePU[k - 1]: wait("C", "PU%d" % (k - 1))
            if k >= 1 and haveRA[k - 1]: wait("C", "RA%d" % (k - 1))
            launch("X(%d)" % k, "C", [K(k + 1, k), Pt(k), K(k + 1, k + 1)], [K(k + 1, k), K(k + 1, k + 1)], k)
            record("X%d" % k, "C"); haveX[k] = 1
        wait("P", "D%d" % k)
        trsm_rows = [i for i in rows_below if not (hasX and i == k + 1)]
        if trsm_rows:
            launch("solve'(%d)" % k, "P", [K(i, k) for i in trsm_rows] + [Pt(k)], [K(i, k) for i in trsm_rows], k)
        record("T%d" % k, "P")
        pu_rows = [i for i in rows_below if i != k + 1] if hasX else []
        if hasX and pu_rows:
            wait("P", "X%d" % k)
            if k >= 1 and haveRA[k - 1]: wait("P", "RA%d" % (k - 1))
            launch("update'(%d)" % k, "P", [K(i, k) for i in pu_rows] + [K(k + 1, k)] + [K(i, k + 1) for i in pu_rows],
                   [K(i, k + 1) for i in pu_rows], k)
            record("PU%d" % k, "P"); havePU[k] = 1
        # trailing update beyond block column k+1
        cols = list(range(k + 2, nb))
        if cols:
            wait("R", "T%d" % k)
            if hasX: wait("R", "X%d" % k)
            ja = k + 2
            ta = [(i, ja) for i in range(ja, nb + 1)]
            launch("rest_a(%d)" % k, "R", [K(i, k) for i, _ in ta] + [K(ja, k)] + [K(i, j) for i, j in ta], [K(i, j) for i, j in ta], k)
            record("RA%d" % k, "R"); haveRA[k] = 1
            tb = [(i, j) for j in range(k + 3, nb) for i in range(j, nb + 1)]
            if tb:
                launch("rest_b(%d)" % k, "R", [K(i, k) for i, _ in tb] + [K(j, k) for _, j in tb] + [K(i, j) for i, j in tb],
                       [K(i, j) for i, j in tb], k)
    record("joinP", "P"); wait("C", "joinP"); record("joinR", "R"); wait("C", "joinR")
    launch("qfill", "C", [Pt(k) for k in range(nb)], [], nb)
    return tasks, edges


def check(nb):
    tasks, edges = build(nb)
    n = len(tasks)
    reach = [set() for _ in range(n)]
    succ = [[] for _ in range(n)]
    for a, b in edges:
        succ[a].append(b)
    for a in range(n - 1, -1, -1):                       # tasks are created in a topological order
        for b in succ[a]:
            reach[a].add(b)
            reach[a] |= reach[b]
    bad = 0
    for a, b in itertools.combinations(range(n), 2):
        ta, tb = tasks[a], tasks[b]
        conflict = (ta[3] & (tb[2] | tb[3])) | (tb[3] & ta[2])
        if conflict and b not in reach[a] and a not in reach[b]:
            bad += 1
            print("nb=%d UNORDERED: %s (%s) vs %s (%s) on %s" % (nb, ta[0], ta[1], tb[0], tb[1], sorted(conflict)[:3]))
    # panels per tile in increasing order: a writer with smaller k must happen before a writer with larger k
    for a, b in itertools.combinations(range(n), 2):
        ta, tb = tasks[a], tasks[b]
        if ta[3] & tb[3] and ta[4] != tb[4]:
            first, second = (a, b) if ta[4] < tb[4] else (b, a)
            if second not in reach[first]:
                bad += 1
                print("nb=%d ORDER: %s should precede %s" % (nb, tasks[first][0], tasks[second][0]))
    print("nb=%d: %d launches, %d happens-before edges, %d problems" % (nb, n, len(edges), bad))
    return bad


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [3, 4, 5, 8, 16]
    sys.exit(1 if sum(check(nb) for nb in sizes) else 0)

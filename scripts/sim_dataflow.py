"""Discrete-event model of potrf_dataflow_kernel's schedule (host only; no GPU): the real ticket list (decoded by
the kernel's own function through gpar_debug_decode_ticket) is played on `grid` workers with per-phase costs
measured on the B200 (scripts/prof_budget.py).  Used to find out where the sweep is chain-bound and what a change
of the ticket order / split-K rule / chain latency would buy before spending GPU time on it.

    python scripts/sim_dataflow.py 8424 [key=value ...]      e.g. fac_flag=20 solve=9
"""
import ctypes as C
import heapq
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpar_b200 import _lib  # noqa: E402

D0, HEAD, PLAIN, PRE = 0, 1, 2, 3
DEFAULT = dict(kt=17.72, kt_y=4.0, epi=8.0, solve=14.0, ticket=0.6, syrk=11.4, asm=3.8, fac_flag=28.0, fac_rest=5.0,
               d0=36.0, refill=1.5, poll=0.2,
               # pipelined HEAD (round 2; pipelined=1): newest k-tile applied on chip, four solve + SYRK blocks behind
               # the panel flags of the previous diagonal tile, which go up fac_panel us apart
               pipelined=1.0, newest=22.0, block=7.5, publish=3.0, fac_panel=9.0)


def tickets(n, nb, batch, grid):
    lib = _lib.load()
    total = lib.gpar_debug_decode_ticket(n, nb, batch, grid, -1, None)
    out = (C.c_int32 * 6)()
    res = []
    for t in range(total):
        lib.gpar_debug_decode_ticket(n, nb, batch, grid, t, out)
        res.append(tuple(out))
    return res


def simulate(n, nb=1, grid=148, P=None, verbose=False):
    P = {**DEFAULT, **(P or {})}
    nt = (n + 127) // 128
    tk = tickets(n, nb, 1, grid)
    flags, waiters = {}, {}
    events = []  # (time, seq, worker)
    seq = [0]
    gens = {}
    next_ticket = [0]
    stats = dict(wait=0.0, busy=0.0)
    ready_time = {}

    def task(kind, i, j, part, nparts):
        yield ("delay", P["ticket"])
        if kind == D0:
            for cb in range(4):
                yield ("delay", P["d0"] / 4)
                yield ("set", ("pf", 0, cb))
            yield ("set", ("r", 0, 0))
            return
        pre = kind == PRE
        nk = i - 1 if pre else j
        k0, k1 = part * nk // nparts, (part + 1) * nk // nparts
        kt = P["kt"] if i < nt else P["kt_y"]
        piped = P["pipelined"] > 0 and kind == HEAD and part + 1 == nparts
        k1a = k1 - 1 if (piped and k1 > k0) else k1
        for k in range(k0, k1a):
            stalled = False
            for key in ((("r", i, k),) if pre else (("r", i, k), ("r", j, k))):
                if key not in flags:
                    stalled = True
                    yield ("wait", key)
            yield ("delay", kt + (P["refill"] if stalled else 0.0))
        if part > 0:
            key = ("pc", i, j, part - 1)
            if key not in flags:
                yield ("wait", key)
        if k1a > k0:
            yield ("delay", P["epi"])
        if pre or part + 1 < nparts:
            yield ("set", ("pc", i, j, part))
            return
        if piped:
            if k1 > k1a:
                for key in (("r", i, k1a), ("r", j, k1a)):
                    if key not in flags:
                        yield ("wait", key)
                yield ("delay", P["newest"])
            for cb in range(4):
                key = ("pf", j, cb)
                if key not in flags:
                    yield ("wait", key)
                yield ("delay", P["block"])
            yield ("delay", P["publish"])
            yield ("set", ("r", i, j))
            k = i
            if k >= 2 and ("pcall", k) not in flags:
                yield ("wait", ("pcall", k))
            yield ("delay", P["asm"])
            for cb in range(3):
                yield ("delay", P["fac_panel"])
                yield ("set", ("pf", k, cb))
            yield ("delay", P["fac_panel"])
            yield ("set", ("pf", k, 3))
            yield ("set", ("r", k, k))
            return
        if ("r", j, j) not in flags:
            yield ("wait", ("r", j, j))
        if kind == HEAD and "head_solve" in P:
            yield ("delay", P["head_solve"])
        else:
            yield ("delay", P["solve"] if i < nt else P["solve"] * 0.5)
        yield ("set", ("r", i, j))
        if kind == HEAD:
            k = i
            yield ("delay", P["syrk"])
            if k >= 2:
                key = ("pcall", k)
                if key not in flags:
                    yield ("wait", key)
            yield ("delay", P["asm"] + P["fac_flag"])
            yield ("set", ("r", k, k))
            yield ("delay", P["fac_rest"])

    pre_parts = {}

    def set_flag(key, now):
        flags[key] = now
        if key[0] == "r":
            ready_time[key[1:]] = now
        if key[0] == "pc" and key[1] == key[2]:  # a PRE part: HEAD waits for all of them
            k = key[1]
            pre_parts[k] = pre_parts.get(k, 0) + 1
            if pre_parts[k] == pre_total[k]:
                set_flag(("pcall", k), now)
        for w in waiters.pop(key, []):
            heapq.heappush(events, (now + P["poll"], seq[0], w))
            seq[0] += 1

    pre_total = {}
    for kind, b, i, j, part, nparts in tk:
        if kind == PRE:
            pre_total[i] = nparts
    for k in range(nt):
        if k not in pre_total and k >= 2:
            pre_total[k] = 0
    for w in range(grid):
        heapq.heappush(events, (0.0, seq[0], w))
        seq[0] += 1
    end = [0.0] * grid
    blocked_since = {}
    while events:
        now, _, w = heapq.heappop(events)
        if w in blocked_since:
            stats["wait"] += now - blocked_since.pop(w)
        g = gens.get(w)
        while True:
            if g is None:
                if next_ticket[0] >= len(tk):
                    end[w] = max(end[w], now)
                    break
                kind, b, i, j, part, nparts = tk[next_ticket[0]]
                next_ticket[0] += 1
                g = gens[w] = task(kind, i, j, part, nparts)
            try:
                op = next(g)
            except StopIteration:
                g = gens[w] = None
                continue
            if op[0] == "delay":
                heapq.heappush(events, (now + op[1], seq[0], w))
                seq[0] += 1
                break
            if op[0] == "wait":
                if op[1] in flags:
                    continue
                waiters.setdefault(op[1], []).append(w)
                blocked_since[w] = now
                break
            if op[0] == "set":
                set_flag(op[1], now)
    span = max(end)
    diag = [ready_time.get((k, k), 0.0) for k in range(nt)]
    res = dict(n=n, span_us=span, wait_per_cta=stats["wait"] / grid, end_min=min(end), end_median=sorted(end)[grid // 2],
               pace=[diag[k] - diag[k - 1] for k in range(3, nt)])
    return res


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8424
    P = {}
    for a in sys.argv[2:]:
        k, v = a.split("=")
        P[k] = float(v)
    r = simulate(n, P=P)
    print(f"n={n}: span {r['span_us']:.0f} us, flag waits {r['wait_per_cta']:.0f} us/CTA, CTA end min {r['end_min']:.0f} "
          f"median {r['end_median']:.0f}")
    print("pace:", " ".join(f"{v:.0f}" for v in r["pace"]))


if __name__ == "__main__":
    main()

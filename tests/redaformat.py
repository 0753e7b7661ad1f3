"""Test helper: parse a serialized REDA program with numpy and put its DFA in a
canonical form (breadth-first from the initial state over bytes 0..255), so
that two programs can be compared up to renumbering of states and classes."""

import numpy as np


def parse(prog: bytes):
    b = np.frombuffer(prog, dtype=np.uint8)
    assert bytes(b[:4]) == b"REDA"
    fmt, max_char, leader_len = int(b[12]), int(b[13]), int(b[14])
    n = int(np.frombuffer(prog[16:20], dtype=np.uint32)[0])
    initial_off = int(np.frombuffer(prog[20:24], dtype=np.uint32)[0])
    leader_off = int(np.frombuffer(prog[24:28], dtype=np.uint32)[0])
    equiv = b[32:288].astype(np.int64)
    leader = b[288:288 + leader_len].astype(np.int64)
    base = 288 + ((leader_len + 7) & ~7)
    row = max_char + 2
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[fmt]
    tab = np.frombuffer(prog[base:base + n * row * fmt], dtype=dt).astype(np.int64).reshape(n, row)
    top = 8 * fmt - 1
    head = tab[:, 0]
    result = head & ((1 << top) - 1)
    dead = head >> top
    nxt = tab[:, 1:] // row
    assert (tab[:, 1:] % row == 0).all()
    assert len(prog) == base + n * row * fmt
    return dict(fmt=fmt, classes=max_char + 1, n=n, init=initial_off // fmt // row,
                leader_state=leader_off // fmt // row, equiv=equiv, leader=leader,
                result=result, dead=dead, next=nxt)


def canonical(prog: bytes):
    """-> (rows, results, dead flags) with states renumbered breadth-first from
    the initial state, transitions expanded to raw bytes."""
    p = parse(prog)
    full = p["next"][:, p["equiv"]]  # n x 256 over raw bytes
    order, seen = [p["init"]], {p["init"]: 0}
    i = 0
    while i < len(order):
        for t in full[order[i]]:
            t = int(t)
            if t not in seen:
                seen[t] = len(order)
                order.append(t)
        i += 1
    remap = np.full(p["n"], -1, dtype=np.int64)
    for old, new in seen.items():
        remap[old] = new
    rows = remap[full[order]]
    return rows, p["result"][order], p["dead"][order], len(order)


def class_partition(prog: bytes):
    p = parse(prog)
    groups = {}
    for byte, c in enumerate(p["equiv"]):
        groups.setdefault(int(c), []).append(byte)
    return sorted(tuple(v) for v in groups.values())


def leader_bytesets(prog: bytes):
    p = parse(prog)
    return [tuple(np.nonzero(p["equiv"] == c)[0].tolist()) for c in p["leader"]]


def same_dfa(a: bytes, b: bytes):
    ra, rb = canonical(a), canonical(b)
    return (ra[3] == rb[3] and np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
            and np.array_equal(ra[2], rb[2]))


def equivalent(a: bytes, b: bytes, starts=False):
    """Do the two programs give the same result id and match end for every
    text and style?  Walk the product automaton from (initial, initial) over raw
    bytes: at every reachable pair the results must agree.  With starts=True
    'is this the initial state' must agree as well - that is what
    Outcome.start_ is derived from, and it is sensitive to whether a minimizer
    happened to merge the initial state with an equivalent twin."""
    pa, pb = parse(a), parse(b)
    fa = pa["next"][:, pa["equiv"]]
    fb = pb["next"][:, pb["equiv"]]
    ia, ib = pa["init"], pb["init"]
    seen = {(ia, ib)}
    todo = [(ia, ib)]
    while todo:
        x, y = todo.pop()
        if pa["result"][x] != pb["result"][y]:
            return False
        if starts and ((x == ia) != (y == ib)):
            return False
        for t in set(zip(fa[x].tolist(), fb[y].tolist())):
            if t not in seen:
                seen.add(t)
                todo.append(t)
    return True

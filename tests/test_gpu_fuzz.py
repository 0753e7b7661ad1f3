"""Seeded fuzz: random rule sets compiled by the product's own compiler, random
texts, every verb and style, replace() and StatefulMatcher - the CUDA path
against the oracle executing the same serialized program.  GPU only."""

import numpy as np
import pytest

from test_gpu_parity import pack

pytestmark = pytest.mark.gpu

VERBS = ["check", "match", "scan", "search"]
STYLES = ["instant", "first", "tangent", "last", "full"]
ATOMS = ["a", "b", "c", "d", "ab", "[a-c]", "[^a]", ".", "x", "\\d", "[0-9a-f]", "abc", "b+"]


def random_regex(rng, depth=0):
    r = rng.random_sample()
    if depth > 2 or r < 0.35:
        return ATOMS[rng.randint(len(ATOMS))]
    if r < 0.55:
        return random_regex(rng, depth + 1) + random_regex(rng, depth + 1)
    if r < 0.70:
        return "(" + random_regex(rng, depth + 1) + "|" + random_regex(rng, depth + 1) + ")"
    if r < 0.80:
        return "(" + random_regex(rng, depth + 1) + ")*"
    if r < 0.88:
        return "(" + random_regex(rng, depth + 1) + ")+"
    if r < 0.94:
        return "(" + random_regex(rng, depth + 1) + ")?"
    return "(" + random_regex(rng, depth + 1) + "){1,3}"


def random_texts(rng, n):
    alphabet = list(b"abcdx0123 \n\xff")
    out = [bytes(rng.choice(alphabet, size=rng.randint(0, 60)).astype(np.uint8)) for _ in range(n)]
    out += [bytes(rng.choice(alphabet[:5], size=300 + rng.randint(0, 40)).astype(np.uint8)) for _ in range(3)]
    return out + [b""]


@pytest.mark.parametrize("seed", [101, 202, 303])
def test_fuzz_random_programs(gpu_red, redlib, oracle_cls, seed):
    rng = np.random.RandomState(seed)
    done = 0
    while done < 50:
        rules = []
        for _ in range(rng.randint(1, 5)):
            flags = [0, 0, 1, 2, 4, 6, 7][rng.randint(7)]
            rules.append((random_regex(rng).encode(), int(rng.randint(1, 6)), flags))
        p = redlib.Parser()
        try:
            for pat, res, flags in rules:
                p.add(pat, res, flags)
            blob = redlib.compileToSerialized(p, [255, 1, 2, 4][rng.randint(4)])
        except redlib.RedExcept:
            continue  # e.g. too many states for a 1-byte format
        done += 1
        prog, orc = gpu_red.Program(blob), oracle_cls(blob)
        texts = random_texts(rng, 40)
        blob_t, starts, lens = pack(texts, lead=b"z" * int(rng.randint(0, 17)))
        for verb in VERBS:
            for sty in STYLES:
                res, ms, me = prog.match_batch_host(blob_t, starts, lens, verb, sty)
                eres, ems, eme = orc.batch(blob_t, starts, lens, verb, sty)
                assert np.array_equal(res, eres) and np.array_equal(ms, ems) and np.array_equal(me, eme), \
                    (seed, rules, verb, sty)
        sty = STYLES[rng.randint(5)]
        mx = [1, 2, 1 << 40][rng.randint(3)]
        out, ostarts, olens, counts = prog.replace_batch_host(blob_t, starts, lens, b"<>", mx, sty)
        for i, t in enumerate(texts):
            cnt, exp = orc.replace(t, b"<>", mx, sty)
            got = out[int(ostarts[i]):int(ostarts[i]) + int(olens[i])].tobytes()
            assert (int(counts[i]), got) == (cnt, exp), (seed, rules, sty, mx, t)
        st, res, hit, hend = prog.advance_batch_host(blob_t, starts, lens)
        for i, t in enumerate(texts):
            _, ores, per = orc.advance(oracle_cls.FRESH, t, per_byte=True)
            last = max([k for k, r in enumerate(per) if r > 0], default=-1)
            assert int(res[i]) == ores and (int(hit[i]), int(hend[i])) == ((per[last], last + 1) if last >= 0 else (0, 0)), \
                (seed, rules, t)
        # the same texts as one stream (no leader programs only take that path; others are refused or routed)
        whole = b"".join(texts)
        import torch
        dt = torch.frombuffer(bytearray(whole) + bytearray(64), dtype=torch.uint8).cuda()[:len(whole)]
        for sty in STYLES:
            assert tuple(prog.match_stream_device(dt, "match", sty)) == orc.run(whole, "match", sty, False), \
                (seed, rules, "stream", sty)
        prog.close()


@pytest.mark.parametrize("seed,hot_states", [(11, None), (12, "6"), (13, "3")])
def test_fuzz_profiled_paths(gpu_red, redlib, oracle_cls, monkeypatch, seed, hot_states):
    """Batches large enough to be profiled (the staged kernel's hot set then
    follows the text: with a cap of a few states lanes keep leaving it) and
    streams long enough to get their own image - random programs, check/match
    in the styles that kernel serves, and the stream in every style."""
    import torch
    if hot_states:
        monkeypatch.setenv("RED_B200_HOT_STATES", hot_states)
        monkeypatch.setenv("RED_B200_HOT_FORCE", "1")
        monkeypatch.setenv("RED_B200_STREAM_HOT", "1")
    rng = np.random.RandomState(seed)
    done = 0
    while done < 8:
        rules = []
        for _ in range(rng.randint(1, 5)):
            flags = [2, 2, 0, 6, 3][rng.randint(5)]
            rules.append((random_regex(rng).encode(), int(rng.randint(1, 6)), flags))
        p = redlib.Parser()
        try:
            for pat, res, flags in rules:
                p.add(pat, res, flags)
            blob = redlib.compileToSerialized(p, 255)
        except redlib.RedExcept:
            continue
        done += 1
        prog, orc = gpu_red.Program(blob), oracle_cls(blob)
        alphabet = list(b"abcdx0123 fe\xc3")
        texts = [bytes(rng.choice(alphabet, size=rng.randint(0, 150)).astype(np.uint8)) for _ in range(9000)]
        blob_t, starts, lens = pack(texts, lead=b"q" * int(rng.randint(0, 17)))
        for verb in ("match", "check"):
            for sty in ("instant", "last", "full"):
                res, _, me = prog.match_batch_host(blob_t, starts, lens, verb, sty, want_starts=False)
                eres, _, eme = orc.batch(blob_t, starts, lens, verb, sty)
                bad = np.nonzero((res != eres) | (me != eme))[0]
                assert bad.size == 0, (seed, rules, verb, sty, bad[:4], res[bad[:4]], eres[bad[:4]])
        whole = b"\n".join(texts) + bytes(rng.choice(alphabet, size=300000).astype(np.uint8))
        dt = torch.frombuffer(bytearray(whole) + bytearray(64), dtype=torch.uint8).cuda()[:len(whole)]
        for sty in STYLES:
            assert tuple(prog.match_stream_device(dt, "match", sty)) == orc.run(whole, "match", sty, False), \
                (seed, rules, "stream", sty)
        prog.close()

"""The oracle (oracle/red_oracle.c) pinned against the reference:
known answers from the reference's own unit tests and outputs of the compiled
reference (tests/golden, made by tools/gen_golden.py).  CPU only."""

import numpy as np
import pytest

from conftest import unb64, unz, rules_from_json

VERBS = ["check", "match", "scan", "search"]
STYLES = ["instant", "first", "tangent", "last", "full"]


def test_matcher_cpp_known_answers(kat, oracle_cls):
    n = 0
    for case in kat["matcher_cpp"]:
        o = oracle_cls(unz(case["program"]))
        for verb, style, text, exp in case["rows"]:
            got = o.run(unb64(text), verb, style)
            for k in range(3):
                if exp[k] is not None:
                    assert got[k] == exp[k], (case["rules"], verb, style, unb64(text), exp, got)
            n += 1
    assert n > 500


def test_omnibus_cpp_known_answers(kat, oracle_cls):
    n = 0
    for rec in kat["omnibus_cpp"]:
        if rec["program"] is None:
            assert rec["text"] is None  # the reference test expects a parse error
            continue
        o = oracle_cls(unz(rec["program"]))
        text = unb64(rec["text"])
        res = o.run(text, "check", "full")[0]
        oc = o.run(text, "match", "full")
        assert res == oc[0]
        assert (res == 1) == rec["match"], (unb64(rec["regex"]), text)
        n += 1
    assert n > 120


def test_reference_vectors(vectors, oracle_cls):
    n = 0
    for case in vectors["cases"]:
        o = oracle_cls(unz(case["program"]))
        assert o.info() == case["info"]
        texts = [unb64(t) for t in case["texts"]]
        for verb in VERBS:
            for sty in STYLES:
                exp = case["expected"][verb][sty]
                for t, e in zip(texts, exp):
                    assert list(o.run(t, verb, sty)) == e, (case["name"], verb, sty, t)
                    n += 1
    assert n > 20000


def test_leader_is_transparent_for_match_and_search(vectors, oracle_cls):
    # For match/search the leader only gates.  check and scan are observably
    # different with it (check starts from the post-leader state without
    # looking at its result; scan skips the bytes a leader comparison ate):
    # both quirks are pinned by the 'leader*' reference vectors.
    for case in vectors["cases"]:
        o = oracle_cls(unz(case["program"]))
        for t in [unb64(t) for t in case["texts"]]:
            for verb in ("match", "search"):
                for sty in STYLES:
                    assert o.run(t, verb, sty, True) == o.run(t, verb, sty, False)


def test_batch_equals_single(vectors, oracle_cls):
    case = next(c for c in vectors["cases"] if c["name"] == "union64")
    o = oracle_cls(unz(case["program"]))
    texts = [unb64(t) for t in case["texts"]]
    blob = b"\n".join(texts)
    starts = np.cumsum([0] + [len(t) + 1 for t in texts[:-1]]).astype(np.uint64)
    lens = np.array([len(t) for t in texts], dtype=np.uint32)
    for verb in VERBS:
        for sty in STYLES:
            res, so, eo = o.batch(np.frombuffer(blob, dtype=np.uint8), starts, lens, verb, sty)
            exp = case["expected"][verb][sty]
            assert [[int(a), int(b), int(c)] for a, b, c in zip(res, so, eo)] == exp


def test_header_validation(vectors, oracle_cls):
    prog = bytearray(unz(vectors["cases"][0]["program"]))
    assert oracle_cls.check_header(bytes(prog)) is None
    assert oracle_cls.check_header(bytes(prog[:100])) == "Serialized DFA: header too short"
    bad = bytearray(prog); bad[0] = ord("X")
    assert oracle_cls.check_header(bytes(bad)) == "Serialized DFA: bad magic number"
    bad = bytearray(prog); bad[4] = 2
    assert oracle_cls.check_header(bytes(bad)) == "Serialized DFA: unrecognized version"
    bad = bytearray(prog); bad[-1] ^= 1
    assert oracle_cls.check_header(bytes(bad)) == "serialized DFA: checksum mismatch"
    with pytest.raises(ValueError):
        oracle_cls(bytes(bad))


def test_oracle_matches_live_reference(vectors, oracle_cls, ref_cls):
    """Where the compiled reference is present: recompile and compare live."""
    rng = np.random.RandomState(5)
    for case in vectors["cases"][:20]:
        prog = ref_cls.compile(rules_from_json(case["rules"]), fmt=case["fmt"])
        assert prog == unz(case["program"])  # the reference is deterministic
        o, r = oracle_cls(prog), ref_cls(prog)
        for _ in range(30):
            t = bytes(rng.choice(list(b"abcdx019 FNewYork"), size=rng.randint(0, 30)).astype(np.uint8))
            for verb in VERBS:
                for sty in STYLES:
                    assert o.run(t, verb, sty) == r.run(t, verb, sty)


def test_match_all_and_collect_vectors(oracle_cls):
    import json, os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "ref_multi.json")) as f:
        doc = json.load(f)
    n = 0
    for case in doc["cases"]:
        o = oracle_cls(unz(case["program"]))
        for t, ma, co in zip(case["texts"], case["match_all"], case["collect"]):
            t = unb64(t)
            assert [list(x) for x in o.match_all(t)] == ma, (case["name"], t)
            assert [list(x) for x in o.collect(t)] == co, (case["name"], t)
            n += 1
    assert n > 100


def test_replace_and_stateful_vectors(extra_vectors, oracle_cls):
    """replace() (Matcher.h:643-707) and StatefulMatcher (Matcher.h:770-794)
    against the reference's outputs, plus the literal expectations of
    test/matcher.cpp:648-690 and :797-818."""
    n = 0
    for case in extra_vectors["cases"]:
        o = oracle_cls(unz(case["program"]))
        repls = [unb64(r) for r in case["repls"]]
        for t, rep, sf in zip(case["texts"], case["replace"], case["stateful"]):
            t = unb64(t)
            for key, (cnt, out) in rep.items():
                sty, mx, ri = key.split("/")
                assert o.replace(t, repls[int(ri)], int(mx), sty) == (cnt, unb64(out)), (case["name"], t, key)
                n += 1
            st, res, per = o.advance(oracle_cls.FRESH, t, per_byte=True)
            assert [o.advance(oracle_cls.FRESH, b"")[1]] + per == sf, (case["name"], t)
            # feeding in pieces carries the state
            st2, _ = o.advance(oracle_cls.FRESH, t[:len(t) // 2])
            assert o.advance(st2, t[len(t) // 2:]) == (st, res)
    assert n > 4000

    def prog_of(name):
        return oracle_cls(unz(next(c for c in extra_vectors["cases"] if c["name"] == name)["program"]))
    o = prog_of("matcher_cpp_replace")
    assert o.replace(b"fooac", b"bar", 9999, "last") == (1, b"foobar")
    assert o.replace(b"fooacz", b"bar", 9999, "last") == (1, b"foobarz")
    assert o.replace(b"xacyabbcz", b",", 9999, "tangent") == (2, b"x,y,z")
    o = prog_of("matcher_cpp_replaceStyles")
    assert o.replace(b"#123defg!", b"xyz", 1, "instant") == (1, b"#xyz23defg!")
    assert o.replace(b"#123defg!", b"xyz", 9999, "instant") == (3, b"#xyzxyzxyzdefg!")
    assert o.replace(b"#123defg!", b"xyz", 9999, "first") == (1, b"#xyzdefg!")
    assert o.replace(b"#123defg!", b"xyz", 9999, "tangent") == (1, b"#xyzefg!")
    assert o.replace(b"#123defg!", b"xyz", 9999, "last") == (1, b"#xyz!")
    assert o.replace(b"#123defg!", b"xyz", 9999, "full") == (0, b"#123defg!")
    assert o.replace(b"#123defg", b"xyz", 9999, "full") == (1, b"#xyz")
    o = prog_of("matcher_cpp_charByChar")
    assert o.advance(oracle_cls.FRESH, b"")[1] == 0
    assert o.advance(oracle_cls.FRESH, b"aleex", per_byte=True)[2] == [0, 0, 1, 1, 2]

#!/usr/bin/env python
"""Generate tests/golden/*.json from the reference itself.

Run here (where /root/reference exists), after `make -C oracle`:

    python tools/gen_golden.py

Two fixtures are written:

  reference_kat.json   known-answer tests transcribed from the reference's own
                       unit tests: test/matcher.cpp (hand-transcribed below as
                       data, every expectation re-checked against the compiled
                       reference before it is written) and the Rec table of
                       test/omnibus.cpp (parsed out of the source at generation
                       time).  Programs are stored as compiled by the reference.
  ref_vectors.json     outputs of the reference for rule sets x texts x all
                       verbs x all styles, including the flume-style sets of
                       one_b200/synth.py and random regexes.

Programs are zlib+base64; texts are base64 (they may hold any byte).
"""

import base64
import json
import os
import re
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.loader import (RefLib, RefError, Oracle, STYLES, VERBS, FMT_AUTO,  # noqa: E402
                           F_IGNORE_CASE, F_LOOSE_START, F_LOOSE_END,
                           LANG_RAW, LANG_AUTO, LANG_GLOB, LANG_EXACT)
from one_b200 import synth  # noqa: E402

REF_TEST_DIR = "/root/reference/quol/red/test"
OUT_DIR = os.path.join(ROOT, "tests", "golden")


def b64(b):
    return base64.b64encode(b).decode()


def zb64(b):
    return base64.b64encode(zlib.compress(b, 9)).decode()


def rule_json(r):
    p = r[0] if isinstance(r[0], bytes) else r[0].encode("latin-1")
    return [b64(p), r[1], r[2], r[3] if len(r) > 3 else LANG_RAW]


# ---- test/matcher.cpp, transcribed ---------------------------------------
# (rules, [(verb, style, text, (result, start, end)) ...]); start/end None =
# not asserted by the reference test.
N3 = [("[0-9]+", 1, 0), ("[0-9]+a", 2, 0), ("[0-9]+abcd", 3, 0)]


def _verify(style, rows):
    # matcher.cpp verifyX tests: check == match.result, search == match
    out = []
    for text, exp in rows:
        out.append(("check", style, text, (exp[0], None, None)))
        out.append(("match", style, text, exp))
        out.append(("search", style, text, exp))
    return out


def _search(style, rows):
    out = []
    for text, exp in rows:
        out.append(("scan", style, text, (exp[0], None, None)))
        out.append(("search", style, text, exp))
    return out


MATCHER_CPP = [
    # TEST_P(MatcherTest, case)
    ([("abc", 1, 0)], [("check", "full", "abc", (1, None, None)), ("check", "full", "aBc", (0, None, None)),
                       ("check", "full", "ABC", (0, None, None)), ("check", "full", "xyz", (0, None, None))]),
    ([("abc", 1, F_IGNORE_CASE)], [("check", "full", "abc", (1, None, None)), ("check", "full", "aBc", (1, None, None)),
                                   ("check", "full", "ABC", (1, None, None)), ("check", "full", "xyz", (0, None, None))]),
    # endmarks
    ([("abe", 1, 0), ("ace", 2, 0)], [("check", "full", "abe", (1, None, None)), ("check", "full", "ace", (2, None, None)),
                                      ("check", "full", "abc", (0, None, None)), ("check", "full", "age", (0, None, None))]),
    # start
    ([("a*", 1, 0)], [("search", "last", "123a", (1, 3, 4)), ("search", "last", "123az", (1, 4, 4))]),
    ([(".*[a-z]+", 1, 0)], [("search", "last", "123a", (1, 3, 4)), ("search", "last", "123az!", (1, 3, 5))]),
    # startEnd
    ([("[^a]*ab*c", 1, 0), ("[^d]*dummy", 2, 0)],
     [("match", "last", "abbc", (1, 0, 4)), ("match", "last", "xabbc", (1, 1, 5)),
      ("match", "last", "xabbcx", (1, 1, 5)), ("match", "last", "xyzabbcxyz", (1, 3, 7))]),
    # matchTangent
    ([("[0-9]+", 1, 0)], [("match", "instant", "0123456789abcdef", (1, None, 1)),
                          ("match", "first", "0123456789abcdef", (1, None, 10)),
                          ("match", "tangent", "0123456789abcdef", (1, None, 10))]),
    # matchLast
    ([("New", 1, F_LOOSE_START), ("New York", 2, F_LOOSE_START), ("York", 3, F_LOOSE_START)],
     [("match", "last", "I love New York.", (2, 7, 15))]),
    # verifyInstant .. verifyFull
    (N3, _verify("instant", [("1", (1, 0, 1)), ("123", (1, 0, 1)), ("123abcd", (1, 0, 1)), ("123abcde", (1, 0, 1))])),
    (N3, _verify("first", [("1", (1, 0, 1)), ("123", (1, 0, 3)), ("123abcd", (1, 0, 3)), ("123XYZ", (1, 0, 3))])),
    (N3, _verify("tangent", [("1", (1, 0, 1)), ("123", (1, 0, 3)), ("123abcd", (2, 0, 4)), ("123XYZ", (1, 0, 3))])),
    (N3, _verify("last", [("1", (1, 0, 1)), ("123", (1, 0, 3)), ("123abcd", (3, 0, 7)), ("123abcde", (3, 0, 7))])),
    (N3, _verify("full", [("1", (1, 0, 1)), ("123", (1, 0, 3)), ("123abcd", (3, 0, 7)), ("123abcde", (0, 0, 0))])),
    # searchInstant .. searchFull
    (N3, _search("instant", [(".,_1", (1, 3, 4)), (".,_123", (1, 3, 4)), (".,_123abcd", (1, 3, 4)), (".,_123abcde", (1, 3, 4))])),
    (N3, _search("first", [(".,_1", (1, 3, 4)), (".,_123", (1, 3, 6)), (".,_123abcd", (1, 3, 6)), (".,_123XYZ", (1, 3, 6))])),
    (N3, _search("tangent", [(".,_1", (1, 3, 4)), (".,_123", (1, 3, 6)), (".,_123abcd", (2, 3, 7)), (".,_123XYZ", (1, 3, 6))])),
    (N3, _search("last", [(".,_1", (1, 3, 4)), (".,_123", (1, 3, 6)), (".,_123abcd", (3, 3, 10)), (".,_123abcde", (3, 3, 10))])),
    (N3, _search("full", [(".,_1", (1, 3, 4)), (".,_123", (1, 3, 6)), (".,_123abcd", (3, 3, 10)), (".,_123abcde", (0, 0, 0))])),
    # check / match (addAuto)
    ([("ab*c", 1, 0, LANG_AUTO), ("ca*b", 2, 0, LANG_AUTO)],
     [("check", "full", "bca", (0, None, None)), ("check", "full", "bac", (1, None, None)),
      ("check", "full", "cab", (2, None, None)), ("match", "full", "bca", (0, None, 0)),
      ("match", "full", "bac", (1, None, 3)), ("match", "full", "cab", (2, None, 3))]),
]


def c_unescape(s):
    """Value of a C string literal body (the subset omnibus.cpp uses)."""
    out = bytearray()
    i = 0
    while i < len(s):
        c = s[i]
        if c != "\\":
            out.append(ord(c))
            i += 1
            continue
        i += 1
        c = s[i]
        simple = {"n": 10, "t": 9, "r": 13, "\\": 92, '"': 34, "0": 0, "a": 7, "b": 8, "v": 11, "f": 12}
        if c == "x":
            j = i + 1
            while j < len(s) and s[j] in "0123456789abcdefABCDEF":
                j += 1
            out.append(int(s[i + 1:j], 16) & 0xff)
            i = j
        else:
            out.append(simple[c])
            i += 1
    return bytes(out)


def parse_omnibus():
    src = open(os.path.join(REF_TEST_DIR, "omnibus.cpp")).read()
    body = src[src.index("Rec testRecs[] = {"):src.index("};", src.index("Rec testRecs[] = {"))]
    body = body.replace("\n    ", " ")
    recs = []
    lit = r'"((?:[^"\\]|\\.)*)"'
    for m in re.finditer(r"Rec\{" + lit + r",\s*(?:" + lit + r"|nullptr),\s*(true|false)\}", body):
        regex = c_unescape(m.group(1))
        text = None if m.group(2) is None else c_unescape(m.group(2))
        recs.append((regex, text, m.group(3) == "true"))
    return recs


def gen_kat():
    cases = []
    nchecked = 0
    for fmt in (FMT_AUTO, 1, 2, 4):
        for rules, rows in MATCHER_CPP:
            try:
                prog = RefLib.compile(rules, fmt=fmt)
            except RefError:
                continue  # a format too small for this DFA, as in the reference's tests
            ref = RefLib(prog)
            jrows = []
            for verb, style, text, exp in rows:
                got = ref.run(text.encode("latin-1"), verb, style)
                for k in range(3):
                    assert exp[k] is None or exp[k] == got[k], (rules, verb, style, text, exp, got)
                nchecked += 1
                jrows.append([verb, style, b64(text.encode("latin-1")), list(exp)])
            cases.append({"source": "test/matcher.cpp", "fmt": fmt, "rules": [rule_json(r) for r in rules],
                          "program": zb64(prog), "rows": jrows})
    # omnibus: addAuto(regex, 1, 0); check/match with styFull; expected (res == 1) == match
    omni = []
    for regex, text, want in parse_omnibus():
        rec = {"regex": b64(regex), "text": None if text is None else b64(text), "match": want}
        try:
            prog = RefLib.compile([(regex, 1, 0, LANG_AUTO)], max_states=200000)
            assert text is not None, regex
            got = RefLib(prog).run(text, "match", "full")
            assert (got[0] == 1) == want, (regex, text, got)
            nchecked += 1
            rec["program"] = zb64(prog)
        except RefError as ex:
            assert ex.code == -3 and text is None, (regex, ex)
            rec["program"] = None
        omni.append(rec)
    doc = {"generator": "tools/gen_golden.py", "matcher_cpp": cases, "omnibus_cpp": omni}
    with open(os.path.join(OUT_DIR, "reference_kat.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("reference_kat.json: %d matcher cases, %d omnibus records, %d expectations verified"
          % (len(cases), len(omni), nchecked))


def random_regex(rng, depth=0):
    atoms = ["a", "b", "c", "d", "ab", "[a-c]", "[^a]", ".", "x", "\\d", "[0-9a-f]"]
    r = rng.random_sample()
    if depth > 2 or r < 0.35:
        return atoms[rng.randint(len(atoms))]
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


def random_text(rng, n):
    alphabet = b"abcdx0123 \n\xff"
    return bytes(alphabet[i] for i in rng.randint(0, len(alphabet), n))


def sweep(ref, texts):
    exp = {}
    for verb in VERBS:
        exp[verb] = {}
        for sty in STYLES:
            exp[verb][sty] = [list(ref.run(t, verb, sty)) for t in texts]
    return exp


def gen_vectors():
    rng = np.random.RandomState(20261018)
    cases = []

    def add(name, rules, texts, fmt=FMT_AUTO):
        prog = RefLib.compile(rules, fmt=fmt)
        ref = RefLib(prog)
        cases.append({"name": name, "fmt": fmt, "rules": [rule_json(r) for r in rules],
                      "program": zb64(prog), "info": Oracle(prog).info(),
                      "texts": [b64(t) for t in texts], "expected": sweep(ref, texts)})

    lines = synth.make_lines(300, seed=99)
    edge = [b"", b"\n", b"a", b"Failed password for", b"xFailed password forx", b"\xff\x00\xfe"]
    add("single_auto", synth.rules_single(), lines[:150] + edge)
    add("union64", synth.rules_union64(), lines + edge)
    add("keywords300", synth.rules_keywords(300), [random_text(rng, 60) for _ in range(40)] +
        [r[0] for r in synth.rules_keywords(300)[:20]] + [b"zz" + r[0] + b"zz" for r in synth.rules_keywords(300)[20:40]])
    add("leader", [("aab", 1, 0)], [b"aaab", b"xaab", b"aab", b"aaaab", b"abaab", b"aabaab", b"", b"aa"])
    add("leader_class", [("foo[0-9]+bar", 7, 0)], [b"foo1bar", b"ffoo12bar", b"foofoo1bar", b"foo", b"fofoo9barx"])
    add("loose_both", [("ab+c", 1, F_LOOSE_START | F_LOOSE_END), ("xyz", 2, F_LOOSE_START | F_LOOSE_END)],
        [b"..abbc..", b"xyzabc", b"abcxyz", b"ab", b"", b"xxyzz"])
    add("newyork", [("New", 1, F_LOOSE_START), ("New York", 2, F_LOOSE_START), ("York", 3, F_LOOSE_START)],
        [b"I love New York.", b"New", b"York New", b"NewYork", b"new york", b"New Yor", b""])
    add("nullable", [("a*", 1, 0), ("(ab)*", 2, 0)], [b"", b"a", b"aaa", b"ab", b"abab", b"aba", b"b", b"ba"])
    add("initloop", [("(ab|c)*", 1, 0)], [b"abc", b"cab", b"ccc", b"abab", b"abx", b"", b"a"])
    add("glob", [("*.[ch]", 1, 0, LANG_GLOB), ("foo?bar*", 2, F_IGNORE_CASE, LANG_GLOB)],
        [b"x.c", b"x.h", b"x.o", b"fooXbar", b"FOO_BARRR", b"foobar", b".c"])
    add("exact", [("foo.*bar", 1, 0, LANG_EXACT), ("plain", 2, F_LOOSE_START | F_LOOSE_END, LANG_EXACT)],
        [b"foo.*bar", b"foobar", b"a plain text", b"plain", b"plai"])
    add("bytes", [("\\x00+\\xff", 1, 0), ("[\\x80-\\xff]+", 2, 0)],
        [b"\x00\xff", b"\x00\x00\xff", b"\x80\x90\xff", b"\xff", b"\x00", b"a\x00\xff"])
    for fmt in (1, 2, 4):
        add("n3_fmt%d" % fmt, [("[0-9]+", 1, 0), ("[0-9]+a", 2, 0), ("[0-9]+abcd", 3, 0)],
            [b"1", b"123", b"123abcd", b"123abcde", b".,_123abcd", b"", b"abcd"], fmt=fmt)
    n = 0
    while n < 40:
        k = rng.randint(1, 4)
        rules = []
        for i in range(k):
            flags = [0, 0, F_LOOSE_START, F_IGNORE_CASE, F_LOOSE_START | F_LOOSE_END][rng.randint(5)]
            rules.append((random_regex(rng), i + 1, flags))
        try:
            prog = RefLib.compile(rules, max_states=5000)
        except RefError:
            continue
        if Oracle(prog).info()["state_cnt"] > 400:
            continue
        texts = [random_text(rng, rng.randint(0, 40)) for _ in range(24)]
        add("random%02d" % n, rules, texts)
        n += 1
    doc = {"generator": "tools/gen_golden.py", "cases": cases}
    path = os.path.join(OUT_DIR, "ref_vectors.json")
    with open(path, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("ref_vectors.json: %d cases, %d bytes" % (len(cases), os.path.getsize(path)))


def gen_multi():
    """matchAll() (Matcher.h:127) and Red::collect() (Red.h:109) outputs."""
    rng = np.random.RandomState(424242)
    cases = []

    def add(name, rules, texts):
        prog = RefLib.compile(rules)
        ref = RefLib(prog)
        cases.append({"name": name, "rules": [rule_json(r) for r in rules], "program": zb64(prog),
                      "texts": [b64(t) for t in texts],
                      "match_all": [ref.multi(t, "match_all") for t in texts],
                      "collect": [ref.multi(t, "collect") for t in texts]})

    # the two matchAll tests of test/matcher.cpp
    add("matcher_cpp_matchAll", [("0", 1, 0), ("0123", 2, 0), ("[0-2]+", 3, 0), ("[3-9]+", 4, 0), ("012345", 5, 0)],
        [b"0123456789", b"", b"0", b"9", b"01", b"33012"])
    add("matcher_cpp_matchAllLoose", [("a+", 1, F_LOOSE_START), ("b+", 2, F_LOOSE_START)],
        [b".aa..b.bb..abba.", b"", b"ab", b"....", b"bbbb"])
    cities = [("tokyo", 1, F_IGNORE_CASE), ("new", 2, F_IGNORE_CASE), ("new york", 3, F_IGNORE_CASE),
              ("york", 4, F_IGNORE_CASE), ("sao paulo", 5, F_IGNORE_CASE), ("ab+", 6, 0)]
    texts = [bytes(rng.choice(list(b"new york tokyoabb NEW sao paulo"), size=rng.randint(0, 60)).astype(np.uint8))
             for _ in range(60)]
    texts += [b"I love New York and Tokyo, not Sao Paulo.", b"newnewnew yorkyork", b"abbbabab"]
    add("cities", cities, texts)
    add("union64_lines", synth.rules_union64(), synth.make_lines(80, seed=5))
    add("leader", [("foo[0-9]+", 1, 0)], [b"foo1", b"xfoo12foo3", b"foofoo9", b"", b"fo"])
    doc = {"generator": "tools/gen_golden.py", "cases": cases}
    path = os.path.join(OUT_DIR, "ref_multi.json")
    with open(path, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("ref_multi.json: %d cases, %d bytes" % (len(cases), os.path.getsize(path)))


def gen_extra():
    """replace() (Matcher.h:100-125) and StatefulMatcher (Matcher.h:770) outputs."""
    rng = np.random.RandomState(77)
    cases = []

    def add(name, rules, texts, repls=(b"X", b"", b"<long replacement>")):
        prog = RefLib.compile(rules)
        ref = RefLib(prog)
        rep = []
        for t in texts:
            per = {}
            for sty in ("instant", "first", "tangent", "last", "full"):
                for mx in (1, 2, 9999):
                    for ri, r in enumerate(repls):
                        cnt, out = ref.replace(t, r, mx, sty)
                        per["%s/%d/%d" % (sty, mx, ri)] = [cnt, b64(out)]
            rep.append(per)
        cases.append({"name": name, "rules": [rule_json(r) for r in rules], "program": zb64(prog),
                      "texts": [b64(t) for t in texts], "repls": [b64(r) for r in repls],
                      "replace": rep, "stateful": [ref.stateful(t) for t in texts]})

    # the expectations of test/matcher.cpp (replace, replaceStyles, charByChar)
    add("matcher_cpp_replace", [("ab*c", 1, 0)], [b"fooac", b"fooacz", b"xacyabbcz", b"", b"abc", b"ab"],
        repls=(b"bar", b",", b""))
    add("matcher_cpp_replaceStyles", [("[0-9]+", 1, 0), ("[0-9]+d", 2, 0), ("[0-9]+defg", 3, 0)],
        [b"#123defg!", b"#123defg", b"1d2d3", b"", b"defg"], repls=(b"xyz", b"", b"0"))
    add("matcher_cpp_charByChar", [("ale+", 1, 0), ("ale*x", 2, 0)], [b"aleex", b"alx", b"aleexx", b"", b"a"])
    cities = [("tokyo", 1, F_IGNORE_CASE), ("new", 2, F_IGNORE_CASE), ("new york", 3, F_IGNORE_CASE),
              ("york", 4, F_IGNORE_CASE), ("sao paulo", 5, F_IGNORE_CASE), ("ab+", 6, 0)]
    texts = [bytes(rng.choice(list(b"new york tokyoabb NEW sao paulo"), size=rng.randint(0, 60)).astype(np.uint8))
             for _ in range(40)]
    texts += [b"I love New York and Tokyo, not Sao Paulo.", b"newnewnew yorkyork", b"abbbabab"]
    add("cities", cities, texts)
    add("leader", [("foo[0-9]+", 1, 0)], [b"foo1", b"xfoo12foo3", b"foofoo9", b"", b"fo", b"foo", b"fofoo77x"])
    add("loose", [("a+", 1, F_LOOSE_START), ("b+", 2, F_LOOSE_START | F_LOOSE_END)],
        [b".aa..b.bb..abba.", b"", b"ab", b"....", b"bbbb", b"a.b"])
    add("nullable", [("a*", 1, 0), ("a*b", 2, 0)], [b"", b"aaa", b"aab", b"baab", b"xaax"])
    add("union64_lines", synth.rules_union64(), synth.make_lines(40, seed=9), repls=(b"#", b""))
    doc = {"generator": "tools/gen_golden.py", "cases": cases}
    path = os.path.join(OUT_DIR, "ref_extra.json")
    with open(path, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("ref_extra.json: %d cases, %d bytes" % (len(cases), os.path.getsize(path)))


if __name__ == "__main__":
    os.makedirs(OUT_DIR, exist_ok=True)
    if "--extra-only" in sys.argv:
        gen_extra()
        sys.exit(0)
    if "--multi-only" not in sys.argv:
        gen_kat()
        gen_vectors()
    gen_multi()
    gen_extra()

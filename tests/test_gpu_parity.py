"""Parity of the CUDA path against the oracle, through the C ABI.  GPU only.

Bit-exact: per line result id, match start and match end, for every verb and
style, on the reference's golden vectors and on seeded synthetic corpora."""

import numpy as np
import pytest

from conftest import unb64, unz

pytestmark = pytest.mark.gpu

VERBS = ["check", "match", "scan", "search"]
STYLES = ["instant", "first", "tangent", "last", "full"]


def pack(texts, sep=b"\n", lead=b""):
    blob = lead + sep.join(texts) + sep
    lens = np.array([len(t) for t in texts], dtype=np.uint32)
    starts = np.zeros(len(texts), dtype=np.uint64)
    pos = len(lead)
    for i, t in enumerate(texts):
        starts[i] = pos
        pos += len(t) + len(sep)
    return np.frombuffer(blob, dtype=np.uint8), starts, lens


def triple(res, ms, me):
    return [[int(a), int(b), int(c)] for a, b, c in zip(res, ms, me)]


def test_known_answers_from_reference_tests(gpu_red, kat):
    n = 0
    for case in kat["matcher_cpp"]:
        prog = gpu_red.Program(unz(case["program"]))
        for verb, style, text, exp in case["rows"]:
            got = prog.run(unb64(text), verb, style)
            for k in range(3):
                if exp[k] is not None:
                    assert got[k] == exp[k], (case["rules"], verb, style, unb64(text), exp, got)
            n += 1
        prog.close()
    assert n > 500


def test_omnibus_known_answers(gpu_red, kat):
    for rec in kat["omnibus_cpp"]:
        if rec["program"] is None:
            continue
        prog = gpu_red.Program(unz(rec["program"]))
        text = unb64(rec["text"])
        res = prog.check(text, "full")
        assert res == prog.match(text, "full").result
        assert (res == 1) == rec["match"], (unb64(rec["regex"]), text)
        prog.close()


def test_reference_vectors_all_verbs_styles(gpu_red, vectors):
    for case in vectors["cases"]:
        prog = gpu_red.Program(unz(case["program"]))
        texts = [unb64(t) for t in case["texts"]]
        # odd lead-in so that line starts hit every alignment
        text, starts, lens = pack(texts, lead=b"#" * 3)
        for verb in VERBS:
            for sty in STYLES:
                res, ms, me = prog.match_batch_host(text, starts, lens, verb, sty)
                assert triple(res, ms, me) == case["expected"][verb][sty], (case["name"], verb, sty)
        prog.close()


@pytest.mark.parametrize("name", ["single_auto", "union64", "keywords300"])
def test_synthetic_corpus_vs_oracle(gpu_red, vectors, oracle_cls, name):
    from one_b200 import synth
    case = next(c for c in vectors["cases"] if c["name"] == name)
    blob = unz(case["program"])
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    text, starts, lens = synth.make_corpus(60000, seed=42)
    for verb in ("check", "match"):
        for sty in STYLES:
            res, ms, me = prog.match_batch_host(text, starts, lens, verb, sty)
            eres, ems, eme = orc.batch(text, starts, lens, verb, sty)
            assert np.array_equal(res, eres), (verb, sty)
            assert np.array_equal(ms, ems) and np.array_equal(me, eme), (verb, sty)
    if name != "keywords300":  # random keywords do not occur in syslog text
        assert (eres > 0).any()
    # sliding verbs on a slice (quadratic in the reference, too)
    sl = slice(0, 4000)
    for verb in ("scan", "search"):
        for sty in ("instant", "last"):
            res, ms, me = prog.match_batch_host(text, starts[sl], lens[sl], verb, sty)
            eres, ems, eme = orc.batch(text, starts[sl], lens[sl], verb, sty)
            assert np.array_equal(res, eres) and np.array_equal(ms, ems) and np.array_equal(me, eme)
    prog.close()


def test_kernel_variants_agree(gpu_red, vectors, oracle_cls, monkeypatch):
    """Persistent-lane schedule (variant 3), first kernel (variant 1), length-sorted
    tiles forced for every style (variant 4, 8 and 3 tiles per group) and the plain
    tile schedule (variant 2): same answers."""
    from one_b200 import synth
    text, starts, lens = synth.make_corpus(30000, seed=13)
    for name in ("union64", "single_auto", "leader_class"):
        blob = unz(next(c for c in vectors["cases"] if c["name"] == name)["program"])
        orc = oracle_cls(blob)
        for variant, tiles in (("3", "0"), ("1", "0"), ("4", "8"), ("4", "3"), ("2", "0")):
            monkeypatch.setenv("RED_B200_VARIANT", variant)
            monkeypatch.setenv("RED_B200_LINES_PER_LANE", tiles)
            prog = gpu_red.Program(blob)
            monkeypatch.delenv("RED_B200_VARIANT")
            monkeypatch.delenv("RED_B200_LINES_PER_LANE")
            for verb in ("check", "match"):
                for sty in STYLES:
                    res, ms, me = prog.match_batch_host(text, starts, lens, verb, sty)
                    eres, ems, eme = orc.batch(text, starts, lens, verb, sty)
                    assert np.array_equal(res, eres) and np.array_equal(ms, ems) and np.array_equal(me, eme), \
                        (name, variant, verb, sty)
            prog.close()


def test_simple_walk_variant_agrees(gpu_red, vectors, oracle_cls, monkeypatch):
    """The per-byte-decision kernel (A/B baseline) gives the same answers."""
    from one_b200 import synth
    case = next(c for c in vectors["cases"] if c["name"] == "union64")
    blob = unz(case["program"])
    monkeypatch.setenv("RED_B200_SIMPLE_WALK", "1")
    monkeypatch.setenv("RED_B200_VARIANT", "1")
    prog = gpu_red.Program(blob)
    monkeypatch.delenv("RED_B200_SIMPLE_WALK")
    monkeypatch.delenv("RED_B200_VARIANT")
    orc = oracle_cls(blob)
    text, starts, lens = synth.make_corpus(20000, seed=3)
    for sty in STYLES:
        res, ms, me = prog.match_batch_host(text, starts, lens, "match", sty)
        eres, ems, eme = orc.batch(text, starts, lens, "match", sty)
        assert np.array_equal(res, eres) and np.array_equal(ms, ems) and np.array_equal(me, eme)
    prog.close()


def test_sliding_kernels_agree(gpu_red, redlib, vectors, oracle_cls, monkeypatch):
    """scan/search: the warp-per-line kernel (32 start positions at a time) and the
    lane-per-line kernel (RED_B200_SIMPLE_WALK=1; also what scan() uses on a program
    with a leader) against the oracle, on tight and loose programs."""
    from one_b200 import synth
    lines = synth.make_lines(6000, seed=23) + [b"", b"x", b"Failed", b"a" * 700 + b"Failed password" + b"b" * 300]
    text, starts, lens = synth.pack_lines(lines)
    p = redlib.Parser()
    for i, ph in enumerate(synth.PHRASES):
        p.addAs(synth.LANG_RAW, ph, i + 1, 0)
    tight = redlib.compileToSerialized(p)
    blobs = [("tight64", tight)] + [(n, unz(next(c for c in vectors["cases"] if c["name"] == n)["program"]))
                                    for n in ("single_auto", "leader_class", "newyork", "nullable")]
    for name, blob in blobs:
        orc = oracle_cls(blob)
        for simple in ("0", "1"):
            monkeypatch.setenv("RED_B200_SIMPLE_WALK", simple)
            prog = gpu_red.Program(blob)
            monkeypatch.delenv("RED_B200_SIMPLE_WALK")
            for verb in ("scan", "search"):
                for sty in STYLES:
                    res, ms, me = prog.match_batch_host(text, starts, lens, verb, sty)
                    eres, ems, eme = orc.batch(text, starts, lens, verb, sty)
                    assert np.array_equal(res, eres) and np.array_equal(ms, ems) and np.array_equal(me, eme), \
                        (name, simple, verb, sty)
            prog.close()


def test_table_layouts_agree(gpu_red, vectors, oracle_cls, monkeypatch):
    """Small DFAs use byte-indexed rows (one load per byte); the class-compressed
    layout and the v1 tile kernel must give the same answers."""
    from one_b200 import synth
    text, starts, lens = synth.make_corpus(20000, seed=5)
    for name in ("single_auto", "n3_fmt1", "newyork", "leader_class"):
        case = next(c for c in vectors["cases"] if c["name"] == name)
        blob = unz(case["program"])
        orc = oracle_cls(blob)
        progs = [gpu_red.Program(blob)]
        assert progs[0].info["row_stride"] == 256
        monkeypatch.setenv("RED_B200_NO_BYTE_ROWS", "1")
        progs.append(gpu_red.Program(blob))
        assert progs[1].info["row_stride"] < 256
        monkeypatch.setenv("RED_B200_VARIANT", "1")
        progs.append(gpu_red.Program(blob))
        monkeypatch.delenv("RED_B200_NO_BYTE_ROWS")
        progs.append(gpu_red.Program(blob))
        monkeypatch.delenv("RED_B200_VARIANT")
        for verb in ("check", "match"):
            for sty in STYLES:
                eres, ems, eme = orc.batch(text, starts, lens, verb, sty)
                for pr in progs:
                    res, ms, me = pr.match_batch_host(text, starts, lens, verb, sty)
                    assert np.array_equal(res, eres) and np.array_equal(ms, ems) and np.array_equal(me, eme), (name, verb, sty)
        for pr in progs:
            pr.close()


def test_global_table_path_agrees(gpu_red, vectors, oracle_cls, monkeypatch):
    """Force the L2/global-table kernels on a DFA that would fit shared memory."""
    from one_b200 import synth
    case = next(c for c in vectors["cases"] if c["name"] == "union64")
    blob = unz(case["program"])
    monkeypatch.setenv("RED_B200_SMEM_TABLE_LIMIT", "1024")
    prog = gpu_red.Program(blob)
    monkeypatch.delenv("RED_B200_SMEM_TABLE_LIMIT")
    assert prog.info["table_mode"] == 2
    orc = oracle_cls(blob)
    text, starts, lens = synth.make_corpus(20000, seed=4)
    for verb in VERBS:
        for sty in ("first", "last"):
            n = 20000 if verb in ("check", "match") else 2000
            res, ms, me = prog.match_batch_host(text, starts[:n], lens[:n], verb, sty)
            eres, ems, eme = orc.batch(text, starts[:n], lens[:n], verb, sty)
            assert np.array_equal(res, eres) and np.array_equal(ms, ems) and np.array_equal(me, eme)
    prog.close()


def test_edge_cases_ragged_and_empty(gpu_red, vectors, oracle_cls):
    case = next(c for c in vectors["cases"] if c["name"] == "union64")
    blob = unz(case["program"])
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    rng = np.random.RandomState(11)
    phrase = b"Failed password for"
    texts = [b"", b"", b"x", phrase, phrase[:-1], b" " * 15 + phrase, b" " * 16 + phrase, b" " * 17 + phrase]
    for ln in list(range(0, 70)) + [255, 256, 257, 1000, 4097, 70000]:
        body = bytes(rng.choice(list(b"abcdefghij klmnop=:"), size=ln).astype(np.uint8))
        texts.append(body)
        if ln > 25:
            k = rng.randint(0, ln - 20)
            texts.append(body[:k] + phrase + body[k + len(phrase):])
    # no separator at all: lines abut, starts unaligned
    text, starts, lens = pack(texts, sep=b"")
    for verb in ("check", "match"):
        for sty in STYLES:
            res, ms, me = prog.match_batch_host(text, starts, lens, verb, sty)
            eres, ems, eme = orc.batch(text, starts, lens, verb, sty)
            assert np.array_equal(res, eres) and np.array_equal(ms, ems) and np.array_equal(me, eme), (verb, sty)
    # zero lines
    res, ms, me = prog.match_batch_host(text, starts[:0], lens[:0], "match", "last")
    assert len(res) == 0
    # lines out of text order and overlapping views
    perm = rng.permutation(len(texts))
    res, ms, me = prog.match_batch_host(text, starts[perm], lens[perm], "match", "last")
    eres, ems, eme = orc.batch(text, starts[perm], lens[perm], "match", "last")
    assert np.array_equal(res, eres) and np.array_equal(me, eme)
    prog.close()


def test_host_batch_chunk_planning(gpu_red, vectors, oracle_cls, monkeypatch):
    """The host pipeline plans its chunks optimistically (lines in text order,
    span = first start .. last end, count by bisection) and a kernel checks on
    the device that every line is inside; chunks that fail are redone with the
    exact plan.  Lines in text order, reversed, shuffled, with views that
    overlap, many small chunks - and RED_B200_EXACT_PLAN=1 - all agree with the
    oracle, for every verb."""
    from one_b200 import synth
    case = next(c for c in vectors["cases"] if c["name"] == "union64")
    blob = unz(case["program"])
    orc = oracle_cls(blob)
    text, starts, lens = synth.make_corpus(60000, seed=77)
    rng = np.random.RandomState(5)
    orders = {"in order": np.arange(len(starts)), "reversed": np.arange(len(starts))[::-1].copy(),
              "shuffled": rng.permutation(len(starts)),
              "mostly ordered": np.arange(len(starts))}
    orders["mostly ordered"][[100, 40000]] = orders["mostly ordered"][[40000, 100]]
    monkeypatch.setenv("RED_B200_HOST_CHUNK_BYTES", str(256 << 10))
    for exact in (False, True):
        if exact:
            monkeypatch.setenv("RED_B200_EXACT_PLAN", "1")
        prog = gpu_red.Program(blob)
        for name, perm in orders.items():
            st, ln = starts[perm].copy(), lens[perm].copy()
            if name == "shuffled":
                ln[::7] = np.minimum(ln[::7] + 40, (len(text) - st[::7]).astype(np.uint32))  # overlapping views
            for verb, sty in (("match", "last"), ("match", "instant"), ("check", "full"), ("search", "first")):
                res, ms, me = prog.match_batch_host(text, st, ln, verb, sty)
                eres, ems, eme = orc.batch(text, st, ln, verb, sty)
                assert np.array_equal(res, eres) and np.array_equal(ms, ems) and np.array_equal(me, eme), \
                    (exact, name, verb, sty)
        prog.close()


def test_device_resident_batch_and_compaction(gpu_red, vectors, oracle_cls):
    import torch
    from one_b200 import synth
    case = next(c for c in vectors["cases"] if c["name"] == "union64")
    blob = unz(case["program"])
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    text, starts, lens = synth.make_corpus(50000, seed=8)
    dt = torch.from_numpy(text.copy()).cuda()
    ds = torch.from_numpy(starts.astype(np.int64)).cuda()
    dl = torch.from_numpy(lens.astype(np.int32)).cuda()
    res, ms, me = prog.match_batch_device(dt, ds, dl, "match", "instant", want_starts=True, want_ends=True)
    torch.cuda.synchronize()
    eres, ems, eme = orc.batch(text, starts, lens, "match", "instant")
    assert np.array_equal(res.cpu().numpy(), eres)
    assert np.array_equal(ms.cpu().numpy().astype(np.uint32), ems)
    assert np.array_equal(me.cpu().numpy().astype(np.uint32), eme)
    idx = gpu_red.compact_matches_device(res)
    assert np.array_equal(idx.cpu().numpy(), np.nonzero(eres > 0)[0])
    # the line index built on the GPU equals the one the generator made
    gs, gl = gpu_red.split_lines_device(dt)
    assert np.array_equal(gs.cpu().numpy().astype(np.uint64), starts)
    assert np.array_equal(gl.cpu().numpy().astype(np.uint32), lens)
    prog.close()


def test_split_lines_edge_cases(gpu_red):
    import torch
    for raw in [b"", b"\n", b"a", b"a\n", b"\n\n", b"ab\ncd", b"ab\n\ncd\n", b"x" * 40000 + b"\n" + b"y" * 33]:
        t = torch.frombuffer(bytearray(raw + b"\0" * 16), dtype=torch.uint8)[:len(raw)].cuda() if raw else \
            torch.empty(0, dtype=torch.uint8, device="cuda")
        gs, gl = gpu_red.split_lines_device(t)
        parts = raw.split(b"\n")
        if raw.endswith(b"\n") or raw == b"":
            parts = parts[:-1]
        exp_lens = [len(p) for p in parts]
        exp_starts = list(np.cumsum([0] + [n + 1 for n in exp_lens[:-1]])) if parts else []
        assert gl.cpu().tolist() == exp_lens, raw[:20]
        assert gs.cpu().tolist() == exp_starts, raw[:20]


def test_style_and_verb_errors(gpu_red, vectors):
    prog = gpu_red.Program(unz(vectors["cases"][0]["program"]))
    with pytest.raises(gpu_red.RedExceptExec):
        prog.run(b"abc", 1, 0)   # styInvalid
    with pytest.raises(gpu_red.RedExceptExec):
        prog.run(b"abc", 1, 6)
    with pytest.raises(gpu_red.RedExceptExec):
        prog.run(b"abc", 9, 4)
    prog.close()


def test_grep_host_blob(gpu_red, vectors, oracle_cls, monkeypatch):
    """Log blob in, matching lines out: index + match + compaction on the GPU."""
    from one_b200 import synth
    case = next(c for c in vectors["cases"] if c["name"] == "union64")
    blob = unz(case["program"])
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    text, starts, lens = synth.make_corpus(70000, seed=31)
    for piece in (None, "65536"):  # default pieces, and many small ones
        if piece:
            monkeypatch.setenv("RED_B200_HOST_CHUNK_BYTES", piece)
        for sty in ("instant", "last"):
            recs, n_lines = prog.grep_host(text, b"\n", sty)
            eres, ems, eme = orc.batch(text, starts, lens, "match", sty)
            hit = np.nonzero(eres > 0)[0]
            assert n_lines == len(starts)
            assert np.array_equal(recs["line_start"], starts[hit])
            assert np.array_equal(recs["line_len"], lens[hit])
            assert np.array_equal(recs["result"], eres[hit])
            assert np.array_equal(recs["match_start"], ems[hit])
            assert np.array_equal(recs["match_end"], eme[hit])
    monkeypatch.delenv("RED_B200_HOST_CHUNK_BYTES")
    # no trailing delimiter, empty lines, nothing matching, capacity too small
    recs, n = prog.grep_host(b"a\n\nFailed password for x\n\nQuarantined", b"\n", "instant")
    assert n == 5 and recs["line_start"].tolist() == [3, 26] and recs["result"].tolist() == [5, 15]
    recs, n = prog.grep_host(b"nothing here\nnor here\n", b"\n", "instant")
    assert n == 2 and len(recs) == 0
    with pytest.raises(gpu_red.RedExceptLimit):
        prog.grep_host(text, b"\n", "instant", capacity=10)
    prog.close()


def test_grep_host_short_lines(gpu_red, vectors, oracle_cls, monkeypatch):
    """Far more lines per piece than the index holds (1-3 byte lines, runs of
    blank lines between ordinary ones): such a piece is taken again in smaller
    pieces, in place, and the matches still come out in text order."""
    from one_b200 import synth
    case = next(c for c in vectors["cases"] if c["name"] == "union64")
    blob = unz(case["program"])
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    rng = np.random.RandomState(6)
    real = synth.make_lines(3000, seed=2)
    lines = []
    for i in range(200000):
        lines.append(real[i % 3000] if i % 50 == 0 else bytes(rng.randint(97, 123, rng.randint(0, 4)).astype(np.uint8)))
    text, starts, lens = synth.pack_lines(lines)
    eres, ems, eme = orc.batch(text, starts, lens, "match", "instant")
    hit = np.nonzero(eres > 0)[0]
    assert hit.size > 100
    for piece in ("65536", "1048576", None):
        if piece:
            monkeypatch.setenv("RED_B200_HOST_CHUNK_BYTES", piece)
        else:
            monkeypatch.delenv("RED_B200_HOST_CHUNK_BYTES")
        recs, n_lines = prog.grep_host(text, b"\n", "instant")
        assert n_lines == len(starts), piece
        assert np.array_equal(recs["line_start"], starts[hit]), piece
        assert np.array_equal(recs["result"], eres[hit]) and np.array_equal(recs["match_end"], eme[hit]), piece
    prog.close()


def test_match_all_and_collect(gpu_red, oracle_cls):
    """Per line all matches: matchAll() and Red::collect() against the reference's
    outputs (fixture) and against the oracle on a seeded corpus."""
    import json, os
    from conftest import GOLDEN
    from one_b200 import synth
    with open(os.path.join(GOLDEN, "ref_multi.json")) as f:
        doc = json.load(f)
    for case in doc["cases"]:
        blob = unz(case["program"])
        prog = gpu_red.Program(blob)
        texts = [unb64(t) for t in case["texts"]]
        text, starts, lens = pack(texts, lead=b"##")
        for which in ("match_all", "collect"):
            counts, spans = prog.multi_batch_host(text, starts, lens, which, max_per_line=24)
            for i, exp in enumerate(case[which]):
                assert counts[i] == len(exp), (case["name"], which, texts[i])
                got = [[int(s["result"]), int(s["start"]), int(s["end"])] for s in spans[i][:counts[i]]]
                assert got == exp, (case["name"], which, texts[i])
        # the single-text conveniences
        assert [list(o) for o in prog.collect(texts[0])] == case["collect"][0]
        assert [list(o) for o in prog.matchAll(texts[0])] == case["match_all"][0]
        prog.close()
    # capped output: counts say how many there were
    case = next(c for c in doc["cases"] if c["name"] == "cities")
    blob = unz(case["program"])
    prog, orc = gpu_red.Program(blob), oracle_cls(blob)
    lines = [b" ".join([b"tokyo", b"new york", b"x"] * 5)] * 100
    text, starts, lens = pack(lines)
    counts, spans = prog.multi_batch_host(text, starts, lens, "collect", max_per_line=3)
    exp = orc.collect(lines[0])
    assert (counts == len(exp)).all() and len(exp) == 10
    assert [[int(s["result"]), int(s["start"]), int(s["end"])] for s in spans[7]] == [list(x) for x in exp[:3]]
    prog.close()

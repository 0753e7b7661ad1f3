"""replace() and StatefulMatcher on the GPU against the reference's outputs
(tests/golden/ref_extra.json) and against the oracle on seeded corpora.
GPU only, through the C ABI."""

import numpy as np
import pytest

from conftest import unb64, unz
from test_gpu_parity import pack

pytestmark = pytest.mark.gpu

STYLES = ["instant", "first", "tangent", "last", "full"]


def test_replace_golden_vectors(gpu_red, extra_vectors):
    n = 0
    for case in extra_vectors["cases"]:
        prog = gpu_red.Program(unz(case["program"]))
        texts = [unb64(t) for t in case["texts"]]
        repls = [unb64(r) for r in case["repls"]]
        blob, starts, lens = pack(texts, lead=b"xx")
        for sty in STYLES:
            for mx in (1, 2, 9999):
                for ri, repl in enumerate(repls):
                    out, ostarts, olens, counts = prog.replace_batch_host(blob, starts, lens, repl, mx, sty)
                    for i, per in enumerate(case["replace"]):
                        cnt, exp = per["%s/%d/%d" % (sty, mx, ri)]
                        got = out[int(ostarts[i]):int(ostarts[i]) + int(olens[i])].tobytes()
                        assert (int(counts[i]), got) == (cnt, unb64(exp)), (case["name"], texts[i], sty, mx, repl)
                        n += 1
        # the single-text entry point
        for t, per in list(zip(texts, case["replace"]))[:6]:
            cnt, exp = per["last/9999/0"]
            assert prog.replace(t, repls[0], 9999, "last") == (cnt, unb64(exp))
        prog.close()
    assert n > 4000


def test_replace_corpus_vs_oracle(gpu_red, vectors, oracle_cls):
    from one_b200 import synth
    lines = synth.make_lines(3000, seed=31) + [b"", b"x"]
    blob, starts, lens = pack(lines)
    for name in ("union64", "single_auto", "newyork", "n3_fmt2"):
        case = next(c for c in vectors["cases"] if c["name"] == name)
        blob_p = unz(case["program"])
        prog, orc = gpu_red.Program(blob_p), oracle_cls(blob_p)
        for sty, mx, repl, delim in (("last", 1 << 40, b"<R>", 10), ("instant", 2, b"", -1),
                                     ("tangent", 1, b"0123456789abcdef0123", 0), ("full", 3, b"#", -1),
                                     ("first", 9999, b"ab", 10)):
            out, ostarts, olens, counts = prog.replace_batch_host(blob, starts, lens, repl, mx, sty, delim)
            pos = 0
            for i, ln in enumerate(lines):
                cnt, exp = orc.replace(ln, repl, mx, sty)
                assert int(ostarts[i]) == pos, (name, sty, i)
                got = out[pos:pos + int(olens[i])].tobytes()
                assert (int(counts[i]), got) == (cnt, exp), (name, sty, i, ln)
                pos += len(exp)
                if delim >= 0:
                    assert out[pos] == delim
                    pos += 1
            assert pos == len(out)
        prog.close()


def test_replace_lane_per_line_kernel_agrees(gpu_red, extra_vectors, monkeypatch):
    """The first replace kernel (one lane per line), kept for A/B behind
    RED_B200_REPLACE_LANE_PER_LINE: same answers as the warp-per-line default."""
    monkeypatch.setenv("RED_B200_REPLACE_LANE_PER_LINE", "1")
    for case in extra_vectors["cases"]:
        prog = gpu_red.Program(unz(case["program"]))
        texts = [unb64(t) for t in case["texts"]]
        repls = [unb64(r) for r in case["repls"]]
        blob, starts, lens = pack(texts)
        for sty in STYLES:
            out, ostarts, olens, counts = prog.replace_batch_host(blob, starts, lens, repls[0], 9999, sty)
            for i, per in enumerate(case["replace"]):
                cnt, exp = per["%s/9999/0" % sty]
                got = out[int(ostarts[i]):int(ostarts[i]) + int(olens[i])].tobytes()
                assert (int(counts[i]), got) == (cnt, unb64(exp)), (case["name"], texts[i], sty)
        prog.close()


def test_stateful_golden_vectors(gpu_red, extra_vectors):
    for case in extra_vectors["cases"]:
        prog = gpu_red.Program(unz(case["program"]))
        for t, sf in zip(case["texts"][:12], case["stateful"]):
            t = unb64(t)
            sm = gpu_red.StatefulMatcher(prog)
            got = [sm.result()] + [sm.advance(b) for b in t]   # byte by byte, as the reference test does
            assert got == sf, (case["name"], t)
            assert sm.result() == sf[-1]
        prog.close()


def test_stateful_segments_vs_oracle(gpu_red, vectors, oracle_cls):
    """Streams fed in ragged segments: state carried across calls, result()
    after each segment, last accepting position inside the segment."""
    from one_b200 import synth
    rng = np.random.RandomState(12)
    lines = synth.make_lines(400, seed=41)
    for name in ("union64", "single_auto", "newyork", "nullable", "initloop"):
        case = next(c for c in vectors["cases"] if c["name"] == name)
        blob_p = unz(case["program"])
        prog, orc = gpu_red.Program(blob_p), oracle_cls(blob_p)
        states = None
        ostates = [oracle_cls.FRESH] * len(lines)
        offs = [0] * len(lines)
        for rnd in range(5):
            segs = []
            for i, ln in enumerate(lines):
                k = len(ln) - offs[i] if rnd == 4 else int(rng.randint(0, 40))
                segs.append(ln[offs[i]:offs[i] + k])
                offs[i] += len(segs[-1])
            blob, starts, lens = pack(segs)
            states, res, hit, hend = prog.advance_batch_host(blob, starts, lens, states)
            for i, sg in enumerate(segs):
                ostates[i], ores, per = orc.advance(ostates[i], sg, per_byte=True)
                assert int(res[i]) == ores, (name, rnd, i)
                last = max([k for k, r in enumerate(per) if r > 0], default=-1)
                exp_hit = (per[last], last + 1) if last >= 0 else (0, 0)
                assert (int(hit[i]), int(hend[i])) == exp_hit, (name, rnd, i, sg)
        prog.close()

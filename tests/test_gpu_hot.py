"""The staged batch kernel (kernels_hot.cu: hot image in shared memory, text
through a ring of bulk-copied slots) against the oracle, through the C ABI.
GPU only.  It serves check/match without start positions for the styles
Instant, Last and Full; every case here asks for exactly that, so that this
kernel - not the one it falls back to - is what answers."""

import numpy as np
import pytest

from conftest import unb64, unz

pytestmark = pytest.mark.gpu

STYLES = ["instant", "last", "full"]
KNOBS = ["RED_B200_HOT", "RED_B200_HOT_STATES", "RED_B200_HOT_PROFILE", "RED_B200_HOT_WARPS",
         "RED_B200_HOT_SLOTS", "RED_B200_HOT_BATCH", "RED_B200_HOT_SLOT_BYTES", "RED_B200_HOT_PRODUCERS", "RED_B200_HOT_DEFER", "RED_B200_HOT_ASCII", "RED_B200_HOT_FORCE"]


def program(gpu_red, monkeypatch, blob, **knobs):
    for k in KNOBS:
        monkeypatch.delenv(k, raising=False)
    knobs.setdefault("force", 1)  # this kernel, also where the launcher would rather take another
    for k, v in knobs.items():
        monkeypatch.setenv("RED_B200_HOT_" + k.upper() if k != "hot" else "RED_B200_HOT", str(v))
    prog = gpu_red.Program(blob)
    for k in KNOBS:
        monkeypatch.delenv(k, raising=False)
    return prog


def agree(prog, orc, text, starts, lens, styles=STYLES, verbs=("match", "check"), what=""):
    for verb in verbs:
        for sty in styles:
            res, _, me = prog.match_batch_host(text, starts, lens, verb, sty, want_starts=False)
            eres, _, eme = orc.batch(text, starts, lens, verb, sty)
            bad = np.nonzero((res != eres) | (me != eme))[0]
            assert bad.size == 0, (what, verb, sty, bad[:5], res[bad[:5]], eres[bad[:5]], me[bad[:5]], eme[bad[:5]])


def blob_of(vectors, name):
    return unz(next(c for c in vectors["cases"] if c["name"] == name)["program"])


def test_reference_vectors(gpu_red, vectors, monkeypatch):
    """Recorded outputs of the reference, every rule set of the fixture."""
    from test_gpu_parity import pack
    for case in vectors["cases"]:
        prog = program(gpu_red, monkeypatch, unz(case["program"]))
        texts = [unb64(t) for t in case["texts"]]
        text, starts, lens = pack(texts, lead=b"#" * 3)
        for verb in ("check", "match"):
            for sty in STYLES:
                res, _, me = prog.match_batch_host(text, starts, lens, verb, sty, want_starts=False)
                exp = case["expected"][verb][sty]
                assert [int(r) for r in res] == [e[0] for e in exp], (case["name"], verb, sty)
                assert [int(e) for e in me] == [e[2] for e in exp], (case["name"], verb, sty)
        prog.close()


@pytest.mark.parametrize("name", ["single_auto", "union64", "keywords300"])
def test_synthetic_corpus(gpu_red, vectors, oracle_cls, monkeypatch, name):
    from one_b200 import synth
    blob = blob_of(vectors, name)
    orc = oracle_cls(blob)
    text, starts, lens = synth.make_corpus(60000, seed=42)
    prog = program(gpu_red, monkeypatch, blob)
    agree(prog, orc, text, starts, lens, what=name)
    assert prog.hot_launch_count == 6  # 2 verbs x 3 styles: this kernel answered, not a fallback
    prog.close()


@pytest.mark.parametrize("states", [1, 3, 16, 100])
def test_small_hot_set_takes_the_full_table(gpu_red, vectors, oracle_cls, monkeypatch, states):
    """With only a few hot states most lanes leave the hot set and come back."""
    from one_b200 import synth
    text, starts, lens = synth.make_corpus(20000, seed=5)
    for name in ("union64", "keywords300", "single_auto"):
        blob = blob_of(vectors, name)
        orc = oracle_cls(blob)
        # defer=1: the careful passes of a tile run together at its end, and a
        # lane that left the hot set walks the rest of its line alone
        for profile, defer in ((0, -1), (1, 1), (0, 1)):
            prog = program(gpu_red, monkeypatch, blob, states=states, profile=profile, defer=defer)
            agree(prog, orc, text, starts, lens, what=(name, states, profile, defer))
            prog.close()


@pytest.mark.parametrize("knobs", [dict(warps=8, producers=1), dict(warps=31, batch=128, slot_bytes=16384),
                                   dict(warps=16, batch=64, producers=4), dict(slots=8, batch=32, warps=4),
                                   dict(slot_bytes=4096, producers=2), dict(slot_bytes=1024, warps=12),
                                   dict(slots=2, producers=4), dict(batch=224, producers=1),
                                   dict(defer=-1), dict(defer=1, batch=96), dict(ascii=-1), dict(ascii=1, profile=0)])
def test_launch_shapes(gpu_red, vectors, oracle_cls, monkeypatch, knobs):
    """Other ring geometries; slots smaller than a batch's text make the producer
    split batches and gather line by line."""
    from one_b200 import synth
    text, starts, lens = synth.make_corpus(30000, seed=9)
    blob = blob_of(vectors, "union64")
    orc = oracle_cls(blob)
    prog = program(gpu_red, monkeypatch, blob, **knobs)
    agree(prog, orc, text, starts, lens, verbs=("match",), what=knobs)
    prog.close()


def test_ragged_batches(gpu_red, vectors, oracle_cls, monkeypatch):
    """Empty lines, one line, counts around the tile and batch sizes, every alignment."""
    from one_b200 import synth
    blob = blob_of(vectors, "union64")
    orc = oracle_cls(blob)
    prog = program(gpu_red, monkeypatch, blob)
    lines = synth.make_lines(700, seed=3)
    for i in range(0, 700, 7):
        lines[i] = b""
    lines[5] = b"x"
    lines[6] = b"Failed password for root"
    for lead in (0, 1, 15, 16, 17):
        text, starts, lens = synth.pack_lines(lines)
        text = np.concatenate([np.full(lead, 35, dtype=np.uint8), text])
        starts = starts + np.uint64(lead)
        for n in (1, 2, 31, 32, 33, 255, 256, 257, 700):
            agree(prog, orc, text, starts[:n], lens[:n], verbs=("match",), what=(lead, n))
    # no separators at all: the lines touch
    text, starts, lens = synth.pack_lines(lines, newline=False)
    agree(prog, orc, text, starts, lens, verbs=("match",), what="touching")
    prog.close()


def test_views_that_are_not_back_to_back(gpu_red, vectors, oracle_cls, monkeypatch):
    """Shuffled, reversed, overlapping and sparse views of a text: gathered line by line."""
    from one_b200 import synth
    blob = blob_of(vectors, "union64")
    orc = oracle_cls(blob)
    text, starts, lens = synth.make_corpus(20000, seed=11)
    rng = np.random.RandomState(1)
    prog = program(gpu_red, monkeypatch, blob)
    perm = rng.permutation(len(starts))
    agree(prog, orc, text, starts[perm], lens[perm], verbs=("match",), what="shuffled")
    agree(prog, orc, text, starts[::-1].copy(), lens[::-1].copy(), verbs=("match",), what="reversed")
    # overlapping windows of 40..300 bytes at random offsets
    n = 15000
    ostarts = rng.randint(0, text.size - 400, n).astype(np.uint64)
    olens = rng.randint(0, 300, n).astype(np.uint32)
    agree(prog, orc, text, ostarts, olens, verbs=("match",), what="overlapping")
    # every 50th line only: spans far wider than a slot
    agree(prog, orc, text, starts[::50].copy(), lens[::50].copy(), verbs=("match",), what="sparse")
    prog.close()


def test_long_lines(gpu_red, vectors, oracle_cls, monkeypatch):
    """Lines of a few hundred bytes (fewer per batch), and lines longer than a
    slot (walked from global memory) between short ones."""
    from one_b200 import synth
    blob = blob_of(vectors, "union64")
    orc = oracle_cls(blob)
    base = synth.make_lines(6000, seed=21)
    rng = np.random.RandomState(2)
    long_lines = []
    for i in range(0, 6000, 3):
        long_lines.append(b" ".join(base[i:i + 3]))
    text, starts, lens = synth.pack_lines(long_lines)
    prog = program(gpu_red, monkeypatch, blob)
    agree(prog, orc, text, starts, lens, what="300-byte lines")
    mixed = []
    for i in range(1500):
        if i % 97 == 0:
            k = rng.randint(200, 1200)
            mixed.append(b" | ".join(base[j] for j in rng.randint(0, 6000, k)))  # 20 .. 130 KB
        else:
            mixed.append(base[i])
    text, starts, lens = synth.pack_lines(mixed)
    assert lens.max() > 70000
    agree(prog, orc, text, starts, lens, what="mixed with very long lines")
    prog.close()


def test_device_resident_batch(gpu_red, vectors, oracle_cls, monkeypatch):
    """red_b200_match_batch_device on torch tensors, text base at an odd address."""
    import torch
    from one_b200 import synth
    blob = blob_of(vectors, "union64")
    orc = oracle_cls(blob)
    text, starts, lens = synth.make_corpus(50000, seed=17)
    prog = program(gpu_red, monkeypatch, blob)
    for shift in (0, 5):
        buf = torch.zeros(text.size + 64, dtype=torch.uint8, device="cuda")
        buf[shift:shift + text.size] = torch.from_numpy(text.copy()).cuda()
        dt = buf[shift:]
        ds = torch.from_numpy(starts.astype(np.int64)).cuda()
        dl = torch.from_numpy(lens.astype(np.int32)).cuda()
        for sty in STYLES:
            res, _, me = prog.match_batch_device(dt, ds, dl, "match", sty)
            torch.cuda.synchronize()
            eres, _, eme = orc.batch(text, starts, lens, "match", sty)
            assert np.array_equal(res.cpu().numpy(), eres), (shift, sty)
            assert np.array_equal(me.cpu().numpy().view(np.uint32), eme), (shift, sty)
    prog.close()


def test_bytes_above_127(gpu_red, vectors, oracle_cls, monkeypatch):
    """Text that is not ASCII: with rows of 128 columns (forced here, and chosen
    or not by the profile) a byte >= 128 takes the lane to the full table."""
    from one_b200 import synth
    rng = np.random.RandomState(4)
    lines = synth.make_lines(20000, seed=8)
    for i in range(0, len(lines), 3):
        ln = bytearray(lines[i])
        for _ in range(rng.randint(1, 6)):
            ln[rng.randint(0, len(ln))] = rng.randint(128, 256)
        lines[i] = bytes(ln)
    lines[7] = bytes(rng.randint(128, 256, 300).astype(np.uint8))
    text, starts, lens = synth.pack_lines(lines)
    for name in ("union64", "single_auto"):
        blob = blob_of(vectors, name)
        orc = oracle_cls(blob)
        for knobs in (dict(ascii=1), dict(ascii=1, defer=1), dict(), dict(ascii=1, profile=0, states=20)):
            prog = program(gpu_red, monkeypatch, blob, **knobs)
            agree(prog, orc, text, starts, lens, what=(name, knobs))
            prog.close()

import sys, os, json, base64, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from one_b200 import red, synth
from oracle.loader import Oracle
kat = json.load(open("tests/golden/reference_kat.json"))
case = kat["matcher_cpp"][0]
blob = zlib.decompress(base64.b64decode(case["program"]))
prog = red.Program(blob); orc = Oracle(blob)
print(prog.info, orc.info())
for verb in ("check","match"):
    for sty in ("full","last","instant"):
        for t in (b"abc", b"abcd", b"ab", b""):
            print(verb, sty, t, "gpu", tuple(prog.run(t, verb, sty)), "oracle", orc.run(t, verb, sty))
import torch
vec = json.load(open("tests/golden/ref_vectors.json"))
case = next(c for c in vec["cases"] if c["name"] == "union64")
blob = zlib.decompress(base64.b64decode(case["program"]))
prog = red.Program(blob); orc = Oracle(blob)
text, starts, lens = synth.make_corpus(2000, seed=1)
eres, ems, eme = orc.batch(text, starts, lens, "match", "instant")
for ws in (False, True):
    res, ms, me = prog.match_batch_host(text, starts, lens, "match", "instant", want_starts=ws)
    print("host want_starts", ws, "res ok", np.array_equal(res, eres), "me ok", np.array_equal(me, eme), "ms ok", ms is None or np.array_equal(ms, ems))
    bad = np.nonzero(res != eres)[0][:5]
    for i in bad: print("  line", i, bytes(text[int(starts[i]):int(starts[i])+int(lens[i])]), res[i], eres[i], me[i], eme[i])
    if ms is not None:
        bad = np.nonzero(ms != ems)[0][:5]
        for i in bad: print("  ms line", i, ms[i], ems[i], res[i], me[i])
    dt = torch.from_numpy(text.copy()).cuda(); ds = torch.from_numpy(starts.astype(np.int64)).cuda(); dl = torch.from_numpy(lens.astype(np.int32)).cuda()
    r2, m2, e2 = prog.match_batch_device(dt, ds, dl, "match", "instant", want_starts=ws)
    torch.cuda.synchronize()
    print("dev want_starts", ws, "res ok", np.array_equal(r2.cpu().numpy(), eres), "me ok", np.array_equal(e2.cpu().numpy().astype(np.uint32), eme), "ms ok", m2 is None or np.array_equal(m2.cpu().numpy().astype(np.uint32), ems))

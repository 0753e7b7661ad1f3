"""Hot-row mirror size sweep for a global-table program (2000 keywords): GB/s per size (GPU box)."""
import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from one_b200 import red, synth
dev = torch.device("cuda", 0)
pool = bench.host_pool(1234)
text, starts, lens, n = bench.device_corpus(pool, 5_000_000, dev)
nbytes = int(lens.sum().item())
res = torch.empty(n, dtype=torch.int32, device=dev); ends = torch.empty(n, dtype=torch.int32, device=dev)
p = red.Parser()
for pat, r, f, lang in synth.rules_keywords(2000):
    p.addAs(lang, pat, r, f)
blob = red.compileToSerialized(p)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for hot in ("0", "16384", "32768", "65536", "98304", "131072", "163840", "207000"):
    os.environ["RED_B200_HOT_ROWS_BYTES"] = hot
    prog = red.Program(blob)
    out = []
    for sty in ("instant", "last"):
        ms = timed(lambda: prog.match_batch_device(text, starts, lens, "match", sty, out=(res, None, ends)))
        out.append("%s %.0f GB/s" % (sty, nbytes / ms / 1e6))
    print("hot rows %7s B:" % hot, ", ".join(out), flush=True)
    prog.close()

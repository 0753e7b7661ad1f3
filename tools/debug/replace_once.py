import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import bench
from one_b200 import red, synth
dev = torch.device("cuda", 0)
pool = bench.host_pool(1234)
text, starts, lens, n = bench.device_corpus(pool, 2_000_000, dev)
p = red.Parser()
for i, ph in enumerate(synth.PHRASES):
    p.addAs(synth.LANG_RAW, ph, i + 1, 0)
prog = red.Program(red.compileToSerialized(p))
for _ in range(2):
    prog.replace_batch_device(text, starts, lens, b"<X>", 1 << 40, "last", 10)
torch.cuda.synchronize()

import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from one_b200 import red, synth
from oracle.loader import Oracle

blob = bench.load_program_blob("union64")
orc = Oracle(blob)
base = synth.make_lines(6000, seed=21)
rng = np.random.RandomState(2)
mixed = []
for i in range(1500):
    if i % 97 == 0:
        k = rng.randint(300, 420)
        mixed.append(b" | ".join(base[j] for j in rng.randint(0, 6000, k)))
    else:
        mixed.append(base[i])
text, starts, lens = synth.pack_lines(mixed)
for cfg in sys.argv[1:] or [""]:
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        os.environ["RED_B200_" + k] = v
    prog = red.Program(blob)
    for rep in range(4):
        for verb in ("match", "check"):
            for sty in ("instant", "last", "full"):
                res, _, me = prog.match_batch_host(text, starts, lens, verb, sty, want_starts=False)
                eres, _, eme = orc.batch(text, starts, lens, verb, sty)
                bad = np.nonzero((res != eres) | (me != eme))[0]
                if bad.size:
                    print(cfg, rep, verb, sty, "bad", bad.size, [(int(i), int(lens[i]), int(res[i]), int(eres[i]), int(me[i]), int(eme[i])) for i in bad[:6]], flush=True)
    print(cfg, "done", prog.hot_launch_count, flush=True)
    prog.close()

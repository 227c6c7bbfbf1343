"""Does a larger frame budget (all 64 utterances in ONE frame-phase group) change results or speed?
Compares item waveforms bit for bit against the default budget and prints step time and device memory."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokorox_b200.synth import ensure_weights, synth_batch  # noqa: E402
import torch
from kokorox_b200.onn import B200Koko

m = B200Koko.new(ensure_weights())
m.set_option("precision", 1)
toks, styles, speeds = synth_batch(64, 510)
res = {}
for mf in (49152, 100000):
    m.set_option("max_frames", mf)
    outs = m.infer_batch(toks, styles, speeds)
    res[mf] = [o.copy() for o in outs]
    del outs
    m.stage(toks, styles, speeds)
    for _ in range(2):
        m.run_staged()
    ts = []
    for _ in range(3):
        m.run_staged()
        ts.append(m.get_stat("gpu_us") / 1e3)
    free, total = torch.cuda.mem_get_info()
    print(f"max_frames={mf}: step {min(ts):.1f} / {sorted(ts)[1]:.1f} ms, launches {m.get_stat('launches')}, "
          f"device memory in use {(total - free) / 2**30:.1f} GiB", flush=True)
same = all(np.array_equal(a, b) for a, b in zip(res[49152], res[100000]))
print("bit-identical across budgets:", same, "finite:", all(np.isfinite(a).all() for a in res[100000]))

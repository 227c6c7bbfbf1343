"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel family of both configurations at sizes that
finish in seconds under the tool.

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokorox_b200.onn import B200Koko  # noqa: E402
from kokorox_b200.synth import ensure_weights, synth_case, synth_voice_table  # noqa: E402


def main():
    m = B200Koko.new(ensure_weights())
    cases = [synth_case(n, 10 + n, 11 + n) for n in (5, 70, 33)]
    for precision in (1, 0):
        m.set_option("precision", precision)
        for _ in range(3):                      # eager, graph capture, graph replay
            y = m.infer_one(cases[1][0], cases[1][1], 1.0)
        outs = m.infer_batch([c[0] for c in cases], [c[1] for c in cases], [1.0, 0.9, 1.2])
        print("precision", precision, "samples", [len(o) for o in outs], "finite", all(np.isfinite(o).all() for o in outs), len(y))
    m.set_option("precision", 1)
    m.set_option("fork_max_batch", 0)           # the non-forked path with a batch that spans two frame groups
    m.set_option("max_frames", 150)
    outs = m.infer_batch([c[0] for c in cases], [c[1] for c in cases], [1.0, 0.9, 1.2])
    pcm = m.infer_batch_pcm16([c[0] for c in cases], [c[1] for c in cases], [1.0, 0.9, 1.2])
    tab = synth_voice_table(3)
    m.load_voices({f"v{i}": tab[i] for i in range(3)})
    v = m.infer_batch_voices([c[0] for c in cases], ["v0", "v1.4+v2.5", "v2"], [1.0, 1.0, 1.0])
    t = [m.submit(c[0], c[1], 1.0) for c in cases]
    r = [m.wait(x) for x in t]
    print("ok", len(outs), len(pcm), len(v), len(r))
    m.close()


if __name__ == "__main__":
    main()

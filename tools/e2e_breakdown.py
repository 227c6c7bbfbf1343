"""Where the end-to-end (host buffers in, host waveform out) time goes: stage / run / fetch, B=64 x 510."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokorox_b200.synth import ensure_weights, synth_batch  # noqa: E402


def main():
    from kokorox_b200.onn import B200Koko
    m = B200Koko.new(ensure_weights())
    m.set_option("precision", 1)
    toks, styles, speeds = synth_batch(64, 510)
    for _ in range(2):
        m.infer_batch(toks, styles, speeds)
    for it in range(3):
        t0 = time.perf_counter()
        m.stage(toks, styles, speeds)
        t1 = time.perf_counter()
        n, _ = m.run_staged()
        t2 = time.perf_counter()
        out, soff, dur = m.fetch_staged(n)
        t3 = time.perf_counter()
        outs = m.infer_batch(toks, styles, speeds)
        t4 = time.perf_counter()
        print(f"stage {1e3*(t1-t0):.1f} ms  run {1e3*(t2-t1):.1f} ms  fetch(pageable dst) {1e3*(t3-t2):.1f} ms  |  infer_batch {1e3*(t4-t3):.1f} ms", flush=True)
    m.close()


if __name__ == "__main__":
    main()

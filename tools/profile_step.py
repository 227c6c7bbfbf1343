"""One small, fixed workload for ncu: B utterances of 510 tokens, two staged runs (warm-up + one).

    python tools/profile_step.py [--batch 8] [--precision 1] [--runs 2]
Prints the launch count per run (use it as ncu -s/-c) and the per-kernel event-timed table.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokorox_b200.synth import ensure_weights, synth_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--tokens", type=int, default=510)
    ap.add_argument("--precision", type=int, default=1)
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--max-frames", type=int, default=0, help="frame budget per frame-phase group (0 = library default)")
    ap.add_argument("--set", action="append", default=[], metavar="OPTION=VALUE", help="session option (kkx_set_option), repeatable")
    a = ap.parse_args()
    from kokorox_b200.onn import B200Koko
    m = B200Koko.new(ensure_weights())
    m.set_option("precision", a.precision)
    if a.max_frames:
        m.set_option("max_frames", a.max_frames)
    for kv in a.set:
        k, v = kv.split("=")
        m.set_option(k, int(v))
    toks, styles, speeds = synth_batch(a.batch, a.tokens)
    m.stage(toks, styles, speeds)
    m.profile_enable(True)
    for _ in range(a.runs):
        total, launches = m.run_staged()
    p = m.profile()
    kern = p["kernels"]
    tot = sum(v[1] for v in kern.values())
    print(json.dumps({"launches_per_run": launches, "audio_s": total / 24000.0, "gpu_ms": p["gpu_us"] / 1e3,
                      "conv_tflops_alg": p["conv_flops"] / 1e12}))
    for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:16s} n={v[0]:4d} total {v[1]/1e3:9.3f} ms  {100*v[1]/tot:5.1f}%")
    m.close()


if __name__ == "__main__":
    main()

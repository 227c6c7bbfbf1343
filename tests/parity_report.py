"""Stage-by-stage parity report: CUDA path (through the C ABI) vs the CPU oracle.

Run on the GPU box:  python tests/parity_report.py [--tokens 50] [--seed 0] [--precision 0]
Prints one line per stage (max-abs error, relative L2) and the duration / waveform verdicts.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.conftest import ensure_weights, make_noise, synth_case  # noqa: E402

STAGES = ["bert", "d_en", "d", "dur_lstm", "dur_logits", "dur_float", "t_en", "idx", "shared_lstm", "F0", "N",
          "dec.encode", "dec.decode.0", "dec.decode.1", "dec.decode.2", "dec.decode.3", "har_source", "har",
          "gen.x_source.0", "gen.ups.0", "gen.stage.0", "gen.x_source.1", "gen.ups.1", "gen.stage.1", "conv_post"]


def compare(name, ref, got):
    ref = np.asarray(ref, dtype=np.float64)
    if got is None:
        return f"{name:18s} MISSING"
    got = np.asarray(got, dtype=np.float64)
    if ref.shape != got.shape:
        return f"{name:18s} SHAPE ref {ref.shape} got {got.shape}"
    d = np.abs(ref - got)
    rel = np.sqrt((d ** 2).sum() / max((ref ** 2).sum(), 1e-30))
    bad = int(np.argmax(d))
    return (f"{name:18s} {str(ref.shape):14s} max_abs {d.max():.3e}  rel_l2 {rel:.3e}  ref_rms {np.sqrt((ref**2).mean()):.3e}"
            f"  nan {int(np.isnan(got).sum())}  argmax {np.unravel_index(bad, ref.shape)}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=50)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--precision", type=int, default=0)
    ap.add_argument("--teacher", action="store_true", help="teacher-force pred_dur/F0/N from the oracle")
    a = ap.parse_args()
    from kokorox_b200.onn import B200Koko
    from kokorox_b200.weightfile import read_weights
    from oracle.kokoro_ref import KokoroOracle
    wp = ensure_weights()
    ids, style = synth_case(a.tokens, a.seed, 100 + a.seed)
    noise = make_noise(50 * len(ids) if a.tokens <= 60 else 12 * len(ids))
    t0 = time.time()
    o = KokoroOracle(read_weights(wp))
    ref = o.forward(ids, style, 1.0, noise=noise, stages=True)
    print(f"oracle: T={ref['T']} in {time.time()-t0:.1f}s")
    m = B200Koko.new(wp)
    m.set_option("precision", a.precision)
    m.debug_enable(True)
    m.set_noise(noise)
    if a.teacher:
        m.set_inject("pred_dur", ref["pred_dur"])
        m.set_inject("F0", ref["stages"]["F0"])
        m.set_inject("N", ref["stages"]["N"])
    t0 = time.time()
    outs, durs = m.infer_batch([ids], [style], [1.0], return_durations=True)
    print(f"cuda: {len(outs[0])} samples in {time.time()-t0:.2f}s, launches {m.get_stat('launches')}, gpu_us {m.get_stat('gpu_us')}")
    print("pred_dur equal:", np.array_equal(durs[0], ref["pred_dur"]), " mismatches:",
          int((durs[0] != ref["pred_dur"]).sum()) if len(durs[0]) == len(ref["pred_dur"]) else "len")
    df = ref["stages"]["dur_float"]
    print("oracle min |frac-0.5| margin:", float(np.abs(df - np.floor(df) - 0.5).min()))
    for s in STAGES:
        print(compare(s, ref["stages"][s], m.debug_stage(s, 0)))
    if len(outs[0]) == len(ref["audio"]):
        print(compare("audio", ref["audio"], outs[0]))
    else:
        print("audio length differs:", len(outs[0]), len(ref["audio"]))


if __name__ == "__main__":
    main()

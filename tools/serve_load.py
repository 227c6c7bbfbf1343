"""Closed-loop concurrency test of the per-request entry point.

c client threads each submit `--requests` 128-token segments back to back through `kkx_infer` (the call a
server thread of the reference makes per sentence, openai lib.rs:400-412 / websocket lib.rs:371-376, where
callers queue on Mutex<Session>, ort_koko.rs:77).  "First audio" = submit -> that segment's full waveform on
the host.  Run once with the reference's one-by-one behaviour and once with the library's request coalescing.

    python tools/serve_load.py [--tokens 128] [--requests 8] [--conc 1,4,16,64] [--out profiles/x.json]
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokorox_b200.synth import ensure_weights  # noqa: E402


def segment(n_tokens, seed):
    rng = np.random.default_rng(seed)
    ids = np.concatenate([[0], rng.integers(1, 178, n_tokens), [0]]).astype(np.int64)
    style = np.random.default_rng(10_000 + seed % 54).normal(0, 0.15, 256).astype(np.float32)
    return ids, style


def run(m, conc, requests, tokens):
    lat = [[] for _ in range(conc)]
    audio = [0.0] * conc
    start = threading.Barrier(conc + 1)

    def client(i):
        segs = [segment(tokens, 3000 + i * requests + r) for r in range(requests)]
        start.wait()
        for ids, style in segs:
            t0 = time.perf_counter()
            wav = m.infer_one(ids, style, 1.0)
            lat[i].append((time.perf_counter() - t0) * 1e3)
            audio[i] += len(wav) / 24000.0
            del wav
    ts = [threading.Thread(target=client, args=(i,)) for i in range(conc)]
    for t in ts:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in ts:
        t.join()
    wall = time.perf_counter() - t0
    allv = np.sort(np.concatenate([np.asarray(x) for x in lat]))
    return {"concurrency": conc, "requests": conc * requests, "audio_s": round(sum(audio), 2),
            "audio_s_per_s": round(sum(audio) / wall, 1), "first_audio_p50_ms": round(float(np.percentile(allv, 50)), 2),
            "first_audio_p95_ms": round(float(np.percentile(allv, 95)), 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=128)
    ap.add_argument("--requests", type=int, default=8)
    ap.add_argument("--conc", default="1,4,16,64")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from kokorox_b200.onn import B200Koko
    m = B200Koko.new(ensure_weights())
    m.set_option("precision", 1)
    rows = []
    for mode, k in (("serial (reference behaviour)", 0), ("coalesce", 64)):
        m.set_option("coalesce", k)
        run(m, 4, 2, a.tokens)   # warm-up: buffers, pinned pool
        for c in [int(x) for x in a.conc.split(",")]:
            b0, r0 = m.get_stat("coalesced_batches"), m.get_stat("coalesced_requests")
            row = run(m, c, a.requests, a.tokens)
            row["mode"] = mode
            nb = m.get_stat("coalesced_batches") - b0
            row["mean_batch"] = round((m.get_stat("coalesced_requests") - r0) / nb, 2) if nb else 1.0
            rows.append(row)
            print(json.dumps(row), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"workload": f"cfg3: {a.tokens}-token segments, closed loop, {a.requests} requests per client",
                       "rows": rows}, f, indent=1)
    m.close()


if __name__ == "__main__":
    main()

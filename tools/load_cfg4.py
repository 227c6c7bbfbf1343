"""BASELINE configs[4] / SURVEY.md 8d cfg4: 4096 synthetic requests of mixed length, voice, speed and language,
sharded by request over the GPUs of one box (no data-path collective).

    python tools/load_cfg4.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/load_cfg4.py

Requests: N ~ U{10..510} tokens (seed 4), one of 54 synthetic voices looked up ON THE DEVICE by the un-padded token
count (koko.rs:1262), speed ~ U[0.8, 1.3]; 10 % are flagged "zh" and go to a second resident session (the reference
switches model files for zh voices, hf_cache.rs / vocab.rs:24-213; same architecture, so the second session loads
the same synthetic weight file).  Each rank takes the requests `shard_requests` gives it, packs them into ragged
batches of at most 64 x 512 tokens and runs them through the public host-buffer call.  Reported: whole-job
audio-s/s = audio of all ranks / max over ranks of the wall time.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokorox_b200.synth import ensure_weights  # noqa: E402

N_REQ = 4096
TOKEN_BUDGET = 64 * 512


def make_requests():
    rng = np.random.default_rng(4)
    lens = rng.integers(10, 511, size=N_REQ)
    voices = rng.integers(0, 54, size=N_REQ)
    speeds = rng.uniform(0.8, 1.3, size=N_REQ).astype(np.float32)
    zh = rng.random(N_REQ) < 0.10
    return lens, voices, speeds, zh


def tokens_of(i, n):
    r = np.random.default_rng(40_000 + i)
    return np.concatenate([[0], r.integers(1, 178, n), [0]]).astype(np.int64)


def pack(idx, lens):
    """Ragged batches under the token budget, longest first so a batch's items have similar length."""
    out, cur, tok = [], [], 0
    for i in sorted(idx, key=lambda i: -lens[i]):
        if cur and (tok + lens[i] + 2 > TOKEN_BUDGET or len(cur) >= 512):
            out.append(cur)
            cur, tok = [], 0
        cur.append(i)
        tok += lens[i] + 2
    if cur:
        out.append(cur)
    return out


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from kokorox_b200.onn import B200Koko
    from kokorox_b200.sharding import reduce_step, shard_requests

    lens, voices, speeds, zh = make_requests()
    mine = shard_requests([int(x) for x in lens], world)[rank]
    table = {f"v{v:02d}": (np.random.default_rng(900 + v).standard_normal((511, 1, 256)) * 0.15).astype(np.float32)
             for v in range(54)}
    sessions = {}
    for lang in ("en", "zh"):
        m = B200Koko.new(ensure_weights(), device=local_rank)
        m.set_option("precision", 1)
        m.load_voices(table)
        sessions[lang] = m
    work = [(lang, b) for lang in ("en", "zh") for b in pack([i for i in mine if bool(zh[i]) == (lang == "zh")], lens)]
    toks = {i: tokens_of(i, int(lens[i])) for i in mine}

    verbose = "--verbose" in sys.argv
    gpu_ms = []

    def run_all():
        audio = 0.0
        for lang, b in work:
            ta = time.perf_counter()
            outs = sessions[lang].infer_batch_voices([toks[i] for i in b], [f"v{voices[i]:02d}" for i in b],
                                                     [float(speeds[i]) for i in b])
            a = sum(len(o) for o in outs) / 24000.0
            audio += a
            gpu_ms.append(sessions[lang].get_stat("gpu_us") / 1e3)
            if verbose and rank == 0:
                print(f"{lang} B={len(b):4d} tok={sum(int(lens[i]) + 2 for i in b):6d} audio={a:8.1f}s "
                      f"wall={(time.perf_counter() - ta) * 1e3:7.1f} ms gpu={gpu_ms[-1]:7.1f} ms", flush=True)
        return audio

    # first pass = cold server (device arenas and the pinned pool grow to their working size), second pass = steady state
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    run_all()
    torch.cuda.synchronize()
    cold = time.perf_counter() - t0
    gpu_ms.clear()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    audio = run_all()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev = torch.device("cuda", local_rank)
    wall_max, (audio_all, req_all, tok_all) = reduce_step(wall, [audio, float(len(mine)), float(sum(lens[i] for i in mine))], dev)
    if rank == 0:
        print(json.dumps({"workload": "cfg4: 4096 requests, N~U{10..510}, 54 device-resident voices, speed~U[0.8,1.3], 10% second session",
                          "n_gpus": world, "requests": int(req_all), "tokens": int(tok_all), "audio_s": round(audio_all, 1),
                          "wall_s": round(wall_max, 3), "cold_first_pass_wall_s_rank0": round(cold, 3), "audio_s_per_s": round(audio_all / wall_max, 1),
                          "requests_per_s": round(req_all / wall_max, 1), "batches_rank0": len(work), "gpu_s_rank0": round(sum(gpu_ms) / 1e3, 3),
                          "timing": "host wall clock around the public host-buffer calls, max over ranks"}), flush=True)
    for m in sessions.values():
        m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Kokoro-82M backend.

Metric (BASELINE.json): audio-seconds generated per second at B=64 x 510 tokens (configs[2]);
plus p50 first-audio latency at B=1 x 510 tokens (configs[1]) as an extra key.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU oracle arm (rank 0 only)

A "step" is one pass of the whole forward over one batch of 64 synthetic utterances
(510 ids ~ U{1..177} seeds 1000+b, styles N(0,0.15^2) seeds 2000+b, speed 1.0 -- SURVEY.md 8d).
`value` times kkx_run_staged (inputs resident in HBM, CUDA events on the library's stream);
`e2e` times B200Koko.infer_batch (host numpy in, pinned host audio out, copies in the timed
region).  Multi-GPU: one process per GPU, each with its own batch (weak scaling, no collective on
the data path -- utterances are independent, koko.rs:947-1191); torch.distributed is used only
for the barrier and the max-over-ranks of the step time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md 8d work model: MACs per token (attention term scaled by N/512) and per frame
MAC_TOKEN_FIXED = 88.86e6 - 9.44e6
MAC_TOKEN_ATTN = 9.44e6
MAC_FRAME = 658.55e6


def synth_batch(B: int, n_tokens: int, rank: int = 0):
    toks, styles = [], []
    for b in range(B):
        rng = np.random.default_rng(1000 + rank * 4096 + b)
        toks.append(np.concatenate([[0], rng.integers(1, 178, n_tokens), [0]]).astype(np.int64))
        styles.append(np.random.default_rng(2000 + rank * 4096 + b).normal(0, 0.15, 256).astype(np.float32))
    return toks, np.stack(styles), np.ones(B, np.float32)


def ensure_weights():
    from tests.conftest import ensure_weights as ew
    return ew()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(n_utts: int, n_tokens: int, threads: int):
    """Times the CPU oracle (B=1 sequential, like the reference's chunk loop koko.rs:947-1191)."""
    from kokorox_b200.weightfile import read_weights
    from oracle.kokoro_ref import KokoroOracle
    o = KokoroOracle(read_weights(ensure_weights()), threads=threads)
    toks, styles, _ = synth_batch(n_utts, n_tokens)
    o.forward(toks[0][:60], styles[0], 1.0, noise_seed=0)  # warm-up (lazy init, thread pools)
    t0 = time.perf_counter()
    audio_s = 0.0
    for b in range(n_utts):
        r = o.forward(toks[b], styles[b], 1.0, noise_seed=b)
        audio_s += len(r["audio"]) / 24000.0
    dt = time.perf_counter() - t0
    return audio_s / dt, audio_s, dt


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from kokorox_b200.weightfile import read_weights
    from oracle.kokoro_ref import KokoroOracle
    o = KokoroOracle(read_weights(ensure_weights()), threads=threads)
    toks, styles, _ = synth_batch(args.batch, args.tokens)
    per_step = 1  # bounded sample: one 510-token utterance of the batch per step
    k = 0
    times, audio = [], []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        a = 0.0
        for _ in range(per_step):
            r = o.forward(toks[k % args.batch], styles[k % args.batch], 1.0, noise_seed=k)
            a += len(r["audio"]) / 24000.0
            k += 1
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt); audio.append(a)
    value = sum(audio) / sum(times)
    sample = f"{per_step} of the {args.batch} utterances per step, B=1 sequential (koko.rs:947-1191), torch fp32 oracle"
    print(json.dumps({
        "impl": "reference", "metric": "audio-seconds per second", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"B={args.batch} x {args.tokens} tokens, ragged durations (BASELINE configs[2]); "
                               "bounded sample per step", "weights": "random-init seed 1234 (no checkpoint, no network)"},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--tokens", type=int, default=510)
    ap.add_argument("--precision", type=int, default=int(os.environ.get("KKX_PRECISION", "1")),
                    help="1 = bf16 tensor-core decoder+generator (default), 0 = fp32 SIMT everywhere")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel timing table of one step here")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from kokorox_b200 import build
    from kokorox_b200.onn import B200Koko, init_ort
    build.build()
    init_ort()
    m = B200Koko.new(ensure_weights(), device=local_rank)
    m.set_option("precision", args.precision)
    toks, styles, speeds = synth_batch(args.batch, args.tokens, rank)
    sum_n = sum(len(t) for t in toks)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM, device-timed
    m.stage(toks, styles, speeds)
    launches = 0
    for _ in range(args.warmup):
        total, launches = m.run_staged()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    gpu_us, totals = [], []
    for _ in range(args.steps):
        total, launches = m.run_staged()
        gpu_us.append(m.get_stat("gpu_us"))
        totals.append(total)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_s = sum(gpu_us) * 1e-6
    frames = m.get_stat("last_frames")
    audio_s_step = totals[-1] / 24000.0

    # one profiled step (not part of the timed region) for the roofline of the dominant kernel
    m.profile_enable(True)
    m.run_staged()
    prof = m.profile()
    m.profile_enable(False)
    if args.profile_out and rank == 0:
        with open(args.profile_out, "w") as f:
            json.dump(prof, f, indent=1)

    # ---------------- e2e: host buffers through the public call, copies inside the timed region
    outs = None
    for _ in range(max(args.warmup, 3)):   # same call pattern as the timed loop: the previous result is still alive
        outs = m.infer_batch(toks, styles, speeds)   # during the next call, so the library cycles two pinned buffers
    barrier()
    t1 = time.perf_counter()
    e2e_audio = 0.0
    for _ in range(args.steps):
        outs = m.infer_batch(toks, styles, speeds)
        e2e_audio += sum(len(o) for o in outs) / 24000.0
    barrier()
    e2e_wall = time.perf_counter() - t1
    h2d = sum_n * 8 + styles.nbytes + speeds.nbytes
    d2h = int(totals[-1]) * 4 + sum_n * 4

    # ---------------- B=1 latency (configs[1]): 510 tokens, mixed style 0.4*A + 0.5*B (koko.rs:1283)
    lat = []
    if rank == 0:
        mix = (0.4 * styles[0] + 0.5 * styles[1]).astype(np.float32)
        for i in range(7):
            ta = time.perf_counter()
            m.infer_batch([toks[0]], [mix], [1.0])
            lat.append((time.perf_counter() - ta) * 1e3)
        lat = lat[2:]

    # ---------------- reductions over ranks (max time, summed work)
    from kokorox_b200.sharding import reduce_step
    step_s = dev_s / args.steps
    dev = torch.device("cuda", local_rank)
    step_s, (audio_all, frames_all, tokens_all) = reduce_step(step_s, [audio_s_step, float(frames), float(sum_n)], dev)
    wall_step_s, _ = reduce_step(wall / args.steps, [0.0], dev)
    e2e_step_s, (e2e_audio_all,) = reduce_step(e2e_wall / args.steps, [e2e_audio / args.steps], dev)

    if rank == 0:
        peaks, peaks_src = load_peaks()
        # whole-step algorithmic work, per rank
        flops_step = 2.0 * (MAC_TOKEN_FIXED * sum_n + MAC_TOKEN_ATTN * sum(len(t) ** 2 for t in toks) / 512.0
                            + MAC_FRAME * frames)
        kern = prof.get("kernels", {})
        top = max(kern.items(), key=lambda kv: kv[1][1]) if kern else ("none", [0, 0.0])
        total_us = sum(v[1] for v in kern.values()) or 1.0
        # roofline of the dominant kernel family.  The generator res-block convs (kernels_arb.cu, "arb_conv")
        # dominate the step; they are tensor-pipe work whose activations (GBs per launch) stream through HBM, so
        # both fractions are reported: algorithmic FLOPs and algorithmic bytes (every tensor once) over the summed,
        # event-timed launch durations of that kernel.
        conv_names = [k for k in kern if k.startswith("conv") or k == "arb_conv"]
        conv_us = sum(kern[k][1] for k in conv_names) or 1.0
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_gbs = peaks.get("hbm_gbs", 6650.0)
        dom = "arb_conv" if "arb_conv" in kern else (top[0] if top[0].startswith("conv") else "conv")
        if dom == "arb_conv":
            dn, dus = kern[dom]
            dflops, dbytes = prof.get("arb_flops", 0.0), prof.get("arb_bytes", 0.0)
        else:
            dn = sum(kern[k][0] for k in conv_names) or 1
            dus, dflops, dbytes = conv_us, prof.get("conv_flops", 0.0), 0.0
        achieved_tf = dflops / (dus * 1e-6) / 1e12
        # DRAM traffic per launch: the ncu --set full capture of this kernel (profiles/) measured
        # dram read+write = 1.00 x the algorithmic bytes; scale this run's per-launch algorithmic bytes by that ratio
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "r1_arb_conv_traffic.json")) as f:
                tj = json.load(f)
            if dom == "arb_conv" and dbytes > 0:
                traffic = tj["traffic_over_algorithmic"] * dbytes / max(dn, 1)
                traffic_src = "ncu dram__bytes_read+write / algorithmic bytes = %.4f (%s)" % (tj["traffic_over_algorithmic"], "profiles/r1_arb_conv_traffic.json")
        except Exception:
            pass
        roofline = {
            "bound": "tensor", "kernel": dom,
            "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": f"{peaks_src} (sustained bf16; kernel timed inside a long step)",
            "launches_per_step": dn, "avg_launch_us": dus / max(dn, 1),
            "alg_flops_per_launch": dflops / max(dn, 1),
            "alg_bytes_per_launch": dbytes / max(dn, 1),
            "hbm_achieved_gbs": dbytes / (dus * 1e-6) / 1e9, "hbm_peak_gbs": peak_gbs,
            "hbm_frac": dbytes / (dus * 1e-6) / 1e9 / peak_gbs,
            "share_of_step": dus / total_us, "top_kernel": top[0], "top_kernel_share": top[1][1] / total_us,
            "all_conv_share_of_step": conv_us / total_us,
            "all_conv_tflops": prof.get("conv_flops", 0.0) / (conv_us * 1e-6) / 1e12,
            "step_alg_tflops": flops_step / step_s / 1e12,
        }
        line = {
            "metric": "audio-seconds per second", "value": audio_all / step_s, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == 0 else "bf16", "data": "synthetic",
            "config": {"workload": f"B={args.batch} x {args.tokens} tokens per GPU, ragged durations (BASELINE configs[2])",
                       "frames_per_step": frames_all, "tokens_per_step": tokens_all,
                       "audio_s_per_step": audio_all, "weights": "random-init seed 1234 (no checkpoint, no network)",
                       "precision": args.precision, "parallelism": f"request-sharded x{world}, no collective",
                       "l2": "working set per step (GBs of activations) is larger than the 126 MB L2; no flush needed"},
            "clocks": clocks,
            "e2e": {"value": e2e_audio_all / e2e_step_s, "unit": "audio-s/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world},
            "gpu_launches": int(launches) * args.steps,
            "wall_ms_per_step": wall_step_s * 1e3,
            "roofline": roofline,
            "realtime_factor": audio_all / step_s,
            "latency_b1_510tok_ms_p50": statistics.median(lat) if lat else None,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_utts = 3
            v, a_s, dt = cpu_oracle_rate(n_utts, args.tokens, threads)
            line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                    "sample": f"{n_utts} of the {args.batch} utterances ({a_s:.1f} audio-s in {dt:.1f} s), "
                                              "B=1 sequential like koko.rs:947-1191, torch fp32 oracle"}
        print(json.dumps(line))
    m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (kokorox_b200.onn ->
libkkx.so), against the CPU oracle on the same seeded inputs and against the committed golden
fixtures.

Tolerances (fp32 configuration, "precision"=0), stated here as the north star asks:
  * integer frame durations and alignment indices: bit-exact;
  * every stage tensor and the 24 kHz waveform with the F0/N curves teacher-forced from the
    oracle: relative L2 <= 1e-4, waveform max-abs <= 1e-3 (measured: ~5e-6 / 2.4e-5);
  * free-running F0/N curves: relative L2 <= 1e-4 (measured 2e-6);
  * free-running waveform: NOT sample-comparable -- the harmonic source integrates F0 into a
    phase (2*pi*cumsum(f0*h/24000)*300 reaches 1e4..2e5 rad), so fp32 rounding differences of
    2e-6 in F0 between any two implementations decorrelate the waveform within seconds (rel-L2
    0.09 at 2.7 s, 0.3 at 33 s).  Checked instead: identical length, finite, and a
    phase-insensitive utterance-averaged power-spectrum distance (< 1 dB).
"""
import os

import numpy as np
import pytest

from tests.conftest import REF_EXAMPLE_IDS, make_noise, synth_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

STAGES = ["bert", "d_en", "d", "dur_lstm", "dur_logits", "dur_float", "t_en", "shared_lstm", "F0", "N", "dec.encode",
          "dec.decode.0", "dec.decode.1", "dec.decode.2", "dec.decode.3", "har_source", "har", "gen.x_source.0",
          "gen.ups.0", "gen.stage.0", "gen.x_source.1", "gen.ups.1", "gen.stage.1", "conv_post"]


def rel_l2(ref, got):
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    return float(np.sqrt(((ref - got) ** 2).sum() / max((ref ** 2).sum(), 1e-30)))


def avg_spectrum_db_diff(a, b, n_fft=1024, hop=512):
    """Mean |dB| difference of the utterance-averaged power spectra (phase-insensitive)."""
    def spec(x):
        n = 1 + (len(x) - n_fft) // hop
        fr = np.stack([x[i * hop:i * hop + n_fft] for i in range(n)]) * np.hanning(n_fft)
        return 10 * np.log10((np.abs(np.fft.rfft(fr, axis=1)) ** 2).mean(axis=0) + 1e-12)
    return float(np.abs(spec(a) - spec(b)).mean())


@pytest.fixture(scope="module")
def model(weights_path):
    from kokorox_b200.onn import B200Koko, init_ort
    init_ort()
    m = B200Koko.new(weights_path)
    m.set_option("precision", 0)
    yield m
    m.close()


def run_cuda(model, ids, style, speed, noise, teacher=None, stages=False):
    model.debug_enable(stages)
    model.set_noise(noise)
    for k in ("pred_dur", "F0", "N"):
        model.set_inject(k, None if teacher is None else teacher[k])
    outs, durs = model.infer_batch([ids], [style], [speed], return_durations=True)
    for k in ("pred_dur", "F0", "N"):
        model.set_inject(k, None)
    return outs[0], durs[0]


@pytest.mark.parametrize("name", ["ref_example", "ref_tokenize", "cfg0"])
def test_golden_fixtures(model, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    noise = make_noise(int(g["noise_frames_max"]), seed=int(g["noise_seed"]))
    audio, dur = run_cuda(model, g["tokens"], g["style"], float(g["speed"]), noise, stages=True)
    assert np.array_equal(dur, g["pred_dur"]), "integer durations must be bit-exact"
    assert audio.shape == g["audio"].shape
    assert rel_l2(g["stage.dur_float"], model.debug_stage("dur_float")) < 1e-5
    assert rel_l2(g["stage.F0"], model.debug_stage("F0")) < 1e-4
    assert rel_l2(g["stage.N"], model.debug_stage("N")) < 1e-4
    # teacher-forced curves -> the waveform is sample-comparable
    audio_t, _ = run_cuda(model, g["tokens"], g["style"], float(g["speed"]), noise,
                          teacher={"pred_dur": g["pred_dur"], "F0": g["stage.F0"], "N": g["stage.N"]}, stages=True)
    assert rel_l2(g["stage.har_source"], model.debug_stage("har_source")) < 1e-5
    assert rel_l2(g["audio"], audio_t) < 1e-4
    assert np.abs(g["audio"] - audio_t).max() < 1e-3


@pytest.mark.parametrize("n_tokens,seed", [(50, 0), (128, 3000), (510, 1)])
def test_stagewise_parity_teacher_forced(model, oracle, n_tokens, seed):
    ids, style = synth_case(n_tokens, seed, 100 + seed)
    noise = make_noise(12 * len(ids) if n_tokens > 60 else 50 * len(ids))
    ref = oracle.forward(ids, style, 1.0, noise=noise, stages=True)
    audio, dur = run_cuda(model, ids, style, 1.0, noise, stages=True,
                          teacher={"pred_dur": ref["pred_dur"], "F0": ref["stages"]["F0"], "N": ref["stages"]["N"]})
    assert np.array_equal(dur, ref["pred_dur"])
    assert np.array_equal(model.debug_stage("idx").astype(np.int64), ref["stages"]["idx"]), "alignment indices bit-exact"
    worst = {}
    for s in STAGES:
        worst[s] = rel_l2(ref["stages"][s], model.debug_stage(s))
    bad = {k: v for k, v in worst.items() if not v < 1e-4}
    assert not bad, bad
    assert rel_l2(ref["audio"], audio) < 1e-4
    assert np.abs(ref["audio"] - audio).max() < 1e-3


@pytest.mark.parametrize("n_tokens,seed,speed", [(50, 0, 1.0), (21, 7, 0.8), (300, 4, 1.3), (510, 1, 1.0), (510, 1000, 1.0)])
def test_durations_bit_exact_free_running(model, oracle, n_tokens, seed, speed):
    ids, style = synth_case(n_tokens, seed, 100 + seed)
    noise = make_noise(12 * len(ids) if n_tokens > 60 else 50 * len(ids))
    ref = oracle.forward(ids, style, speed, noise=noise, stages=True)
    audio, dur = run_cuda(model, ids, style, speed, noise, stages=True)
    df = ref["stages"]["dur_float"]
    margin = np.abs(df - np.floor(df) - 0.5)
    # any token whose oracle decision margin is inside the fp32 noise floor would be a coin flip
    # between two correct fp32 implementations; none of the seeded cases has one
    assert margin.min() > 5e-5, margin.min()
    assert np.array_equal(dur, ref["pred_dur"]), np.flatnonzero(dur != ref["pred_dur"])
    assert np.array_equal(model.debug_stage("idx").astype(np.int64), ref["stages"]["idx"])
    assert rel_l2(df, model.debug_stage("dur_float")) < 1e-5
    assert rel_l2(ref["stages"]["F0"], model.debug_stage("F0")) < 1e-4
    assert rel_l2(ref["stages"]["N"], model.debug_stage("N")) < 1e-4
    assert audio.shape == ref["audio"].shape and np.isfinite(audio).all()
    assert avg_spectrum_db_diff(ref["audio"], audio) < 1.0
    assert abs(float(np.sqrt((audio ** 2).mean())) / float(np.sqrt((ref["audio"] ** 2).mean())) - 1) < 0.05


def test_ragged_batch_equals_single_calls(model):
    cases = [synth_case(n, s, 200 + s) for n, s in ((21, 1), (128, 2), (50, 3), (1, 4), (300, 5))]
    speeds = [1.0, 0.9, 1.2, 1.0, 1.1]
    noise = make_noise(12 * 302 + 50 * 60)
    model.debug_enable(False)
    model.set_noise(noise)
    singles = [model.infer_batch([c[0]], [c[1]], [sp], return_durations=True) for c, sp in zip(cases, speeds)]
    outs, durs = model.infer_batch([c[0] for c in cases], [c[1] for c in cases], speeds, return_durations=True)
    for b in range(len(cases)):
        assert np.array_equal(durs[b], singles[b][1][0])
        assert len(outs[b]) == 600 * int(durs[b].sum())
        assert np.array_equal(outs[b], singles[b][0][0]), f"item {b}: batched result differs from the B=1 call"
    # frame-budget grouping must not change results either
    model.set_option("max_frames", 700)
    outs2 = model.infer_batch([c[0] for c in cases], [c[1] for c in cases], speeds)
    model.set_option("max_frames", 49152)
    for a, b in zip(outs, outs2):
        assert np.array_equal(a, b)


def test_reference_call_shape_and_edge_cases(model):
    from kokorox_b200.onn import KkxError
    style = synth_case(1, 0, 100)[1]
    model.set_noise(None)
    # the reference's own example call (ort_koko.rs:46): tokens [[...23 ids...]], styles [[256]], speed
    y = model.infer([REF_EXAMPLE_IDS], [style.tolist()], 1.0)
    assert y.ndim == 1 and y.dtype == np.float32 and len(y) % 600 == 0 and len(y) >= 600 * 23
    # empty phoneme string -> just the two pads (koko.rs:1168-1173)
    y2, d2 = model.infer_batch([[0, 0]], [style], [1.0], return_durations=True)
    assert len(y2[0]) == 600 * int(d2[0].sum()) and (d2[0] >= 1).all()
    # maximum length
    ids, st = synth_case(510, 9, 10)
    y3, d3 = model.infer_batch([ids], [st], [1.0], return_durations=True)
    assert len(d3[0]) == 512 and len(y3[0]) == 600 * int(d3[0].sum())
    # on-device noise generator is deterministic per seed
    y4 = model.infer([REF_EXAMPLE_IDS], [style], 1.0)
    assert np.array_equal(y, y4)
    for bad_tokens, bad_speed in (([0, 178, 0], 1.0), ([0, -1, 0], 1.0), ([0] * 513, 1.0), ([0, 5, 0], 0.0)):
        with pytest.raises(KkxError):
            model.infer([bad_tokens], [style], bad_speed)
    with pytest.raises(KkxError):
        model.infer([[0, 5, 0]], [style[:100]], 1.0)


def test_full_size_batch_properties(model):
    # BASELINE configs[2] at reduced B (fp32 path): size-independent properties
    B = 8
    cases = [synth_case(510, 1000 + b, 2000 + b) for b in range(B)]
    model.set_noise(None)
    outs, durs = model.infer_batch([c[0] for c in cases], [c[1] for c in cases], [1.0] * B, return_durations=True)
    for b in range(B):
        assert (durs[b] >= 1).all() and (durs[b] <= 50).all()
        assert len(outs[b]) == 600 * int(durs[b].sum())
        assert np.isfinite(outs[b]).all() and np.abs(outs[b]).max() < 50
    one = model.infer_batch([cases[3][0]], [cases[3][1]], [1.0])
    assert np.array_equal(one[0], outs[3])


# ---------------------------------------------------------------------------- bf16 tensor-core configuration
# "precision"=1: decoder + generator convs run on tcgen05 with bf16 operands and fp32 (TMEM)
# accumulation; the whole predictor (ALBERT -> durations, F0/N) stays fp32.  Stated tolerance:
# durations / indices bit-exact, F0/N rel-L2 <= 1e-4, every decoder/generator stage and the
# teacher-forced waveform rel-L2 <= 3e-2 (measured ~8e-3), waveform max-abs <= 0.15 (signal
# rms ~0.09, peak ~0.45).
@pytest.mark.parametrize("n_tokens,seed", [(50, 0), (510, 1)])
def test_bf16_tensor_core_path_parity(model, oracle, n_tokens, seed):
    ids, style = synth_case(n_tokens, seed, 100 + seed)
    noise = make_noise(12 * len(ids) if n_tokens > 60 else 50 * len(ids))
    ref = oracle.forward(ids, style, 1.0, noise=noise, stages=True)
    model.set_option("precision", 1)
    try:
        audio_free, dur = run_cuda(model, ids, style, 1.0, noise, stages=True)
        assert np.array_equal(dur, ref["pred_dur"])
        assert rel_l2(ref["stages"]["F0"], model.debug_stage("F0")) < 1e-4
        assert rel_l2(ref["stages"]["N"], model.debug_stage("N")) < 1e-4
        audio, dur = run_cuda(model, ids, style, 1.0, noise, stages=True,
                              teacher={"pred_dur": ref["pred_dur"], "F0": ref["stages"]["F0"], "N": ref["stages"]["N"]})
        worst = {s: rel_l2(ref["stages"][s], model.debug_stage(s)) for s in STAGES}
        bad = {k: v for k, v in worst.items() if not v < 3e-2}
        assert not bad, bad
        assert rel_l2(ref["audio"], audio) < 3e-2
        assert np.abs(ref["audio"] - audio).max() < 0.15
        # free-running in the configuration that ships: same phase-insensitive checks as the fp32 free-running test
        assert audio_free.shape == audio.shape and np.isfinite(audio_free).all()
        assert avg_spectrum_db_diff(ref["audio"], audio_free) < 1.0
        assert abs(float(np.sqrt((audio_free ** 2).mean())) / float(np.sqrt((ref["audio"] ** 2).mean())) - 1) < 0.05
    finally:
        model.set_option("precision", 0)


def test_benched_configuration_is_parity_tested(model, oracle):
    """The exact workload bench.py measures: kokorox_b200.synth.synth_batch(64, 510), "precision" = 1, default frame
    budget -> the frame phase runs as two groups.  Checked against the oracle and against B = 1 calls:
    (a) integer durations + alignment indices bit-exact for four items incl. the first and last item of each group;
    (b) those items bit-identical to their own B = 1 call; (c) teacher-forced waveform of the LAST item (second
    group) within the bf16 tolerance rel-L2 3e-2 / max-abs 0.15; free-running spectrum within 1 dB."""
    from kokorox_b200.synth import synth_batch
    toks, styles, speeds = synth_batch(64, 510)
    noise = make_noise(12 * 512)
    model.set_option("precision", 1)
    try:
        model.debug_enable(False)
        model.set_noise(noise)
        for k in ("pred_dur", "F0", "N"):
            model.set_inject(k, None)
        outs, durs = model.infer_batch(toks, styles, speeds, return_durations=True)
        outs = [o.copy() for o in outs]
        groups = model.get_stat("frame_groups")
        assert groups >= 2, "the benched batch is expected to need two frame groups"
        first = [model.get_stat(f"group_first:{g}") for g in range(groups)]
        assert first[0] == 0 and all(0 < f < 64 for f in first[1:])
        picks = sorted({0, first[1] - 1, first[1], 63})
        refs = {}
        for b in picks:
            refs[b] = oracle.forward(toks[b], styles[b], 1.0, noise=noise, stages=True)
            assert np.array_equal(durs[b], refs[b]["pred_dur"]), f"item {b}: durations differ from the oracle"
            assert len(outs[b]) == 600 * int(refs[b]["pred_dur"].sum())
            one, d1 = model.infer_batch([toks[b]], [styles[b]], [1.0], return_durations=True)
            assert np.array_equal(d1[0], durs[b])
            assert np.array_equal(one[0], outs[b]), f"item {b}: B=64 result differs from its own B=1 call"
            assert avg_spectrum_db_diff(refs[b]["audio"], outs[b]) < 1.0
        # alignment indices + teacher-forced waveform of an item of the SECOND group, inside the full batch
        b = 63
        model.debug_enable(True, item=b)
        model.set_inject("F0", refs[b]["stages"]["F0"], item=b)
        model.set_inject("N", refs[b]["stages"]["N"], item=b)
        outs_t, durs_t = model.infer_batch(toks, styles, speeds, return_durations=True)
        assert np.array_equal(model.debug_stage("idx", b).astype(np.int64), refs[b]["stages"]["idx"])
        assert model.debug_stage("idx", 0) is None                       # only the selected item is kept
        assert rel_l2(refs[b]["stages"]["F0"], model.debug_stage("F0", b)) == 0.0
        worst = {s: rel_l2(refs[b]["stages"][s], model.debug_stage(s, b)) for s in STAGES if s not in ("F0", "N")}
        bad = {k: v for k, v in worst.items() if not v < 3e-2}
        assert not bad, bad
        assert rel_l2(refs[b]["audio"], outs_t[b]) < 3e-2
        assert np.abs(refs[b]["audio"] - outs_t[b]).max() < 0.15
        for o in picks[:-1]:                                            # the other items are untouched by the injection
            assert np.array_equal(outs_t[o], outs[o])
        # the first item of group 2 as well (indices only)
        model.set_inject("F0", None, item=b)
        model.set_inject("N", None, item=b)
        model.debug_enable(True, item=first[1])
        model.infer_batch(toks, styles, speeds)
        assert np.array_equal(model.debug_stage("idx", first[1]).astype(np.int64), refs[first[1]]["stages"]["idx"])
    finally:
        model.debug_enable(False)
        for k in ("F0", "N"):
            model.set_inject(k, None, item=63)
        model.set_noise(None)
        model.set_option("precision", 0)


@pytest.mark.parametrize("precision", [0, 1])
def test_stft_replicate_padding_convention(model, weights, precision):
    """SURVEY hard part 6: the conv-based STFT of ONNX exports pads by replication, torch.stft by reflection.  Both
    are selectable ("stft_replicate"); the replicate mode is checked against KokoroOracle(stft_pad_mode="replicate")."""
    from oracle.kokoro_ref import KokoroOracle
    orc = KokoroOracle(weights, stft_pad_mode="replicate")
    ids, style = synth_case(50, 0, 100)
    noise = make_noise(50 * len(ids))
    ref = orc.forward(ids, style, 1.0, noise=noise, stages=True)
    tol = 1e-4 if precision == 0 else 3e-2
    model.set_option("precision", precision)
    model.set_option("stft_replicate", 1)
    try:
        audio, dur = run_cuda(model, ids, style, 1.0, noise, stages=True,
                              teacher={"pred_dur": ref["pred_dur"], "F0": ref["stages"]["F0"], "N": ref["stages"]["N"]})
        assert np.array_equal(dur, ref["pred_dur"])
        assert rel_l2(ref["stages"]["har"], model.debug_stage("har")) < 1e-4
        assert rel_l2(ref["audio"], audio) < tol
        # and the two conventions really differ at the utterance edges
        model.set_option("stft_replicate", 0)
        audio_reflect, _ = run_cuda(model, ids, style, 1.0, noise, stages=True,
                                    teacher={"pred_dur": ref["pred_dur"], "F0": ref["stages"]["F0"], "N": ref["stages"]["N"]})
        assert not np.array_equal(model.debug_stage("har")[:3], ref["stages"]["har"][:3])
        assert audio_reflect.shape == audio.shape
    finally:
        model.set_option("stft_replicate", 0)
        model.set_option("precision", 0)


def test_asynchronous_submit_and_wait(model):
    """kkx_submit / kkx_poll / kkx_wait (websocket lib.rs:371-376): one thread keeps two tickets in
    flight -- sentence k+1 is queued while k is still running -- and every result equals the blocking call."""
    import time
    from kokorox_b200.onn import KkxError
    cases = [synth_case(n, 600 + n, 700 + n) for n in (128, 40, 128, 90, 17)]
    speeds = [1.0, 1.1, 0.9, 1.0, 1.2]
    model.set_noise(None)
    model.set_option("precision", 1)
    try:
        alone = [model.infer_one(c[0], c[1], sp, return_durations=True) for c, sp in zip(cases, speeds)]
        t0 = model.submit(cases[0][0], cases[0][1], speeds[0])
        t1 = model.submit(cases[1][0], cases[1][1], speeds[1])       # second ticket while the first is in flight
        assert t0 != t1
        a0, d0 = model.wait(t0, return_durations=True)
        t2 = model.submit(cases[2][0], cases[2][1], speeds[2])       # "send" sentence 0 while 1 and 2 synthesise
        a1, d1 = model.wait(t1, return_durations=True)
        a2 = model.wait(t2)
        assert np.array_equal(a0, alone[0][0]) and np.array_equal(d0, alone[0][1])
        assert np.array_equal(a1, alone[1][0]) and np.array_equal(d1, alone[1][1])
        assert np.array_equal(a2, alone[2][0])
        # a burst leaves as ragged batches; poll turns true without blocking
        before = model.get_stat("async_batches")
        ts = [model.submit(c[0], c[1], sp) for c, sp in zip(cases, speeds)]
        deadline = time.time() + 30
        while not all(model.poll(t) for t in ts) and time.time() < deadline:
            time.sleep(0.002)
        assert all(model.poll(t) for t in ts)
        res = [model.wait(t) for t in ts]
        assert all(np.array_equal(r, a[0]) for r, a in zip(res, alone))
        assert model.get_stat("async_batches") - before < len(ts)    # at least two requests shared a batch
        assert model.get_stat("async_requests") >= 8
        # errors: bad requests are rejected at submit, tickets are redeemed once
        with pytest.raises(KkxError):
            model.submit([0, 999, 0], cases[0][1], 1.0)
        with pytest.raises(KkxError):
            model.submit(cases[0][0], cases[0][1], 0.0)
        with pytest.raises(KkxError):
            model.wait(ts[0])
        # blocking calls and tickets interleave on one session
        t = model.submit(cases[3][0], cases[3][1], speeds[3])
        y = model.infer_one(cases[4][0], cases[4][1], speeds[4])
        assert np.array_equal(y, alone[4][0]) and np.array_equal(model.wait(t), alone[3][0])
    finally:
        model.set_option("precision", 0)


def test_latency_path_graphs_and_forks_do_not_change_results(model):
    """Single-utterance calls replay the token phase from a CUDA graph (captured on the second call with a given
    token count) and small batches run independent branches on two streams; neither may change a single bit."""
    model.set_noise(None)
    model.debug_enable(False)                  # stage dumps (left on by earlier tests) disable both mechanisms
    model.set_option("precision", 1)
    try:
        for n in (50, 128, 510):
            a, sa = synth_case(n, 40 + n, 41 + n)
            b, sb = synth_case(n, 42 + n, 43 + n)            # same token count, different content / style / speed
            model.set_option("latency_graphs", 0)
            model.set_option("fork_max_batch", 0)
            want_a = model.infer_one(a, sa, 1.0, return_durations=True)
            want_b = model.infer_one(b, sb, 1.15, return_durations=True)
            model.set_option("latency_graphs", 1)
            model.set_option("fork_max_batch", 4)
            r0 = model.get_stat("graph_replays")
            got = [model.infer_one(a, sa, 1.0, return_durations=True) for _ in range(3)]   # eager, capture, replay
            got_b = model.infer_one(b, sb, 1.15, return_durations=True)                     # replay with new inputs
            assert model.get_stat("graph_replays") >= r0 + 3
            for y, d in got:
                assert np.array_equal(d, want_a[1]) and np.array_equal(y, want_a[0])
            assert np.array_equal(got_b[1], want_b[1]) and np.array_equal(got_b[0], want_b[0])
        # a small batch with forks == the same batch without
        cases = [synth_case(n, 50 + n, 51 + n) for n in (33, 140, 77)]
        model.set_option("fork_max_batch", 0)
        ref = [o.copy() for o in model.infer_batch([c[0] for c in cases], [c[1] for c in cases], [1.0, 0.9, 1.2])]
        model.set_option("fork_max_batch", 4)
        got = model.infer_batch([c[0] for c in cases], [c[1] for c in cases], [1.0, 0.9, 1.2])
        assert all(np.array_equal(x, y) for x, y in zip(ref, got))
        # the fp32 verification configuration takes the same paths
        model.set_option("precision", 0)
        a, sa = synth_case(60, 1, 2)
        model.set_option("latency_graphs", 0)
        model.set_option("fork_max_batch", 0)
        want = model.infer_one(a, sa, 1.0).copy()
        model.set_option("latency_graphs", 1)
        model.set_option("fork_max_batch", 4)
        for _ in range(3):
            assert np.array_equal(model.infer_one(a, sa, 1.0), want)
    finally:
        model.set_option("latency_graphs", 1)
        model.set_option("fork_max_batch", 4)
        model.set_option("precision", 0)


@pytest.mark.parametrize("n_tokens,seed", [(50, 0), (510, 1)])
def test_bf16_residual_stream_parity(model, oracle, n_tokens, seed):
    """"stream_bf16": the generator res-blocks keep their residual stream in bf16 between iterations.  Same bars as the
    tensor-core configuration: every stage and the teacher-forced waveform within rel-L2 3e-2 / max-abs 0.15 of the
    oracle, and a ragged batch still equals its single calls bit for bit."""
    ids, style = synth_case(n_tokens, seed, 100 + seed)
    noise = make_noise(12 * len(ids) if n_tokens > 60 else 50 * len(ids))
    ref = oracle.forward(ids, style, 1.0, noise=noise, stages=True)
    model.set_option("precision", 1)
    model.set_option("stream_bf16", 1)
    try:
        audio, dur = run_cuda(model, ids, style, 1.0, noise, stages=True,
                              teacher={"pred_dur": ref["pred_dur"], "F0": ref["stages"]["F0"], "N": ref["stages"]["N"]})
        assert np.array_equal(dur, ref["pred_dur"])
        worst = {s: rel_l2(ref["stages"][s], model.debug_stage(s)) for s in STAGES}
        print("stream_bf16 stage errors:", {k: round(v, 5) for k, v in worst.items() if k.startswith("gen") or k == "conv_post"},
              "audio", rel_l2(ref["audio"], audio))
        bad = {k: v for k, v in worst.items() if not v < 3e-2}
        assert not bad, bad
        assert rel_l2(ref["audio"], audio) < 3e-2
        assert np.abs(ref["audio"] - audio).max() < 0.15
        if n_tokens == 50:
            model.debug_enable(False)
            model.set_noise(None)
            cases = [synth_case(n, s, 300 + s) for n, s in ((40, 1), (200, 2), (90, 3))]
            singles = [model.infer_batch([c[0]], [c[1]], [1.0])[0].copy() for c in cases]
            outs = model.infer_batch([c[0] for c in cases], [c[1] for c in cases], [1.0] * 3)
            assert all(np.array_equal(a, b) for a, b in zip(outs, singles))
    finally:
        model.set_option("stream_bf16", 0)
        model.set_option("precision", 0)


def test_fused_noise_conv_is_closer_to_the_oracle(model, oracle):
    """"fuse_noise_stats": the generator's Conv1d(22, 128, k = 1) over the STFT frames in fp32 with the AdaIN statistics of
    its output from the same pass, against bf16 operands on the tensor cores + a separate statistics pass.  The stage the
    conv feeds (gen.x_source.1 = noise_res[1] of it) must stay inside the tensor-core bar either way and must not get
    worse with the fp32 conv; the waveform stays inside 3e-2 / 0.15."""
    ids, style = synth_case(120, 5, 105)
    noise = make_noise(50 * len(ids))
    ref = oracle.forward(ids, style, 1.0, noise=noise, stages=True)
    teacher = {"pred_dur": ref["pred_dur"], "F0": ref["stages"]["F0"], "N": ref["stages"]["N"]}
    model.set_option("precision", 1)
    try:
        err = {}
        for fused in (0, 1):
            model.set_option("fuse_noise_stats", fused)
            audio, dur = run_cuda(model, ids, style, 1.0, noise, stages=True, teacher=teacher)
            assert np.array_equal(dur, ref["pred_dur"])
            err[fused] = (rel_l2(ref["stages"]["gen.x_source.1"], model.debug_stage("gen.x_source.1")), rel_l2(ref["audio"], audio))
            assert err[fused][0] < 3e-2 and err[fused][1] < 3e-2, err
            assert np.abs(ref["audio"] - audio).max() < 0.15
        print("x_source.1 / audio rel-L2, bf16 tensor-core conv vs fp32 fused conv:", err)
        assert err[1][0] <= err[0][0] * 1.05, err
    finally:
        model.set_option("fuse_noise_stats", 1)
        model.set_option("precision", 0)


@pytest.mark.parametrize("opts", [{"attention_umma": 0, "split_f16": 0}, {"attention_umma": 1, "split_f16": 0},
                                  {"attention_umma": 0, "split_f16": 1}, {"attention_umma": 1, "split_f16": 1}])
@pytest.mark.parametrize("n_tokens,seed,speed", [(50, 0, 1.0), (300, 4, 1.3), (510, 1, 1.0), (510, 1000, 1.0)])
def test_round2_kernels_keep_durations_bit_exact(model, oracle, opts, n_tokens, seed, speed):
    """The tcgen05 attention kernel and the split-FP16 GEMMs (both on by default) replace fp32-grade kernels on the
    path that decides the integer durations: every combination with their round-1 counterparts (mma.sync attention,
    split-TF32 planes) must match the oracle bit for bit, with dur_float as close as the round-1 kernels get it."""
    ids, style = synth_case(n_tokens, seed, 100 + seed)
    noise = make_noise(12 * len(ids) if n_tokens > 60 else 50 * len(ids))
    ref = oracle.forward(ids, style, speed, noise=noise, stages=True)
    model.set_option("precision", 1)
    for k, v in opts.items():
        model.set_option(k, v)
    try:
        audio, dur = run_cuda(model, ids, style, speed, noise, stages=True)
        assert np.array_equal(dur, ref["pred_dur"]), np.flatnonzero(dur != ref["pred_dur"])
        assert rel_l2(ref["stages"]["dur_float"], model.debug_stage("dur_float")) < 1e-5
        assert rel_l2(ref["stages"]["bert"], model.debug_stage("bert")) < 1e-4
        assert rel_l2(ref["stages"]["F0"], model.debug_stage("F0")) < 1e-4
        assert rel_l2(ref["stages"]["N"], model.debug_stage("N")) < 1e-4
        assert audio.shape == ref["audio"].shape and np.isfinite(audio).all()
    finally:
        for k in opts:
            model.set_option(k, 1)          # the library defaults
        model.set_option("precision", 0)


def test_pair_gemm_and_plane_fusion_are_bit_identical(model):
    """Round-2 GEMM path of the big-batch case: CTA-pair kernel (tcgen05 cta_group::2, "gemm_pair") and operand planes
    handed from LayerNorm / the FFN GEMM straight to the next GEMM ("fuse_planes").  Both only change WHERE the same
    arithmetic happens, so a batch large enough to take the pair kernel (>= 148 output tiles per GEMM) must produce the
    same bits with either switched off -- and each utterance the same bits as when it is run alone (single-tile
    kernels, no pair, planes from LayerNorm only)."""
    model.set_noise(None)
    model.debug_enable(False)
    model.set_option("precision", 1)
    try:
        cases = [synth_case(510 if i % 3 else 200 + 7 * i, 700 + i, 800 + i) for i in range(36)]
        toks, styles = [c[0] for c in cases], [c[1] for c in cases]
        speeds = [1.0 + 0.01 * (i % 5) for i in range(36)]
        outs = {}
        for pair, planes, cpair in ((1, 1, 1), (0, 0, 0), (1, 0, 1), (0, 1, 0), (1, 1, 0)):
            model.set_option("gemm_pair", pair)
            model.set_option("fuse_planes", planes)
            model.set_option("conv_pair", cpair)       # bf16 decoder convs on CTA pairs (256 x 256 tiles)
            outs[(pair, planes, cpair)] = [o.copy() for o in model.infer_batch(toks, styles, speeds)]
        ref = outs[(0, 0, 0)]
        for key, got in outs.items():
            assert len(got) == len(ref)
            assert all(np.array_equal(x, y) for x, y in zip(ref, got)), f"gemm_pair, fuse_planes, conv_pair = {key} changed the result"
        model.set_option("gemm_pair", 1)
        model.set_option("fuse_planes", 1)
        model.set_option("conv_pair", 1)
        for i in (0, 1, 17):
            one = model.infer_one(toks[i], styles[i], speeds[i])
            assert np.array_equal(one, ref[i]), f"utterance {i} alone differs from the batch"
    finally:
        model.set_option("gemm_pair", 1)
        model.set_option("fuse_planes", 1)
        model.set_option("conv_pair", 1)
        model.set_option("precision", 0)


def test_phase_fused_upsampling_is_bit_identical(model):
    """"fuse_phases": the 10 + 6 ConvTranspose1d phase convs of the generator as one launch each (phase = fastest grid
    dimension, weights stacked along the map's rows).  Same kernel, same accumulation order -> the same bits."""
    model.set_noise(None)
    model.debug_enable(False)
    model.set_option("precision", 1)
    try:
        cases = [synth_case(n, 60 + n, 61 + n) for n in (33, 140, 77, 510)]
        speeds = [1.0, 0.9, 1.2, 1.0]
        model.set_option("fuse_phases", 0)
        ref = [o.copy() for o in model.infer_batch([c[0] for c in cases], [c[1] for c in cases], speeds)]
        model.set_option("fuse_phases", 1)
        # "ups_phase_loop": the stage-1 launch as one CTA per m-tile looping over the 6 phases (1: two-stage ring, 2: three
        # stages), as persistent CTAs with one phase's weights resident (3), or as one CTA per (m-tile, phase) (0) -- the same MMAs in the same order per output element
        for loop in (0, 1, 2, 3):
            model.set_option("ups_phase_loop", loop)
            got = model.infer_batch([c[0] for c in cases], [c[1] for c in cases], speeds)
            assert all(np.array_equal(x, y) for x, y in zip(ref, got)), f"ups_phase_loop = {loop} changed the result"
        model.set_option("ups_phase_loop", 3)
        one = model.infer_one(cases[1][0], cases[1][1], speeds[1])
        assert np.array_equal(one, ref[1])
    finally:
        model.set_option("fuse_phases", 1)          # the library defaults
        model.set_option("ups_phase_loop", 3)
        model.set_option("precision", 0)


def test_call_sequence_and_limits(model):
    """the staged API cannot be driven into a stale state, speed and frame counts are bounded."""
    from kokorox_b200.onn import KkxError
    ids, style = synth_case(30, 1, 2)
    model.set_noise(None)
    model.stage([ids], [style], [1.0])
    n, _ = model.run_staged()
    good = model.fetch_staged(n)[0].copy()
    # a rejected stage leaves the previous batch runnable and fetchable
    with pytest.raises(KkxError):
        model.stage([ids, [0, 500, 0]], [style, style], [1.0, 1.0])
    assert np.array_equal(model.fetch_staged(n)[0], good)
    n2, _ = model.run_staged()
    assert n2 == n and np.array_equal(model.fetch_staged(n2)[0], good)
    # fetch before any run of a newly staged batch is a state error, not a stale read
    model.stage([ids[:10].tolist() + [0]], [style], [1.0])
    with pytest.raises(KkxError) as e:
        model.fetch_staged(n)
    assert e.value.code == -5
    for bad_speed in (0.0, -1.0, 0.05, 11.0, float("nan"), float("inf")):
        with pytest.raises(KkxError):
            model.infer_one(ids, style, bad_speed)
    # the slowest accepted speed: 10x longer durations, still bounded
    y, d = model.infer_one(ids[:12].tolist() + [0], style, 0.1, return_durations=True)
    assert len(y) == 600 * int(d.sum()) and d.max() <= 500
    # an injected duration table that would exceed the per-utterance frame cap is refused
    with pytest.raises(KkxError):
        model.set_inject("pred_dur", np.full(len(ids), 501, np.int32))
    with pytest.raises(KkxError):
        model.set_inject("pred_dur", np.zeros(len(ids), np.int32))


def test_bf16_batch_equals_single(model):
    cases = [synth_case(n, s, 300 + s) for n, s in ((40, 1), (200, 2), (90, 3))]
    model.set_option("precision", 1)
    try:
        model.set_noise(None)
        singles = [model.infer_batch([c[0]], [c[1]], [1.0]) for c in cases]
        outs = model.infer_batch([c[0] for c in cases], [c[1] for c in cases], [1.0] * 3)
        for b in range(3):
            assert np.array_equal(outs[b], singles[b][0])
    finally:
        model.set_option("precision", 0)


def test_many_short_requests_bf16(model):
    # BASELINE configs[3]/[4] flavour: hundreds of short, mixed-length, mixed-speed requests in one call (more items
    # than one frame group may hold).  Every item must equal its own B=1 call bit for bit, and lengths must follow
    # the integer durations.
    rng = np.random.default_rng(4)
    B = 600
    lens = rng.integers(3, 41, size=B)
    cases = [synth_case(int(n), 5000 + i, 6000 + (i % 54)) for i, n in enumerate(lens)]
    speeds = rng.uniform(0.8, 1.3, size=B).astype(np.float32).tolist()
    model.set_option("precision", 1)
    try:
        model.set_noise(None)
        outs, durs = model.infer_batch([c[0] for c in cases], [c[1] for c in cases], speeds, return_durations=True)
        assert len(outs) == B
        for b in range(B):
            assert len(outs[b]) == 600 * int(durs[b].sum()) and np.isfinite(outs[b]).all()
        for b in (0, 17, 311, 599):
            one, d1 = model.infer_batch([cases[b][0]], [cases[b][1]], [speeds[b]], return_durations=True)
            assert np.array_equal(d1[0], durs[b])
            assert np.array_equal(one[0], outs[b]), f"item {b}: batched result differs from the B=1 call"
    finally:
        model.set_option("precision", 0)


# ---------------------------------------------------------------------------- "next" rows
def _voices(n=5, seed=77):
    rng = np.random.default_rng(seed)
    names = ["af_sky", "af_nicole", "am_echo", "jf_alpha", "zf_xiaoxiao"][:n]
    return {nm: (rng.standard_normal((511, 1, 256)) * 0.15).astype(np.float32) for nm in names}


def test_device_voice_table_and_mix_styles(model):
    # TTSKoko::load_voices + mix_styles (koko.rs:1255-1334) on the device: the device-mixed style must equal the
    # reference loop (f32, un-normalised portions = weight * 0.1, row = un-padded token count) bit for bit, so the
    # audio equals the host-mixed call exactly.
    from kokorox_b200.onn import KkxError
    model.load_voices(_voices())
    toks = [synth_case(n, 900 + n, 1)[0] for n in (30, 77, 12)]
    names = ["af_sky", "af_sky.4+af_nicole.5", "jf_alpha.3+am_echo.3+zf_xiaoxiao.4"]
    speeds = [1.0, 0.9, 1.1]
    model.set_noise(None)
    host_styles = [model.mix_styles(nm, len(t) - 2)[0] for nm, t in zip(names, toks)]
    # the reference arithmetic, spelled out for the two-voice mix (koko.rs:1296-1302)
    tab = model._voice_table
    ids = model._voice_ids
    row = len(toks[1]) - 2
    want = np.zeros(256, np.float32)
    want = want + tab[ids["af_sky"], row] * np.float32(np.float32(4.0) * np.float32(0.1))
    want = want + tab[ids["af_nicole"], row] * np.float32(np.float32(5.0) * np.float32(0.1))
    assert np.array_equal(want.astype(np.float32), host_styles[1])
    ref, rd = model.infer_batch(toks, host_styles, speeds, return_durations=True)
    got, gd = model.infer_batch_voices(toks, names, speeds, return_durations=True)
    for b in range(3):
        assert np.array_equal(rd[b], gd[b])
        assert np.array_equal(ref[b], got[b]), f"item {b}: device-mixed style differs from the host-mixed one"
    with pytest.raises(KkxError):
        model.infer_batch_voices(toks[:1], ["no_such_voice"], [1.0])
    with pytest.raises(KkxError):
        model.infer_batch_voices(toks[:1], ["af_sky.x"], [1.0])


def test_pcm16_output_fused_into_istft(model):
    # f32 -> i16 like kokorox-websocket/src/lib.rs:699-703: (s.clamp(-1, 1) * 32767) as i16 (truncation)
    cases = [synth_case(n, 40 + n, 41 + n) for n in (25, 140)]
    model.set_noise(None)
    f = model.infer_batch([c[0] for c in cases], [c[1] for c in cases], [1.0, 1.2])
    p = model.infer_batch_pcm16([c[0] for c in cases], [c[1] for c in cases], [1.0, 1.2])
    for a, b in zip(f, p):
        assert b.dtype == np.int16 and len(a) == len(b)
        want = np.trunc(np.clip(a, -1.0, 1.0).astype(np.float32) * np.float32(32767.0)).astype(np.int16)
        assert np.array_equal(want, b)


def _run_threads(fn, n):
    import threading
    res, errs = [None] * n, [None] * n

    def work(i):
        try:
            res[i] = fn(i)
        except Exception as e:  # noqa: BLE001 -- collected and asserted by the caller
            errs[i] = e
    ts = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return res, errs


def test_concurrent_callers_are_coalesced_into_one_batch(model):
    # SURVEY 8f row 4 / 8b "thread-safe": 16 threads call kkx_infer at once, the way the reference's server
    # threads call OrtKoko::infer (openai lib.rs:400-412) and queue on Mutex<Session> (ort_koko.rs:77).  With
    # "coalesce" on they leave as ONE ragged batch; every caller must get exactly its own B=1 result.
    from kokorox_b200.onn import KkxError
    n = 16
    rng = np.random.default_rng(11)
    cases = [synth_case(int(k), 7000 + i, 7100 + i) for i, k in enumerate(rng.integers(4, 120, size=n))]
    speeds = rng.uniform(0.8, 1.3, size=n).astype(np.float32).tolist()
    model.set_noise(None)
    model.set_option("precision", 1)
    try:
        alone = [model.infer_one(cases[i][0], cases[i][1], speeds[i], return_durations=True) for i in range(n)]
        b0, r0 = model.get_stat("coalesced_batches"), model.get_stat("coalesced_requests")
        assert b0 == 0 and r0 == 0          # off by default: the reference's one-by-one behaviour
        model.set_option("coalesce", n)
        model.set_option("coalesce_wait_us", 500000)   # the first caller waits until all 16 have queued
        res, errs = _run_threads(lambda i: model.infer_one(cases[i][0], cases[i][1], speeds[i], return_durations=True), n)
        assert all(e is None for e in errs), errs
        assert model.get_stat("coalesced_requests") == n
        assert model.get_stat("coalesced_batches") == 1 and model.get_stat("coalesced_largest") == n
        for i in range(n):
            assert np.array_equal(res[i][1], alone[i][1])
            assert np.array_equal(res[i][0], alone[i][0]), f"caller {i}: coalesced result differs from its own call"
        # the shared buffer returns to the pool with its last view; the next round reuses it
        del res
        res, errs = _run_threads(lambda i: model.infer_one(cases[i][0], cases[i][1], speeds[i]), n)
        assert all(e is None for e in errs), errs
        assert all(np.array_equal(res[i], alone[i][0]) for i in range(n))
        # a bad request fails alone (id 9999 is outside the vocabulary); its batch mates still get audio
        before = model.get_stat("coalesced_requests")

        def one(i):
            ids = list(cases[i][0])
            if i == 5:
                ids[2] = 9999
            return model.infer_one(ids, cases[i][1], speeds[i])
        res, errs = _run_threads(one, n)
        assert isinstance(errs[5], KkxError) and res[5] is None
        assert all(errs[i] is None and np.array_equal(res[i], alone[i][0]) for i in range(n) if i != 5)
        assert model.get_stat("coalesced_requests") == before + n - 1
        # no waiting window: still correct, batches form from whoever queued behind the running step
        model.set_option("coalesce_wait_us", 0)
        res, errs = _run_threads(lambda i: model.infer_one(cases[i][0], cases[i][1], speeds[i]), n)
        assert all(e is None for e in errs), errs
        assert all(np.array_equal(res[i], alone[i][0]) for i in range(n))
    finally:
        model.set_option("coalesce", 0)
        model.set_option("coalesce_wait_us", 0)
        model.set_option("precision", 0)


def test_oversized_batch_runs_in_passes(model):
    # a call may carry more tokens than one pass holds ("max_tokens"): the library runs several passes and
    # concatenates; every waveform, offset and duration must equal the single-pass result bit for bit
    rng = np.random.default_rng(21)
    B = 14
    cases = [synth_case(int(n), 8000 + i, 8100 + i) for i, n in enumerate(rng.integers(20, 200, size=B))]
    speeds = rng.uniform(0.8, 1.3, size=B).astype(np.float32).tolist()
    model.set_noise(None)
    model.set_option("precision", 1)
    try:
        ref, rdur = model.infer_batch([c[0] for c in cases], [c[1] for c in cases], speeds, return_durations=True)
        ref = [r.copy() for r in ref]
        model.set_option("max_tokens", 512)
        got, gdur = model.infer_batch([c[0] for c in cases], [c[1] for c in cases], speeds, return_durations=True)
        assert len(got) == B
        for b in range(B):
            assert np.array_equal(gdur[b], rdur[b])
            assert np.array_equal(got[b], ref[b]), f"item {b} differs when the batch is split into passes"
        with pytest.raises(Exception):
            model.set_option("max_tokens", 10)
    finally:
        model.set_option("max_tokens", 40960)
        model.set_option("precision", 0)


def test_two_sessions_run_concurrently(model, weights_path):
    # one process may hold several sessions (one per GPU, or en + zh models on one GPU: koko.rs TTSManager); two
    # threads driving two sessions at the same time must each get exactly what they get alone
    from kokorox_b200.onn import B200Koko
    other = B200Koko.new(weights_path)
    try:
        cases = [synth_case(60 + 7 * i, 9000 + i, 9100 + i) for i in range(6)]
        for m in (model, other):
            m.set_noise(None)
            m.set_option("precision", 1)
        alone = [model.infer_one(c[0], c[1], 1.0) .copy() for c in cases]

        def work(i):
            m = model if i % 2 == 0 else other
            return [m.infer_one(c[0], c[1], 1.0).copy() for c in cases]
        res, errs = _run_threads(work, 2)
        assert all(e is None for e in errs), errs
        for r in res:
            assert all(np.array_equal(a, b) for a, b in zip(r, alone))
    finally:
        other.close()
        model.set_option("precision", 0)

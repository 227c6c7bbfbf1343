// loadgen.cpp -- native closed-loop load generator for BASELINE configs[3] (streaming / pipe mode: sentence-chunked
// 128-token segments, first-audio latency under concurrent load on one B200).
//
// It is a CLIENT of the C ABI (include/kkx.h), doing what the reference's server threads do: every client thread
// submits its segments back to back through kkx_infer -- the call `TTSKoko::tts_raw_audio` makes per sentence
// (kokorox-websocket/src/lib.rs:371-376, kokorox-openai/src/lib.rs:400-412), where the reference's callers queue on
// Mutex<Session> (ort_koko.rs:77).  "First audio" = submit -> that segment's complete waveform in host memory.
// Three modes on the same session:
//   serial    the reference's behaviour (one step per request, callers wait their turn)
//   coalesce  kkx_set_option("coalesce", 64): queued callers leave as one ragged batch
//   async     kkx_submit / kkx_wait with two tickets in flight per client (sentence k+1 synthesises while k is "sent")
// Output: one JSON object per (mode, concurrency) on stdout.
//   kkx_loadgen <model file> [--device 0] [--tokens 128] [--requests 8] [--conc 1,4,16,64] [--modes serial,coalesce,async]
#include "../../include/kkx.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

namespace {
using Clock = std::chrono::steady_clock;

struct Segment { std::vector<int64_t> ids; std::vector<float> style; };

Segment make_segment(int n_tokens, uint64_t seed) {
  // synthetic phoneme ids ~ U{1..177} wrapped in the 0 pads of koko.rs:1168-1173; style ~ N(0, 0.15^2) out of 54 voices
  std::mt19937_64 rng(seed);
  std::uniform_int_distribution<int> id(1, 177);
  Segment s;
  s.ids.push_back(0);
  for (int i = 0; i < n_tokens; i++) s.ids.push_back(id(rng));
  s.ids.push_back(0);
  std::mt19937_64 vr(10000 + seed % 54);
  std::normal_distribution<float> nd(0.f, 0.15f);
  s.style.resize(256);
  for (auto& v : s.style) v = nd(vr);
  return s;
}

struct Barrier {
  std::mutex m; std::condition_variable cv; int waiting = 0, target; long gen = 0;
  explicit Barrier(int n) : target(n) {}
  void wait() {
    std::unique_lock<std::mutex> l(m);
    const long g = gen;
    if (++waiting == target) { waiting = 0; gen++; cv.notify_all(); }
    else cv.wait(l, [&] { return gen != g; });
  }
};

struct Row { double p50, p95, audio_s, wall_s; long requests, failures; };

double pct(std::vector<double>& v, double q) {
  if (v.empty()) return 0;
  std::sort(v.begin(), v.end());
  const double pos = q * (v.size() - 1);
  const size_t i = (size_t)pos;
  const double f = pos - i;
  return i + 1 < v.size() ? v[i] * (1 - f) + v[i + 1] * f : v[i];
}

Row run(kkx_ctx* ctx, int conc, int requests, int tokens, bool use_async, uint64_t seed0) {
  std::vector<std::vector<double>> lat(conc);
  std::vector<double> audio(conc, 0.0);
  std::atomic<long> failures{0};
  Barrier start(conc + 1);
  std::vector<std::thread> th;
  for (int c = 0; c < conc; c++) {
    th.emplace_back([&, c] {
      std::vector<Segment> segs;
      for (int r = 0; r < requests; r++) segs.push_back(make_segment(tokens, seed0 + (uint64_t)c * requests + r));
      start.wait();
      if (!use_async) {
        for (auto& s : segs) {
          const auto t0 = Clock::now();
          float* wav = nullptr; int64_t n = 0;
          const int rc = kkx_infer(ctx, s.ids.data(), (int32_t)s.ids.size(), s.style.data(), 1.0f, &wav, &n, nullptr);
          lat[c].push_back(std::chrono::duration<double, std::milli>(Clock::now() - t0).count());
          if (rc != KKX_OK) { failures++; continue; }
          audio[c] += (double)n / KKX_SAMPLE_RATE;
          kkx_release(ctx, wav);
        }
      } else {
        // two tickets in flight: submit k+1, then wait for k (the WebSocket handler would be sending k meanwhile)
        std::vector<kkx_ticket> tk(segs.size(), 0);
        std::vector<Clock::time_point> t0(segs.size());
        auto submit = [&](size_t k) {
          t0[k] = Clock::now();
          if (kkx_submit(ctx, segs[k].ids.data(), (int32_t)segs[k].ids.size(), segs[k].style.data(), 1.0f, &tk[k]) != KKX_OK) { failures++; tk[k] = 0; }
        };
        if (!segs.empty()) submit(0);
        for (size_t k = 0; k < segs.size(); k++) {
          if (k + 1 < segs.size()) submit(k + 1);
          if (!tk[k]) continue;
          float* wav = nullptr; int64_t n = 0;
          const int rc = kkx_wait(ctx, tk[k], &wav, &n, nullptr);
          lat[c].push_back(std::chrono::duration<double, std::milli>(Clock::now() - t0[k]).count());
          if (rc != KKX_OK) { failures++; continue; }
          audio[c] += (double)n / KKX_SAMPLE_RATE;
          kkx_release(ctx, wav);
        }
      }
    });
  }
  start.wait();
  const auto t0 = Clock::now();
  for (auto& t : th) t.join();
  Row r;
  r.wall_s = std::chrono::duration<double>(Clock::now() - t0).count();
  std::vector<double> all;
  r.audio_s = 0;
  for (int c = 0; c < conc; c++) { all.insert(all.end(), lat[c].begin(), lat[c].end()); r.audio_s += audio[c]; }
  r.requests = (long)all.size();
  r.failures = failures.load();
  r.p50 = pct(all, 0.50);
  r.p95 = pct(all, 0.95);
  return r;
}

std::vector<std::string> split(const std::string& s) {
  std::vector<std::string> out;
  size_t i = 0;
  while (i <= s.size()) {
    const size_t j = std::min(s.find(',', i), s.size());
    if (j > i) out.push_back(s.substr(i, j - i));
    i = j + 1;
  }
  return out;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: kkx_loadgen <model file> [--device D] [--tokens N] [--requests R] [--conc 1,4,16,64] [--modes serial,coalesce,async]\n");
    return 2;
  }
  int device = 0, tokens = 128, requests = 8;
  std::string conc = "1,4,16,64", modes = "serial,coalesce,async";
  for (int i = 2; i + 1 < argc; i += 2) {
    const std::string k = argv[i];
    if (k == "--device") device = atoi(argv[i + 1]);
    else if (k == "--tokens") tokens = atoi(argv[i + 1]);
    else if (k == "--requests") requests = atoi(argv[i + 1]);
    else if (k == "--conc") conc = argv[i + 1];
    else if (k == "--modes") modes = argv[i + 1];
    else { fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
  }
  kkx_ctx* ctx = nullptr;
  if (kkx_create(argv[1], device, &ctx) != KKX_OK) { fprintf(stderr, "kkx_create: %s\n", kkx_last_error(nullptr)); return 1; }
  // (a new session runs the tensor-core configuration, "precision" = 1)
  for (const std::string& mode : split(modes)) {
    const bool use_async = mode == "async";
    kkx_set_option(ctx, "coalesce", mode == "coalesce" ? 64 : 0);
    run(ctx, 4, 2, tokens, use_async, 99000);   // warm-up: arenas, pinned pool, kernel attributes
    for (const std::string& cs : split(conc)) {
      const int c = atoi(cs.c_str());
      if (c < 1) continue;
      const int64_t b0 = kkx_get_stat(ctx, use_async ? "async_batches" : "coalesced_batches");
      const int64_t r0 = kkx_get_stat(ctx, use_async ? "async_requests" : "coalesced_requests");
      const Row r = run(ctx, c, requests, tokens, use_async, 3000);
      const int64_t nb = kkx_get_stat(ctx, use_async ? "async_batches" : "coalesced_batches") - b0;
      const int64_t nr = kkx_get_stat(ctx, use_async ? "async_requests" : "coalesced_requests") - r0;
      printf("{\"mode\": \"%s\", \"concurrency\": %d, \"requests\": %ld, \"failures\": %ld, \"audio_s\": %.2f, \"wall_s\": %.4f, "
             "\"audio_s_per_s\": %.1f, \"first_audio_p50_ms\": %.2f, \"first_audio_p95_ms\": %.2f, \"mean_batch\": %.2f}\n",
             mode.c_str(), c, r.requests, r.failures, r.audio_s, r.wall_s, r.audio_s / r.wall_s, r.p50, r.p95,
             nb > 0 ? (double)nr / nb : 1.0);
      fflush(stdout);
    }
  }
  kkx_destroy(ctx);
  return 0;
}

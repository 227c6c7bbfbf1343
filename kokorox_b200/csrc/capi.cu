// capi.cu -- the extern "C" boundary (include/kkx.h).  Every entry point catches all C++
// exceptions and converts them to a negative return code + message: nothing unwinds or aborts
// across the ABI (the reference `expect`s on bad outputs, ort_koko.rs:82,86, and its CLI turns a
// panic into abort(); a library must not).
#include "../../include/kkx.h"
#include "model.h"
#include <cstring>
#include <mutex>
#include <map>
#include <set>
#include <sstream>

using namespace kkx;

struct kkx_ctx {
  std::unique_ptr<Model> model;
  std::mutex mu;  // single-flight like the reference's Mutex<Session> (ort_koko.rs:14,77-78)
  std::string err;
  std::map<float*, size_t> pinned;       // audio buffers handed out -> capacity (floats)
  std::vector<std::pair<float*, size_t>> pinned_free;  // returned buffers kept for reuse
};

static thread_local std::string g_err;

template <class F>
static int guarded(kkx_ctx* ctx, F&& f) {
  std::string msg;
  int rc = KKX_OK;
  try {
    f();
    return KKX_OK;
  } catch (const ArgError& e) { rc = KKX_ERR_ARG; msg = e.what();
  } catch (const IoError& e) { rc = KKX_ERR_IO; msg = e.what();
  } catch (const CudaError& e) { rc = KKX_ERR_CUDA; msg = e.what();
  } catch (const std::exception& e) { rc = KKX_ERR_CUDA; msg = e.what();
  } catch (...) { rc = KKX_ERR_CUDA; msg = "unknown error"; }
  g_launch_stats = nullptr;
  g_dry_run = false;
  if (ctx) ctx->err = msg;
  g_err = msg;
  return rc;
}

extern "C" {

KKX_API const char* kkx_version(void) { return "kkx 0.1 sm_100a"; }

KKX_API int kkx_init(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    g_err = std::string("no CUDA device available: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
            " (the B200 backend has no CPU fallback)";
    cudaGetLastError();
    return KKX_ERR_NO_DEVICE;
  }
  return KKX_OK;
}

KKX_API int kkx_create(const char* weights_path, int device_ordinal, kkx_ctx** out) {
  if (out) *out = nullptr;
  if (!weights_path || !out) { g_err = "kkx_create: null argument"; return KKX_ERR_ARG; }
  int rc = kkx_init();
  if (rc != KKX_OK) return rc;
  kkx_ctx* ctx = new (std::nothrow) kkx_ctx();
  if (!ctx) { g_err = "out of memory"; return KKX_ERR_CUDA; }
  rc = guarded(nullptr, [&] { ctx->model.reset(new Model(weights_path, device_ordinal)); });
  if (rc != KKX_OK) { delete ctx; return rc; }
  *out = ctx;
  return KKX_OK;
}

KKX_API void kkx_destroy(kkx_ctx* ctx) {
  if (!ctx) return;
  try {
    if (ctx->model) cudaSetDevice(ctx->model->device());
    for (auto& p : ctx->pinned) cudaFreeHost(p.first);
    for (auto& p : ctx->pinned_free) cudaFreeHost(p.first);
    ctx->model.reset();
  } catch (...) {}
  delete ctx;
}

KKX_API const char* kkx_last_error(const kkx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

static int check_ctx(kkx_ctx* ctx) {
  if (!ctx) { g_err = "null ctx"; return KKX_ERR_ARG; }
  if (!ctx->model) { ctx->err = "Session is not initialized."; return KKX_ERR_STATE; }
  return KKX_OK;
}

// shared tail of the infer entry points: run the staged batch and hand out a pooled pinned buffer
static float* take_pinned(kkx_ctx* ctx, size_t need_floats, size_t* cap_out) {
  float* host = nullptr;
  size_t cap = 0;
  for (size_t i = 0; i < ctx->pinned_free.size(); i++)
    if (ctx->pinned_free[i].second >= need_floats) {
      host = ctx->pinned_free[i].first; cap = ctx->pinned_free[i].second;
      ctx->pinned_free.erase(ctx->pinned_free.begin() + i);
      break;
    }
  if (!host) {
    cap = std::max<size_t>(need_floats + need_floats / 8, 1024);
    KKX_CUDA(cudaMallocHost(&host, cap * sizeof(float)));
  }
  *cap_out = cap;
  return host;
}

KKX_API void kkx_release(kkx_ctx* ctx, float* audio);

KKX_API int kkx_infer_batch(kkx_ctx* ctx, int32_t batch, const int64_t* tokens, const int32_t* tok_offsets,
                    const float* styles, const float* speeds, float** out_audio,
                    int64_t* out_sample_offsets, int32_t* out_pred_dur) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (out_audio) *out_audio = nullptr;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded(ctx, [&] {
    if (!out_audio) throw ArgError("out_audio is null");
    Model& m = *ctx->model;
    m.stage(batch, tokens, tok_offsets, styles, speeds);
    // the pinned result buffer is handed to the model as a sink: every frame group's audio is copied out on a
    // second stream while the next group computes
    float* host = nullptr;
    size_t cap = 0;
    m.host_sink = [&](long long n) { host = take_pinned(ctx, (size_t)n, &cap); return host; };
    try {
      m.run();
      m.host_sink = nullptr;
      const long long n = m.total_samples();
      if (!host) host = take_pinned(ctx, (size_t)n, &cap);
      m.fetch(m.sink_filled() ? nullptr : host, n, out_sample_offsets, out_pred_dur);
    } catch (...) { m.host_sink = nullptr; if (host) cudaFreeHost(host); throw; }
    ctx->pinned[host] = cap;
    *out_audio = host;
  });
}

KKX_API int kkx_load_voices(kkx_ctx* ctx, const float* table, int32_t n_voices) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded(ctx, [&] { ctx->model->load_voices(table, n_voices); });
}

KKX_API int kkx_infer_batch_voices(kkx_ctx* ctx, int32_t batch, const int64_t* tokens, const int32_t* tok_offsets,
                           const int32_t* mix_offsets, const int32_t* voice_ids, const float* voice_portions,
                           const int32_t* style_rows, const float* speeds, float** out_audio,
                           int64_t* out_sample_offsets, int32_t* out_pred_dur) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (out_audio) *out_audio = nullptr;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded(ctx, [&] {
    if (!out_audio) throw ArgError("out_audio is null");
    Model& m = *ctx->model;
    m.stage_voices(batch, tokens, tok_offsets, mix_offsets, voice_ids, voice_portions, style_rows, speeds);
    m.run();
    const long long n = m.total_samples();
    size_t cap = 0;
    float* host = take_pinned(ctx, (size_t)n, &cap);
    try {
      m.fetch(host, n, out_sample_offsets, out_pred_dur);
    } catch (...) { cudaFreeHost(host); throw; }
    ctx->pinned[host] = cap;
    *out_audio = host;
  });
}

KKX_API int kkx_infer_batch_pcm16(kkx_ctx* ctx, int32_t batch, const int64_t* tokens, const int32_t* tok_offsets,
                          const float* styles, const float* speeds, int16_t** out_pcm,
                          int64_t* out_sample_offsets, int32_t* out_pred_dur) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (out_pcm) *out_pcm = nullptr;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded(ctx, [&] {
    if (!out_pcm) throw ArgError("out_pcm is null");
    Model& m = *ctx->model;
    m.stage(batch, tokens, tok_offsets, styles, speeds);
    m.set_pcm16(true);
    try { m.run(); } catch (...) { m.set_pcm16(false); throw; }
    m.set_pcm16(false);
    const long long n = m.total_samples();
    size_t cap = 0;
    float* host = take_pinned(ctx, (size_t)(n + 1) / 2, &cap);   // n int16 = n/2 floats of the same pool
    try {
      m.fetch_pcm16(reinterpret_cast<short*>(host), n, out_sample_offsets, out_pred_dur);
    } catch (...) { cudaFreeHost(host); throw; }
    ctx->pinned[host] = cap;
    *out_pcm = reinterpret_cast<int16_t*>(host);
  });
}

KKX_API void kkx_release_pcm16(kkx_ctx* ctx, int16_t* pcm) { kkx_release(ctx, reinterpret_cast<float*>(pcm)); }

KKX_API int kkx_infer(kkx_ctx* ctx, const int64_t* tokens, int32_t n_tokens, const float* style256,
              float speed, float** out_audio, int64_t* out_samples, int32_t* out_pred_dur) {
  if (out_samples) *out_samples = 0;
  const int32_t offs[2] = {0, n_tokens};
  int64_t soff[2] = {0, 0};
  const int rc = kkx_infer_batch(ctx, 1, tokens, offs, style256, &speed, out_audio, soff, out_pred_dur);
  if (rc == KKX_OK && out_samples) *out_samples = soff[1];
  return rc;
}

KKX_API void kkx_release(kkx_ctx* ctx, float* audio) {
  if (!ctx || !audio) return;
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->pinned.find(audio);
  if (it != ctx->pinned.end()) {
    if (ctx->pinned_free.size() < 4) {
      ctx->pinned_free.push_back({it->first, it->second});
    } else {
      if (ctx->model) cudaSetDevice(ctx->model->device());
      cudaFreeHost(audio);
    }
    ctx->pinned.erase(it);
  }
}

KKX_API int kkx_stage_batch(kkx_ctx* ctx, int32_t batch, const int64_t* tokens, const int32_t* tok_offsets,
                    const float* styles, const float* speeds) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded(ctx, [&] { ctx->model->stage(batch, tokens, tok_offsets, styles, speeds); });
}

KKX_API int kkx_run_staged(kkx_ctx* ctx, int64_t* out_total_samples, int64_t* out_launches) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded(ctx, [&] {
    ctx->model->run();
    if (out_total_samples) *out_total_samples = ctx->model->total_samples();
    if (out_launches) *out_launches = ctx->model->stats.launches;
  });
}

KKX_API int kkx_fetch_staged(kkx_ctx* ctx, float* dst_audio, int64_t capacity, int64_t* out_sample_offsets,
                     int32_t* out_pred_dur) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded(ctx, [&] { ctx->model->fetch(dst_audio, capacity, out_sample_offsets, out_pred_dur); });
}

KKX_API int kkx_set_option(kkx_ctx* ctx, const char* key, int64_t value) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded(ctx, [&] {
    if (!key) throw ArgError("null key");
    Options& o = ctx->model->opt;
    const std::string k(key);
    if (k == "precision") { if (value < 0 || value > 1) throw ArgError("precision must be 0 or 1"); o.precision = (int)value; }
    else if (k == "noise_seed") o.noise_seed = (unsigned long long)value;
    else if (k == "max_frames") { if (value < 1) throw ArgError("max_frames must be >= 1"); o.max_frames = (int)value; }
    else if (k == "stft_replicate") o.stft_replicate = value ? 1 : 0;
    else throw ArgError("unknown option: " + k);
  });
}

KKX_API int64_t kkx_get_stat(kkx_ctx* ctx, const char* key) {
  if (check_ctx(ctx) || !key) return -1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  const std::string k(key);
  if (k == "launches") return ctx->model->stats.launches;
  if (k == "last_frames") return ctx->model->last_frames;
  if (k == "gpu_us") return (int64_t)ctx->model->last_gpu_us;
  if (k == "precision") return ctx->model->opt.precision;
  return -1;
}

KKX_API int kkx_profile_enable(kkx_ctx* ctx, int enable) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->model->stats.profile = enable != 0;
  return KKX_OK;
}

KKX_API int64_t kkx_profile_json(kkx_ctx* ctx, char* buf, int64_t capacity) {
  if (check_ctx(ctx)) return -1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  std::ostringstream os;
  os << "{\"conv_flops\": " << ctx->model->stats.conv_flops << ", \"arb_flops\": " << ctx->model->stats.arb_flops
     << ", \"arb_bytes\": " << ctx->model->stats.arb_bytes << ", \"gpu_us\": " << ctx->model->last_gpu_us
     << ", \"kernels\": {";
  bool first = true;
  for (auto& kv : ctx->model->prof_) {
    if (!first) os << ", ";
    first = false;
    os << "\"" << kv.first << "\": [" << kv.second.first << ", " << kv.second.second << "]";
  }
  os << "}}";
  const std::string s = os.str();
  if (buf && capacity > 0) {
    const size_t n = std::min<size_t>(s.size(), (size_t)capacity - 1);
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (int64_t)s.size();
}

KKX_API int kkx_set_noise(kkx_ctx* ctx, const float* noise, int64_t n) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded(ctx, [&] { ctx->model->set_noise(noise, n); });
}

KKX_API int kkx_set_inject(kkx_ctx* ctx, const char* name, const void* data, int64_t count) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded(ctx, [&] {
    if (!name) throw ArgError("null name");
    ctx->model->set_inject(name, count > 0 ? data : nullptr, count);
  });
}

KKX_API int kkx_debug_enable(kkx_ctx* ctx, int enable) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->model->set_debug(enable != 0);
  return KKX_OK;
}

KKX_API int64_t kkx_debug_stage(kkx_ctx* ctx, const char* name, int32_t item, float* dst, int64_t capacity,
                        int64_t* rows, int64_t* cols) {
  if (check_ctx(ctx) || !name) return -1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  const DebugStage* s = ctx->model->debug_stage(name, item);
  if (!s) { ctx->err = std::string("no such debug stage: ") + name; return -1; }
  if (rows) *rows = s->rows;
  if (cols) *cols = s->cols;
  const int64_t n = (int64_t)s->data.size();
  if (dst && capacity > 0) memcpy(dst, s->data.data(), (size_t)std::min(n, capacity) * sizeof(float));
  return n;
}

}  // extern "C"

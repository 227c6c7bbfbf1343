/*
 * kkx.h -- C ABI of the B200-native Kokoro-82M inference backend.
 *
 * Drop-in boundary for the ONNX-Runtime session wrapper of byteowlz/kokorox
 * (/root/reference/kokorox/src/onn/).  Each entry point names the reference interface it
 * replaces.  Plain pointers and sizes only; every function is callable from Rust FFI
 * (`extern "C"`), cgo, or ctypes.  No function aborts or throws across the ABI: all failures
 * are a negative return code plus a message retrievable with kkx_last_error().
 *
 * Thread safety: a kkx_ctx may be shared between threads (the reference shares one
 * `Arc<OrtKoko>` between tokio workers, ort_koko.rs:13-18, koko.rs:36-45); calls on one ctx
 * are serialised internally, exactly like the reference's `Mutex<Session>` (ort_koko.rs:77-78).
 * Use one ctx per GPU for multi-GPU request sharding.
 */
#ifndef KKX_H_
#define KKX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kkx_ctx kkx_ctx;

#if defined(__GNUC__)
#define KKX_API __attribute__((visibility("default")))
#else
#define KKX_API
#endif

#define KKX_OK 0
#define KKX_ERR_ARG (-1)      /* bad argument (null pointer, token id out of 0..177, n_tokens > 512, speed outside [0.1, 10]) */
#define KKX_ERR_IO (-2)       /* model file missing / malformed / tensor missing or unresolved  */
#define KKX_ERR_CUDA (-3)     /* CUDA runtime error (message has the cudaError string)          */
#define KKX_ERR_NO_DEVICE (-4)/* no CUDA device / device ordinal out of range -- there is NO CPU fallback */
#define KKX_ERR_STATE (-5)    /* ctx not initialised ("Session is not initialized.", ort_koko.rs:88-90), or a call
                               * out of sequence (run before stage, fetch before run, ticket of a destroyed ctx) */

#define KKX_STYLE_DIM 256     /* ort_koko.rs:61-65: style tensor [B,256]                        */
#define KKX_MAX_TOKENS 512    /* ALBERT max_position_embeddings; koko.rs:781-783 chunks to <=500 */
#define KKX_SAMPLE_RATE 24000 /* koko.rs:59                                                     */

/* Replaces kokorox::onn::init_ort(dylib_path) (onn/mod.rs:19-49): process-wide one-time init.
 * Checks that a CUDA driver and at least one sm_100 device are present.  Optional: kkx_create
 * performs the same check.  Returns KKX_OK or KKX_ERR_NO_DEVICE. */
KKX_API int kkx_init(void);

/* Replaces OrtKoko::new(model_path) -> OrtBase::load_model (ort_koko.rs:31-35,
 * ort_base.rs:14-39, `commit_from_file(model_path)`): loads the model file onto GPU `device_ordinal` and builds
 * every derived weight layout.  `weights_path` is what the reference passes -- the downloaded ONNX file
 * (kokoro-v1.0.onnx, onnx/model.onnx or one of the fp16 / q8 / uint8 / q4 variants of hf_cache.rs:135-144; the
 * initialisers are read straight from the protobuf, named through the node graph and dequantised at load) -- or a
 * KKXW file written by kokorox_b200/weightfile.py / convert.py.  The format is detected from the file's magic.
 * Sessions created from the same (file, device) share one device-resident weight set (the reference's servers hold
 * two or three sessions of one model, koko main.rs:1477,1596).  A new session runs the benchmarked configuration
 * ("precision" = 1, see kkx_set_option).  On failure *out is NULL and kkx_last_error(NULL) has the reason; for an
 * ONNX file that does not resolve it lists the tensors that could not be found. */
KKX_API int kkx_create(const char* weights_path, int device_ordinal, kkx_ctx** out);

/* Host-only helper (no GPU needed): reads `src_path` exactly as kkx_create would (ONNX or KKXW) and writes the
 * recovered Kokoro-82M state dict as a KKXW file, e.g. to convert a downloaded model_q8f16.onnx once.
 * *out_tensors (nullable) receives the tensor count. */
KKX_API int kkx_convert_model_file(const char* src_path, const char* dst_kkxw_path, int32_t* out_tensors);

/* Drop of OrtKoko (koko.rs:1338-1375 `cleanup`): frees all device and pinned host memory. */
KKX_API void kkx_destroy(kkx_ctx* ctx);

/* Error string of the CALLING THREAD's last failed call (the ctx argument is accepted for symmetry and may be
 * NULL): thread-local, so threads sharing one ctx never see -- or race with -- each other's messages; read it
 * on the thread that got the error code.  The Rust shim turns rc<0 into Err(kkx_last_error()), matching
 * the Result<_, String> / Box<dyn Error> returns of ort_koko.rs:31,42. */
KKX_API const char* kkx_last_error(const kkx_ctx* ctx);

/* Replaces OrtKoko::infer(tokens, styles, speed) at B=1 (ort_koko.rs:37-91; call site
 * koko.rs:1168-1177).
 *   tokens     [n_tokens] i64, INCLUDING the leading and trailing 0 pad (koko.rs:1168-1173);
 *              2 <= n_tokens <= 512, ids in 0..177 (tts/vocab.rs:5-20)   -- "input_ids"
 *   style256   [256] f32 (koko.rs:1255-1306 mix_styles row)             -- "style"
 *   speed      in [0.1, 10] (durations are 50/speed frames per token at most) -- "speed"
 *   out_audio  receives a library-owned pinned host buffer of *out_samples f32 (24 kHz mono,
 *              = 600 * sum(pred_dur)); valid until kkx_release(ctx, ptr) -- outputs[0]
 *   out_pred_dur  nullable; [n_tokens] predicted integer frame durations (not observable
 *              through the reference graph; exposed for parity tests).
 * Thread-safe.  By default concurrent callers run one after another, like the reference's
 * Mutex<Session> (ort_koko.rs:14,77).  With option "coalesce" = K > 1 the callers that are waiting while a
 * step runs are merged into ONE ragged batch of up to K utterances (the batching queue SURVEY 8f row 4 asks the
 * servers for, openai lib.rs:400-412): each caller still gets exactly the waveform it would get alone (batched
 * and single results are bit-identical), as a view into a shared pinned buffer that is recycled once every
 * caller of that batch has called kkx_release.  A request that fails validation fails alone. */
KKX_API int kkx_infer(kkx_ctx* ctx, const int64_t* tokens, int32_t n_tokens, const float* style256,
              float speed, float** out_audio, int64_t* out_samples, int32_t* out_pred_dur);

/* Ragged batch of independent utterances (new capability; the reference only ever calls infer
 * with B=1, koko.rs:1175, and serialises callers on a mutex).  Item b uses
 * tokens[tok_offsets[b] .. tok_offsets[b+1]), styles[b*256 ..], speeds[b]; result b is
 * (*out_audio)[out_sample_offsets[b] .. out_sample_offsets[b+1]) and equals what kkx_infer
 * returns for that item alone.  out_pred_dur (nullable) is indexed like tokens. */
KKX_API int kkx_infer_batch(kkx_ctx* ctx, int32_t batch, const int64_t* tokens, const int32_t* tok_offsets,
                    const float* styles, const float* speeds, float** out_audio,
                    int64_t* out_sample_offsets, int32_t* out_pred_dur);

/* Returns an audio buffer obtained from kkx_infer / kkx_infer_batch / kkx_wait to the library. */
KKX_API void kkx_release(kkx_ctx* ctx, float* audio);

/* ---- asynchronous form of kkx_infer.  The reference's WebSocket server synthesises and sends
 * sentence by sentence, `for sentence ... { synthesize_and_send_chunk(...).await }` (kokorox-websocket/src/lib.rs:
 * 371-376), so sentence k+1 cannot start while k is on the wire, and its servers block one thread per request on
 * Mutex<Session> (openai lib.rs:400-412, ort_koko.rs:77).
 *   kkx_submit  copies the request (the caller's buffers are free again on return), queues it and returns a
 *               ticket at once; a worker thread owned by the ctx runs whatever is queued as ragged batches of up
 *               to "async_batch" utterances (each result is bit-identical to the caller's own kkx_infer).
 *   kkx_poll    1 = finished (kkx_wait will not block), 0 = still queued / running, <0 = unknown ticket.
 *   kkx_wait    blocks until the request has finished and redeems the ticket (once): same outputs as kkx_infer;
 *               out_pred_dur (nullable) must hold n_tokens entries.  A failed request returns its own error code
 *               and message here.  Release the audio with kkx_release. */
typedef int64_t kkx_ticket;
KKX_API int kkx_submit(kkx_ctx* ctx, const int64_t* tokens, int32_t n_tokens, const float* style256, float speed,
                       kkx_ticket* out_ticket);
KKX_API int kkx_poll(kkx_ctx* ctx, kkx_ticket ticket);
KKX_API int kkx_wait(kkx_ctx* ctx, kkx_ticket ticket, float** out_audio, int64_t* out_samples,
                     int32_t* out_pred_dur);

/* ---- "next" rows of the hot-path scope: the host work either side of the call.
 *
 * kkx_load_voices replaces TTSKoko::load_voices (koko.rs:1308-1334): uploads the voice table once,
 *   table [n_voices][511][256] f32 (row = un-padded token count, koko.rs:1262).
 * kkx_infer_batch_voices replaces TTSKoko::mix_styles + OrtKoko::infer (koko.rs:1161-1180, 1255-1306): like
 *   kkx_infer_batch, but the style of item b is mixed ON THE DEVICE from the table:
 *     style_b[j] = sum over i in [mix_offsets[b], mix_offsets[b+1]) of table[voice_ids[i]][style_rows[b]][j] * voice_portions[i]
 *   in that order with separate multiply and add, i.e. bit-identical to the reference loop (koko.rs:1296-1302).
 *   A single voice ("af_sky") is one entry with portion 1.0; a mix "af_sky.4+af_nicole.5" is two entries with
 *   portions 0.4, 0.5 (portion = weight * 0.1, NOT renormalised, koko.rs:1283). */
KKX_API int kkx_load_voices(kkx_ctx* ctx, const float* table, int32_t n_voices);
KKX_API int kkx_infer_batch_voices(kkx_ctx* ctx, int32_t batch, const int64_t* tokens, const int32_t* tok_offsets,
                           const int32_t* mix_offsets, const int32_t* voice_ids, const float* voice_portions,
                           const int32_t* style_rows, const float* speeds, float** out_audio,
                           int64_t* out_sample_offsets, int32_t* out_pred_dur);

/* kkx_infer_batch_pcm16: like kkx_infer_batch, but the waveform comes back as 16-bit PCM,
 *   pcm = trunc(clamp(s, -1, 1) * 32767) -- the f32 -> i16 conversion of the reference's WebSocket server
 *   (kokorox-websocket/src/lib.rs:699-703) fused into the iSTFT kernel; half the device->host bytes.
 *   Release the buffer with kkx_release_pcm16. */
KKX_API int kkx_infer_batch_pcm16(kkx_ctx* ctx, int32_t batch, const int64_t* tokens, const int32_t* tok_offsets,
                          const float* styles, const float* speeds, int16_t** out_pcm,
                          int64_t* out_sample_offsets, int32_t* out_pred_dur);
KKX_API void kkx_release_pcm16(kkx_ctx* ctx, int16_t* pcm);

/* ---- output containers (host-side byte work after the pcm16 / f32 result; no GPU involved, no ctx needed).
 * kkx_wav_header_pcm16: the 44-byte RIFF/WAVE header of the WebSocket server's `encode_audio`
 *   (kokorox-websocket/src/lib.rs:707-731): PCM format 1, mono, 16 bit, sizes filled in from n_samples.
 * kkx_wav_header_f32_stream: the streaming header of utils/wav.rs:19-43: IEEE-float format 3, both size fields
 *   0xFFFFFFFF placeholders; f32 samples follow as little-endian bytes (wav.rs:45-50), i.e. the result buffer as is.
 * kkx_encode_wav16_base64: `encode_audio` end to end for a pcm16 result: base64 (standard alphabet, '=' padded)
 *   of header + samples.  Returns the number of characters (without the terminating NUL, which is written when
 *   capacity allows); writes nothing if capacity is too small, so call with dst = NULL to size the buffer. */
KKX_API int32_t kkx_wav_header_pcm16(uint8_t* dst44, int64_t n_samples, int32_t sample_rate);
KKX_API int32_t kkx_wav_header_f32_stream(uint8_t* dst44, int32_t channels, int32_t sample_rate);
KKX_API int64_t kkx_encode_wav16_base64(const int16_t* pcm, int64_t n_samples, int32_t sample_rate, char* dst,
                                int64_t capacity);

/* ---- device-resident variant (bench.py `value`: inputs already in HBM, output stays in HBM).
 * Stages the batch on the device once; kkx_run_staged() then runs the whole forward with no
 * host<->device payload traffic (only the per-item frame counts cross, 4 bytes per item).
 * Returns total samples in *out_total_samples and kernel launches in *out_launches
 * (either may be NULL). */
KKX_API int kkx_stage_batch(kkx_ctx* ctx, int32_t batch, const int64_t* tokens, const int32_t* tok_offsets,
                    const float* styles, const float* speeds);
KKX_API int kkx_run_staged(kkx_ctx* ctx, int64_t* out_total_samples, int64_t* out_launches);
/* Copies the result of the last kkx_run_staged to host memory. */
KKX_API int kkx_fetch_staged(kkx_ctx* ctx, float* dst_audio, int64_t capacity, int64_t* out_sample_offsets,
                     int32_t* out_pred_dur);

/* ---- options.  Known keys:
 *   "precision"   1 (DEFAULT, the benchmarked configuration) = tcgen05 tensor cores: bf16 decoder + generator,
 *                 split-TF32 predictor; 0 = fp32 SIMT everywhere (verification mode: tightest parity, several
 *                 times slower)
 *   "noise_seed"  seed of the on-device Philox N(0,1) generator for the SineGen noise
 *   "max_frames"  frame budget per frame-phase group (memory control for large batches)
 *   "max_tokens"  token budget per pass of one kkx_infer_batch call (default 40960): larger batches are run in
 *                 several passes and concatenated, so a call may carry any number of utterances
 *   "stft_replicate" 0 = reflect edge padding (upstream STFT), 1 = replicate (conv-STFT export)
 *   "coalesce"    0/1 = off; K > 1 (<= 512) = merge up to K concurrent kkx_infer callers per step
 *   "coalesce_wait_us"  how long the caller that found the queue idle waits for company (default 0: no
 *                 added latency -- batches form from whatever queued up behind the running step)
 *   "async_batch" most kkx_submit requests the worker merges into one ragged batch (default 64)
 *   "latency_graphs" 0/1: replay the token phase of single-utterance calls from a CUDA graph keyed by the token
 *                 count (default 1; captured on the second call with a given count)
 *   "fork_max_batch" largest batch whose independent branches (text encoder | ALBERT, F0 | N, harmonic source |
 *                 decoder) run on two streams (default 4; 0 = never).  Results do not depend on either option.
 *   kernel selection, all default 1 and none of them changes a bit of the result (each names the round-2 kernel it
 *   switches on; 0 selects the kernel it replaced -- for A/B measurements and the bit-identity tests):
 *   "attention_umma" (tcgen05 attention), "split_f16" (fp16 instead of tf32 operand planes; fp32-grade either way, but
 *   not the same bits), "gemm_pair" (split-precision GEMMs on CTA pairs, tcgen05 cta_group::2), "conv_pair" (the
 *   decoder's wide bf16 convs on CTA pairs), "fuse_planes" (LayerNorm / FFN / QKV GEMM write the next kernel's operand
 *   planes directly), "fuse_phases" (ConvTranspose1d phases in one launch), "ups_phase_loop" (default 3: the stage-1
 *   up-sampling conv on persistent CTAs with resident weights; 1 / 2 = one CTA per row tile looping over the phases,
 *   0 = one CTA per (tile, phase); same bits); "lstm_fast_gates" (SFU gate functions in the LSTM recurrence of the
 *   tensor-core configuration; fp32-grade, not the same bits as libm) and "fuse_noise_stats" (the generator's 22 -> 128
 *   noise conv in fp32 together with the statistics of its output instead of bf16 operands on the tensor cores);
 *   "stream_bf16" (default 0) keeps the res-block residual stream in bf16.
 * kkx_get_stat keys: "launches", "last_frames", "gpu_us", "precision", "coalesced_batches",
 * "coalesced_requests", "coalesced_largest", "async_batches", "async_requests", "frame_groups",
 * "group_first:<g>" (first item of frame group g of the last run), "weights_sessions" (sessions sharing this
 * ctx's device weight set), "weights_bytes", "weights_from_onnx", "graph_replays". */
KKX_API int kkx_set_option(kkx_ctx* ctx, const char* key, int64_t value);
KKX_API int64_t kkx_get_stat(kkx_ctx* ctx, const char* key); /* "launches", "last_frames", "gpu_us" ... */

/* Per-kernel device timing of the last run (CUDA events on the library's stream, one after every
 * launch).  kkx_profile_json writes {"conv_flops": F, "gpu_us": T, "kernels": {name: [launches,
 * total_us]}} and returns the full length of the text. */
KKX_API int kkx_profile_enable(kkx_ctx* ctx, int enable);
KKX_API int64_t kkx_profile_json(kkx_ctx* ctx, char* buf, int64_t capacity);

/* Library / build identification, e.g. "kkx 0.1 sm_100a". */
KKX_API const char* kkx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* KKX_H_ */

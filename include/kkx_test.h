/*
 * kkx_test.h -- test hooks of libkkx.so (parity harness only; the Rust shim binds nothing from
 * this header): whole-model hooks (noise / teacher forcing / stage dumps) and op-level entry
 * points.  The op-level calls take host pointers in and out, one ragged item; each allocates
 * device scratch, runs ONE hand-written kernel and copies the result back, so that tests/ can
 * compare a kernel against the matching torch op in isolation.
 */
#ifndef KKX_TEST_H_
#define KKX_TEST_H_
#include "kkx.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---- whole-model hooks (moved here from kkx.h: the parity harness is their only user) -------------------
 * kkx_set_noise: explicit SineGen noise, element (sample t, harmonic h) at noise[t*9+h]; n = number of floats;
 *   n == 0 returns to the on-device generator.  With a batch every item reads the same buffer from offset 0.
 * kkx_set_inject_item: teacher-force an intermediate of batch item `item` (any frame group) in every following
 *   call: name is "pred_dur" (int32 data, n_tokens entries in 1..500), "F0" or "N" (f32 data, 2T entries);
 *   count == 0 clears that item's entry.  kkx_set_inject is the item-0 form.
 * kkx_debug_enable / kkx_debug_select: keep stage tensors alive for kkx_debug_stage -- of every item, or of one
 *   item only (a B = 64 x 510 batch would otherwise hold tens of GB of host copies).
 * kkx_debug_stage: after a call, copy a named stage tensor of item `item` to host.  Returns the number of floats the
 *   stage holds (rows*cols) or <0; rows/cols are written if non-NULL; at most `capacity` floats are copied (dst may
 *   be NULL to query the size).  Stage names match oracle/kokoro_ref.py (`bert`, `d`, `dur_logits`, `F0`, ...). */
KKX_API int kkx_set_noise(kkx_ctx* ctx, const float* noise, int64_t n);
KKX_API int kkx_set_inject(kkx_ctx* ctx, const char* name, const void* data, int64_t count);
KKX_API int kkx_set_inject_item(kkx_ctx* ctx, int32_t item, const char* name, const void* data, int64_t count);
KKX_API int kkx_debug_enable(kkx_ctx* ctx, int enable);
KKX_API int kkx_debug_select(kkx_ctx* ctx, int enable, int32_t item);
KKX_API int64_t kkx_debug_stage(kkx_ctx* ctx, const char* name, int32_t item, float* dst, int64_t capacity,
                        int64_t* rows, int64_t* cols);

/* Generic shifted-GEMM kernel (Conv1d / ConvTranspose1d phase / Linear), see kernels.h ConvArgs.
 * in [rows_in, ldi]; w [ks][Ci][Co]; out [out_rows, Co] (pre-filled by the caller; rows not
 * written keep their value); res [ceil(out_rows >> res_shift), Co] or NULL. */
KKX_API int kkx_test_conv(int device, const float* in, int rows_in, int ldi, int Ci, const float* w,
                          const float* bias, int Co, int ks, int dil, int pad, int stride,
                          const float* pscale, const float* pshift, int pact, float pslope,
                          const float* palpha, int m_len, int ors, int oro, int out_rows,
                          const float* res, int res_rows, int res_shift, float oscale, int eact,
                          int accumulate, float* out);
/* xproj [N,2048], whhT [2][256][1024] -> out [N,512] */
KKX_API int kkx_test_lstm(int device, const float* xproj, const float* whhT, int N, float* out);
/* ragged batch: xproj [rows,2048], items at off[b] with len[b] rows -> out [rows,512] (pre-zeroed) */
KKX_API int kkx_test_lstm_batch(int device, const float* xproj, const float* whhT, int B, const int* off,
                                const int* len, int rows, float* out);
/* the same with an explicit kernel variant (1 = cluster kernel, SFU gate functions; 0 = cluster kernel, libm gate
   functions; -1 = plain one-CTA-per-item kernel) and the average device time of `reps` further launches */
KKX_API int kkx_test_lstm_batch_v(int device, const float* xproj, const float* whhT, int B, const int* off,
                                  const int* len, int rows, int variant, int reps, float* out, float* ms);
/* Conv1d(22, 128, k = 1) fused with the chunk statistics of its output (kernels_signal.cu): x [rows,24] (22 used),
   w [22][128], items at off[b] (len[b] rows) -> out [rows,128] (pre-filled by the caller), part [B][nchunk][2][128] with
   nchunk = ceil(max_len / 128); part_ref receives what launch_colstats computes from `out` */
KKX_API int kkx_test_pointwise_conv_stats(int device, const float* x, const float* w, const float* bias, int B,
                                          const int* off, const int* len, int rows, int max_len, float* out, float* part,
                                          float* part_ref);
/* row LayerNorm of x (+ res) [rows,C] with optional affine w / b [C], per-item AdaLN (ada = gamma [C] then beta [C]),
   LeakyReLU slope (1 = none); optional fp16 hi / lo operand planes of 16 * result (raw half bits) */
KKX_API int kkx_test_layernorm(int device, const float* x, const float* res, const float* w, const float* b,
                               const float* ada, int rows, int C, float eps, float slope, float* out,
                               unsigned short* pl_hi, unsigned short* pl_lo);
/* im2col operand producer of noise_convs[0] (22 of 24 input columns, k = 12, stride 6, pad 3): in [rows_in,24] fp32 with
   items at in_off[b] (in_len[b] rows) -> out [rows_out,Cpad] bf16 bits (pre-filled by the caller) with items at out_off[b]
   (out_len[b] rows) and 32 zeroed gap rows around each; generic = 1 forces the element-wise kernel */
KKX_API int kkx_test_im2col(int device, const float* in, int rows_in, int B, const int* in_off, const int* in_len,
                            const int* out_off, const int* out_len, int rows_out, int max_out_len, int Cpad, int generic,
                            unsigned short* out);
/* qkv [N,2304] -> ctx [N,768] */
KKX_API int kkx_test_attention(int device, const float* qkv, int N, float* ctx);
/* ragged batch through the tcgen05 / TMEM attention kernel (kernels_attn.cu), or the mma.sync kernel (umma = 0):
 * qkv [rows,2304] with items at off[b] (len[b] rows) -> ctx [rows,768] (rows outside the items are left untouched) */
KKX_API int kkx_test_attention_batch(int device, const float* qkv, int B, const int* off, const int* len, int rows,
                                     int umma, float* ctx);
/* x [L,C] + style row (gamma|beta, 2C) -> scale [C], shift [C] (InstanceNorm stats + AdaIN) */
KKX_API int kkx_test_adain_coef(int device, const float* x, int L, int C, const float* gamma_beta,
                                float* scale, float* shift);
/* Tensor-core (tcgen05/TMEM/TMA) shifted GEMM, stride 1: x [L, Ci] fp32 is rounded to bf16 by the
 * operand-producer kernel (optional per-channel scale/shift + activation), w [Co][ks][Ci] fp32 is
 * rounded to bf16 on the host; fp32 accumulation and epilogue.  out [out_rows, Co] pre-filled. */
KKX_API int kkx_test_conv_tc(int device, const float* x, int L, int Ci, const float* w, const float* bias,
                             int Co, int ks, int dil, int pad, const float* pscale, const float* pshift,
                             int pact, float pslope, const float* palpha, int m_len, int ors, int oro,
                             int out_rows, const float* res, int res_rows, int res_shift, float oscale,
                             int accumulate, float* out);
/* Split-TF32 tensor-core GEMM/conv (fp32-grade accuracy on tcgen05 kind::tf32): x [L,Ci], w [Co][ks][Ci]
 * fp32, nprod = 3 or 4 partial products, eact 0 / 3 (gelu_new).  out [L, Co]. */
KKX_API int kkx_test_conv_tf32(int device, const float* x, int L, int Ci, const float* w, const float* bias,
                               int Co, int ks, int dil, int pad, int nprod, int eact, float* out);
/* The same with split-FP16 operand planes ("3xFP16": fp16 hi/lo planes, kind::f16 MMAs, power-of-two scaling undone in
 * the epilogue); problems with >= 148 output tiles take the persistent kernel, smaller ones the single-tile kernel. */
KKX_API int kkx_test_conv_f16x3(int device, const float* x, int L, int Ci, const float* w, const float* bias,
                                int Co, int ks, int dil, int pad, int eact, float* out);
/* The same with the kernel chosen by the caller -- 1 = single-tile, 2 = persistent, 3 = persistent CTA-pair
 * (tcgen05 cta_group::2) -- plus an optional residual [L, Co] and output scale: out = (conv + bias (+act) + res) * oscale.
 * The three kernels accumulate every output element in the same order; tests compare them bit for bit. */
KKX_API int kkx_test_conv_f16x3_k(int device, const float* x, int L, int Ci, const float* w, const float* bias,
                                  int Co, int ks, int dil, int pad, int eact, int kernel, const float* res, float oscale,
                                  float* out);
/* Fused generator res-block conv (kernels_arb.cu): out = (conv1d(snake(x*scale_b+shift_b), w, dilation) + bias + res)
 * * oscale (+ out when accumulate); B ragged items packed along rows (x, res, out: [sum lens, C]); C in
 * {128, 256}; w [C][ks][C]; x optionally rounded to bf16 first (in_bf16); result as fp32 or, want_bf16, the
 * unscaled bf16 output widened to fp32.  sums (nullable) [B][2][C] = column sums of (conv+bias+res) and of
 * its square.  desc_mode selects the UMMA base-offset convention for row-shifted operand views. */
KKX_API int kkx_test_arb_conv(int device, const float* x, int B, const int* lens, int C, int in_bf16,
                              const float* scale, const float* shift, const float* alpha, const float* w,
                              const float* bias, int ks, int dil, const float* res, float oscale,
                              int accumulate, int want_bf16, float* out, float* sums, int desc_mode);
/* Host-only: the library's tensor spec table (onnx_loader.cu kokoro_tensor_specs) as text lines "name d0 d1 ...\n";
 * returns the full length, writes at most capacity-1 characters + NUL. */
KKX_API int64_t kkx_test_tensor_specs(char* buf, int64_t capacity);
/* The same kernel on a bf16 residual stream: x and res are rounded to bf16 first; res == NULL is conv1 (bf16 x -> bf16
 * out), res != NULL is conv2 (bf16 residual -> the next bf16 x when want_bf16, else the fp32 block output). */
KKX_API int kkx_test_arb_conv_stream(int device, const float* x, int B, const int* lens, int C, const float* scale,
                                     const float* shift, const float* alpha, const float* w, const float* bias, int ks,
                                     int dil, const float* res, float oscale, int accumulate, int want_bf16, float* out,
                                     float* sums);
KKX_API const char* kkx_test_last_error(void);

#ifdef __cplusplus
}
#endif
#endif

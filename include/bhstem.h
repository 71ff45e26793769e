/* bhstem.h -- C ABI of libbhstem.so: the encoder's convolutional stem on B200 (sm_100a), the
 * consumer of the log-mel frames libbhmel.so produces.
 *
 * Replaces, in the reference encoder (ref: osuT5/osuT5/model/custom_transformers/
 * modeling_ropewhisper.py:1135-1136 construction, :1206-1209 forward; the stock HF WhisperEncoder
 * has the same stem):
 *
 *     inputs_embeds = gelu(conv1(input_features))      conv1 = Conv1d(C_in, D, kernel 3, padding 1)
 *     inputs_embeds = gelu(conv2(inputs_embeds))       conv2 = Conv1d(D, D, kernel 3, stride 2, padding 1)
 *     inputs_embeds = inputs_embeds.permute(0, 2, 1)   -> [B, T/2, D]
 *
 * as two implicit GEMMs on the tcgen05 tensor cores (bf16 operands, fp32 accumulation in tensor
 * memory), operands staged by TMA straight from the channels-last activations -- the three taps
 * are three shifted views of the same matrix, the zero padding is TMA's out-of-bounds fill -- with
 * bias, GELU and the bf16 conversion fused into the epilogue.  The input is the channels-last
 * encoder input [B][T][C_in] that bhmel_forward_encoder_input (BHMEL_LAYOUT_BTC) writes, so the
 * reference's swapaxes before the stem and permute after it both disappear.
 *
 * Arithmetic contract (what the reference computes under bf16): per output element, fp32
 * accumulation of bf16 products plus the bias, rounded to bf16 (the convolution's output), then the
 * erf GELU evaluated in fp32 on that bf16 value, rounded to bf16.  erf is Abramowitz-Stegun 7.1.26
 * (|error| <= 1.5e-7); because GELU's input is always a bf16 value the result is checked for ALL
 * 65 536 inputs: identical to torch's fp32 erf GELU after rounding, except a handful of inputs in the
 * tail x <= -3.5 where |difference| <= 4e-6 (tests/test_oracle.py, tests/test_gpu_stem.py).
 *
 * Plain C: pointers, sizes, a CUDA stream.  The caller owns every device buffer; the handle owns
 * the packed weights.  Launches on the caller's stream, never synchronises, allocates nothing in
 * bhstem_forward.  Every int-returning function returns 0 on success, a BHSTEM_E* code otherwise;
 * message through bhstem_last_error() (thread-local).
 */
#ifndef BHSTEM_H_
#define BHSTEM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BHSTEM_VERSION 2   /* 2: bhstem_prepare_split / bhstem_forward_split, options 3-5 */

#define BHSTEM_OK 0
#define BHSTEM_EINVAL 1
#define BHSTEM_ECUDA 2
#define BHSTEM_EDEVICE 3

typedef struct bhstem_handle bhstem_handle;

/* Builds a stem for C_in input channels and D model channels on the CURRENT device from HOST
 * float32 parameters in torch's Conv1d layout (state-dict keys conv1.weight [D][C_in][3],
 * conv1.bias [D], conv2.weight [D][D][3], conv2.bias [D]; ref: modeling_ropewhisper.py:1135-1136).
 * Weights and biases are rounded to bf16 (what `.to(bfloat16)` of the model does, ref:
 * inference.py:486-489); the weights are repacked tap-major [3][D][C].  Requires C_in % 8 == 0 and D % 128 == 0. */
int bhstem_create(int32_t c_in, int32_t d_model, const float* conv1_weight, const float* conv1_bias,
                  const float* conv2_weight, const float* conv2_bias, bhstem_handle** out);
void bhstem_destroy(bhstem_handle* h);

/* x       DEVICE bf16 [B][T][C_in], channels last, contiguous (T even, T >= 2)
 * hidden  DEVICE bf16 [B][T][D] scratch for gelu(conv1) (caller-owned; stays L2-resident for
 *         inference-sized batches)
 * y       DEVICE bf16 [B][T/2][D] = gelu(conv2(gelu(conv1(x^T))))^T
 * Two kernel launches on `stream`. */
int bhstem_forward(bhstem_handle* h, const void* x, int64_t B, int64_t T, void* hidden, void* y, void* stream);

/* conv1 or conv2 alone (stage 1 / 2): for tests and profiling.  Stage 1: in [B][T][C_in] ->
 * out [B][T][D]; stage 2: in [B][T][D] -> out [B][T/2][D]. */
int bhstem_forward_stage(bhstem_handle* h, int32_t stage, const void* in, int64_t B, int64_t T, void* out,
                         void* stream);

/* Split conv1 (SURVEY.md 8f N1 + N3 fused: the encoder input is never materialised).
 *
 * The reference feeds conv1 the concatenation [mel | cond] where the conditioning embeddings are ONE vector per
 * window repeated over its T frames (ref: osuT5/osuT5/model/modeling_mapperatorinator.py:368-370,
 * `c.unsqueeze(1).expand((-1, frames.shape[1], -1))`, then swapaxes :375-376 and the encoder's conv1,
 * modeling_ropewhisper.py:1206).  A convolution is linear in its input channels, so the time-constant channels
 * contribute the same three per-tap sums S_tap[b][n] = sum_c W[n][n_var + c][tap] * cond[b][c] at every frame
 * (two of them at a window's first and last frame, where one tap reads the zero padding).  bhstem_forward_split
 * evaluates those sums once per (window, output channel) -- the same bf16 x bf16 products, accumulated in fp32 --
 * adds them to the bias, and runs the tcgen05 implicit GEMM over the n_var time-varying channels only:
 *     3 * n_var instead of 3 * c_in products per output element (80 of 464 channels at the reference's dims)
 * with the same epilogue (fp32 accumulator + bias -> bf16 -> GELU -> bf16).  Results equal bhstem_forward on the
 * materialised [B][T][c_in] input up to the fp32 summation order (same tolerance as against the reference's
 * cuDNN convolution; tests/test_gpu_stem.py).
 *
 * bhstem_prepare_split: once per handle, before the first bhstem_forward_split (allocates the repacked
 * [3][D][n_var] weights; synchronous; not to be raced with launches on the same handle).  n_var % 8 == 0,
 * 8 <= n_var < c_in. */
int bhstem_prepare_split(bhstem_handle* h, int32_t n_var);
/* x_var   DEVICE bf16 [B][T][n_var], channels last, contiguous: the time-varying channels (the log-mel frames,
 *         e.g. written by bhmel_forward_ex with BHMEL_OUT_BF16)
 * cond    DEVICE bf16 [B][c_in - n_var]: the concatenated conditioning embeddings of each window
 * bias3   DEVICE float32 [B][3][D] scratch owned by the caller (interior / first-frame / last-frame bias)
 * hidden  DEVICE bf16 [B][T][D] scratch, y DEVICE bf16 [B][T/2][D] as for bhstem_forward.
 * Three kernel launches on `stream` (folded bias, conv1, conv2). */
int bhstem_forward_split(bhstem_handle* h, const void* x_var, const void* cond, int64_t B, int64_t T, float* bias3,
                         void* hidden, void* y, void* stream);

/* Kernel schedule of this handle (A/B experiments and wider models; all variants give the same results
 * within the bf16 tolerance).  The choice is explicit -- no environment variable is read by the library. */
#define BHSTEM_OPT_VARIANT 1
#define BHSTEM_VARIANT_TAP_BOXES 0    /* one TMA box per (tap, channel step) */
#define BHSTEM_VARIANT_SHARED_TAPS 1  /* one CTA per tile: one staged block, three row-shifted descriptors */
#define BHSTEM_VARIANT_CTA_PAIRS 2    /* default: tcgen05.mma.cta_group::2 CTA pairs where d_model % 256 == 0 and the SM
                                         count is even (else the shared-tap kernel runs) */
/* 1 (default): kernels are launched with programmatic stream serialisation, so a kernel's prologue (barrier
 * init, TMEM allocation) overlaps the tail of the previous kernel in the stream; every global access still
 * waits for that kernel to complete (griddepcontrol.wait).  0: plain stream order. */
#define BHSTEM_OPT_PDL 2
/* Epilogue warps of the CTA-pair kernel, one byte per stage: bits 0-7 conv1, 8-15 conv2, 16-23 the split conv1;
 * each 8 (two warps per TMEM lane quarter, 8-stage weight ring) or 16 (four per quarter, 6-stage ring).  Default
 * 8 / 8 / 16: the full convolutions are bound by the tensor pipe, the split conv1 (15 MMA steps per tile) by the
 * epilogue's GELU.  Same results bit for bit. */
#define BHSTEM_OPT_EPILOGUE_WARPS 3
/* 1 (default): a launch with so few 256-column tiles that half the SMs would idle (one window of conv2: 48 tiles
 * on 148 SMs) runs 128-column tiles on the one-CTA kernel instead -- twice the tiles, half the work each, same
 * bits.  0: always the handle's tile width. */
#define BHSTEM_OPT_SMALL_BATCH_TILES 4
/* Ring depths of the CTA-pair kernel with 8 epilogue warps, a 3-bit mask (bit 0 conv1, bit 1 conv2, bit 2 split
 * conv1 when it runs with 8 epilogue warps): 1 = 3 activation stages + 6 weight stages, 0 = 2 + 8 (same shared
 * memory; conv1, which stages only the 136-row block, gets 5 / 3 activation stages out of the same space).  Default
 * 3: the third activation stage covers the load latency of the streamed operand (conv2 -6 %, conv1 -4..8 %).  Same
 * bits. */
#define BHSTEM_OPT_DEEP_A_RING 5
int bhstem_set_option(bhstem_handle* h, int32_t option, int64_t value);

int bhstem_version(void);
const char* bhstem_last_error(void);
/* Kernel launches issued through this handle so far. */
int64_t bhstem_launch_count(const bhstem_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* BHSTEM_H_ */

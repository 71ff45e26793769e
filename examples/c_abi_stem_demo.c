/* Plain-C use of both C ABIs (include/bhmel.h, include/bhstem.h): the serving chain of the reference's encoder
 * front -- log-mel frontend, conditioning channels, conv stem -- without Python or torch, in its two forms:
 *
 *   full    bhmel_forward_encoder_input  ->  [B][T][80 + 384] bf16  ->  bhstem_forward
 *   split   bhmel_forward_ex (bf16 frames [B][T][80])  +  cond [B][384]  ->  bhstem_forward_split
 *           (the time-constant conditioning channels folded into a per-window bias; the encoder input is never built)
 *
 *   nvcc -x cu examples/c_abi_stem_demo.c -Iinclude -Lbeatheritage_b200 -lbhmel -lbhstem \
 *        -Xlinker -rpath=$PWD/beatheritage_b200 -o /tmp/c_abi_stem_demo && /tmp/c_abi_stem_demo
 *
 * Two windows of 65 408 samples (512 frames: the reference's src_seq_len 512 configuration), whisper-small widths,
 * pseudo-random bf16-representable parameters.  Checks that both forms agree within the stem's bf16 tolerance
 * (|diff| <= 2^-6 |y| + 4e-3: only the fp32 summation order inside conv1 differs) and prints the launch counts.
 * Exit code 0 on success.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bhmel.h"
#include "bhstem.h"

#define CHECK_MEL(call)                                                                                  \
  do {                                                                                                   \
    int rc__ = (call);                                                                                   \
    if (rc__ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, bhmel_last_error()); return 1; } \
  } while (0)
#define CHECK_STEM(call)                                                                                  \
  do {                                                                                                    \
    int rc__ = (call);                                                                                    \
    if (rc__ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, bhstem_last_error()); return 1; } \
  } while (0)
#define CHECK_CUDA(call)                                                                          \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e__)); return 2; } \
  } while (0)

static uint16_t f32_to_bf16(float f) {          /* round to nearest even, like tensor.to(bfloat16) */
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint32_t rng_state = 12345u;
static float rnd(void) {                         /* uniform in [-1, 1) */
  rng_state = rng_state * 1664525u + 1013904223u;
  return (float)(rng_state >> 8) / 8388608.0f - 1.0f;
}

int main(void) {
  const int64_t B = 2, N = 65408;
  const int32_t n_mels = 80, n_cond = 384, c_in = n_mels + n_cond, d_model = 768;
  bhmel_params prm;
  memset(&prm, 0, sizeof(prm));
  prm.sample_rate = 16000; prm.n_fft = 1024; prm.hop_length = 128; prm.n_mels = n_mels;
  prm.f_min = 20.0; prm.f_max = 8000.0; prm.pad_mode = BHMEL_PAD_REFLECT; prm.log_scale = 1;
  bhmel_handle* mel = NULL;
  CHECK_MEL(bhmel_create(&prm, &mel));
  const int64_t T = bhmel_num_frames(mel, N);    /* 512 */
  if (T % 2) { fprintf(stderr, "expected an even frame count, got %lld\n", (long long)T); return 3; }

  /* stem parameters in torch's Conv1d layout, bf16-representable like a model cast with .to(bfloat16) */
  const size_t n_w1 = (size_t)d_model * c_in * 3, n_w2 = (size_t)d_model * d_model * 3;
  float* w1 = (float*)malloc(n_w1 * sizeof(float));
  float* w2 = (float*)malloc(n_w2 * sizeof(float));
  float* b1 = (float*)malloc(d_model * sizeof(float));
  float* b2 = (float*)malloc(d_model * sizeof(float));
  const float s1 = 1.0f / sqrtf(3.0f * c_in), s2 = 1.0f / sqrtf(3.0f * d_model);
  for (size_t i = 0; i < n_w1; ++i) w1[i] = bf16_to_f32(f32_to_bf16(rnd() * s1));
  for (size_t i = 0; i < n_w2; ++i) w2[i] = bf16_to_f32(f32_to_bf16(rnd() * s2));
  for (int i = 0; i < d_model; ++i) { b1[i] = bf16_to_f32(f32_to_bf16(rnd() * s1)); b2[i] = bf16_to_f32(f32_to_bf16(rnd() * s2)); }
  bhstem_handle* stem = NULL;
  CHECK_STEM(bhstem_create(c_in, d_model, w1, b1, w2, b2, &stem));
  CHECK_STEM(bhstem_prepare_split(stem, n_mels));           /* once per handle */

  /* two windows of a two-tone signal, one conditioning vector per window */
  float* x = (float*)malloc((size_t)(B * N) * sizeof(float));
  for (int64_t i = 0; i < B * N; ++i) {
    const float t = (float)(i % N) / 16000.0f;
    x[i] = 0.4f * sinf(6.2831853f * 440.0f * t) + 0.3f * sinf(6.2831853f * (1500.0f + 700.0f * (float)(i / N)) * t) + 0.05f * rnd();
  }
  uint16_t* cond = (uint16_t*)malloc((size_t)(B * n_cond) * sizeof(uint16_t));
  for (int64_t i = 0; i < B * n_cond; ++i) cond[i] = f32_to_bf16(rnd() * 1.5f);

  float* d_x = NULL;
  void *d_cond = NULL, *d_enc = NULL, *d_frames = NULL, *d_hidden = NULL, *d_y_full = NULL, *d_y_split = NULL;
  float* d_bias3 = NULL;
  const size_t n_y = (size_t)(B * (T / 2) * d_model);
  CHECK_CUDA(cudaMalloc((void**)&d_x, (size_t)(B * N) * sizeof(float)));
  CHECK_CUDA(cudaMalloc(&d_cond, (size_t)(B * n_cond) * 2));
  CHECK_CUDA(cudaMalloc(&d_enc, (size_t)(B * T * c_in) * 2));
  CHECK_CUDA(cudaMalloc(&d_frames, (size_t)(B * T * n_mels) * 2));
  CHECK_CUDA(cudaMalloc(&d_hidden, (size_t)(B * T * d_model) * 2));
  CHECK_CUDA(cudaMalloc(&d_y_full, n_y * 2));
  CHECK_CUDA(cudaMalloc(&d_y_split, n_y * 2));
  CHECK_CUDA(cudaMalloc((void**)&d_bias3, (size_t)(B * 3 * d_model) * sizeof(float)));
  CHECK_CUDA(cudaMemcpy(d_x, x, (size_t)(B * N) * sizeof(float), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_cond, cond, (size_t)(B * n_cond) * 2, cudaMemcpyHostToDevice));

  /* full form: assembled channels-last encoder input, two stem launches */
  bhmel_encoder_input_desc enc;
  memset(&enc, 0, sizeof(enc));
  enc.y = d_enc; enc.dtype = BHMEL_OUT_BF16; enc.layout = BHMEL_LAYOUT_BTC; enc.cond = d_cond; enc.n_cond = n_cond;
  CHECK_MEL(bhmel_forward_encoder_input(mel, d_x, B, N, N, &enc, NULL));
  CHECK_STEM(bhstem_forward(stem, d_enc, B, T, d_hidden, d_y_full, NULL));
  /* split form: dense bf16 frames, folded bias + split conv1 + conv2 */
  bhmel_out_desc out;
  memset(&out, 0, sizeof(out));
  out.y = d_frames; out.dtype = BHMEL_OUT_BF16;
  CHECK_MEL(bhmel_forward_ex(mel, d_x, B, N, N, &out, NULL));
  CHECK_STEM(bhstem_forward_split(stem, d_frames, d_cond, B, T, d_bias3, d_hidden, d_y_split, NULL));
  CHECK_CUDA(cudaDeviceSynchronize());

  uint16_t* y_full = (uint16_t*)malloc(n_y * 2);
  uint16_t* y_split = (uint16_t*)malloc(n_y * 2);
  CHECK_CUDA(cudaMemcpy(y_full, d_y_full, n_y * 2, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(y_split, d_y_split, n_y * 2, cudaMemcpyDeviceToHost));
  size_t differ = 0, nonzero = 0;
  double worst = 0.0;
  for (size_t i = 0; i < n_y; ++i) {
    const float a = bf16_to_f32(y_full[i]), b = bf16_to_f32(y_split[i]);
    if (!isfinite(a) || !isfinite(b)) { fprintf(stderr, "non-finite output at %zu\n", i); return 4; }
    const double diff = fabs((double)a - (double)b), bound = fabs((double)a) * 0.015625 + 4e-3;
    if (diff > bound) { fprintf(stderr, "element %zu: full %g split %g\n", i, a, b); return 5; }
    if (diff > worst) worst = diff;
    differ += y_full[i] != y_split[i];
    nonzero += a != 0.0f;
  }
  printf("frames per window: %lld, stem output [%lld][%lld][%d]; split vs full: %zu of %zu elements differ, max |diff| %.4g; "
         "launches: frontend %lld, stem %lld\n",
         (long long)T, (long long)B, (long long)(T / 2), d_model, differ, n_y, worst, (long long)bhmel_launch_count(mel),
         (long long)bhstem_launch_count(stem));
  if (nonzero < n_y / 2 || differ * 10 > n_y) return 6;
  if (bhstem_launch_count(stem) != 5) return 7;               /* 2 (full) + 3 (split) */
  bhstem_destroy(stem);
  bhmel_destroy(mel);
  cudaFree(d_x); cudaFree(d_cond); cudaFree(d_enc); cudaFree(d_frames); cudaFree(d_hidden); cudaFree(d_y_full);
  cudaFree(d_y_split); cudaFree(d_bias3);
  free(w1); free(w2); free(b1); free(b2); free(x); free(cond); free(y_full); free(y_split);
  return 0;
}

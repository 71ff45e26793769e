/* Plain-C use of the libbhmel C ABI (include/bhmel.h): what a non-Python host would bind.
 *
 *   nvcc -x cu examples/c_abi_demo.c -Iinclude -Lbeatheritage_b200 -lbhmel -Xlinker -rpath=$PWD/beatheritage_b200 -o /tmp/c_abi_demo
 *   /tmp/c_abi_demo
 *
 * Runs the P0 frontend on a 440 Hz tone (2 windows of 65 536 samples), once from device buffers and
 * once through the host-buffer entry, checks both agree bit for bit and that the tone's mel band
 * dominates, and prints the frame count.  Exit code 0 on success.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bhmel.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc__ = (call);                                                       \
    if (rc__ != 0) {                                                         \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, bhmel_last_error()); \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(void) {
  const int64_t B = 2, N = 65536;
  bhmel_params prm;
  memset(&prm, 0, sizeof(prm));
  prm.sample_rate = 16000; prm.n_fft = 1024; prm.hop_length = 128; prm.n_mels = 80;
  prm.f_min = 20.0; prm.f_max = 8000.0; prm.pad_mode = BHMEL_PAD_REFLECT; prm.log_scale = 1;
  bhmel_handle* h = NULL;
  CHECK(bhmel_create(&prm, &h));
  const int64_t T = bhmel_num_frames(h, N);
  const size_t n_in = (size_t)(B * N), n_out = (size_t)(B * T * prm.n_mels);
  float* x = (float*)malloc(n_in * sizeof(float));
  float* y_dev_path = (float*)malloc(n_out * sizeof(float));
  float* y_host_path = (float*)malloc(n_out * sizeof(float));
  for (size_t i = 0; i < n_in; ++i) x[i] = 0.5f * sinf(2.0f * 3.14159265358979f * 440.0f * (float)(i % N) / 16000.0f);

  float *dx = NULL, *dy = NULL;
  if (cudaMalloc((void**)&dx, n_in * sizeof(float)) != cudaSuccess || cudaMalloc((void**)&dy, n_out * sizeof(float)) != cudaSuccess) return 2;
  cudaMemcpy(dx, x, n_in * sizeof(float), cudaMemcpyHostToDevice);
  CHECK(bhmel_forward(h, dx, B, N, N, dy, NULL));                 /* default stream */
  if (cudaDeviceSynchronize() != cudaSuccess) return 3;
  cudaMemcpy(y_dev_path, dy, n_out * sizeof(float), cudaMemcpyDeviceToHost);
  CHECK(bhmel_forward_host(h, x, B, N, N, y_host_path));
  if (memcmp(y_dev_path, y_host_path, n_out * sizeof(float)) != 0) { fprintf(stderr, "device and host paths differ\n"); return 4; }

  /* the loudest mel band of a middle frame must be the one containing 440 Hz (band 13 or 14 for P0: mel(440 Hz) = 549.7, band centres every 34.67 mel from 31.75) */
  const float* fr = y_dev_path + (size_t)(T / 2) * prm.n_mels;
  int best = 0;
  for (int m = 1; m < prm.n_mels; ++m) if (fr[m] > fr[best]) best = m;
  printf("frames per window: %lld, loudest mel band of the 440 Hz tone: %d (log1p power %.3f), launches: %lld\n",
         (long long)T, best, fr[best], (long long)bhmel_launch_count(h));
  if (best < 12 || best > 15) return 5;
  bhmel_destroy(h);
  cudaFree(dx); cudaFree(dy); free(x); free(y_dev_path); free(y_host_path);
  return 0;
}

// lsu_probe.cu -- shared-memory / shuffle throughput of the patterns the FFT role uses, on one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lsu_probe lsu_probe.cu && ./lsu_probe
// For W = 1, 2, 4, 8, 16 concurrently running warps (one CTA, one SM) it reports cycles per
// repetition of:  T32  one 32x32 plane transpose (32 STS.32, syncwarp, 32 LDS.32, syncwarp)
//                 T64  two planes at once as float2 (32 STS.64 + 32 LDS.64)
//                 T128 four planes (two pairs' worth) as float4 (32 STS.128 + 32 LDS.128)
//                 SHF  32 shuffles from lane (32-lane)&31
//                 LD4  32 LDS.128 at the mel role's row pitch (532 floats)
// and the implied shared wavefronts per cycle SM-wide.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kReps = 200;

template <int MODE>
__global__ void probe(float* out, long long* cyc) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float r[32], i2[32], a3[32], a4[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) { r[k] = lane + k; i2[k] = lane - k; a3[k] = lane * k; a4[k] = k; }
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {
    float* rows = sm + warp * (32 * 33);
    for (int rep = 0; rep < kReps; ++rep) {
#pragma unroll
      for (int k = 0; k < 32; ++k) rows[k * 33 + lane] = r[k];
      __syncwarp();
#pragma unroll
      for (int n = 0; n < 32; ++n) r[n] = rows[lane * 33 + n] + 1.f;
      __syncwarp();
    }
  } else if (MODE == 1) {
    float2* rows = reinterpret_cast<float2*>(sm) + warp * (32 * 33);
    for (int rep = 0; rep < kReps; ++rep) {
#pragma unroll
      for (int k = 0; k < 32; ++k) rows[k * 33 + lane] = make_float2(r[k], i2[k]);
      __syncwarp();
#pragma unroll
      for (int n = 0; n < 32; ++n) { float2 v = rows[lane * 33 + n]; r[n] = v.x + 1.f; i2[n] = v.y; }
      __syncwarp();
    }
  } else if (MODE == 2) {
    float4* rows = reinterpret_cast<float4*>(sm) + warp * (32 * 33);
    for (int rep = 0; rep < kReps; ++rep) {
#pragma unroll
      for (int k = 0; k < 32; ++k) rows[k * 33 + lane] = make_float4(r[k], i2[k], a3[k], a4[k]);
      __syncwarp();
#pragma unroll
      for (int n = 0; n < 32; ++n) { float4 v = rows[lane * 33 + n]; r[n] = v.x + 1.f; i2[n] = v.y; a3[n] = v.z; a4[n] = v.w; }
      __syncwarp();
    }
  } else if (MODE == 3) {
    const int src = (32 - lane) & 31;
    for (int rep = 0; rep < kReps; ++rep) {
#pragma unroll
      for (int k = 0; k < 32; ++k) r[k] = __shfl_sync(0xffffffffu, r[k], src) + 1.f;
    }
  } else if (MODE == 7) {   // STS.32 only, conflict free
    float* rows = sm + warp * (32 * 33);
    for (int rep = 0; rep < kReps; ++rep) {
#pragma unroll
      for (int k = 0; k < 32; ++k) rows[k * 33 + lane + (rep & 1) * 2048] = r[k];
      __syncwarp();
    }
  } else if (MODE == 8) {   // LDS.32 only, conflict free (pitch 33)
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(sm)) + (warp * 32 * 33 + lane * 33) * 4;
    for (int rep = 0; rep < kReps; ++rep) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + k * 4 + (rep & 1) * 8192) : "memory");
        r[k] += v;
      }
    }
  } else if (MODE == 4 || MODE == 5 || MODE == 6) {
    // MODE 4: LDS.128, lane = row at the mel role's pitch (532 floats); MODE 5: LDS.128, lanes contiguous;
    // MODE 6: LDS.128 broadcast (all lanes one address, the mel weights)
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(sm)) +
                          (MODE == 4 ? lane * 532 * 4 : MODE == 5 ? lane * 16 : 0) + warp * 16;
    for (int rep = 0; rep < kReps; ++rep) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(base + (MODE == 5 ? k * 512 : k * 16 * 4) + (rep & 1) * 128)
                     : "memory");
        r[k] += v.x + v.y + v.z + v.w;
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) s += r[k] + i2[k] + a3[k] + a4[k];
  out[threadIdx.x] = s;
  if (lane == 0) cyc[warp] = t1 - t0;
}

template <int MODE>
void run(const char* name, double wf_per_rep) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 * 1024); cudaMalloc(&cyc, 8 * 32);
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int W : {1, 2, 4, 8, 16}) {
    if (MODE == 2 && W > 8) continue;   // 16 warps x 16.9 KB would not fit
    long long h[32];
    for (int pass = 0; pass < 2; ++pass) {
      probe<MODE><<<1, W * 32, 200 * 1024>>>(out, cyc);
      cudaMemcpy(h, cyc, 8 * W, cudaMemcpyDeviceToHost);
    }
    long long mx = 0;
    for (int w = 0; w < W; ++w) mx = h[w] > mx ? h[w] : mx;
    const double per = double(mx) / kReps;
    printf("%-5s W=%2d  %8.1f cycles/rep/warp   %6.3f wavefronts/cycle SM-wide\n", name, W, per, wf_per_rep * W / per);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("%s: %s\n", name, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("T32", 64);
  run<1>("T64", 128);
  run<2>("T128", 256);
  run<3>("SHF", 32);
  run<4>("LD4p", 128);
  run<5>("LD4c", 128);
  run<6>("LD4b", 32);
  run<7>("ST1", 32);
  run<8>("LD1", 32);
  return 0;
}

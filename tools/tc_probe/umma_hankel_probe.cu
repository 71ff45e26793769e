// Experimental probe (NOT part of the product): can a tcgen05 TF32 MMA read an A operand whose rows
// OVERLAP in shared memory?  A[m][k] = S[4m + k] -- the Hankel/Toeplitz structure a framed signal has
// (frame m starts 4 elements after frame m-1).  With a K-major, no-swizzle descriptor the canonical
// core matrix is 8 rows x 16 bytes with a 16-byte row pitch, which is exactly a frame pitch of
// 4 TF32 elements: LBO (next core matrix along K) = 16 B, SBO (next 8 rows) = 128 B.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe tools/tc_probe/umma_hankel_probe.cu
//
// Prints max |D - expected| for D[128][16] = A[128][K] * B[K][16], K = 8 and K = 32 (4 accumulating
// MMAs), and the cycles per MMA for a burst of small MMAs.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, sm_100 version field = 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;   // descriptor version
  return d;
}

constexpr int M = 128, N = 16, KMAX = 32;
constexpr int S_LEN = 4 * (M - 1) + KMAX + 8;

__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ S_g, const float* __restrict__ Bt_g /*[N][KMAX]*/,
                                                float* __restrict__ out8, float* __restrict__ out32,
                                                long long* __restrict__ cycles, int burst) {
  __shared__ __align__(128) float S[S_LEN];
  __shared__ __align__(128) float Bs[KMAX / 4 * 2 * 32];   // core matrices [(n/8)][(k/4)][8 n][4 k]
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < S_LEN; i += 128) S[i] = S_g[i];
  for (int i = tid; i < N * KMAX; i += 128) {
    const int n = i / KMAX, k = i % KMAX;
    Bs[((n / 8) * (KMAX / 4) + (k / 4)) * 32 + (n % 8) * 4 + (k % 4)] = Bt_g[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> async proxy (MMA) reads
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem_base;

  // instruction descriptor: D=F32, A=B=TF32, both K-major, N>>3, M>>4
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
                         (static_cast<uint32_t>(M >> 4) << 24);
  const uint32_t a0 = smem_u32(S), b0 = smem_u32(Bs);
  uint32_t parity = 0;

  for (int test = 0; test < 2; ++test) {
    const int ksteps = test == 0 ? 1 : KMAX / 8;
    if (tid == 0) {
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t adesc = make_desc(a0 + ks * 32, 16, 128);                 // rows overlap: pitch 16 B
        const uint64_t bdesc = make_desc(b0 + ks * 2 * 128, 128, (KMAX / 4) * 128);
        const uint32_t acc = ks > 0;
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(taddr),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs
    {
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(&bar)), "r"(parity)
                     : "memory");
      }
      parity ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t r[16];
    const uint32_t lane_addr = taddr + (static_cast<uint32_t>(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(lane_addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float* out = test == 0 ? out8 : out32;
    for (int j = 0; j < 16; ++j) out[tid * N + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }

  // throughput of a burst of small MMAs (N = 16, K = 8), one commit at the end
  if (tid == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < burst; ++i) {
      const uint64_t adesc = make_desc(a0 + (i & 3) * 32, 16, 128);
      const uint64_t bdesc = make_desc(b0 + (i & 3) * 2 * 128, 128, (KMAX / 4) * 128);
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(taddr),
          "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok)
                   : "r"(smem_u32(&bar)), "r"(parity)
                   : "memory");
    }
    const long long t2 = clock64();
    cycles[0] = t1 - t0;   // issue
    cycles[1] = t2 - t0;   // issue + drain
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(taddr));
}

int main() {
  std::vector<float> S(S_LEN), Bt(N * KMAX);
  // values exactly representable in TF32 (10-bit mantissa): small multiples of 1/64
  for (int i = 0; i < S_LEN; ++i) S[i] = static_cast<float>((i * 37 + 11) % 129 - 64) / 64.0f;
  for (int i = 0; i < N * KMAX; ++i) Bt[i] = static_cast<float>((i * 53 + 5) % 65 - 32) / 32.0f;
  float *dS, *dB, *d8, *d32;
  long long* dc;
  cudaMalloc(&dS, S_LEN * 4); cudaMalloc(&dB, N * KMAX * 4); cudaMalloc(&d8, M * N * 4); cudaMalloc(&d32, M * N * 4);
  cudaMalloc(&dc, 16);
  cudaMemcpy(dS, S.data(), S_LEN * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bt.data(), N * KMAX * 4, cudaMemcpyHostToDevice);
  const int burst = 512;
  probe<<<1, 128>>>(dS, dB, d8, d32, dc, burst);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> o8(M * N), o32(M * N);
  long long cyc[2];
  cudaMemcpy(o8.data(), d8, M * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(o32.data(), d32, M * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost);
  for (int K : {8, 32}) {
    const std::vector<float>& o = K == 8 ? o8 : o32;
    double worst = 0;
    int bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += static_cast<double>(S[4 * m + k]) * Bt[n * KMAX + k];
        const double err = std::abs(ref - o[m * N + n]);
        if (err > worst) worst = err;
        if (err > 1e-4) ++bad;
      }
    printf("K=%2d: max |D - expected| = %.3e, mismatches %d / %d   (D[5][3] = %f)\n", K, worst, bad, M * N, o[5 * N + 3]);
  }
  printf("burst of %d MMAs (M=128, N=16, K=8, TF32): issue %lld cycles (%.1f / MMA), issue+drain %lld cycles (%.1f / MMA)\n",
         burst, cyc[0], static_cast<double>(cyc[0]) / burst, cyc[1], static_cast<double>(cyc[1]) / burst);
  return 0;
}

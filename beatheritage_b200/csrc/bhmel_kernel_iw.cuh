// bhmel_kernel_iw.cuh -- "independent warps" schedule of the fused log-mel kernel (sm_100a).
//
// Same arithmetic as bhmel_logmel_kernel (bhmel_kernel.cuh) -- identical FFT passes, pair
// separation, 4-bin-group mel accumulation and log epilogue, so the results are bit-identical --
// but no CTA-wide barrier anywhere: every warp of the persistent CTA is its own pipeline.
//
//   warp tile   8 consecutive frames (4 frame pairs) of one input row.
//   stage 1     the warp's 1920-sample span is staged in its PRIVATE double-buffered shared
//               memory by one TMA bulk copy (cp.async.bulk + mbarrier) issued one tile ahead;
//               edge / unaligned tiles use per-element cp.async with the reflect / zero mapping.
//   stage 2     per pair: 1024-pt complex FFT as 32 x 32 (in-register passes, warp transpose
//               through private padded shared memory, window and twiddles in registers), pair
//               separation by shuffle, |X|^2 of both frames written to two private P rows
//               (aliasing the transpose scratch).
//   stage 3     the same warp projects its two P rows on the mel filterbank: lanes 0-15 frame A,
//               lanes 16-31 frame B, lane i owns filter 16 r + i in round r (tables from
//               make_rounds: warp-uniform trip count, conflict-free 128-bit loads), log1p, and
//               stores [frame][16 filters] runs straight to global memory (64-byte segments).
//
// Because warps never wait for each other, the latency-bound stage 3 of one warp overlaps the
// issue-bound stage 2 of the others.
#pragma once
#include "bhmel_kernel.cuh"

namespace bhmel {
namespace iw {

constexpr int kWTileF = 8;                                  // frames per warp tile
constexpr int kWPairs = kWTileF / 2;
constexpr int kWSpan = (kWTileF - 1) * kHop + kNfft;        // 1920 samples
constexpr int kWSpanBytes = kWSpan * 4;                     // 7680 (multiple of 16)
constexpr int kIwWarps = 8;
constexpr int kIwThreads = kIwWarps * 32;
constexpr int kWtCap4 = 1536;                               // float4 of round weights kept in shared memory
constexpr int kRoundCap = 64;                               // n_mels <= 1024

struct IwParams {
  KParams k;                 // x, strides, N, T, y, win_half, tw, n_mels, pad/log flags, n_tiles, tiles_per_row
  const RoundDesc* rounds;   // [n_rounds]
  const int* g0;             // [n_rounds * 16]
  const float4* wt;          // round weights
  int n_rounds;
  int n_wt4;
};

struct SmemIW {
  float2 scr[kIwWarps][32 * kScrPitch];      // 67 584 B  transpose scratch; first 2*516 floats double as P rows
  float span[kIwWarps][2][kWSpan];           // 122 880 B per-warp double-buffered spans
  float4 wt[kWtCap4];                        // 24 576 B
  int g0[kRoundCap * 16];                    //  4 096 B
  int2 rounds[kRoundCap];                    //    512 B
  unsigned long long bar_bulk[kIwWarps][2];  // count 1  (TMA path)
  unsigned long long bar_gen[kIwWarps][2];   // count 32 (cp.async path)
};

__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// Issue the asynchronous fill of one private span buffer.  Returns true for the TMA path.
__device__ __forceinline__ bool issue_warp_span(const KParams& p, long long tile, float* dst,
                                                unsigned long long* bar_bulk, unsigned long long* bar_gen, int lane) {
  const long long r = tile / p.tiles_per_row;
  const int tb = static_cast<int>(tile - r * p.tiles_per_row);
  const long long s0 = static_cast<long long>(tb) * (kWTileF * kHop) - kNfft / 2;
  const long long row_off = p.row0 + r * p.row_stride;
  long long valid = p.n_total - row_off;
  valid = valid < 0 ? 0 : (valid > p.N ? p.N : valid);
  const float* row = p.x + row_off;
  const bool interior = s0 >= 0 && s0 + kWSpan <= valid;
  if (p.use_bulk && interior && ((reinterpret_cast<uintptr_t>(row + s0) & 15) == 0)) {
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(bar_bulk, kWSpanBytes);
      bulk_g2s(dst, row + s0, kWSpanBytes, bar_bulk);
    }
    return true;
  }
  if (interior) {
    const float* src = row + s0;
#pragma unroll 4
    for (int e = lane; e < kWSpan; e += 32) cp_async_4(dst + e, src + e, 4);
  } else {
    const long long N = p.N;
    for (int e = lane; e < kWSpan; e += 32) {
      long long i = s0 + e;
      if (i < 0) i = p.pad_reflect ? -i : -1;
      else if (i >= N) i = p.pad_reflect ? 2 * (N - 1) - i : -1;
      const bool ok = (i >= 0) && (i < valid);
      cp_async_4(dst + e, row + (ok ? i : 0), ok ? 4 : 0);
    }
  }
  cp_async_arrive_noinc(bar_gen);
  return false;
}

template <bool kLog, bool kSmemW>
__device__ __forceinline__ void mel_rounds(const float* __restrict__ prows, const SmemIW& S, const IwParams& q,
                                           long long yrow, bool frame_ok, int lane) {
  const int li = lane & 15;
  const float4* prow = reinterpret_cast<const float4*>(prows + (lane >> 4) * kPPitch);
  const float4* wt = kSmemW ? S.wt : q.wt;
#pragma unroll 1
  for (int r = 0; r < q.n_rounds; ++r) {
    const int2 rd = S.rounds[r];
    const float4* pp = prow + S.g0[r * 16 + li];
    const float4* wp = wt + rd.x + li;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 2
    for (int g = 0; g < rd.y; ++g) {
      const float4 w4 = ldw<kSmemW>(wp + g * 16);
      const float4 x4 = pp[g];
      a0 = fmaf(x4.x, w4.x, a0);
      a1 = fmaf(x4.y, w4.y, a1);
      a2 = fmaf(x4.z, w4.z, a2);
      a3 = fmaf(x4.w, w4.w, a3);
    }
    float v = (a0 + a1) + (a2 + a3);
    if constexpr (kLog) v = fast_log1p(v);
    const int m = r * 16 + li;
    if (frame_ok && m < q.k.n_mels) store_out(q.k, yrow + m, v);
  }
}

template <bool kLog>
__global__ void __launch_bounds__(kIwThreads, 1) bhmel_logmel_iw_kernel(const __grid_constant__ IwParams q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemIW& S = *reinterpret_cast<SmemIW*>(smem_raw);
  const KParams& p = q.k;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  // one-time CTA setup (the only CTA-wide barrier of the kernel)
  const bool wt_in_smem = q.n_wt4 <= kWtCap4;
  if (wt_in_smem)
    for (int i = tid; i < q.n_wt4; i += kIwThreads) S.wt[i] = q.wt[i];
  for (int i = tid; i < q.n_rounds * 16; i += kIwThreads) S.g0[i] = q.g0[i];
  for (int i = tid; i < q.n_rounds; i += kIwThreads) S.rounds[i] = make_int2(q.rounds[i].woff4, q.rounds[i].ng);
  if (lane == 0) {
    mbar_init(&S.bar_bulk[warp][0], 1);
    mbar_init(&S.bar_bulk[warp][1], 1);
    mbar_init(&S.bar_gen[warp][0], 32);
    mbar_init(&S.bar_gen[warp][1], 32);
    fence_mbar_init();
  }
  __syncthreads();

  float wreg[32], twr[32], twi[32];
#pragma unroll
  for (int m = 0; m < 32; ++m) {
    wreg[m] = __ldg(p.win_half + lane + 32 * m);
    const float2 t = __ldg(p.tw + m * 32 + lane);
    twr[m] = t.x;
    twi[m] = t.y;
  }
  float2* scr = S.scr[warp];
  float* prows = reinterpret_cast<float*>(scr);          // two P rows alias the transpose scratch
  const int src = (32 - lane) & 31;
  const long long wstride = static_cast<long long>(gridDim.x) * kIwWarps;
  long long tile = static_cast<long long>(blockIdx.x) * kIwWarps + warp;
  uint32_t par_bulk = 0, par_gen = 0;                    // bit b = next parity to wait for on buffer b
  bool cur_bulk = false;
  if (tile < p.n_tiles)
    cur_bulk = issue_warp_span(p, tile, S.span[warp][0], &S.bar_bulk[warp][0], &S.bar_gen[warp][0], lane);

#pragma unroll 1
  for (int it = 0; tile < p.n_tiles; tile += wstride, ++it) {
    const int b = it & 1;
    // prefetch the next tile into the other buffer (every lane finished reading it last iteration)
    __syncwarp();
    const long long next_tile = tile + wstride;
    bool next_bulk = false;
    if (next_tile < p.n_tiles)
      next_bulk = issue_warp_span(p, next_tile, S.span[warp][b ^ 1], &S.bar_bulk[warp][b ^ 1],
                                  &S.bar_gen[warp][b ^ 1], lane);
    // wait for this tile's span
    if (cur_bulk) {
      mbar_wait(&S.bar_bulk[warp][b], (par_bulk >> b) & 1);
      par_bulk ^= 1u << b;
    } else {
      mbar_wait(&S.bar_gen[warp][b], (par_gen >> b) & 1);
      par_gen ^= 1u << b;
    }
    const float* span = S.span[warp][b];
    const long long r = tile / p.tiles_per_row;
    const int t0 = static_cast<int>(tile - r * p.tiles_per_row) * kWTileF;
    const long long ytile = r * p.y_row_pitch + static_cast<long long>(t0) * p.y_frame_pitch;

#pragma unroll 1
    for (int j = 0; j < kWPairs; ++j) {
      float ar[32], ai[32];
      {
        float v[36];
        const float* sp = span + (2 * j) * kHop + lane;
#pragma unroll
        for (int m = 0; m < 36; ++m) v[m] = sp[32 * m];
        fft32_pass_a(v, wreg, ar, ai);
      }
      __syncwarp();   // the previous pair's mel stage has finished reading the P rows (aliased below)
#pragma unroll
      for (int k = 0; k < 32; ++k) scr[k * kScrPitch + lane] = make_float2(ar[k], ai[k]);
      __syncwarp();
      float br[32], bi[32];
      {
        float ur[32], ui[32];
#pragma unroll
        for (int n = 0; n < 32; ++n) {
          const float2 u = scr[lane * kScrPitch + n];
          ur[n] = u.x;
          ui[n] = u.y;
        }
        fft32_pass_b(ur, ui, twr, twi, br, bi);
      }
      __syncwarp();   // every lane has read its transposed column before the P rows overwrite it
      float* Pa = prows + lane;
      float* Pb = Pa + kPPitch;
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const int s = 31 - k2;
        float pr = __shfl_sync(0xffffffffu, br[s], src);
        float pi = __shfl_sync(0xffffffffu, bi[s], src);
        if (lane == 0) {
          pr = br[(s + 1) & 31];
          pi = bi[(s + 1) & 31];
        }
        const float a1 = br[k2] + pr, a2 = bi[k2] - pi;
        const float b1 = bi[k2] + pi, b2 = pr - br[k2];
        Pa[32 * k2] = fmaf(a1, a1, a2 * a2);
        Pb[32 * k2] = fmaf(b1, b1, b2 * b2);
      }
      if (lane == 0) {
        const float zr = 2.f * br[16], zi = 2.f * bi[16];
        Pa[512] = zr * zr;
        Pb[512] = zi * zi;
      }
      if (lane < 2 * (kPPitch - kBins)) {   // zero the pad bins 513..515 of both rows
        const int row = lane / (kPPitch - kBins), c = lane % (kPPitch - kBins);
        prows[row * kPPitch + kBins + c] = 0.f;
      }
      __syncwarp();
      const int tf = t0 + 2 * j + (lane >> 4);            // this lane's frame
      const long long yrow = ytile + static_cast<long long>(2 * j + (lane >> 4)) * p.y_frame_pitch;
      if (wt_in_smem) mel_rounds<kLog, true>(prows, S, q, yrow, tf < p.T, lane);
      else mel_rounds<kLog, false>(prows, S, q, yrow, tf < p.T, lane);
    }
    cur_bulk = next_bulk;
  }
}

}  // namespace iw
}  // namespace bhmel

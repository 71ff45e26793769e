// bhmel_kernel_ws.cuh -- warp-specialised variant of the fused log-mel kernel (sm_100a).
//
// Same arithmetic as bhmel_logmel_kernel (bhmel_kernel.cuh: identical FFT passes, pairing,
// banded mel, log epilogue -- results are bit-identical), different schedule: the three stages
// run concurrently on different warps of one persistent CTA per SM and hand tiles to each other
// through mbarriers instead of CTA-wide barriers, so the latency-bound mel/store stage fills the
// issue slots the FFT warps leave idle.
//
//   warps 0-7   (2 warpgroups, setmaxnreg -> 200)  FFT role: one frame PAIR per warp per tile
//   warps 8-11  (1 warpgroup,  setmaxnreg -> 104)  producer + mel + store role
//
//   tile = 16 consecutive frames of one input row.
//   span[2]   2944-sample spans, filled by the producer role (TMA bulk copy, or per-element
//             cp.async with reflect / zero mapping) two tiles ahead      -> span_full[2]
//   P[2]      power spectra [16][516], written by the FFT warps          -> p_full[2]
//             and released by the mel warps                              -> p_empty[2]
//   mel role  lane = (frame 0..15, pair member 0..1): each lane owns one (frame, filter) dot
//             product per pair descriptor; staging + coalesced store inside the warpgroup
//             (named barrier 1).
#pragma once
#include "bhmel_kernel.cuh"

namespace bhmel {
namespace ws {

constexpr int kTile = 16;
constexpr int kSpanW = (kTile - 1) * kHop + kNfft;   // 2944 samples
constexpr int kSpanWBytes = kSpanW * 4;              // 11776 (multiple of 16)
constexpr int kFftWarps = 8;
constexpr int kMelWarps = 4;
constexpr int kMelThreads = kMelWarps * 32;
constexpr int kThreadsW = (kFftWarps + kMelWarps) * 32;   // 384
// setmaxnreg budget: the CTA is launched with 168 registers/thread (65536 / 384, rounded down to a
// multiple of 8); registers only move between warpgroups of the CTA, so
// 2 * kFftRegs + kMelRegs must not exceed 3 * 168 = 504 or the FFT warps wait forever.
constexpr int kLaunchRegs = 168;
constexpr int kFftRegs = 200;
constexpr int kMelRegs = 104;
static_assert(2 * kFftRegs + kMelRegs <= 3 * kLaunchRegs, "setmaxnreg pool would deadlock");

struct SmemWS {
  float P[2][kTile * kPPitch];                // 66 048 B
  float2 scr[kFftWarps][32 * kScrPitch];      // 67 584 B
  float span[2][kSpanW];                      // 23 552 B
  float out[kTile * kOutPitch];               //  6 208 B
  float fw[kFwCap];                           // 16 384 B
  int4 pairs[kPairCap];                       //  8 192 B
  unsigned long long span_full[2];
  unsigned long long p_full[2];
  unsigned long long p_empty[2];
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive on `bar` once all cp.async issued so far by this thread have landed (counts as one of
// the barrier's expected arrivals).
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mel_group_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(kMelThreads) : "memory"); }

// Producer: every thread of the mel warpgroup takes part, so span_full always sees 128 arrivals.
__device__ __forceinline__ void issue_span(const KParams& p, long long tile, float* dst, unsigned long long* bar,
                                           int mt) {
  const long long r = tile / p.tiles_per_row;
  const int tb = static_cast<int>(tile - r * p.tiles_per_row);
  const long long s0 = static_cast<long long>(tb) * (kTile * kHop) - kNfft / 2;
  const long long row_off = p.row0 + r * p.row_stride;
  long long valid = p.n_total - row_off;
  valid = valid < 0 ? 0 : (valid > p.N ? p.N : valid);
  const float* row = p.x + row_off;
  const bool interior = s0 >= 0 && s0 + kSpanW <= valid;
  if (p.use_bulk && interior && ((reinterpret_cast<uintptr_t>(row + s0) & 15) == 0)) {
    if (mt == 0) {
      fence_proxy_async();
      mbar_expect_tx(bar, kSpanWBytes);
      bulk_g2s(dst, row + s0, kSpanWBytes, bar);
    } else {
      mbar_arrive(bar);
    }
    return;
  }
  if (interior) {   // unaligned but fully inside the row: plain element copies
    const float* src = row + s0;
    for (int e = mt; e < kSpanW; e += kMelThreads) cp_async_4(dst + e, src + e, 4);
  } else {
    const long long N = p.N;
    for (int e = mt; e < kSpanW; e += kMelThreads) {
      long long i = s0 + e;
      if (i < 0) i = p.pad_reflect ? -i : -1;
      else if (i >= N) i = p.pad_reflect ? 2 * (N - 1) - i : -1;
      const bool ok = (i >= 0) && (i < valid);
      cp_async_4(dst + e, row + (ok ? i : 0), ok ? 4 : 0);
    }
  }
  cp_async_arrive_noinc(bar);
}

template <int NG, bool kSmemW>
__device__ __forceinline__ float band_dot1(const float4* __restrict__ pp, const float4* __restrict__ wp) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const float4 w4 = ldw<kSmemW>(wp + 2 * g);   // pair weights are interleaved A0 B0 A1 B1 ...
    const float4 x4 = pp[g];
    a0 = fmaf(x4.x, w4.x, a0); a1 = fmaf(x4.y, w4.y, a1); a2 = fmaf(x4.z, w4.z, a2); a3 = fmaf(x4.w, w4.w, a3);
  }
  return (a0 + a1) + (a2 + a3);
}
template <bool kSmemW>
__device__ __forceinline__ float band_dot1_n(const float4* __restrict__ pp, const float4* __restrict__ wp, int ng) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 2
  for (int g = 0; g < ng; ++g) {
    const float4 w4 = ldw<kSmemW>(wp + 2 * g);
    const float4 x4 = pp[g];
    a0 = fmaf(x4.x, w4.x, a0); a1 = fmaf(x4.y, w4.y, a1); a2 = fmaf(x4.z, w4.z, a2); a3 = fmaf(x4.w, w4.w, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

template <bool kSmemW, bool kLog>
__device__ __forceinline__ void mel_chunk_ws(const float4* __restrict__ prow, const int4* __restrict__ pd, int npairs,
                                             const float* __restrict__ wbase, float* __restrict__ orow, int mw,
                                             int member) {
  int4 dnext = pd[mw < npairs ? mw : 0];
#pragma unroll 1
  for (int q = mw; q < npairs; q += kMelWarps) {
    const int4 d = dnext;
    if (q + kMelWarps < npairs) dnext = pd[q + kMelWarps];
    const unsigned g0 = member ? (static_cast<unsigned>(d.x) >> 16) : (static_cast<unsigned>(d.x) & 0xFFFFu);
    const unsigned col = member ? (static_cast<unsigned>(d.w) >> 16) : (static_cast<unsigned>(d.w) & 0xFFFFu);
    const float4* wp = reinterpret_cast<const float4*>(wbase + d.y) + member;
    const float4* pp = prow + g0;
    float v;
    if constexpr (!kSmemW) {
      v = band_dot1_n<kSmemW>(pp, wp, d.z);
    } else
    switch (d.z) {
      case 0: v = 0.f; break;
      case 1: v = band_dot1<1, kSmemW>(pp, wp); break;
      case 2: v = band_dot1<2, kSmemW>(pp, wp); break;
      case 3: v = band_dot1<3, kSmemW>(pp, wp); break;
      case 4: v = band_dot1<4, kSmemW>(pp, wp); break;
      case 5: v = band_dot1<5, kSmemW>(pp, wp); break;
      case 6: v = band_dot1<6, kSmemW>(pp, wp); break;
      case 7: v = band_dot1<7, kSmemW>(pp, wp); break;
      case 8: v = band_dot1<8, kSmemW>(pp, wp); break;
      case 9: v = band_dot1<9, kSmemW>(pp, wp); break;
      case 10: v = band_dot1<10, kSmemW>(pp, wp); break;
      default: v = band_dot1_n<kSmemW>(pp, wp, d.z); break;
    }
    if constexpr (kLog) v = __logf(1.0f + v);
    if (col != 0xFFFFu) orow[col] = v;
  }
}

template <bool kLog>
__global__ void __launch_bounds__(kThreadsW, 1) bhmel_logmel_ws_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemWS& S = *reinterpret_cast<SmemWS*>(smem_raw);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  const bool fw_in_smem = p.n_weights <= kFwCap;
  if (fw_in_smem)
    for (int i = tid; i < p.n_weights; i += kThreadsW) S.fw[i] = p.weights[i];
  for (int i = tid; i < p.n_pairs; i += kThreadsW) S.pairs[i] = reinterpret_cast<const int4*>(p.pairs)[i];
  for (int i = tid; i < 2 * kTile * (kPPitch - kBins); i += kThreadsW) {
    const int rowi = i / (kPPitch - kBins);   // 0 .. 2*kTile-1 over both buffers
    (&S.P[0][0])[rowi * kPPitch + kBins + i % (kPPitch - kBins)] = 0.f;
  }
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&S.span_full[b], kMelThreads);
      mbar_init(&S.p_full[b], kFftWarps);
      mbar_init(&S.p_empty[b], kMelWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp < kFftWarps) {
    // =============================== FFT role ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kFftRegs));
    float wreg[32], twr[32], twi[32];
#pragma unroll
    for (int m = 0; m < 32; ++m) {
      wreg[m] = __ldg(p.win_half + lane + 32 * m);
      const float2 t = __ldg(p.tw + m * 32 + lane);
      twr[m] = t.x;
      twi[m] = t.y;
    }
    float2* scr = S.scr[warp];
    const int src = (32 - lane) & 31;
    int it = 0;
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&S.span_full[b], ph);
      float ar[32], ai[32];
      {
        float v[36];
        const float* sp = S.span[b] + (2 * warp) * kHop + lane;
#pragma unroll
        for (int m = 0; m < 36; ++m) v[m] = sp[32 * m];
        fft32_pass_a(v, wreg, ar, ai);
      }
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
        __syncwarp();
        fft32_pass_b(ur, ui, twr, twi, br, bi);
      }
      mbar_wait(&S.p_empty[b], ph ^ 1);   // the mel role has released this P buffer (tile it-2)
      float* Pa = S.P[b] + (2 * warp) * kPPitch + lane;
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
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.p_full[b]);
    }
  } else {
    // ======================= producer + mel + store role =======================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kMelRegs));
    const int mt = tid - kFftWarps * 32;
    const int mw = warp - kFftWarps;
    const int f = lane & 15;
    const int member = lane >> 4;
    {
      long long t0 = blockIdx.x;
      if (t0 < p.n_tiles) issue_span(p, t0, S.span[0], &S.span_full[0], mt);
      t0 += gridDim.x;
      if (t0 < p.n_tiles) issue_span(p, t0, S.span[1], &S.span_full[1], mt);
    }
    int it = 0;
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&S.p_full[b], ph);
      // every FFT warp is done with tile `it`, so span[b] is free: fetch the tile two steps ahead
      const long long tile2 = tile + 2 * static_cast<long long>(gridDim.x);
      if (tile2 < p.n_tiles) issue_span(p, tile2, S.span[b], &S.span_full[b], mt);

      const long long r = tile / p.tiles_per_row;
      const int t0 = static_cast<int>(tile - r * p.tiles_per_row) * kTile;
      const long long frames_left = p.T - t0;
      const int nf = frames_left < kTile ? static_cast<int>(frames_left) : kTile;
      float* ybase = p.y + (r * p.T + t0) * static_cast<long long>(p.n_mels);
      const float4* prow = reinterpret_cast<const float4*>(S.P[b] + f * kPPitch);
      float* orow = S.out + f * kOutPitch;
      for (int mc = 0, c = 0; mc < p.n_mels; mc += kMChunk, ++c) {
        const int mcount = (p.n_mels - mc) < kMChunk ? (p.n_mels - mc) : kMChunk;
        const int4* pd = S.pairs + c * (kMChunk / 2);
        if (fw_in_smem) mel_chunk_ws<true, kLog>(prow, pd, (mcount + 1) >> 1, S.fw, orow, mw, member);
        else mel_chunk_ws<false, kLog>(prow, pd, (mcount + 1) >> 1, p.weights, orow, mw, member);
        if (mc + kMChunk >= p.n_mels) {   // last chunk: this warp no longer reads P[b]
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.p_empty[b]);
        }
        mel_group_sync();   // staging complete
        {
          constexpr int kFr = kTile / kMelWarps, kCo = kMChunk / 32;   // 4 frames x 3 column steps
          float vals[kFr][kCo];
#pragma unroll
          for (int a = 0; a < kFr; ++a)
#pragma unroll
            for (int bb = 0; bb < kCo; ++bb) vals[a][bb] = S.out[(mw + a * kMelWarps) * kOutPitch + lane + 32 * bb];
#pragma unroll
          for (int a = 0; a < kFr; ++a) {
            const int fr = mw + a * kMelWarps;
            float* yrow = ybase + static_cast<long long>(fr) * p.n_mels + mc;
#pragma unroll
            for (int bb = 0; bb < kCo; ++bb) {
              const int c2 = lane + 32 * bb;
              if (fr < nf && c2 < mcount) yrow[c2] = vals[a][bb];
            }
          }
        }
        mel_group_sync();   // staging free again
      }
    }
  }
}

}  // namespace ws
}  // namespace bhmel

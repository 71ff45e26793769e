// bhmel_kernel_ws.cuh -- warp-specialised schedule of the fused log-mel kernel (sm_100a).
//
// Same arithmetic as bhmel_logmel_kernel (bhmel_kernel.cuh: identical FFT passes, pair
// separation, paired banded mel, log epilogue -- results are bit-identical), different schedule:
// the stages run concurrently on different warps of one persistent 512-thread CTA per SM and
// hand 32-frame tiles to each other through mbarriers, so the latency-bound mel / store stage
// fills the issue slots the FFT warps leave idle instead of alternating with them.
//
//   warps 0-7   (2 warpgroups, setmaxnreg -> 184)  FFT role: two frame pairs per warp per tile
//   warps 8-15  (2 warpgroups, setmaxnreg ->  72)  producer + mel + store role
//
//   span[2]   4992-sample spans filled by the producer role one tile ahead of the FFT warps
//             (TMA bulk copy, or per-element cp.async with the reflect / zero mapping) -> span_full[2]
//   P[2]      power spectra [32][532] double buffered: FFT warps fill -> p_full[2], mel warps
//             release -> p_empty[2].  The two rows a frame pair will write double as that pair's
//             transpose scratch (real and imaginary planes are transposed one after the other), so
//             the double buffer costs no extra shared memory.
//   mel role  lane = frame, warp-uniform broadcast weights, statically unrolled pair dot
//             products (mel_chunk of bhmel_kernel.cuh), staging + coalesced store inside the
//             256-thread role group (named barrier 1).
#pragma once
#include "bhmel_kernel.cuh"

namespace bhmel {
// Output path of the direct static mel stages.  Each mel warp owns a staging block [32 frames][kStageCols
// filters] (row pitch kStagePitch floats: pitch / 4 odd, so the lane = frame 128-bit stores are conflict
// free); four finished filters of this lane's frame go in with one store (mel_stage4), and a full block
// is written out by the same warp as row segments of up to 128 contiguous bytes (mel_flush, one shared
// loop that also applies the log1p epilogue) -- float32, or bfloat16 rounded to nearest even.
constexpr int kStageCols = 32;
constexpr int kStagePitch = kStageCols + 4;
__device__ __forceinline__ void mel_stage4(float* __restrict__ srow, int col, float a, float b, float c, float d) {
  BH_CHECK(col >= 0 && col % 4 == 0 && col + 4 <= kStageCols);
  *reinterpret_cast<float4*>(srow + col) = make_float4(a, b, c, d);
}
// Writes the first `ncols` (multiple of 4) columns of the warp's staging block to rows 0..nf-1 of the
// tile: lane -> 16-byte column group lane & 7 of rows (lane >> 3) + 4 i, so one instruction moves four
// row segments of up to 128 contiguous bytes.
template <bool kBf16, bool kLog, int kPitch = kStagePitch>
__device__ __forceinline__ void mel_flush(const float* __restrict__ stage, void* __restrict__ ytile, long long fpitch,
                                          int m0, int ncols, int nf, int lane, bool vec, long long y_room = 0) {
  BH_CHECK(ncols % 4 == 0 && ncols + 4 <= kPitch && nf >= 1 && nf <= 32);
  __syncwarp();
  const int c = 4 * (lane & 7);
  if (c < ncols) {
    const float* sp = stage + (lane >> 3) * kPitch + c;
    long long off = static_cast<long long>(lane >> 3) * fpitch + m0 + c;
#pragma unroll
    for (int i = 0; i < kTileF / 4; ++i, sp += 4 * kPitch, off += 4 * fpitch) {
      if ((lane >> 3) + 4 * i < nf) {
        BH_CHECK(off >= 0 && off + 4 <= y_room);
        BH_CHECK(!vec || (kBf16 ? (reinterpret_cast<uintptr_t>(static_cast<__nv_bfloat16*>(ytile) + off) & 7) == 0
                                : (reinterpret_cast<uintptr_t>(static_cast<float*>(ytile) + off) & 15) == 0));
        float4 v = *reinterpret_cast<const float4*>(sp);
        if constexpr (kLog) {
          v.x = fast_log1p(v.x);
          v.y = fast_log1p(v.y);
          v.z = fast_log1p(v.z);
          v.w = fast_log1p(v.w);
        }
        if constexpr (kBf16) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
          __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(ytile) + off;
          if (vec) {
            uint2 q;
            q.x = *reinterpret_cast<const unsigned*>(&lo);
            q.y = *reinterpret_cast<const unsigned*>(&hi);
            *reinterpret_cast<uint2*>(yp) = q;
          } else {   // rows not 8-byte aligned (odd channel offset): element stores, same values
            yp[0] = lo.x; yp[1] = lo.y; yp[2] = hi.x; yp[3] = hi.y;
          }
        } else {
          float* yp = static_cast<float*>(ytile) + off;
          if (vec) {
            *reinterpret_cast<float4*>(yp) = v;
          } else {
            yp[0] = v.x; yp[1] = v.y; yp[2] = v.z; yp[3] = v.w;
          }
        }
      }
    }
  }
  __syncwarp();
}

}  // namespace bhmel

#include "mel_static_gen.h"

namespace bhmel {
namespace ws {

constexpr int kPPitchW = 532;                          // 532/4 odd -> LDS.128 conflict free; 2 rows >= 4224 B
constexpr int kFftWarps = 8;
#ifndef BHMEL_MEL_WARPS
#define BHMEL_MEL_WARPS 8
#endif
constexpr int kMelWarps = BHMEL_MEL_WARPS;                // 8 (2 warpgroups) or 4 (1 warpgroup)
constexpr int kMelThreads = kMelWarps * 32;
constexpr int kThreadsW = (kFftWarps + kMelWarps) * 32;   // 512
constexpr int kPlanePitch = 33;                        // floats per row of a transposed plane
// setmaxnreg budget: the CTA is launched with 128 registers/thread (65536 / 512); registers only
// move between the warpgroups of the CTA, so 2 * kFftRegs + 2 * kMelRegs must not exceed 4 * 128.
constexpr int kLaunchRegs = (65536 / kThreadsW) / 8 * 8;   // 128 for 512 threads, 168 for 384
#ifndef BHMEL_FFT_REGS
#define BHMEL_FFT_REGS 184
#endif
constexpr int kFftRegs = BHMEL_FFT_REGS;                   // tunable at build time (A/B experiments)
constexpr int kMelGroups = kMelWarps / 4;
constexpr int kMelRegs = ((2 + kMelGroups) * kLaunchRegs - 2 * kFftRegs) / kMelGroups / 8 * 8;   // 72 for 184 / 8 warps
static_assert(2 * kFftRegs + kMelGroups * kMelRegs <= (2 + kMelGroups) * kLaunchRegs, "setmaxnreg pool would deadlock");
static_assert(kMelRegs >= 24 && kMelRegs <= 256 && kFftRegs <= 256, "setmaxnreg range");
static_assert(2 * kPPitchW * 4 >= 32 * kPlanePitch * 4, "a pair's two P rows must hold one transposed plane");
static_assert(kMelWarps * kTileF * kStagePitch * 4 <= (kTileF * kOutPitch + kFwCap) * 4 + kPairCap * 16 &&
                  (kTileF * kOutPitch * 4) % 16 == 0 && (kFwCap * 4) % 16 == 0,
              "the direct stages' staging blocks overlay out / fw / pairs, which must be contiguous");

struct SmemWS {
  float P[2][kTileF * kPPitchW];              // 136 192 B
  float span[2][kSpan + 4];                   //  39 968 B  (+4: slack for the aligned-down TMA copy of unaligned rows)
  float out[kTileF * kOutPitch];              //  12 416 B  } generic / hybrid stages; the direct stages use the
  float fw[kFwCap];                           //  16 384 B  } same 36 992 bytes as eight private staging blocks
  int4 pairs[kPairCap];                       //   8 192 B  } [32][kStagePitch] (mel_stage())
  __device__ float* mel_stage() { return out; }
  unsigned long long span_full[2];
  unsigned long long p_full[2];
  unsigned long long p_empty[2];
  int span_delta[2];                          // element offset of the tile's first sample inside span[b]
};

#ifdef BHMEL_TRACE   // phase timeline of block 0 (tools/trace_phases.py); never defined in the shipped build
constexpr int kTraceIt0 = 64, kTraceIts = 4, kTraceEv = 16;
__device__ long long g_trace[16 * kTraceIts * kTraceEv];
#define BHMEL_TR(ev)                                                                                  \
  do {                                                                                                \
    if (blockIdx.x == 0 && lane == 0 && it >= kTraceIt0 && it < kTraceIt0 + kTraceIts)                \
      g_trace[(warp * kTraceIts + (it - kTraceIt0)) * kTraceEv + (ev)] = clock64();                   \
  } while (0)
#else
#define BHMEL_TR(ev) do {} while (0)
#endif

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mel_group_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(kMelThreads) : "memory"); }

// Producer: every thread of the mel role takes part, so span_full always sees kMelThreads arrivals.
__device__ __forceinline__ void issue_span(const KParams& p, long long tile, float* dst, unsigned long long* bar,
                                           int* delta_out, int mt) {
  const long long r = tile / p.tiles_per_row;
  const int tb = static_cast<int>(tile - r * p.tiles_per_row);
  const long long s0 = static_cast<long long>(tb) * (kTileF * kHop) - kNfft / 2;
  const long long row_off = p.row0 + r * p.row_stride;
  long long valid = p.n_total - row_off;
  valid = valid < 0 ? 0 : (valid > p.N ? p.N : valid);
  const float* row = p.x + row_off;
  const bool interior = s0 >= 0 && s0 + kSpan <= valid;
  // TMA needs a 16-byte aligned source: copy from the address rounded DOWN to 16 bytes (delta = 0..3
  // extra leading samples, 16 extra bytes in total) and let the FFT warps read from span + delta.
  // This keeps windows gathered at odd strides (52 415 in the reference) on the bulk-copy path.
  const int delta = static_cast<int>((reinterpret_cast<uintptr_t>(row + s0) >> 2) & 3);
  const bool bulk = p.use_bulk && interior && (reinterpret_cast<uintptr_t>(p.x) & 3) == 0 &&
                    (delta == 0 || (s0 >= delta && s0 - delta + kSpan + 4 <= valid));
  if (bulk) {
    if (mt == 0) {
      const uint32_t bytes = delta ? kSpanBytes + 16 : kSpanBytes;
      *delta_out = delta;
      BH_CHECK(s0 - delta >= 0 && s0 - delta + bytes / 4 <= valid && bytes / 4 <= kSpan + 4);
      BH_CHECK((reinterpret_cast<uintptr_t>(row + s0 - delta) & 15) == 0 && (smem_u32(dst) & 15) == 0);
      fence_proxy_async();
      mbar_expect_tx(bar, bytes);
      bulk_g2s(dst, row + s0 - delta, bytes, bar);
    } else {
      mbar_arrive(bar);
    }
    return;
  }
  if (mt == 0) {
    *delta_out = 0;
    __threadfence_block();   // performed before this thread's (asynchronous) arrival below
  }
  if (interior) {   // fully inside the row but not coverable by an aligned bulk copy: element copies
    const float* src = row + s0;
    BH_CHECK(s0 >= 0 && s0 + kSpan <= valid);
    for (int e = mt; e < kSpan; e += kMelThreads) cp_async_4(dst + e, src + e, 4);
  } else {
    const long long N = p.N;
    for (int e = mt; e < kSpan; e += kMelThreads) {
      long long i = s0 + e;
      if (i < 0) i = p.pad_reflect ? -i : -1;
      else if (i >= N) i = p.pad_reflect ? 2 * (N - 1) - i : -1;
      const bool ok = (i >= 0) && (i < valid);
      BH_CHECK(!ok || (i < N && row_off + i < p.n_total));
      cp_async_4(dst + e, row + (ok ? i : 0), ok ? 4 : 0);
    }
  }
  cp_async_arrive_noinc(bar);
}

// kStatic selects the mel stage: 0 = generic (pair descriptors + weight table, any filterbank);
// otherwise the id of a baked reference filterbank (bhmel_fb_baked.h kBakedFbs):
// 1 = P0: mel warps 0..kStaticP0Warps-1 run generated straight-line code (mel_static_gen.h: weights
// as FFMA immediates, every power block read once) for filters 0..kStaticP0Filters-1, the other mel
// warps run the generic stage on the pair tables of the remaining filters;
// >= 2 (P128, P1, T5; kStaticP0Direct = P0 again, for A/B runs): the DIRECT form -- all eight mel
// warps run generated code, one block per warp and tile, results leave through the warp's private
// staging block (mel_stage4 / mel_flush): no chunk loop, no cross-warp barriers.  kBf16 is the output type
// of the direct forms (the others read p.y_bf16).
// kStatic <= 1 is bit-identical to the generic stage; the direct forms sum each filter in one or two
// chains instead of four and agree with it to a few ulp.
template <bool kLog, int kStatic = 0, bool kBf16 = false>
__global__ void __launch_bounds__(kThreadsW, 1) bhmel_logmel_ws_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemWS& S = *reinterpret_cast<SmemWS*>(smem_raw);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  static_assert(kStatic == 0 || (kMelWarps == 8 && kStaticP0Warps < kMelWarps), "static mel stage: 8 mel warps");
  static_assert(kStatic >= 0 && kStatic <= kStaticP0Direct, "kStatic is 0 or the id of a baked filterbank");
  constexpr bool kDirect = kStatic >= 2;
  static_assert(kDirect || !kBf16, "kBf16 only selects the store type of the direct forms");
  const bool fw_in_smem = p.n_weights <= kFwCap;   // the host guarantees this for kStatic != 0
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next kernel's CTAs may take this SM as soon as we exit
  if (fw_in_smem)
    for (int i = tid; i < p.n_weights; i += kThreadsW) S.fw[i] = p.weights[i];
  for (int i = tid; i < p.n_pairs; i += kThreadsW) S.pairs[i] = reinterpret_cast<const int4*>(p.pairs)[i];
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
    const int src = (32 - lane) & 31;
    int it = 0;
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&S.span_full[b], ph);
      mbar_wait(&S.p_empty[b], ph ^ 1);   // the mel role has released this P buffer (tile it-2)
      BHMEL_TR(0);
      const float* span_b = S.span[b] + S.span_delta[b];
#ifdef BHMEL_DEBUG_SKIP_FFT
      if (tile < 0)
#endif
#pragma unroll 1
      for (int j = warp; j < kPairs; j += kFftWarps) {
        float* rows = S.P[b] + (2 * j) * kPPitchW;      // this pair's two P rows; scratch until written
        float ar[32], ai[32];
        {
          float v[36];
          const float* sp = span_b + (2 * j) * kHop + lane;
          BH_CHECK(S.span_delta[b] >= 0 && S.span_delta[b] <= 3 && (sp - S.span[b]) + 32 * 35 < kSpan + 4);
#pragma unroll
          for (int m = 0; m < 36; ++m) v[m] = sp[32 * m];
          fft32_pass_a(v, wreg, ar, ai);
        }
        BHMEL_TR(1 + 5 * (j / kFftWarps));
        // transpose the real plane, then the imaginary plane, through the pair's own rows
        float ur[32], ui[32];
        BH_CHECK((rows - S.P[b]) + 31 * kPlanePitch + 31 < kTileF * kPPitchW && 31 * kPlanePitch + 31 < 2 * kPPitchW);
#pragma unroll
        for (int k = 0; k < 32; ++k) rows[k * kPlanePitch + lane] = ar[k];
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 32; ++n) ur[n] = rows[lane * kPlanePitch + n];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) rows[k * kPlanePitch + lane] = ai[k];
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 32; ++n) ui[n] = rows[lane * kPlanePitch + n];
        __syncwarp();
        BHMEL_TR(2 + 5 * (j / kFftWarps));
        float br[32], bi[32];
        fft32_pass_b(ur, ui, twr, twi, br, bi);
        BHMEL_TR(3 + 5 * (j / kFftWarps));
        float* Pa = rows + lane;
        float* Pb = Pa + kPPitchW;
        BH_CHECK((Pb - S.P[b]) + 32 * 15 < kTileF * kPPitchW && (lane != 0 || (Pb - S.P[b]) + 512 < kTileF * kPPitchW) &&
                 kBins + 2 < kPPitchW);
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
        if (lane < 6) rows[(lane / 3) * kPPitchW + kBins + lane % 3] = 0.f;   // bins 513..515 read as zero weights' partners
        BHMEL_TR(4 + 5 * (j / kFftWarps));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.p_full[b]);
    }
  } else {
    // ======================= producer + mel + store role =======================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kMelRegs));
    // Programmatic dependent launch (launch(), BHMEL_OPT_PDL): this grid may have become resident while the
    // previous kernel in the stream was draining.  Everything above touches only the handle's constant tables;
    // the samples are read and the output written by this role alone, after the previous grid has completed.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int mt = tid - kFftWarps * 32;
    const int mw = warp - kFftWarps;
    {
      long long t0 = blockIdx.x;
      if (t0 < p.n_tiles) issue_span(p, t0, S.span[0], &S.span_full[0], &S.span_delta[0], mt);
      t0 += gridDim.x;
      if (t0 < p.n_tiles) issue_span(p, t0, S.span[1], &S.span_full[1], &S.span_delta[1], mt);
    }
    // tile -> (row, 32-frame block) is tracked incrementally: tile advances by gridDim.x per step
    const int tpr = p.tiles_per_row;
    const int step_r = static_cast<int>(gridDim.x / tpr), step_tb = static_cast<int>(gridDim.x % tpr);
    int r = static_cast<int>(blockIdx.x / tpr);   // rows fit an int (checked by the host)
    int tb = static_cast<int>(blockIdx.x % tpr);
    int it = 0;
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      BHMEL_TR(0);
      mbar_wait_idle(&S.p_full[b], ph);
      BHMEL_TR(1);
      // every FFT warp is done with tile `it`, so span[b] is free: fetch the tile two steps ahead
      const long long tile2 = tile + 2 * static_cast<long long>(gridDim.x);
      if (tile2 < p.n_tiles) issue_span(p, tile2, S.span[b], &S.span_full[b], &S.span_delta[b], mt);

      const int t0 = tb * kTileF;
      const long long frames_left = p.T - t0;
      const int nf = frames_left < kTileF ? static_cast<int>(frames_left) : kTileF;
      const long long ybase = static_cast<long long>(r) * p.y_row_pitch + static_cast<long long>(t0) * p.y_frame_pitch;
      const float4* prow = reinterpret_cast<const float4*>(S.P[b] + lane * kPPitchW);
      float* orow = S.out + lane * kOutPitch;
      if constexpr (kDirect) {
        void* ytile = kBf16 ? static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(p.y) + ybase) : static_cast<void*>(p.y + ybase);
        float* stage = S.mel_stage() + mw * (kTileF * kStagePitch);
        constexpr int kParts = mel_direct_parts<kStatic>();
#pragma unroll 1
        for (int part = 0; part < kParts; ++part) {
#ifdef BHMEL_DEBUG_ONE_BLOCK   // timing experiment (wrong results): every mel warp runs warp 0's code -> 1/8 of the footprint
          mel_direct<kStatic>(prow, stage + lane * kStagePitch, 0, part);
#else
          mel_direct<kStatic>(prow, stage + lane * kStagePitch, mw, part);
#endif
          if (part == kParts - 1) {   // this warp no longer reads P[b]
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.p_empty[b]);
          }
          int m0, ncols;
          mel_direct_run<kStatic>(mw * kParts + part, m0, ncols);
          BH_CHECK(m0 + ncols <= p.n_mels && (stage - S.mel_stage()) + kTileF * kStagePitch <= kMelWarps * kTileF * kStagePitch);
          mel_flush<kBf16, kLog>(stage, ytile, p.y_frame_pitch, m0, ncols, nf, lane, p.y_vec_ok != 0, p.y_limit - ybase);
        }
      } else {
#ifdef BHMEL_DEBUG_SKIP_MEL
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.p_empty[b]);
      if (tile < 0)
#endif
      for (int mc = 0, c = 0; mc < p.n_mels; mc += kMChunk, ++c) {
        const int mcount = (p.n_mels - mc) < kMChunk ? (p.n_mels - mc) : kMChunk;
        const int4* pd = S.pairs + c * (kMChunk / 2);
#ifdef BHMEL_DEBUG_SKIP_DOTS
        if (tile < 0)
#endif
        if constexpr (kStatic == 1) {
          if (mw < kStaticP0Warps) mel_static_P0<kLog>(prow, orow, mw);
          else mel_chunk<true, kLog, kMelWarps - kStaticP0Warps>(prow, pd, p.n_pairs, S.fw, orow, mw - kStaticP0Warps);
        }
        else if (fw_in_smem) mel_chunk<true, kLog, kMelWarps>(prow, pd, (mcount + 1) >> 1, S.fw, orow, mw);
        else mel_chunk<false, kLog, kMelWarps>(prow, pd, (mcount + 1) >> 1, p.weights, orow, mw);
        BHMEL_TR(2);
        if (mc + kMChunk >= p.n_mels) {   // last chunk: this warp no longer reads P[b]
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.p_empty[b]);
        }
        mel_group_sync();   // staging complete
        BHMEL_TR(3);
        {
          constexpr int kFr = kTileF / kMelWarps, kCo = kMChunk / 32;   // 4 frames x 3 column steps
          float vals[kFr][kCo];
#pragma unroll
          for (int a = 0; a < kFr; ++a)
#pragma unroll
            for (int bb = 0; bb < kCo; ++bb) vals[a][bb] = S.out[(mw + a * kMelWarps) * kOutPitch + lane + 32 * bb];
          const long long fpitch = p.y_frame_pitch;
          const long long y0 = ybase + static_cast<long long>(mw) * fpitch + mc + lane;
          BH_CHECK((mw + (kFr - 1) * kMelWarps) * kOutPitch + lane + 32 * (kCo - 1) < kTileF * kOutPitch);
#ifdef BHMEL_BOUNDS
          for (int a = 0; a < kFr; ++a)
            for (int bb = 0; bb < kCo; ++bb)
              if (mw + a * kMelWarps < nf && lane + 32 * bb < mcount)
                BH_CHECK(y0 + a * kMelWarps * fpitch + 32 * bb >= 0 && y0 + a * kMelWarps * fpitch + 32 * bb < p.y_limit);
#endif
          if (!p.y_bf16) {
            float* yp = p.y + y0;
#pragma unroll
            for (int a = 0; a < kFr; ++a, yp += kMelWarps * fpitch) {
              if (mw + a * kMelWarps < nf) {
#pragma unroll
                for (int bb = 0; bb < kCo; ++bb)
                  if (lane + 32 * bb < mcount) yp[32 * bb] = vals[a][bb];
              }
            }
          } else {
            __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + y0;
#pragma unroll
            for (int a = 0; a < kFr; ++a, yp += kMelWarps * fpitch) {
              if (mw + a * kMelWarps < nf) {
#pragma unroll
                for (int bb = 0; bb < kCo; ++bb)
                  if (lane + 32 * bb < mcount) yp[32 * bb] = __float2bfloat16_rn(vals[a][bb]);
              }
            }
          }
        }
        BHMEL_TR(4);
        mel_group_sync();   // staging free again
        BHMEL_TR(5);
      }
      }
      r += step_r;
      tb += step_tb;
      if (tb >= tpr) { tb -= tpr; ++r; }
    }
  }
}

}  // namespace ws
}  // namespace bhmel

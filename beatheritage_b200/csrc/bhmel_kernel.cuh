// bhmel_kernel.cuh -- the fused log-mel kernel for sm_100a.
//
// Replaces, in ONE launch, what the reference reaches through
// osuT5/osuT5/model/spectrogram.py:79-82 (transform -> log1p -> permute), i.e.
// F.pad(reflect|constant) -> torch.stft (frame gather * Hann, 1024-pt rFFT) -> |X|^2 ->
// matmul(fb) -> log1p -> [B, T, n_mels]   (SURVEY.md 8a rows a2..a9).
//
// Work decomposition
//   tile        = 32 consecutive frames of one input row (window); persistent CTAs (one per SM)
//                 walk tiles with a static stride.
//   stage 1     the 4992-sample span a tile needs is staged in shared memory, double buffered:
//                 interior, 16-byte aligned tiles by ONE cp.async.bulk (TMA bulk copy,
//                 mbarrier completion); edge / unaligned tiles by per-element cp.async with the
//                 reflect / zero index mapping (so every window is padded on its own).
//   stage 2     one warp transforms a PAIR of frames as one 1024-pt complex FFT split 32 x 32:
//                 lane = fast sample index, in-register 32-pt pass A over the slow index with
//                 the Hann window fused into the first butterflies, warp transpose through
//                 padded shared memory, in-register 32-pt pass B with the inter-pass twiddle
//                 fused, then the two real spectra are separated with one shuffle per value and
//                 |X|^2 goes to the tile's power buffer P[32][516].
//   stage 3     banded mel projection with lane = frame (weights are warp-uniform, P rows are
//                 read with conflict-free 128-bit loads), log1p epilogue, staged through shared
//                 memory so the [frames, n_mels] tile is written with coalesced stores.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "bhmel_tables.h"

#ifndef BHMEL_HD
#define BHMEL_HD __device__ __forceinline__
#endif
#include "fft32_gen.h"

// -DBHMEL_BOUNDS (debug library libbhmel_bounds.so, built by beatheritage_b200.build and exercised by
// tests/test_gpu_bounds.py): every shared- and global-memory index of the staging, transpose, power-buffer,
// mel and store paths is asserted; a violation prints the expression and traps (compute-sanitizer is
// closed on the GPU pool this was developed on).  Compiled out of the shipped library.
#ifdef BHMEL_BOUNDS
#include <cstdio>
#define BH_CHECK(cond)                                                                                          \
  do {                                                                                                          \
    if (!(cond)) {                                                                                              \
      printf("BHMEL_BOUNDS violated: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__,            \
             static_cast<int>(blockIdx.x), static_cast<int>(threadIdx.x));                                     \
      __trap();                                                                                                 \
    }                                                                                                           \
  } while (0)
#else
#define BH_CHECK(cond) do {} while (0)
#endif

namespace bhmel {

constexpr int kTileF = 32;                                // frames per tile
constexpr int kSpan = (kTileF - 1) * kHop + kNfft;        // 4992 samples
constexpr int kSpanBytes = kSpan * 4;                     // 19968 (multiple of 16)
constexpr int kPPitch = 516;                              // 513 bins + pad; /4 odd -> LDS.128 conflict-free
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kScrPitch = 33;                             // complex elements per transpose row
constexpr int kMChunk = kMelChunk;                        // mel filters per epilogue chunk (96)
constexpr int kOutPitch = kMChunk + 1;                    // odd -> conflict-free lane=frame writes
constexpr int kPairs = kTileF / 2;
constexpr int kFwCap = 4096;                              // floats of filter weights kept in shared memory
constexpr int kPairCap = 512;                             // filter-pair descriptors (n_mels <= 1024)

struct SmemLayout {
  float P[kTileF * kPPitch];                  // 66 048 B  power spectra of the tile
  float2 scr[kWarps][32 * kScrPitch];         // 67 584 B  per-warp transpose scratch
  float span[2][kSpan];                       // 39 936 B  double-buffered sample span
  float out[kTileF * kOutPitch];              // 12 416 B  epilogue staging
  float fw[kFwCap];                           // 16 384 B  banded filter weights (when they fit)
  int4 pairs[kPairCap];                       //  8 192 B  filter-pair descriptors (PairDesc)
  unsigned long long mbar[2];
};

struct KParams {
  const float* x;          // module mode: [B][row_stride]; gather mode: the song
  long long row_stride;    // elements between consecutive rows (gather: window stride)
  long long row0;          // offset of row 0 inside x (gather: first_offset)
  long long n_total;       // samples at absolute index >= n_total read as zero
  long long N;             // logical row length (samples per window)
  long long T;             // frames per row = N / hop + 1
  long long n_tiles;
  float* y;                // output base: element (row r, frame t, mel m) at r*y_row_pitch + t*y_frame_pitch + m
  long long y_row_pitch;   // elements between rows     (T * n_mels for the plain [B][T][n_mels] layout)
  long long y_frame_pitch; // elements between frames   (n_mels for the plain layout)
  const float* win_half;   // [1024]
  const float2* tw;        // [32*32]
  const PairDesc* pairs;   // [n_pairs]  (bhmel_tables.h make_pairs)
  const float* weights;    // interleaved pair weights
  int B;
  int tiles_per_row;
  int n_mels;
  int pad_reflect;
  int log_scale;
  int use_bulk;
  int n_weights;           // floats in `weights` (multiple of 8)
  int n_pairs;
  int y_bf16;              // 0: float32 output, 1: bfloat16 (round to nearest even)
  long long y_limit;       // one past the largest output element index the call may write (BH_CHECK only)
  int y_vec_ok;            // direct mel stages: output rows are 16-byte (f32) / 8-byte (bf16) aligned -> vector stores
};

// Output element store: float32 or bfloat16, any frame / row pitch (N1: writes the mel channels
// straight into a wider encoder-input buffer).
__device__ __forceinline__ void store_out(const KParams& p, long long idx, float v) {
  if (p.y_bf16) reinterpret_cast<__nv_bfloat16*>(p.y)[idx] = __float2bfloat16_rn(v);
  else p.y[idx] = v;
}

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async_4(void* dst, const float* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(dst)), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// The spin loop is written in C++ (not inside the asm block) so the compiler knows about the
// branch and reconverges the warp afterwards.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Waiting for a barrier that is expected to take a while (the mel role waiting for the FFT role's next
// tile): every poll is an issue slot taken from the warps that do the work -- ncu showed 28 % of the
// kernel's executed instructions in these loops -- so the poll carries a suspend-time hint (the warp is
// parked in hardware until the phase completes or the time is up) and backs off with nanosleep.
#ifndef BHMEL_IDLE_HINT_NS
#define BHMEL_IDLE_HINT_NS 20000
#endif
#ifndef BHMEL_IDLE_SLEEP_NS
#define BHMEL_IDLE_SLEEP_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait_hint(unsigned long long* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_idle(unsigned long long* bar, uint32_t parity) {
  while (!mbar_try_wait_hint(bar, parity, BHMEL_IDLE_HINT_NS)) {
    if (BHMEL_IDLE_SLEEP_NS > 0) __nanosleep(BHMEL_IDLE_SLEEP_NS);
  }
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }

// ---------------------------------------------------------------- stage 1: span staging
// Issues the asynchronous fill of span buffer `dst` for `tile`; returns true when the TMA bulk
// path was used (caller then waits on the mbarrier), false for the cp.async path.
__device__ __forceinline__ bool stage_span(const KParams& p, long long tile, float* dst,
                                           unsigned long long* bar, int tid) {
  const long long r = tile / p.tiles_per_row;
  const int tb = static_cast<int>(tile - r * p.tiles_per_row);
  const long long s0 = static_cast<long long>(tb) * (kTileF * kHop) - kNfft / 2;   // first sample of the span
  const long long row_off = p.row0 + r * p.row_stride;
  long long valid = p.n_total - row_off;                  // real samples available in this row
  valid = valid < 0 ? 0 : (valid > p.N ? p.N : valid);
  const float* row = p.x + row_off;
  const bool bulk = p.use_bulk && s0 >= 0 && s0 + kSpan <= valid &&
                    ((reinterpret_cast<uintptr_t>(row + s0) & 15) == 0);
  if (bulk) {
    if (tid == 0) {
      fence_proxy_async();
      mbar_expect_tx(bar, kSpanBytes);
      bulk_g2s(dst, row + s0, kSpanBytes, bar);
    }
    return true;
  }
  const long long N = p.N;
  for (int e = tid; e < kSpan; e += kThreads) {
    long long i = s0 + e;
    if (i < 0) i = p.pad_reflect ? -i : -1;
    else if (i >= N) i = p.pad_reflect ? 2 * (N - 1) - i : -1;
    const bool ok = (i >= 0) && (i < valid);
    cp_async_4(dst + e, row + (ok ? i : 0), ok ? 4 : 0);   // src-size 0 -> zero fill
  }
  cp_async_commit();
  return false;
}

// ---------------------------------------------------------------- stage 3 helpers
// log1p epilogue: the argument 1 + v is >= 1 (never denormal), so the bare lg2.approx is enough
// -- no range-fix instructions like __logf emits.  |error| < 3e-6 in the log domain for v < 1e8.
__device__ __forceinline__ float fast_log1p(float v) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + v));
  return r * 0.693147180559945f;
}

// Dot products of one filter PAIR (NG groups of 4 bins each, zero padded to equal length) with
// this lane's frame: P rows via conflict-free LDS.128, weights via warp-uniform 128-bit loads
// (shared memory when the table fits, else the read-only global path).  Eight independent FMA
// chains per lane.
template <bool kSmemW>
__device__ __forceinline__ float4 ldw(const float4* p) {
  if constexpr (kSmemW) return *p;
  else return __ldg(p);
}
template <int NG, bool kSmemW>
__device__ __forceinline__ void band_dot2(const float4* __restrict__ pa, const float4* __restrict__ pb,
                                          const float4* __restrict__ wp, float& va, float& vb) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const float4 wa = ldw<kSmemW>(wp + 2 * g), wb = ldw<kSmemW>(wp + 2 * g + 1);
    const float4 xa = pa[g], xb = pb[g];
    a0 = fmaf(xa.x, wa.x, a0); a1 = fmaf(xa.y, wa.y, a1); a2 = fmaf(xa.z, wa.z, a2); a3 = fmaf(xa.w, wa.w, a3);
    b0 = fmaf(xb.x, wb.x, b0); b1 = fmaf(xb.y, wb.y, b1); b2 = fmaf(xb.z, wb.z, b2); b3 = fmaf(xb.w, wb.w, b3);
  }
  va = (a0 + a1) + (a2 + a3);
  vb = (b0 + b1) + (b2 + b3);
}
template <bool kSmemW>
__device__ __forceinline__ void band_dot2_n(const float4* __restrict__ pa, const float4* __restrict__ pb,
                                            const float4* __restrict__ wp, int ng, float& va, float& vb) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll 2
  for (int g = 0; g < ng; ++g) {
    const float4 wa = ldw<kSmemW>(wp + 2 * g), wb = ldw<kSmemW>(wp + 2 * g + 1);
    const float4 xa = pa[g], xb = pb[g];
    a0 = fmaf(xa.x, wa.x, a0); a1 = fmaf(xa.y, wa.y, a1); a2 = fmaf(xa.z, wa.z, a2); a3 = fmaf(xa.w, wa.w, a3);
    b0 = fmaf(xb.x, wb.x, b0); b1 = fmaf(xb.y, wb.y, b1); b2 = fmaf(xb.z, wb.z, b2); b3 = fmaf(xb.w, wb.w, b3);
  }
  va = (a0 + a1) + (a2 + a3);
  vb = (b0 + b1) + (b2 + b3);
}

// Fall-through form of the pair dot product: one copy of the group body per possible remaining
// count, entered at `ng` and run to the end, so the hot mel code is ~10 group bodies instead of
// 1 + 2 + ... + 10 (the kernel's hot loop has to fit the 32 KB instruction cache next to the FFT
// role's ~24 KB; DESIGN.md "instruction footprint").  Groups are addressed back from the END of
// the band, which keeps the accumulation order ascending -- bit-identical to band_dot2<NG>.
template <bool kSmemW>
__device__ __forceinline__ void band_dot2_ft(const float4* __restrict__ pa, const float4* __restrict__ pb,
                                             const float4* __restrict__ wp, int ng, float& va, float& vb) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
  const float4* pae = pa + ng;
  const float4* pbe = pb + ng;
  const float4* wpe = wp + 2 * ng;
#define BHMEL_GROUP(L)                                                                                   \
  {                                                                                                      \
    const float4 wa = ldw<kSmemW>(wpe - 2 * (L)), wb = ldw<kSmemW>(wpe - 2 * (L) + 1);                   \
    const float4 xa = pae[-(L)], xb = pbe[-(L)];                                                         \
    a0 = fmaf(xa.x, wa.x, a0); a1 = fmaf(xa.y, wa.y, a1); a2 = fmaf(xa.z, wa.z, a2); a3 = fmaf(xa.w, wa.w, a3); \
    b0 = fmaf(xb.x, wb.x, b0); b1 = fmaf(xb.y, wb.y, b1); b2 = fmaf(xb.z, wb.z, b2); b3 = fmaf(xb.w, wb.w, b3); \
  }
  switch (ng) {
    default:
      for (int L = ng; L > 10; --L) BHMEL_GROUP(L)
      [[fallthrough]];
    case 10: BHMEL_GROUP(10) [[fallthrough]];
    case 9: BHMEL_GROUP(9) [[fallthrough]];
    case 8: BHMEL_GROUP(8) [[fallthrough]];
    case 7: BHMEL_GROUP(7) [[fallthrough]];
    case 6: BHMEL_GROUP(6) [[fallthrough]];
    case 5: BHMEL_GROUP(5) [[fallthrough]];
    case 4: BHMEL_GROUP(4) [[fallthrough]];
    case 3: BHMEL_GROUP(3) [[fallthrough]];
    case 2: BHMEL_GROUP(2) [[fallthrough]];
    case 1: BHMEL_GROUP(1) [[fallthrough]];
    case 0: break;
  }
#undef BHMEL_GROUP
  va = (a0 + a1) + (a2 + a3);
  vb = (b0 + b1) + (b2 + b3);
}

// One chunk (<= kMChunk output columns) of the mel projection for the 32 frames of the tile:
// warp w takes pairs w, w + kWarps, ...; results go to the staging buffer [frame][column].
template <bool kSmemW, bool kLog, int kStride = kWarps>
__device__ __forceinline__ void mel_chunk(const float4* __restrict__ prow, const int4* __restrict__ pd, int npairs,
                                          const float* __restrict__ wbase, float* __restrict__ orow, int warp) {
  int4 dnext = pd[warp < npairs ? warp : 0];              // PairDesc, warp-uniform; prefetched one step ahead
#pragma unroll 1
  for (int q = warp; q < npairs; q += kStride) {
    const int4 d = dnext;
    if (q + kStride < npairs) dnext = pd[q + kStride];
    const float4* wp = reinterpret_cast<const float4*>(wbase + d.y);
    const float4* pa = prow + (d.x & 0xFFFF);
    const float4* pb = prow + (static_cast<unsigned>(d.x) >> 16);
    BH_CHECK(d.z >= 0 && (d.x & 0xFFFF) + d.z <= (kBins + 3) / 4 && (static_cast<unsigned>(d.x) >> 16) + d.z <= (kBins + 3) / 4);
    BH_CHECK(d.y >= 0 && d.y % 4 == 0);
    float va, vb;
#if defined(BHMEL_MEL_COMPACT)
    constexpr int kForm = 1;
#elif defined(BHMEL_MEL_UNROLLED)
    constexpr int kForm = kSmemW ? 2 : 1;   // rare path (huge / dense filterbanks): keep the code small
#else
    constexpr int kForm = kSmemW ? 0 : 1;
#endif
    if constexpr (kForm == 0) {
      band_dot2_ft<kSmemW>(pa, pb, wp, d.z, va, vb);
    } else if constexpr (kForm == 1) {
      band_dot2_n<kSmemW>(pa, pb, wp, d.z, va, vb);
    } else {
    switch (d.z) {
      case 0: va = 0.f; vb = 0.f; break;
      case 1: band_dot2<1, kSmemW>(pa, pb, wp, va, vb); break;
      case 2: band_dot2<2, kSmemW>(pa, pb, wp, va, vb); break;
      case 3: band_dot2<3, kSmemW>(pa, pb, wp, va, vb); break;
      case 4: band_dot2<4, kSmemW>(pa, pb, wp, va, vb); break;
      case 5: band_dot2<5, kSmemW>(pa, pb, wp, va, vb); break;
      case 6: band_dot2<6, kSmemW>(pa, pb, wp, va, vb); break;
      case 7: band_dot2<7, kSmemW>(pa, pb, wp, va, vb); break;
      case 8: band_dot2<8, kSmemW>(pa, pb, wp, va, vb); break;
      case 9: band_dot2<9, kSmemW>(pa, pb, wp, va, vb); break;
      case 10: band_dot2<10, kSmemW>(pa, pb, wp, va, vb); break;
      default: band_dot2_n<kSmemW>(pa, pb, wp, d.z, va, vb); break;
    }
    }
    if constexpr (kLog) {
      va = fast_log1p(va);
      vb = fast_log1p(vb);
    }
    const int ca = d.w & 0xFFFF, cb = static_cast<unsigned>(d.w) >> 16;
    BH_CHECK(ca < kMChunk && (cb == 0xFFFF || cb < kMChunk));
    orow[ca] = va;
    if (cb != 0xFFFF) orow[cb] = vb;
  }
}

// ---------------------------------------------------------------- the kernel
template <bool kLog>
__global__ void __launch_bounds__(kThreads, 1) bhmel_logmel_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemLayout& S = *reinterpret_cast<SmemLayout*>(smem_raw);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  // one-time: this lane's window samples and inter-pass twiddles live in registers for the whole
  // kernel (one CTA per SM leaves 255 registers per thread); filter tables -> shared memory when
  // they fit; zero the pad columns of P; init mbarriers
  float wreg[32], twr[32], twi[32];
#pragma unroll
  for (int m = 0; m < 32; ++m) {
    wreg[m] = __ldg(p.win_half + lane + 32 * m);
    const float2 t = __ldg(p.tw + m * 32 + lane);
    twr[m] = t.x;
    twi[m] = t.y;
  }
  const bool fw_in_smem = p.n_weights <= kFwCap;
  if (fw_in_smem)
    for (int i = tid; i < p.n_weights; i += kThreads) S.fw[i] = p.weights[i];
  for (int i = tid; i < p.n_pairs; i += kThreads) S.pairs[i] = reinterpret_cast<const int4*>(p.pairs)[i];
  for (int i = tid; i < kTileF * (kPPitch - kBins); i += kThreads)
    S.P[(i / (kPPitch - kBins)) * kPPitch + kBins + i % (kPPitch - kBins)] = 0.f;
  if (tid == 0) {
    mbar_init(&S.mbar[0], 1);
    mbar_init(&S.mbar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  uint32_t parity0 = 0, parity1 = 0;
  long long tile = blockIdx.x;
  int buf = 0;
  bool cur_bulk = false;
  if (tile < p.n_tiles) cur_bulk = stage_span(p, tile, S.span[0], &S.mbar[0], tid);

  for (; tile < p.n_tiles; tile += gridDim.x, buf ^= 1) {
    // ---- wait for this tile's span ------------------------------------------------------
    cp_async_wait_all();
    if (cur_bulk) {   // warp 0 observes the TMA completion; the barrier below publishes it to the CTA
      if (buf == 0) { if (warp == 0) mbar_wait(&S.mbar[0], parity0); parity0 ^= 1; }
      else          { if (warp == 0) mbar_wait(&S.mbar[1], parity1); parity1 ^= 1; }
    }
    __syncthreads();   // span[buf] visible to all; every thread is done with the previous tile

    // ---- prefetch the next tile's span into the other buffer -----------------------------
    const long long next_tile = tile + gridDim.x;
    bool next_bulk = false;
    if (next_tile < p.n_tiles)
      next_bulk = stage_span(p, next_tile, S.span[buf ^ 1], &S.mbar[buf ^ 1], tid);

    // ---- stage 2: FFT of frame pairs ------------------------------------------------------
    const float* span = S.span[buf];
    float2* scr = S.scr[warp];
#pragma unroll 1
    for (int j = warp; j < kPairs; j += kWarps) {
      float ar[32], ai[32];
      {
        float v[36];
        const float* sp = span + (2 * j) * kHop + lane;
#pragma unroll
        for (int m = 0; m < 36; ++m) v[m] = sp[32 * m];
        fft32_pass_a(v, wreg, ar, ai);
      }
      // transpose: thread (lane = n2) holds Y[k1] -> thread (lane = k1) holds Y[n2]
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
        __syncwarp();   // all lanes have read the scratch before the next pair overwrites it
        fft32_pass_b(ur, ui, twr, twi, br, bi);
      }
      // Z[k], k = lane + 32*k2.  Partner bin 1024-k lives in lane (32-lane)&31, slot 31-k2
      // (lane 0: its own slot (32-k2)&31).  Xa = (Z[k] + conj Z[N-k])/2, Xb = (Z[k] - conj Z[N-k])/2i;
      // the window carries the factor 1/2, so the sums below are Xa, Xb themselves.
      const int src = (32 - lane) & 31;
      float* Pa = S.P + (2 * j) * kPPitch + lane;
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
      if (lane == 0) {   // bin 512 = (k1 = 0, k2 = 16) is its own partner
        const float zr = 2.f * br[16], zi = 2.f * bi[16];
        Pa[512] = zr * zr;
        Pb[512] = zi * zi;
      }
    }
    __syncthreads();   // P complete

    // ---- stage 3: banded mel projection + log1p + coalesced store ------------------------
    const long long r = tile / p.tiles_per_row;
    const int t0 = static_cast<int>(tile - r * p.tiles_per_row) * kTileF;
    const long long frames_left = p.T - t0;
    const int nf = frames_left < kTileF ? static_cast<int>(frames_left) : kTileF;
    const long long ybase = r * p.y_row_pitch + static_cast<long long>(t0) * p.y_frame_pitch;
    const float4* prow = reinterpret_cast<const float4*>(S.P + lane * kPPitch);
    for (int mc = 0, c = 0; mc < p.n_mels; mc += kMChunk, ++c) {
      const int mcount = (p.n_mels - mc) < kMChunk ? (p.n_mels - mc) : kMChunk;
      if (mc > 0) __syncthreads();   // previous chunk's staging fully stored
      const int4* pd = S.pairs + c * (kMChunk / 2);
      float* orow = S.out + lane * kOutPitch;
      if (fw_in_smem) mel_chunk<true, kLog>(prow, pd, (mcount + 1) >> 1, S.fw, orow, warp);
      else mel_chunk<false, kLog>(prow, pd, (mcount + 1) >> 1, p.weights, orow, warp);
      __syncthreads();
      // coalesced store of the staged [frames][columns] block: all loads first, then all stores
      {
        constexpr int kFr = kTileF / kWarps, kCo = kMChunk / 32;   // 4 frames x 3 column steps per thread
        float vals[kFr][kCo];
#pragma unroll
        for (int a = 0; a < kFr; ++a)
#pragma unroll
          for (int b = 0; b < kCo; ++b) vals[a][b] = S.out[(warp + a * kWarps) * kOutPitch + lane + 32 * b];
#pragma unroll
        for (int a = 0; a < kFr; ++a) {
          const int f = warp + a * kWarps;
          const long long yrow = ybase + static_cast<long long>(f) * p.y_frame_pitch + mc;
#pragma unroll
          for (int b = 0; b < kCo; ++b) {
            const int c2 = lane + 32 * b;
            if (f < nf && c2 < mcount) store_out(p, yrow + c2, vals[a][b]);
          }
        }
      }
    }
    cur_bulk = next_bulk;
  }
}

}  // namespace bhmel

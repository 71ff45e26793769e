// CPU lane-emulator of bhmel_logmel_kernel -- TEST INFRASTRUCTURE ONLY (never shipped, never
// on the product path).  It executes the kernel's algorithm warp by warp on the host with the
// SAME generated FFT passes (fft32_gen.h), the SAME constant tables (bhmel_tables.h) and the same
// index arithmetic (span staging with reflect/zero mapping, 32-frame tiles, frame pairs, the
// lane <-> bin pairing of the two-real-FFTs-in-one-complex trick, banded mel, log epilogue), so
// the no-GPU test suite can check everything except CUDA-specific mechanics (barriers, TMA).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

#define BHMEL_HD static inline
#include "../../beatheritage_b200/csrc/fft32_gen.h"
#include "../../beatheritage_b200/csrc/bhmel_tables.h"
#include "../../beatheritage_b200/csrc/bhmel_fb_baked.h"

// host shims for the generated static mel stage (mel_static_gen.h is device code)
struct float4 { float x, y, z, w; };
#define __device__
#define __forceinline__ inline
#define __constant__
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __fmul_rn(float a, float b) { return a * b; }   // built with -ffp-contract=off
namespace bhmel {
static inline float fast_log1p(float v) { return logf(1.0f + v); }
// direct static stages: the staging block of one frame (lane) is a plain array here, mel_flush copies it out
constexpr int kStageCols = 32, kStagePitch = kStageCols + 4;
static inline void mel_stage4(float* srow, int col, float a, float b, float c, float d) {
  srow[col] = a; srow[col + 1] = b; srow[col + 2] = c; srow[col + 3] = d;
}
}  // namespace bhmel
#include "../../beatheritage_b200/csrc/mel_static_gen.h"

namespace {
constexpr int kTileF = 32, kSpan = (kTileF - 1) * bhmel::kHop + bhmel::kNfft, kPPitch = 516, kScrPitch = 33;
}

extern "C" int bhmel_emu_forward(const float* x, long long B, long long N, long long row_stride,
                                 long long row0, long long n_total, int n_mels, const float* fb_in,
                                 const float* window_in, double f_min, double f_max, int sample_rate,
                                 int pad_reflect, int log_scale, int exact_log1p, float* y) {
  using namespace bhmel;
  std::vector<float> fb = fb_in ? std::vector<float>(fb_in, fb_in + (size_t)kBins * n_mels)
                                : make_mel_fb(n_mels, f_min, f_max, sample_rate);
  std::vector<float> win = window_in ? std::vector<float>(window_in, window_in + kNfft) : make_hann_window();
  std::vector<float> win_half(kNfft);
  for (int i = 0; i < kNfft; ++i) win_half[i] = 0.5f * win[i];
  std::vector<float> tw = make_twiddles();
  PairTables pt = make_pairs(fb.data(), n_mels);
  RoundTables rt = make_rounds(fb.data(), n_mels);
  const bool use_rounds = (exact_log1p & 2) != 0;   // bit 1: independent-warp kernel's round tables
  // bit 2: the warp-specialised kernel's generated mel stage of a baked filterbank: the direct form, or
  // (with bit 3) P0's hybrid form -- generated code for filters < kStaticP0Filters, pair tables for the rest
  const bool use_static = (exact_log1p & 4) != 0;
  const bool p0_hybrid = (exact_log1p & 8) != 0;    // bit 3: P0 takes its hybrid form (BHMEL_OPT_STATIC_MEL = 2) instead of the direct one
  exact_log1p &= 1;
  PairTables pt_rem;
  int baked_id = 0;
  if (use_static) {
    for (int t = 0; t < kNumBakedFbs && !baked_id; ++t) {   // same bit-for-bit match as the library's match_baked_fb
      const BakedFb& b = kBakedFbs[t];
      if (b.n_mels != n_mels) continue;
      std::vector<uint32_t> want((size_t)kBins * n_mels, 0u);
      int pos = 0;
      for (int m = 0; m < n_mels; ++m)
        for (int j = 0; j < b.count[m]; ++j) want[(size_t)(b.start[m] + j) * n_mels + m] = b.bits[pos++];
      const uint32_t* got = reinterpret_cast<const uint32_t*>(fb.data());
      bool same = true;
      for (size_t i = 0; i < want.size() && same; ++i) same = want[i] == got[i] || ((want[i] | got[i]) & 0x7fffffffu) == 0;
      if (same) baked_id = b.id;
    }
    if (!baked_id) return 2;   // not a baked table
    if (baked_id == 1 && !p0_hybrid) baked_id = kStaticP0Direct;
    if (baked_id == 1) pt_rem = make_pairs(fb.data(), n_mels, kStaticP0Filters);
  }

  const long long T = N / kHop + 1;
  const int tiles_per_row = (int)((T + kTileF - 1) / kTileF);
  std::vector<float> span(kSpan), P((size_t)kTileF * kPPitch, 0.f);
  std::vector<float> scr_r(32 * kScrPitch), scr_i(32 * kScrPitch);

  for (long long r = 0; r < B; ++r) {
    for (int tb = 0; tb < tiles_per_row; ++tb) {
      // ---- stage 1 (same mapping as stage_span's cp.async path)
      const long long s0 = (long long)tb * (kTileF * kHop) - kNfft / 2;
      const long long row_off = row0 + r * row_stride;
      long long valid = n_total - row_off;
      valid = valid < 0 ? 0 : (valid > N ? N : valid);
      const float* row = x + row_off;
      for (int e = 0; e < kSpan; ++e) {
        long long i = s0 + e;
        if (i < 0) i = pad_reflect ? -i : -1;
        else if (i >= N) i = pad_reflect ? 2 * (N - 1) - i : -1;
        const bool ok = (i >= 0) && (i < valid);
        span[e] = ok ? row[i] : 0.f;
      }
      // ---- stage 2
      for (int j = 0; j < kTileF / 2; ++j) {
        float ar[32][32], ai[32][32];     // [lane][slot]
        for (int lane = 0; lane < 32; ++lane) {
          float v[36], w[32];
          for (int m = 0; m < 36; ++m) v[m] = span[(2 * j) * kHop + lane + 32 * m];
          for (int m = 0; m < 32; ++m) w[m] = win_half[lane + 32 * m];
          fft32_pass_a(v, w, ar[lane], ai[lane]);
          for (int k = 0; k < 32; ++k) {
            scr_r[k * kScrPitch + lane] = ar[lane][k];
            scr_i[k * kScrPitch + lane] = ai[lane][k];
          }
        }
        float br[32][32], bi[32][32];
        for (int lane = 0; lane < 32; ++lane) {
          float ur[32], ui[32], tr[32], ti[32];
          for (int n = 0; n < 32; ++n) {
            ur[n] = scr_r[lane * kScrPitch + n];
            ui[n] = scr_i[lane * kScrPitch + n];
            tr[n] = tw[(n * 32 + lane) * 2 + 0];
            ti[n] = tw[(n * 32 + lane) * 2 + 1];
          }
          fft32_pass_b(ur, ui, tr, ti, br[lane], bi[lane]);
        }
        for (int lane = 0; lane < 32; ++lane) {
          const int src = (32 - lane) & 31;
          float* Pa = P.data() + (size_t)(2 * j) * kPPitch + lane;
          float* Pb = Pa + kPPitch;
          for (int k2 = 0; k2 < 16; ++k2) {
            const int s = 31 - k2;
            float pr = br[src][s], pi = bi[src][s];      // __shfl_sync(.., src)
            if (lane == 0) { pr = br[0][(s + 1) & 31]; pi = bi[0][(s + 1) & 31]; }
            const float a1 = br[lane][k2] + pr, a2 = bi[lane][k2] - pi;
            const float b1 = bi[lane][k2] + pi, b2 = pr - br[lane][k2];
            Pa[32 * k2] = fmaf(a1, a1, a2 * a2);
            Pb[32 * k2] = fmaf(b1, b1, b2 * b2);
          }
          if (lane == 0) {
            const float zr = 2.f * br[0][16], zi = 2.f * bi[0][16];
            Pa[512] = zr * zr;
            Pb[512] = zi * zi;
          }
        }
      }
      // ---- stage 3
      const int t0 = tb * kTileF;
      const int nf = (int)((T - t0) < kTileF ? (T - t0) : kTileF);
      for (int f = 0; f < nf; ++f) {
        const float* prow = P.data() + (size_t)f * kPPitch;
        float* yrow = y + ((r * T + t0 + f) * (long long)n_mels);
        if (use_static && baked_id >= 2) {   // direct forms: one generated block per mel warp, stores to the frame row
          std::vector<float> tmp(n_mels);
          float srow[kStagePitch];   // this frame's row of the warp's staging block
          const float4* pr = reinterpret_cast<const float4*>(prow);
          auto run_set = [&](auto tag) {
            constexpr int kS = decltype(tag)::value;
            constexpr int kParts = mel_direct_parts<kS>();
            for (int mw = 0; mw < 8; ++mw)
              for (int part = 0; part < kParts; ++part) {
                mel_direct<kS>(pr, srow, mw, part);
                int m0, ncols;
                mel_direct_run<kS>(mw * kParts + part, m0, ncols);
                for (int j = 0; j < ncols; ++j) tmp[m0 + j] = srow[j];   // mel_flush
              }
          };
          if (baked_id == 2) run_set(std::integral_constant<int, 2>{});
          if (baked_id == 3) run_set(std::integral_constant<int, 3>{});
          if (baked_id == 4) run_set(std::integral_constant<int, 4>{});
          if (baked_id == 5) run_set(std::integral_constant<int, 5>{});
          for (int m = 0; m < n_mels; ++m) {
            float v = tmp[m];
            if (log_scale) v = exact_log1p ? log1pf(v) : logf(1.0f + v);
            yrow[m] = v;
          }
        } else if (use_static) {
          float orow[96];
          for (int mw = 0; mw < kStaticP0Warps; ++mw)
            mel_static_P0<false>(reinterpret_cast<const float4*>(prow), orow, mw);
          for (int m = 0; m < kStaticP0Filters; ++m) {
            float v = orow[m];
            if (log_scale) v = exact_log1p ? log1pf(v) : logf(1.0f + v);
            yrow[m] = v;
          }
          for (const PairDesc& d : pt_rem.pairs) {
            const float* pa = prow + 4 * (d.g0 & 0xFFFF);
            const float* pb = prow + 4 * ((unsigned)d.g0 >> 16);
            const float* w = pt_rem.weights.data() + d.woff;
            float a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
            for (int g = 0; g < d.ng; ++g)
              for (int e = 0; e < 4; ++e) {
                a[e] = fmaf(pa[4 * g + e], w[8 * g + e], a[e]);
                b[e] = fmaf(pb[4 * g + e], w[8 * g + 4 + e], b[e]);
              }
            float va = (a[0] + a[1]) + (a[2] + a[3]), vb = (b[0] + b[1]) + (b[2] + b[3]);
            if (log_scale) {
              va = exact_log1p ? log1pf(va) : logf(1.0f + va);
              vb = exact_log1p ? log1pf(vb) : logf(1.0f + vb);
            }
            const int ca = d.mcol & 0xFFFF, cb = (unsigned)d.mcol >> 16;
            yrow[ca] = va;
            if (cb != 0xFFFF) yrow[cb] = vb;
          }
        } else if (use_rounds) {
          for (size_t r = 0; r < rt.rounds.size(); ++r)
            for (int li = 0; li < 16; ++li) {
              const int m = (int)r * 16 + li;
              const float* pp = prow + 4 * rt.g0[r * 16 + li];
              const float* wp = rt.weights.data() + 4 * ((size_t)rt.rounds[r].woff4 + li);
              float a[4] = {0, 0, 0, 0};
              for (int g = 0; g < rt.rounds[r].ng; ++g)
                for (int e = 0; e < 4; ++e) a[e] = fmaf(pp[4 * g + e], wp[64 * g + e], a[e]);
              float v = (a[0] + a[1]) + (a[2] + a[3]);
              if (log_scale) v = exact_log1p ? log1pf(v) : logf(1.0f + v);
              if (m < n_mels) yrow[m] = v;
            }
        } else
        for (int mc = 0, c = 0; mc < n_mels; mc += kMelChunk, ++c) {
          for (int q = pt.chunk_start[c]; q < pt.chunk_start[c + 1]; ++q) {
            const PairDesc d = pt.pairs[q];
            const float* pa = prow + 4 * (d.g0 & 0xFFFF);
            const float* pb = prow + 4 * ((unsigned)d.g0 >> 16);
            const float* w = pt.weights.data() + d.woff;
            float a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
            for (int g = 0; g < d.ng; ++g)
              for (int e = 0; e < 4; ++e) {
                a[e] = fmaf(pa[4 * g + e], w[8 * g + e], a[e]);
                b[e] = fmaf(pb[4 * g + e], w[8 * g + 4 + e], b[e]);
              }
            float va = (a[0] + a[1]) + (a[2] + a[3]), vb = (b[0] + b[1]) + (b[2] + b[3]);
            if (log_scale) {
              va = exact_log1p ? log1pf(va) : logf(1.0f + va);
              vb = exact_log1p ? log1pf(vb) : logf(1.0f + vb);
            }
            const int ca = d.mcol & 0xFFFF, cb = (unsigned)d.mcol >> 16;
            yrow[mc + ca] = va;
            if (cb != 0xFFFF) yrow[mc + cb] = vb;
          }
        }
      }
    }
  }
  return 0;
}

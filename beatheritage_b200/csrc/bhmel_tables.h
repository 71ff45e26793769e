// Host-side constant-table builders for libbhmel (plain C++, also used by the CPU lane emulator).
//
// What each table restates in the reference stack (ref = BeatHeritage repo root):
//   window     torch.hann_window(n_fft, periodic=True), the `window` buffer torchaudio's
//              Spectrogram registers (ref: osuT5/osuT5/model/spectrogram.py:40-49)
//   filterbank torchaudio.functional.melscale_fbanks(n_freqs, f_min, f_max, n_mels, sr,
//              norm=None, mel_scale="htk") -> the `fb` buffer of MelScale
//   twiddles   W_1024^(a*b), a,b in [0,32): the inter-pass factors of the 32x32 split
//   bands      per-filter contiguous non-zero range of fb, padded to groups of 4 bins so the
//              kernel can use 128-bit shared-memory loads
#pragma once
#include <cmath>
#include <utility>
#include <cstdint>
#include <vector>

namespace bhmel {

constexpr int kNfft = 1024;
constexpr int kHop = 128;
constexpr int kBins = kNfft / 2 + 1;   // 513

inline std::vector<float> make_hann_window() {
  std::vector<float> w(kNfft);
  for (int n = 0; n < kNfft; ++n)
    w[n] = static_cast<float>(0.5 - 0.5 * std::cos(2.0 * M_PI * n / kNfft));
  return w;
}

// [32][32] complex (cos, -sin) of 2*pi*a*b/1024, laid out [a][b] (symmetric).
inline std::vector<float> make_twiddles() {
  std::vector<float> t(32 * 32 * 2);
  for (int a = 0; a < 32; ++a)
    for (int b = 0; b < 32; ++b) {
      const double th = 2.0 * M_PI * static_cast<double>(a * b) / kNfft;
      t[(a * 32 + b) * 2 + 0] = static_cast<float>(std::cos(th));
      t[(a * 32 + b) * 2 + 1] = static_cast<float>(-std::sin(th));
    }
  return t;
}

// htk mel filterbank, norm=None, evaluated in double and rounded once (torchaudio evaluates in
// fp32; the two agree to ~1e-5 absolute, far inside the 1e-3 parity bar -- callers that need the
// reference's exact buffer pass it through bhmel_params.fb / bhmel_set_fb).
inline std::vector<float> make_mel_fb(int n_mels, double f_min, double f_max, int sample_rate) {
  auto hz2mel = [](double f) { return 2595.0 * std::log10(1.0 + f / 700.0); };
  auto mel2hz = [](double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); };
  const double m_min = hz2mel(f_min), m_max = hz2mel(f_max);
  std::vector<double> f_pts(n_mels + 2);
  for (int i = 0; i < n_mels + 2; ++i)
    f_pts[i] = mel2hz(m_min + (m_max - m_min) * i / (n_mels + 1));
  std::vector<float> fb(static_cast<size_t>(kBins) * n_mels, 0.f);
  const double nyq = sample_rate / 2;   // integer division like `sample_rate // 2`
  for (int k = 0; k < kBins; ++k) {
    const double f = nyq * k / (kBins - 1);
    for (int m = 0; m < n_mels; ++m) {
      const double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
      const double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
      const double v = std::fmax(0.0, std::fmin(down, up));
      fb[static_cast<size_t>(k) * n_mels + m] = static_cast<float>(v);
    }
  }
  return fb;
}

// Per-filter band descriptor consumed by the kernel's mel stage.
struct FilterBand {
  int32_t g0;      // first 4-bin group (bin index / 4)
  int32_t ng;      // number of 4-bin groups (0 for an all-zero filter)
  int32_t woff;    // offset (in floats, multiple of 4) of this filter's weights
  int32_t pad;
};

struct BandTables {
  std::vector<FilterBand> bands;   // [n_mels]
  std::vector<float> weights;      // 4-aligned groups, zero padded
  int max_groups = 0;
};

// fb: [kBins][n_mels] row-major.  Works for ANY matrix (a dense filter simply gets one long band).
inline BandTables make_bands(const float* fb, int n_mels) {
  BandTables t;
  t.bands.resize(n_mels);
  for (int m = 0; m < n_mels; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < kBins; ++k)
      if (fb[static_cast<size_t>(k) * n_mels + m] != 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    FilterBand b{0, 0, static_cast<int32_t>(t.weights.size()), 0};
    if (first >= 0) {
      b.g0 = first / 4;
      b.ng = last / 4 - b.g0 + 1;
      for (int k = b.g0 * 4; k < (b.g0 + b.ng) * 4; ++k)
        t.weights.push_back(k < kBins ? fb[static_cast<size_t>(k) * n_mels + m] : 0.f);
    }
    if (b.ng > t.max_groups) t.max_groups = b.ng;
    t.bands[m] = b;
  }
  if (t.weights.empty()) t.weights.assign(4, 0.f);
  return t;
}

// ---- paired band tables ------------------------------------------------------------------
// The kernel's mel stage handles two filters per step (two independent FMA chains per lane).
// Filters are taken chunk by chunk (kMelChunk output columns at a time, the size of the epilogue
// staging buffer), sorted by band length inside the chunk and paired with their neighbour; both
// members of a pair are padded with zero weights to the longer band, so one statically unrolled
// loop serves both.  Weights of a pair are interleaved group by group: A0 B0 A1 B1 ...
constexpr int kMelChunk = 96;

struct PairDesc {
  int32_t g0;      // first 4-bin group of A (low 16 bits) and of B (high 16 bits)
  int32_t woff;    // offset (floats) of the pair's interleaved weights
  int32_t ng;      // groups per member (after padding to the longer one)
  int32_t mcol;    // output column inside the chunk: A (low 16 bits), B (high 16 bits; 0xFFFF = none)
};

struct PairTables {
  std::vector<PairDesc> pairs;        // all chunks, concatenated
  std::vector<int32_t> chunk_start;   // [n_chunks + 1] index into pairs
  std::vector<float> weights;
  int max_groups = 0;
};

// first_filter > 0 builds the tables for filters [first_filter, n_mels) only (the remainder the
// generic stage handles next to a statically scheduled mel stage); output columns stay absolute.
inline PairTables make_pairs(const float* fb, int n_mels, int first_filter = 0) {
  const BandTables bt = make_bands(fb, n_mels);
  constexpr int kGroups = (kBins + 3) / 4;   // 129 groups cover bins 0..515
  PairTables t;
  t.chunk_start.push_back(0);
  for (int mc = 0; mc < n_mels; mc += kMelChunk) {
    const int cnt = (n_mels - mc) < kMelChunk ? (n_mels - mc) : kMelChunk;
    std::vector<int> order;
    for (int i = 0; i < cnt; ++i)
      if (mc + i >= first_filter) order.push_back(mc + i);
    // longest bands first (stable -> deterministic)
    const int cnt_sel = static_cast<int>(order.size());
    for (int i = 1; i < cnt_sel; ++i)
      for (int j = i; j > 0 && bt.bands[order[j]].ng > bt.bands[order[j - 1]].ng; --j) std::swap(order[j], order[j - 1]);
    for (int i = 0; i < cnt_sel; i += 2) {
      const int ma = order[i], mb = (i + 1 < cnt_sel) ? order[i + 1] : -1;
      const FilterBand a = bt.bands[ma];
      const FilterBand b = mb >= 0 ? bt.bands[mb] : FilterBand{0, 0, 0, 0};
      const int ng = a.ng > b.ng ? a.ng : b.ng;
      auto place = [&](const FilterBand& f) { return (f.g0 + ng > kGroups) ? kGroups - ng : f.g0; };
      const int ga = place(a), gb = place(b);
      PairDesc d;
      d.g0 = ga | (gb << 16);
      d.woff = static_cast<int32_t>(t.weights.size());
      d.ng = ng;
      d.mcol = (ma - mc) | ((mb >= 0 ? (mb - mc) : 0xFFFF) << 16);
      for (int g = 0; g < ng; ++g) {
        for (int which = 0; which < 2; ++which) {
          const FilterBand& f = which == 0 ? a : b;
          const int gbase = which == 0 ? ga : gb;
          const int grp = gbase + g;                 // absolute group index of this slot
          for (int e = 0; e < 4; ++e) {
            float w = 0.f;
            if (grp >= f.g0 && grp < f.g0 + f.ng) w = bt.weights[f.woff + (grp - f.g0) * 4 + e];
            t.weights.push_back(w);
          }
        }
      }
      if (ng > t.max_groups) t.max_groups = ng;
      t.pairs.push_back(d);
    }
    t.chunk_start.push_back(static_cast<int32_t>(t.pairs.size()));
  }
  if (t.weights.empty()) t.weights.assign(8, 0.f);
  return t;
}

// ---- round tables (independent-warp kernel) ------------------------------------------------
// In the independent-warp kernel one warp projects its own frame PAIR: lanes 0-15 take frame A,
// lanes 16-31 frame B, and lane i (mod 16) owns filter 16*r + i in round r.  All 16 filters of a
// round are zero padded to the longest band of the round, so the dot-product loop is warp
// uniform; weights are stored lane-interleaved ([round][group][lane 0..15] float4) so the 128-bit
// weight loads are conflict free, and each filter's first group is shifted (inside the slack the
// padding gives) so that the 8 lanes of a shared-memory phase hit different bank quads of P.
struct RoundDesc {
  int32_t woff4;   // offset of the round's weights in float4 units
  int32_t ng;      // groups per filter in this round
};

struct RoundTables {
  std::vector<RoundDesc> rounds;      // [ceil(n_mels / 16)]
  std::vector<int32_t> g0;            // [rounds * 16] first 4-bin group per (round, lane)
  std::vector<float> weights;         // float4-granular, [round][group][16 lanes][4]
};

inline RoundTables make_rounds(const float* fb, int n_mels) {
  const BandTables bt = make_bands(fb, n_mels);
  constexpr int kGroups = (kBins + 3) / 4;   // 129
  RoundTables t;
  const int n_rounds = (n_mels + 15) / 16;
  t.g0.assign(static_cast<size_t>(n_rounds) * 16, 0);
  for (int r = 0; r < n_rounds; ++r) {
    int ng = 0;
    for (int i = 0; i < 16; ++i) {
      const int m = 16 * r + i;
      if (m < n_mels && bt.bands[m].ng > ng) ng = bt.bands[m].ng;
    }
    RoundDesc rd{static_cast<int32_t>(t.weights.size() / 4), ng};
    // choose the (shifted) first group of every filter; tightest intervals first
    int lo[16], hi[16], order[16], chosen[16];
    for (int i = 0; i < 16; ++i) {
      const int m = 16 * r + i;
      const FilterBand f = (m < n_mels) ? bt.bands[m] : FilterBand{0, 0, 0, 0};
      lo[i] = f.g0 + f.ng - ng;
      if (lo[i] < 0) lo[i] = 0;
      hi[i] = f.g0;
      if (hi[i] > kGroups - ng) hi[i] = kGroups - ng;
      if (hi[i] < lo[i]) hi[i] = lo[i];
      order[i] = i;
    }
    for (int a = 1; a < 16; ++a)
      for (int b = a; b > 0 && (hi[order[b]] - lo[order[b]]) < (hi[order[b - 1]] - lo[order[b - 1]]); --b)
        std::swap(order[b], order[b - 1]);
    bool used[2][8] = {{false}};
    for (int a = 0; a < 16; ++a) {
      const int i = order[a], grp = i >> 3;
      int pick = hi[i];
      for (int c = hi[i]; c >= lo[i]; --c)
        if (!used[grp][c & 7]) { pick = c; break; }
      used[grp][pick & 7] = true;
      chosen[i] = pick;
      t.g0[static_cast<size_t>(r) * 16 + i] = pick;
    }
    for (int g = 0; g < ng; ++g)
      for (int i = 0; i < 16; ++i) {
        const int m = 16 * r + i;
        const FilterBand f = (m < n_mels) ? bt.bands[m] : FilterBand{0, 0, 0, 0};
        const int grp = chosen[i] + g;
        for (int e = 0; e < 4; ++e) {
          float w = 0.f;
          if (grp >= f.g0 && grp < f.g0 + f.ng) w = bt.weights[f.woff + (grp - f.g0) * 4 + e];
          t.weights.push_back(w);
        }
      }
    t.rounds.push_back(rd);
  }
  if (t.weights.empty()) t.weights.assign(64, 0.f);
  return t;
}

}  // namespace bhmel

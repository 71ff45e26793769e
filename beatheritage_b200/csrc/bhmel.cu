// libbhmel.so -- C ABI implementation (see include/bhmel.h for the contract and the reference
// file:line each entry point replaces).
#include "../../include/bhmel.h"

#include <cuda_runtime.h>

#include <atomic>
#include <climits>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "bhmel_fb_baked.h"
#include "bhmel_kernel_iw.cuh"
#include "bhmel_kernel_ws.cuh"


namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define BH_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return fail(BHMEL_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));         \
  } while (0)

#ifndef BHMEL_HOST_SLOTS
#define BHMEL_HOST_SLOTS 3
#endif
#ifndef BHMEL_HOST_CHUNK_MB
#define BHMEL_HOST_CHUNK_MB 16
#endif
#ifndef BHMEL_HOST_CHUNK_MB_PCM
#define BHMEL_HOST_CHUNK_MB_PCM 32
#endif
#ifndef BHMEL_HOST_TAPER
#define BHMEL_HOST_TAPER 1
#endif
constexpr int kHostSlots = BHMEL_HOST_SLOTS;

}  // namespace

// int16 PCM -> float32 with a per-row multiplier: float32(pcm) * scale, one rounding, exactly what
// numpy does for `samples.astype(np.float32) * np.float32(scale)` (ref: data_utils.py:95-97).
__global__ void bhmel_pcm16_to_f32_kernel(const int16_t* __restrict__ in, float* __restrict__ out,
                                          const float* __restrict__ scales, long long n_per_row, long long total) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / n_per_row;
    out[i] = static_cast<float>(in[i]) * (scales ? scales[r] : 1.0f);
  }
}

// max |pcm| over a device-resident int16 song (grid-stride, 8 samples per 128-bit load where the
// pointer allows), then scale = 1.0f / max in float32 -- ref: data_utils.py:94-96.
__global__ void bhmel_absmax_pcm16_kernel(const int16_t* __restrict__ in, long long n, unsigned* __restrict__ out_max) {
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  unsigned m = 0;
  auto upd = [&m](int v) {
    const unsigned a = static_cast<unsigned>(v < 0 ? -v : v);   // |-32768| = 32768 fits
    m = a > m ? a : m;
  };
  const uintptr_t addr = reinterpret_cast<uintptr_t>(in);
  long long head = ((16 - (addr & 15)) & 15) / 2;               // samples before the first 16-byte boundary
  if (head > n) head = n;
  for (long long i = tid; i < head; i += nthreads) upd(in[i]);
  const long long nvec = (n - head) / 8;
  const int4* v = reinterpret_cast<const int4*>(in + head);
  for (long long i = tid; i < nvec; i += nthreads) {
    const int4 q = v[i];
    const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      upd(static_cast<int16_t>(w[k] & 0xFFFF));
      upd(static_cast<int16_t>(static_cast<unsigned>(w[k]) >> 16));
    }
  }
  for (long long i = head + nvec * 8 + tid; i < n; i += nthreads) upd(in[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned other = __shfl_xor_sync(0xffffffffu, m, o);
    m = other > m ? other : m;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out_max, m);
}

// In place: the 4 bytes at `scale` hold max|pcm| as an unsigned integer on entry and 1.0f / max on exit.
__global__ void bhmel_peak_to_scale_kernel(float* __restrict__ scale) {
  const unsigned m = *reinterpret_cast<const unsigned*>(scale);
  *scale = 1.0f / static_cast<float>(m);   // 1/0 -> +inf, like the reference's division by zero
}

// ---- encoder-input assembly helpers (bhmel_forward_encoder_input); T = float or __nv_bfloat16 ----
// BTC: y[(b*Tn + t)*C + n_mels + c] = cond[b*n_cond + c].  One thread per 16-byte chunk of a frame's
// conditioning slice when everything is 16-byte aligned (the reference shapes: 80 + 384 channels),
// else one thread per element.
template <typename T, bool kVec>
__global__ void bhmel_cond_fill_btc_kernel(T* __restrict__ y, const T* __restrict__ cond, long long rows /* B*Tn */,
                                           long long Tn, int C, int n_mels, int n_cond) {
  constexpr int kPer = kVec ? 16 / static_cast<int>(sizeof(T)) : 1;
  const int per_row = n_cond / kPer;
  const long long total = rows * per_row;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long row = i / per_row;
    const int k = static_cast<int>(i - row * per_row);
    const long long b = row / Tn;
    if constexpr (kVec) {
      const int4 v = *reinterpret_cast<const int4*>(cond + b * n_cond + k * kPer);
      *reinterpret_cast<int4*>(y + row * C + n_mels + k * kPer) = v;
    } else {
      y[row * C + n_mels + k] = cond[b * n_cond + k];
    }
  }
}

// BCT: y[b][c][t] = mel[b][t][c] for c < n_mels (32 x 32 tile transpose through shared memory),
// cond[b][c - n_mels] for the rest.  Grid: (tiles of 32 frames, tiles of 32 channels, B).
template <typename T>
__global__ void bhmel_assemble_bct_kernel(T* __restrict__ y, const T* __restrict__ mel, const T* __restrict__ cond,
                                          long long Tn, int C, int n_mels, int n_cond) {
  __shared__ T tile[32][33];
  const long long b = blockIdx.z;
  const long long t0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  T* yb = y + b * C * Tn;
  if (c0 < n_mels) {   // a tile that holds mel channels (n_mels need not be a multiple of 32)
    const T* mb = mel + b * Tn * n_mels;
    for (int j = ty; j < 32; j += 8) {
      const long long t = t0 + j;
      const int c = c0 + tx;
      if (t < Tn && c < n_mels) tile[j][tx] = mb[t * n_mels + c];
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int c = c0 + j;
      const long long t = t0 + tx;
      if (t < Tn && c < C) yb[static_cast<long long>(c) * Tn + t] = c < n_mels ? tile[tx][j] : cond[b * n_cond + (c - n_mels)];
    }
  } else {
    for (int j = ty; j < 32; j += 8) {
      const int c = c0 + j;
      const long long t = t0 + tx;
      if (t < Tn && c < C) yb[static_cast<long long>(c) * Tn + t] = cond[b * n_cond + (c - n_mels)];
    }
  }
}

struct bhmel_handle {
  int device = 0;
  int num_sms = 0;
  bhmel_params prm{};
  std::vector<float> fb;       // host copy [513][n_mels]
  std::vector<float> window;   // host copy [1024]
  float* d_win = nullptr;      // 0.5 * window
  float2* d_tw = nullptr;
  bhmel::PairDesc* d_pairs = nullptr;
  float* d_weights = nullptr;
  int n_weights = 0;
  int n_pairs = 0;
  bhmel::PairDesc* d_pairs_rem = nullptr;   // baked filterbank: generic tables of the filters the static stage leaves
  float* d_weights_rem = nullptr;
  int n_weights_rem = 0;
  int n_pairs_rem = 0;
  bhmel::RoundDesc* d_rounds = nullptr;   // independent-warp kernel tables
  int* d_g0 = nullptr;
  float* d_wt = nullptr;
  int n_rounds = 0;
  int n_wt4 = 0;
  std::atomic<int64_t> launches{0};
  int use_bulk = 1;
  int kernel_variant = BHMEL_KERNEL_WARP_SPECIALIZED;
  int baked_fb = 0;      // id of the baked table (bhmel_fb_baked.h) the filterbank equals bit for bit: 1 = P0, 2.. = all-static sets; 0 = none
  int pdl = 1;           // BHMEL_OPT_PDL: launch the warp-specialised kernel with programmatic stream serialisation
  int static_mel = 1;    // BHMEL_OPT_STATIC_MEL: 0 generic stage always, 1 the baked table's direct form, 2 P0 takes its hybrid form
  // bhmel_forward_host pipeline (lazily created)
  std::mutex host_mu;
  cudaStream_t hs[kHostSlots] = {};
  float* d_in[kHostSlots] = {};
  float* d_out[kHostSlots] = {};
  int16_t* d_pcm[kHostSlots] = {};
  float* d_scales = nullptr;
  size_t cap_in = 0, cap_out = 0, cap_pcm = 0, cap_scales = 0;
  // bhmel_forward_encoder_input / bhmel_forward_gather_pcm16 scratch (lazily created, only when the caller passes none)
  void* d_stage = nullptr;     // bhmel_forward_encoder_input, BCT layout: [B][T][n_mels] of the output dtype
  size_t cap_stage = 0;
  float* d_song = nullptr;     // float32 copy of the int16 song the gather kernel reads
  size_t cap_song = 0;
};

namespace {

// Which baked filterbank (bhmel_fb_baked.h) the host table equals bit for bit: its id, or 0 for none.
int match_baked_fb(const std::vector<float>& fb, int n_mels) {
  for (int t = 0; t < kNumBakedFbs; ++t) {
    const BakedFb& b = kBakedFbs[t];
    if (b.n_mels != n_mels) continue;
    std::vector<uint32_t> want(static_cast<size_t>(bhmel::kBins) * n_mels, 0u);
    int pos = 0;
    for (int m = 0; m < n_mels; ++m)
      for (int j = 0; j < b.count[m]; ++j) want[static_cast<size_t>(b.start[m] + j) * n_mels + m] = b.bits[pos++];
    // value equality with -0 == +0: torchaudio's P1 / T5 tables carry a negative zero at [0][0]
    // (max(0, min(down, up)) of a -0 slope); a zero weight of either sign contributes nothing
    const uint32_t* got = reinterpret_cast<const uint32_t*>(fb.data());
    bool same = true;
    for (size_t i = 0; i < want.size() && same; ++i) same = want[i] == got[i] || ((want[i] | got[i]) & 0x7fffffffu) == 0;
    if (same) return b.id;
  }
  return 0;
}

int upload_filterbank(bhmel_handle* h) {
  h->baked_fb = match_baked_fb(h->fb, h->prm.n_mels);
  bhmel::PairTables t = bhmel::make_pairs(h->fb.data(), h->prm.n_mels);
  if (t.pairs.size() > static_cast<size_t>(bhmel::kPairCap))
    return fail(BHMEL_EINVAL, "too many filter pairs for the kernel's descriptor table");
  // The tables may be in use by kernels in flight on any stream of this device.
  BH_CUDA(cudaDeviceSynchronize());
  if (h->d_pairs) cudaFree(h->d_pairs);
  if (h->d_weights) cudaFree(h->d_weights);
  h->d_pairs = nullptr;
  h->d_weights = nullptr;
  BH_CUDA(cudaMalloc(&h->d_pairs, t.pairs.size() * sizeof(bhmel::PairDesc)));
  BH_CUDA(cudaMalloc(&h->d_weights, t.weights.size() * sizeof(float)));
  BH_CUDA(cudaMemcpy(h->d_pairs, t.pairs.data(), t.pairs.size() * sizeof(bhmel::PairDesc),
                     cudaMemcpyHostToDevice));
  BH_CUDA(cudaMemcpy(h->d_weights, t.weights.data(), t.weights.size() * sizeof(float),
                     cudaMemcpyHostToDevice));
  h->n_weights = static_cast<int>(t.weights.size());
  h->n_pairs = static_cast<int>(t.pairs.size());
  if (h->d_pairs_rem) cudaFree(h->d_pairs_rem);
  if (h->d_weights_rem) cudaFree(h->d_weights_rem);
  h->d_pairs_rem = nullptr;
  h->d_weights_rem = nullptr;
  if (h->baked_fb == 1) {
    bhmel::PairTables r = bhmel::make_pairs(h->fb.data(), h->prm.n_mels, bhmel::kStaticP0Filters);
    if (r.weights.size() > static_cast<size_t>(bhmel::kFwCap) || r.pairs.size() > static_cast<size_t>(bhmel::kPairCap)) {
      h->baked_fb = 0;   // cannot happen for the baked table; keep the generic stage if it ever does
    } else {
      BH_CUDA(cudaMalloc(&h->d_pairs_rem, r.pairs.size() * sizeof(bhmel::PairDesc)));
      BH_CUDA(cudaMalloc(&h->d_weights_rem, r.weights.size() * sizeof(float)));
      BH_CUDA(cudaMemcpy(h->d_pairs_rem, r.pairs.data(), r.pairs.size() * sizeof(bhmel::PairDesc), cudaMemcpyHostToDevice));
      BH_CUDA(cudaMemcpy(h->d_weights_rem, r.weights.data(), r.weights.size() * sizeof(float), cudaMemcpyHostToDevice));
      h->n_weights_rem = static_cast<int>(r.weights.size());
      h->n_pairs_rem = static_cast<int>(r.pairs.size());
    }
  }

  bhmel::RoundTables rt = bhmel::make_rounds(h->fb.data(), h->prm.n_mels);
  if (rt.rounds.size() > static_cast<size_t>(bhmel::iw::kRoundCap))
    return fail(BHMEL_EINVAL, "too many filter rounds for the kernel's descriptor table");
  if (h->d_rounds) cudaFree(h->d_rounds);
  if (h->d_g0) cudaFree(h->d_g0);
  if (h->d_wt) cudaFree(h->d_wt);
  h->d_rounds = nullptr;
  h->d_g0 = nullptr;
  h->d_wt = nullptr;
  BH_CUDA(cudaMalloc(&h->d_rounds, rt.rounds.size() * sizeof(bhmel::RoundDesc)));
  BH_CUDA(cudaMalloc(&h->d_g0, rt.g0.size() * sizeof(int)));
  BH_CUDA(cudaMalloc(&h->d_wt, rt.weights.size() * sizeof(float)));
  BH_CUDA(cudaMemcpy(h->d_rounds, rt.rounds.data(), rt.rounds.size() * sizeof(bhmel::RoundDesc),
                     cudaMemcpyHostToDevice));
  BH_CUDA(cudaMemcpy(h->d_g0, rt.g0.data(), rt.g0.size() * sizeof(int), cudaMemcpyHostToDevice));
  BH_CUDA(cudaMemcpy(h->d_wt, rt.weights.data(), rt.weights.size() * sizeof(float), cudaMemcpyHostToDevice));
  h->n_rounds = static_cast<int>(rt.rounds.size());
  h->n_wt4 = static_cast<int>(rt.weights.size() / 4);
  return BHMEL_OK;
}

int upload_window(bhmel_handle* h) {
  std::vector<float> half(bhmel::kNfft);
  for (int i = 0; i < bhmel::kNfft; ++i) half[i] = 0.5f * h->window[i];   // exact scaling
  BH_CUDA(cudaDeviceSynchronize());
  if (!h->d_win) BH_CUDA(cudaMalloc(&h->d_win, bhmel::kNfft * sizeof(float)));
  BH_CUDA(cudaMemcpy(h->d_win, half.data(), bhmel::kNfft * sizeof(float), cudaMemcpyHostToDevice));
  return BHMEL_OK;
}

int check_device(const bhmel_handle* h) {
  int dev = -1;
  BH_CUDA(cudaGetDevice(&dev));
  if (dev != h->device)
    return fail(BHMEL_EDEVICE, "handle was created on device " + std::to_string(h->device) +
                                   " but the current device is " + std::to_string(dev));
  return BHMEL_OK;
}

struct OutSpec {
  void* y;
  int bf16;
  long long frame_pitch;   // 0 -> n_mels
  long long row_pitch;     // 0 -> T * frame_pitch
};

int launch(bhmel_handle* h, const float* x, long long row_stride, long long row0, long long n_total,
           long long B, long long N, OutSpec out, cudaStream_t stream) {
  void* y = out.y;
  if (!h) return fail(BHMEL_EINVAL, "null handle");
  if (!x || !y) return fail(BHMEL_EINVAL, "null data pointer");
  if (B <= 0 || N <= 0) return fail(BHMEL_ESHAPE, "batch and sample count must be positive");
  if (B > INT_MAX) return fail(BHMEL_ESHAPE, "batch too large");
  if (h->prm.pad_mode == BHMEL_PAD_REFLECT && N <= bhmel::kNfft / 2)
    return fail(BHMEL_ESHAPE,
                "reflect padding needs more than n_fft/2 = 512 samples per row (got " +
                    std::to_string(N) + "); torch.nn.functional.pad raises for the reference too");
  if (int rc = check_device(h)) return rc;

  bhmel::KParams p{};
  p.x = x;
  p.row_stride = row_stride;
  p.row0 = row0;
  p.n_total = n_total;
  p.N = N;
  p.T = N / bhmel::kHop + 1;
  const bool iw = h->kernel_variant == BHMEL_KERNEL_INDEPENDENT_WARPS;
  const bool ws = h->kernel_variant == BHMEL_KERNEL_WARP_SPECIALIZED;
  const int tile_frames = iw ? bhmel::iw::kWTileF : bhmel::kTileF;
  p.tiles_per_row = static_cast<int>((p.T + tile_frames - 1) / tile_frames);
  p.n_tiles = static_cast<long long>(p.tiles_per_row) * B;
  p.y = static_cast<float*>(y);
  p.y_frame_pitch = out.frame_pitch > 0 ? out.frame_pitch : h->prm.n_mels;
  p.y_row_pitch = out.row_pitch > 0 ? out.row_pitch : p.T * p.y_frame_pitch;
  p.y_bf16 = out.bf16;
  p.y_limit = (B - 1) * p.y_row_pitch + (p.T - 1) * p.y_frame_pitch + h->prm.n_mels;
  if (p.y_frame_pitch < h->prm.n_mels || p.y_row_pitch < p.T * p.y_frame_pitch)
    return fail(BHMEL_EINVAL, "output pitches too small for [T][n_mels]");
  p.win_half = h->d_win;
  p.tw = h->d_tw;
  p.pairs = h->d_pairs;
  p.weights = h->d_weights;
  p.B = static_cast<int>(B);
  p.n_mels = h->prm.n_mels;
  p.pad_reflect = h->prm.pad_mode == BHMEL_PAD_REFLECT;
  p.log_scale = h->prm.log_scale != 0;
  p.use_bulk = h->use_bulk;
  p.n_weights = h->n_weights;
  p.n_pairs = h->n_pairs;

  if (iw) {
    bhmel::iw::IwParams q{};
    q.k = p;
    q.rounds = h->d_rounds;
    q.g0 = h->d_g0;
    q.wt = reinterpret_cast<const float4*>(h->d_wt);
    q.n_rounds = h->n_rounds;
    q.n_wt4 = h->n_wt4;
    const long long ctas = (p.n_tiles + bhmel::iw::kIwWarps - 1) / bhmel::iw::kIwWarps;
    const unsigned grid = static_cast<unsigned>(ctas < h->num_sms ? ctas : h->num_sms);
    if (p.log_scale)
      bhmel::iw::bhmel_logmel_iw_kernel<true><<<grid, bhmel::iw::kIwThreads, sizeof(bhmel::iw::SmemIW), stream>>>(q);
    else
      bhmel::iw::bhmel_logmel_iw_kernel<false><<<grid, bhmel::iw::kIwThreads, sizeof(bhmel::iw::SmemIW), stream>>>(q);
  } else if (ws) {
    const unsigned grid = static_cast<unsigned>(p.n_tiles < h->num_sms ? p.n_tiles : h->num_sms);
    // Which mel stage: 0 generic, 1 P0 hybrid, >= 2 a direct form (vector stores from registers: the
    // output rows must be 16-byte (float32) / 8-byte (bfloat16) aligned, else the generic stage runs).
    // BHMEL_OPT_STATIC_MEL: 0 pair tables always; 1 (default) the baked table's direct form; 2 P0 takes its
    // hybrid form instead (bit-identical to the pair tables, the round-1 default)
    int st = h->static_mel ? h->baked_fb : 0;
    if (st == 1 && h->static_mel != 2) st = bhmel::kStaticP0Direct;
    {   // direct forms: 16-byte (f32) / 8-byte (bf16) stores when every output row allows it, else element stores
      const uintptr_t align = out.bf16 ? 8 : 16;
      p.y_vec_ok = (reinterpret_cast<uintptr_t>(y) % align) == 0 && p.y_frame_pitch % 4 == 0 && p.y_row_pitch % 4 == 0;
    }
    constexpr size_t smem = sizeof(bhmel::ws::SmemWS);
    if (st == 1) {   // hybrid: the generic stage only sees the filters the generated code leaves
      p.pairs = h->d_pairs_rem;
      p.weights = h->d_weights_rem;
      p.n_pairs = h->n_pairs_rem;
      p.n_weights = h->n_weights_rem;
    } else if (st >= 2) {   // direct: no tables at all
      p.n_pairs = 0;
      p.n_weights = 0;
    }
    // Programmatic stream serialisation: the kernel's prologue (table staging, barrier init) may overlap the
    // tail of the previous kernel in the stream; its mel / producer role waits (griddepcontrol.wait) before the
    // first sample is read or the first output written.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(bhmel::ws::kThreadsW);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute pdl_attr[1];
    pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = pdl_attr;
    cfg.numAttrs = h->pdl ? 1 : 0;
#define BHMEL_WS_LAUNCH(LOG, ST, BF) BH_CUDA(cudaLaunchKernelEx(&cfg, bhmel::ws::bhmel_logmel_ws_kernel<LOG, ST, BF>, p))
#define BHMEL_WS_DIRECT(ST)                                        \
  case ST:                                                         \
    if (p.log_scale) {                                             \
      if (p.y_bf16) BHMEL_WS_LAUNCH(true, ST, true);               \
      else BHMEL_WS_LAUNCH(true, ST, false);                       \
    } else {                                                       \
      if (p.y_bf16) BHMEL_WS_LAUNCH(false, ST, true);              \
      else BHMEL_WS_LAUNCH(false, ST, false);                      \
    }                                                              \
    break
    switch (st) {
      case 1:
        if (p.log_scale) BHMEL_WS_LAUNCH(true, 1, false);
        else BHMEL_WS_LAUNCH(false, 1, false);
        break;
      BHMEL_WS_DIRECT(2);
      BHMEL_WS_DIRECT(3);
      BHMEL_WS_DIRECT(4);
      BHMEL_WS_DIRECT(5);
      default:
        if (p.log_scale) BHMEL_WS_LAUNCH(true, 0, false);
        else BHMEL_WS_LAUNCH(false, 0, false);
        break;
    }
#undef BHMEL_WS_DIRECT
#undef BHMEL_WS_LAUNCH
  } else {
    const unsigned grid = static_cast<unsigned>(p.n_tiles < h->num_sms ? p.n_tiles : h->num_sms);
    if (p.log_scale)
      bhmel::bhmel_logmel_kernel<true><<<grid, bhmel::kThreads, sizeof(bhmel::SmemLayout), stream>>>(p);
    else
      bhmel::bhmel_logmel_kernel<false><<<grid, bhmel::kThreads, sizeof(bhmel::SmemLayout), stream>>>(p);
  }
  BH_CUDA(cudaGetLastError());
  h->launches.fetch_add(1, std::memory_order_relaxed);
  return BHMEL_OK;
}

template <typename T>
int encoder_input_tail(bhmel_handle* h, const bhmel_encoder_input_desc* d, const T* staged, long long B, long long Tn,
                       cudaStream_t s) {
  const int n_mels = h->prm.n_mels, n_cond = static_cast<int>(d->n_cond), C = n_mels + n_cond;
  T* y = static_cast<T*>(d->y);
  const T* cond = static_cast<const T*>(d->cond);
  if (d->layout == BHMEL_LAYOUT_BTC) {
    if (n_cond == 0) return BHMEL_OK;
    constexpr int kPer = 16 / static_cast<int>(sizeof(T));
    const bool vec = n_cond % kPer == 0 && n_mels % kPer == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(cond) & 15) == 0;
    const long long total = B * Tn * (vec ? n_cond / kPer : n_cond);
    const long long want = (total + 255) / 256;
    const unsigned blocks = static_cast<unsigned>(want < h->num_sms * 16 ? want : h->num_sms * 16);
    if (vec) bhmel_cond_fill_btc_kernel<T, true><<<blocks, 256, 0, s>>>(y, cond, B * Tn, Tn, C, n_mels, n_cond);
    else bhmel_cond_fill_btc_kernel<T, false><<<blocks, 256, 0, s>>>(y, cond, B * Tn, Tn, C, n_mels, n_cond);
  } else {
    const dim3 grid(static_cast<unsigned>((Tn + 31) / 32), static_cast<unsigned>((C + 31) / 32), static_cast<unsigned>(B));
    bhmel_assemble_bct_kernel<T><<<grid, dim3(32, 8), 0, s>>>(y, staged, cond, Tn, C, n_mels, n_cond);
  }
  BH_CUDA(cudaGetLastError());
  h->launches.fetch_add(1, std::memory_order_relaxed);
  return BHMEL_OK;
}

}  // namespace

extern "C" {

int bhmel_version(void) { return BHMEL_VERSION; }

#ifdef BHMEL_TRACE   // debug builds only: fetch the phase timeline recorded by block 0 of the ws kernel
int bhmel_debug_trace(long long* out, int n) {
  const size_t bytes = sizeof(long long) * static_cast<size_t>(n);
  if (bytes > sizeof(bhmel::ws::g_trace)) return BHMEL_EINVAL;
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, bhmel::ws::g_trace, bytes) == cudaSuccess ? BHMEL_OK : BHMEL_ECUDA;
}
#endif

const char* bhmel_last_error(void) { return g_err.c_str(); }

void bhmel_kernel_info(int32_t* smem_bytes, int32_t* threads, int32_t* tile_frames) {
  if (smem_bytes) *smem_bytes = static_cast<int32_t>(sizeof(bhmel::ws::SmemWS));
  if (threads) *threads = bhmel::ws::kThreadsW;
  if (tile_frames) *tile_frames = bhmel::kTileF;
}

int bhmel_create(const bhmel_params* prm, bhmel_handle** out) {
  if (!prm || !out) return fail(BHMEL_EINVAL, "null argument");
  *out = nullptr;
  if (prm->n_fft != bhmel::kNfft || prm->hop_length != bhmel::kHop)
    return fail(BHMEL_EINVAL, "only n_fft=1024 / hop_length=128 are compiled in (constant in every "
                              "reference config); got n_fft=" + std::to_string(prm->n_fft) +
                                  " hop_length=" + std::to_string(prm->hop_length));
  if (prm->n_mels < 1 || prm->n_mels > 1024) return fail(BHMEL_EINVAL, "n_mels must be in [1, 1024]");
  if (prm->pad_mode != BHMEL_PAD_CONSTANT && prm->pad_mode != BHMEL_PAD_REFLECT)
    return fail(BHMEL_EINVAL, "pad_mode must be BHMEL_PAD_CONSTANT or BHMEL_PAD_REFLECT");
  if (prm->sample_rate <= 0) return fail(BHMEL_EINVAL, "sample_rate must be positive");
  if (!prm->fb && !(prm->f_max > prm->f_min && prm->f_min >= 0))
    return fail(BHMEL_EINVAL, "need 0 <= f_min < f_max");

  int dev = 0;
  BH_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop{};
  BH_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(BHMEL_EDEVICE, std::string("libbhmel is built for sm_100a only; device is ") + prop.name +
                                   " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")");
  BH_CUDA(cudaFuncSetAttribute(bhmel::bhmel_logmel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(sizeof(bhmel::SmemLayout))));
  BH_CUDA(cudaFuncSetAttribute(bhmel::bhmel_logmel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(sizeof(bhmel::SmemLayout))));
#define BHMEL_WS_ATTR(ST, BF)                                                                                             \
  BH_CUDA(cudaFuncSetAttribute(bhmel::ws::bhmel_logmel_ws_kernel<true, ST, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                               static_cast<int>(sizeof(bhmel::ws::SmemWS))));                                               \
  BH_CUDA(cudaFuncSetAttribute(bhmel::ws::bhmel_logmel_ws_kernel<false, ST, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                               static_cast<int>(sizeof(bhmel::ws::SmemWS))))
  static_assert(bhmel::kStaticP0Direct == 5, "launch() and bhmel_create() enumerate the direct forms 2..5");
  BHMEL_WS_ATTR(0, false);
  BHMEL_WS_ATTR(1, false);
  BHMEL_WS_ATTR(2, false); BHMEL_WS_ATTR(2, true);
  BHMEL_WS_ATTR(3, false); BHMEL_WS_ATTR(3, true);
  BHMEL_WS_ATTR(4, false); BHMEL_WS_ATTR(4, true);
  BHMEL_WS_ATTR(5, false); BHMEL_WS_ATTR(5, true);
#undef BHMEL_WS_ATTR
  BH_CUDA(cudaFuncSetAttribute(bhmel::iw::bhmel_logmel_iw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(sizeof(bhmel::iw::SmemIW))));
  BH_CUDA(cudaFuncSetAttribute(bhmel::iw::bhmel_logmel_iw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(sizeof(bhmel::iw::SmemIW))));

  bhmel_handle* h = new bhmel_handle();
  h->device = dev;
  h->num_sms = prop.multiProcessorCount;
  h->prm = *prm;
  h->prm.fb = nullptr;
  h->prm.window = nullptr;
  const size_t fb_elems = static_cast<size_t>(bhmel::kBins) * prm->n_mels;
  if (prm->fb) h->fb.assign(prm->fb, prm->fb + fb_elems);
  else h->fb = bhmel::make_mel_fb(prm->n_mels, prm->f_min, prm->f_max, prm->sample_rate);
  if (prm->window) h->window.assign(prm->window, prm->window + bhmel::kNfft);
  else h->window = bhmel::make_hann_window();

  auto cleanup = [&](int rc) {
    bhmel_destroy(h);
    return rc;
  };
  std::vector<float> tw = bhmel::make_twiddles();
  cudaError_t e = cudaMalloc(&h->d_tw, tw.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_tw, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cleanup(fail(BHMEL_ECUDA, std::string("twiddle upload: ") + cudaGetErrorString(e)));
  if (int rc = upload_window(h)) return cleanup(rc);
  if (int rc = upload_filterbank(h)) return cleanup(rc);
  *out = h;
  return BHMEL_OK;
}

void bhmel_destroy(bhmel_handle* h) {
  if (!h) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  for (int i = 0; i < kHostSlots; ++i) {
    if (h->hs[i]) {
      cudaStreamSynchronize(h->hs[i]);
      cudaStreamDestroy(h->hs[i]);
    }
    if (h->d_in[i]) cudaFree(h->d_in[i]);
    if (h->d_out[i]) cudaFree(h->d_out[i]);
    if (h->d_pcm[i]) cudaFree(h->d_pcm[i]);
  }
  if (h->d_scales) cudaFree(h->d_scales);
  if (h->d_stage) cudaFree(h->d_stage);
  if (h->d_song) cudaFree(h->d_song);
  if (h->d_win) cudaFree(h->d_win);
  if (h->d_tw) cudaFree(h->d_tw);
  if (h->d_pairs) cudaFree(h->d_pairs);
  if (h->d_weights) cudaFree(h->d_weights);
  if (h->d_pairs_rem) cudaFree(h->d_pairs_rem);
  if (h->d_weights_rem) cudaFree(h->d_weights_rem);
  if (h->d_rounds) cudaFree(h->d_rounds);
  if (h->d_g0) cudaFree(h->d_g0);
  if (h->d_wt) cudaFree(h->d_wt);
  if (prev >= 0) cudaSetDevice(prev);
  delete h;
}

int bhmel_set_fb(bhmel_handle* h, const float* fb_host) {
  if (!h || !fb_host) return fail(BHMEL_EINVAL, "null argument");
  if (int rc = check_device(h)) return rc;
  h->fb.assign(fb_host, fb_host + static_cast<size_t>(bhmel::kBins) * h->prm.n_mels);
  return upload_filterbank(h);
}

int bhmel_set_window(bhmel_handle* h, const float* window_host) {
  if (!h || !window_host) return fail(BHMEL_EINVAL, "null argument");
  if (int rc = check_device(h)) return rc;
  h->window.assign(window_host, window_host + bhmel::kNfft);
  return upload_window(h);
}

int bhmel_get_fb(const bhmel_handle* h, float* fb_host) {
  if (!h || !fb_host) return fail(BHMEL_EINVAL, "null argument");
  std::memcpy(fb_host, h->fb.data(), h->fb.size() * sizeof(float));
  return BHMEL_OK;
}

int bhmel_get_window(const bhmel_handle* h, float* window_host) {
  if (!h || !window_host) return fail(BHMEL_EINVAL, "null argument");
  std::memcpy(window_host, h->window.data(), h->window.size() * sizeof(float));
  return BHMEL_OK;
}

int64_t bhmel_num_frames(const bhmel_handle* h, int64_t n_samples) {
  (void)h;
  return n_samples < 0 ? 0 : n_samples / bhmel::kHop + 1;
}

int bhmel_set_option(bhmel_handle* h, int32_t option, int64_t value) {
  if (!h) return fail(BHMEL_EINVAL, "null handle");
  switch (option) {
    case BHMEL_OPT_BULK_COPY:
      h->use_bulk = value != 0;
      return BHMEL_OK;
    case BHMEL_OPT_KERNEL:
      if (value != BHMEL_KERNEL_BARRIER && value != BHMEL_KERNEL_INDEPENDENT_WARPS &&
          value != BHMEL_KERNEL_WARP_SPECIALIZED)
        return fail(BHMEL_EINVAL, "unknown kernel variant");
      h->kernel_variant = static_cast<int>(value);
      return BHMEL_OK;
    case BHMEL_OPT_STATIC_MEL:
      if (value < 0 || value > 2) return fail(BHMEL_EINVAL, "BHMEL_OPT_STATIC_MEL takes 0, 1 or 2");
      h->static_mel = static_cast<int>(value);
      return BHMEL_OK;
    case BHMEL_OPT_PDL:
      if (value != 0 && value != 1) return fail(BHMEL_EINVAL, "BHMEL_OPT_PDL takes 0 or 1");
      h->pdl = static_cast<int>(value);
      return BHMEL_OK;
    default:
      return fail(BHMEL_EINVAL, "unknown option " + std::to_string(option));
  }
}

int64_t bhmel_launch_count(const bhmel_handle* h) { return h ? h->launches.load() : 0; }

int bhmel_forward(bhmel_handle* h, const float* x, int64_t B, int64_t N, int64_t x_row_stride, float* y,
                  void* stream) {
  if (x_row_stride < N) return fail(BHMEL_EINVAL, "x_row_stride must be >= N");
  return launch(h, x, x_row_stride, 0, LLONG_MAX, B, N, OutSpec{y, 0, 0, 0}, static_cast<cudaStream_t>(stream));
}

int bhmel_forward_ex(bhmel_handle* h, const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                     const bhmel_out_desc* out, void* stream) {
  if (!out) return fail(BHMEL_EINVAL, "null output descriptor");
  if (x_row_stride < N) return fail(BHMEL_EINVAL, "x_row_stride must be >= N");
  if (out->dtype != BHMEL_OUT_F32 && out->dtype != BHMEL_OUT_BF16) return fail(BHMEL_EINVAL, "unknown output dtype");
  return launch(h, x, x_row_stride, 0, LLONG_MAX, B, N,
                OutSpec{out->y, out->dtype == BHMEL_OUT_BF16, out->frame_pitch, out->row_pitch},
                static_cast<cudaStream_t>(stream));
}

int bhmel_forward_encoder_input(bhmel_handle* h, const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                                const bhmel_encoder_input_desc* d, void* stream) {
  if (!h) return fail(BHMEL_EINVAL, "null handle");
  if (!d || !d->y) return fail(BHMEL_EINVAL, "null output descriptor / pointer");
  if (x_row_stride < N) return fail(BHMEL_EINVAL, "x_row_stride must be >= N");
  if (d->dtype != BHMEL_OUT_F32 && d->dtype != BHMEL_OUT_BF16) return fail(BHMEL_EINVAL, "unknown output dtype");
  if (d->layout != BHMEL_LAYOUT_BTC && d->layout != BHMEL_LAYOUT_BCT) return fail(BHMEL_EINVAL, "unknown layout");
  if (d->n_cond < 0 || d->n_cond > 65536) return fail(BHMEL_EINVAL, "n_cond out of range");
  if (d->n_cond > 0 && !d->cond) return fail(BHMEL_EINVAL, "null conditioning pointer");
  if (B > 65535 && d->layout == BHMEL_LAYOUT_BCT) return fail(BHMEL_ESHAPE, "batch too large for the BCT layout");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool bf16 = d->dtype == BHMEL_OUT_BF16;
  const long long Tn = N / bhmel::kHop + 1;
  const long long C = h->prm.n_mels + d->n_cond;
  const void* staged = nullptr;
  if (d->layout == BHMEL_LAYOUT_BTC) {
    if (int rc = launch(h, x, x_row_stride, 0, LLONG_MAX, B, N, OutSpec{d->y, bf16 ? 1 : 0, C, Tn * C}, s)) return rc;
  } else {
    if (B <= 0 || N <= 0) return fail(BHMEL_ESHAPE, "batch and sample count must be positive");
    if (int rc = check_device(h)) return rc;
    void* stage = d->scratch;
    if (!stage) {
      std::lock_guard<std::mutex> lock(h->host_mu);
      const size_t need = static_cast<size_t>(B) * Tn * h->prm.n_mels * (bf16 ? 2 : 4);
      if (need > h->cap_stage) {
        BH_CUDA(cudaDeviceSynchronize());   // the old scratch may still be in use by kernels in flight
        if (h->d_stage) cudaFree(h->d_stage);
        h->d_stage = nullptr;
        h->cap_stage = 0;
        BH_CUDA(cudaMalloc(&h->d_stage, need));
        h->cap_stage = need;
      }
      stage = h->d_stage;
    }
    if (int rc = launch(h, x, x_row_stride, 0, LLONG_MAX, B, N, OutSpec{stage, bf16 ? 1 : 0, 0, 0}, s)) return rc;
    staged = stage;
  }
  if (bf16) return encoder_input_tail<__nv_bfloat16>(h, d, static_cast<const __nv_bfloat16*>(staged), B, Tn, s);
  return encoder_input_tail<float>(h, d, static_cast<const float*>(staged), B, Tn, s);
}

int bhmel_forward_gather(bhmel_handle* h, const float* song, int64_t n_song, int64_t first_offset,
                         int64_t stride, int64_t W, int64_t window_len, float* y, void* stream) {
  if (n_song < 0 || first_offset < 0 || stride <= 0)
    return fail(BHMEL_EINVAL, "need n_song >= 0, first_offset >= 0, stride > 0");
  return launch(h, song, stride, first_offset, n_song, W, window_len, OutSpec{y, 0, 0, 0},
                static_cast<cudaStream_t>(stream));
}

int bhmel_peak_scale_pcm16(bhmel_handle* h, const int16_t* pcm_dev, int64_t n, float* scale_dev, void* stream) {
  if (!h) return fail(BHMEL_EINVAL, "null handle");
  if (!pcm_dev || !scale_dev) return fail(BHMEL_EINVAL, "null data pointer");
  if (n <= 0) return fail(BHMEL_ESHAPE, "sample count must be positive");
  if (reinterpret_cast<uintptr_t>(pcm_dev) & 1) return fail(BHMEL_EINVAL, "pcm_dev must be 2-byte aligned");
  if (int rc = check_device(h)) return rc;
  // The reduction runs in the caller's own 4 bytes (integer max first, converted in place), so the
  // entry owns no handle state: any number of threads / streams may share one handle.
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned* peak = reinterpret_cast<unsigned*>(scale_dev);
  BH_CUDA(cudaMemsetAsync(peak, 0, sizeof(unsigned), s));
  const long long work = (n + 8 * 256 - 1) / (8 * 256);
  const unsigned blocks = static_cast<unsigned>(work < h->num_sms * 8 ? (work < 1 ? 1 : work) : h->num_sms * 8);
  bhmel_absmax_pcm16_kernel<<<blocks, 256, 0, s>>>(pcm_dev, n, peak);
  bhmel_peak_to_scale_kernel<<<1, 1, 0, s>>>(scale_dev);
  BH_CUDA(cudaGetLastError());
  h->launches.fetch_add(2, std::memory_order_relaxed);
  return BHMEL_OK;
}

int bhmel_forward_gather_pcm16(bhmel_handle* h, const int16_t* song_dev, int64_t n_song, const float* scale_dev,
                               int64_t first_offset, int64_t stride, int64_t W, int64_t window_len,
                               float* y, float* scratch, void* stream) {
  if (!h) return fail(BHMEL_EINVAL, "null handle");
  if (!song_dev) return fail(BHMEL_EINVAL, "null data pointer");
  if (n_song <= 0 || first_offset < 0 || stride <= 0)
    return fail(BHMEL_EINVAL, "need n_song > 0, first_offset >= 0, stride > 0");
  if (int rc = check_device(h)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!scratch) {
    std::lock_guard<std::mutex> lock(h->host_mu);
    const size_t need = static_cast<size_t>(n_song) * sizeof(float);
    if (need > h->cap_song) {
      BH_CUDA(cudaDeviceSynchronize());   // the old scratch may still be read by a kernel in flight
      if (h->d_song) cudaFree(h->d_song);
      h->d_song = nullptr;
      h->cap_song = 0;
      BH_CUDA(cudaMalloc(&h->d_song, need));
      h->cap_song = need;
    }
    scratch = h->d_song;
  }
  const long long work = (n_song + 255) / 256;
  const unsigned blocks = static_cast<unsigned>(work < h->num_sms * 8 ? work : h->num_sms * 8);
  // one "row" of n_song samples: the per-row scale pointer is the song's scale (or none)
  bhmel_pcm16_to_f32_kernel<<<blocks, 256, 0, s>>>(song_dev, scratch, scale_dev, n_song, n_song);
  BH_CUDA(cudaGetLastError());
  h->launches.fetch_add(1, std::memory_order_relaxed);
  return launch(h, scratch, stride, first_offset, n_song, W, window_len, OutSpec{y, 0, 0, 0}, s);
}

// Row chunks of one bhmel_forward_host_ex call.  ~16 MB of fp32 / ~32 MB of int16 source per chunk (8 / 32
// model-context windows) keeps both copy engines busy; the first and last chunks are tapered (1/8, 1/4, 1/2 of a
// chunk) because nothing overlaps the very first host->device copy and the very last device->host copy of a call
// (measured on B200 with tools/host_chunk_probe.py: fp32 12 / 16 / 20 / 24 / 32 MB chunks 10.63 / 10.44-10.51 /
// 10.53 / 10.57 / 10.77-11.0 ms per 256 windows, int16 16 / 32 / 40 / 48 / 64 MB 5.77 / 5.60 / 5.60 / 5.90 / 6.26 ms;
// the taper is worth 0.3 % / 1.2 %).
static std::vector<int64_t> host_chunk_rows(int64_t B, size_t row_src, bool pcm) {
  int64_t rows = static_cast<int64_t>((static_cast<size_t>(pcm ? BHMEL_HOST_CHUNK_MB_PCM : BHMEL_HOST_CHUNK_MB) << 20) / row_src);
  if (rows < 1) rows = 1;
  if (rows > B) rows = B;
  if (rows * 2 > B && B >= 2 * kHostSlots) rows = (B + 2 * kHostSlots - 1) / (2 * kHostSlots);
  std::vector<int64_t> chunk_rows;
  int64_t head[3], n_head = 0, taper_total = 0;
  if (BHMEL_HOST_TAPER && B >= 6 * rows)
    for (int64_t r = rows >= 8 ? rows / 8 : 1; r < rows && n_head < 3; r *= 2) { head[n_head++] = r; taper_total += r; }
  for (int64_t i = 0; i < n_head; ++i) chunk_rows.push_back(head[i]);
  int64_t middle = B - 2 * taper_total;
  while (middle > 0) {
    const int64_t r = middle < rows ? middle : rows;
    chunk_rows.push_back(r);
    middle -= r;
  }
  for (int64_t i = n_head - 1; i >= 0; --i) chunk_rows.push_back(head[i]);
  return chunk_rows;
}

int bhmel_forward_host_ex(bhmel_handle* h, const bhmel_host_io* io, int64_t B, int64_t N, int64_t x_row_stride) {
  if (!h) return fail(BHMEL_EINVAL, "null handle");
  if (!io || !io->x_host || !io->y_host) return fail(BHMEL_EINVAL, "null data pointer");
  if (B <= 0 || N <= 0) return fail(BHMEL_ESHAPE, "batch and sample count must be positive");
  if (x_row_stride < N) return fail(BHMEL_EINVAL, "x_row_stride must be >= N");
  if (io->x_dtype != BHMEL_IN_F32 && io->x_dtype != BHMEL_IN_PCM16) return fail(BHMEL_EINVAL, "unknown input dtype");
  if (io->y_dtype != BHMEL_OUT_F32 && io->y_dtype != BHMEL_OUT_BF16) return fail(BHMEL_EINVAL, "unknown output dtype");
  if (int rc = check_device(h)) return rc;
  std::lock_guard<std::mutex> lock(h->host_mu);

  const bool pcm = io->x_dtype == BHMEL_IN_PCM16;
  const bool bf16 = io->y_dtype == BHMEL_OUT_BF16;
  const int64_t T = N / bhmel::kHop + 1;
  const size_t row_f32 = static_cast<size_t>(N) * sizeof(float);
  const size_t row_src = static_cast<size_t>(N) * (pcm ? sizeof(int16_t) : sizeof(float));
  const size_t row_out = static_cast<size_t>(T) * h->prm.n_mels * (bf16 ? 2 : 4);
  const std::vector<int64_t> chunk_rows = host_chunk_rows(B, row_src, pcm);
  int64_t rows = 1;
  for (int64_t r : chunk_rows) rows = r > rows ? r : rows;
  const size_t need_in = static_cast<size_t>(rows) * row_f32;
  const size_t need_out = static_cast<size_t>(rows) * static_cast<size_t>(T) * h->prm.n_mels * 4;
  const size_t need_pcm = pcm ? static_cast<size_t>(rows) * N * sizeof(int16_t) : 0;
  for (int i = 0; i < kHostSlots; ++i)
    if (!h->hs[i]) BH_CUDA(cudaStreamCreateWithFlags(&h->hs[i], cudaStreamNonBlocking));
  if (need_in > h->cap_in || need_out > h->cap_out || need_pcm > h->cap_pcm) {
    for (int i = 0; i < kHostSlots; ++i) {
      BH_CUDA(cudaStreamSynchronize(h->hs[i]));
      if (h->d_in[i]) cudaFree(h->d_in[i]);
      if (h->d_out[i]) cudaFree(h->d_out[i]);
      if (h->d_pcm[i]) cudaFree(h->d_pcm[i]);
      h->d_in[i] = h->d_out[i] = nullptr;
      h->d_pcm[i] = nullptr;
    }
    h->cap_in = h->cap_out = h->cap_pcm = 0;
    const size_t ci = need_in, co = need_out, cp = need_pcm;
    for (int i = 0; i < kHostSlots; ++i) {
      BH_CUDA(cudaMalloc(&h->d_in[i], ci));
      BH_CUDA(cudaMalloc(&h->d_out[i], co));
      if (cp) BH_CUDA(cudaMalloc(&h->d_pcm[i], cp));
    }
    h->cap_in = ci;
    h->cap_out = co;
    h->cap_pcm = cp;
  }
  const float* d_scales = nullptr;
  if (pcm && io->scales) {
    if (static_cast<size_t>(B) * sizeof(float) > h->cap_scales) {
      if (h->d_scales) cudaFree(h->d_scales);
      h->d_scales = nullptr;
      h->cap_scales = 0;
      BH_CUDA(cudaMalloc(&h->d_scales, static_cast<size_t>(B) * sizeof(float)));
      h->cap_scales = static_cast<size_t>(B) * sizeof(float);
    }
    for (int i = 0; i < kHostSlots; ++i) BH_CUDA(cudaStreamSynchronize(h->hs[i]));
    BH_CUDA(cudaMemcpy(h->d_scales, io->scales, static_cast<size_t>(B) * sizeof(float), cudaMemcpyHostToDevice));
    d_scales = h->d_scales;
  }
  // Every exit below this point first drains the private streams: the caller owns the host buffers again the
  // moment this function returns, with or without an error.
  auto enqueue = [&]() -> int {
    int slot = 0;
    int64_t b0 = 0;
    for (size_t c = 0; c < chunk_rows.size(); b0 += chunk_rows[c], ++c, slot = (slot + 1) % kHostSlots) {
      const int64_t nb = chunk_rows[c];
      cudaStream_t s = h->hs[slot];
      if (pcm) {
        const int16_t* src = static_cast<const int16_t*>(io->x_host) + b0 * x_row_stride;
        BH_CUDA(cudaMemcpy2DAsync(h->d_pcm[slot], row_src, src, static_cast<size_t>(x_row_stride) * sizeof(int16_t),
                                  row_src, static_cast<size_t>(nb), cudaMemcpyHostToDevice, s));
        const long long total = static_cast<long long>(nb) * N;
        const unsigned blocks = static_cast<unsigned>(h->num_sms * 8);
        bhmel_pcm16_to_f32_kernel<<<blocks, 256, 0, s>>>(h->d_pcm[slot], h->d_in[slot],
                                                         d_scales ? d_scales + b0 : nullptr, N, total);
        BH_CUDA(cudaGetLastError());
        h->launches.fetch_add(1, std::memory_order_relaxed);
      } else {
        const float* src = static_cast<const float*>(io->x_host) + b0 * x_row_stride;
        BH_CUDA(cudaMemcpy2DAsync(h->d_in[slot], row_f32, src, static_cast<size_t>(x_row_stride) * sizeof(float),
                                  row_f32, static_cast<size_t>(nb), cudaMemcpyHostToDevice, s));
      }
      if (int rc = launch(h, h->d_in[slot], N, 0, LLONG_MAX, nb, N, OutSpec{h->d_out[slot], bf16 ? 1 : 0, 0, 0}, s))
        return rc;
      BH_CUDA(cudaMemcpyAsync(static_cast<char*>(io->y_host) + static_cast<size_t>(b0) * row_out, h->d_out[slot],
                              static_cast<size_t>(nb) * row_out, cudaMemcpyDeviceToHost, s));
    }
    return BHMEL_OK;
  };
  const int rc = enqueue();
  cudaError_t sync_err = cudaSuccess;
  for (int i = 0; i < kHostSlots; ++i) {
    const cudaError_t e = cudaStreamSynchronize(h->hs[i]);
    if (e != cudaSuccess && sync_err == cudaSuccess) sync_err = e;
  }
  if (rc != BHMEL_OK) return rc;      // bhmel_last_error() already holds the first failure
  if (sync_err != cudaSuccess) return fail(BHMEL_ECUDA, std::string("bhmel_forward_host: ") + cudaGetErrorString(sync_err));
  return BHMEL_OK;
}

int64_t bhmel_host_chunk_plan(int64_t B, int64_t N, int32_t x_dtype, int64_t* rows_out, int64_t cap) {
  if (B <= 0 || N <= 0 || (x_dtype != BHMEL_IN_F32 && x_dtype != BHMEL_IN_PCM16)) return 0;
  const bool pcm = x_dtype == BHMEL_IN_PCM16;
  const std::vector<int64_t> c = host_chunk_rows(B, static_cast<size_t>(N) * (pcm ? sizeof(int16_t) : sizeof(float)), pcm);
  for (size_t i = 0; i < c.size() && rows_out && static_cast<int64_t>(i) < cap; ++i) rows_out[i] = c[i];
  return static_cast<int64_t>(c.size());
}

int bhmel_forward_host(bhmel_handle* h, const float* x_host, int64_t B, int64_t N, int64_t x_row_stride,
                       float* y_host) {
  bhmel_host_io io{};
  io.x_host = x_host;
  io.x_dtype = BHMEL_IN_F32;
  io.scales = nullptr;
  io.y_host = y_host;
  io.y_dtype = BHMEL_OUT_F32;
  return bhmel_forward_host_ex(h, &io, B, N, x_row_stride);
}

}  // extern "C"

/* bhmel.h -- C ABI of libbhmel.so: the B200-native (sm_100a) replacement for the hot path of
 * BeatHeritage's audio frontend, osuT5/osuT5/model/spectrogram.py::MelSpectrogram
 * (reference file:line cited per entry point below; "ref:" paths are relative to the
 * reference repository root).
 *
 * All entry points are plain C: pointers, sizes and a CUDA stream handle -- no torch types.
 * Device pointers are owned by the caller; the handle owns only its constant tables (and, for
 * bhmel_forward_host, its private staging buffers).  bhmel_forward / bhmel_forward_gather
 * launch on the caller's stream, never synchronise the device and allocate nothing.
 * Every function returning int returns 0 on success and a BHMEL_E* code otherwise; the
 * message is available through bhmel_last_error() (thread-local).
 * The library is re-entrant: no global mutable state besides the thread-local error string.
 */
#ifndef BHMEL_H_
#define BHMEL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BHMEL_VERSION 101          /* 0.1.1 */

#define BHMEL_OK 0
#define BHMEL_EINVAL 1             /* bad argument / unsupported parameter combination */
#define BHMEL_ECUDA 2              /* CUDA runtime error (message carries cudaGetErrorString) */
#define BHMEL_ESHAPE 3             /* input too short for reflect padding, empty input, ... */
#define BHMEL_EDEVICE 4            /* not an sm_100 device / wrong current device */

#define BHMEL_PAD_CONSTANT 0       /* ref: configs/model/default.yaml:27 pad_mode 'constant' */
#define BHMEL_PAD_REFLECT 1        /* ref: configs/model/whisper_small_v2.yaml:21 pad_mode 'reflect' */

/* Output element type of the forward calls (the reference always returns float32,
 * ref: osuT5/osuT5/model/spectrogram.py:79-83). */
#define BHMEL_OUT_F32 0
#define BHMEL_OUT_BF16 1

typedef struct bhmel_handle bhmel_handle;

/* Mirrors the constructor arguments of the reference module
 * (ref: osuT5/osuT5/model/spectrogram.py:8-19; values from configs/model/default.yaml:18-27 and
 * configs/model/whisper_small_v2.yaml:16-21).  n_fft must be 1024 and hop_length 128 (constant in
 * every reference config); 1 <= n_mels <= 1024. */
typedef struct bhmel_params {
  int32_t sample_rate;     /* 16000 */
  int32_t n_fft;           /* 1024 ("n_ftt" in the reference signature) */
  int32_t hop_length;      /* 128 */
  int32_t n_mels;          /* 80 / 128 / 388 / 512 */
  double f_min;            /* Hz */
  double f_max;            /* Hz */
  int32_t pad_mode;        /* BHMEL_PAD_* (torch.stft center=True padding) */
  int32_t log_scale;       /* non-zero: log1p epilogue (ref: spectrogram.py:80-81) */
  const float* fb;         /* optional HOST [n_fft/2+1][n_mels] row-major filterbank override (the
                              state-dict buffer transform.mel_scale.fb); NULL -> htk/norm=None
                              triangular bank as torchaudio.functional.melscale_fbanks builds it */
  const float* window;     /* optional HOST [n_fft] window override (state-dict buffer
                              transform.spectrogram.window); NULL -> periodic Hann */
} bhmel_params;

/* Replaces MelSpectrogram.__init__ (ref: spectrogram.py:8-61): validates the parameters and
 * builds the device-resident constant tables (window, twiddles, banded filterbank) on the
 * CURRENT CUDA device.  */
int bhmel_create(const bhmel_params* params, bhmel_handle** out);
void bhmel_destroy(bhmel_handle* h);

/* Replaces the persistent buffers of the reference module (state_dict keys
 * transform.spectrogram.window / transform.mel_scale.fb): set from / copy to HOST arrays.
 * Setting rebuilds the derived device tables (synchronises the handle's device). */
int bhmel_set_fb(bhmel_handle* h, const float* fb_host /* [n_fft/2+1][n_mels] */);
int bhmel_set_window(bhmel_handle* h, const float* window_host /* [n_fft] */);
int bhmel_get_fb(const bhmel_handle* h, float* fb_host);
int bhmel_get_window(const bhmel_handle* h, float* window_host);

/* Number of output frames for an input of n_samples: n_samples / hop + 1
 * (ref: spectrogram.py:71 "n_frames = n_samples // hop_length + 1"). */
int64_t bhmel_num_frames(const bhmel_handle* h, int64_t n_samples);

/* Replaces MelSpectrogram.forward (ref: spectrogram.py:63-83):
 *   x  DEVICE float32 [B][N] with row stride x_row_stride (elements, >= N)
 *   y  DEVICE float32 [B][N/hop+1][n_mels], contiguous -- the encoder's [batch, frames, n_mels]
 * One fused kernel: centre pad (reflect/constant) + framing + Hann + 1024-pt real FFT +
 * |X|^2 + mel filterbank + log1p.  Asynchronous on `stream` (a cudaStream_t).
 * Errors: BHMEL_ESHAPE if N <= n_fft/2 with reflect padding (the reference raises RuntimeError
 * from F.pad), or B/N <= 0. */
int bhmel_forward(bhmel_handle* h, const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                  float* y, void* stream);

/* Output descriptor for bhmel_forward_ex: element (row r, frame t, mel m) is written at
 * y[r * row_pitch + t * frame_pitch + m] (pitches in ELEMENTS of dtype; 0 selects the dense
 * [B][T][n_mels] layout).  With frame_pitch > n_mels the mel channels land inside a wider
 * per-frame record, e.g. the encoder input [B][T][n_mels + cond] the reference builds with
 * `.to(dtype)` + `torch.cat` (ref: osuT5/osuT5/model/modeling_mapperatorinator.py:352, 369-370). */
typedef struct bhmel_out_desc {
  void* y;               /* DEVICE pointer to element (0, 0, 0) */
  int32_t dtype;         /* BHMEL_OUT_F32 or BHMEL_OUT_BF16 (round to nearest even, == tensor.to(bfloat16)) */
  int64_t frame_pitch;   /* >= n_mels, or 0 */
  int64_t row_pitch;     /* >= T * frame_pitch, or 0 */
} bhmel_out_desc;

/* bhmel_forward with a typed / pitched output (next-row N1 of SURVEY.md 8f: removes the caller's
 * dtype cast and concat passes over the encoder input).  Same arithmetic; the fp32 result is
 * converted once at the store. */
int bhmel_forward_ex(bhmel_handle* h, const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                     const bhmel_out_desc* out, void* stream);

/* Encoder-input assembly (next-row N1 of SURVEY.md 8f).  Replaces, after MelSpectrogram.forward,
 *   frames = frames.to(dtype); conds_expanded = [c.unsqueeze(1).expand(-1, T, -1) ...];
 *   inputs_embeds = torch.concatenate([frames] + conds_expanded, dim=-1)       -> layout BTC
 *   inputs_embeds = torch.swapaxes(inputs_embeds, 1, 2)                        -> layout BCT
 * (ref: osuT5/osuT5/model/modeling_mapperatorinator.py:351-352, 368-376): the mel channels are
 * written once, already converted, into channels 0..n_mels-1 of the [B][T][C] / [B][C][T] buffer
 * (C = n_mels + n_cond) and the conditioning vector of each batch row is broadcast over its T
 * frames into channels n_mels..C-1, so the dtype cast, expand, concatenate and transpose passes
 * over the encoder input disappear.  Bit-identical to those torch ops. */
#define BHMEL_LAYOUT_BTC 0   /* [B][T][C] channels last  (what torch.concatenate builds)              */
#define BHMEL_LAYOUT_BCT 1   /* [B][C][T] channels first (contiguous form of the swapaxes view)      */
typedef struct bhmel_encoder_input_desc {
  void* y;            /* DEVICE, contiguous B*T*C elements of dtype in `layout`                      */
  int32_t dtype;      /* BHMEL_OUT_F32 or BHMEL_OUT_BF16                                              */
  int32_t layout;     /* BHMEL_LAYOUT_BTC or BHMEL_LAYOUT_BCT                                         */
  const void* cond;   /* DEVICE [B][n_cond] of dtype: the concatenated conditioning embeddings; NULL
                         only if n_cond == 0                                                          */
  int64_t n_cond;     /* >= 0                                                                          */
  void* scratch;      /* BCT only: DEVICE B*T*n_mels elements of dtype owned by the caller, or NULL to
                         use a handle-owned buffer (grown on demand; then not re-entrant per handle)   */
} bhmel_encoder_input_desc;
/* BTC: the fused kernel stores straight into y (pitched), one broadcast-fill kernel follows.
 * BCT: the fused kernel stores [B][T][n_mels] of dtype into the scratch, one assembly kernel
 * transposes it and fills the conditioning rows. */
int bhmel_forward_encoder_input(bhmel_handle* h, const float* x, int64_t B, int64_t N, int64_t x_row_stride,
                                const bhmel_encoder_input_desc* out, void* stream);

/* Fused segmentation + forward.  Replaces Preprocessor.segment/window followed by forward
 * (ref: osuT5/osuT5/inference/preprocessor.py:58-71, 94-102): window w (0 <= w < W) covers
 * song[first_offset + w*stride ... + window_len), samples at or beyond n_song read as zero
 * (the right padding segment() applies); every window is centre-padded on its own exactly as
 * forward does.   song DEVICE float32 [n_song];  y DEVICE float32 [W][window_len/hop+1][n_mels]. */
int bhmel_forward_gather(bhmel_handle* h, const float* song, int64_t n_song, int64_t first_offset,
                         int64_t stride, int64_t W, int64_t window_len, float* y, void* stream);

/* Peak normalisation scalar of an int16 PCM song that is resident on the DEVICE (next-row N2 of
 * SURVEY.md 8f).  Writes  scale = 1.0f / max|pcm[i]|  (float32 division, i < n) to *scale_dev --
 * the factor of the reference's `samples *= 1.0 / np.max(np.abs(samples))` after its
 * `.astype(np.float32)` (ref: osuT5/osuT5/dataset/data_utils.py:94-96).  An all-zero song gives
 * +inf exactly like the reference's division by zero (its samples then become NaN).
 * pcm_dev DEVICE int16 [n];  scale_dev DEVICE float32 [1], 4-byte aligned: it doubles as the
 * reduction's scratch (integer max first, converted in place), so the call owns no handle state
 * and is re-entrant across threads and streams sharing one handle.  Asynchronous on `stream`. */
int bhmel_peak_scale_pcm16(bhmel_handle* h, const int16_t* pcm_dev, int64_t n, float* scale_dev, void* stream);

/* bhmel_forward_gather for a song kept on the device as int16 PCM (2 bytes per sample resident
 * instead of 4): sample i is used as  float32(pcm[i]) * *scale_dev  (scale_dev == NULL: 1.0), i.e.
 * the reference's int16 -> float32 cast + peak normalisation (data_utils.py:94-96) followed by
 * Preprocessor.segment/window + forward (preprocessor.py:58-71, 94-102).  The conversion runs as a
 * bandwidth-bound pre-pass into a float32 scratch of n_song samples, then the fused kernel gathers
 * the windows from it; results are bit-identical to bhmel_forward_gather on the converted song.
 * scratch: DEVICE float32 [n_song] owned by the caller, or NULL to use a handle-owned buffer (grown
 * on demand -- the only allocation this entry may make -- and then not re-entrant per handle). */
int bhmel_forward_gather_pcm16(bhmel_handle* h, const int16_t* song_dev, int64_t n_song, const float* scale_dev,
                               int64_t first_offset, int64_t stride, int64_t W, int64_t window_len,
                               float* y, float* scratch, void* stream);

/* Same contract as bhmel_forward with HOST buffers (pinned memory recommended): chunks the
 * batch, and overlaps host->device copy, the kernel and device->host copy on private streams.
 * Synchronous: returns when y_host is complete.  This is the end-to-end entry the plugin uses
 * when the caller's data lives on the host (ref: osuT5/dataloading.py:128-130 calls the module
 * with CPU tensors; osuT5/osuT5/inference/server.py:42 does the H2D copy for inference). */
int bhmel_forward_host(bhmel_handle* h, const float* x_host, int64_t B, int64_t N,
                       int64_t x_row_stride, float* y_host);

/* The row chunks bhmel_forward_host / bhmel_forward_host_ex split a [B][N] batch into (x_dtype BHMEL_IN_F32 or
 * BHMEL_IN_PCM16): chunk i covers rows_out[i] consecutive rows; chunks go round-robin over the handle's private
 * streams, each as one host->device copy, the kernel(s), one device->host copy.  Returns the number of chunks
 * (0 for invalid arguments) and fills at most `cap` entries; rows_out may be NULL.  Pure host arithmetic -- it
 * exists so that a benchmark can time plain copies in exactly the pattern the entry uses (bench.py,
 * e2e.plain_copy_ceiling). */
int64_t bhmel_host_chunk_plan(int64_t B, int64_t N, int32_t x_dtype, int64_t* rows_out, int64_t cap);

/* Tuning / debugging switches (per handle).  BHMEL_OPT_BULK_COPY: 1 (default) stages aligned
 * interior tiles with the TMA bulk copy, 0 forces the per-element cp.async path everywhere. */
#define BHMEL_OPT_BULK_COPY 1
/* BHMEL_OPT_KERNEL: which schedule of the fused kernel to launch.  All three run the same
 * arithmetic and produce bit-identical results:
 *   BHMEL_KERNEL_WARP_SPECIALIZED (default) 32-frame tiles; 8 FFT warps and 8 producer/mel/store
 *       warps of one 512-thread CTA run concurrently and hand tiles over through mbarriers;
 *   BHMEL_KERNEL_BARRIER            32-frame tiles; 8 warps step through the stages together,
 *       separated by CTA barriers;
 *   BHMEL_KERNEL_INDEPENDENT_WARPS  8-frame per-warp tiles; every warp is its own
 *       load -> FFT -> mel -> store pipeline, no CTA barriers. */
#define BHMEL_OPT_KERNEL 2
#define BHMEL_KERNEL_BARRIER 0
#define BHMEL_KERNEL_INDEPENDENT_WARPS 1
#define BHMEL_KERNEL_WARP_SPECIALIZED 2
/* BHMEL_OPT_STATIC_MEL: 1 (default) lets the warp-specialised kernel run the mel stage as generated
 * straight-line code (weights as instruction immediates, every power-spectrum block read once)
 * whenever the handle's filterbank equals, bit for bit, a table baked into the library: the
 * filterbanks of every parameter set in the reference's configs (80 / 128 htk mels 20..8000 Hz,
 * 388 / 512 htk mels 0..8000 Hz at 16 kHz; ref: configs/model/whisper_small_v2.yaml:16-21,
 * configs/train/tiny_dist22.yaml:10-12, configs/model/default.yaml:18-27, configs/model/t5_small.yaml:9-10).
 * They take the "direct" form (one or two summation chains per filter, results staged per warp and
 * written as row segments -- 16-byte / 8-byte vector stores when the output rows are aligned that way,
 * element stores otherwise, same values; a few ulp from the generic stage, same tolerance against the
 * reference).
 * 0 forces the generic descriptor-driven stage (pair tables); 2 gives the 80-mel table its "hybrid" form
 * instead (half the mel warps on generated code, half on pair tables: bit-identical to the generic stage).
 * Any other filterbank always takes the generic stage. */
#define BHMEL_OPT_STATIC_MEL 3
/* BHMEL_OPT_PDL: 1 (default) launches the warp-specialised kernel with programmatic stream serialisation, so
 * its prologue (constant-table staging, barrier init) overlaps the tail of the previous kernel in the stream.
 * The kernel reads no sample and writes no output before that kernel has completed (griddepcontrol.wait);
 * results and stream-order semantics are unchanged.  0: plain stream order. */
#define BHMEL_OPT_PDL 4
int bhmel_set_option(bhmel_handle* h, int32_t option, int64_t value);

/* Host-buffer entry with typed input / output (next rows N1 + N2 of SURVEY.md 8f).
 *   x_dtype BHMEL_IN_F32   : x_host is float32 [B][x_row_stride], exactly bhmel_forward_host.
 *   x_dtype BHMEL_IN_PCM16 : x_host is int16 PCM [B][x_row_stride]; row r is converted on the device as
 *                            float32(pcm) * scales[r] -- the reference's int16 -> float32 cast and
 *                            peak normalisation `samples *= 1.0 / max(abs(samples))`
 *                            (ref: osuT5/osuT5/dataset/data_utils.py:95-97) -- so only 2 bytes per sample
 *                            cross PCIe.  scales: HOST float32 [B], or NULL for 1.0.
 *   y_dtype BHMEL_OUT_F32 / BHMEL_OUT_BF16 : y_host is [B][N/hop+1][n_mels] of that type.
 * Synchronous; chunks are pipelined over private streams like bhmel_forward_host. */
#define BHMEL_IN_F32 0
#define BHMEL_IN_PCM16 1
typedef struct bhmel_host_io {
  const void* x_host;
  int32_t x_dtype;
  const float* scales;
  void* y_host;
  int32_t y_dtype;
} bhmel_host_io;
int bhmel_forward_host_ex(bhmel_handle* h, const bhmel_host_io* io, int64_t B, int64_t N, int64_t x_row_stride);

/* Introspection used by tests and bench.py. */
int bhmel_version(void);
const char* bhmel_last_error(void);
/* Number of kernel launches issued through this handle so far (bench.py's gpu_launches). */
int64_t bhmel_launch_count(const bhmel_handle* h);
/* Static facts about the default (warp-specialised) kernel: dynamic shared memory bytes,
 * threads per CTA, frames per tile. Any pointer may be NULL. */
void bhmel_kernel_info(int32_t* smem_bytes, int32_t* threads, int32_t* tile_frames);

#ifdef __cplusplus
}
#endif
#endif /* BHMEL_H_ */

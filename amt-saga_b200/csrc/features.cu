// K5: feature gather for the classifiers.
// Replaces, from the producer loop /root/reference/training.py:333-388:
//   compress_bands(mag, bands) [util_audio.py:436-466]  -> saga_compress_bands_exec
//   resize(...) + section_power + log10(1000 x + 1)/max + (angle(ph)+3.15)/6.3
//     [util_audio.py:469-507, :334-349; training.py:347-363]   -> saga_short_window_exec
//   librosa.feature.spectral_flatness [util_audio.py:330-332]  -> saga_spectral_flatness_exec
// All operate on frame-major storage (saga_b200.h) and write frame-major outputs.
#include "saga_common.cuh"

namespace saga {

constexpr int MAX_BANDS = 128;
struct BandEdges { int e[MAX_BANDS + 1]; };

// one warp per frame; band b = mean of bins [e[b], e[b+1]) (np.mean of the slice), times scale
__global__ void __launch_bounds__(256) compress_bands_kernel(const float* __restrict__ mag, float* __restrict__ out,
                                                             const BandEdges edges, int n_bands, int n_frames,
                                                             int64_t P, int64_t clip_stride, int64_t out_P,
                                                             int64_t out_clip_stride, const float* __restrict__ inv_scale) {
  const int clip = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + warp;
  if (t >= n_frames) return;
  const float* row = mag + clip * clip_stride + (int64_t)t * P;
  float* orow = out + clip * out_clip_stride + (int64_t)t * out_P;
  const float sc = inv_scale ? inv_scale[clip] : 1.0f;
  for (int b = 0; b < n_bands; ++b) {
    const int k0 = edges.e[b], k1 = edges.e[b + 1];
    float s = 0.f;
    for (int k = k0 + lane; k < k1; k += 32) s += row[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) orow[b] = (k1 > k0) ? (s / (float)(k1 - k0)) * sc : __int_as_float(0x7fc00000);  // mean of empty slice = nan
  }
  for (int64_t k = n_bands + lane; k < out_P; k += 32) orow[k] = 0.f;
}

// Short-window gather for one window: output column j takes source frame src_frames[j] (-1 = zeros,
// the t == 0 branch of _resize); rows are bins [band_min, band_min + n_rows) zero-padded past n_bins.
//   out_lin  = mag / ref                        (training.py:352, :361)
//   out_log  = log10(1000 mag + 1) / max(.)     (training.py:350-351, :359-360)
//   out_phase = (angle(ph) + 3.15) / 6.3        (training.py:362-363)
// One CTA per window (blockIdx.x): window w reads mag + w * clip_stride, its own source frames
// src_frames[w * n_cols ..], band_min_dev[w] / inv_ref_dev[w] when given (else the scalar arguments), and writes
// out_* + w * out_clip_stride.
__global__ void __launch_bounds__(256) short_window_kernel(const float* __restrict__ mag, const float2* __restrict__ ph,
                                                           const int* __restrict__ src_frames, int n_cols, int band_min,
                                                           int n_rows, int n_bins, int64_t P, float inv_ref,
                                                           float* __restrict__ out_lin, float* __restrict__ out_log,
                                                           float* __restrict__ out_phase, int64_t out_P,
                                                           int64_t clip_stride = 0, int64_t out_clip_stride = 0,
                                                           const int* __restrict__ band_min_dev = nullptr,
                                                           const float* __restrict__ inv_ref_dev = nullptr) {
  __shared__ float red[8];
  const int w = blockIdx.x;
  mag += (int64_t)w * clip_stride;
  if (ph) ph += (int64_t)w * clip_stride;
  src_frames += (int64_t)w * n_cols;
  if (out_lin) out_lin += (int64_t)w * out_clip_stride;
  if (out_log) out_log += (int64_t)w * out_clip_stride;
  if (out_phase) out_phase += (int64_t)w * out_clip_stride;
  if (band_min_dev) band_min = band_min_dev[w];
  if (inv_ref_dev) inv_ref = inv_ref_dev[w];
  const int n = n_rows * n_cols;
  float vmax = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int j = i / n_rows, r = i % n_rows;          // output is frame-major: column j, row r
    const int k = band_min + r;
    const int t = src_frames[j];
    const bool in = (t >= 0 && k < n_bins && k >= 0);
    const float m = in ? mag[(int64_t)t * P + k] : 0.f;
    if (out_lin) out_lin[(int64_t)j * out_P + r] = m * inv_ref;
    if (out_log) {
      const float l = log10f(fmaf(m, 1000.0f, 1.0f));
      out_log[(int64_t)j * out_P + r] = l;
      vmax = fmaxf(vmax, l);
    }
    if (out_phase) {
      float a = 0.f;
      if (in && ph) {
        const float2 z = ph[(int64_t)t * P + k];
        a = atan2f(z.y, z.x);
      }
      out_phase[(int64_t)j * out_P + r] = (a + 3.15f) / 6.3f;
    }
  }
  if (!out_log) return;
  vmax = warp_max(vmax);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = vmax;
  __syncthreads();
  float m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __threadfence_block();
  for (int i = threadIdx.x; i < n; i += 256) {
    const int j = i / n_rows, r = i % n_rows;
    out_log[(int64_t)j * out_P + r] /= m;              // 0/0 = nan for an all-zero tile, as in numpy
  }
}

// flatness[t] = exp(mean_k log(max(amin, S^2))) / mean_k max(amin, S^2); one warp per frame (float accumulators
// in double for the means: librosa computes this in the STFT's float32/float64 mix, tolerance 1e-4)
__global__ void __launch_bounds__(256) flatness_kernel(const float* __restrict__ mag, float* __restrict__ out, int n_bins,
                                                       int n_frames, int64_t P, int64_t clip_stride, float amin) {
  const int clip = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + warp;
  if (t >= n_frames) return;
  const float* row = mag + clip * clip_stride + (int64_t)t * P;
  double slog = 0.0, slin = 0.0;
  for (int k = lane; k < n_bins; k += 32) {
    const float p = fmaxf(amin, row[k] * row[k]);
    slog += (double)logf(p);
    slin += (double)p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    slog += __shfl_xor_sync(0xffffffffu, slog, o);
    slin += __shfl_xor_sync(0xffffffffu, slin, o);
  }
  if (lane == 0) out[(int64_t)clip * n_frames + t] = (float)(exp(slog / n_bins) / (slin / n_bins));
}

// Column gather of a batch of frame-major images: out[w][j][k] = in[w][src[w][j]][k] * scale[w] for k < n_bins
// (src = -1: zeros; columns k in [n_bins, out_pitch) are written as 0).  This is `C[:, s:t]` + `_resize` +
// `/ ref_C` of the producer loop's five slice_C calls (util_audio.py:431-434, :384-409; training.py:340-388) for a
// whole batch of windows.  One warp per output column.
__global__ void __launch_bounds__(256) gather_frames_kernel(const float* __restrict__ in, const int* __restrict__ src,
                                                            const float* __restrict__ scale, float* __restrict__ out,
                                                            int n_cols, int n_bins, int64_t P, int64_t clip_stride,
                                                            int64_t out_P, int64_t out_clip_stride, int n_windows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t col = (int64_t)blockIdx.x * 8 + warp;
  if (col >= (int64_t)n_windows * n_cols) return;
  const int w = (int)(col / n_cols), j = (int)(col - (int64_t)w * n_cols);
  const int t = src[col];
  const float sc = scale ? scale[w] : 1.0f;
  const float* row = in + (int64_t)w * clip_stride + (int64_t)t * P;
  float* orow = out + (int64_t)w * out_clip_stride + (int64_t)j * out_P;
  for (int k = lane; k < out_P; k += 32) orow[k] = (t >= 0 && k < n_bins) ? row[k] * sc : 0.f;
}

}  // namespace saga

using namespace saga;

extern "C" int saga_gather_frames_exec(const float* in, const int32_t* src_frames, const float* scale, float* out,
                                       int n_windows, int n_cols, int n_bins, int64_t frame_pitch, int64_t clip_stride,
                                       int64_t out_pitch, int64_t out_clip_stride, void* stream) {
  if (!in || !src_frames || !out) return set_error(SAGA_ERR_INVALID, "gather_frames_exec: null argument");
  if (n_cols < 1 || n_bins < 1 || out_pitch < n_bins || frame_pitch < n_bins)
    return set_error(SAGA_ERR_INVALID, "gather_frames_exec: bad shape");
  if (n_windows <= 0) return SAGA_OK;
  const int64_t cols = (int64_t)n_windows * n_cols;
  gather_frames_kernel<<<(unsigned)((cols + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      in, src_frames, scale, out, n_cols, n_bins, frame_pitch, clip_stride, out_pitch, out_clip_stride, n_windows);
  SAGA_LAUNCH_CHECK();
  return SAGA_OK;
}

extern "C" int saga_short_window_batch_exec(const float* mag, const void* phase, int64_t clip_stride, int64_t frame_pitch,
                                            const int32_t* src_frames, int n_cols, const int32_t* band_min, int band_min_all,
                                            int n_rows, int n_bins, const float* inv_ref, float inv_ref_all,
                                            float* out_lin, float* out_log, float* out_phase, int64_t out_pitch,
                                            int64_t out_clip_stride, int n_windows, void* stream) {
  if (!mag || !src_frames) return set_error(SAGA_ERR_INVALID, "short_window_batch_exec: null argument");
  if (n_cols < 1 || n_rows < 1 || out_pitch < n_rows) return set_error(SAGA_ERR_INVALID, "short_window_batch_exec: bad shape");
  if (n_windows <= 0) return SAGA_OK;
  short_window_kernel<<<n_windows, 256, 0, (cudaStream_t)stream>>>(
      mag, (const float2*)phase, src_frames, n_cols, band_min_all, n_rows, n_bins, frame_pitch, inv_ref_all, out_lin,
      out_log, out_phase, out_pitch, clip_stride, out_clip_stride, band_min, inv_ref);
  SAGA_LAUNCH_CHECK();
  return SAGA_OK;
}

extern "C" int saga_compress_bands_exec(const float* mag, float* out, const int32_t* band_edges_host, int n_bands,
                                        int n_clips, int n_frames, int64_t frame_pitch, int64_t clip_stride,
                                        int64_t out_pitch, int64_t out_clip_stride, const float* inv_scale,
                                        void* stream) {
  if (!mag || !out || !band_edges_host) return set_error(SAGA_ERR_INVALID, "compress_bands_exec: null argument");
  if (n_bands < 1 || n_bands > MAX_BANDS) return set_error(SAGA_ERR_UNSUPPORTED, "compress_bands_exec: 1..%d bands", MAX_BANDS);
  if (out_pitch < n_bands) return set_error(SAGA_ERR_INVALID, "compress_bands_exec: out_pitch < n_bands");
  BandEdges e;
  for (int i = 0; i <= n_bands; ++i) {
    e.e[i] = band_edges_host[i];
    if (e.e[i] < 0 || e.e[i] > frame_pitch || (i && e.e[i] < e.e[i - 1]))
      return set_error(SAGA_ERR_INVALID, "compress_bands_exec: band edges must be non-decreasing within the row");
  }
  if (n_clips <= 0 || n_frames <= 0) return SAGA_OK;
  dim3 grid((n_frames + 7) / 8, n_clips);
  compress_bands_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mag, out, e, n_bands, n_frames, frame_pitch, clip_stride,
                                                                out_pitch, out_clip_stride, inv_scale);
  SAGA_LAUNCH_CHECK();
  return SAGA_OK;
}

extern "C" int saga_short_window_exec(const float* mag, const void* phase, const int32_t* src_frames, int n_cols,
                                      int band_min, int n_rows, int n_bins, int64_t frame_pitch, float inv_ref,
                                      float* out_lin, float* out_log, float* out_phase, int64_t out_pitch, void* stream) {
  if (!mag || !src_frames) return set_error(SAGA_ERR_INVALID, "short_window_exec: null argument");
  if (n_cols < 1 || n_rows < 1 || out_pitch < n_rows) return set_error(SAGA_ERR_INVALID, "short_window_exec: bad shape");
  short_window_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(mag, (const float2*)phase, src_frames, n_cols, band_min, n_rows,
                                                           n_bins, frame_pitch, inv_ref, out_lin, out_log, out_phase,
                                                           out_pitch);
  SAGA_LAUNCH_CHECK();
  return SAGA_OK;
}

extern "C" int saga_spectral_flatness_exec(const float* mag, float* flatness_out, int n_clips, int n_bins, int n_frames,
                                           int64_t frame_pitch, int64_t clip_stride, float amin, void* stream) {
  if (!mag || !flatness_out) return set_error(SAGA_ERR_INVALID, "spectral_flatness_exec: null argument");
  if (n_clips <= 0 || n_frames <= 0) return SAGA_OK;
  dim3 grid((n_frames + 7) / 8, n_clips);
  flatness_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mag, flatness_out, n_bins, n_frames, frame_pitch, clip_stride, amin);
  SAGA_LAUNCH_CHECK();
  return SAGA_OK;
}

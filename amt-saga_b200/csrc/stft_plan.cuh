// Plan object shared by the STFT translation units (stft.cu: generic kernels + C ABI; stft_ring.cu: the
// ring-buffered n_fft 2048 / hop 512 kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct saga_stft_plan {
  int n_fft, hop, center, M;
  int n_pass;
  int radix[3];
  float* d_window;   // n_fft floats
  float* d_window_half;  // 0.5 * window (forward kernel)
  float2* d_tw[3];   // per pass: [R][L/R] inter-pass twiddles (last pass: unused)
  float2* d_twN;     // M/2 + 1 entries exp(-2*pi*i*k/n_fft)
  float2* d_tw_eo;   // n_fft 4096 only: [32][32] exp(-2*pi*i*rp*j/1024), first-pass twiddles of the two 1024-point halves
  int warps;         // warps per CTA
  int frames_per_cta;
  int span_alloc;    // floats reserved for the staged span
  size_t smem_bytes;
  int default_window;     // periodic Hann (w[n + N/2] = 1 - w[n]): the ring kernel's fused first stage relies on it
  float2* d_ring_tables;  // shared-memory image of the ring kernel's tables (stft_ring.cu), or NULL
  float2* d_iring_tables; // the same for the inverse ring kernel (istft_ring.cu), or NULL
  float* d_wsq;           // window^2 (inverse ring kernel, clip-edge blocks)
};

namespace saga {

struct StftArgs {
  const float* wav;
  const int64_t* clip_offsets;
  const int64_t* clip_lens;
  float* mag_out;
  float2* phase_out;
  float2* cplx_out;
  float* frame_max_out;
  float* clip_max_out;
  const float* window;
  const float2* tw0;
  const float2* tw1;
  const float2* twN;
  const float2* tw_eo;
  int64_t frame_pitch, out_clip_stride;
  int hop, center, frames_per_cta, span_alloc, tiles_per_clip, max_frames;
};

// stft_ring.cu
bool stft_ring_supported(const saga_stft_plan* p);
int stft_ring_build_tables(saga_stft_plan* p);
int launch_stft_ring(const saga_stft_plan* p, const StftArgs& a, int n_clips, int64_t max_frames, cudaStream_t st);

// istft_ring.cu
bool istft_ring_supported(const saga_stft_plan* p);
int istft_ring_build_tables(saga_stft_plan* p);
int launch_istft_ring(const saga_stft_plan* p, const void* cplx_in, const float* mag_in, const void* phase_in,
                      int n_clips, int n_frames, int64_t frame_pitch, int64_t in_clip_stride, float* wav_out,
                      int64_t wav_clip_stride, cudaStream_t st);

}  // namespace saga

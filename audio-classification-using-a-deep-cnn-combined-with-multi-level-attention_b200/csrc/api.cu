// extern "C" boundary (include/vggish_mla_b200.h): argument checks, error strings, dispatch into the kernels.
#include "../../include/vggish_mla_b200.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "igemm_sm100.cuh"
#include "kernels.cuh"

namespace {
thread_local char g_api_err[768] = "";

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_api_err, sizeof g_api_err, fmt, ap);
  va_end(ap);
  return 1;
}
int fail_from(const char* where, const char* inner) { return fail("%s: %s", where, inner); }

cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
}  // namespace

namespace vmb {
void set_api_error(const char* msg) { snprintf(g_api_err, sizeof g_api_err, "%s", msg); }
}  // namespace vmb

extern "C" {

const char* vmb_last_error(void) { return g_api_err; }
int vmb_abi_version(void) { return 2; }

int vmb_device_arch(int device) {
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess) {
    fail("vmb_device_arch: no CUDA device %d (%s)", device, cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  return major * 10 + minor;
}

int vmb_igemm_pair_enable(int on) { return vmb::igemm_set_pair(on); }
int vmb_igemm_halo_enable(int on) { return vmb::igemm_set_halo(on); }

long long vmb_num_frames(long long n_samples) {
  // mel_features.py:42  1 + int(floor((num_samples - window_length) / hop_length)), floor division
  const long long d = n_samples - 400;
  long long q = d / 160;
  if (d % 160 != 0 && d < 0) --q;
  return 1 + q;
}

long long vmb_num_examples(long long n_samples) {
  const long long f = vmb_num_frames(n_samples);
  if (f < 96) return f < 0 ? -1 : 0;
  return 1 + (f - 96) / 96;  // vggish_input.py:73-76 via mel_features.py:42 with window = hop = 96
}

static int logmel_common(const char* who, bool tensor_core, const float* wave, long long n_clips,
                         long long samples_per_clip, long long clip_stride, long long frames_out, float* logmel,
                         void* stream) {
  if (n_clips < 0 || frames_out < 0) return fail("%s: negative size", who);
  const long long nf = vmb_num_frames(samples_per_clip);
  if (nf < 1) return fail("%s: %lld samples is shorter than one 400-sample window", who, samples_per_clip);
  if (frames_out > nf) return fail("%s: frames_out %lld > available frames %lld", who, frames_out, nf);
  if (clip_stride < samples_per_clip) return fail("%s: clip_stride < samples_per_clip", who);
  if (n_clips == 0 || frames_out == 0) return 0;
  if (!wave || !logmel) return fail("%s: null pointer", who);
  const int rc = tensor_core
                     ? vmb::logmel_tc_forward(wave, n_clips, samples_per_clip, clip_stride, frames_out, logmel, S(stream))
                     : vmb::logmel_forward(wave, n_clips, samples_per_clip, clip_stride, frames_out, logmel, S(stream));
  if (rc) return fail_from(who, vmb::kernels_last_error());
  return 0;
}

int vmb_logmel(const float* wave, long long n_clips, long long samples_per_clip, long long clip_stride,
               long long frames_out, float* logmel, void* stream) {
  return logmel_common("vmb_logmel", true, wave, n_clips, samples_per_clip, clip_stride, frames_out, logmel, stream);
}

int vmb_logmel_pcm16(const int16_t* pcm, long long n_clips, long long samples_per_clip, long long clip_stride,
                     long long frames_out, float* logmel, void* stream) {
  const char* who = "vmb_logmel_pcm16";
  if (n_clips < 0 || frames_out < 0) return fail("%s: negative size", who);
  const long long nf = vmb_num_frames(samples_per_clip);
  if (nf < 1) return fail("%s: %lld samples is shorter than one 400-sample window", who, samples_per_clip);
  if (frames_out > nf) return fail("%s: frames_out %lld > available frames %lld", who, frames_out, nf);
  if (clip_stride < samples_per_clip) return fail("%s: clip_stride < samples_per_clip", who);
  if (n_clips == 0 || frames_out == 0) return 0;
  if (!pcm || !logmel) return fail("%s: null pointer", who);
  if (vmb::logmel_tc_forward_pcm16(pcm, n_clips, samples_per_clip, clip_stride, frames_out, logmel, S(stream)))
    return fail_from(who, vmb::kernels_last_error());
  return 0;
}

int vmb_logmel_cudacore(const float* wave, long long n_clips, long long samples_per_clip, long long clip_stride,
                        long long frames_out, float* logmel, void* stream) {
  return logmel_common("vmb_logmel_cudacore", false, wave, n_clips, samples_per_clip, clip_stride, frames_out, logmel,
                       stream);
}

int vmb_stft_magnitude(const double* signal, long long n_samples, double* mag, void* stream) {
  if (n_samples < 0) return fail("vmb_stft_magnitude: negative size");
  if (vmb_num_frames(n_samples) < 1) return fail("vmb_stft_magnitude: %lld samples is shorter than one 400-sample window", n_samples);
  if (!signal || !mag) return fail("vmb_stft_magnitude: null pointer");
  if (vmb::stft_magnitude_f64(signal, n_samples, mag, S(stream))) return fail_from("vmb_stft_magnitude", vmb::kernels_last_error());
  return 0;
}

int vmb_spec_tiles(const float* examples, long long n_clips, int n_examples_per_clip, int n_frames, int overlap,
                   float* out, void* stream) {
  if (n_clips < 0) return fail("vmb_spec_tiles: negative n_clips");
  if (n_examples_per_clip < 0 || n_examples_per_clip > 4) return fail("vmb_spec_tiles: 0..4 examples per clip (a 4 s clip)");
  if (overlap ? n_frames < 4 : (n_frames < 1 || n_frames > 4))
    return fail("vmb_spec_tiles: overlap needs >= 4 frames, contiguous tiling at most 4 (dataset.py:169-172)");
  if (n_clips == 0) return 0;
  if ((!examples && n_examples_per_clip) || !out) return fail("vmb_spec_tiles: null pointer");
  const int step = overlap ? (384 - 96) / (n_frames - 1) : 96;   // dataset.py:356-359, :362-363
  if (vmb::spec_tiles(examples, n_clips, n_examples_per_clip, n_frames, step, out, S(stream)))
    return fail_from("vmb_spec_tiles", vmb::kernels_last_error());
  return 0;
}

int vmb_front_end_tables(double* hann400, double* mel257x64) {
  vmb::front_end_tables_host(hann400, mel257x64);
  return 0;
}

static int conv1_common(const char* who, bool tensor_core, const float* examples, const float* w, const float* b,
                        void* out, long long n, void* stream, int dtype = 0) {
  if (n < 0) return fail("%s: negative n", who);
  if (dtype != 0 && dtype != 1) return fail("%s: dtype must be 0 (bf16) or 1 (fp16)", who);
  if (n == 0) return 0;
  if (!examples || !w || !b || !out) return fail("%s: null pointer", who);
  const int rc = tensor_core ? vmb::conv1_tc_relu_pool(examples, w, b, out, n, S(stream), false, dtype)
                             : vmb::conv1_relu_pool(examples, w, b, out, n, S(stream));
  if (rc) return fail_from(who, vmb::kernels_last_error());
  return 0;
}

int vmb_conv1_relu_pool(const float* examples, const float* w, const float* b, void* out, long long n, void* stream) {
  return conv1_common("vmb_conv1_relu_pool", true, examples, w, b, out, n, stream);
}

int vmb_conv1_relu_pool_ex(const float* examples, const float* w, const float* b, void* out, long long n, int dtype,
                           void* stream) {
  return conv1_common("vmb_conv1_relu_pool_ex", true, examples, w, b, out, n, stream, dtype);
}

int vmb_conv1_relu_pool_cudacore(const float* examples, const float* w, const float* b, void* out, long long n,
                                 void* stream) {
  return conv1_common("vmb_conv1_relu_pool_cudacore", false, examples, w, b, out, n, stream);
}

int vmb_conv3x3_relu_ex(const void* act, const void* w, const float* bias, void* out, long long n, int H, int W,
                        int C_in, int C_out, int pool, int dtype, void* stream) {
  if (n < 0 || n > 0x7fffffffLL / 4096) return fail("vmb_conv3x3_relu: bad n %lld", n);
  if (dtype != 0 && dtype != 1) return fail("vmb_conv3x3_relu: dtype must be 0 (bf16) or 1 (fp16)");
  if (n == 0) return 0;
  if (!act || !w || !bias || !out) return fail("vmb_conv3x3_relu: null pointer");
  if (pool && ((H | W) & 1)) return fail("vmb_conv3x3_relu: pooling needs even H, W");
  if (vmb::igemm_conv3x3(act, w, bias, out, int(n), H, W, C_in, C_out, pool, S(stream), dtype))
    return fail_from("vmb_conv3x3_relu", vmb::igemm_last_error());
  return 0;
}

int vmb_conv3x3_relu(const void* act, const void* w, const float* bias, void* out, long long n, int H, int W, int C_in,
                     int C_out, int pool, void* stream) {
  return vmb_conv3x3_relu_ex(act, w, bias, out, n, H, W, C_in, C_out, pool, 0, stream);
}

int vmb_linear_ex(const void* a, const void* w, const float* bias, void* out, int out_f32, int relu, long long M, int N,
                  int K, int dtype, void* stream) {
  if (M < 0 || M > 0x7fffffffLL) return fail("vmb_linear: bad M %lld", M);
  if (dtype != 0 && dtype != 1) return fail("vmb_linear: dtype must be 0 (bf16) or 1 (fp16)");
  if (M == 0) return 0;
  if (!a || !w || !bias || !out) return fail("vmb_linear: null pointer");
  if (vmb::igemm_linear(a, w, bias, out, out_f32, relu, int(M), N, K, S(stream), dtype))
    return fail_from("vmb_linear", vmb::igemm_last_error());
  return 0;
}

int vmb_linear(const void* a, const void* w, const float* bias, void* out, int out_f32, int relu, long long M, int N,
               int K, void* stream) {
  return vmb_linear_ex(a, w, bias, out, out_f32, relu, M, N, K, 0, stream);
}

int vmb_postprocess(const float* emb, const float* eigen, const float* means, float* out_f32, uint8_t* out_u8,
                    long long n, void* stream) {
  if (n < 0) return fail("vmb_postprocess: negative n");
  if (n == 0) return 0;
  if (!emb || !eigen || !means) return fail("vmb_postprocess: null pointer");
  if (vmb::postprocess(emb, eigen, means, out_f32, out_u8, n, S(stream)))
    return fail_from("vmb_postprocess", vmb::kernels_last_error());
  return 0;
}

}  // extern "C"

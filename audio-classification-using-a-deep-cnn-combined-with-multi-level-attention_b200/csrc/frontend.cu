// Log-mel front end: framing + periodic Hann + 512-point real DFT magnitude + HTK mel projection + log.
// Replaces torchvggish/mel_features.py:21-45, 48-68, 71-92, 114-189, 192-223 (float64 numpy) with one fused
// kernel.  Frame i of a clip starts at sample 160*i (mel_features.py:42-45); nothing is padded.
//
// Only DFT bins kBinLo .. kBinLo+kNumBins-1 are evaluated: the 257x64 HTK mel matrix is exactly zero outside
// bins 5..239 for (16 kHz, 125 Hz, 7500 Hz) — checked at table-build time — so the other bins cannot reach
// the output (mel_features.py:155-189).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

#include "kernels.cuh"

namespace vmb {

namespace {
thread_local char g_kerr[512] = "";
}
const char* kernels_last_error() { return g_kerr; }
void set_kernel_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_kerr, sizeof g_kerr, fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_kernel_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

namespace {

constexpr int kWin = 400;      // int(round(16000 * 0.025))          mel_features.py:212
constexpr int kHop = 160;      // int(round(16000 * 0.010))          mel_features.py:213
constexpr int kFft = 512;      // 2 ** ceil(log2(400))               mel_features.py:214
constexpr int kBins = 257;     // fft/2 + 1
constexpr int kMel = 64;       // vggish_params.NUM_MEL_BINS
constexpr int kBinLo = 4;      // first evaluated DFT bin
constexpr int kNumBins = 236;  // evaluated bins 4..239
constexpr double kLogOffset = 0.01;
constexpr double kPi = 3.14159265358979323846;

// ------------------------------------------------------------------ host tables (float64, as numpy builds them)
void hann_host(double* w) {
  // 0.5 - (0.5 * np.cos(2 * np.pi / window_length * np.arange(window_length)))   mel_features.py:67-68
  const double step = 2 * kPi / kWin;
  for (int n = 0; n < kWin; ++n) w[n] = 0.5 - (0.5 * std::cos(step * n));
}

double hz_to_mel(double hz) { return 1127.0 * std::log(1.0 + (hz / 700.0)); }  // mel_features.py:110-111

void linspace(double start, double stop, int num, double* out) {
  // numpy.linspace(endpoint=True): arange(num) * step + start, last element forced to `stop`
  const double step = (stop - start) / (num - 1);
  for (int i = 0; i < num; ++i) out[i] = i * step + start;
  out[num - 1] = stop;
}

void mel_matrix_host(double* m /*[257][64]*/) {
  // spectrogram_to_mel_matrix(64, 257, 16000, 125, 7500)   mel_features.py:155-189
  const double nyquist = 16000 / 2.;
  std::vector<double> bins_hz(kBins), bins_mel(kBins), edges(kMel + 2);
  linspace(0.0, nyquist, kBins, bins_hz.data());
  for (int i = 0; i < kBins; ++i) bins_mel[i] = hz_to_mel(bins_hz[i]);
  linspace(hz_to_mel(125.0), hz_to_mel(7500.0), kMel + 2, edges.data());
  for (int i = 0; i < kMel; ++i) {
    const double lo = edges[i], ce = edges[i + 1], up = edges[i + 2];
    for (int k = 0; k < kBins; ++k) {
      const double ls = (bins_mel[k] - lo) / (ce - lo);
      const double us = (up - bins_mel[k]) / (up - ce);
      m[k * kMel + i] = std::fmax(0.0, std::fmin(ls, us));
    }
  }
  for (int i = 0; i < kMel; ++i) m[i] = 0.0;  // DC row, mel_features.py:188
}

struct DeviceTables {
  float* basis = nullptr;  // [400][2][kNumBins]: hann[n]*cos(2 pi k n / 512), hann[n]*sin(..), k = kBinLo..
  float* melw = nullptr;   // [kNumBins][64]
};

std::mutex g_tab_mu;
DeviceTables g_tabs[64];

int get_tables(DeviceTables** out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_kernel_error("front end: cudaGetDevice failed");
    return 1;
  }
  std::lock_guard<std::mutex> lk(g_tab_mu);
  DeviceTables& t = g_tabs[dev];
  if (!t.basis) {
    std::vector<double> hann(kWin), mel(kBins * kMel);
    hann_host(hann.data());
    mel_matrix_host(mel.data());
    for (int k = 0; k < kBins; ++k)
      if (k < kBinLo || k >= kBinLo + kNumBins)
        for (int i = 0; i < kMel; ++i)
          if (mel[k * kMel + i] != 0.0) {
            set_kernel_error("front end: mel matrix has weight outside the evaluated bin range (bin %d)", k);
            return 1;
          }
    std::vector<float> basis(size_t(kWin) * 2 * kNumBins), melw(size_t(kNumBins) * kMel);
    for (int n = 0; n < kWin; ++n)
      for (int j = 0; j < kNumBins; ++j) {
        // exact argument reduction: (k*n) mod 512 before scaling by 2 pi / 512
        const int kn = ((kBinLo + j) * n) % kFft;
        const double ang = 2 * kPi * kn / kFft;
        basis[(size_t(n) * 2 + 0) * kNumBins + j] = float(hann[n] * std::cos(ang));
        basis[(size_t(n) * 2 + 1) * kNumBins + j] = float(hann[n] * std::sin(ang));
      }
    for (int j = 0; j < kNumBins; ++j)
      for (int i = 0; i < kMel; ++i) melw[size_t(j) * kMel + i] = float(mel[(kBinLo + j) * kMel + i]);
    float *db = nullptr, *dm = nullptr;
    if (cudaMalloc(&db, basis.size() * 4) != cudaSuccess || cudaMalloc(&dm, melw.size() * 4) != cudaSuccess ||
        cudaMemcpy(db, basis.data(), basis.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(dm, melw.data(), melw.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
      set_kernel_error("front end: table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
      return 1;
    }
    t.basis = db;
    t.melw = dm;
  }
  *out = &t;
  return 0;
}

// ------------------------------------------------------------------ kernel
constexpr int kFT = 32;                          // frames per CTA
constexpr int kSeg = (kFT - 1) * kHop + kWin;    // samples a CTA touches (5360)
constexpr int kBinGroups = kNumBins / 4;         // 59 groups of 4 bins
constexpr int kFrameGroups = 4;                  // x 8 frames
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
logmel_kernel(const float* __restrict__ wave, long long n_samples, long long clip_stride, long long frames_out,
              const float* __restrict__ basis, const float* __restrict__ melw, float* __restrict__ out) {
  // one buffer, two lives: the waveform segment during the DFT, the magnitudes afterwards
  __shared__ float buf[kFT * (kNumBins + 1)];
  static_assert(kFT * (kNumBins + 1) >= kSeg, "segment must fit in the magnitude buffer");
  float* xs = buf;
  float (*mag)[kNumBins + 1] = reinterpret_cast<float (*)[kNumBins + 1]>(buf);
  const int tid = threadIdx.x;
  const long long clip = blockIdx.y;
  const long long f0 = static_cast<long long>(blockIdx.x) * kFT;
  const float* src = wave + clip * clip_stride;
  const long long s0 = f0 * kHop;
  for (int i = tid; i < kSeg; i += kThreads) xs[i] = (s0 + i < n_samples) ? __ldg(src + s0 + i) : 0.f;
  __syncthreads();

  const bool dft_thread = tid < kBinGroups * kFrameGroups;
  const int bg = tid % kBinGroups, fg = tid / kBinGroups;
  float re[8][4], im[8][4];
  if (dft_thread) {
#pragma unroll
    for (int f = 0; f < 8; ++f)
#pragma unroll
      for (int j = 0; j < 4; ++j) re[f][j] = im[f][j] = 0.f;
    const float* xp = xs + fg * 8 * kHop;
    const float4* bc = reinterpret_cast<const float4*>(basis) + bg;
    const float4* bs = reinterpret_cast<const float4*>(basis + kNumBins) + bg;
#pragma unroll 2
    for (int n = 0; n < kWin; ++n) {
      const float4 c = __ldg(bc + size_t(n) * (2 * kNumBins / 4));
      const float4 s = __ldg(bs + size_t(n) * (2 * kNumBins / 4));
#pragma unroll
      for (int f = 0; f < 8; ++f) {
        const float x = xp[f * kHop + n];
        re[f][0] = fmaf(x, c.x, re[f][0]);
        re[f][1] = fmaf(x, c.y, re[f][1]);
        re[f][2] = fmaf(x, c.z, re[f][2]);
        re[f][3] = fmaf(x, c.w, re[f][3]);
        im[f][0] = fmaf(x, s.x, im[f][0]);
        im[f][1] = fmaf(x, s.y, im[f][1]);
        im[f][2] = fmaf(x, s.z, im[f][2]);
        im[f][3] = fmaf(x, s.w, im[f][3]);
      }
    }
  }
  __syncthreads();  // every thread is done reading xs before it is overwritten with magnitudes
  if (dft_thread) {
#pragma unroll
    for (int f = 0; f < 8; ++f)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        mag[fg * 8 + f][bg * 4 + j] = sqrtf(fmaf(re[f][j], re[f][j], im[f][j] * im[f][j]));
  }
  __syncthreads();

  const int m = tid & 63, fq = tid >> 6;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int k = 0; k < kNumBins; ++k) {
    const float w = __ldg(melw + k * kMel + m);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(mag[fq + 4 * i][k], w, acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long f = f0 + fq + 4 * i;
    if (f < frames_out) out[(clip * frames_out + f) * kMel + m] = logf(acc[i] + float(kLogOffset));
  }
}

}  // namespace

void front_end_tables_host(double* hann400, double* mel257x64) {
  if (hann400) hann_host(hann400);
  if (mel257x64) mel_matrix_host(mel257x64);
}

int logmel_forward(const float* wave, long long n_clips, long long samples_per_clip, long long clip_stride,
                   long long frames_out, float* logmel, cudaStream_t stream) {
  DeviceTables* t = nullptr;
  if (get_tables(&t)) return 1;
  const long long tiles = (frames_out + kFT - 1) / kFT;
  if (tiles > 0x7fffffffLL || n_clips > 65535) {
    // grid.y limit: callers batch clips in chunks (the host layer does); keep the kernel simple.
    set_kernel_error("logmel_forward: n_clips %lld > 65535 per launch or too many frame tiles", n_clips);
    return 1;
  }
  dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(n_clips));
  logmel_kernel<<<grid, kThreads, 0, stream>>>(wave, samples_per_clip, clip_stride, frames_out, t->basis, t->melw,
                                               logmel);
  count_launch();
  return check_launch("logmel_kernel");
}

}  // namespace vmb

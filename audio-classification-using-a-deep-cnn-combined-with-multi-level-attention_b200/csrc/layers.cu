// VGGish layers that are not GEMM-shaped: the C_in = 1 first conv (K = 9, CUDA cores), the PCA/quantise
// postprocessor, and the one-time weight re-layout kernels.
#include <cuda_bf16.h>

#include "kernels.cuh"
#include "sm100_ptx.cuh"

namespace vmb {

namespace {

// ------------------------------------------------------------------ conv1 + ReLU + 2x2 maxpool
// features.0/1/2 of make_layers() (vggish.py:108-118): Conv2d(1, 64, 3, padding=1) -> ReLU -> MaxPool2d(2, 2).
// One CTA = 8 pooled rows of one example (16 input rows + halo), one thread = one pooled pixel, all 64 channels.
constexpr int kInH = 96, kInW = 64, kC1 = 64;
constexpr int kPRows = 8;                 // pooled rows per CTA
constexpr int kTileH = 2 * kPRows + 2;    // 18 input rows with halo
constexpr int kTileW = kInW + 2;          // 66

__global__ void __launch_bounds__(256)
conv1_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
             __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[kTileH][kTileW + 1];
  __shared__ float ws[kC1 * 9];
  __shared__ float bs[kC1];
  const int tid = threadIdx.x;
  const long long n = blockIdx.y;
  const int pr0 = blockIdx.x * kPRows;  // first pooled row
  const int r0 = 2 * pr0 - 1;           // first input row in the tile (may be -1)
  const float* src = x + n * (kInH * kInW);
  for (int i = tid; i < kTileH * kTileW; i += 256) {
    const int r = i / kTileW, c = i - r * kTileW;
    const int gr = r0 + r, gc = c - 1;
    tile[r][c] = (gr >= 0 && gr < kInH && gc >= 0 && gc < kInW) ? __ldg(src + gr * kInW + gc) : 0.f;
  }
  for (int i = tid; i < kC1 * 9; i += 256) ws[i] = __ldg(w + i);
  if (tid < kC1) bs[tid] = __ldg(b + tid);
  __syncthreads();

  const int pw = tid & 31, ph = tid >> 5;  // pooled pixel inside the CTA tile
  float p[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) p[i][j] = tile[2 * ph + i][2 * pw + j];

  __nv_bfloat16* dst = out + ((n * (kInH / 2) + pr0 + ph) * (kInW / 2) + pw) * kC1;
#pragma unroll 1
  for (int c0 = 0; c0 < kC1; c0 += 8) {
    uint32_t pk[4];
#pragma unroll
    for (int cc = 0; cc < 8; cc += 2) {
      float r2[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float* wc = ws + (c0 + cc + u) * 9;
        float m = -INFINITY;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            float a = bs[c0 + cc + u];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) a = fmaf(p[dy + ky][dx + kx], wc[ky * 3 + kx], a);
            m = fmaxf(m, a);
          }
        r2[u] = fmaxf(m, 0.f);
      }
      pk[cc >> 1] = pack_bf16x2(r2[0], r2[1]);
    }
    *reinterpret_cast<uint4*>(dst + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ------------------------------------------------------------------ postprocessor (vggish.py:87-102)
// out[j] = round((clamp(sum_i E[j][i] * (x[i] - mu[i]), -2, 2) + 2) * 63.75), round-half-even like torch.round.
__global__ void __launch_bounds__(128)
postprocess_kernel(const float* __restrict__ emb, const float* __restrict__ eig, const float* __restrict__ mu,
                   float* __restrict__ out_f32, uint8_t* __restrict__ out_u8, long long n) {
  constexpr int kRows = 8;
  __shared__ float xc[kRows][128];
  const int j = threadIdx.x;
  const long long r0 = static_cast<long long>(blockIdx.x) * kRows;
  const float m = __ldg(mu + j);
#pragma unroll
  for (int r = 0; r < kRows; ++r) xc[r][j] = (r0 + r < n) ? __ldg(emb + (r0 + r) * 128 + j) - m : 0.f;
  __syncthreads();
  float acc[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) acc[r] = 0.f;
  const float4* e4 = reinterpret_cast<const float4*>(eig + j * 128);
  for (int i = 0; i < 32; ++i) {
    const float4 e = __ldg(e4 + i);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const float4 xv = *reinterpret_cast<const float4*>(&xc[r][4 * i]);
      acc[r] = fmaf(e.x, xv.x, acc[r]);
      acc[r] = fmaf(e.y, xv.y, acc[r]);
      acc[r] = fmaf(e.z, xv.z, acc[r]);
      acc[r] = fmaf(e.w, xv.w, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    if (r0 + r >= n) break;
    const float c = fminf(fmaxf(acc[r], -2.0f), 2.0f);
    const float qv = rintf((c - (-2.0f)) * (255.0f / (2.0f - (-2.0f))));
    if (out_f32) out_f32[(r0 + r) * 128 + j] = qv;
    if (out_u8) out_u8[(r0 + r) * 128 + j] = static_cast<uint8_t>(qv);
  }
}

// ------------------------------------------------------------------ UrbanSound8K tiling (dataset.py:318-326, :355-363)
// create_spec (torchvggish branch): the <= 4 examples of a 4 s clip are laid side by side as a (64 mel, 384 time)
// spectrogram, missing examples are zero; split(): T windows of 96 time steps, `step` apart.
// out[clip][f][m][t] = spec[m][f * step + t],  spec[m][tau] = examples[clip][tau / 96][tau % 96][m].
__global__ void __launch_bounds__(256)
spec_tiles_kernel(const float* __restrict__ ex, long long n_clips, int n_ex, int T, int step, float* __restrict__ out) {
  const long long total = n_clips * T * 64 * 96;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % 96);
    const int m = static_cast<int>((i / 96) % 64);
    const int f = static_cast<int>((i / (96 * 64)) % T);
    const long long clip = i / (96LL * 64 * T);
    const int tau = f * step + t;
    const int e = tau / 96;
    out[i] = (tau < 384 && e < n_ex) ? __ldg(ex + ((clip * n_ex + e) * 96 + (tau - e * 96)) * 64 + m) : 0.f;
  }
}

// ------------------------------------------------------------------ weight re-layout
__device__ __forceinline__ uint16_t to16(float v, int f16) {
  return f16 ? static_cast<uint16_t>(pack_f16x2(v, 0.f) & 0xFFFFu) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
}

__global__ void relayout_conv_kernel(const float* __restrict__ w, uint16_t* __restrict__ o, int C_out, int C_in, int f16) {
  const long long total = static_cast<long long>(C_out) * 9 * C_in;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = int(i % C_in);
    const int tap = int((i / C_in) % 9);
    const long long oc = i / (9LL * C_in);
    o[i] = to16(__ldg(w + (oc * C_in + c) * 9 + tap), f16);
  }
}

__global__ void relayout_conv_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o, int C_out, int C_in) {
  const long long K = 9LL * C_in, total = static_cast<long long>(C_out) * K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = int(i % C_in);
    const int tap = int((i / C_in) % 9);
    const long long oc = i / K;
    const float v = __ldg(w + (oc * C_in + c) * 9 + tap);
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    o[oc * 2 * K + (i - oc * K)] = h;
    o[oc * 2 * K + K + (i - oc * K)] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

__global__ void split_planes_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, long long rows, long long cols) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    const float v = __ldg(s + i);
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    d[r * 2 * cols + c] = h;
    d[r * 2 * cols + cols + c] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ s, uint16_t* __restrict__ d, long long n, int f16) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    d[i] = to16(__ldg(s + i), f16);
}

}  // namespace

int conv1_relu_pool(const float* examples, const float* w, const float* b, void* out, long long n,
                    cudaStream_t stream) {
  for (long long done = 0; done < n;) {  // grid.y limit
    const long long chunk = (n - done) < 65535 ? (n - done) : 65535;
    dim3 grid(kInH / 2 / kPRows, static_cast<unsigned>(chunk));
    conv1_kernel<<<grid, 256, 0, stream>>>(examples + done * kInH * kInW, w, b,
                                           static_cast<__nv_bfloat16*>(out) + done * (kInH / 2) * (kInW / 2) * kC1);
    count_launch();
    if (check_launch("conv1_kernel")) return 1;
    done += chunk;
  }
  return 0;
}

int postprocess(const float* emb, const float* eigen, const float* means, float* out_f32, uint8_t* out_u8,
                long long n, cudaStream_t stream) {
  const long long blocks = (n + 7) / 8;
  postprocess_kernel<<<static_cast<unsigned>(blocks), 128, 0, stream>>>(emb, eigen, means, out_f32, out_u8, n);
  count_launch();
  return check_launch("postprocess_kernel");
}

int spec_tiles(const float* examples, long long n_clips, int n_ex, int T, int step, float* out, cudaStream_t stream) {
  const long long total = n_clips * T * 64 * 96;
  const unsigned grid = static_cast<unsigned>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  spec_tiles_kernel<<<grid, 256, 0, stream>>>(examples, n_clips, n_ex, T, step, out);
  count_launch();
  return check_launch("spec_tiles_kernel");
}

int relayout_conv_weight(const float* w_oihw, void* w_bf16, int C_out, int C_in, cudaStream_t stream, int fmt) {
  relayout_conv_kernel<<<1024, 256, 0, stream>>>(w_oihw, static_cast<uint16_t*>(w_bf16), C_out, C_in, fmt == kFmtF16);
  count_launch();
  return check_launch("relayout_conv_kernel");
}

int relayout_conv_weight_split(const float* w_oihw, void* planes, int C_out, int C_in, cudaStream_t stream) {
  relayout_conv_split_kernel<<<1024, 256, 0, stream>>>(w_oihw, static_cast<__nv_bfloat16*>(planes), C_out, C_in);
  count_launch();
  return check_launch("relayout_conv_split_kernel");
}

int split_f32_to_planes(const float* src, void* planes, long long rows, long long cols, cudaStream_t stream) {
  split_planes_kernel<<<2048, 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(planes), rows, cols);
  count_launch();
  return check_launch("split_planes_kernel");
}

int cast_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t stream, int fmt) {
  cast_bf16_kernel<<<2048, 256, 0, stream>>>(src, static_cast<uint16_t*>(dst), n, fmt == kFmtF16);
  count_launch();
  return check_launch("cast_bf16_kernel");
}

}  // namespace vmb

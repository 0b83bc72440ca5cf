// VGGish first layer on the tensor cores: Conv2d(1, 64, 3, padding=1) + ReLU + MaxPool2d(2, 2)
// (reference torchvggish/vggish.py:108-118, features.0/1/2): examples fp32 [n][96][64] -> NHWC bf16 [n][48][32][64].
//
// With C_in = 1 the GEMM K is only 9, so the operand tiles are built by hand instead of by TMA.  One CTA tile = 128
// POOLED pixels (4 pooled rows x 32 columns of one example).  Each producer thread writes ONE row: the 4x4 input
// window of its pooled pixel, split into bf16 hi / lo halves,
//     A[pixel]      = [ win_hi(16) | win_lo(16) | win_hi(16) | 1 | 1 | 0 ... ]          (64 bf16, K-major, no-swizzle
// and the 3x3 kernel is scattered, once per CTA, into four weight tiles — one per position (dy, dx) of the 2x2    core-matrix
// pooling window —                                                                                               layout)
//     B_pos[ch][..] = w_hi at slots (dy+ky)*4 + dx+kx of the first two blocks, w_lo in the third, then b_hi, b_lo,
// so 4 x 4 tcgen05.mma (M = 128, N = 64, K = 16) leave conv + bias of window position `pos` for all 64 channels in
// TMEM column block `pos` (input AND weights keep 16 mantissa bits through the hi/lo splits; the bias is added in the
// fp32 accumulator).  In the epilogue thread = pooled pixel = TMEM lane: the max over the four column blocks is the
// max-pool (no shuffles), then ReLU, bf16 (or hi | lo planes for the accuracy mode), one 128-byte store per thread.
// Warp-specialised and persistent (one CTA per SM): two producer groups of four warps build alternate tiles (each
// thread reads its 4x4 input window straight from global / L1, one tile ahead), warp 8 issues the MMAs, warps 9-12
// run the epilogue; the A tiles and the TMEM accumulators are two-deep rings.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>

#include "igemm_sm100.cuh"
#include "kernels.cuh"
#include "sm100_ptx.cuh"

namespace vmb {

namespace {

constexpr int kH = 96, kW = 64, kC = 64;
constexpr int kPH = kH / 2, kPW = kW / 2;        // 48 x 32 pooled
constexpr int kRowsPerTile = 4;                  // pooled rows per CTA tile (x 32 columns = 128 pooled pixels)
constexpr int kTilesPerExample = kPH / kRowsPerTile;  // 12
constexpr int kK = 64;                           // GEMM K: [win_hi(16) | win_lo(16) | win_hi(16) | 1 | 1 | 0 ...]
constexpr int kLbo = 128;                        // bytes between the 16-byte K chunks of one 8-row group
constexpr int kSbo = kK / 8 * kLbo;              // 1024 bytes between 8-row groups (8 chunks x 128 B)
constexpr int kATile = 128 / 8 * kSbo;           // 16 KiB: one im2col-free tile (the 4x4 windows of 128 pooled pixels)
constexpr int kBTile = kC / 8 * kSbo;            // 8 KiB per pooling-window position
constexpr int kProducerThreads = 128;            // per producer group
constexpr int kProducerGroups = 2;               // warps 0-3 and 4-7: group g builds the tiles with (iteration & 1) == g
constexpr int kMmaWarp = 8;                      // warp 8 issues the MMAs
#ifndef VMB_CONV1_EPI_WARPS
#define VMB_CONV1_EPI_WARPS 8
#endif
constexpr int kEpiWarps = VMB_CONV1_EPI_WARPS;                     // warps 9-16: epilogue, two per TMEM lane quarter (read-out paced)
constexpr int kThreads = (kMmaWarp + 1 + kEpiWarps) * 32;
constexpr int kStages = 2;                       // A-tile ring (one stage per producer group) and TMEM accumulator ring
constexpr int kTmemCols = 512;                   // 2 buffers x 4 positions x 64 channels
constexpr int kOutTile = 128 * kC * 2;            // 16 KiB: the bf16 output of one tile = 128 pooled pixels x 64 channels,
                                                 // one contiguous block of the NHWC tensor (4 full pooled rows)
template <bool SPLIT_OUT>
constexpr int conv1_smem_bytes() { return 1024 + kStages * kATile + 4 * kBTile + (SPLIT_OUT ? 4 : 2) * kOutTile; }

// K-major, no swizzle: element (row r, 16-byte chunk j) at (r / 8) * SBO + j * LBO + (r % 8) * 16
__device__ __forceinline__ uint64_t umma_desc_kmajor_noswizzle(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(kLbo >> 4) << 16;
  d |= static_cast<uint64_t>(kSbo >> 4) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ uint32_t bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }

// The 4x4 input window of one pooled pixel (rows 2*prow-1 .. 2*prow+2, columns 2*pcol-1 .. 2*pcol+2), zero outside the
// image: read straight from global / L1 (neighbouring threads overlap), issued one tile ahead of its use.
__device__ __forceinline__ void load_window(const float* __restrict__ x, long long tile, int pr, int pc, float (&v)[16]) {
  const long long n = tile / kTilesPerExample;
  const int tr = static_cast<int>(tile - n * kTilesPerExample);
  const int r0 = 2 * (tr * kRowsPerTile + pr) - 1, c0 = 2 * pc - 1;
  const float* src = x + n * (kH * kW);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gr = r0 + i;
    const bool rok = gr >= 0 && gr < kH;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = c0 + j;
      v[4 * i + j] = (rok && gc >= 0 && gc < kW) ? __ldg(src + gr * kW + gc) : 0.f;
    }
  }
}

template <bool SPLIT_OUT>
__global__ void __launch_bounds__(kThreads, 1)
conv1_tc_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                const __grid_constant__ CUtensorMap tmap_out, long long n_tiles, int f16_out, int* sat_flag) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;                                   // [kStages][kATile]
  uint8_t* b_smem = smem + kStages * kATile;                // [4 positions][kBTile]
  uint8_t* o_smem = b_smem + 4 * kBTile;                    // [2 buffers][hi (, lo)][kOutTile], 128-byte swizzle
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], tmem_full[kStages], tmem_empty[kStages];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- one-time setup: the four B tiles (one per pooling-window position), the constant chunks of A, barriers, TMEM
  // B_pos[channel][k]: the 3x3 kernel scattered to the window slots (dy+ky, dx+kx) it touches for position (dy, dx):
  // w_hi against win_hi, w_hi against win_lo, w_lo against win_hi, then b_hi, b_lo against the two ones.
  if (tid < 4 * kC) {
    const int pos = tid / kC, c = tid % kC, dy = pos >> 1, dx = pos & 1;
    uint32_t k[kK];
#pragma unroll
    for (int i = 0; i < kK; ++i) k[i] = 0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float wv = __ldg(w + c * 9 + ky * 3 + kx);
        const __nv_bfloat16 wh = __float2bfloat16_rn(wv);
        const int slot = (dy + ky) * 4 + dx + kx;
        k[slot] = __bfloat16_as_ushort(wh);
        k[16 + slot] = k[slot];
        k[32 + slot] = bf16_bits(wv - __bfloat162float(wh));
      }
    const float bias = __ldg(b + c);
    const __nv_bfloat16 bh = __float2bfloat16_rn(bias);
    k[48] = __bfloat16_as_ushort(bh);
    k[49] = bf16_bits(bias - __bfloat162float(bh));
    uint8_t* row = b_smem + pos * kBTile + (c / 8) * kSbo + (c % 8) * 16;
#pragma unroll
    for (int j = 0; j < kK / 8; ++j)
      *reinterpret_cast<uint4*>(row + j * kLbo) =
          make_uint4(k[8 * j] | (k[8 * j + 1] << 16), k[8 * j + 2] | (k[8 * j + 3] << 16),
                     k[8 * j + 4] | (k[8 * j + 5] << 16), k[8 * j + 6] | (k[8 * j + 7] << 16));
  }
  if (tid < kProducerThreads * kStages) {   // chunks 6 (1, 1, 0 ...) and 7 (zeros) of every A row never change
    const int stg = tid / kProducerThreads, r = tid % kProducerThreads;
    uint8_t* row = a_smem + stg * kATile + (r / 8) * kSbo + (r % 8) * 16;
    *reinterpret_cast<uint4*>(row + 6 * kLbo) = make_uint4(0x3F803F80u, 0, 0, 0);   // bf16 1.0, 1.0
    *reinterpret_cast<uint4*>(row + 7 * kLbo) = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], kProducerThreads);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&tmem_slot, kTmemCols);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);   // warp-uniform by construction
  pdl_wait();   // the set-up above read only the layer's constant weights; the examples are read below

  if (warp < 4 * kProducerGroups) {
    // ------------------------------------------------------------------ producers: 4x4 window -> four im2col rows
    const int group = warp >> 2;                     // owns A stage `group`
    const int row = tid & (kProducerThreads - 1);    // pooled pixel within the tile = A row = TMEM lane
    const int pr = row >> 5, pc = row & 31;
    const long long stride = static_cast<long long>(gridDim.x) * kProducerGroups;
    long long tile = blockIdx.x + static_cast<long long>(group) * gridDim.x;
    float win[16];
    if (tile < n_tiles) load_window(x, tile, pr, pc, win);
    uint32_t use = 0;
    for (; tile < n_tiles; tile += stride, ++use) {
      // hi / lo halves of the window, two per register in window order (the GEMM K order)
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(win[2 * j], win[2 * j + 1]);
        hi[j] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[j] = pack_bf16x2(win[2 * j] - __low2float(h2), win[2 * j + 1] - __high2float(h2));
      }
      if (tile + stride < n_tiles) load_window(x, tile + stride, pr, pc, win);   // in flight during the stores
      mbar_wait(&empty_bar[group], (use & 1) ^ 1);         // the MMAs that read this stage last time are done
      uint8_t* dst = a_smem + group * kATile + (row / 8) * kSbo + (row % 8) * 16;
      const uint4 h0 = make_uint4(hi[0], hi[1], hi[2], hi[3]), h1 = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      *reinterpret_cast<uint4*>(dst + 0 * kLbo) = h0;
      *reinterpret_cast<uint4*>(dst + 1 * kLbo) = h1;
      *reinterpret_cast<uint4*>(dst + 2 * kLbo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(dst + 3 * kLbo) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      *reinterpret_cast<uint4*>(dst + 4 * kLbo) = h0;
      *reinterpret_cast<uint4*>(dst + 5 * kLbo) = h1;
      fence_proxy_async_smem();        // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[group]);
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    // The four weight tiles (one per pooling-window position) are stacked in shared memory exactly like one 256-row B
    // operand, and all four multiply the same A tile: one N = 256 MMA per K step instead of four N = 64 ones (A is read
    // from shared memory once instead of four times: 12 KB instead of 24 KB per K step).
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 4 * kC);
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t stage = it & 1, ph = (it >> 1) & 1;
      mbar_wait(&tmem_empty[stage], ph ^ 1);
      mbar_wait(&full_bar[stage], ph);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint64_t a_desc = umma_desc_kmajor_noswizzle(smem_u32(a_smem + stage * kATile));
        const uint64_t b_desc = umma_desc_kmajor_noswizzle(smem_u32(b_smem));
        {
#pragma unroll
          for (int ks = 0; ks < kK / 16; ++ks)   // K step 16 = two chunks = 2 * LBO = 256 bytes (>> 4 = 16)
            umma_bf16_ss(tmem_base + stage * 256, a_desc + ks * (2 * kLbo >> 4), b_desc + ks * (2 * kLbo >> 4), idesc, ks);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[stage]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue: pool = max over 4 column blocks
    // The results are staged in shared memory (row = pooled pixel = 128 bytes, 128-byte swizzle so that the lanes of a
    // quarter-warp hit different banks) and leave as ONE TMA store per tile: the tile's output is a contiguous 16 KB
    // block of the NHWC tensor, and 32-byte stores at a 128-byte stride were what throttled this kernel.
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int half = (warp - kMmaWarp - 1) >> 2;  // the two warps of a quarter alternate over the 16-channel chunks
    const int row = q * 32 + lane;                // pooled pixel within the tile
    const bool store_thread = (warp == kMmaWarp + 1) && lane == 0;
    constexpr int kBufBytes = (SPLIT_OUT ? 2 : 1) * kOutTile;
    uint32_t it = 0;
    uint32_t sat_max = 0;   // fp16 output: running maximum of the packed results
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t stage = it & 1, ph = (it >> 1) & 1;
      uint8_t* obuf = o_smem + (it & 1) * kBufBytes;
      // the store of two tiles ago has finished reading this buffer
      if (store_thread) bulk_wait_group_read<1>();
      asm volatile("bar.sync 2, %0;" ::"n"(kEpiWarps * 32) : "memory");
      mbar_wait(&tmem_full[stage], ph);
      tc_fence_after_sync();
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + stage * 256;
#pragma unroll 1
      for (int ch = half; ch < 4; ch += kEpiWarps / 4) {      // 16 channels at a time keeps the register count low
        uint32_t v0[16], v1[16], v2[16], v3[16];
        tmem_ld_32x16(t_lane + 0 * kC + ch * 16, v0);
        tmem_ld_32x16(t_lane + 1 * kC + ch * 16, v1);
        tmem_ld_32x16(t_lane + 2 * kC + ch * 16, v2);
        tmem_ld_32x16(t_lane + 3 * kC + ch * 16, v3);
        tmem_ld_wait();
        uint32_t pk[8], pl[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = fmaxf(fmaxf(fmaxf(__uint_as_float(v0[2 * j]), __uint_as_float(v1[2 * j])),
                                      fmaxf(__uint_as_float(v2[2 * j]), __uint_as_float(v3[2 * j]))), 0.f);
          const float c = fmaxf(fmaxf(fmaxf(__uint_as_float(v0[2 * j + 1]), __uint_as_float(v1[2 * j + 1])),
                                      fmaxf(__uint_as_float(v2[2 * j + 1]), __uint_as_float(v3[2 * j + 1]))), 0.f);
          if (!SPLIT_OUT && f16_out) {
            pk[j] = pack_f16x2(a, c);
            sat_max = max_f16x2(sat_max, pk[j]);
          } else {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, c);
            pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
            if (SPLIT_OUT) pl[j] = pack_bf16x2(a - __low2float(h2), c - __high2float(h2));
          }
        }
        // 16-byte pieces 2*ch and 2*ch+1 of this pixel's 128-byte row, at the swizzled positions TMA expects
        uint8_t* orow = obuf + row * 128;
        const int s0 = ((2 * ch) ^ (row & 7)) * 16, s1 = ((2 * ch + 1) ^ (row & 7)) * 16;
        *reinterpret_cast<uint4*>(orow + s0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(orow + s1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (SPLIT_OUT) {
          *reinterpret_cast<uint4*>(orow + kOutTile + s0) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
          *reinterpret_cast<uint4*>(orow + kOutTile + s1) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[stage]);
      fence_proxy_async_smem();                   // the generic-proxy writes above, before the async-proxy (TMA) read
      asm volatile("bar.sync 3, %0;" ::"n"(kEpiWarps * 32) : "memory");
      if (store_thread) {
        // output rows: bf16 NHWC [n][48*32][64], or hi | lo planes [n][2][48*32][64]; tile = 128 consecutive rows
        const long long n = tile / kTilesPerExample;
        const int tr = static_cast<int>(tile - n * kTilesPerExample);
        constexpr int kRowsPerImg = kPH * kPW;
        if (SPLIT_OUT) {
          tma_store_2d(&tmap_out, obuf, 0, static_cast<int>((2 * n) * kRowsPerImg + tr * 128));
          tma_store_2d(&tmap_out, obuf + kOutTile, 0, static_cast<int>((2 * n + 1) * kRowsPerImg + tr * 128));
        } else {
          tma_store_2d(&tmap_out, obuf, 0, static_cast<int>(n * kRowsPerImg + tr * 128));
        }
        bulk_commit_group();
      }
    }
    if (store_thread) bulk_wait_group<0>();       // shared memory stays valid until the last store has read it
    if (sat_flag && saturated_f16x2(sat_max)) *reinterpret_cast<volatile int*>(sat_flag) = 1;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

int conv1_tc_relu_pool(const float* examples, const float* w, const float* b, void* out, long long n,
                       cudaStream_t stream, bool split_out, int fmt, int* sat_flag) {
  const long long tiles = n * kTilesPerExample;
  if (tiles <= 0) return 0;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(tiles, num_sms()));
  static std::atomic<unsigned long long> attr_set{0};   // one bit per device: the attribute is per device
  if (device_needs_setup(attr_set)) {
    if (cudaFuncSetAttribute(conv1_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             conv1_smem_bytes<false>()) != cudaSuccess ||
        cudaFuncSetAttribute(conv1_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             conv1_smem_bytes<true>()) != cudaSuccess) {
      device_setup_failed(attr_set);
      set_kernel_error("conv1: cannot set the dynamic shared memory size");
      return 1;
    }
  }
  // the output as [rows = n * (1 or 2 planes) * 48 * 32][64 channels]; one box = one tile = 128 rows
  CUtensorMap to;
  {
    const long long rows = n * (split_out ? 2 : 1) * kPH * kPW;
    if (rows > 0x7fffffffLL) {
      set_kernel_error("conv1: too many output rows for one call");
      return 1;
    }
    uint64_t dims[2] = {uint64_t(kC), uint64_t(rows)};
    uint64_t str[1] = {uint64_t(kC) * 2};
    uint32_t box[2] = {uint32_t(kC), 128};
    if (make_tmap_bf16(&to, out, 2, dims, str, box)) {
      set_kernel_error("conv1: %s", igemm_last_error());
      return 1;
    }
  }
  const cudaError_t e = split_out ? launch_pdl(conv1_tc_kernel<true>, dim3(grid), dim3(kThreads), conv1_smem_bytes<true>(),
                                               stream, examples, w, b, to, tiles, 0, static_cast<int*>(nullptr))
                                  : launch_pdl(conv1_tc_kernel<false>, dim3(grid), dim3(kThreads), conv1_smem_bytes<false>(),
                                               stream, examples, w, b, to, tiles, fmt == kFmtF16 ? 1 : 0, sat_flag);
  count_launch();
  if (e != cudaSuccess) {
    set_kernel_error("conv1_tc_kernel: %s", cudaGetErrorString(e));
    return 1;
  }
  return check_launch("conv1_tc_kernel");
}

}  // namespace vmb

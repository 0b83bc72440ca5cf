// CTA-pair (cta_group::2) variant of the implicit-GEMM engine: the bf16 conv3x3 / FC layers of the VGGish body
// (reference vggish.py:13-19, :108-118) with one 256-row x BLOCK_N tile per pair of SMs.
//
// Why a second kernel: with one CTA per 128 x 256 tile every K block moves 16 KB of A + 32 KB of B from L2 to shared
// memory for 512 tensor-pipe cycles (96 B/clk/SM), and the layers with a short K loop or a large output were bound by
// that L2 traffic (plus the epilogue's stores contending with it), not by the tensor pipe.  A pair of CTAs on one TPC
// computes a 256 x 256 tile with tcgen05.mma.cta_group::2: each CTA loads its own 128 rows of A and only HALF of the
// B tile (its 128 of the 256 output channels), the MMA reads both halves across the pair.  Per SM that is 32 KB per
// K block (64 B/clk), six pipeline stages instead of four, and half the shared-memory reads for B.
//
// Protocol (leader = CTA rank 0 of the cluster):
//   producers (both CTAs, one elected lane)  wait own empty[s] -> TMA (cta_group::2) into OWN smem, bytes credited to
//                                            the LEADER's full[s]; the leader's producer posts the expect_tx for both
//   MMA issuer (leader only)                 wait full[s] -> 4 x tcgen05.mma.cta_group::2 -> commit multicast to
//                                            empty[s] of both CTAs; after the last K block commit multicast to
//                                            tmem_full[acc] of both CTAs
//   epilogue (both CTAs, 8 warps each)       wait own tmem_full[acc] -> TMEM -> bias/ReLU/(pool) -> global; arrive on
//                                            the LEADER's tmem_empty[acc] (count 16)
// Everything else (K order, NHWC boxes with TMA zero-fill as the conv padding, pooling by warp shuffles, full-sector
// stores) is the single-CTA kernel's scheme; see igemm_sm100.cuh.
#include <cstdio>

#include "igemm_sm100.cuh"
#include "kernels.cuh"
#include "sm100_ptx.cuh"

namespace vmb {

namespace {

constexpr int kBlockM = 128;                       // rows per CTA (256 per pair)
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;     // 16 KiB
constexpr int kQuarterBytes = 32 * kBlockK * 2;
constexpr int kEpiWarps = 8;                       // two per TMEM lane quarter: the short-K layers are epilogue-paced
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kNumThreads = 64 + kEpiThreads;

// MT = 128-row sub-tiles per CTA that share one B stage (MT = 2 for C_out = 128: a 512 x 128 pair tile)
// HALO: as in the single-CTA kernel (igemm_sm100.cu, Cfg) — a stage is one haloed activation box (10 rows x 16 pixels x
// 64 channels, loaded at column offset dx) plus this CTA's halves of the three weight tiles (dy = -1, 0, 1; dx): 68 KB
// for 12 MMAs (1536 tensor cycles), 44 instead of 64 B/clk/SM from L2.
constexpr int kHaloRows = 10, kHaloWb = 16;
constexpr int kHaloBytes = kHaloRows * kHaloWb * kBlockK * 2;   // 20480
template <int BLOCK_N, int MT, bool HALO = false>
struct PairCfg {
  static constexpr int kBHalfBytes = (BLOCK_N / 2) * kBlockK * 2;   // this CTA's half of the B tile
  static constexpr int kStageBytes = HALO ? MT * kHaloBytes + 3 * kBHalfBytes : MT * kABytes + kBHalfBytes;
  static constexpr int kStages = (HALO ? 208 : 200) * 1024 / kStageBytes;   // 6 x 32 KB, or 3 x 68 KB with HALO
  static_assert(kStages >= 2, "need a double-buffered stage ring");
  static constexpr int kTmemCols = 2 * MT * BLOCK_N;                // two accumulator buffers
  static_assert(kTmemCols <= 512, "TMEM holds 512 columns");
  static constexpr int kBiasBytes = 2 * BLOCK_N * 4;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kBiasBytes + kBarBytes;
};

template <int BLOCK_N, int MT, bool CONV, bool POOL, bool BIG, bool HALO = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
igemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const IgemmParams p) {
  static_assert(!HALO || (CONV && BIG), "HALO is a variant of the big-box conv");
  using C = PairCfg<BLOCK_N, MT, HALO>;
  extern __shared__ uint8_t smem_raw[];
  // both CTAs of the pair must use identical shared-memory offsets (the MMA and the multicast commits address the
  // peer's memory by offset): the dynamic segment starts at the same offset in both, so the same rounding applies
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint8_t* stage_base = smem;
  float* bias_s = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + C::kBiasBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full = empty_bar + C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_m_pairs = (p.num_m_tiles + 2 * MT - 1) / (2 * MT);
  const int all_tiles = num_m_pairs * p.num_n_tiles;     // pair tiles (2 * MT * 128 rows x BLOCK_N)
  const int num_tiles = (p.tile_count > 0 && p.tile_count < all_tiles) ? p.tile_count : all_tiles;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);    // used in the leader only: its producer's arrive.expect_tx
      mbar_init(&empty_bar[s], 1);   // the leader's multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);   // the leader's multicast commit
      mbar_init(&tmem_empty[s], 2 * kEpiWarps);  // used in the leader only: the epilogue warps of both CTAs
    }
    mbar_fence_init_cluster();
  }
  __syncthreads();
  cluster_sync_all();                // barriers of both CTAs exist before anything signals across the pair
  if (warp == 1) tmem_alloc_pair(tmem_slot, C::kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      const uint32_t full0 = mapa_u32(smem_u32(&full_bar[0]), 0);   // the leader's full barriers
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        const int m_pair = tile / p.num_n_tiles;
        const int n_tile = tile - m_pair * p.num_n_tiles;
        const int m_tile = (m_pair * 2 + static_cast<int>(rank)) * MT;   // may run past the end: TMA zero-fills
        int bx[4 * MT], by[4 * MT], bn[4 * MT];
        if (CONV && BIG) {
#pragma unroll
          for (int sub = 0; sub < MT; ++sub) {
            const int g = m_tile + sub;
            const int n_img = g / p.boxes_per_img;
            const int r = g - n_img * p.boxes_per_img;
            const int yy = r / p.boxes_per_row;
            bn[sub] = n_img;
            by[sub] = yy * 4 * p.Hb;
            bx[sub] = (r - yy * p.boxes_per_row) * p.Wb;
          }
        } else if (CONV) {
#pragma unroll
          for (int q = 0; q < 4 * MT; ++q) {
            const int g = m_tile * 4 + q;
            const int n_img = g / p.boxes_per_img;
            const int r = g - n_img * p.boxes_per_img;
            const int yy = r / p.boxes_per_row;
            bn[q] = n_img;
            by[q] = yy * p.Hb;
            bx[q] = (r - yy * p.boxes_per_row) * p.Wb;
          }
        }
        const int b_row = n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2);
        if (HALO) {
          for (int s = 0; s < 3 * p.cblks; ++s) {
            const int cbh = s / 3, dxi = s - cbh * 3;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* a_dst = stage_base + stage * C::kStageBytes;
            uint8_t* b_dst = a_dst + MT * kHaloBytes;
            const uint32_t full = full0 + stage * 8;
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
#pragma unroll
            for (int sub = 0; sub < MT; ++sub)
              tma_load_4d_pair(a_dst + sub * kHaloBytes, &tmap_a, full, cbh * kBlockK, bx[sub] + dxi - 1, by[sub] - 1,
                               bn[sub]);
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi)
              tma_load_2d_pair(b_dst + dyi * C::kBHalfBytes, &tmap_b, full, ((dyi * 3 + dxi) * p.cblks + cbh) * kBlockK,
                               b_row);
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        int tap = 0, cb = 0;   // K order (channel block, dx, dy): see the single-CTA kernel; tap = dxi * 3 + dyi
        int b_kb = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = stage_base + stage * C::kStageBytes;
          uint8_t* b_dst = a_dst + MT * kABytes;
          const uint32_t full = full0 + stage * 8;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
          if (CONV) {
            const int dxi = tap / 3, dyi = tap - dxi * 3;
            const int dh = dyi - 1, dw = dxi - 1;
            b_kb = (dyi * 3 + dxi) * p.cblks + cb;
            if (BIG) {
#pragma unroll
              for (int sub = 0; sub < MT; ++sub)
                tma_load_4d_pair(a_dst + sub * kABytes, &tmap_a, full, cb * kBlockK, bx[sub] + dw, by[sub] + dh,
                                 bn[sub]);
            } else {
#pragma unroll
              for (int q = 0; q < 4 * MT; ++q)
                tma_load_4d_pair(a_dst + q * kQuarterBytes, &tmap_a, full, cb * kBlockK, bx[q] + dw, by[q] + dh, bn[q]);
            }
            if (++tap == 9) { tap = 0; ++cb; }
          } else {
            b_kb = kb;
#pragma unroll
            for (int sub = 0; sub < MT; ++sub)
              tma_load_2d_pair(a_dst + sub * kABytes, &tmap_a, full, kb * kBlockK, (m_tile + sub) * kBlockM);
          }
          tma_load_2d_pair(b_dst, &tmap_b, full, b_kb * kBlockK, b_row);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
      // producer tail: every stage this CTA filled has been released, i.e. all of the leader's multicast commits to THIS
      // CTA's empty barriers have landed before the CTA can reach the teardown barrier and exit
      for (int i = 0; i < C::kStages; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = p.f16 ? umma_idesc_f16_f32(2 * kBlockM, BLOCK_N) : umma_idesc_bf16_f32(2 * kBlockM, BLOCK_N);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * (MT * BLOCK_N);
        if (HALO) {
          const int n_st = 3 * p.cblks;
          for (int s = 0; s < n_st; ++s) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t a_addr = smem_u32(stage_base + stage * C::kStageBytes);
              const uint32_t b_addr = a_addr + MT * kHaloBytes;
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi) {
                const uint64_t b_desc = umma_desc_kmajor_sw128(b_addr + dyi * C::kBHalfBytes);
#pragma unroll
                for (int sub = 0; sub < MT; ++sub) {
                  const uint64_t a_desc =
                      umma_desc_kmajor_sw128(a_addr + sub * kHaloBytes + dyi * (kHaloWb * kBlockK * 2));
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k)
                    umma_bf16_ss_pair(d_tmem + sub * BLOCK_N, a_desc + 2 * k, b_desc + 2 * k, idesc, (s | dyi | k) != 0);
                }
              }
              umma_commit_pair(&empty_bar[stage], 0b11);
              if (s == n_st - 1) umma_commit_pair(&tmem_full[acc], 0b11);
            }
            __syncwarp();
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(stage_base + stage * C::kStageBytes);
            const uint64_t b_desc = umma_desc_kmajor_sw128(a_addr + MT * kABytes);
#pragma unroll
            for (int sub = 0; sub < MT; ++sub) {
              const uint64_t a_desc = umma_desc_kmajor_sw128(a_addr + sub * kABytes);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_bf16_ss_pair(d_tmem + sub * BLOCK_N, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            }
            umma_commit_pair(&empty_bar[stage], 0b11);
            if (kb == p.num_kb - 1) umma_commit_pair(&tmem_full[acc], 0b11);
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9 of both CTAs)
    // warp w may read TMEM lanes 32*(w%4)..+31; the two warps of a quarter split the tile's sub-tiles (MT = 2) or
    // alternate over its 32-column chunks (MT = 1)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ep_tid = threadIdx.x - 64;
    const uint32_t tmem_empty0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    uint32_t it = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++it) {
      const int m_pair = tile / p.num_n_tiles;
      const int n_tile = tile - m_pair * p.num_n_tiles;
      const int m_tile0 = (m_pair * 2 + static_cast<int>(rank)) * MT;
      const int n0 = n_tile * BLOCK_N;
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      float* bias_t = bias_s + acc * BLOCK_N;
      for (int i = ep_tid; i < BLOCK_N; i += kEpiThreads) bias_t[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
#pragma unroll 1
      for (int mt = (MT == 2 ? half : 0); mt < (MT == 2 ? half + 1 : 1); ++mt) {
      const int m_tile = m_tile0 + mt;
      bool valid;
      size_t out_off;
      int sub = 0;
      if (CONV) {
        const int g = BIG ? m_tile : m_tile * 4 + q;
        const int n_img = g / p.boxes_per_img;
        const int r = g - n_img * p.boxes_per_img;
        const int yy = r / p.boxes_per_row;
        const int hh = lane / p.Wb, ww = lane - hh * p.Wb;
        const int h = BIG ? yy * 4 * p.Hb + q * p.Hb + hh : yy * p.Hb + hh;
        const int w = (r - yy * p.boxes_per_row) * p.Wb + ww;
        valid = n_img < p.M;
        if (POOL) {
          sub = (ww & 1) | ((hh & 1) << 1);
          out_off = static_cast<size_t>(n_img) * p.out_img_stride +
                    (static_cast<size_t>(h >> 1) * (p.W >> 1) + (w >> 1)) * p.N + n0;
        } else {
          out_off = static_cast<size_t>(n_img) * p.out_img_stride + (static_cast<size_t>(h) * p.W + w) * p.N + n0;
        }
      } else {
        const int row = m_tile * kBlockM + q * 32 + lane;
        valid = row < p.M;
        out_off = static_cast<size_t>(row) * p.ldo + n0;
      }

      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * (MT * BLOCK_N) + mt * BLOCK_N;
      uint32_t sat_max = 0;   // fp16 mode: running maximum of the packed outputs
#pragma unroll 1
      for (int ch = (MT == 2 ? 0 : half); ch < BLOCK_N / 32; ch += (MT == 2 ? 1 : 2)) {
        uint32_t v[32];
        tmem_ld_32x32(t_addr + ch * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_t + ch * 32 + j);
          f[j + 0] = __uint_as_float(v[j + 0]) + b4.x;
          f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
          f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
          f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
        }
        if (p.relu && !POOL) {   // the pooled path applies ReLU after the pooling (fewer values)
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        void* outp = static_cast<uint16_t*>(p.out) + out_off + ch * 32;
        if (p.f16)
          sat_max = max_f16x2(sat_max, store_row32_16bit<true, POOL>(f, sub, p.Wb, p.relu != 0, valid, outp));
        else
          store_row32_16bit<false, POOL>(f, sub, p.Wb, p.relu != 0, valid, outp);
      }
      if (p.f16 && p.sat_flag && saturated_f16x2(sat_max)) *reinterpret_cast<volatile int*>(p.sat_flag) = 1;
      }  // mt
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tmem_empty0 + acc * 8);
    }
  }

  // nobody leaves (or frees tensor memory) while the partner may still read this CTA's shared memory or signal its
  // barriers
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, C::kTmemCols);
  }
}

thread_local char g_pair_err[512] = "";

template <int BLOCK_N, int MT, bool CONV, bool POOL, bool BIG, bool HALO = false>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const IgemmParams& p, cudaStream_t stream) {
  auto kern = igemm_pair_kernel<BLOCK_N, MT, CONV, POOL, BIG, HALO>;
  static std::atomic<unsigned long long> attr_set{0};   // one bit per device: the attribute is per device
  constexpr int smem = PairCfg<BLOCK_N, MT, HALO>::kSmemBytes;
  if (device_needs_setup(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      device_setup_failed(attr_set);
      snprintf(g_pair_err, sizeof g_pair_err, "cudaFuncSetAttribute(pair, smem=%d): %s", smem, cudaGetErrorString(e));
      return 1;
    }
  }
  const int all_pair_tiles = ((p.num_m_tiles + 2 * MT - 1) / (2 * MT)) * p.num_n_tiles;
  const int pair_tiles = (p.tile_count > 0 && p.tile_count < all_pair_tiles) ? p.tile_count : all_pair_tiles;
  const int max_pairs = num_sms() / 2;
  const int grid = 2 * (pair_tiles < max_pairs ? pair_tiles : max_pairs);
  // cluster shape comes from __cluster_dims__
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(kNumThreads), smem, stream, ta, tb, p);
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_pair_err, sizeof g_pair_err, "igemm pair launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

}  // namespace

const char* igemm_pair_last_error() { return g_pair_err; }

int igemm_pair_linear(const void* a, const void* w, const float* bias, void* out, int relu, int M, int N, int K,
                      cudaStream_t stream, int fmt, int* sat_flag) {
  if (M <= 0) return 0;
  if (K % kBlockK != 0 || N % 256 != 0) {
    snprintf(g_pair_err, sizeof g_pair_err, "igemm_pair_linear: need K %% 64 == 0 and N %% 256 == 0");
    return 1;
  }
  CUtensorMap ta, tb;
  {
    uint64_t dims[2] = {uint64_t(K), uint64_t(M)};
    uint64_t str[1] = {uint64_t(K) * 2};
    uint32_t box[2] = {kBlockK, kBlockM};
    if (make_tmap_bf16(&ta, a, 2, dims, str, box)) return 2;
  }
  {
    uint64_t dims[2] = {uint64_t(K), uint64_t(N)};
    uint64_t str[1] = {uint64_t(K) * 2};
    uint32_t box[2] = {kBlockK, 128};
    if (make_tmap_bf16(&tb, w, 2, dims, str, box)) return 2;
  }
  IgemmParams p{};
  p.M = M;
  p.N = N;
  p.num_kb = K / kBlockK;
  p.num_m_tiles = (M + kBlockM - 1) / kBlockM;
  p.num_n_tiles = N / 256;
  p.relu = relu;
  p.ldo = N;
  p.bias = bias;
  p.out = out;
  p.f16 = fmt == kFmtF16;
  p.sat_flag = sat_flag;
  // Wave quantisation: T tiles on P SM pairs run as ceil(T / P) rounds.  When the last round would be less than half
  // full (fc1 / fc2 at 2 560 rows: 160 tiles on 74 pairs = 2 full rounds + 12 tiles), the full rounds go to the pair
  // kernel and the left-over tiles — one or two rectangles of the output — to the single-CTA kernel with 128 x 128
  // tiles, which spreads them over all SMs at half the tile duration.  Same K order, so every output element is
  // the same sum either way (bit-identical; tested).
  const int P = num_sms() / 2;
  const int nn = p.num_n_tiles, T = ((p.num_m_tiles + 1) / 2) * nn;
  const int left = T > P ? T % P : 0;
  if (left == 0 || 2 * left > P) return launch_pair<256, 1, false, false, false>(ta, tb, p, stream);
  const int full = T - left;
  p.tile_count = full;
  if (launch_pair<256, 1, false, false, false>(ta, tb, p, stream)) return 1;
  const char* a8 = static_cast<const char*>(a);
  const char* w8 = static_cast<const char*>(w);
  char* o8 = static_cast<char*>(out);
  int r = full / nn;
  const int c = full - r * nn;
  if (c > 0) {   // the rest of the partly covered row of tiles
    const int row0 = r * 256, rows = (M - row0 < 256) ? M - row0 : 256, col0 = c * 256;
    if (igemm_linear_rect(a8 + size_t(row0) * K * 2, w8 + size_t(col0) * K * 2, bias ? bias + col0 : nullptr,
                             o8 + (size_t(row0) * N + col0) * 2, N, relu, rows, N - col0, K, stream, fmt, sat_flag)) {
      snprintf(g_pair_err, sizeof g_pair_err, "%s", igemm_last_error());
      return 1;
    }
    ++r;
  }
  if (r * 256 < M) {   // whole rows of tiles below
    const int row0 = r * 256;
    if (igemm_linear_rect(a8 + size_t(row0) * K * 2, w, bias, o8 + size_t(row0) * N * 2, N, relu, M - row0, N, K, stream, fmt,
                          sat_flag)) {
      snprintf(g_pair_err, sizeof g_pair_err, "%s", igemm_last_error());
      return 1;
    }
  }
  return 0;
}

int igemm_pair_conv3x3(const void* act, const void* w, const float* bias, void* out, int n_img, int H, int W, int C_in,
                       int C_out, int pool, cudaStream_t stream, int fmt, int* sat_flag) {
  if (n_img <= 0) return 0;
  const int Wb = (W % 16 == 0) ? 16 : 8;
  const int Hb = 32 / Wb;
  if (C_in % kBlockK != 0 || C_out % 256 != 0 || W % Wb != 0 || H % Hb != 0) {
    snprintf(g_pair_err, sizeof g_pair_err, "igemm_pair_conv3x3: unsupported geometry");
    return 1;
  }
  const int block_n = (C_out % 256 == 0) ? 256 : 128;
  const int K = 9 * C_in;
  const bool big = H % (4 * Hb) == 0;
  CUtensorMap ta, tb;
  {
    uint64_t dims[4] = {uint64_t(C_in), uint64_t(W), uint64_t(H), uint64_t(n_img)};
    uint64_t str[3] = {uint64_t(C_in) * 2, uint64_t(W) * C_in * 2, uint64_t(H) * W * C_in * 2};
    uint32_t box[4] = {kBlockK, uint32_t(Wb), uint32_t(big ? 4 * Hb : Hb), 1};
    if (make_tmap_bf16(&ta, act, 4, dims, str, box)) return 2;
  }
  {
    uint64_t dims[2] = {uint64_t(K), uint64_t(C_out)};
    uint64_t str[1] = {uint64_t(K) * 2};
    uint32_t box[2] = {kBlockK, uint32_t(block_n / 2)};
    if (make_tmap_bf16(&tb, w, 2, dims, str, box)) return 2;
  }
  IgemmParams p{};
  p.M = n_img;
  p.N = C_out;
  p.num_kb = K / kBlockK;
  p.cblks = C_in / kBlockK;
  p.H = H;
  p.W = W;
  p.Hb = Hb;
  p.Wb = Wb;
  p.boxes_per_row = W / Wb;
  p.big_box = big ? 1 : 0;
  p.boxes_per_img = big ? (H / (4 * Hb)) * (W / Wb) : (H / Hb) * (W / Wb);
  p.total_boxes = n_img * p.boxes_per_img;
  p.num_m_tiles = big ? p.total_boxes : (p.total_boxes + 3) / 4;
  p.num_n_tiles = C_out / block_n;
  p.relu = 1;
  p.ldo = C_out;
  p.bias = bias;
  p.out = out;
  p.f16 = fmt == kFmtF16;
  p.sat_flag = sat_flag;
  p.out_img_stride = static_cast<long long>(pool ? (H / 2) * (W / 2) : H * W) * C_out;
  if (big && Wb == kHaloWb && 4 * Hb + 2 == kHaloRows && igemm_use_halo()) {
    CUtensorMap th;
    uint64_t dims[4] = {uint64_t(C_in), uint64_t(W), uint64_t(H), uint64_t(n_img)};
    uint64_t str[3] = {uint64_t(C_in) * 2, uint64_t(W) * C_in * 2, uint64_t(H) * W * C_in * 2};
    uint32_t box[4] = {kBlockK, uint32_t(kHaloWb), uint32_t(kHaloRows), 1};
    if (make_tmap_bf16(&th, act, 4, dims, str, box)) return 2;
    return pool ? launch_pair<256, 1, true, true, true, true>(th, tb, p, stream)
                : launch_pair<256, 1, true, false, true, true>(th, tb, p, stream);
  }
  if (big)
    return pool ? launch_pair<256, 1, true, true, true>(ta, tb, p, stream)
                : launch_pair<256, 1, true, false, true>(ta, tb, p, stream);
  return pool ? launch_pair<256, 1, true, true, false>(ta, tb, p, stream)
              : launch_pair<256, 1, true, false, false>(ta, tb, p, stream);
}

}  // namespace vmb

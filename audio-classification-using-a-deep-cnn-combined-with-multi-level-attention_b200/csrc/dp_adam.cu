// Data-parallel head training (BASELINE configs[4]; the reference's optimiser step, train.py:133-138, with the gradient
// averaging that data parallelism adds): gradient reduce-scatter + Adam + parameter all-gather in ONE kernel over NVLink
// peer memory.
//
// Every rank owns an "arena" in device memory — params | grads (two buffers, used alternately) | flags — allocated by
// this library and opened by every other rank of the box through CUDA IPC, so a kernel on rank r can load and store
// any rank's arena directly (NVLink 5 through the NVSwitch: 1.8 us per dependent access, 775 GB/s per direction).
// After its backward pass a rank enqueues vmb_dp_adam_step, which runs three kernels on the caller's stream:
//
//   1. dp_barrier_kernel            one warp: writes the epoch into its flag slot in EVERY rank's arena, then waits
//                                   until every rank's epoch has arrived in its own arena — all gradients are final
//   2. dp_reduce_adam_gather_kernel the rank's slice [rank * chunk, (rank + 1) * chunk) of the bucket: loads the slice
//                                   of all `world` gradient buffers (peer loads, fixed rank order, so every element has
//                                   one well-defined sum), scales by 1 / world, applies torch.optim.Adam's update to
//                                   the slice with the rank's own slice of the moments, and stores the new parameters
//                                   into every rank's params (peer stores)
//   3. dp_barrier_kernel            every rank's slice has landed everywhere: the next forward may read params
//
// Against all-reduce + Adam on every rank (what NCCL + vmb_adam_step do: 109 us + 19 us at 8 ranks for the 7.96 MB
// bucket) each rank moves 7/8 of ONE bucket in and 7/8 of one bucket out instead of two bucket-sized ring passes, runs
// 1/8 of the Adam arithmetic, keeps 1/8 of the moments, and nothing but two flag exchanges is latency-bound.  The
// parameters of all ranks stay bit-identical by construction: each element is computed once and broadcast.
//
// Gradients alternate between two buffers: rank A may start the backward pass of step i + 1 (which clears and rewrites
// its gradients) while a slower rank B is still reading A's gradients of step i; with two buffers the rewrite of buffer
// (i mod 2) happens in step i + 2, after A has passed barrier 1 of step i + 1 — which B only signals after its own
// kernel 2 of step i.
//
// A rank that never arrives (a crashed peer) must not hang the GPU: the barrier gives up after kBarrierTimeoutNs, sets
// a flag the host reads with vmb_dp_status, and lets the stream continue.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/vggish_mla_b200.h"
#include "kernels.cuh"

namespace vmb {
void set_api_error(const char* msg);
}

namespace {

constexpr int kMaxRanks = 8;
constexpr int kFlagStride = 16;                          // u64 per flag slot: one 128-byte line each
constexpr unsigned long long kBarrierTimeoutNs = 10000000000ull;

int fail(const char* fmt, const char* detail = "") {
  char buf[640];
  snprintf(buf, sizeof buf, fmt, detail);
  vmb::set_api_error(buf);
  return 1;
}

struct DpPeers {
  float* params[kMaxRanks];
  const float* grads[kMaxRanks];          // the buffer of the current parity
  unsigned long long* flags[kMaxRanks];   // flags[r][src * kFlagStride]: epoch last signalled by rank src to rank r
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One warp.  Lane r < world signals rank r and waits for rank r.
__global__ void dp_barrier_kernel(DpPeers peers, int rank, int world, unsigned long long epoch, int* timed_out) {
  const int r = threadIdx.x;
  if (r < world) {
    // everything this rank's earlier kernels wrote (its gradients / its parameter slice in every arena) is ordered
    // before the flag at system scope
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.flags[r] + rank * kFlagStride), "l"(epoch) : "memory");
    const unsigned long long* mine = peers.flags[rank] + r * kFlagStride;
    const unsigned long long t0 = globaltimer_ns();
    unsigned long long seen;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
      if (seen >= epoch) break;
      if (globaltimer_ns() - t0 > kBarrierTimeoutNs) {
        *timed_out = 1;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncwarp();
  __threadfence_system();
}

__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;   // not through L1: peer lines are cached there and the two gradient buffers are re-read every other step
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, float lr, float b1, float b2, float eps,
                                          float wd, float bc1, float bc2_sqrt) {
  if (wd != 0.f) g = fmaf(wd, p, g);
  m = fmaf(b1, m, (1.f - b1) * g);       // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(b2, v, (1.f - b2) * g * g);   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  return p - (lr / bc1) * (m / denom);
}

// Elements [begin, end) of the bucket (multiples of 4).  The same arithmetic as adam_kernel (mla_train.cu) on
// g = (sum over ranks, in rank order) * gscale.
template <int WORLD>
__global__ void __launch_bounds__(256)
dp_reduce_adam_gather_kernel(DpPeers peers, int rank, long long begin, long long end, float* __restrict__ m,
                             float* __restrict__ v, float lr, float b1, float b2, float eps, float wd, float bc1,
                             float bc2_sqrt, float gscale) {
  for (long long i = begin + 4 * (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x); i < end;
       i += 4LL * gridDim.x * blockDim.x) {
    float4 g[WORLD];
#pragma unroll
    for (int r = 0; r < WORLD; ++r) g[r] = ld_peer_f4(peers.grads[r] + i);   // WORLD loads in flight per thread
    float4 s = g[0];
#pragma unroll
    for (int r = 1; r < WORLD; ++r) {
      s.x += g[r].x;
      s.y += g[r].y;
      s.z += g[r].z;
      s.w += g[r].w;
    }
    float4 p = *reinterpret_cast<const float4*>(peers.params[rank] + i);
    float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
    p.x = adam_one(p.x, s.x * gscale, mm.x, vv.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
    p.y = adam_one(p.y, s.y * gscale, mm.y, vv.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
    p.z = adam_one(p.z, s.z * gscale, mm.z, vv.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
    p.w = adam_one(p.w, s.w * gscale, mm.w, vv.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
#pragma unroll
    for (int r = 0; r < WORLD; ++r) *reinterpret_cast<float4*>(peers.params[r] + i) = p;   // own arena included
  }
}

size_t up256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

struct vmb_dp {
  int rank = 0, world = 1, device = 0;
  long long n = 0, npad = 0, chunk = 0;
  char* arena = nullptr;          // this rank's allocation
  size_t off_grads[2] = {0, 0}, off_flags = 0, bytes = 0;
  char* peer[kMaxRanks] = {};     // every rank's arena as mapped here (peer[rank] == arena)
  bool connected = false;
  unsigned long long epoch = 0;
  int* timed_out = nullptr;       // device flag
};

extern "C" {

void vmb_dp_slice(long long n_params, int world, int rank, long long* begin, long long* end) {
  // equal chunks of a multiple of 4 elements (16-byte vector accesses); the padded tail belongs to the last rank
  const long long npad = (n_params + 3) / 4 * 4;
  long long chunk = world > 0 ? (npad / 4 + world - 1) / world * 4 : npad;
  long long b = std::min<long long>(npad, chunk * rank), e = std::min<long long>(npad, chunk * (rank + 1));
  if (begin) *begin = b;
  if (end) *end = e;
}

int vmb_dp_create(vmb_dp_t** handle, long long n_params, int rank, int world, void* ipc_handle_out) {
  if (!handle || !ipc_handle_out) return fail("vmb_dp_create: null argument");
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail("vmb_dp_create: 1..8 ranks of one box");
  if (n_params < 1) return fail("vmb_dp_create: empty bucket");
  static_assert(sizeof(cudaIpcMemHandle_t) == VMB_DP_IPC_HANDLE_BYTES, "IPC handle size");
  vmb_dp* h = new vmb_dp();
  h->rank = rank;
  h->world = world;
  h->n = n_params;
  h->npad = (n_params + 3) / 4 * 4;
  cudaGetDevice(&h->device);
  const size_t arr = up256(size_t(h->npad) * 4);
  h->off_grads[0] = arr;
  h->off_grads[1] = 2 * arr;
  h->off_flags = 3 * arr;
  h->bytes = 3 * arr + size_t(kMaxRanks) * kFlagStride * 8;
  if (cudaMalloc(reinterpret_cast<void**>(&h->arena), h->bytes) != cudaSuccess ||
      cudaMemset(h->arena, 0, h->bytes) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&h->timed_out), 4) != cudaSuccess ||
      cudaMemset(h->timed_out, 0, 4) != cudaSuccess) {
    fail("vmb_dp_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(h->arena);
    cudaFree(h->timed_out);
    delete h;
    return 1;
  }
  cudaIpcMemHandle_t ipc;
  if (cudaIpcGetMemHandle(&ipc, h->arena) != cudaSuccess) {
    fail("vmb_dp_create: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(h->arena);
    cudaFree(h->timed_out);
    delete h;
    return 1;
  }
  memcpy(ipc_handle_out, &ipc, sizeof ipc);
  h->peer[rank] = h->arena;
  h->connected = world == 1;
  *handle = h;
  return 0;
}

int vmb_dp_connect(vmb_dp_t* h, const void* all_handles) {
  if (!h || !all_handles) return fail("vmb_dp_connect: null argument");
  if (h->connected) return 0;
  for (int r = 0; r < h->world; ++r) {
    if (r == h->rank) continue;
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, static_cast<const char*>(all_handles) + size_t(r) * sizeof ipc, sizeof ipc);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      char msg[160];
      snprintf(msg, sizeof msg, "rank %d cannot map rank %d's arena: %s", h->rank, r, cudaGetErrorString(cudaGetLastError()));
      return fail("vmb_dp_connect: %s", msg);
    }
    h->peer[r] = static_cast<char*>(p);
  }
  h->connected = true;
  return 0;
}

float* vmb_dp_params(vmb_dp_t* h) { return h ? reinterpret_cast<float*>(h->arena) : nullptr; }

float* vmb_dp_grads(vmb_dp_t* h, int parity) {
  return h ? reinterpret_cast<float*>(h->arena + h->off_grads[parity & 1]) : nullptr;
}

int vmb_dp_adam_step(vmb_dp_t* h, int parity, float* exp_avg, float* exp_avg_sq, float lr, float beta1, float beta2,
                     float eps, float weight_decay, long long step, void* stream) {
  if (!h || !exp_avg || !exp_avg_sq) return fail("vmb_dp_adam_step: null pointer");
  if (!h->connected) return fail("vmb_dp_adam_step: vmb_dp_connect has not been called");
  if (step < 1) return fail("vmb_dp_adam_step: step counts from 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DpPeers peers{};
  for (int r = 0; r < h->world; ++r) {
    peers.params[r] = reinterpret_cast<float*>(h->peer[r]);
    peers.grads[r] = reinterpret_cast<const float*>(h->peer[r] + h->off_grads[parity & 1]);
    peers.flags[r] = reinterpret_cast<unsigned long long*>(h->peer[r] + h->off_flags);
  }
  long long begin = 0, end = 0;
  vmb_dp_slice(h->n, h->world, h->rank, &begin, &end);
  const float bc1 = 1.f - std::pow(beta1, static_cast<float>(step));
  const float bc2 = 1.f - std::pow(beta2, static_cast<float>(step));
  const float gscale = 1.f / static_cast<float>(h->world);
  dp_barrier_kernel<<<1, 32, 0, st>>>(peers, h->rank, h->world, ++h->epoch, h->timed_out);
  vmb::count_launch();
  if (end > begin) {
    const unsigned grid = static_cast<unsigned>(std::min<long long>(((end - begin) / 4 + 255) / 256, 148LL * 4));
#define VMB_DP_LAUNCH(W)                                                                                                   \
  dp_reduce_adam_gather_kernel<W><<<grid, 256, 0, st>>>(peers, h->rank, begin, end, exp_avg, exp_avg_sq, lr, beta1, beta2, \
                                                         eps, weight_decay, bc1, std::sqrt(bc2), gscale)
    switch (h->world) {
      case 1: VMB_DP_LAUNCH(1); break;
      case 2: VMB_DP_LAUNCH(2); break;
      case 3: VMB_DP_LAUNCH(3); break;
      case 4: VMB_DP_LAUNCH(4); break;
      case 5: VMB_DP_LAUNCH(5); break;
      case 6: VMB_DP_LAUNCH(6); break;
      case 7: VMB_DP_LAUNCH(7); break;
      default: VMB_DP_LAUNCH(8); break;
    }
#undef VMB_DP_LAUNCH
    vmb::count_launch();
  }
  dp_barrier_kernel<<<1, 32, 0, st>>>(peers, h->rank, h->world, ++h->epoch, h->timed_out);
  vmb::count_launch();
  if (vmb::check_launch("dp_reduce_adam_gather_kernel")) return fail("vmb_dp_adam_step: %s", vmb::kernels_last_error());
  return 0;
}

int vmb_dp_status(vmb_dp_t* h) {
  if (!h) return fail("vmb_dp_status: null handle");
  int flag = 0;
  if (cudaMemcpy(&flag, h->timed_out, 4, cudaMemcpyDeviceToHost) != cudaSuccess)
    return fail("vmb_dp_status: %s", cudaGetErrorString(cudaGetLastError()));
  if (flag) return fail("vmb_dp_status: a cross-GPU barrier timed out (a rank did not arrive within 10 s)");
  return 0;
}

void vmb_dp_disconnect(vmb_dp_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < h->world; ++r)
    if (r != h->rank && h->peer[r]) {
      cudaIpcCloseMemHandle(h->peer[r]);
      h->peer[r] = nullptr;
    }
  h->connected = h->world == 1;
}

void vmb_dp_destroy(vmb_dp_t* h) {
  if (!h) return;
  vmb_dp_disconnect(h);
  cudaFree(h->arena);
  cudaFree(h->timed_out);
  delete h;
}

}  // extern "C"

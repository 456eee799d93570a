// Peer-memory exchange kernels for data-parallel BatchNorm (SURVEY.md §8e coupling 1): the per-channel sums of the
// recognizer's train-mode BatchNorm layers are summed over the ranks INSIDE the kernel that consumes them, through
// mailboxes in peer-mapped HBM (NVLink 5 / NVSwitch stores), instead of an NCCL launch + stream fork/join per layer.
//
// Protocol (flagged stores, as NCCL's LL): a mailbox entry is 16 bytes {value0, epoch, value1, epoch}; each 8-byte
// half carries its own flag, so an entry is valid as soon as both flags equal the epoch — no fences, no separate
// signal.  Every rank writes its entry i into the mailbox of EVERY rank (its own included) and then spins on its own
// mailbox until the world's entries for this epoch have landed; all ranks add the entries in rank order, so the sums
// are bit-identical everywhere.  Mailboxes are indexed [slot][epoch parity][source rank][entry]; a slot belongs to
// one call site (one BatchNorm layer, one direction), its epoch lives in device memory and advances by one per
// launch, so captured launches stay correct over CUDA-graph replays.  Parity double-buffering makes a slot safe to
// reuse back to back: a peer can only be one epoch ahead of the slowest rank.
// A rank that waits longer than HWG_PEER_TIMEOUT_NS sets *fault and carries on (never hangs the GPU).
#include "common.cuh"

namespace hwg {
namespace {

constexpr int PEER_ENTRIES = HWG_PEER_SLOT_VALUES / 2;      // 16-byte entries per (slot, parity, source)
constexpr unsigned long long HWG_PEER_TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;

struct PeerArgs {
  const unsigned long long* mailboxes;   // device array [world]: base address of rank r's mailbox in THIS address space
  int world, rank, slot;
  unsigned* epochs;                      // device [slots], local
  int* fault;                            // device, local
};

__device__ __forceinline__ void st_entry(uint4* p, float a, float b, unsigned e) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(e),
               "r"(__float_as_uint(b)), "r"(e)
               : "memory");
}
__device__ __forceinline__ uint4 ld_entry(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ size_t entry_offset(const PeerArgs& pa, unsigned epoch, int src, int i) {
  return ((size_t)(pa.slot * 2 + (int)(epoch & 1u)) * pa.world + src) * PEER_ENTRIES + i;
}

// Sum of (a, b) over the ranks for entry i of this launch's epoch.  Called by every thread that owns an entry.
__device__ __forceinline__ float2 peer_sum_entry(const PeerArgs& pa, unsigned epoch, int i, float a, float b) {
  for (int r = 0; r < pa.world; ++r)
    st_entry(reinterpret_cast<uint4*>(pa.mailboxes[r]) + entry_offset(pa, epoch, pa.rank, i), a, b, epoch);
  float sa = 0.f, sb = 0.f;
  const uint4* mine = reinterpret_cast<const uint4*>(pa.mailboxes[pa.rank]);
  unsigned long long t0 = 0;
  for (int r = 0; r < pa.world; ++r) {
    const uint4* src = mine + entry_offset(pa, epoch, r, i);
    uint4 v = ld_entry(src);
    unsigned spins = 0;
    while (v.y != epoch || v.w != epoch) {
      if ((++spins & 1023u) == 0) {
        const unsigned long long t = now_ns();
        if (t0 == 0) t0 = t;
        if (t - t0 > HWG_PEER_TIMEOUT_NS) { atomicExch(pa.fault, 1); break; }
      }
      v = ld_entry(src);
    }
    sa += __uint_as_float(v.x);
    sb += __uint_as_float(v.z);
  }
  return make_float2(sa, sb);
}

// ---- BatchNorm coefficients over the joint batch of the ranks (forward) ------------------------------------
__global__ void __launch_bounds__(PEER_ENTRIES)
bn_coeffs_peer_kernel(const float* __restrict__ stats, int N, int C, float global_count,
                      const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ rmean,
                      float* __restrict__ rvar, float momentum, float eps, float* __restrict__ coef,
                      float* __restrict__ save, PeerArgs pa) {
  const unsigned epoch = pa.epochs[pa.slot] + 1u;
  const int c = threadIdx.x;
  if (c < C) {
    float s1 = 0.f, s2 = 0.f;
    for (int n = 0; n < N; ++n) { s1 += stats[((size_t)n * C + c) * 2]; s2 += stats[((size_t)n * C + c) * 2 + 1]; }
    const float2 tot = peer_sum_entry(pa, epoch, c, s1, s2);
    const float mean = tot.x / global_count;
    const float var = fmaxf(tot.y / global_count - mean * mean, 0.f);
    if (rmean) {
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * var * (global_count / fmaxf(global_count - 1.f, 1.f));
    }
    const float rstd = rsqrtf(var + eps);
    const float a = (weight ? weight[c] : 1.f) * rstd;
    coef[2 * c] = a;
    coef[2 * c + 1] = (bias ? bias[c] : 0.f) - mean * a;
    if (save) { save[2 * c] = mean; save[2 * c + 1] = rstd; }
  }
  __syncthreads();                       // every thread has read the old epoch before it advances
  if (threadIdx.x == 0) pa.epochs[pa.slot] = epoch;
}

// ---- in-place sum of a small fp32 vector over the ranks (backward: the [C,2] sums of bn_bwd_reduce) ---------
__global__ void __launch_bounds__(PEER_ENTRIES)
peer_allreduce_kernel(const float* in, float* out, int n, PeerArgs pa) {   // in == out allowed
  const unsigned epoch = pa.epochs[pa.slot] + 1u;
  const int i = threadIdx.x;
  if (2 * i < n) {
    const bool two = 2 * i + 1 < n;
    const float2 tot = peer_sum_entry(pa, epoch, i, in[2 * i], two ? in[2 * i + 1] : 0.f);
    out[2 * i] = tot.x;
    if (two) out[2 * i + 1] = tot.y;
  }
  __syncthreads();
  if (threadIdx.x == 0) pa.epochs[pa.slot] = epoch;
}

int check_peer(const void* mailboxes, int world, int rank, int slot, int slots, const void* epochs, const void* fault,
               const char* who) {
  HWG_REQUIRE(mailboxes && epochs && fault, "%s: peer exchange state missing", who);
  HWG_REQUIRE(world >= 1 && world <= HWG_PEER_MAX_WORLD && rank >= 0 && rank < world, "%s: bad world/rank %d/%d", who,
              world, rank);
  HWG_REQUIRE(slot >= 0 && slot < slots, "%s: slot %d outside [0,%d)", who, slot, slots);
  return HWG_OK;
}

}  // namespace
}  // namespace hwg

using namespace hwg;

extern "C" int64_t hwg_peer_mailbox_bytes(int world, int slots) {
  if (world < 1 || world > HWG_PEER_MAX_WORLD || slots < 1) return -1;
  return (int64_t)slots * 2 * world * PEER_ENTRIES * 16;
}

extern "C" int hwg_peer_enable_access(int peer_device) {
  int dev = 0;
  HWG_CUDA(cudaGetDevice(&dev));
  if (dev == peer_device) return HWG_OK;
  int can = 0;
  HWG_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  HWG_REQUIRE(can, "hwg_peer_enable_access: device %d cannot access device %d", dev, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return HWG_OK; }
  HWG_CUDA(e);
  return HWG_OK;
}

extern "C" int hwg_bn_coeffs_peer(const float* stats, int N, int C, int64_t global_count, const float* weight,
                                  const float* bias, float* running_mean, float* running_var, float momentum,
                                  float eps, float* coef, float* save_mean_rstd, const uint64_t* peer_mailboxes,
                                  int world, int rank, int slot, int slots, uint32_t* epochs, int* fault,
                                  void* stream) {
  HWG_REQUIRE(stats && coef && N > 0 && C > 0 && global_count > 0, "hwg_bn_coeffs_peer: bad argument");
  HWG_REQUIRE(C <= PEER_ENTRIES, "hwg_bn_coeffs_peer: C=%d exceeds the %d entries of a mailbox slot", C, PEER_ENTRIES);
  HWG_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "hwg_bn_coeffs_peer: running_mean/var must come together");
  if (int rc = check_peer(peer_mailboxes, world, rank, slot, slots, epochs, fault, "hwg_bn_coeffs_peer")) return rc;
  PeerArgs pa{reinterpret_cast<const unsigned long long*>(peer_mailboxes), world, rank, slot, epochs, fault};
  bn_coeffs_peer_kernel<<<1, PEER_ENTRIES, 0, (cudaStream_t)stream>>>(stats, N, C, (float)global_count, weight, bias,
                                                                      running_mean, running_var, momentum, eps, coef,
                                                                      save_mean_rstd, pa);
  return check_launch("bn_coeffs_peer_kernel");
}

extern "C" int hwg_peer_allreduce_f32(const float* in, float* out, int n, const uint64_t* peer_mailboxes, int world,
                                      int rank, int slot, int slots, uint32_t* epochs, int* fault, void* stream) {
  HWG_REQUIRE(in && out && n > 0 && n <= HWG_PEER_SLOT_VALUES, "hwg_peer_allreduce_f32: n=%d outside (0,%d]", n,
              HWG_PEER_SLOT_VALUES);
  if (int rc = check_peer(peer_mailboxes, world, rank, slot, slots, epochs, fault, "hwg_peer_allreduce_f32")) return rc;
  PeerArgs pa{reinterpret_cast<const unsigned long long*>(peer_mailboxes), world, rank, slot, epochs, fault};
  peer_allreduce_kernel<<<1, PEER_ENTRIES, 0, (cudaStream_t)stream>>>(in, out, n, pa);
  return check_launch("peer_allreduce_kernel");
}

// Bulk-staged streaming for the HBM-bound element-wise / reduction passes (sm_100a).
//
// A register-only streaming loop keeps (threads x loads-in-flight x 16 B) bytes outstanding per SM; at the ~100
// registers these kernels need that is 16-64 KB and they measured 1.8-3.2 TB/s next to a 6.1 TB/s copy.  Here the
// TMA engine streams the inputs instead: one thread issues 1-D bulk copies (cp.async.bulk, global -> shared,
// completion on an mbarrier) ST_STAGES-1 chunks ahead into a shared-memory ring, every thread then reads its
// 16-byte items from shared memory.  Bytes in flight per SM = resident blocks x (ST_STAGES-1) x chunk bytes
// (3 blocks x 2 stages x 8-16 KB per SM), independent of the register budget.
#pragma once
#include "sm100.cuh"

namespace hwg {

constexpr int ST_THREADS = 256;
constexpr int ST_ITEMS = 2;                        // 16-byte items per thread per chunk
constexpr int ST_CHUNK = ST_THREADS * ST_ITEMS;    // items per chunk and stream (8 KB)
constexpr int ST_STAGES = 3;
constexpr int ST_BAR_SLOTS = (ST_STAGES + 1) / 2 * 2;   // mbarrier slots, padded so that what follows stays 16-byte aligned

__device__ __forceinline__ void bulk_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   sm100::smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(sm100::smem_u32(bar))
               : "memory");
}

// per-channel constants live in shared memory; volatile: re-read at every use instead of being hoisted into registers
__device__ __forceinline__ void ld8s(const float* p, float (&f)[8]) {
  // volatile: re-read at every use instead of being hoisted into 8 live registers per constant vector
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(a));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(a));
}
__device__ __forceinline__ void ld4s(const float* p, float (&f)[4]) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(a));
}
constexpr size_t stream_smem_bytes(int nstream) { return (size_t)ST_STAGES * nstream * ST_CHUNK * 16 + 128; }

// Streams `total` 16-byte items of NS input arrays through the ring; this block handles chunks
// first_chunk, first_chunk + chunk_step, ...  body(item, v[NS]) is called once per item by exactly one thread
// (thread t handles items chunk*ST_CHUNK + i*ST_THREADS + t).  `ring` must be 128-byte aligned shared memory of
// stream_smem_bytes(NS) - 128 bytes, `bars` ST_STAGES mbarriers; all threads of the block must call this.
template <int NS, class Body>
__device__ __forceinline__ void stream_chunks(const uint4* const (&src)[NS], long long total, long long first_chunk,
                                              long long chunk_step, unsigned char* ring, uint64_t* bars, Body body) {
  using namespace sm100;
  const long long nchunks = (total + ST_CHUNK - 1) / ST_CHUNK;
  const long long nk = first_chunk < nchunks ? (nchunks - first_chunk + chunk_step - 1) / chunk_step : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ST_STAGES; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](long long k) {
    const long long chunk = first_chunk + k * chunk_step;
    const long long rem = total - chunk * ST_CHUNK;
    const uint32_t bytes = (uint32_t)(rem < ST_CHUNK ? rem : ST_CHUNK) * 16u;
    const int st = (int)(k % ST_STAGES);
    mbar_expect_tx(&bars[st], bytes * NS);
#pragma unroll
    for (int q = 0; q < NS; ++q)
      bulk_load_1d(ring + ((size_t)st * NS + q) * (ST_CHUNK * 16), src[q] + chunk * ST_CHUNK, bytes, &bars[st]);
  };
  if (threadIdx.x == 0)
    for (long long k = 0; k < ST_STAGES - 1 && k < nk; ++k) issue(k);
  for (long long k = 0; k < nk; ++k) {
    if (threadIdx.x == 0 && k + ST_STAGES - 1 < nk) issue(k + ST_STAGES - 1);   // its stage was drained in iteration k-1
    const int st = (int)(k % ST_STAGES);
    mbar_wait(&bars[st], (uint32_t)((k / ST_STAGES) & 1));
    const long long chunk = first_chunk + k * chunk_step;
#pragma unroll
    for (int i = 0; i < ST_ITEMS; ++i) {
      const int local = i * ST_THREADS + (int)threadIdx.x;
      const long long item = chunk * ST_CHUNK + local;
      if (item < total) {
        uint4 v[NS];
#pragma unroll
        for (int q = 0; q < NS; ++q)
          v[q] = *reinterpret_cast<const uint4*>(ring + ((size_t)st * NS + q) * (ST_CHUNK * 16) + (size_t)local * 16);
        body(item, v);
      }
    }
    __syncthreads();   // every thread is done with stage st before it is refilled
  }
}

}  // namespace hwg

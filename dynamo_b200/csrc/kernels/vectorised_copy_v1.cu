// vectorized_copy.fatbin -- drop-in for the v1 block manager's embedded kernel module.
//
// Dynamo's v1 KVBM does not link libkvbm_kernels; it cuModuleLoadData()s a fatbin and looks up the symbol
// `vectorised_copy(void** src, void** dst, size_t size, int num_pairs)`, launching it with
// grid = min(1024, num_pairs), block = 256, 0 bytes of dynamic shared memory
// (/root/reference/lib/llm/src/block_manager/block/transfer/cuda.rs:85-153, :521-548); the module can be
// replaced at run time through DYN_FATBIN_PATH (cuda.rs:588-611).  This file builds that module for sm_100a
// with the same TMA ring as the library kernels, squeezed into the 48 KiB of static shared memory the
// fixed launch configuration allows: 8 warps x (2 slots x 2 KiB + descriptor ring + barriers).
#include "copy_engine.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kStages = 2;
constexpr uint32_t kTile = 2048;

struct PairGenV1 {
  void* const* src_ptrs;
  void* const* dst_ptrs;
  size_t copy_size;
  uint32_t tiles_per_pair;
  __device__ __forceinline__ void get(uint32_t item, kvbm::Piece& p) const
  {
    const uint32_t pair = item / tiles_per_pair;
    const uint32_t t = item - pair * tiles_per_pair;
    const size_t off = static_cast<size_t>(t) * kTile;
    p.src = static_cast<const uint8_t*>(src_ptrs[pair]) + off;
    p.dst[0] = static_cast<uint8_t*>(dst_ptrs[pair]) + off;
    const size_t left = copy_size - off;
    p.bytes = left < kTile ? static_cast<uint32_t>(left) : kTile;
    p.ndst = 1;
    p.layer = 0;
  }
};

}  // namespace

extern "C" __global__ void __launch_bounds__(256)
vectorised_copy(void** src, void** dst, size_t size, int num_pairs)
{
  using namespace kvbm;
  __shared__ __align__(128) uint8_t slots[kWarps][kStages][kTile];
  __shared__ __align__(16) uint8_t desc[kWarps][desc_bytes_per_warp(1)];
  __shared__ uint64_t bars[kWarps][kStages];
  const int warp = threadIdx.x >> 5;
  const int warps_here = blockDim.x >> 5;
  if (warp >= kWarps || size == 0 || num_pairs <= 0) return;  // any extra warps of an unexpected launch idle
  if ((threadIdx.x & 31) == 0) {
    for (int s = 0; s < kStages; ++s) ptx::mbar_init(ptx::smem_addr(&bars[warp][s]), 1);
    ptx::mbar_fence_init();
  }
  __syncwarp();
  const uint64_t tpp = (size + kTile - 1) / kTile;
  const uint64_t total64 = tpp * static_cast<uint64_t>(num_pairs);
  if (tpp >= (1ull << 31) || total64 >= (1ull << 32)) {
    // beyond the 32-bit item space of the ring: plain warp-strided SIMT copy of whole pairs
    for (int pair = blockIdx.x * warps_here + warp; pair < num_pairs; pair += gridDim.x * warps_here) {
      const uint8_t* s = static_cast<const uint8_t*>(src[pair]);
      uint8_t* d = static_cast<uint8_t*>(dst[pair]);
      for (size_t off = 0; off < size; off += (1u << 30)) {
        const size_t left = size - off;
        warp_copy_simt(d + off, s + off, left < (1u << 30) ? static_cast<uint32_t>(left) : (1u << 30), threadIdx.x & 31);
      }
    }
    return;
  }
  PairGenV1 gen{src, dst, size, static_cast<uint32_t>(tpp)};
  StreamSync ss{};
  ss.layer_end = 1;
  RingParams rp{kStages, 1, kTile, 0, true, 0, 0};
  const int nw = warps_here < kWarps ? warps_here : kWarps;
  const uint32_t first = warp * gridDim.x + blockIdx.x;
  warp_ring<0>(gen, first, gridDim.x * nw, static_cast<uint32_t>(total64), &slots[warp][0][0], nullptr, &bars[warp][0],
               &desc[warp][0], 1, rp, ss);
}

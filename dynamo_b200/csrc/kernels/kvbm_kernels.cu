// libkvbm_kernels.so -- sm_100a KV-block transfer kernels behind the reference C ABI
// (see include/kvbm_kernels.h for the contract and the reference lines each entry replaces).
#include <cuda_runtime_api.h>

#include "../../../include/kvbm_kernels.h"
#include "copy_engine.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <vector>

namespace kvbm {

// ------------------------------------------------------------------------------------------------
// launch bookkeeping
// ------------------------------------------------------------------------------------------------
static std::atomic<uint64_t> g_launches{0};

struct DeviceInfo {
  int sm_count = 0;
  int max_smem_optin = 0;
};

static cudaError_t preload_kernels();

static cudaError_t device_info(DeviceInfo* out)
{
  constexpr int kMaxDev = 64;
  static std::atomic<int> sm[kMaxDev];
  static std::atomic<int> smem[kMaxDev];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDev) return cudaErrorInvalidDevice;
  int s = sm[dev].load(std::memory_order_acquire);
  if (s == 0) {
    int a = 0, b = 0;
    if ((e = cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&b, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    // CUDA loads kernels lazily and a first-time load can wait for the device to go idle.  A transfer kernel
    // that is spinning on layer-ready flags would then deadlock against the very first launch of the tiny
    // flag kernel that is supposed to release it -- so every kernel of this library is loaded up front.
    if ((e = preload_kernels()) != cudaSuccess) return e;
    smem[dev].store(b, std::memory_order_release);
    sm[dev].store(a, std::memory_order_release);
    s = a;
  }
  out->sm_count = s;
  out->max_smem_optin = smem[dev].load(std::memory_order_acquire);
  return cudaSuccess;
}

static int default_ctas(const DeviceInfo& di) { return (di.sm_count + 1) / 2; }

// Ring geometry of one launch.
struct RingCfg {
  int warps;        // W
  int stages;       // S (input slots per warp)
  int pending;      // P (stores allowed to keep draining their slot); loads ahead A = S - P
  uint32_t tile;    // source bytes per slot
  uint32_t smem;    // dynamic shared memory bytes
  uint32_t out_tile;  // cast only
};

constexpr uint32_t kBarBytesPerWarp = kMaxStages * 8;
// defaults tuned on B200 (profiles/r01_sweep_*.json)
constexpr int kDefaultWarps = 4;
constexpr int kDefaultStages = 3;
constexpr uint32_t kDefaultTile = 16384;

static uint32_t round_up(uint32_t x, uint32_t m) { return (x + m - 1) / m * m; }

// cast: 0 none, 1 up (out = 2x), 2 down (out = x/2)
static RingCfg make_ring(const DeviceInfo& di, uint32_t unit_bytes, int warps, int stages, int tile,
                         int cast, int pending = 0, int ndst = 1)
{
  RingCfg c{};
  c.warps = warps > 0 ? std::min(warps, 16) : kDefaultWarps;
  // default tile: the whole unit when it is small, else kDefaultTile pieces
  uint32_t t = tile > 0 ? static_cast<uint32_t>(tile) : std::min<uint32_t>(std::max<uint32_t>(unit_bytes, 16), kDefaultTile);
  t = round_up(t, 32);
  const uint32_t budget = static_cast<uint32_t>(di.max_smem_optin) - 1024;
  for (;;) {
    const uint32_t out = cast == 1 ? 2 * t : (cast == 2 ? t / 2 : 0);
    const uint32_t fixed = c.warps * (kBarBytesPerWarp + desc_bytes_per_warp(ndst) + 2 * out);
    int s = stages > 0 ? stages : kDefaultStages;
    s = std::min(s, kMaxStages);
    while (s > 2 && fixed + c.warps * s * t > budget) --s;
    if (fixed + c.warps * s * t <= budget) {
      c.stages = s;
      c.pending = pending > 0 ? std::min(pending, s - 1) : std::max(1, s / 2);
      c.tile = t;
      c.out_tile = out;
      c.smem = fixed + c.warps * s * t;
      return c;
    }
    if (t > 1024)
      t = round_up(t / 2, 32);
    else if (c.warps > 1)
      c.warps /= 2;
    else {
      c.stages = 2;
      c.pending = 1;
      c.tile = t;
      c.out_tile = out;
      c.smem = fixed + 2 * t;
      return c;
    }
  }
}

// shared-memory carve-up: [bars: W * kMaxStages * 8][descriptor rings: W * desc][in slots: W * S * tile][out slots: W * 2 * out_tile]
struct SmemView {
  uint64_t* bars;
  uint8_t* desc;
  uint8_t* in;
  uint8_t* out;
};

__device__ __forceinline__ SmemView carve(uint8_t* base, int W, int S, uint32_t tile, uint32_t out_tile, int ndst)
{
  const int warp = threadIdx.x >> 5;
  SmemView v;
  v.bars = reinterpret_cast<uint64_t*>(base) + warp * kMaxStages;
  const uint32_t db = desc_bytes_per_warp(ndst);
  v.desc = base + W * kBarBytesPerWarp + warp * db;
  uint8_t* in0 = base + W * (kBarBytesPerWarp + db);
  v.in = in0 + static_cast<size_t>(warp) * S * tile;
  v.out = in0 + static_cast<size_t>(W) * S * tile + static_cast<size_t>(warp) * 2 * out_tile;
  return v;
}

__device__ __forceinline__ void init_bars(uint64_t* bars, int S)
{
  if ((threadIdx.x & 31) == 0) {
    for (int s = 0; s < S; ++s) ptx::mbar_init(ptx::smem_addr(bars + s), 1);
    ptx::mbar_fence_init();
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// K1: pointer-pair copy (legacy ABI).  item = (pair, tile-within-pair)
// ------------------------------------------------------------------------------------------------
struct PairGen {
  void* const* src_ptrs;
  void* const* dst_ptrs;
  size_t copy_size;
  uint32_t tiles_per_pair;
  uint32_t tile;
  __device__ __forceinline__ void get(uint32_t item, Piece& p) const
  {
    const uint32_t pair = item / tiles_per_pair;
    const uint32_t t = item - pair * tiles_per_pair;
    const size_t off = static_cast<size_t>(t) * tile;
    p.src = static_cast<const uint8_t*>(src_ptrs[pair]) + off;
    p.dst[0] = static_cast<uint8_t*>(dst_ptrs[pair]) + off;
    const size_t left = copy_size - off;
    p.bytes = left < tile ? static_cast<uint32_t>(left) : tile;
    p.ndst = 1;
    p.layer = 0;
  }
};

__global__ void __launch_bounds__(512, 1)
kvbm_pair_copy_kernel(PairGen gen, uint32_t total, int S, int P, uint32_t tile, int allow_tma)
{
  extern __shared__ __align__(128) uint8_t smem[];
  const int W = blockDim.x >> 5;
  SmemView v = carve(smem, W, S, tile, 0, 1);
  init_bars(v.bars, S);
  StreamSync ss{};
  ss.layer_end = 1;
  RingParams rp{S, P, tile, 0, allow_tma != 0, 0, 0};
  // interleave warps of different CTAs over neighbouring items: item i -> CTA (i % grid), warp (i / grid) % W
  const uint32_t first = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
  warp_ring<0>(gen, first, gridDim.x * W, total, v.in, v.out, v.bars, v.desc, 1, rp, ss);
}

// ------------------------------------------------------------------------------------------------
// v2: block-table (paged) gather -> push -> scatter, addresses computed on the device
// ------------------------------------------------------------------------------------------------
struct PagedArgs {
  kvbm_paged_layout src;
  kvbm_paged_layout dst[kMaxDst];
  const int32_t* src_ids[kMaxDst];
  const int32_t* dst_ids[kMaxDst];
  int ndst;        // destinations
  int replicate;   // 1: all destinations share src_ids[0] -> read once, store ndst times
  uint32_t n_blocks;
  uint32_t layer_begin, n_layers;  // layers [layer_begin, layer_begin + n_layers)
  uint32_t outer;
  uint32_t tiles_per_region;
  uint32_t tile;      // source bytes per piece
  uint32_t dst_num, dst_den;  // destination bytes = source bytes * dst_num / dst_den (cast)
};

struct PagedGen {
  PagedArgs a;
  // item order (layer-major so a layer completes early for streaming):
  //   item = (((layer * n_blocks + block) * outer + o) * fan + d) * tiles_per_region + t
  // where fan = ndst for distinct payloads (destinations interleaved -> all NVLink ports busy) and 1
  // when replicating.
  __device__ __forceinline__ void get(uint32_t item, Piece& p) const
  {
    uint32_t r = item / a.tiles_per_region;
    const uint32_t t = item - r * a.tiles_per_region;
    uint32_t d = 0;
    if (!a.replicate) {
      const uint32_t r2 = r / a.ndst;
      d = r - r2 * a.ndst;
      r = r2;
    }
    uint32_t r2 = r / a.outer;
    const uint32_t o = r - r2 * a.outer;
    r = r2;
    r2 = r / a.n_blocks;
    const uint32_t blk = r - r2 * a.n_blocks;
    const uint32_t layer = a.layer_begin + r2;

    const uint64_t soff = static_cast<uint64_t>(t) * a.tile;
    const uint32_t left = a.src.region_bytes - static_cast<uint32_t>(soff);
    p.bytes = left < a.tile ? left : a.tile;
    p.layer = static_cast<int>(layer);
    const uint64_t doff = soff * a.dst_num / a.dst_den;
    const int32_t sb = __ldg(a.src_ids[d] + blk);
    p.src = reinterpret_cast<const uint8_t*>(__ldg(a.src.layer_base + layer) + static_cast<uint64_t>(sb) * a.src.block_stride +
                                             static_cast<uint64_t>(o) * a.src.outer_stride + soff);
    if (a.replicate) {
      p.ndst = a.ndst;
#pragma unroll
      for (int k = 0; k < kMaxDst; ++k) {
        if (k < a.ndst) {
          const int32_t db = __ldg(a.dst_ids[k] + blk);
          p.dst[k] = reinterpret_cast<uint8_t*>(__ldg(a.dst[k].layer_base + layer) +
                                                static_cast<uint64_t>(db) * a.dst[k].block_stride +
                                                static_cast<uint64_t>(o) * a.dst[k].outer_stride + doff);
        }
      }
    } else {
      p.ndst = 1;
      const int32_t db = __ldg(a.dst_ids[d] + blk);
      // select destination d without dynamic indexing of a kernel-parameter array of structs
      const uint64_t* lb = a.dst[0].layer_base;
      uint64_t bs = a.dst[0].block_stride, os = a.dst[0].outer_stride;
#pragma unroll
      for (int k = 1; k < kMaxDst; ++k)
        if (k == static_cast<int>(d)) {
          lb = a.dst[k].layer_base;
          bs = a.dst[k].block_stride;
          os = a.dst[k].outer_stride;
        }
      p.dst[0] = reinterpret_cast<uint8_t*>(__ldg(lb + layer) + static_cast<uint64_t>(db) * bs +
                                            static_cast<uint64_t>(o) * os + doff);
    }
  }
};

struct PagedSyncArgs {
  const uint32_t* layer_ready;
  uint32_t* workspace;
  uint32_t* done_flag[kMaxDst];
  uint32_t* layer_done[kMaxDst];
  uint32_t* completion_flag;
  uint32_t completion_value;
  uint32_t epoch;
  int num_layers_total;
  uint64_t gate_timeout_ns;
  int num_flag_dsts;  // destinations that get flags (== data destinations unless multicast)
};

__device__ __forceinline__ StreamSync make_sync(const PagedSyncArgs& s, const PagedArgs& a, int W)
{
  StreamSync ss{};
  ss.gate = s.layer_ready != nullptr;
  ss.want_layers = false;
  ss.want_done = s.completion_flag != nullptr || s.layer_ready != nullptr;  // gated launches always count their warps (abort cleanup)
  ss.layer_ready = s.layer_ready;
  ss.workspace = s.workspace;
  ss.epoch = s.epoch;
  ss.gate_timeout_ns = s.gate_timeout_ns;
  ss.completion_flag = s.completion_flag;
  ss.completion_value = s.completion_value;
  ss.total_warps = gridDim.x * W;
  ss.ndst = s.num_flag_dsts;
  ss.num_layers = s.num_layers_total;
  ss.layer_begin = static_cast<int>(a.layer_begin);
  ss.layer_end = static_cast<int>(a.layer_begin + a.n_layers);
#pragma unroll
  for (int d = 0; d < kMaxDst; ++d) {
    ss.done_flag[d] = s.done_flag[d];
    ss.layer_done[d] = s.layer_done[d];
    if (d < s.num_flag_dsts && s.layer_done[d] != nullptr) ss.want_layers = true;
    if (d < s.num_flag_dsts && s.done_flag[d] != nullptr) ss.want_done = true;
  }
  return ss;
}

template <int CAST>
__global__ void __launch_bounds__(512, 1)
kvbm_paged_copy_kernel(const __grid_constant__ PagedGen gen, const __grid_constant__ PagedSyncArgs sync,
                       uint32_t total, int S, int P, uint32_t out_tile, int allow_tma, int cache_hint, int variant)
{
  extern __shared__ __align__(128) uint8_t smem[];
  const int W = blockDim.x >> 5;
  const uint32_t tile = gen.a.tile;
  const int ring_ndst = gen.a.replicate ? gen.a.ndst : 1;
  SmemView v = carve(smem, W, S, tile, out_tile, ring_ndst);
  init_bars(v.bars, S);
  const StreamSync ss = make_sync(sync, gen.a, W);
  const uint32_t first = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
  const uint32_t stride = gridDim.x * W;
  RingParams rp{S, P, tile, out_tile, allow_tma != 0, cache_hint, CAST == KVBM_CAST_NONE ? variant : 0};
  warp_ring<CAST>(gen, first, stride, total, v.in, v.out, v.bars, v.desc, ring_ndst, rp, ss);
}

// ------------------------------------------------------------------------------------------------
// K2 / K3: block stacks <-> universal.  Index peel identical in meaning to the reference
// (tensor_kernels.cu:150-228) but done per 16-byte vector when a head row allows it, and with
// 32-bit arithmetic; one thread moves VEC bytes.
// ------------------------------------------------------------------------------------------------
template <int VEC, bool TO_UNIVERSAL>
__global__ void __launch_bounds__(256)
kvbm_permute_kernel(void* const* universal_ptrs, void* const* block_ptrs, uint64_t total_units,
                    uint32_t units_per_block, uint32_t nh, uint32_t nl, uint32_t no, uint32_t nt,
                    uint32_t hdv /* row length in VEC units */, int layout)
{
  const uint64_t stride = static_cast<uint64_t>(blockDim.x) * gridDim.x;
  for (uint64_t u = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < total_units; u += stride) {
    const uint32_t b = static_cast<uint32_t>(u / units_per_block);
    const uint32_t residual = static_cast<uint32_t>(u - static_cast<uint64_t>(b) * units_per_block);
    uint32_t tmp = residual;
    const uint32_t hd_i = tmp % hdv;
    tmp /= hdv;
    const uint32_t nt_i = tmp % nt;
    tmp /= nt;
    const uint32_t no_i = tmp % no;
    tmp /= no;
    const uint32_t nl_i = tmp % nl;
    const uint32_t nh_i = tmp / nl;
    const uint32_t chunk_off = layout == KVBM_BLOCK_LAYOUT_NHD ? ((nt_i * nh) + nh_i) * hdv + hd_i
                                                               : ((nh_i * nt) + nt_i) * hdv + hd_i;
    uint8_t* chunk = static_cast<uint8_t*>(block_ptrs[static_cast<size_t>(b) * nl * no + nl_i * no + no_i]) +
                     static_cast<size_t>(chunk_off) * VEC;
    uint8_t* uni = static_cast<uint8_t*>(universal_ptrs[b]) + static_cast<size_t>(residual) * VEC;
    uint8_t* d = TO_UNIVERSAL ? uni : chunk;
    const uint8_t* s = TO_UNIVERSAL ? chunk : uni;
    const bool aligned = ((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(s)) & (VEC - 1)) == 0;
    if (!aligned) {
      // buffers only promise element alignment (>= 2 B): move the unit as 16-bit pieces
#pragma unroll
      for (int k = 0; k < VEC / 2; ++k)
        reinterpret_cast<uint16_t*>(d)[k] = reinterpret_cast<const uint16_t*>(s)[k];
    } else if (VEC == 16)
      *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
    else if (VEC == 8)
      *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(s);
    else if (VEC == 4)
      *reinterpret_cast<uint32_t*>(d) = *reinterpret_cast<const uint32_t*>(s);
    else
      *reinterpret_cast<uint16_t*>(d) = *reinterpret_cast<const uint16_t*>(s);
  }
}

static size_t dtype_size(int dtype)
{
  switch (dtype) {
    case KVBM_DTYPE_F16:
    case KVBM_DTYPE_BF16:
      return 2;
    case KVBM_DTYPE_F32:
      return 4;
    case KVBM_DTYPE_F64:
      return 8;
    default:
      return 0;
  }
}

template <bool TO_UNIVERSAL>
static cudaError_t launch_permute(void* const* universal_ptrs, void* const* block_ptrs, size_t num_blocks,
                                  size_t nh, size_t nl, size_t no, size_t nt, size_t hd, int dtype,
                                  int layout, cudaStream_t stream)
{
  const size_t elem = dtype_size(dtype);
  if (elem == 0) return cudaErrorInvalidValue;  // tensor_kernels.cu:327
  const size_t total_per_block = nh * nl * no * nt * hd;
  if (total_per_block * num_blocks == 0) return cudaSuccess;  // tensor_kernels.cu:238-240
  if (!block_ptrs || !universal_ptrs) return cudaErrorInvalidValue;
  if (layout != KVBM_BLOCK_LAYOUT_NHD && layout != KVBM_BLOCK_LAYOUT_HND) return cudaErrorInvalidValue;
  if (total_per_block * elem >= (1ull << 32)) return cudaErrorInvalidValue;

  // The vector must divide a head row AND every buffer must be aligned to it.  Chunk/universal
  // buffers come from cudaMalloc / torch (>= 256 B aligned) in practice, but the ABI takes raw
  // pointers, so only widen past the element size when the row allows it and keep 2/4/8-byte
  // vectors naturally aligned with the element type.
  const size_t row = hd * elem;
  int vec = static_cast<int>(elem);
  if (row % 16 == 0)
    vec = 16;
  else if (row % 8 == 0 && elem <= 8)
    vec = std::max<int>(vec, 8);
  else if (row % 4 == 0 && elem <= 4)
    vec = std::max<int>(vec, 4);
  const uint32_t hdv = static_cast<uint32_t>(row / vec);
  const uint32_t units_per_block = static_cast<uint32_t>(total_per_block * elem / vec);
  const uint64_t total_units = static_cast<uint64_t>(units_per_block) * num_blocks;
  DeviceInfo di;
  cudaError_t e = device_info(&di);
  if (e != cudaSuccess) return e;
  const uint64_t want = (total_units + 255) / 256;
  const int grid = static_cast<int>(std::min<uint64_t>(want, static_cast<uint64_t>(di.sm_count) * 16));
  auto go = [&](auto kern) {
    kern<<<grid, 256, 0, stream>>>(universal_ptrs, block_ptrs, total_units, units_per_block,
                                   static_cast<uint32_t>(nh), static_cast<uint32_t>(nl),
                                   static_cast<uint32_t>(no), static_cast<uint32_t>(nt), hdv, layout);
  };
  switch (vec) {
    case 16: go(kvbm_permute_kernel<16, TO_UNIVERSAL>); break;
    case 8: go(kvbm_permute_kernel<8, TO_UNIVERSAL>); break;
    case 4: go(kvbm_permute_kernel<4, TO_UNIVERSAL>); break;
    default: go(kvbm_permute_kernel<2, TO_UNIVERSAL>); break;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// tiny flag kernels
// ------------------------------------------------------------------------------------------------
__global__ void kvbm_set_flags_kernel(uint32_t* flags, int first, int count, uint32_t value)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    __threadfence_system();
    ptx::st_release_sys(flags + first + i, value);
  }
}
__global__ void kvbm_wait_flag_kernel(const uint32_t* flag, uint32_t value)
{
  while (ptx::ld_acquire_sys(flag) < value) __nanosleep(100);
}

static cudaError_t preload_kernels()
{
  cudaFuncAttributes attr;
  cudaError_t e;
#define KVBM_PRELOAD(k) \
  if ((e = cudaFuncGetAttributes(&attr, k)) != cudaSuccess) return e;
  KVBM_PRELOAD(kvbm_pair_copy_kernel)
  KVBM_PRELOAD(kvbm_paged_copy_kernel<KVBM_CAST_NONE>)
  KVBM_PRELOAD(kvbm_paged_copy_kernel<KVBM_CAST_FP8E4M3_TO_BF16>)
  KVBM_PRELOAD(kvbm_paged_copy_kernel<KVBM_CAST_BF16_TO_FP8E4M3>)
  KVBM_PRELOAD(kvbm_set_flags_kernel)
  KVBM_PRELOAD(kvbm_wait_flag_kernel)
  KVBM_PRELOAD((kvbm_permute_kernel<16, true>))
  KVBM_PRELOAD((kvbm_permute_kernel<8, true>))
  KVBM_PRELOAD((kvbm_permute_kernel<4, true>))
  KVBM_PRELOAD((kvbm_permute_kernel<2, true>))
  KVBM_PRELOAD((kvbm_permute_kernel<16, false>))
  KVBM_PRELOAD((kvbm_permute_kernel<8, false>))
  KVBM_PRELOAD((kvbm_permute_kernel<4, false>))
  KVBM_PRELOAD((kvbm_permute_kernel<2, false>))
#undef KVBM_PRELOAD
  return cudaSuccess;
}

// Stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32) through the driver entry point:
// the stream's front-end writes / polls the word itself -- no kernel launch, no SM occupied by a spinner.
typedef int (*stream_memop_fn)(cudaStream_t, unsigned long long, uint32_t, unsigned int);
static stream_memop_fn g_write32 = nullptr, g_wait32 = nullptr;
static std::atomic<int> g_memops_state{0};  // 0 unknown, 1 available, -1 unavailable

static bool stream_memops()
{
  int st = g_memops_state.load(std::memory_order_acquire);
  if (st == 0) {
    void *w = nullptr, *q = nullptr;
    cudaDriverEntryPointQueryResult r1, r2;
    bool ok = cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &r1) == cudaSuccess && w &&
              cudaGetDriverEntryPoint("cuStreamWaitValue32", &q, cudaEnableDefault, &r2) == cudaSuccess && q;
    (void)cudaGetLastError();
    g_write32 = reinterpret_cast<stream_memop_fn>(w);
    g_wait32 = reinterpret_cast<stream_memop_fn>(q);
    st = ok ? 1 : -1;
    g_memops_state.store(st, std::memory_order_release);
  }
  return st == 1;
}

template <class K>
static cudaError_t set_smem(K kernel, uint32_t bytes)
{
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
}

}  // namespace kvbm

using namespace kvbm;

// =================================================================================================
// C ABI -- Part 1 (reference symbols)
// =================================================================================================
extern "C" cudaError_t
kvbm_kernels_launch_vectorized_copy(void** src_ptrs, void** dst_ptrs, size_t copy_size_bytes, int num_pairs,
                                    cudaStream_t stream)
{
  if (num_pairs == 0 || copy_size_bytes == 0) return cudaSuccess;  // tensor_kernels.cu:555-557
  if (!src_ptrs || !dst_ptrs) return cudaErrorInvalidValue;        // :559-561
  if (num_pairs < 0) return cudaErrorInvalidValue;
  DeviceInfo di;
  cudaError_t e = device_info(&di);
  if (e != cudaSuccess) return e;

  const uint32_t unit = copy_size_bytes > (1u << 20) ? (1u << 20) : static_cast<uint32_t>(copy_size_bytes);
  RingCfg rc = make_ring(di, unit, 0, 0, 0, 0);
  const uint64_t tiles_per_pair = (copy_size_bytes + rc.tile - 1) / rc.tile;
  const uint64_t total = tiles_per_pair * static_cast<uint64_t>(num_pairs);
  if (tiles_per_pair >= (1ull << 31) || total >= (1ull << 32)) return cudaErrorInvalidValue;

  PairGen gen{src_ptrs, dst_ptrs, copy_size_bytes, static_cast<uint32_t>(tiles_per_pair), rc.tile};
  const uint64_t ctas_needed = (total + rc.warps - 1) / rc.warps;
  const int grid = static_cast<int>(std::min<uint64_t>(ctas_needed, static_cast<uint64_t>(default_ctas(di))));
  if ((e = set_smem(kvbm_pair_copy_kernel, rc.smem)) != cudaSuccess) return e;
  kvbm_pair_copy_kernel<<<grid, rc.warps * 32, rc.smem, stream>>>(gen, static_cast<uint32_t>(total), rc.stages,
                                                                 rc.pending, rc.tile, 1);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();  // :570
}

extern "C" cudaError_t
kvbm_kernels_memcpy_batch(const void* const* src_ptrs, void* const* dst_ptrs, size_t size_per_copy,
                          size_t num_copies, int mode, cudaStream_t stream)
{
  if (num_copies == 0 || size_per_copy == 0) return cudaSuccess;  // tensor_kernels.cu:396-398
  if (!src_ptrs || !dst_ptrs) return cudaErrorInvalidValue;       // :400-402

  auto one_by_one = [&]() -> cudaError_t {
    for (size_t i = 0; i < num_copies; ++i) {
      cudaError_t e = cudaMemcpyAsync(dst_ptrs[i], src_ptrs[i], size_per_copy, cudaMemcpyDefault, stream);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  if (mode == KVBM_MEMCPY_FALLBACK_ONLY) return one_by_one();

#if CUDART_VERSION >= 12090
  // Host tables are consumed before returning (the caller may free them immediately: cuda.rs:185-204).
  std::vector<void*> srcs(num_copies), dsts(num_copies);
  std::vector<size_t> sizes(num_copies, size_per_copy), attr_idx(num_copies, 0);
  for (size_t i = 0; i < num_copies; ++i) {
    srcs[i] = const_cast<void*>(src_ptrs[i]);
    dsts[i] = dst_ptrs[i];
  }
  cudaMemcpyAttributes attr = {};
  attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
#if CUDART_VERSION >= 13000
  cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), num_copies, &attr, attr_idx.data(), 1, stream);
#else
  size_t fail_idx = 0;
  cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), num_copies, &attr, attr_idx.data(), 1,
                                       &fail_idx, stream);
#endif
  if (e == cudaErrorNotSupported || e == cudaErrorInvalidValue) {
    if (mode == KVBM_MEMCPY_BATCH_WITHOUT_FALLBACK) return e;
    (void)cudaGetLastError();
    return one_by_one();
  }
  return e;
#else
  if (mode == KVBM_MEMCPY_BATCH_WITHOUT_FALLBACK) return cudaErrorNotSupported;
  return one_by_one();
#endif
}

extern "C" cudaError_t
kvbm_kernels_launch_universal_from_block(void* const* universal_ptrs, const void* const* block_ptrs,
                                         size_t num_blocks, size_t nh, size_t nl, size_t no, size_t nt, size_t hd,
                                         int dtype, int layout, cudaStream_t stream)
{
  return launch_permute<true>(universal_ptrs, const_cast<void* const*>(reinterpret_cast<const void* const*>(block_ptrs)),
                              num_blocks, nh, nl, no, nt, hd, dtype, layout, stream);
}

extern "C" cudaError_t
kvbm_kernels_launch_block_from_universal(const void* const* universal_ptrs, void* const* block_ptrs,
                                         size_t num_blocks, size_t nh, size_t nl, size_t no, size_t nt, size_t hd,
                                         int dtype, int layout, cudaStream_t stream)
{
  return launch_permute<false>(const_cast<void* const*>(reinterpret_cast<const void* const*>(universal_ptrs)), block_ptrs,
                               num_blocks, nh, nl, no, nt, hd, dtype, layout, stream);
}

extern "C" bool
kvbm_kernels_has_memcpy_batch_async(void)
{
#if CUDART_VERSION >= 12090
  return true;
#else
  return false;
#endif
}

extern "C" bool
kvbm_kernels_is_stub_build(void)
{
  return false;
}

// =================================================================================================
// C ABI -- Part 2 (v2 extensions)
// =================================================================================================
extern "C" cudaError_t
kvbm_kernels_paged_copy_v2(const kvbm_paged_layout* src, const kvbm_paged_dst* dsts, int num_dsts, int num_blocks,
                           int layer_begin, int layer_end, int cast_mode, const kvbm_paged_copy_opts* opts,
                           cudaStream_t stream)
{
  if (num_blocks == 0 || num_dsts == 0 || layer_end == layer_begin) return cudaSuccess;
  if (!src || !dsts) return cudaErrorInvalidValue;
  if (num_blocks < 0 || num_dsts < 0 || num_dsts > KVBM_MAX_DESTINATIONS) return cudaErrorInvalidValue;
  if (layer_begin < 0 || layer_end < layer_begin || static_cast<uint32_t>(layer_end) > src->num_layers)
    return cudaErrorInvalidValue;
  if (cast_mode < KVBM_CAST_NONE || cast_mode > KVBM_CAST_BF16_TO_FP8E4M3) return cudaErrorInvalidValue;
  if (!src->layer_base || src->region_bytes == 0) return cudaErrorInvalidValue;
  const uint32_t num = cast_mode == KVBM_CAST_FP8E4M3_TO_BF16 ? 2 : 1;
  const uint32_t den = cast_mode == KVBM_CAST_BF16_TO_FP8E4M3 ? 2 : 1;
  if (cast_mode == KVBM_CAST_BF16_TO_FP8E4M3 && (src->region_bytes & 1)) return cudaErrorInvalidValue;

  kvbm_paged_copy_opts o{};
  if (opts) o = *opts;

  // NVLS: dsts[0] is addressed through a multicast mapping and carries the payload for every bound device; the
  // other entries only name the flags of the remaining receivers.
  const bool mc = o.multicast != 0;
  if (mc && (cast_mode != KVBM_CAST_NONE || (src->region_bytes & 3) || (src->block_stride & 3) || (src->outer_stride & 3) ||
             (dsts[0].layout.block_stride & 15) || (dsts[0].layout.outer_stride & 15)))
    return cudaErrorInvalidValue;
  const int data_dsts = mc ? 1 : num_dsts;

  PagedGen gen{};
  PagedSyncArgs sync{};
  gen.a.src = *src;
  gen.a.ndst = data_dsts;
  gen.a.replicate = 1;
  sync.num_flag_dsts = num_dsts;
  for (int d = 0; d < num_dsts; ++d) {
    const kvbm_paged_dst& D = dsts[d];
    sync.done_flag[d] = D.done_flag;
    sync.layer_done[d] = D.layer_done_flags;
    if (d >= data_dsts) continue;
    // same compatibility rules as execute_cuda_transfer (executor/cuda.rs:52-67) + region size match
    // (executor/memcpy.rs:143-153), adjusted for the element-width change of a cast
    if (!D.layout.layer_base || !D.src_block_ids || !D.dst_block_ids) return cudaErrorInvalidValue;
    if (D.layout.num_layers != src->num_layers || D.layout.outer_dim != src->outer_dim) return cudaErrorInvalidValue;
    if (static_cast<uint64_t>(D.layout.region_bytes) * den != static_cast<uint64_t>(src->region_bytes) * num)
      return cudaErrorInvalidValue;
    gen.a.dst[d] = D.layout;
    gen.a.src_ids[d] = D.src_block_ids;
    gen.a.dst_ids[d] = D.dst_block_ids;
    if (D.src_block_ids != dsts[0].src_block_ids) gen.a.replicate = 0;
  }
  if (data_dsts == 1) gen.a.replicate = 1;
  sync.completion_flag = o.completion_flag;
  sync.completion_value = o.completion_value;
  sync.layer_ready = o.layer_ready_flags;
  sync.workspace = o.sync_workspace;
  sync.epoch = o.epoch;
  sync.num_layers_total = static_cast<int>(src->num_layers);
  sync.gate_timeout_ns = static_cast<uint64_t>(o.gate_timeout_ms > 0 ? o.gate_timeout_ms : 10000) * 1000000ull;
  if (o.layer_ready_flags && !o.sync_workspace) return cudaErrorInvalidValue;  // the abort word lives in the workspace
  bool needs_ws = o.completion_flag != nullptr;
  for (int d = 0; d < num_dsts; ++d)
    if (dsts[d].done_flag || dsts[d].layer_done_flags) needs_ws = true;
  if (needs_ws && !o.sync_workspace) return cudaErrorInvalidValue;

  DeviceInfo di;
  cudaError_t e = device_info(&di);
  if (e != cudaSuccess) return e;
  RingCfg rc = make_ring(di, src->region_bytes, o.warps_per_cta, o.stages, o.tile_bytes, cast_mode, o.stores_in_flight,
                         gen.a.replicate ? data_dsts : 1);

  gen.a.n_blocks = static_cast<uint32_t>(num_blocks);
  gen.a.layer_begin = static_cast<uint32_t>(layer_begin);
  gen.a.n_layers = static_cast<uint32_t>(layer_end - layer_begin);
  gen.a.outer = src->outer_dim;
  gen.a.tile = rc.tile;
  gen.a.tiles_per_region = (src->region_bytes + rc.tile - 1) / rc.tile;
  gen.a.dst_num = num;
  gen.a.dst_den = den;
  const uint64_t fan = gen.a.replicate ? 1 : static_cast<uint64_t>(data_dsts);
  const uint64_t total = static_cast<uint64_t>(gen.a.n_layers) * gen.a.n_blocks * gen.a.outer * fan * gen.a.tiles_per_region;
  if (total >= (1ull << 32)) return cudaErrorInvalidValue;

  const uint64_t ctas_needed = (total + rc.warps - 1) / rc.warps;
  // Default: one CTA per TPC (half the SMs).  Measured on B200 (profiles/r01_sweep_n1_fine.json): 74 CTAs x 4 warps
  // move 5.96 TB/s r+w on a same-GPU copy vs 5.70 with 148, and 32 CTAs already saturate NVLink -- and the other
  // half of the chip stays free for whatever the engine is running.
  int cap = o.max_ctas > 0 ? o.max_ctas : (cast_mode == KVBM_CAST_NONE ? default_ctas(di) : di.sm_count);  // the cast is ALU work: use every SM
  const int grid = static_cast<int>(std::min<uint64_t>(ctas_needed, static_cast<uint64_t>(cap)));
  const int allow_tma = o.force_simt ? 0 : 1;
  const uint32_t total32 = static_cast<uint32_t>(total);

  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t err = set_smem(kern, rc.smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, rc.warps * 32, rc.smem, stream>>>(gen, sync, total32, rc.stages, rc.pending, rc.out_tile, allow_tma,
                                                   o.cache_hint, mc ? (o.multicast == 2 ? 0 : 4) : o.variant);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
  };
  switch (cast_mode) {
    case KVBM_CAST_NONE: return launch(kvbm_paged_copy_kernel<KVBM_CAST_NONE>);
    case KVBM_CAST_FP8E4M3_TO_BF16: return launch(kvbm_paged_copy_kernel<KVBM_CAST_FP8E4M3_TO_BF16>);
    default: return launch(kvbm_paged_copy_kernel<KVBM_CAST_BF16_TO_FP8E4M3>);
  }
}

extern "C" cudaError_t
kvbm_kernels_set_flags(uint32_t* flags, int first, int count, uint32_t value, cudaStream_t stream)
{
  if (count == 0) return cudaSuccess;
  if (!flags || count < 0 || first < 0) return cudaErrorInvalidValue;
  DeviceInfo di;
  cudaError_t e = device_info(&di);
  if (e != cudaSuccess) return e;
  if (count == 1 && stream_memops()) {
    // CU_STREAM_WRITE_VALUE_DEFAULT (0): memory fence before the write => release semantics for prior work
    if (g_write32(stream, reinterpret_cast<unsigned long long>(flags + first), value, 0) == 0) return cudaSuccess;
  }
  kvbm_set_flags_kernel<<<(count + 127) / 128, 128, 0, stream>>>(flags, first, count, value);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

extern "C" cudaError_t
kvbm_kernels_wait_flag(const uint32_t* flag, uint32_t value, cudaStream_t stream)
{
  if (!flag) return cudaErrorInvalidValue;
  DeviceInfo di;
  cudaError_t e = device_info(&di);
  if (e != cudaSuccess) return e;
  if (stream_memops()) {
    // CU_STREAM_WAIT_VALUE_GEQ (0): wait until (int32)(*flag - value) >= 0
    if (g_wait32(stream, reinterpret_cast<unsigned long long>(flag), value, 0) == 0) return cudaSuccess;
  }
  kvbm_wait_flag_kernel<<<1, 1, 0, stream>>>(flag, value);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

extern "C" cudaError_t
kvbm_kernels_stream_wait_event(cudaStream_t stream, void* event)
{
  if (!event) return cudaErrorInvalidValue;
  return cudaStreamWaitEvent(stream, static_cast<cudaEvent_t>(event), 0);
}

extern "C" uint64_t
kvbm_kernels_launch_count(void)
{
  return g_launches.load(std::memory_order_relaxed);
}

extern "C" const char*
kvbm_kernels_build_info(void)
{
  return "libkvbm_kernels sm_100a tma=cp.async.bulk cast=e4m3<->bf16";
}

// The warp-level TMA ring that every KV transfer kernel in this library is built from.
//
// One warp owns `S` shared-memory slots of `tile` bytes and walks its share of the work items:
//
//     HBM (paged source) --cp.async.bulk--> smem slot --cp.async.bulk--> HBM / NVLink peer (1..N)
//
// Lane 0 drives both directions asynchronously (loads complete on per-slot mbarriers, stores are
// tracked by bulk async-groups); the other lanes only work for pieces that cannot use TMA (pointer
// or size not a multiple of 16 B -> vectorised SIMT ladder) and for the fused fp8<->bf16 cast.
// A CTA is W such warps, each with a private ring, so a single SM keeps W*(S-1) tiles in flight.
#pragma once
#include "ptx.cuh"

namespace kvbm {

constexpr int kMaxDst = 8;
constexpr int kMaxStages = 16;

// One unit of work: `bytes` (of source) from `src` to each of dst[0..ndst).
struct Piece {
  const uint8_t* src;
  uint8_t* dst[kMaxDst];
  uint32_t bytes;  // source bytes
  int ndst;
  int layer;  // for layer-streaming flags (0 when unused)
};

// ------------------------------------------------------------------------------------------
// SIMT ladder for pieces TMA cannot take.  Same alignment contract as the reference K1 kernel
// (/root/reference/lib/kvbm-kernels/cuda/tensor_kernels.cu:511-540): the widest vector both
// pointers allow, byte tail.  Whole warp cooperates, 4 independent 16 B loads in flight per lane.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_copy_simt(uint8_t* dst, const uint8_t* src, uint32_t bytes,
                                               int lane)
{
  const uintptr_t both = reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src);
  uint32_t done = 0;
  if ((both & 15) == 0) {
    const uint32_t n = bytes >> 4;
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    uint32_t i = lane;
    for (; i + 96 < n; i += 128) {
      uint4 a = ptx::ld_stream_v4(s + i), b = ptx::ld_stream_v4(s + i + 32);
      uint4 c = ptx::ld_stream_v4(s + i + 64), e = ptx::ld_stream_v4(s + i + 96);
      ptx::st_stream_v4(d + i, a);
      ptx::st_stream_v4(d + i + 32, b);
      ptx::st_stream_v4(d + i + 64, c);
      ptx::st_stream_v4(d + i + 96, e);
    }
    for (; i < n; i += 32) ptx::st_stream_v4(d + i, ptx::ld_stream_v4(s + i));
    done = n << 4;
  } else if ((both & 7) == 0) {
    const uint32_t n = bytes >> 3;
    const uint2* s = reinterpret_cast<const uint2*>(src);
    uint2* d = reinterpret_cast<uint2*>(dst);
#pragma unroll 4
    for (uint32_t i = lane; i < n; i += 32) d[i] = s[i];
    done = n << 3;
  } else if ((both & 3) == 0) {
    const uint32_t n = bytes >> 2;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
#pragma unroll 4
    for (uint32_t i = lane; i < n; i += 32) d[i] = s[i];
    done = n << 2;
  }
#pragma unroll 4
  for (uint32_t i = done + lane; i < bytes; i += 32) dst[i] = src[i];
}

__device__ __forceinline__ bool piece_tma_ok(const Piece& p)
{
  uintptr_t m = reinterpret_cast<uintptr_t>(p.src) | p.bytes;
#pragma unroll
  for (int d = 0; d < kMaxDst; ++d)
    if (d < p.ndst) m |= reinterpret_cast<uintptr_t>(p.dst[d]);
  return (m & 15) == 0 && p.bytes != 0;
}

// Layer-streaming hooks shared by the rings.  All members may be null.
struct StreamSync {
  const uint32_t* layer_ready;  // wait until layer_ready[l] >= epoch before reading layer l
  uint32_t* workspace;          // [num_layers + 1] zeroed counters (last = whole transfer)
  uint32_t* done_flag[kMaxDst];
  uint32_t* layer_done[kMaxDst];
  uint32_t* completion_flag;    // extra whole-transfer flag (host-mapped memory allowed)
  uint32_t completion_value;
  uint32_t epoch;
  uint32_t total_warps;
  int ndst;
  int num_layers;  // counters per layer live at workspace[l]; whole-transfer counter at [num_layers]
  int layer_begin, layer_end;
  uint64_t gate_timeout_ns;  // a gated wait longer than this aborts the transfer instead of spinning forever
  bool gate;         // layer_ready present: reads of a layer wait for its flag
  bool want_layers;  // some destination wants per-layer done flags
  bool want_done;    // some whole-transfer flag is wanted
};

// Ring geometry / behaviour of one launch (uniform across the grid).
struct RingParams {
  int S;              // input slots per warp
  int P;              // byte-exact ring: stores allowed to keep draining (loads ahead A = S - P)
  uint32_t tile_in;   // source bytes per slot
  uint32_t tile_out;  // cast rings: bytes per output slot (2 of them)
  bool allow_tma;
  int cache_hint;     // bit0 evict_first loads, bit1 evict_first stores
  int variant;        // 0 = TMA load + TMA store; 1 = TMA load + SIMT store from smem; 4 = TMA load + multimem.st (NVLS);
                      // 2 = loads only (diagnostic), 3 = stores only (diagnostic)
};

// Flag polls are relaxed loads; one acquire fence follows a successful probe (an ld.acquire.sys per poll would
// invalidate L1 every time and, measured, slowed the co-resident compute kernel).
// Returns false when the flag was not released within gate_timeout_ns: a spinning kernel that waits for work
// which can never be submitted (e.g. because a lazy module load is itself waiting for this kernel) must not hang
// the GPU; the warp then records the abort in the workspace and stops.
__device__ __forceinline__ bool wait_layer_ready(const StreamSync& ss, int layer, int lane)
{
  uint32_t ok = 1;
  if (lane == 0) {
    const uint64_t t0 = ptx::globaltimer_ns();
    uint32_t polls = 0;
    while (ptx::ld_relaxed_sys(ss.layer_ready + layer) < ss.epoch) {
      __nanosleep(128);
      if ((++polls & 1023u) == 0 && ptx::globaltimer_ns() - t0 > ss.gate_timeout_ns) {
        ok = 0;
        if (ss.workspace != nullptr) atomicExch(ss.workspace + ss.num_layers + 1, 1u);
        break;
      }
    }
    ptx::fence_acq_rel_sys();
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  return ok != 0;
}

// non-blocking probe, warp-uniform result
__device__ __forceinline__ bool layer_is_ready(const StreamSync& ss, int layer, int lane)
{
  uint32_t v = 0;
  if (lane == 0) {
    v = ptx::ld_relaxed_sys(ss.layer_ready + layer);
    if (v >= ss.epoch) ptx::fence_acq_rel_sys();
  }
  v = __shfl_sync(0xffffffffu, v, 0);
  return v >= ss.epoch;
}

// This warp has *completed* (stores landed) everything it owns in layers [from, to).
__device__ __forceinline__ void arrive_layers(const StreamSync& ss, int from, int to, int lane)
{
  if (lane != 0 || !ss.want_layers) return;
  for (int l = from; l < to; ++l) {
    uint32_t old = ptx::atom_add_acq_rel_gpu(ss.workspace + l, 1u);
    if (old == ss.total_warps - 1) {
      ss.workspace[l] = 0;  // leave the workspace zeroed for the next launch
      ptx::fence_acq_rel_sys();  // once per layer for the whole grid: everything acquired above is released below
#pragma unroll
      for (int d = 0; d < kMaxDst; ++d)
        if (d < ss.ndst && ss.layer_done[d] != nullptr) ptx::st_release_sys(ss.layer_done[d] + l, ss.epoch);
    }
  }
}

__device__ __forceinline__ void arrive_transfer(const StreamSync& ss, int lane)
{
  if (lane != 0 || !ss.want_done) return;
  uint32_t old = ptx::atom_add_acq_rel_gpu(ss.workspace + ss.num_layers, 1u);
  if (old == ss.total_warps - 1) {
    ss.workspace[ss.num_layers] = 0;
    const bool aborted = ss.gate && atomicExch(ss.workspace + ss.num_layers + 1, 0u) != 0;
    if (aborted)  // warps that gave up skipped their layer arrivals: leave every counter zeroed for the next launch
      for (int l = ss.layer_begin; l < ss.layer_end; ++l) ss.workspace[l] = 0;
    ptx::fence_acq_rel_sys();
    if (!aborted) {
#pragma unroll
      for (int d = 0; d < kMaxDst; ++d)
        if (d < ss.ndst && ss.done_flag[d] != nullptr) ptx::st_release_sys(ss.done_flag[d], ss.epoch);
    }
    // 0xFFFFFFFF = "this transfer gave up waiting for a layer": destinations are NOT told it completed
    if (ss.completion_flag != nullptr) ptx::st_release_sys(ss.completion_flag, aborted ? 0xFFFFFFFFu : ss.completion_value);
  }
}

// Every store this warp issued so far has landed and is ordered before the arrive that follows:
// lane 0 waits for its bulk groups and crosses the async->generic proxy; __syncwarp orders the other lanes'
// SIMT stores before lane 0's release atomic (which is cumulative).  No per-lane MEMBAR.SYS: on an SM shared
// with a compute kernel those stalled both kernels.
__device__ __forceinline__ void drain_stores(int lane)
{
  if (lane == 0) {
    ptx::bulk_wait<0>();
    ptx::fence_proxy_async_global();
  }
  __syncwarp();
}

// smem -> global by the whole warp (variant 1): 16 B per lane, 4 independent stores in flight
__device__ __forceinline__ void warp_store_from_smem(uint8_t* dst, uint32_t src_smem, uint32_t bytes, int lane)
{
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint32_t n = bytes >> 4;
  for (uint32_t i = lane; i < n; i += 32) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src_smem + i * 16));
    ptx::st_stream_v4(d + i, v);
  }
}

// smem -> NVLink multicast address (variant 4): one store, the switch delivers it to every bound device
__device__ __forceinline__ void warp_store_from_smem_mc(uint8_t* dst, uint32_t src_smem, uint32_t bytes, int lane)
{
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint32_t n = bytes >> 4;
  for (uint32_t i = lane; i < n; i += 32) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src_smem + i * 16));
    ptx::multimem_st_v4(d + i, v);
  }
}
// global -> multicast for pieces the TMA load cannot take (the launcher guarantees 4-byte granularity)
__device__ __forceinline__ void warp_copy_simt_mc(uint8_t* dst, const uint8_t* src, uint32_t bytes, int lane)
{
  const uintptr_t both = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | bytes;
  if ((both & 15) == 0) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (uint32_t i = lane; i < (bytes >> 4); i += 32) ptx::multimem_st_v4(d + i, ptx::ld_stream_v4(s + i));
  } else {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    for (uint32_t i = lane; i < (bytes >> 2); i += 32) ptx::multimem_st_b32(d + i, s[i]);
  }
}

template <bool UP>
__device__ __forceinline__ void convert_smem(uint32_t in_smem, uint32_t out_smem, uint32_t src_bytes, int lane);
template <bool UP>
__device__ __forceinline__ void warp_cast_simt(uint8_t* dst, const uint8_t* src, uint32_t src_bytes, int lane);

// ------------------------------------------------------------------------------------------
// Per-warp descriptor ring.  Address generation (a handful of integer divisions and two dependent
// table loads per item) is far too slow for the single instruction stream that drives the TMA
// pipeline -- with one warp per scheduler every dependent instruction costs its full latency, which
// measured ~2 us per item.  So the 32 lanes generate the descriptors of 32 consecutive items at
// once (SIMD), park them in shared memory, and the pipeline reads them back with 3-4 LDS per item.
// Layout (SoA, kDescRing entries each):  src[u64] | meta[u32 bytes|ok<<31, i32 layer] | dst[d][u64]...
// ------------------------------------------------------------------------------------------
constexpr int kDescRing = 64;
__host__ __device__ constexpr uint32_t desc_bytes_per_warp(int ndst) { return kDescRing * (16u + 8u * static_cast<uint32_t>(ndst)); }

struct DescRing {
  uint64_t* src;
  uint2* meta;
  uint64_t* dst;  // [ndst][kDescRing]
  __device__ __forceinline__ DescRing(uint8_t* mem)
      : src(reinterpret_cast<uint64_t*>(mem)),
        meta(reinterpret_cast<uint2*>(mem + kDescRing * 8)),
        dst(reinterpret_cast<uint64_t*>(mem + kDescRing * 16))
  {
  }
  __device__ __forceinline__ void put(uint32_t j, const Piece& p, bool ok) const
  {
    const uint32_t i = j & (kDescRing - 1);
    src[i] = reinterpret_cast<uint64_t>(p.src);
    meta[i] = make_uint2(p.bytes | (ok ? 0x80000000u : 0u), static_cast<uint32_t>(p.layer));
#pragma unroll
    for (int d = 0; d < kMaxDst; ++d)
      if (d < p.ndst) dst[d * kDescRing + i] = reinterpret_cast<uint64_t>(p.dst[d]);
  }
  // returns eligibility; fills src/bytes/layer (+dst when want_dst)
  __device__ __forceinline__ bool get(uint32_t j, int ndst, Piece& p, bool want_dst) const
  {
    const uint32_t i = j & (kDescRing - 1);
    const uint2 m = meta[i];
    p.src = reinterpret_cast<const uint8_t*>(src[i]);
    p.bytes = m.x & 0x7fffffffu;
    p.layer = static_cast<int>(m.y);
    p.ndst = ndst;
    if (want_dst) {
#pragma unroll
      for (int d = 0; d < kMaxDst; ++d)
        if (d < ndst) p.dst[d] = reinterpret_cast<uint8_t*>(dst[d * kDescRing + i]);
    }
    return (m.x & 0x80000000u) != 0;
  }
};

// ------------------------------------------------------------------------------------------
// The ring.  CAST: 0 byte-exact, 1 fp8->bf16, 2 bf16->fp8.  Gen::get(item, Piece&) is warp-uniform.
//   first/stride/total : this warp's arithmetic progression of item indices
//   in_slots           : S * tile_in bytes private to this warp;  out_slots: 2 * tile_out (cast only)
//   bars               : S mbarriers private to this warp (initialised, count 1)
// Loads are issued ahead without ever blocking on a layer's ready flag: when the item to be stored next
// is itself gated, the warp first drains and publishes the layers it has finished (so a consumer that
// releases layer l+1 only after seeing layer l done cannot deadlock), then blocks.
// ------------------------------------------------------------------------------------------
template <int CAST, class Gen>
__device__ __forceinline__ void warp_ring(const Gen& gen, uint32_t first, uint32_t stride, uint32_t total,
                                          uint8_t* in_slots, uint8_t* out_slots, uint64_t* bars,
                                          uint8_t* desc_mem, int ndst, const RingParams& rp,
                                          const StreamSync& ss)
{
  const int lane = threadIdx.x & 31;
  const int S = rp.S;
  const DescRing ring(desc_mem);
  uint32_t filled = 0;  // descriptors exist for this warp's items [max(0, filled - kDescRing), filled)
  const bool sync_slot = CAST != 0 || rp.variant == 1 || rp.variant == 2 || rp.variant == 4;  // input slot is released synchronously
  const uint32_t ahead = sync_slot ? static_cast<uint32_t>(S) : static_cast<uint32_t>(S - rp.P);
  const uint32_t n_my = total > first ? (total - first + stride - 1) / stride : 0;
  const uint32_t in0 = ptx::smem_addr(in_slots);
  const uint32_t out0 = ptx::smem_addr(out_slots);
  const uint32_t bar0 = ptx::smem_addr(bars);
  const uint64_t policy = rp.cache_hint ? ptx::policy_evict_first() : 0;
  uint32_t phase = 0;       // bit s = parity the next wait on slot s expects
  uint32_t next_load = 0;   // items [0, next_load) have had their load issued (or need none)
  int ready_layer = ss.gate ? ss.layer_begin - 1 : 0x7fffffff;
  int open_layer = ss.layer_begin;  // lowest layer this warp has not yet arrived for
  bool aborted = false;

  auto eligible = [&](const Piece& p) {
    bool ok = rp.allow_tma && piece_tma_ok(p);
    if (CAST != 0) ok = ok && (p.bytes & 31) == 0;  // both sides whole 16 B vectors
    return ok;
  };
  // Descriptor generation is software-pipelined: the table loads of the NEXT batch of 32 items are issued
  // kPrefetchLead items early (results stay in registers, nothing waits on them) and only written to the ring
  // when the batch is actually needed, so their DRAM/L2 latency never drains the TMA pipeline.
  constexpr uint32_t kPrefetchLead = 16;
  Piece pre;
  bool pre_valid = false;
  auto prefetch = [&]() {  // all 32 lanes
    const uint32_t j = filled + lane;
    pre.bytes = 0;
    pre.ndst = 0;
    pre.layer = 0;
    pre.src = nullptr;
    if (j < n_my) gen.get(first + j * stride, pre);
    pre_valid = true;
  };
  auto fill = [&]() {  // all 32 lanes: publish the prefetched descriptors of the next 32 items
    if (!pre_valid) prefetch();
    __syncwarp();  // lanes that ran ahead must not overwrite entries a slower lane still reads
    const uint32_t j = filled + lane;
    if (j < n_my) ring.put(j, pre, eligible(pre));
    filled += 32;
    pre_valid = false;
    __syncwarp();
  };
  auto pump = [&](uint32_t limit) {
    while (next_load < n_my && next_load < limit) {
      Piece p;
      const bool ok = ring.get(next_load, ndst, p, false);
      if (p.layer > ready_layer) {
        if (!layer_is_ready(ss, p.layer, lane)) break;
        ready_layer = p.layer;
      }
      if (rp.variant != 3 && ok && lane == 0) {
        const int s = next_load % S;
        ptx::mbar_arrive_expect_tx(bar0 + 8 * s, p.bytes);
        if (rp.cache_hint & 1)
          ptx::bulk_g2s_hint(in0 + s * rp.tile_in, p.src, p.bytes, bar0 + 8 * s, policy);
        else
          ptx::bulk_g2s(in0 + s * rp.tile_in, p.src, p.bytes, bar0 + 8 * s);
      }
      ++next_load;
    }
  };
  auto publish_upto = [&](int layer) {
    drain_stores(lane);
    arrive_layers(ss, open_layer, layer, lane);
    open_layer = layer;
  };

  for (uint32_t q = 0; q < n_my; ++q) {
    while (filled < n_my && filled <= q + ahead) fill();
    if (!pre_valid && filled < n_my && filled <= q + ahead + kPrefetchLead) prefetch();
    pump(q + ahead);
    Piece p;
    const bool ok = ring.get(q, ndst, p, true);
    if (next_load <= q) {  // this item's layer has not been released yet: publish what is finished, then block
      if (ss.want_layers) publish_upto(p.layer);
      if (!wait_layer_ready(ss, p.layer, lane)) {  // gate timeout: abandon the rest (the abort is recorded)
        aborted = true;
        break;
      }
      ready_layer = p.layer;
      pump(q + ahead);
    } else if (ss.want_layers && p.layer > open_layer) {
      publish_upto(p.layer);
    }
    const int s = q % S;
    const uint32_t slot = in0 + s * rp.tile_in;
    if (ok) {
      if (CAST == 0 && rp.variant == 0) {
        if (lane == 0) {
          ptx::mbar_wait(bar0 + 8 * s, (phase >> s) & 1u);
#pragma unroll
          for (int d = 0; d < kMaxDst; ++d)
            if (d < p.ndst) {
              if (rp.cache_hint & 2)
                ptx::bulk_s2g_hint(p.dst[d], slot, p.bytes, policy);
              else
                ptx::bulk_s2g(p.dst[d], slot, p.bytes);
            }
        }
        phase ^= 1u << s;
      } else if (CAST == 0 && rp.variant == 3) {  // stores only: whatever is in the slot
        if (lane == 0) {
#pragma unroll
          for (int d = 0; d < kMaxDst; ++d)
            if (d < p.ndst) ptx::bulk_s2g(p.dst[d], slot, p.bytes);
        }
      } else if (CAST == 0) {  // variants 1, 2: every lane observes the landed tile
        ptx::mbar_wait(bar0 + 8 * s, (phase >> s) & 1u);
        phase ^= 1u << s;
        if (rp.variant == 1) {
          for (int d = 0; d < p.ndst; ++d) warp_store_from_smem(p.dst[d], slot, p.bytes, lane);
        } else if (rp.variant == 4) {
          warp_store_from_smem_mc(p.dst[0], slot, p.bytes, lane);
        }
        __syncwarp();
      } else {
        const uint32_t dst_bytes = CAST == 1 ? p.bytes * 2 : p.bytes / 2;
        const uint32_t ob = out0 + (q & 1) * rp.tile_out;
        if (lane == 0) ptx::bulk_wait_read<1>();  // store q-2 (same out buffer) has been read out
        __syncwarp();
        ptx::mbar_wait(bar0 + 8 * s, (phase >> s) & 1u);  // every lane reads the slot
        phase ^= 1u << s;
        convert_smem<CAST == 1>(slot, ob, p.bytes, lane);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int d = 0; d < kMaxDst; ++d)
            if (d < p.ndst) ptx::bulk_s2g(p.dst[d], ob, dst_bytes);
        }
      }
    } else {
      for (int d = 0; d < p.ndst; ++d) {
        if (CAST == 0 && rp.variant == 4)
          warp_copy_simt_mc(p.dst[d], p.src, p.bytes, lane);
        else if (CAST == 0)
          warp_copy_simt(p.dst[d], p.src, p.bytes, lane);
        else
          warp_cast_simt<CAST == 1>(p.dst[d], p.src, p.bytes, lane);
      }
    }
    if (lane == 0) {
      ptx::bulk_commit();  // always one group per item (possibly empty) so the counts hold
      // byte-exact TMA ring: stores q-P+1..q may still drain; store q-P has left slot (q-P)%S == (q+A)%S
      if (!sync_slot) ptx::bulk_wait_read_n(rp.P);
    }
  }

  drain_stores(lane);
  if (!aborted) arrive_layers(ss, open_layer, ss.layer_end, lane);  // never publish layers that were not copied
  arrive_transfer(ss, lane);
}

// ------------------------------------------------------------------------------------------
// Conversion helpers of the cast rings.
// ------------------------------------------------------------------------------------------
template <bool UP>
__device__ __forceinline__ void convert_smem(uint32_t in_smem, uint32_t out_smem, uint32_t src_bytes,
                                             int lane)
{
  if (UP) {
    // 16 fp8 in (16 B) -> 16 bf16 out (32 B) per lane-iteration
    for (uint32_t off = lane * 16; off < src_bytes; off += 32 * 16) {
      uint4 v;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(in_smem + off));
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t o[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        o[2 * k] = ptx::e4m3x2_to_bf16x2(static_cast<uint16_t>(w[k] & 0xffff));
        o[2 * k + 1] = ptx::e4m3x2_to_bf16x2(static_cast<uint16_t>(w[k] >> 16));
      }
      asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(out_smem + 2 * off), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
      asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(out_smem + 2 * off + 16), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
    }
  } else {
    // 16 bf16 in (32 B) -> 16 fp8 out (16 B)
    for (uint32_t off = lane * 32; off < src_bytes; off += 32 * 32) {
      uint4 a, b;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(in_smem + off));
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(in_smem + off + 16));
      uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o[k] = static_cast<uint32_t>(ptx::bf16x2_to_e4m3x2(w[2 * k])) |
               (static_cast<uint32_t>(ptx::bf16x2_to_e4m3x2(w[2 * k + 1])) << 16);
      asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(out_smem + off / 2), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
    }
  }
}

// SIMT conversion straight from global to global for pieces TMA cannot take (any alignment).
template <bool UP>
__device__ __forceinline__ void warp_cast_simt(uint8_t* dst, const uint8_t* src, uint32_t src_bytes,
                                               int lane)
{
  if (UP) {
    uint16_t* d = reinterpret_cast<uint16_t*>(dst);
    const uint32_t pairs = src_bytes >> 1;
    for (uint32_t i = lane; i < pairs; i += 32) {
      uint16_t two = static_cast<uint16_t>(src[2 * i]) | (static_cast<uint16_t>(src[2 * i + 1]) << 8);
      uint32_t r = ptx::e4m3x2_to_bf16x2(two);
      d[2 * i] = static_cast<uint16_t>(r & 0xffff);
      d[2 * i + 1] = static_cast<uint16_t>(r >> 16);
    }
    if ((src_bytes & 1) && lane == 0)
      d[src_bytes - 1] = static_cast<uint16_t>(ptx::e4m3x2_to_bf16x2(src[src_bytes - 1]) & 0xffff);
  } else {
    const uint16_t* s = reinterpret_cast<const uint16_t*>(src);
    const uint32_t elems = src_bytes >> 1;
    for (uint32_t i = lane; i < elems; i += 32)
      dst[i] = static_cast<uint8_t>(ptx::bf16x2_to_e4m3x2(s[i]) & 0xff);
  }
}

}  // namespace kvbm

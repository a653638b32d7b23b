// libkvbm_physical.so -- host half of the KV transfer path (C++ restatement of Dynamo's kvbm-physical crate
// for this path; see include/kvbm_physical.h for the reference lines every entry point replaces).
#include <cuda_runtime_api.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <memory>
#include <mutex>
#include <random>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <sched.h>
#include <unistd.h>

#include "layout.hpp"
#include "serialized_layout.hpp"

namespace kvbm_host {

static thread_local std::string g_last_error;

// Identity of this address space for layout blobs: the pid alone collides across PID namespaces (two containers can
// both be pid 1), so a per-process random nonce rides in the upper 32 bits of the header's pid field.
static uint64_t process_identity()
{
  static const uint64_t id = [] {
    std::random_device rd;
    uint64_t nonce = (static_cast<uint64_t>(rd()) << 32 | rd()) ^
                     static_cast<uint64_t>(std::chrono::steady_clock::now().time_since_epoch().count());
    nonce ^= nonce >> 29;
    return (static_cast<uint64_t>(getpid()) & 0xffffffffull) | ((nonce & 0xffffffffull) << 32);
  }();
  return id;
}

static int fail(int code, const std::string& msg)
{
  g_last_error = msg;
  return code;
}
int set_last_error(int code, const std::string& msg) { return fail(code, msg); }
static int fail_cuda(cudaError_t e, const char* what)
{
  g_last_error = std::string(what) + " failed: " + cudaGetErrorName(e) + " (" + std::to_string(static_cast<int>(e)) + ")";
  (void)cudaGetLastError();
  return KVBM_ERR_CUDA;
}
#define CU(call)                                   \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// validation.rs
// ---------------------------------------------------------------------------------------------------
// Fast path of the three checks below for the common case (every id in range): one bitmap over the destination's
// block ids instead of a hash set -- 256 ids cost ~0.3 us instead of ~10 us, which is a tenth of a whole 512 MiB
// same-GPU transfer.  Returns false when anything is wrong (or an id is out of range); the reference-ordered slow
// path then produces the exact error.
static bool validate_fast(const size_t* src_ids, const size_t* dst_ids, size_t n, size_t src_blocks, size_t dst_blocks,
                          bool same_layout)
{
  thread_local std::vector<uint64_t> bits;
  const size_t words = (dst_blocks + 63) / 64;
  if (words > (1u << 20)) return false;  // absurdly large pools: take the general path
  if (bits.size() < words) bits.resize(words, 0);
  bool ok = true;
  size_t marked = 0;
  for (; marked < n; ++marked) {
    const size_t d = dst_ids[marked];
    if (d >= dst_blocks || src_ids[marked] >= src_blocks) {
      ok = false;
      break;
    }
    uint64_t& w = bits[d >> 6];
    const uint64_t m = 1ull << (d & 63);
    if (w & m) {
      ok = false;
      break;
    }
    w |= m;
  }
  if (ok && same_layout)
    for (size_t i = 0; i < n; ++i) {
      const size_t v = src_ids[i];
      if (v < dst_blocks && (bits[v >> 6] >> (v & 63)) & 1) {
        ok = false;
        break;
      }
    }
  for (size_t i = 0; i < marked; ++i) bits[dst_ids[i] >> 6] = 0;  // leave the scratch bitmap clean
  return ok;
}

static int validate_block_transfer(const size_t* src_ids, size_t n_src, const size_t* dst_ids, size_t n_dst,
                                   size_t src_blocks, size_t dst_blocks, bool same_layout)
{
  if (n_src != n_dst)  // validation.rs:177-183
    return fail(KVBM_ERR_LENGTH_MISMATCH, "Block ID lists have mismatched lengths: src=" + std::to_string(n_src) +
                                              ", dst=" + std::to_string(n_dst) + ", bounce=None");
  if (n_src == 0 || validate_fast(src_ids, dst_ids, n_src, src_blocks, dst_blocks, same_layout)) return KVBM_OK;
  std::unordered_set<size_t> seen;
  seen.reserve(n_dst * 2);
  for (size_t i = 0; i < n_dst; ++i)  // validate_dst_unique, validation.rs:55-70
    if (!seen.insert(dst_ids[i]).second)
      return fail(KVBM_ERR_DUPLICATE_DST, "Destination block IDs are not unique: duplicates = [" + std::to_string(dst_ids[i]) + "]");
  if (same_layout) {  // validate_disjoint_same_layout, validation.rs:112-136
    for (size_t i = 0; i < n_src; ++i)
      if (seen.count(src_ids[i]))
        return fail(KVBM_ERR_OVERLAP, "Source and destination blocks overlap (same layout): overlapping = [" + std::to_string(src_ids[i]) + "]");
  }
  for (size_t i = 0; i < n_src; ++i) {  // validate_block_ids_in_range, validation.rs:139-158
    if (src_ids[i] >= src_blocks)
      return fail(KVBM_ERR_RANGE, "Block ID " + std::to_string(src_ids[i]) + " out of range for source (max=" + std::to_string(src_blocks) + ")");
    if (dst_ids[i] >= dst_blocks)
      return fail(KVBM_ERR_RANGE, "Block ID " + std::to_string(dst_ids[i]) + " out of range for destination (max=" + std::to_string(dst_blocks) + ")");
  }
  return KVBM_OK;
}

// ---------------------------------------------------------------------------------------------------
// strategy.rs:138-210 (local-to-local rows).  Device index is ignored exactly like strategy.rs:168.
// ---------------------------------------------------------------------------------------------------
// select_remote_strategy (strategy.rs:213-243): the source is local, the destination another agent's
static int select_remote_strategy(int src, const kvbm_transfer_capabilities* caps, kvbm_transfer_plan* out)
{
  if (src == KVBM_STORAGE_SYSTEM || src == KVBM_STORAGE_PINNED) {
    *out = kvbm_transfer_plan{0, KVBM_STRATEGY_NIXL_WRITE, 0, KVBM_STRATEGY_INVALID};
  } else if (src == KVBM_STORAGE_DEVICE) {
    if (caps && caps->allow_gpu_rdma)
      *out = kvbm_transfer_plan{0, KVBM_STRATEGY_NIXL_WRITE, 0, KVBM_STRATEGY_INVALID};
    else
      *out = kvbm_transfer_plan{1, KVBM_STRATEGY_CUDA_ASYNC_D2H, KVBM_STORAGE_PINNED, KVBM_STRATEGY_NIXL_WRITE};
  } else {
    *out = kvbm_transfer_plan{1, KVBM_STRATEGY_NIXL_WRITE, KVBM_STORAGE_PINNED, KVBM_STRATEGY_NIXL_WRITE};  // Disk -> Remote
  }
  return KVBM_OK;
}

static int select_direct_strategy(int src, int dst, const kvbm_transfer_capabilities* caps, kvbm_transfer_plan* out, bool dst_is_remote = false)
{
  if (dst_is_remote) {  // strategy.rs:147-150
    if (src < 0 || src > KVBM_STORAGE_DISK || dst < 0 || dst > KVBM_STORAGE_DISK) return fail(KVBM_ERR, "unknown StorageKind");
    return select_remote_strategy(src, caps, out);
  }
  auto direct = [&](int s) {
    *out = kvbm_transfer_plan{0, s, 0, KVBM_STRATEGY_INVALID};
    return KVBM_OK;
  };
  auto two_hop = [&](int first, int second) {
    *out = kvbm_transfer_plan{1, first, KVBM_STORAGE_PINNED, second};
    return KVBM_OK;
  };
  const bool gds = caps && caps->allow_gds;
  const bool host_s = src == KVBM_STORAGE_SYSTEM || src == KVBM_STORAGE_PINNED;
  const bool host_d = dst == KVBM_STORAGE_SYSTEM || dst == KVBM_STORAGE_PINNED;
  if (src < 0 || src > KVBM_STORAGE_DISK || dst < 0 || dst > KVBM_STORAGE_DISK) return fail(KVBM_ERR, "unknown StorageKind");
  if (host_s && host_d) return direct(KVBM_STRATEGY_MEMCPY);
  if (src == KVBM_STORAGE_SYSTEM && dst == KVBM_STORAGE_DEVICE) return fail(KVBM_ERR_UNSUPPORTED, "System to Device transfers are not supported");
  if (src == KVBM_STORAGE_PINNED && dst == KVBM_STORAGE_DEVICE) return direct(KVBM_STRATEGY_CUDA_ASYNC_H2D);
  if (src == KVBM_STORAGE_DEVICE && dst == KVBM_STORAGE_SYSTEM) return fail(KVBM_ERR_UNSUPPORTED, "Device to System transfers are not supported");
  if (src == KVBM_STORAGE_DEVICE && dst == KVBM_STORAGE_PINNED) return direct(KVBM_STRATEGY_CUDA_ASYNC_D2H);
  if (src == KVBM_STORAGE_DEVICE && dst == KVBM_STORAGE_DEVICE) return direct(KVBM_STRATEGY_CUDA_ASYNC_D2D);
  if (host_s && dst == KVBM_STORAGE_DISK) return direct(KVBM_STRATEGY_NIXL_WRITE);
  if (src == KVBM_STORAGE_DISK && host_d) return direct(KVBM_STRATEGY_NIXL_READ_FLIPPED);
  if (src == KVBM_STORAGE_DISK && dst == KVBM_STORAGE_DISK) return two_hop(KVBM_STRATEGY_NIXL_READ_FLIPPED, KVBM_STRATEGY_NIXL_WRITE);
  if (src == KVBM_STORAGE_DEVICE && dst == KVBM_STORAGE_DISK)
    return gds ? direct(KVBM_STRATEGY_NIXL_WRITE) : two_hop(KVBM_STRATEGY_CUDA_ASYNC_D2H, KVBM_STRATEGY_NIXL_WRITE);
  if (src == KVBM_STORAGE_DISK && dst == KVBM_STORAGE_DEVICE)
    return gds ? direct(KVBM_STRATEGY_NIXL_READ) : two_hop(KVBM_STRATEGY_NIXL_READ_FLIPPED, KVBM_STRATEGY_CUDA_ASYNC_H2D);
  return fail(KVBM_ERR, "unreachable strategy");
}

// select_strategy (strategy.rs:78-108) + select_remote_strategy_v2 (:245-281): locality = whose agent owns the layout
static int select_strategy(int src, bool src_local, int dst, bool dst_local, const kvbm_transfer_capabilities* caps, kvbm_transfer_plan* out)
{
  if (src < 0 || src > KVBM_STORAGE_DISK || dst < 0 || dst > KVBM_STORAGE_DISK) return fail(KVBM_ERR, "unknown StorageKind");
  if (!src_local && !dst_local) return fail(KVBM_ERR_UNSUPPORTED, "Both src and dst are remote - this is not supported.");
  if (src_local && dst_local) return select_direct_strategy(src, dst, caps, out);
  if (src == KVBM_STORAGE_DISK || dst == KVBM_STORAGE_DISK)
    return fail(KVBM_ERR_UNSUPPORTED, "Neither local nor remote disk transfers are supported over NIXL at this time.");
  if (!(caps && caps->allow_gpu_rdma) && (src == KVBM_STORAGE_DEVICE || dst == KVBM_STORAGE_DEVICE))
    return fail(KVBM_ERR_UNSUPPORTED, "GPU RDMA is disabled - this transfer requires GPU RDMA.");
  *out = kvbm_transfer_plan{0, src_local ? KVBM_STRATEGY_NIXL_WRITE : KVBM_STRATEGY_NIXL_READ_FLIPPED, 0, KVBM_STRATEGY_INVALID};
  return KVBM_OK;
}

// ---------------------------------------------------------------------------------------------------
// KvBlockLayout::requires_transform (kv_block_layout.rs:107-119) and select_transform_kernel (executor/mod.rs:46-100)
// ---------------------------------------------------------------------------------------------------
static bool kv_known(int k) { return k >= KVBM_KV_UNIVERSAL_TP && k <= KVBM_KV_CUSTOM; }
static bool kv_requires_transform(int a, int b)
{
  if (a == KVBM_KV_CUSTOM || b == KVBM_KV_CUSTOM) return true;  // the int does not carry Custom's dimension order: never assume equality
  if (kv_known(a) && kv_known(b)) return a != b;  // dim orders of the four named formats are pairwise different
  if (!kv_known(a) && !kv_known(b)) return false; // Unknown -> Unknown: compatible (the reference warns)
  return true;                                    // Unknown <-> known: conservative
}
static int select_transform_kernel(int src, int dst)
{
  if (!kv_requires_transform(src, dst)) return KVBM_TRANSFORM_NONE;
  if (!kv_known(src) || !kv_known(dst)) return KVBM_TRANSFORM_UNSUPPORTED;
  const bool s_op = src == KVBM_KV_OPERATIONAL_NHD || src == KVBM_KV_OPERATIONAL_HND;
  const bool d_op = dst == KVBM_KV_OPERATIONAL_NHD || dst == KVBM_KV_OPERATIONAL_HND;
  const bool s_un = src == KVBM_KV_UNIVERSAL_TP || src == KVBM_KV_UNIVERSAL_PP;
  const bool d_un = dst == KVBM_KV_UNIVERSAL_TP || dst == KVBM_KV_UNIVERSAL_PP;
  if (s_op && d_un) return KVBM_TRANSFORM_BLOCK_TO_UNIVERSAL;
  if (s_un && d_op) return KVBM_TRANSFORM_UNIVERSAL_TO_BLOCK;
  if (s_op && d_op) return KVBM_TRANSFORM_OPERATIONAL_TRANSPOSE;
  return KVBM_TRANSFORM_UNSUPPORTED;  // Custom, and Universal <-> Universal (a TODO in the reference, :87-91)
}
static const char* kv_name(int k)
{
  static const char* n[] = {"Unknown", "UniversalTP", "UniversalPP", "OperationalHND", "OperationalNHD", "Custom"};
  return n[k >= 0 && k <= 5 ? k : 0];
}

// ---------------------------------------------------------------------------------------------------
// TransferManager
// ---------------------------------------------------------------------------------------------------
constexpr int kSlots = 64;     // in-flight transfers before the host has to wait for the oldest
constexpr int kStreams = 4;    // context.rs:286-292: 4 h2d + 4 d2h streams, round-robin

struct Slot {
  int32_t* pinned_ids = nullptr;  // staging for the narrowed block tables (H2D source)
  int32_t* dev_ids = nullptr;
  size_t cap = 0;                 // capacity in int32 entries
  uint32_t* host_flag = nullptr;  // mapped pinned word the kernel's last warp writes
  uint32_t* dev_ws = nullptr;     // sync workspace (zeroed, kernel leaves it zeroed)
  size_t ws_cap = 0;
  cudaEvent_t ev = nullptr;
  std::atomic<uint64_t> seq{0};   // owner (read lock-free by kvbm_notification_is_complete)
  bool in_flight = false;
  bool owns_ids = false;          // false: pinned_ids / dev_ids point into the manager's arenas
  bool owns_ws = false;
  bool owns_flag = false;
};
constexpr size_t kArenaIds = 4096;  // int32 entries per slot carved from the arenas (16 KiB: 2 x 2048-block tables)
constexpr size_t kArenaWs = 512;    // workspace words per slot

#pragma pack(push, 1)
struct BlobHeader {
  char magic[8];
  uint32_t version;
  uint32_t fully_contiguous;
  uint32_t block_dim;
  uint32_t storage;
  int32_t device_id;
  uint32_t n_allocs;
  uint64_t worker_id;
  uint64_t pid;
  uint64_t cfg[9];
};
struct BlobAlloc {
  uint64_t addr;
  uint64_t size;
  uint64_t offset_in_ipc;   // addr - base of the cudaMalloc allocation the IPC handle names
  uint32_t has_ipc;
  unsigned char ipc[64];
};
#pragma pack(pop)
constexpr uint32_t kBlobVersion = 1;
static const char kMagic[8] = {'K', 'V', 'B', 'M', 'L', 'A', 'Y', '1'};

}  // namespace kvbm_host

using namespace kvbm_host;

struct kvbm_transfer_manager {
  int device = -1;       // < 0: host-only
  uint64_t worker_id = 0;
  std::mutex mu;
  std::unordered_map<uint64_t, std::unique_ptr<Layout>> layouts;
  uint16_t next_layout_id = 1;
  cudaStream_t h2d[kStreams] = {};
  cudaStream_t d2h[kStreams] = {};
  std::atomic<uint32_t> rr_h2d{0}, rr_d2h{0};
  Slot slots[kSlots];
  // one pinned and one device allocation back all slots (128 cudaHostAlloc + 128 cudaMalloc per manager took seconds
  // on a multi-GPU box); a slot only allocates privately when a transfer outgrows its share
  uint8_t* pinned_arena = nullptr;
  uint8_t* dev_arena = nullptr;
  uint64_t next_seq = 1;
  std::atomic<uint64_t> bytes_moved{0}, h2d_bytes{0};
  // CUDA IPC mappings opened by import_metadata, keyed by the 64-byte handle: an allocation shared by several
  // imported layouts (e.g. two tensors in one allocator segment) is mapped once; closed when the manager dies.
  std::unordered_map<std::string, void*> ipc_cache;
  // (agent name, worker id) of every SerializedLayout imported so far (manager/mod.rs:572-584 refuses a second load)
  std::set<std::pair<std::string, uint64_t>> loaded_remotes;
  // TransferCapabilities (transfer/strategy.rs:245-278).  allow_gpu_rdma = peers may be written directly over NVLink
  // (the default here); 0 forces the reference's TwoHop plan Device -> Pinned -> Device through a bounce buffer.
  kvbm_transfer_capabilities caps{0, 1};

  Layout* find(kvbm_layout_handle h)
  {
    auto it = layouts.find(h);
    return it == layouts.end() ? nullptr : it->second.get();
  }
};

namespace kvbm_host {

struct DeviceGuard {
  int prev = -1;
  bool active = false;
  explicit DeviceGuard(int dev)
  {
    if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      active = cudaSetDevice(dev) == cudaSuccess;
    }
  }
  ~DeviceGuard()
  {
    if (active) cudaSetDevice(prev);
  }
};

static int upload_layer_base(kvbm_transfer_manager* m, Layout* L)
{
  if (m->device < 0) return KVBM_OK;
  DeviceGuard g(m->device);
  CU(cudaMalloc(reinterpret_cast<void**>(&L->dev_layer_base), L->layer_base.size() * sizeof(uint64_t)));
  CU(cudaMemcpy(L->dev_layer_base, L->layer_base.data(), L->layer_base.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
  return KVBM_OK;
}

// LayoutHandle = (worker_id, layout_id: u16) (manager/handle.rs:16-50).  Ids of unregistered layouts are reused; a live
// layout is never overwritten: 0 = "the 16-bit id space of this worker is exhausted".
static kvbm_layout_handle add_layout(kvbm_transfer_manager* m, Layout&& L)
{
  std::lock_guard<std::mutex> lk(m->mu);
  for (uint32_t tries = 0; tries < 65535; ++tries) {
    uint16_t id = m->next_layout_id++;
    if (id == 0) id = m->next_layout_id++;
    const kvbm_layout_handle h = (m->worker_id << 16) | id;
    if (m->layouts.find(h) != m->layouts.end()) continue;
    m->layouts[h] = std::make_unique<Layout>(std::move(L));
    return h;
  }
  return 0;
}

static kvbm_paged_layout descriptor(const Layout& L)
{
  kvbm_paged_layout d{};
  d.layer_base = L.dev_layer_base;
  d.block_stride = L.block_stride;
  d.outer_stride = L.outer_stride;
  d.region_bytes = static_cast<uint32_t>(L.region);
  d.num_layers = static_cast<uint32_t>(L.cfg.num_layers);
  d.outer_dim = static_cast<uint32_t>(L.cfg.outer_dim);
  d.num_blocks = static_cast<uint32_t>(L.cfg.num_blocks);
  return d;
}

// memcpy.rs:98-165
// executor/memcpy.rs:30-165.  The reference copies chunk after chunk on the calling thread; here the chunk list is
// resolved (and fully validated) first and large requests are split over host threads -- destination blocks are
// unique (validate_block_transfer), so the pieces are disjoint.  Still synchronous: returns when every byte is written.
struct HostCopy {
  void* dst;
  const void* src;
  size_t bytes;
};

static unsigned memcpy_threads()
{
  static const unsigned n = [] {
    if (const char* e = std::getenv("KVBM_MEMCPY_THREADS")) {
      const long v = std::atol(e);
      if (v >= 1) return static_cast<unsigned>(std::min<long>(v, 256));
    }
    const unsigned hw = std::thread::hardware_concurrency();
    return std::max(1u, std::min(hw ? hw : 1u, 32u));   // memory-bound: more threads than memory channels buys nothing
  }();
  return n;
}

static void run_host_copies(const std::vector<HostCopy>& cs)
{
  size_t total = 0;
  for (const HostCopy& c : cs) total += c.bytes;
  const unsigned want = memcpy_threads();
  constexpr size_t kParallelFrom = 8u << 20;   // below this the thread start-up costs more than it saves
  if (want <= 1 || total < kParallelFrom || cs.size() < 2) {
    for (const HostCopy& c : cs) std::memcpy(c.dst, c.src, c.bytes);
    return;
  }
  const unsigned T = static_cast<unsigned>(std::min<size_t>(want, cs.size()));
  auto work = [&](unsigned t) {
    const size_t lo = cs.size() * t / T, hi = cs.size() * (t + 1) / T;
    for (size_t i = lo; i < hi; ++i) std::memcpy(cs[i].dst, cs[i].src, cs[i].bytes);
  };
  std::vector<std::thread> th;
  th.reserve(T - 1);
  for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
}

static int host_memcpy_transfer(const Layout& S, const Layout& D, const size_t* sid, const size_t* did, size_t n,
                                bool has_range, size_t lb, size_t le)
{
  const bool full = !has_range || (lb == 0 && le == S.cfg.num_layers);
  std::string why;
  std::vector<HostCopy> cs;
  if (full && S.fully_contiguous && D.fully_contiguous) {  // can_use_whole_block_transfer, transfer/mod.rs:150-173
    const size_t bytes = bytes_per_block(S.cfg);
    cs.reserve(n);
    for (size_t i = 0; i < n; ++i) {
      uintptr_t s, d;
      int rc;
      if ((rc = S.memory_region(sid[i], 0, 0, &s, nullptr, &why))) return fail(rc, why);
      if ((rc = D.memory_region(did[i], 0, 0, &d, nullptr, &why))) return fail(rc, why);
      cs.push_back({reinterpret_cast<void*>(d), reinterpret_cast<const void*>(s), bytes});
    }
    run_host_copies(cs);
    return KVBM_OK;
  }
  cs.reserve(n * (le - lb) * S.cfg.outer_dim);
  for (size_t i = 0; i < n; ++i)
    for (size_t l = lb; l < le; ++l)
      for (size_t o = 0; o < S.cfg.outer_dim; ++o) {
        uintptr_t s, d;
        size_t ss, ds;
        int rc;
        if ((rc = S.memory_region(sid[i], l, o, &s, &ss, &why))) return fail(rc, why);
        if ((rc = D.memory_region(did[i], l, o, &d, &ds, &why))) return fail(rc, why);
        if (ss != ds)
          return fail(KVBM_ERR_INCOMPATIBLE, "Memory region size mismatch at block=(" + std::to_string(sid[i]) + "," + std::to_string(did[i]) +
                                                 "), layer=" + std::to_string(l) + ", outer=" + std::to_string(o) + ": src=" + std::to_string(ss) +
                                                 ", dst=" + std::to_string(ds));
        cs.push_back({reinterpret_cast<void*>(d), reinterpret_cast<const void*>(s), ss});
      }
  run_host_copies(cs);
  return KVBM_OK;
}

// (Re)size one slot's resources.  cudaMalloc / cudaHostAlloc / cudaMemset synchronise the device, which would
// deadlock against a gated (layer-streaming) transfer that is still spinning on its ready flags -- so every slot
// is provisioned when the manager is created and this only runs again for unusually large requests.
static int ensure_slot(Slot& sl, size_t ids_needed, size_t ws_needed)
{
  if (!sl.ev) CU(cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
  if (!sl.host_flag) {
    CU(cudaHostAlloc(reinterpret_cast<void**>(&sl.host_flag), 64, cudaHostAllocMapped | cudaHostAllocPortable));
    *sl.host_flag = 0;
    sl.owns_flag = true;
  }
  if (sl.cap < ids_needed) {
    if (sl.owns_ids) {
      if (sl.pinned_ids) cudaFreeHost(sl.pinned_ids);
      if (sl.dev_ids) cudaFree(sl.dev_ids);
    }
    sl.pinned_ids = nullptr;
    sl.dev_ids = nullptr;
    sl.cap = 0;
    size_t cap = 4096;
    while (cap < ids_needed) cap *= 2;
    CU(cudaHostAlloc(reinterpret_cast<void**>(&sl.pinned_ids), cap * sizeof(int32_t), cudaHostAllocPortable));
    sl.owns_ids = true;
    CU(cudaMalloc(reinterpret_cast<void**>(&sl.dev_ids), cap * sizeof(int32_t)));
    sl.cap = cap;
  }
  if (sl.ws_cap < ws_needed) {
    if (sl.owns_ws && sl.dev_ws) cudaFree(sl.dev_ws);
    sl.dev_ws = nullptr;
    sl.ws_cap = 0;
    size_t cap = 512;
    while (cap < ws_needed) cap *= 2;
    CU(cudaMalloc(reinterpret_cast<void**>(&sl.dev_ws), cap * sizeof(uint32_t)));
    sl.owns_ws = true;
    CU(cudaMemset(sl.dev_ws, 0, cap * sizeof(uint32_t)));
    sl.ws_cap = cap;
  }
  return KVBM_OK;
}

// Carve every slot's default share out of two arenas (called once, device current).
static int provision_slots(kvbm_transfer_manager* m)
{
  const size_t pin_per = kArenaIds * sizeof(int32_t) + 64;                       // ids + flag word (own cache line)
  const size_t dev_per = kArenaIds * sizeof(int32_t) + kArenaWs * sizeof(uint32_t);
  CU(cudaHostAlloc(reinterpret_cast<void**>(&m->pinned_arena), pin_per * kSlots, cudaHostAllocMapped | cudaHostAllocPortable));
  CU(cudaMalloc(reinterpret_cast<void**>(&m->dev_arena), dev_per * kSlots));
  CU(cudaMemset(m->dev_arena, 0, dev_per * kSlots));
  std::memset(m->pinned_arena, 0, pin_per * kSlots);
  for (int i = 0; i < kSlots; ++i) {
    Slot& sl = m->slots[i];
    uint8_t* hp = m->pinned_arena + pin_per * i;
    uint8_t* dp = m->dev_arena + dev_per * i;
    sl.pinned_ids = reinterpret_cast<int32_t*>(hp);
    sl.host_flag = reinterpret_cast<uint32_t*>(hp + kArenaIds * sizeof(int32_t));
    sl.dev_ids = reinterpret_cast<int32_t*>(dp);
    sl.dev_ws = reinterpret_cast<uint32_t*>(dp + kArenaIds * sizeof(int32_t));
    sl.cap = kArenaIds;
    sl.ws_cap = kArenaWs;
    CU(cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
  }
  return KVBM_OK;
}

// Acquire a slot whose previous transfer has completed (waits for the oldest if all are busy).
static int acquire_slot(kvbm_transfer_manager* m, size_t ids_needed, size_t ws_needed, Slot** out, uint64_t* seq)
{
  const uint64_t s = m->next_seq++;
  Slot& sl = m->slots[(s - 1) % kSlots];
  if (sl.in_flight) {
    CU(cudaEventSynchronize(sl.ev));
    sl.in_flight = false;
  }
  int rc = ensure_slot(sl, ids_needed, ws_needed);
  if (rc) return rc;
  // the previous owner has completed: clear its completion word (an abort marker must not be mistaken for ours) BEFORE
  // the new owner is published
  __atomic_store_n(sl.host_flag, 0u, __ATOMIC_RELAXED);
  sl.seq.store(s, std::memory_order_release);
  *out = &sl;
  *seq = s;
  return KVBM_OK;
}

static int check_compat(const Layout& S, const Layout& D, int cast)
{
  if (S.cfg.num_layers != D.cfg.num_layers)  // executor/cuda.rs:52-58, memcpy.rs:49-55
    return fail(KVBM_ERR_INCOMPATIBLE, "Layouts have incompatible layer counts: src=" + std::to_string(S.cfg.num_layers) +
                                           ", dst=" + std::to_string(D.cfg.num_layers));
  if (S.cfg.outer_dim != D.cfg.outer_dim)  // cuda.rs:60-66
    return fail(KVBM_ERR_INCOMPATIBLE, "Layouts have incompatible outer dimensions: src=" + std::to_string(S.cfg.outer_dim) +
                                           ", dst=" + std::to_string(D.cfg.outer_dim));
  const size_t num = cast == KVBM_CAST_FP8E4M3_TO_BF16 ? 2 : 1, den = cast == KVBM_CAST_BF16_TO_FP8E4M3 ? 2 : 1;
  if (S.region * num != D.region * den)
    return fail(KVBM_ERR_INCOMPATIBLE, "Memory region size mismatch: src=" + std::to_string(S.region) + ", dst=" + std::to_string(D.region));
  if (cast == KVBM_CAST_FP8E4M3_TO_BF16 && (S.cfg.dtype_width_bytes != 1 || D.cfg.dtype_width_bytes != 2))
    return fail(KVBM_ERR_INCOMPATIBLE, "fp8->bf16 cast needs dtype widths 1 -> 2");
  if (cast == KVBM_CAST_BF16_TO_FP8E4M3 && (S.cfg.dtype_width_bytes != 2 || D.cfg.dtype_width_bytes != 1))
    return fail(KVBM_ERR_INCOMPATIBLE, "bf16->fp8 cast needs dtype widths 2 -> 1");
  return KVBM_OK;
}

// One block-table kernel launch (execute_cuda_transfer, executor/cuda.rs:39-156) on `stream`, optionally ordered after
// `wait_ev`.  The manager mutex is held by the caller.
static int launch_cuda(kvbm_transfer_manager* m, Layout* S, Layout* const* D, int nd, const size_t* const* src_ids,
                       const size_t* const* dst_ids, size_t n, bool replicate, size_t lb, size_t le,
                       const kvbm_transfer_options& o, cudaStream_t stream, cudaEvent_t wait_ev, Slot** slot_out, uint64_t* seq_out)
{
  // block tables: narrow to int32 into pinned staging, async upload on the transfer's stream (no host sync;
  // the reference blocks on pointers_transfered_event.synchronize(), cuda.rs:324)
  const size_t lists = replicate ? static_cast<size_t>(nd) + 1 : 2 * static_cast<size_t>(nd);
  Slot* sl;
  uint64_t seq;
  int rc = acquire_slot(m, lists * n, S->cfg.num_layers + 4, &sl, &seq);
  if (rc) return rc;
  if (wait_ev) CU(cudaStreamWaitEvent(stream, wait_ev, 0));
  auto narrow = [&](const size_t* ids, int32_t* dstp) { for (size_t i = 0; i < n; ++i) dstp[i] = static_cast<int32_t>(ids[i]); };
  std::vector<const int32_t*> dev_src(nd), dev_dst(nd);
  size_t k = 0;
  if (replicate) {
    narrow(src_ids[0], sl->pinned_ids);
    for (int d = 0; d < nd; ++d) dev_src[d] = sl->dev_ids;
    k = 1;
  }
  for (int d = 0; d < nd; ++d) {
    if (!replicate) {
      narrow(src_ids[d], sl->pinned_ids + k * n);
      dev_src[d] = sl->dev_ids + k * n;
      ++k;
    }
    narrow(dst_ids[d], sl->pinned_ids + k * n);
    dev_dst[d] = sl->dev_ids + k * n;
    ++k;
  }
  CU(cudaMemcpyAsync(sl->dev_ids, sl->pinned_ids, k * n * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  m->h2d_bytes += k * n * sizeof(int32_t);

  kvbm_paged_layout sdesc = descriptor(*S);
  kvbm_paged_dst dd[KVBM_MAX_DESTINATIONS] = {};
  for (int d = 0; d < nd; ++d) {
    dd[d].layout = descriptor(*D[d]);
    dd[d].src_block_ids = dev_src[d];
    dd[d].dst_block_ids = dev_dst[d];
    dd[d].done_flag = o.per_dst_done_flags ? o.per_dst_done_flags[d] : ((d == 0) ? o.done_flag : nullptr);
    dd[d].layer_done_flags = o.per_dst_layer_done_flags ? o.per_dst_layer_done_flags[d] : ((d == 0) ? o.layer_done_flags : nullptr);
  }
  kvbm_paged_copy_opts ko{};
  ko.epoch = o.epoch;
  ko.layer_ready_flags = o.layer_ready_flags;
  ko.sync_workspace = sl->dev_ws;
  ko.max_ctas = o.max_ctas;
  ko.gate_timeout_ms = o.gate_timeout_ms;
  ko.gate_mode = o.gate_mode;
  ko.multicast = o.multicast;
  ko.completion_flag = sl->host_flag;
  ko.completion_value = static_cast<uint32_t>(seq);
  cudaError_t e = kvbm_kernels_paged_copy_v2(&sdesc, dd, nd, static_cast<int>(n), static_cast<int>(lb), static_cast<int>(le), o.cast_mode, &ko, stream);
  if (e != cudaSuccess) return fail_cuda(e, "kvbm_kernels_paged_copy_v2");
  CU(cudaEventRecord(sl->ev, stream));
  sl->in_flight = true;
  for (int d = 0; d < nd; ++d) m->bytes_moved += n * (le - lb) * S->cfg.outer_dim * D[d]->region;
  *slot_out = sl;
  *seq_out = seq;
  return KVBM_OK;
}

// The layout-transforming transfer: what select_transform_kernel's result would drive if the reference ran it inside
// execute_transfer (it rejects the pair instead, transfer/mod.rs:128-147).  One kvbm_kernels_paged_permute launch:
// block tables uploaded like launch_cuda's, no pointer tables, completion word written by a trailing signal kernel.
static int launch_transform(kvbm_transfer_manager* m, Layout* S, Layout* D, int src_kv, int dst_kv, const size_t* src_ids,
                            const size_t* dst_ids, size_t n, size_t lb, size_t le, const kvbm_transfer_options& o,
                            cudaStream_t stream, Slot** slot_out, uint64_t* seq_out)
{
  Slot* sl;
  uint64_t seq;
  int rc = acquire_slot(m, 2 * n, 4, &sl, &seq);
  if (rc) return rc;
  for (size_t i = 0; i < n; ++i) {
    sl->pinned_ids[i] = static_cast<int32_t>(src_ids[i]);
    sl->pinned_ids[n + i] = static_cast<int32_t>(dst_ids[i]);
  }
  CU(cudaMemcpyAsync(sl->dev_ids, sl->pinned_ids, 2 * n * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  m->h2d_bytes += 2 * n * sizeof(int32_t);
  kvbm_permute_side ps{descriptor(*S), sl->dev_ids, src_kv}, pd{descriptor(*D), sl->dev_ids + n, dst_kv};
  const size_t nh = S->cfg.num_heads;
  const size_t row = S->cfg.inner_dim / nh * S->cfg.dtype_width_bytes;
  cudaError_t e = kvbm_kernels_paged_permute(&ps, &pd, static_cast<int>(n), static_cast<int>(lb), static_cast<int>(le),
                                             static_cast<uint32_t>(nh), static_cast<uint32_t>(S->cfg.page_size),
                                             static_cast<uint32_t>(row), o.done_flag, o.epoch, sl->host_flag,
                                             static_cast<uint32_t>(seq), stream);
  if (e != cudaSuccess) return fail_cuda(e, "kvbm_kernels_paged_permute");
  CU(cudaEventRecord(sl->ev, stream));
  sl->in_flight = true;
  m->bytes_moved += n * (le - lb) * S->cfg.outer_dim * D->region;
  *slot_out = sl;
  *seq_out = seq;
  return KVBM_OK;
}

static bool cuda_strategy(int s) { return s == KVBM_STRATEGY_CUDA_ASYNC_H2D || s == KVBM_STRATEGY_CUDA_ASYNC_D2H || s == KVBM_STRATEGY_CUDA_ASYNC_D2D; }

// execute_two_hop_transfer (executor/mod.rs:477-571) with handle_buffered_transfer's two bounce groups (:357-416): the
// bounce blocks are split in two halves; chunk i goes src -> half (i & 1) -> dst.  The reference runs two tokio tasks
// that each await their hops; here both hops are stream work ordered by events, so chunk i+1's first hop overlaps chunk
// i's second hop and the host never blocks.  One bounce block = strictly sequential, like the reference (:526-548).
static int execute_two_hop(kvbm_transfer_manager* m, Layout* S, Layout* D, Layout* B, const size_t* src_ids, const size_t* dst_ids,
                           size_t n, const size_t* bounce_ids, size_t nb, size_t lb, size_t le, const kvbm_transfer_plan& plan,
                           const kvbm_transfer_options& o, kvbm_notification* out)
{
  auto pick = [&](int strategy) { return strategy == KVBM_STRATEGY_CUDA_ASYNC_D2H ? m->d2h[m->rr_d2h++ % kStreams] : m->h2d[m->rr_h2d++ % kStreams]; };
  cudaStream_t s1 = pick(plan.first), s2 = pick(plan.second);
  nb = std::min(nb, n);
  const size_t half = nb >= 2 ? nb / 2 : nb;
  const size_t group_begin[2] = {0, half};
  const size_t group_len[2] = {half, nb >= 2 ? nb - half : 0};
  const int groups = nb >= 2 ? 2 : 1;
  cudaEvent_t hop2_done[2] = {nullptr, nullptr};
  kvbm_transfer_options o1 = o, o2 = o;
  o1.done_flag = nullptr;        // completion signals belong to the hop that lands the bytes at the destination
  o1.layer_done_flags = nullptr;
  o1.per_dst_done_flags = nullptr;
  o1.per_dst_layer_done_flags = nullptr;
  o2.layer_ready_flags = nullptr;  // ... and gating to the hop that reads the source
  uint64_t last = 0;
  size_t pos = 0;
  for (int c = 0; pos < n; ++c) {
    const int g = c % groups;
    const size_t cnt = std::min(group_len[g], n - pos);
    const size_t* bids = bounce_ids + group_begin[g];
    const size_t* sp = src_ids + pos;
    const size_t* dp = dst_ids + pos;
    Slot *a, *b;
    uint64_t sa, sb;
    Layout* Bp = B;
    int rc = launch_cuda(m, S, &Bp, 1, &sp, &bids, cnt, true, lb, le, o1, s1, hop2_done[g], &a, &sa);
    if (rc) return rc;
    Layout* Dp = D;
    kvbm_transfer_options o2c = o2;
    if (pos + cnt < n) o2c.done_flag = nullptr;  // only the last chunk announces the whole transfer
    rc = launch_cuda(m, B, &Dp, 1, &bids, &dp, cnt, true, lb, le, o2c, s2, a->ev, &b, &sb);
    if (rc) return rc;
    hop2_done[g] = b->ev;
    last = sb;
    pos += cnt;
  }
  if (out) *out = last;
  return KVBM_OK;
}

static int execute(kvbm_transfer_manager* m, kvbm_layout_handle src_h, int nd, const kvbm_layout_handle* dst_h,
                   const size_t* const* src_ids, const size_t* const* dst_ids, size_t n, bool replicate,
                   const kvbm_transfer_options* opts_in, kvbm_notification* out)
{
  kvbm_transfer_options o{};
  if (opts_in) o = *opts_in;
  if (out) *out = 0;
  if (nd < 1 || nd > KVBM_MAX_DESTINATIONS) return fail(KVBM_ERR, "number of destinations must be in 1..=8");
  std::lock_guard<std::mutex> lk(m->mu);
  Layout* S = m->find(src_h);
  if (!S) return fail(KVBM_ERR_HANDLE, "invalid source handle");
  if (S->unmapped) return fail(KVBM_ERR_UNSUPPORTED, "source layout is a descriptor of another process's host memory: not addressable from this process");
  std::vector<Layout*> D(nd);
  for (int d = 0; d < nd; ++d) {
    D[d] = m->find(dst_h[d]);
    if (!D[d]) return fail(KVBM_ERR_HANDLE, "invalid destination handle");
    if (D[d]->unmapped) return fail(KVBM_ERR_UNSUPPORTED, "destination layout is a descriptor of another process's host memory: not addressable from this process");
    int rc = check_compat(*S, *D[d], o.cast_mode);
    if (rc) return rc;
    if (n && (!src_ids || !dst_ids || !src_ids[d] || !dst_ids[d])) return fail(KVBM_ERR, "null block id list");
    rc = validate_block_transfer(src_ids[d], n, dst_ids[d], n, S->cfg.num_blocks, D[d]->cfg.num_blocks, src_h == dst_h[d]);
    if (rc) return rc;
  }
  for (int d = 0; d < nd; ++d)  // select_strategy, strategy.rs:86-90
    if (S->remote && D[d]->remote) return fail(KVBM_ERR_UNSUPPORTED, "Both src and dst are remote - this is not supported.");
  // effective_src_layout / effective_dst_layout (executor/mod.rs:103-119): the option overrides the layout's own format.
  // The reference rejects every pair that needs a transformation (validate_layout_compatibility, transfer/mod.rs:128-147);
  // here the pairs select_transform_kernel names -- plus UniversalTP <-> UniversalPP, its TODO -- run as ONE permuting launch.
  const int src_kv = o.src_kv_layout ? o.src_kv_layout : S->kv_block_layout;
  int transform = KVBM_TRANSFORM_NONE;
  for (int d = 0; d < nd; ++d) {
    const int dst_kv = o.dst_kv_layout ? o.dst_kv_layout : D[d]->kv_block_layout;
    int t = select_transform_kernel(src_kv, dst_kv);
    if (t == KVBM_TRANSFORM_UNSUPPORTED && (src_kv == KVBM_KV_UNIVERSAL_TP || src_kv == KVBM_KV_UNIVERSAL_PP) &&
        (dst_kv == KVBM_KV_UNIVERSAL_TP || dst_kv == KVBM_KV_UNIVERSAL_PP))
      t = KVBM_TRANSFORM_UNIVERSAL_TO_UNIVERSAL;
    if (t == KVBM_TRANSFORM_UNSUPPORTED)
      return fail(KVBM_ERR_UNSUPPORTED, std::string("Layout transformation not supported: src=") + kv_name(src_kv) + ", dst=" + kv_name(dst_kv));
    if (t != KVBM_TRANSFORM_NONE) {
      if (nd != 1) return fail(KVBM_ERR_UNSUPPORTED, "a layout-transforming transfer takes one destination");
      if (o.cast_mode != KVBM_CAST_NONE) return fail(KVBM_ERR_UNSUPPORTED, "the fused cast and a layout transformation cannot be combined");
      if (o.layer_ready_flags || o.layer_done_flags || o.per_dst_done_flags || o.per_dst_layer_done_flags || o.multicast)
        return fail(KVBM_ERR_UNSUPPORTED, "layer streaming / multicast are not available on a layout-transforming transfer");
      const kvbm_layout_config &sc = S->cfg, &dc = D[d]->cfg;
      if (!sc.num_heads || !dc.num_heads) return fail(KVBM_ERR_CONFIG, "num_heads_required_for_kv_block_layout");  // config.rs:118-126
      if (sc.num_heads != dc.num_heads || sc.page_size != dc.page_size || sc.inner_dim != dc.inner_dim || sc.dtype_width_bytes != dc.dtype_width_bytes)
        return fail(KVBM_ERR_INCOMPATIBLE, "a layout transformation needs equal num_heads, page_size, inner_dim and dtype on both sides");
      if (sc.inner_dim % sc.num_heads) return fail(KVBM_ERR_CONFIG, "inner_dim_must_be_divisible_by_num_heads");
      const size_t row = sc.inner_dim / sc.num_heads * sc.dtype_width_bytes;
      if (row < 16 || row > 65536 || (row & 15))
        return fail(KVBM_ERR_UNSUPPORTED, "layout transformation: head_dim * dtype width must be a multiple of 16 bytes in 16..65536, got " + std::to_string(row));
      const bool s_uni = src_kv == KVBM_KV_UNIVERSAL_TP || src_kv == KVBM_KV_UNIVERSAL_PP;
      const bool d_uni = dst_kv == KVBM_KV_UNIVERSAL_TP || dst_kv == KVBM_KV_UNIVERSAL_PP;
      if ((s_uni && !S->fully_contiguous) || (d_uni && !D[d]->fully_contiguous))
        return fail(KVBM_ERR_INCOMPATIBLE, "a universal KV block layout needs a fully contiguous pool");
      auto aligned16 = [](const Layout& L) {
        if ((L.block_stride | L.outer_stride) & 15) return false;
        for (uint64_t b : L.layer_base)
          if (b & 15) return false;
        return true;
      };
      if (!aligned16(*S) || !aligned16(*D[d])) return fail(KVBM_ERR_UNSUPPORTED, "layout transformation needs 16-byte aligned pools and strides");
    }
    transform = t;
  }
  const int dst_kv0 = o.dst_kv_layout ? o.dst_kv_layout : D[0]->kv_block_layout;
  size_t lb = 0, le = S->cfg.num_layers;
  if (o.has_layer_range) {
    lb = o.layer_begin;
    le = o.layer_end;
    if (lb > le || le > S->cfg.num_layers)
      return fail(KVBM_ERR_RANGE, "Layer range " + std::to_string(lb) + ".." + std::to_string(le) + " exceeds num_layers " + std::to_string(S->cfg.num_layers));
  }
  if (n == 0 || lb == le) return KVBM_OK;

  // select_strategy (strategy.rs:78-108): every layout handled here is local or peer-mapped
  kvbm_transfer_plan plan{};
  for (int d = 0; d < nd; ++d) {
    kvbm_transfer_plan p{};
    int rc = select_direct_strategy(S->storage, D[d]->storage, &m->caps, &p);
    if (rc) return rc;
    // strategy.rs:222-233,245-278: a device pool of ANOTHER GPU is "remote"; without GPU RDMA it is reached through
    // pinned host memory: TwoHop { CudaAsyncD2H, Pinned, CudaAsyncH2D }
    if (!p.two_hop && p.first == KVBM_STRATEGY_CUDA_ASYNC_D2D && !m->caps.allow_gpu_rdma &&
        (S->device_id != D[d]->device_id || S->remote != D[d]->remote))
      p = kvbm_transfer_plan{1, KVBM_STRATEGY_CUDA_ASYNC_D2H, KVBM_STORAGE_PINNED, KVBM_STRATEGY_CUDA_ASYNC_H2D};
    if (d && (p.first != plan.first || p.two_hop != plan.two_hop)) return fail(KVBM_ERR_UNSUPPORTED, "destinations need different strategies");
    plan = p;
  }
  if (plan.two_hop) {
    if (!cuda_strategy(plan.first) || !cuda_strategy(plan.second))
      return fail(KVBM_ERR_UNSUPPORTED, "two-hop plans with a NIXL (disk / remote agent) leg are outside this library");
    if (m->device < 0) return fail(KVBM_ERR_CUDA, "this TransferManager was created without a CUDA device; CUDA strategies have no CPU fallback");
    if (nd != 1) return fail(KVBM_ERR_UNSUPPORTED, "two-hop transfers take one destination");
    if (o.use_caller_stream) return fail(KVBM_ERR_UNSUPPORTED, "Two-hop transfers don't support caller-provided streams");  // executor/mod.rs:441
    if (o.cast_mode != KVBM_CAST_NONE) return fail(KVBM_ERR_UNSUPPORTED, "the fused cast is not available on two-hop transfers");
    if (transform != KVBM_TRANSFORM_NONE) return fail(KVBM_ERR_UNSUPPORTED, "a layout transformation is not available on two-hop transfers");
    if (!o.bounce_layout || !o.bounce_block_ids || o.num_bounce_blocks == 0)
      return fail(KVBM_ERR, "Two-hop transfers require a bounce buffer.");  // executor/mod.rs:514-519
    Layout* B = m->find(o.bounce_layout);
    if (!B) return fail(KVBM_ERR_HANDLE, "invalid bounce buffer handle");
    if (B->storage != plan.bounce_location) return fail(KVBM_ERR, "Bounce buffer layout does not match bounce location.");  // :521-527
    int rc = check_compat(*S, *B, KVBM_CAST_NONE);
    if (rc) return rc;
    if ((rc = check_compat(*B, *D[0], KVBM_CAST_NONE))) return rc;
    std::unordered_set<size_t> seen;
    for (size_t i = 0; i < o.num_bounce_blocks; ++i) {
      if (o.bounce_block_ids[i] >= B->cfg.num_blocks) return fail(KVBM_ERR_RANGE, "bounce block id out of range");
      if (!seen.insert(o.bounce_block_ids[i]).second) return fail(KVBM_ERR_DUPLICATE_DST, "duplicate bounce block id");
    }
    DeviceGuard g(m->device);
    return execute_two_hop(m, S, D[0], B, src_ids[0], dst_ids[0], n, o.bounce_block_ids, o.num_bounce_blocks, lb, le, plan, o, out);
  }

  if (plan.first == KVBM_STRATEGY_MEMCPY) {
    if (o.cast_mode != KVBM_CAST_NONE) return fail(KVBM_ERR_UNSUPPORTED, "the fused cast exists only on the CUDA strategies");
    if (transform != KVBM_TRANSFORM_NONE) return fail(KVBM_ERR_UNSUPPORTED, "layout transformations exist only on the CUDA strategies");
    if (o.use_caller_stream) return fail(KVBM_ERR_UNSUPPORTED, "cuda_stream option is not supported for Memcpy strategy");  // executor/mod.rs:272-276
    for (int d = 0; d < nd; ++d) {
      int rc = host_memcpy_transfer(*S, *D[d], src_ids[d], dst_ids[d], n, o.has_layer_range != 0, lb, le);
      if (rc) return rc;
      m->bytes_moved += n * (le - lb) * S->cfg.outer_dim * D[d]->region;
    }
    return KVBM_OK;  // synchronous: TransferCompleteNotification::completed() (memcpy.rs:91-92)
  }
  if (!cuda_strategy(plan.first))
    return fail(KVBM_ERR_UNSUPPORTED, "NIXL strategies are outside this library (the NVLink peer path replaces them)");
  if (m->device < 0) return fail(KVBM_ERR_CUDA, "this TransferManager was created without a CUDA device; CUDA strategies have no CPU fallback");

  DeviceGuard g(m->device);
  // stream: caller's, or round-robin from the pool (cuda.rs:82-90: D2H pool for D2H, H2D pool otherwise)
  cudaStream_t stream;
  if (o.use_caller_stream)
    stream = o.cuda_stream;
  else if (plan.first == KVBM_STRATEGY_CUDA_ASYNC_D2H)
    stream = m->d2h[m->rr_d2h++ % kStreams];
  else
    stream = m->h2d[m->rr_h2d++ % kStreams];
  Slot* sl;
  uint64_t seq;
  int rc = transform != KVBM_TRANSFORM_NONE ? launch_transform(m, S, D[0], src_kv, dst_kv0, src_ids[0], dst_ids[0], n, lb, le, o, stream, &sl, &seq)
                                            : launch_cuda(m, S, D.data(), nd, src_ids, dst_ids, n, replicate, lb, le, o, stream, nullptr, &sl, &seq);
  if (rc) return rc;
  // caller-provided stream: caller manages sync, completed() is returned (cuda.rs:139-141)
  if (out) *out = o.use_caller_stream ? 0 : seq;
  return KVBM_OK;
}

}  // namespace kvbm_host

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" const char* kvbm_last_error(void) { return g_last_error.c_str(); }

extern "C" int kvbm_layout_config_validate(const kvbm_layout_config* cfg)
{
  try {
  if (!cfg) return fail(KVBM_ERR, "null config");
  std::string why;
  int rc = validate_config(*cfg, &why);
  return rc ? fail(rc, why) : KVBM_OK;
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}
extern "C" size_t kvbm_layout_required_bytes(const kvbm_layout_config* cfg) { return cfg ? required_bytes(*cfg) : 0; }
extern "C" size_t kvbm_layout_bytes_per_block(const kvbm_layout_config* cfg) { return cfg ? bytes_per_block(*cfg) : 0; }

extern "C" int kvbm_select_direct_strategy(int src_kind, int dst_kind, const kvbm_transfer_capabilities* caps, kvbm_transfer_plan* out)
{
  try {
  if (!out) return fail(KVBM_ERR, "null plan");
  return select_direct_strategy(src_kind, dst_kind, caps, out);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_select_direct_strategy_remote(int src_kind, int dst_kind, int dst_is_remote, const kvbm_transfer_capabilities* caps,
                                                  kvbm_transfer_plan* out)
{
  try {
  if (!out) return fail(KVBM_ERR, "null plan");
  return select_direct_strategy(src_kind, dst_kind, caps, out, dst_is_remote != 0);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_select_strategy(int src_kind, int src_is_local, int dst_kind, int dst_is_local, const kvbm_transfer_capabilities* caps,
                                    kvbm_transfer_plan* out)
{
  try {
  if (!out) return fail(KVBM_ERR, "null plan");
  return select_strategy(src_kind, src_is_local != 0, dst_kind, dst_is_local != 0, caps, out);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_manager_select_strategy(kvbm_transfer_manager* m, kvbm_layout_handle src, kvbm_layout_handle dst, kvbm_transfer_plan* out)
{
  try {
  if (!m || !out) return fail(KVBM_ERR, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  Layout *S = m->find(src), *D = m->find(dst);
  if (!S || !D) return fail(KVBM_ERR_HANDLE, "invalid layout handle");
  return select_strategy(S->storage, !S->remote, D->storage, !D->remote, &m->caps, out);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_validate_block_transfer(const size_t* src_ids, size_t n_src, const size_t* dst_ids, size_t n_dst,
                                            size_t src_num_blocks, size_t dst_num_blocks, int same_layout)
{
  try {
  return validate_block_transfer(src_ids, n_src, dst_ids, n_dst, src_num_blocks, dst_num_blocks, same_layout != 0);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_manager_create(int cuda_device_id, uint64_t worker_id, kvbm_transfer_manager** out)
{
  try {
  if (!out) return fail(KVBM_ERR, "null out");
  auto m = std::make_unique<kvbm_transfer_manager>();
  m->device = cuda_device_id;
  m->worker_id = worker_id & 0xffffffffffffull;
  if (cuda_device_id >= 0) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
    if (cuda_device_id >= count) return fail(KVBM_ERR_CUDA, "CUDA device " + std::to_string(cuda_device_id) + " does not exist");
    DeviceGuard g(cuda_device_id);
    for (int i = 0; i < kStreams; ++i) {
      CU(cudaStreamCreateWithFlags(&m->h2d[i], cudaStreamNonBlocking));
      CU(cudaStreamCreateWithFlags(&m->d2h[i], cudaStreamNonBlocking));
    }
    int rc = provision_slots(m.get());
    if (rc) return rc;
  }
  *out = m.release();
  return KVBM_OK;
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" void kvbm_manager_destroy(kvbm_transfer_manager* m)
{
  if (!m) return;
  if (m->device >= 0) {
    DeviceGuard g(m->device);
    cudaDeviceSynchronize();
    for (auto& kv : m->layouts)
      if (kv.second->dev_layer_base) cudaFree(kv.second->dev_layer_base);
    for (auto& kv : m->ipc_cache) cudaIpcCloseMemHandle(kv.second);
    for (auto& s : m->slots) {
      if (s.ev) cudaEventDestroy(s.ev);
      if (s.owns_ids && s.pinned_ids) cudaFreeHost(s.pinned_ids);
      if (s.owns_ids && s.dev_ids) cudaFree(s.dev_ids);
      if (s.owns_flag && s.host_flag) cudaFreeHost(s.host_flag);
      if (s.owns_ws && s.dev_ws) cudaFree(s.dev_ws);
    }
    if (m->pinned_arena) cudaFreeHost(m->pinned_arena);
    if (m->dev_arena) cudaFree(m->dev_arena);
    for (int i = 0; i < kStreams; ++i) {
      if (m->h2d[i]) cudaStreamDestroy(m->h2d[i]);
      if (m->d2h[i]) cudaStreamDestroy(m->d2h[i]);
    }
  }
  delete m;
}

static int finish_register(kvbm_transfer_manager* m, Layout&& L, int storage_kind, int device_id, kvbm_layout_handle* out)
{
  if (storage_kind < KVBM_STORAGE_SYSTEM || storage_kind > KVBM_STORAGE_DISK) return fail(KVBM_ERR, "unknown StorageKind");
  if (L.region >= (1ull << 32)) return fail(KVBM_ERR_CONFIG, "a (block, layer, outer) region must be < 4 GiB");
  L.storage = storage_kind;
  L.device_id = device_id;
  int rc = upload_layer_base(m, &L);
  if (rc) return rc;
  uint64_t* dev_table = L.dev_layer_base;
  *out = add_layout(m, std::move(L));
  if (*out == 0) {
    if (dev_table) cudaFree(dev_table);
    return fail(KVBM_ERR, "layout id space exhausted: 65535 layouts are registered on this worker");
  }
  return KVBM_OK;
}

extern "C" int kvbm_manager_register_fully_contiguous(kvbm_transfer_manager* m, const kvbm_layout_config* cfg, void* base, size_t size,
                                                      int storage_kind, int device_id, kvbm_layout_handle* out)
{
  try {
  if (!m || !cfg || !out) return fail(KVBM_ERR, "null argument");
  Layout L;
  std::string why;
  int rc = make_fully_contiguous(*cfg, reinterpret_cast<uintptr_t>(base), size, &L, &why);
  if (rc) return fail(rc, why);
  return finish_register(m, std::move(L), storage_kind, device_id, out);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_manager_register_layer_separate(kvbm_transfer_manager* m, const kvbm_layout_config* cfg, void* const* layer_bases,
                                                    const size_t* layer_sizes, int block_dim, int storage_kind, int device_id,
                                                    kvbm_layout_handle* out)
{
  try {
  if (!m || !cfg || !out || !layer_bases || !layer_sizes) return fail(KVBM_ERR, "null argument");
  std::vector<uintptr_t> bases(cfg->num_layers);
  for (size_t i = 0; i < cfg->num_layers; ++i) bases[i] = reinterpret_cast<uintptr_t>(layer_bases[i]);
  Layout L;
  std::string why;
  int rc = make_layer_separate(*cfg, bases.data(), layer_sizes, cfg->num_layers, block_dim, &L, &why);
  if (rc) return fail(rc, why);
  return finish_register(m, std::move(L), storage_kind, device_id, out);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_manager_unregister(kvbm_transfer_manager* m, kvbm_layout_handle h)
{
  try {
  if (!m) return fail(KVBM_ERR, "null manager");
  std::lock_guard<std::mutex> lk(m->mu);
  auto it = m->layouts.find(h);
  if (it == m->layouts.end()) return fail(KVBM_ERR_HANDLE, "invalid handle");
  if (m->device >= 0) {
    DeviceGuard g(m->device);
    for (auto& s : m->slots)
      if (s.in_flight) {
        cudaEventSynchronize(s.ev);
        s.in_flight = false;
      }
    if (it->second->dev_layer_base) cudaFree(it->second->dev_layer_base);
  }
  m->layouts.erase(it);
  return KVBM_OK;
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_layout_memory_region(kvbm_transfer_manager* m, kvbm_layout_handle h, size_t block, size_t layer, size_t outer,
                                         uintptr_t* addr, size_t* size)
{
  try {
  if (!m || !addr) return fail(KVBM_ERR, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  Layout* L = m->find(h);
  if (!L) return fail(KVBM_ERR_HANDLE, "invalid handle");
  std::string why;
  int rc = L->memory_region(block, layer, outer, addr, size, &why);
  return rc ? fail(rc, why) : KVBM_OK;
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_layout_is_fully_contiguous(kvbm_transfer_manager* m, kvbm_layout_handle h)
{
  if (!m) return -1;
  std::lock_guard<std::mutex> lk(m->mu);
  Layout* L = m->find(h);
  return L ? (L->fully_contiguous ? 1 : 0) : -1;
}

extern "C" int kvbm_manager_enable_peer_access(kvbm_transfer_manager* m, int peer_device)
{
  try {
  if (!m || m->device < 0) return fail(KVBM_ERR_CUDA, "manager has no CUDA device");
  if (peer_device == m->device) return KVBM_OK;
  DeviceGuard g(m->device);
  int can = 0;
  CU(cudaDeviceCanAccessPeer(&can, m->device, peer_device));
  if (!can) return fail(KVBM_ERR_UNSUPPORTED, "device " + std::to_string(m->device) + " cannot map device " + std::to_string(peer_device) + " (no NVLink/P2P path)");
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();
    return KVBM_OK;
  }
  if (e != cudaSuccess) return fail_cuda(e, "cudaDeviceEnablePeerAccess");
  return KVBM_OK;
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

// base address + size of the cudaMalloc allocation containing p (driver entry point, no -lcuda needed)
static int allocation_range(uintptr_t p, uintptr_t* base, size_t* size)
{
  typedef int (*fn_t)(unsigned long long*, size_t*, unsigned long long);
  static fn_t fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q);
    if (e != cudaSuccess || !f) return fail_cuda(e, "cudaGetDriverEntryPoint(cuMemGetAddressRange)");
    fn = reinterpret_cast<fn_t>(f);
  }
  unsigned long long b = 0;
  size_t s = 0;
  int r = fn(&b, &s, static_cast<unsigned long long>(p));
  if (r != 0) return fail(KVBM_ERR_CUDA, "cuMemGetAddressRange failed: " + std::to_string(r));
  *base = static_cast<uintptr_t>(b);
  *size = s;
  return KVBM_OK;
}

extern "C" int kvbm_manager_export_metadata(kvbm_transfer_manager* m, kvbm_layout_handle h, void* buf, size_t cap, size_t* len)
{
  try {
  if (!m || !len) return fail(KVBM_ERR, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  Layout* L = m->find(h);
  if (!L) return fail(KVBM_ERR_HANDLE, "invalid handle");
  const size_t need = sizeof(BlobHeader) + L->allocs.size() * sizeof(BlobAlloc);
  *len = need;
  if (!buf) return KVBM_OK;
  if (cap < need) return fail(KVBM_ERR, "buffer too small for metadata");
  BlobHeader hd{};
  std::memcpy(hd.magic, kMagic, 8);
  hd.version = kBlobVersion;
  hd.fully_contiguous = L->fully_contiguous;
  hd.block_dim = static_cast<uint32_t>(L->block_dim);
  hd.storage = static_cast<uint32_t>(L->storage);
  hd.device_id = L->device_id;
  hd.n_allocs = static_cast<uint32_t>(L->allocs.size());
  hd.worker_id = m->worker_id;
  hd.pid = process_identity();
  const kvbm_layout_config& c = L->cfg;
  const uint64_t cfg[9] = {c.num_blocks, c.num_layers, c.outer_dim, c.page_size, c.inner_dim, c.alignment, c.dtype_width_bytes, c.num_heads,
                           static_cast<uint64_t>(c.allow_fp8 ? 1 : 0) | (static_cast<uint64_t>(L->kv_block_layout) << 8)};
  std::memcpy(hd.cfg, cfg, sizeof(cfg));
  auto* p = static_cast<unsigned char*>(buf);
  std::memcpy(p, &hd, sizeof(hd));
  p += sizeof(hd);
  DeviceGuard g(L->storage == KVBM_STORAGE_DEVICE ? L->device_id : -1);
  for (const Allocation& a : L->allocs) {
    BlobAlloc ba{};
    ba.addr = a.addr;
    ba.size = a.size;
    if (L->storage == KVBM_STORAGE_DEVICE && !L->remote) {
      uintptr_t base;
      size_t sz;
      int rc = allocation_range(a.addr, &base, &sz);
      if (rc) return rc;
      cudaIpcMemHandle_t ih;
      cudaError_t e = cudaIpcGetMemHandle(&ih, reinterpret_cast<void*>(base));
      if (e != cudaSuccess) return fail_cuda(e, "cudaIpcGetMemHandle");
      static_assert(sizeof(ih) == 64, "cudaIpcMemHandle_t is 64 bytes");
      std::memcpy(ba.ipc, &ih, 64);
      ba.has_ipc = 1;
      ba.offset_in_ipc = a.addr - base;
    }
    std::memcpy(p, &ba, sizeof(ba));
    p += sizeof(ba);
  }
  return KVBM_OK;
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

static int import_metadata_impl(kvbm_transfer_manager* m, const void* buf, size_t len, const void* const* local_bases, size_t num_local_bases,
                                kvbm_layout_handle* out);

extern "C" int kvbm_manager_import_metadata(kvbm_transfer_manager* m, const void* buf, size_t len, kvbm_layout_handle* out)
{
  try {
  return import_metadata_impl(m, buf, len, nullptr, 0, out);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

// Host pools of another process that THIS process has mapped itself (POSIX / SysV shared memory, a hugetlbfs file, memory
// shared across fork): the importer names its own view of every allocation.  The CPU twin of the CUDA-IPC mapping above --
// what makes BASELINE configs[0] (two worker processes, CPU memcpy hand-off) a one-sided push like the NVLink path.
extern "C" int kvbm_manager_import_metadata_mapped(kvbm_transfer_manager* m, const void* buf, size_t len, const void* const* local_bases,
                                                   size_t num_local_bases, kvbm_layout_handle* out)
{
  try {
  if (!local_bases || num_local_bases == 0) return fail(KVBM_ERR, "local_bases: one address per allocation of the layout");
  return import_metadata_impl(m, buf, len, local_bases, num_local_bases, out);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

static int import_metadata_impl(kvbm_transfer_manager* m, const void* buf, size_t len, const void* const* local_bases, size_t num_local_bases,
                                kvbm_layout_handle* out)
{
  if (!m || !buf || !out) return fail(KVBM_ERR, "null argument");
  if (len < sizeof(BlobHeader)) return fail(KVBM_ERR, "metadata blob truncated");
  BlobHeader hd;
  std::memcpy(&hd, buf, sizeof(hd));
  if (std::memcmp(hd.magic, kMagic, 8) != 0) return fail(KVBM_ERR, "not a KVBM layout blob");
  if (hd.version != kBlobVersion)  // layout/serialize.rs version check
    return fail(KVBM_ERR_VERSION, "Unsupported layout descriptor version " + std::to_string(hd.version) + " (expected " + std::to_string(kBlobVersion) + ")");
  if (len < sizeof(BlobHeader) + hd.n_allocs * sizeof(BlobAlloc)) return fail(KVBM_ERR, "metadata blob truncated");
  kvbm_layout_config cfg{};
  cfg.num_blocks = hd.cfg[0];
  cfg.num_layers = hd.cfg[1];
  cfg.outer_dim = hd.cfg[2];
  cfg.page_size = hd.cfg[3];
  cfg.inner_dim = hd.cfg[4];
  cfg.alignment = hd.cfg[5];
  cfg.dtype_width_bytes = hd.cfg[6];
  cfg.num_heads = hd.cfg[7];
  cfg.allow_fp8 = static_cast<int>(hd.cfg[8] & 0xff);
  const bool same_process = hd.pid == process_identity();
  std::vector<uintptr_t> bases(hd.n_allocs);
  std::vector<size_t> sizes(hd.n_allocs);
  bool unmapped = false;
  const auto* pa = static_cast<const unsigned char*>(buf) + sizeof(BlobHeader);
  DeviceGuard g(m->device);
  for (uint32_t i = 0; i < hd.n_allocs; ++i) {
    BlobAlloc ba;
    std::memcpy(&ba, pa + i * sizeof(BlobAlloc), sizeof(ba));
    sizes[i] = ba.size;
    if (same_process) {
      bases[i] = ba.addr;  // same address space (peer access is enabled with kvbm_manager_enable_peer_access)
    } else if (!ba.has_ipc && local_bases) {
      if (hd.storage == KVBM_STORAGE_DEVICE) return fail(KVBM_ERR_UNSUPPORTED, "device allocations travel as CUDA IPC handles, not as caller-provided mappings");
      if (num_local_bases != hd.n_allocs)
        return fail(KVBM_ERR, "local_bases has " + std::to_string(num_local_bases) + " entries, the layout has " + std::to_string(hd.n_allocs) + " allocations");
      if (!local_bases[i]) return fail(KVBM_ERR, "null local mapping for allocation " + std::to_string(i));
      bases[i] = reinterpret_cast<uintptr_t>(local_bases[i]);  // this process's own mapping of the peer's shared memory
    } else if (!ba.has_ipc) {
      // Another process's System / Pinned pool (or device memory that could not be exported): its virtual addresses mean
      // nothing here.  The layout is registered as a DESCRIPTOR (geometry, handle, memory_region arithmetic) that no
      // transfer may touch -- the role NIXL's remote-descriptor-only entries play (manager/mod.rs:519-633).
      bases[i] = ba.addr;
      unmapped = true;
    } else {
      if (m->device < 0) return fail(KVBM_ERR_CUDA, "importing a device layout needs a CUDA manager");
      cudaIpcMemHandle_t ih;
      std::memcpy(&ih, ba.ipc, 64);
      const std::string key(reinterpret_cast<const char*>(ba.ipc), 64);
      void* mapped = nullptr;
      {
        std::lock_guard<std::mutex> lk(m->mu);
        auto it = m->ipc_cache.find(key);
        if (it != m->ipc_cache.end()) mapped = it->second;
      }
      if (!mapped) {
        cudaError_t e = cudaIpcOpenMemHandle(&mapped, ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail_cuda(e, "cudaIpcOpenMemHandle");
        std::lock_guard<std::mutex> lk(m->mu);
        m->ipc_cache[key] = mapped;
      }
      bases[i] = reinterpret_cast<uintptr_t>(mapped) + ba.offset_in_ipc;
    }
  }
  Layout L;
  std::string why;
  int rc = hd.fully_contiguous ? make_fully_contiguous(cfg, bases.empty() ? 0 : bases[0], sizes.empty() ? 0 : sizes[0], &L, &why)
                               : make_layer_separate(cfg, bases.data(), sizes.data(), hd.n_allocs, static_cast<int>(hd.block_dim), &L, &why);
  if (rc) return fail(rc, why);
  L.remote = !same_process;
  L.unmapped = unmapped;
  L.kv_block_layout = static_cast<int>((hd.cfg[8] >> 8) & 0xff);
  return finish_register(m, std::move(L), static_cast<int>(hd.storage), hd.device_id, out);
}

// FullyContiguousLayoutBuilder::kv_block_layout / LayerSeparateLayoutBuilder::inner_shape (fully_contiguous.rs:83-88,
// layer_separate.rs:91-101): the format of one block of a registered layout; Unknown until set.
extern "C" int kvbm_manager_set_kv_block_layout(kvbm_transfer_manager* m, kvbm_layout_handle h, int kv_layout)
{
  try {
  if (!m) return fail(KVBM_ERR, "null manager");
  if (kv_layout < KVBM_KV_UNKNOWN || kv_layout > KVBM_KV_OPERATIONAL_NHD) return fail(KVBM_ERR, "unknown KvBlockLayout");
  std::lock_guard<std::mutex> lk(m->mu);
  Layout* L = m->find(h);
  if (!L) return fail(KVBM_ERR_HANDLE, "invalid layout handle");
  if (kv_layout != KVBM_KV_UNKNOWN) {
    // LayoutConfig::validate_for_kv_block_layout (config.rs:114-140); the head dimension is inner_dim / num_heads, the
    // figure consistent with required_bytes (:64-71) and with the chunk size K2 / K3 take
    if (!L->cfg.num_heads) return fail(KVBM_ERR_CONFIG, "num_heads_required_for_kv_block_layout");
    if (L->cfg.inner_dim % L->cfg.num_heads) return fail(KVBM_ERR_CONFIG, "inner_dim_must_be_divisible_by_num_heads");
    if ((kv_layout == KVBM_KV_UNIVERSAL_TP || kv_layout == KVBM_KV_UNIVERSAL_PP) && !L->fully_contiguous)
      return fail(KVBM_ERR_CONFIG, "universal KV block layouts exist only on fully contiguous layouts");  // layer_separate.rs:91-101
  }
  L->kv_block_layout = kv_layout;
  return KVBM_OK;
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_manager_kv_block_layout(kvbm_transfer_manager* m, kvbm_layout_handle h)
{
  if (!m) return -1;
  std::lock_guard<std::mutex> lk(m->mu);
  Layout* L = m->find(h);
  return L ? L->kv_block_layout : -1;
}

extern "C" int kvbm_select_transform_kernel(int src_kv_layout, int dst_kv_layout) { return select_transform_kernel(src_kv_layout, dst_kv_layout); }
extern "C" int kvbm_kv_layout_requires_transform(int a, int b) { return kv_requires_transform(a, b) ? 1 : 0; }

extern "C" int kvbm_manager_set_capabilities(kvbm_transfer_manager* m, const kvbm_transfer_capabilities* caps)
{
  if (!m || !caps) return fail(KVBM_ERR, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  m->caps = *caps;
  return KVBM_OK;
}

extern "C" int kvbm_manager_execute_transfer(kvbm_transfer_manager* m, kvbm_layout_handle src, const size_t* src_ids, kvbm_layout_handle dst,
                                             const size_t* dst_ids, size_t n, const kvbm_transfer_options* opts, kvbm_notification* out)
{
  try {
  if (!m) return fail(KVBM_ERR, "null manager");
  const size_t* s[1] = {src_ids};
  const size_t* d[1] = {dst_ids};
  return execute(m, src, 1, &dst, s, d, n, true, opts, out);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_manager_execute_fanout(kvbm_transfer_manager* m, kvbm_layout_handle src, int num_dsts, const kvbm_layout_handle* dsts,
                                           const size_t* const* src_ids, const size_t* const* dst_ids, size_t n, int replicate,
                                           const kvbm_transfer_options* opts, kvbm_notification* out)
{
  try {
  if (!m || !dsts) return fail(KVBM_ERR, "null argument");
  bool rep = replicate != 0;
  if (!rep && src_ids) {
    rep = true;
    for (int d = 1; d < num_dsts; ++d)
      if (src_ids[d] != src_ids[0]) rep = false;
  }
  return execute(m, src, num_dsts, dsts, src_ids, dst_ids, n, rep, opts, out);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_notification_is_complete(kvbm_transfer_manager* m, kvbm_notification n)
{
  if (!m) return -1;
  if (n == 0) return 1;
  // Lock-free against execute() recycling the slot: the owner word is atomic, the completion word lives at a fixed
  // address for the manager's lifetime, and the owner is re-checked after the completion word was read.
  Slot& sl = m->slots[(n - 1) % kSlots];
  const uint64_t s1 = sl.seq.load(std::memory_order_acquire);
  if (s1 != n) return s1 > n ? 1 : -1;  // slot recycled by a later transfer => ours completed
  const uint32_t v = __atomic_load_n(sl.host_flag, __ATOMIC_ACQUIRE);
  if (sl.seq.load(std::memory_order_acquire) != n) return 1;
  if (v == 0xFFFFFFFFu) return -2;  // the gated kernel gave up waiting for a layer_ready flag
  return v == static_cast<uint32_t>(n) ? 1 : 0;
}

extern "C" int kvbm_notification_wait(kvbm_transfer_manager* m, kvbm_notification n, int64_t timeout_us)
{
  try {
  if (!m) return fail(KVBM_ERR, "null manager");
  const auto t0 = std::chrono::steady_clock::now();
  unsigned spins = 0;
  for (;;) {
    int c = kvbm_notification_is_complete(m, n);
    if (c == 1) return KVBM_OK;
    if (c == -2) return fail(KVBM_ERR_TIMEOUT, "gated transfer aborted: a layer_ready flag was not released within the gate timeout");
    if (c < 0) return fail(KVBM_ERR_HANDLE, "unknown notification");
    if ((++spins & 63) == 0) {
      if (timeout_us >= 0 &&
          std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() > timeout_us)
        return fail(KVBM_ERR_TIMEOUT, "transfer did not complete in time");
      sched_yield();
    }
  }
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" uint64_t kvbm_manager_bytes_moved(kvbm_transfer_manager* m) { return m ? m->bytes_moved.load() : 0; }
extern "C" uint64_t kvbm_manager_h2d_bytes(kvbm_transfer_manager* m) { return m ? m->h2d_bytes.load() : 0; }

// =====================================================================================================
// SerializedLayout / LayoutDescriptor in the reference's wire formats; codec in serialized_layout.hpp
// =====================================================================================================
namespace kvbm_host {

static const char kTransportMagic[8] = {'K', 'V', 'B', 'M', 'I', 'P', 'C', '1'};

static std::string agent_name_of(uint64_t worker_id) { return "kvbm-b200-ipc-" + std::to_string(worker_id); }

// PhysicalLayout::to_descriptor (layout/physical.rs:163-187)
static kvbm_wire::Descriptor to_wire(const Layout& L, const std::string& agent)
{
  kvbm_wire::Descriptor d;
  d.version = 1;
  d.num_blocks = L.cfg.num_blocks;
  d.num_layers = L.cfg.num_layers;
  d.outer_dim = L.cfg.outer_dim;
  d.page_size = L.cfg.page_size;
  d.inner_dim = L.cfg.inner_dim;
  d.alignment = L.cfg.alignment == 0 ? 1 : L.cfg.alignment;
  d.dtype_width_bytes = L.cfg.dtype_width_bytes;
  d.has_num_heads = L.cfg.num_heads != 0;
  d.num_heads = L.cfg.num_heads;
  switch (L.storage) {
    case KVBM_STORAGE_PINNED: d.location = kvbm_wire::kPinned; break;
    case KVBM_STORAGE_DEVICE:
      d.location = kvbm_wire::kDevice;
      d.location_arg = static_cast<uint64_t>(L.device_id);
      break;
    case KVBM_STORAGE_DISK: d.location = kvbm_wire::kDisk; break;
    default: d.location = kvbm_wire::kSystem;
  }
  d.agent_name = agent;
  d.mem_type = L.storage == KVBM_STORAGE_DEVICE ? kvbm_wire::kVram : L.storage == KVBM_STORAGE_DISK ? kvbm_wire::kFile : kvbm_wire::kDram;
  d.nixl_device_id = L.storage == KVBM_STORAGE_DEVICE ? static_cast<uint64_t>(L.device_id) : 0;
  for (const Allocation& a : L.allocs) d.regions.push_back({a.addr, a.size});
  d.fully_contiguous = L.fully_contiguous;
  d.block_dim = L.block_dim == KVBM_BLOCK_IS_SECOND_DIM ? 1u : 0u;
  d.kv_block_layout = L.kv_block_layout >= KVBM_KV_UNIVERSAL_TP && L.kv_block_layout <= KVBM_KV_CUSTOM ? static_cast<uint32_t>(L.kv_block_layout - 1) : kvbm_wire::kKvUnknown;
  return d;
}

static kvbm_layout_config config_of(const kvbm_wire::Descriptor& d)
{
  kvbm_layout_config c{};
  c.num_blocks = d.num_blocks;
  c.num_layers = d.num_layers;
  c.outer_dim = d.outer_dim;
  c.page_size = d.page_size;
  c.inner_dim = d.inner_dim;
  c.alignment = d.alignment;
  c.dtype_width_bytes = d.dtype_width_bytes;
  c.num_heads = d.has_num_heads ? d.num_heads : 0;
  c.allow_fp8 = d.dtype_width_bytes == 1;
  return c;
}

// the structural checks of PhysicalLayout::from_descriptor (layout/physical.rs:203-262)
static int check_descriptor(const kvbm_wire::Descriptor& d)
{
  if (d.version > 1) return fail(KVBM_ERR_VERSION, "Unsupported serialization version: " + std::to_string(d.version) + ". Maximum supported: 1");
  if (d.fully_contiguous && d.regions.size() != 1)
    return fail(KVBM_ERR_CONFIG, "FullyContiguous layout requires exactly 1 memory region, got " + std::to_string(d.regions.size()));
  if (!d.fully_contiguous && d.regions.size() != d.num_layers)
    return fail(KVBM_ERR_CONFIG, "LayerSeparate layout requires " + std::to_string(d.num_layers) + " memory regions (one per layer), got " +
                                     std::to_string(d.regions.size()));
  return KVBM_OK;
}

static int copy_out(const void* data, size_t need, void* buf, size_t cap, size_t* len)
{
  *len = need;
  if (!buf) return KVBM_OK;
  if (cap < need) return fail(KVBM_ERR, "buffer too small");
  std::memcpy(buf, data, need);
  return KVBM_OK;
}

}  // namespace kvbm_host

extern "C" int kvbm_layout_descriptor_json(kvbm_transfer_manager* m, kvbm_layout_handle h, char* buf, size_t cap, size_t* len)
{
  try {
  if (!m || !len) return fail(KVBM_ERR, "null argument");
  std::string js;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    Layout* L = m->find(h);
    if (!L) return fail(KVBM_ERR_HANDLE, "invalid handle");
    js = kvbm_wire::descriptor_to_json(to_wire(*L, agent_name_of(m->worker_id)));
  }
  return copy_out(js.data(), js.size(), buf, cap, len);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

extern "C" int kvbm_manager_import_descriptor_json(kvbm_transfer_manager* m, const char* json, size_t len, kvbm_layout_handle* out)
{
  try {
  if (!m || !json || !out) return fail(KVBM_ERR, "null argument");
  kvbm_wire::Descriptor d;
  std::string why;
  if (!kvbm_wire::descriptor_from_json(json, len, &d, &why)) return fail(KVBM_ERR, why);
  int rc = check_descriptor(d);
  if (rc) return rc;
  const kvbm_layout_config cfg = config_of(d);
  Layout L;
  if (d.fully_contiguous) {
    rc = make_fully_contiguous(cfg, static_cast<uintptr_t>(d.regions[0].addr), d.regions[0].size, &L, &why);
  } else {
    std::vector<uintptr_t> bases;
    std::vector<size_t> sizes;
    for (const auto& g : d.regions) {
      bases.push_back(static_cast<uintptr_t>(g.addr));
      sizes.push_back(g.size);
    }
    rc = make_layer_separate(cfg, bases.data(), sizes.data(), bases.size(), d.block_dim ? KVBM_BLOCK_IS_SECOND_DIM : KVBM_BLOCK_IS_FIRST_DIM, &L, &why);
  }
  if (rc) return fail(rc, why);
  L.kv_block_layout = d.kv_block_layout <= kvbm_wire::kCustom ? static_cast<int>(d.kv_block_layout) + 1 : KVBM_KV_UNKNOWN;
  const int storage = d.location == kvbm_wire::kDevice ? KVBM_STORAGE_DEVICE
                      : d.location == kvbm_wire::kPinned ? KVBM_STORAGE_PINNED
                      : d.location == kvbm_wire::kDisk   ? KVBM_STORAGE_DISK
                                                         : KVBM_STORAGE_SYSTEM;
  return finish_register(m, std::move(L), storage, d.location == kvbm_wire::kDevice ? static_cast<int>(d.location_arg) : 0, out);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

// TransferManager::export_metadata (manager/mod.rs:112, registry :519-556): every local host / device layout, one blob.
// `nixl_metadata` carries this library's transport records instead of a NIXL agent's: per layout one KVBMLAY1 blob with a
// CUDA IPC handle per allocation, so the importing process can map the pools and reach them over NVLink.
extern "C" int kvbm_manager_export_serialized_layout(kvbm_transfer_manager* m, void* buf, size_t cap, size_t* len)
{
  try {
  if (!m || !len) return fail(KVBM_ERR, "null argument");
  std::vector<kvbm_layout_handle> handles;
  kvbm_wire::Bundle b;
  b.worker_id = m->worker_id;
  b.agent_name = agent_name_of(m->worker_id);
  {
    std::lock_guard<std::mutex> lk(m->mu);
    for (auto& kv : m->layouts)
      if (!kv.second->remote && kv.second->storage != KVBM_STORAGE_DISK) handles.push_back(kv.first);
    std::sort(handles.begin(), handles.end());
    for (kvbm_layout_handle h : handles) {
      kvbm_wire::Logical l;
      l.worker_id = m->worker_id;
      l.layout_id = static_cast<uint16_t>(h & 0xffff);
      l.logical_type = kvbm_wire::kG2;  // LocalLayoutDescriptor::new_with_default_type (manager/metadata.rs:70-76)
      l.layout = to_wire(*m->find(h), b.agent_name);
      b.layouts.push_back(std::move(l));
    }
  }
  kvbm_wire::Writer t;
  t.raw(kTransportMagic, 8);
  t.le(handles.size(), 4);
  for (kvbm_layout_handle h : handles) {
    size_t need = 0;
    int rc = kvbm_manager_export_metadata(m, h, nullptr, 0, &need);
    if (rc) return rc;
    std::vector<uint8_t> one(need);
    rc = kvbm_manager_export_metadata(m, h, one.data(), one.size(), &need);
    if (rc) return rc;
    t.le(one.size(), 4);
    t.raw(one.data(), one.size());
  }
  b.nixl_metadata = std::move(t.out);
  const std::vector<uint8_t> bytes = kvbm_wire::encode_bundle(b);
  return copy_out(bytes.data(), bytes.size(), buf, cap, len);
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

// TransferManager::import_metadata (manager/mod.rs:130, registry :572-633)
extern "C" int kvbm_manager_import_serialized_layout(kvbm_transfer_manager* m, const void* buf, size_t len, kvbm_layout_handle* out, size_t cap,
                                                     size_t* n_out)
{
  try {
  if (!m || !buf || !n_out) return fail(KVBM_ERR, "null argument");
  kvbm_wire::Bundle b;
  std::string why;
  if (!kvbm_wire::decode_bundle(buf, len, &b, &why)) return fail(KVBM_ERR, why);
  const auto key = std::make_pair(b.agent_name, b.worker_id);
  {
    std::lock_guard<std::mutex> lk(m->mu);
    if (m->loaded_remotes.count(key))
      return fail(KVBM_ERR, "Remote worker already loaded: " + b.agent_name + " (worker_id=" + std::to_string(b.worker_id) + ")");
  }
  *n_out = b.layouts.size();
  if (!out) return KVBM_OK;  // size query
  if (cap < b.layouts.size()) return fail(KVBM_ERR, "handle array too small");
  // transport records
  const std::vector<uint8_t>& t = b.nixl_metadata;
  if (t.size() < 12 || std::memcmp(t.data(), kTransportMagic, 8) != 0)
    return fail(KVBM_ERR_UNSUPPORTED, "failed to load remote NIXL metadata: the blob carries no KVBM IPC transport records (a NIXL agent's "
                                      "metadata cannot be used by this library)");
  uint32_t count;
  std::memcpy(&count, t.data() + 8, 4);
  if (count != b.layouts.size()) return fail(KVBM_ERR, "transport record count does not match the layout count");
  size_t off = 12;
  std::vector<kvbm_layout_handle> done;
  auto rollback = [&](int rc) {
    const std::string msg = kvbm_last_error();
    for (kvbm_layout_handle h : done) kvbm_manager_unregister(m, h);
    return fail(rc, msg);
  };
  for (size_t i = 0; i < b.layouts.size(); ++i) {
    const kvbm_wire::Descriptor& d = b.layouts[i].layout;
    int rc = check_descriptor(d);
    if (rc) return rollback(rc);
    uint32_t blen;
    if (t.size() - off < 4) return rollback(fail(KVBM_ERR, "transport records truncated"));
    std::memcpy(&blen, t.data() + off, 4);
    off += 4;
    if (t.size() - off < blen) return rollback(fail(KVBM_ERR, "transport records truncated"));
    kvbm_layout_handle h = 0;
    rc = kvbm_manager_import_metadata(m, t.data() + off, blen, &h);
    off += blen;
    if (rc) return rollback(rc);
    done.push_back(h);
    // the descriptor and its transport record must describe the same layout
    bool same;
    {
      std::lock_guard<std::mutex> lk(m->mu);
      const Layout* L = m->find(h);
      const kvbm_layout_config c = config_of(d);
      same = L && L->cfg.num_blocks == c.num_blocks && L->cfg.num_layers == c.num_layers && L->cfg.outer_dim == c.outer_dim &&
             L->cfg.page_size == c.page_size && L->cfg.inner_dim == c.inner_dim && L->cfg.dtype_width_bytes == c.dtype_width_bytes &&
             L->fully_contiguous == d.fully_contiguous && L->allocs.size() == d.regions.size() &&
             (d.fully_contiguous || (L->block_dim == KVBM_BLOCK_IS_SECOND_DIM) == (d.block_dim == 1));
    }
    if (!same) return rollback(fail(KVBM_ERR_INCOMPATIBLE, "failed to reconstruct layout: descriptor and transport record disagree"));
  }
  for (size_t i = 0; i < done.size(); ++i) out[i] = done[i];
  std::lock_guard<std::mutex> lk(m->mu);
  m->loaded_remotes.insert(key);
  return KVBM_OK;
  } catch (const std::exception& e) {
    return kvbm_host::set_last_error(KVBM_ERR, std::string("internal error: ") + e.what());
  } catch (...) {
    return kvbm_host::set_last_error(KVBM_ERR, "internal error");
  }
}

/*
 * kvbm_physical.h -- C ABI of libkvbm_physical.so: the host half of the KV transfer path.
 *
 * It restates, in C++ behind plain-C entry points, what Dynamo's Rust crate lib/kvbm-physical does for
 * this path (no Rust toolchain exists in the build image; see INTEGRATION.md for the Rust `extern "C"`
 * block a maintainer would add):
 *
 *   LayoutConfig + validation            lib/kvbm-physical/src/layout/config.rs:14-56,165-180
 *   FullyContiguous / LayerSeparate      layout/fully_contiguous.rs:141-185, layout/layer_separate.rs:155-210
 *   Layout::memory_region                layout/mod.rs:73-78
 *   validate_block_transfer              transfer/validation.rs:55-225
 *   select_strategy / direct strategy    transfer/strategy.rs:78-108,138-210
 *   TransferOptions                      transfer/options.rs:27-81
 *   TransferManager::{register_layout, export_metadata, import_metadata, execute_transfer}
 *                                        manager/mod.rs:101,112,130,227
 *   TransferCompleteNotification         transfer/context.rs:438-470 (event -> 1 ms poller) -- replaced by an
 *                                        in-band completion flag the kernel writes to pinned host memory
 *   layer-wise onboard                   lib/kvbm-engine/src/worker/physical.rs:277-346
 *   CollectiveOps::broadcast             lib/kvbm-engine/src/collectives/mod.rs:75-106 (ncclBcast per region)
 *                                        -- replaced by one replicate launch over NVLink peer mappings
 *
 * Threading: a kvbm_transfer_manager may be used from any number of threads concurrently (the reference is called from
 * arbitrary tokio workers); checked under ThreadSanitizer by tests/c/race_check.cpp.
 * Error style: every call returns KVBM_OK (0) or an error code; kvbm_last_error() gives the thread's last
 * message (house style of lib/bindings/c/src/lib.rs:74-78: OK=0, ERR!=0).  Nothing aborts or throws.
 */
#ifndef KVBM_PHYSICAL_H
#define KVBM_PHYSICAL_H

#include <stddef.h>
#include <stdint.h>

#include "kvbm_kernels.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
  KVBM_OK = 0,
  KVBM_ERR = 1,                 /* generic */
  KVBM_ERR_CONFIG = 2,          /* LayoutConfig validation failed / memory too small */
  KVBM_ERR_RANGE = 3,           /* block / layer / outer id out of range */
  KVBM_ERR_LENGTH_MISMATCH = 4, /* BlockValidationError::LengthMismatch */
  KVBM_ERR_DUPLICATE_DST = 5,   /* BlockValidationError::DuplicateDestinationBlocks */
  KVBM_ERR_OVERLAP = 6,         /* BlockValidationError::OverlappingBlocks */
  KVBM_ERR_INCOMPATIBLE = 7,    /* layer count / outer dim / region size mismatch */
  KVBM_ERR_UNSUPPORTED = 8,     /* strategy this library does not implement (NIXL, disk, System<->Device) */
  KVBM_ERR_CUDA = 9,            /* a CUDA call failed (message carries the cudaError_t) */
  KVBM_ERR_HANDLE = 10,         /* unknown layout handle / notification */
  KVBM_ERR_TIMEOUT = 11,
  KVBM_ERR_VERSION = 12,        /* metadata blob version mismatch (layout/serialize.rs version check) */
};

/* dynamo_memory::StorageKind (lib/memory/src/lib.rs) */
enum { KVBM_STORAGE_SYSTEM = 0, KVBM_STORAGE_PINNED = 1, KVBM_STORAGE_DEVICE = 2, KVBM_STORAGE_DISK = 3 };
/* BlockDimension (layout/config.rs:151-163) */
enum { KVBM_BLOCK_IS_FIRST_DIM = 0, KVBM_BLOCK_IS_SECOND_DIM = 1 };
/* TransferStrategy (transfer/strategy.rs:20-58) */
enum {
  KVBM_STRATEGY_MEMCPY = 0,
  KVBM_STRATEGY_CUDA_ASYNC_H2D = 1,
  KVBM_STRATEGY_CUDA_ASYNC_D2H = 2,
  KVBM_STRATEGY_CUDA_ASYNC_D2D = 3,
  KVBM_STRATEGY_NIXL_READ = 4,
  KVBM_STRATEGY_NIXL_WRITE = 5,
  KVBM_STRATEGY_NIXL_READ_FLIPPED = 6,
  KVBM_STRATEGY_INVALID = 7,
};

typedef struct kvbm_layout_config {
  size_t num_blocks;
  size_t num_layers;
  size_t outer_dim;         /* 1 or 2 */
  size_t page_size;
  size_t inner_dim;
  size_t alignment;         /* power of two; 0 is read as 1 (the builder default) */
  size_t dtype_width_bytes; /* 2,4,8 as the reference; 1 accepted only with allow_fp8 */
  size_t num_heads;         /* 0 = None */
  int allow_fp8;            /* extension: fp8 KV pools (dtype_width_bytes == 1) */
} kvbm_layout_config;

/* A transfer plan as select_strategy returns it (transfer/strategy.rs:60-76). */
typedef struct kvbm_transfer_plan {
  int two_hop;          /* 0 = Direct(first) */
  int first;            /* KVBM_STRATEGY_* */
  int bounce_location;  /* KVBM_STORAGE_* (two-hop only) */
  int second;           /* two-hop only */
} kvbm_transfer_plan;

typedef struct kvbm_transfer_capabilities {
  int allow_gds;
  int allow_gpu_rdma;
} kvbm_transfer_capabilities;

/* TransferOptions (transfer/options.rs:27-81) + the cast / streaming extensions. */
typedef struct kvbm_transfer_options {
  int has_layer_range;
  size_t layer_begin, layer_end;
  cudaStream_t cuda_stream;     /* non-NULL: caller manages sync, notification is already complete */
  int use_caller_stream;        /* set when cuda_stream is meaningful (NULL is the legacy default stream) */
  int cast_mode;                /* KVBM_CAST_* */
  int max_ctas;                 /* 0 = whole GPU; else cap so attention keeps its SMs */
  const uint32_t* layer_ready_flags; /* nullable: device flags released by the producer (attention) per layer */
  uint32_t* layer_done_flags;        /* nullable: device-visible flags on the destination, set per layer */
  uint32_t epoch;
  int gate_timeout_ms;               /* 0 = 10 s; see kvbm_paged_copy_opts.gate_timeout_ms */
  int multicast;                     /* non-zero: the destination layout is registered over a kvbm_mc_group_map() range */
  uint32_t* done_flag;               /* nullable: word in the DESTINATION GPU's memory that receives `epoch` once every byte of
                                        the transfer has landed -- the role of TransferOptions::nixl_write_notification
                                        (options.rs:36-43: "delivered to the remote node after the RDMA write completes") */
  /* BounceBuffer (options.rs:45-51): a registered layout + the blocks of it a two-hop transfer may stage through */
  uint64_t bounce_layout;            /* kvbm_layout_handle; 0 = none */
  const size_t* bounce_block_ids;
  size_t num_bounce_blocks;
  int src_kv_layout, dst_kv_layout;  /* KvBlockLayout overrides, KVBM_KV_* (options.rs:63-80): 0 = the layout's own.  A pair that
                                        needs a transformation is EXECUTED as one permuting launch (the reference rejects it,
                                        transfer/mod.rs:128-147): see kvbm_select_transform_kernel */
  int gate_mode;                     /* see kvbm_paged_copy_opts.gate_mode */
  /* fan-out only: per-destination signals (arrays of num_dsts pointers, entries may be NULL).  When set they replace
   * done_flag / layer_done_flags, which only reach destination 0.  With `multicast` they name the flags of every receiver of
   * the one multicast write -- what a staged broadcast needs (disagg.StagedBroadcast). */
  uint32_t* const* per_dst_done_flags;
  uint32_t* const* per_dst_layer_done_flags;
} kvbm_transfer_options;

typedef struct kvbm_transfer_manager kvbm_transfer_manager;
typedef uint64_t kvbm_layout_handle;   /* (worker_id << 16) | layout_id, cf. manager/handle.rs:16-50 */
typedef uint64_t kvbm_notification;    /* 0 = already complete */

const char* kvbm_last_error(void);

/* ---- pure host logic (no GPU needed) ---- */
int kvbm_layout_config_validate(const kvbm_layout_config* cfg);
size_t kvbm_layout_required_bytes(const kvbm_layout_config* cfg);   /* config.rs:64-71 */
size_t kvbm_layout_bytes_per_block(const kvbm_layout_config* cfg);  /* config.rs:76-82 */
/* select_direct_strategy (strategy.rs:138-210).  Device ids are ignored exactly as strategy.rs:168 does. */
int kvbm_select_direct_strategy(int src_kind, int dst_kind, const kvbm_transfer_capabilities* caps,
                                kvbm_transfer_plan* out);
/* The same with select_direct_strategy's `dst_is_remote` argument (strategy.rs:138-150, select_remote_strategy :213-243): a local
 * source to another agent's memory is NixlWrite, staged through Pinned for Device (without GPU RDMA) and Disk. */
int kvbm_select_direct_strategy_remote(int src_kind, int dst_kind, int dst_is_remote, const kvbm_transfer_capabilities* caps,
                                       kvbm_transfer_plan* out);
/* select_strategy (strategy.rs:78-108) with select_remote_strategy_v2 (:245-281): both local = the table above; exactly one
 * side local = NixlWrite (push) or NixlReadFlipped (pull) -- executed here as the push / pull launch over the IPC-mapped pool;
 * both remote, Disk, or a Device side without allow_gpu_rdma = the reference's errors. */
int kvbm_select_strategy(int src_kind, int src_is_local, int dst_kind, int dst_is_local, const kvbm_transfer_capabilities* caps,
                         kvbm_transfer_plan* out);
/* validate_block_transfer, debug flavour (validation.rs:168-225). */
int kvbm_validate_block_transfer(const size_t* src_ids, size_t n_src, const size_t* dst_ids, size_t n_dst,
                                 size_t src_num_blocks, size_t dst_num_blocks, int same_layout);

/* ---- TransferManager ---- */
/* cuda_device_id < 0: host-only manager (Memcpy strategy only; every CUDA strategy returns KVBM_ERR_CUDA). */
int kvbm_manager_create(int cuda_device_id, uint64_t worker_id, kvbm_transfer_manager** out);
void kvbm_manager_destroy(kvbm_transfer_manager* m);

/* register_layout over caller-owned memory (the engine's KV tensors).  device_id is where the memory lives
 * (may differ from the manager's device: a peer GPU reachable over NVLink). */
int kvbm_manager_register_fully_contiguous(kvbm_transfer_manager* m, const kvbm_layout_config* cfg, void* base,
                                           size_t size, int storage_kind, int device_id,
                                           kvbm_layout_handle* out);
int kvbm_manager_register_layer_separate(kvbm_transfer_manager* m, const kvbm_layout_config* cfg,
                                         void* const* layer_bases, const size_t* layer_sizes, int block_dim,
                                         int storage_kind, int device_id, kvbm_layout_handle* out);
int kvbm_manager_unregister(kvbm_transfer_manager* m, kvbm_layout_handle h);
int kvbm_layout_memory_region(kvbm_transfer_manager* m, kvbm_layout_handle h, size_t block, size_t layer,
                              size_t outer, uintptr_t* addr, size_t* size);
int kvbm_layout_is_fully_contiguous(kvbm_transfer_manager* m, kvbm_layout_handle h);

/* Enable NVLink peer mappings from the manager's device to `peer_device` (same process). */
int kvbm_manager_enable_peer_access(kvbm_transfer_manager* m, int peer_device);

/* export_metadata / import_metadata (manager/mod.rs:112,130): a self-describing blob carrying the layout
 * geometry and, for Device storage, one CUDA IPC handle per allocation, so another PROCESS on the same
 * NVSwitch domain can map the pool and have the transfer kernel store into it directly.
 * kvbm_manager_export_metadata with buf == NULL returns the needed size in *len. */
int kvbm_manager_export_metadata(kvbm_transfer_manager* m, kvbm_layout_handle h, void* buf, size_t cap,
                                 size_t* len);
int kvbm_manager_import_metadata(kvbm_transfer_manager* m, const void* buf, size_t len, kvbm_layout_handle* out);
/* The same for HOST pools (System / Pinned) of another process that this process has mapped itself -- shared memory, a file
 * mapping, memory shared across fork: local_bases[i] is this process's address of the layout's i-th allocation (1 for fully
 * contiguous, num_layers for layer-separate).  Without it such a layout is imported as a descriptor only (no transfer may touch
 * it).  The CPU twin of the CUDA-IPC mapping: BASELINE configs[0]'s two-process memcpy hand-off is a one-sided push too. */
int kvbm_manager_import_metadata_mapped(kvbm_transfer_manager* m, const void* buf, size_t len, const void* const* local_bases,
                                        size_t num_local_bases, kvbm_layout_handle* out);

/* The reference's own wire formats for the handshake:
 *   kvbm_manager_export_serialized_layout / import_serialized_layout = TransferManager::export_metadata / import_metadata
 *   (manager/mod.rs:112,130): ONE blob for every local host/device layout, encoded exactly as SerializedLayout::pack
 *   (manager/metadata.rs:120-134: bincode 2 standard config over RdmaLayoutDescriptors).  Its `nixl_metadata` byte
 *   field carries this library's CUDA-IPC transport records instead of a NIXL agent's metadata; a second import of
 *   the same (agent, worker_id) fails like the reference ("Remote worker already loaded").  import with out == NULL
 *   returns the number of layouts in *n_out.
 *   kvbm_layout_descriptor_json / kvbm_manager_import_descriptor_json = LayoutDescriptor::to_json / from_json +
 *   PhysicalLayout::from_descriptor (layout/serialize.rs:96-137, layout/physical.rs:203-262); JSON-imported layouts
 *   keep the addresses as written (same address space).  Handles on the wire are the reference's u128
 *   (worker_id | layout_id << 64, manager/handle.rs:16-24); this ABI's 64-bit handle is (worker_id << 16) | layout_id. */
int kvbm_manager_export_serialized_layout(kvbm_transfer_manager* m, void* buf, size_t cap, size_t* len);
int kvbm_manager_import_serialized_layout(kvbm_transfer_manager* m, const void* buf, size_t len, kvbm_layout_handle* out,
                                          size_t cap, size_t* n_out);
int kvbm_layout_descriptor_json(kvbm_transfer_manager* m, kvbm_layout_handle h, char* buf, size_t cap, size_t* len);
int kvbm_manager_import_descriptor_json(kvbm_transfer_manager* m, const char* json, size_t len, kvbm_layout_handle* out);

/* select_strategy for two registered layouts (local = registered or imported in-process; remote = imported from another process) */
int kvbm_manager_select_strategy(kvbm_transfer_manager* m, kvbm_layout_handle src, kvbm_layout_handle dst, kvbm_transfer_plan* out);

/* TransformKernel (transfer/executor/mod.rs:27-41) as select_transform_kernel (:46-100) returns it. */
enum {
  KVBM_TRANSFORM_NONE = 0,
  KVBM_TRANSFORM_BLOCK_TO_UNIVERSAL = 1,
  KVBM_TRANSFORM_UNIVERSAL_TO_BLOCK = 2,
  KVBM_TRANSFORM_OPERATIONAL_TRANSPOSE = 3,
  KVBM_TRANSFORM_UNSUPPORTED = 4,
  KVBM_TRANSFORM_UNIVERSAL_TO_UNIVERSAL = 5, /* extension: UniversalTP <-> UniversalPP, a TODO in the reference (:87-91);
                                                never returned by kvbm_select_transform_kernel, but executed by transfers */
};
int kvbm_select_transform_kernel(int src_kv_layout, int dst_kv_layout);       /* pure; KVBM_KV_* in, KVBM_TRANSFORM_* out */
int kvbm_kv_layout_requires_transform(int a, int b);                          /* kv_block_layout.rs:107-119 */
/* The format of one block of a registered layout (FullyContiguousLayoutBuilder::kv_block_layout, fully_contiguous.rs:83-88;
 * LayerSeparateLayoutBuilder::inner_shape, layer_separate.rs:91-101).  Needs num_heads in the config; universal formats need
 * a fully contiguous layout.  Travels with export_metadata / the SerializedLayout. */
int kvbm_manager_set_kv_block_layout(kvbm_transfer_manager* m, kvbm_layout_handle h, int kv_layout);
int kvbm_manager_kv_block_layout(kvbm_transfer_manager* m, kvbm_layout_handle h);  /* KVBM_KV_*, -1 = unknown handle */

/* TransferCapabilities (transfer/strategy.rs:245-278).  Default {allow_gds 0, allow_gpu_rdma 1}.  With allow_gpu_rdma = 0 a
 * Device -> Device transfer between different GPUs follows the reference's TwoHop plan (strategy.rs:222-233):
 * CudaAsyncD2H into the bounce buffer named in the options (Pinned), then CudaAsyncH2D into the destination; the bounce blocks
 * are split into two groups and the hops of consecutive chunks overlap (executor/mod.rs:357-416,477-571).  Without a bounce
 * buffer such a transfer fails with "Two-hop transfers require a bounce buffer." as in the reference. */
int kvbm_manager_set_capabilities(kvbm_transfer_manager* m, const kvbm_transfer_capabilities* caps);

/* execute_transfer (manager/mod.rs:227-303).  ids are host arrays, consumed before return. */
int kvbm_manager_execute_transfer(kvbm_transfer_manager* m, kvbm_layout_handle src, const size_t* src_ids,
                                  kvbm_layout_handle dst, const size_t* dst_ids, size_t n,
                                  const kvbm_transfer_options* opts, kvbm_notification* out);

/* 1 -> N in ONE launch.  src_ids[d] == src_ids[0] for all d (same pointer or same contents flag
 * `replicate`) reads HBM once and stores N times: the replacement of CollectiveOps::broadcast. */
int kvbm_manager_execute_fanout(kvbm_transfer_manager* m, kvbm_layout_handle src, int num_dsts,
                                const kvbm_layout_handle* dsts, const size_t* const* src_ids,
                                const size_t* const* dst_ids, size_t n, int replicate,
                                const kvbm_transfer_options* opts, kvbm_notification* out);

/* Completion: the transfer's last warp writes a flag in pinned host memory (no cudaEventQuery polling). */
int kvbm_notification_is_complete(kvbm_transfer_manager* m, kvbm_notification n);  /* 1 / 0 / <0 error */
int kvbm_notification_wait(kvbm_transfer_manager* m, kvbm_notification n, int64_t timeout_us);

/* ---- NVLS multicast groups: identical KV blocks to N GPUs with ONE write per tile --------------------------------
 * Replaces CollectiveOps::broadcast / the grouped ncclBcast per region (lib/kvbm-engine/src/collectives/mod.rs:75-106,
 * nccl.rs:421-462).  Every receiver binds a pool allocation of the same size to one CUmulticastObject; the sender maps
 * the object and passes a layout over that mapping with kvbm_transfer_options.multicast (host API) or
 * kvbm_paged_copy_opts.multicast (kernel ABI).  Block b of the layout lands at the same offset in every bound pool.
 * Order of calls (all participating processes): create | import_fd  ->  add_device (every device, by its owner)
 * -> [barrier] -> bind_local (each receiver) -> [barrier] -> map (the sender).  bind_local / map block inside the
 * driver until every device has been added.  Needs NVSwitch + CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED. */
typedef struct kvbm_mc_group kvbm_mc_group;
int kvbm_mc_supported(int device);   /* 1 / 0 (kvbm_last_error says why) */
/* bytes_per_device is rounded up to the multicast granularity (kvbm_mc_group_size).  shareable != 0 allows export_fd. */
int kvbm_mc_group_create(int num_devices, size_t bytes_per_device, int shareable, kvbm_mc_group** out);
int kvbm_mc_group_export_fd(kvbm_mc_group* g, int* fd);   /* POSIX fd; pass it to the peers over SCM_RIGHTS */
int kvbm_mc_group_import_fd(int fd, int num_devices, size_t bytes_per_device, kvbm_mc_group** out);
size_t kvbm_mc_group_size(const kvbm_mc_group* g);
int kvbm_mc_group_add_device(kvbm_mc_group* g, int device);
/* allocate this device's pool (cuMemCreate), map it for the device and bind it at offset 0 of the object */
int kvbm_mc_group_bind_local(kvbm_mc_group* g, int device, void** unicast_ptr);
/* bind the first kvbm_mc_group_size() bytes of memory the ENGINE owns (cuMulticastBindAddr): the range must be backed by the
 * driver's VMM API (cuMemCreate / cuMemMap, e.g. PyTorch expandable segments) and aligned to the multicast granularity */
int kvbm_mc_group_bind_addr(kvbm_mc_group* g, int device, void* ptr, size_t bytes);
/* map the multicast object for `device` (the sender); stores to the returned range reach every bound pool */
int kvbm_mc_group_map(kvbm_mc_group* g, int device, void** multicast_ptr);
void kvbm_mc_group_destroy(kvbm_mc_group* g);

/* accounting for the bench */
uint64_t kvbm_manager_bytes_moved(kvbm_transfer_manager* m);
uint64_t kvbm_manager_h2d_bytes(kvbm_transfer_manager* m);   /* block-table uploads */

#ifdef __cplusplus
}
#endif
#endif /* KVBM_PHYSICAL_H */

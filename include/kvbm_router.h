/*
 * kvbm_router.h -- C ABI of libkvbm_router.so: the KV-aware routing index that turns prefix hits into the block
 * tables the transfer path moves (BASELINE configs[4] "kv_router RadixTree prefix-hit routing").
 *
 * C++ restatement of (relative to /root/reference/lib/kv-router/src):
 *   protocols.rs:20-25      XXH3_SEED = 1337, compute_hash = xxh3_64_with_seed
 *   protocols.rs:74-133     compute_block_hash_for_seq (full blocks only, LoRA name mixed into the seed, eagle window)
 *   protocols.rs:135-172    compute_seq_hash_for_block (rolling parent || block hash)
 *   indexer/radix_tree.rs   RadixTree::{find_matches :165-306, apply_event :315-450, remove_worker, clear_all_blocks}
 *   active_set.rs:9-40      reconcile_active_workers
 * XXH3 itself is the third-party xxHash library (the reference pins the xxhash-rust crate 0.8); this build uses the
 * single-header xxhash.h that ships inside the image's pyarrow wheel.
 */
#ifndef KVBM_ROUTER_H
#define KVBM_ROUTER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  KVR_OK = 0,
  KVR_ERR_PARENT_BLOCK_NOT_FOUND = 1,   /* KvCacheEventError::ParentBlockNotFound */
  KVR_ERR_BLOCK_NOT_FOUND = 2,          /* KvCacheEventError::BlockNotFound */
  KVR_ERR_INVALID_BLOCK_SEQUENCE = 3,   /* KvCacheEventError::InvalidBlockSequence */
  KVR_ERR_ARGUMENT = 4,
};

#define KVR_XXH3_SEED 1337ull

/* compute_hash (protocols.rs:23-25) */
uint64_t kvr_compute_hash(const void* data, size_t len);
/* compute_block_hash_for_seq without multimodal info; returns the number of hashes written (<= cap).
 * lora_name may be NULL / "" (base model).  is_eagle: window = block_size + 1, stride = block_size. */
size_t kvr_compute_block_hash_for_seq(const uint32_t* tokens, size_t n_tokens, uint32_t kv_block_size,
                                      const char* lora_name, int is_eagle, uint64_t* out, size_t cap);
/* compute_seq_hash_for_block */
void kvr_compute_seq_hash_for_block(const uint64_t* block_hashes, size_t n, uint64_t* out);

/* Threading: a kvr_radix_tree is NOT internally synchronised -- like the reference's RadixTree (Rc<RefCell<..>> nodes owned by
 * one indexer task, indexer/kv_indexer.rs) it belongs to one thread at a time.  The dynamo_kv_event_* publisher below IS
 * thread-safe and serialises the events it applies to an attached tree (tests/c/race_check.cpp under ThreadSanitizer). */
typedef struct kvr_radix_tree kvr_radix_tree;

/* expiration_ms < 0: no frequency tracking (RadixTree::new); else new_with_frequency(Some(duration)) */
kvr_radix_tree* kvr_tree_create(int64_t expiration_ms);
void kvr_tree_destroy(kvr_radix_tree* t);

/* RouterEvent{worker_id, KvCacheEvent{event_id, data: Stored{parent_hash?, blocks:[{block_hash, tokens_hash}]}, dp_rank}} */
int kvr_tree_apply_stored(kvr_radix_tree* t, uint64_t worker_id, uint32_t dp_rank, uint64_t event_id, int has_parent,
                          uint64_t parent_hash, size_t n_blocks, const uint64_t* block_hashes,
                          const uint64_t* tokens_hashes);
int kvr_tree_apply_removed(kvr_radix_tree* t, uint64_t worker_id, uint32_t dp_rank, uint64_t event_id, size_t n_blocks,
                           const uint64_t* block_hashes);
/* Cleared event: registers (worker_id, dp_rank) like every apply_event does, then clears every dp rank of worker_id */
int kvr_tree_apply_cleared(kvr_radix_tree* t, uint64_t worker_id, uint32_t dp_rank);
void kvr_tree_remove_worker(kvr_radix_tree* t, uint64_t worker_id);
void kvr_tree_remove_worker_dp_rank(kvr_radix_tree* t, uint64_t worker_id, uint32_t dp_rank);
void kvr_tree_clear_all_blocks(kvr_radix_tree* t, uint64_t worker_id);
/* get_workers: sorted unique worker ids; returns count (writes up to cap) */
size_t kvr_tree_get_workers(kvr_radix_tree* t, uint64_t* out, size_t cap);

/* current_size (radix_tree.rs:567-569): blocks held, summed over (worker, dp_rank) -- what the pruning policy watches */
size_t kvr_tree_current_size(kvr_radix_tree* t);
/* dump_tree_as_events (radix_tree.rs:505-565): the tree as single-block Stored events in breadth-first order, event ids
 * 0..n-1; replaying them into an empty tree rebuilds it (router replica sync, kv_indexer.rs:243).  Returns the number of
 * events; writes the first `cap`. */
typedef struct kvr_dump_event {
  uint64_t worker_id;
  uint32_t dp_rank;
  uint32_t has_parent;
  uint64_t event_id;
  uint64_t parent_hash; /* ExternalSequenceBlockHash of the parent block (has_parent) */
  uint64_t block_hash;  /* ExternalSequenceBlockHash */
  uint64_t tokens_hash; /* LocalBlockHash */
} kvr_dump_event;
size_t kvr_tree_dump_events(kvr_radix_tree* t, kvr_dump_event* out, size_t cap);

/* introspection used by the tests (the reference's tests read `trie.lookup` / `trie.root` directly) */
int64_t kvr_tree_lookup_size(kvr_radix_tree* t, uint64_t worker_id, uint32_t dp_rank); /* -1: worker not in lookup */
size_t kvr_tree_lookup_len(kvr_radix_tree* t);
/* node reached from the root along `path` (tokens hashes); returns 0 and fills counts, or -1 if the path does not exist */
int kvr_tree_node_info(kvr_radix_tree* t, const uint64_t* path, size_t n, size_t* n_workers, size_t* n_children);

/* find_matches(sequence, early_exit) -> OverlapScores.  Arrays are filled up to their capacity; the return value is the
 * number of scored workers; *n_freq receives the number of frequency entries written. */
size_t kvr_tree_find_matches(kvr_radix_tree* t, const uint64_t* sequence, size_t n, int early_exit, uint64_t* worker_ids,
                             uint32_t* dp_ranks, uint32_t* scores, uint64_t* tree_sizes, size_t cap, uint64_t* frequencies,
                             size_t freq_cap, size_t* n_freq);

/* ---- KV-cache event publisher: the reference's C ABI (lib/bindings/c/src/lib.rs:74-78,112-116,177-194,328-391) ----
 * dynamo_llm_init creates the process-wide publisher (component NULL/"" -> "backend"; a second init keeps the first);
 * publish_stored hashes each block's tokens (compute_block_hash_for_seq, optional LoRA name) and publishes ONE Stored event
 * with the blocks whose num_block_tokens equal kv_block_size -- the first partial block ends the list, exactly as
 * kv_event_create_stored_from_parts does; parent_hash NULL = None.  Publishing before init returns ERR (the reference
 * unwraps).  Events are RouterEvent{worker_id, storage_tier: device, event: KvCacheEvent{.., dp_rank: 0}}
 * (lib/kv-router/src/protocols.rs:473-520,695-735).  Dynamo sends them over its runtime's "kv-events" subject; this library
 * owns no runtime, so the host registers the sinks: an in-process RadixTree (the events are applied as the indexer would)
 * and/or a callback that receives each event as the JSON serde_json writes for RouterEvent. */
enum { DYNAMO_LLM_OK = 0, DYNAMO_LLM_ERR = 1 };
uint32_t dynamo_llm_init(const char* namespace_c_str, const char* component_c_str, uint32_t kv_block_size);
uint32_t dynamo_llm_shutdown(void);
uint32_t dynamo_llm_load_publisher_create(void);
uint32_t dynamo_kv_event_publish_stored(uint64_t event_id, const uint32_t* token_ids, const size_t* num_block_tokens,
                                        const uint64_t* block_ids, size_t num_blocks, const uint64_t* parent_hash,
                                        const char* lora_name);
uint32_t dynamo_kv_event_publish_removed(uint64_t event_id, const uint64_t* block_ids, size_t num_blocks);
/* sinks / identity (extensions: in Dynamo the runtime provides them) */
typedef void (*dynamo_kv_event_callback)(const char* router_event_json, size_t len, void* user);
uint32_t dynamo_kv_event_set_worker_id(uint64_t worker_id);
uint32_t dynamo_kv_event_attach_tree(kvr_radix_tree* tree);   /* NULL detaches */
uint32_t dynamo_kv_event_subscribe(dynamo_kv_event_callback cb, void* user);
uint64_t dynamo_kv_event_published_count(void);

#ifdef __cplusplus
}
#endif
#endif

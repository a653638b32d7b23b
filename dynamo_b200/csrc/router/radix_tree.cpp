// libkvbm_router.so -- RadixTree prefix index + XXH3 block hashing (see include/kvbm_router.h for the reference lines).
#define XXH_INLINE_ALL
#include <xxhash.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <deque>
#include <memory>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "../../../include/kvbm_router.h"

namespace {

struct Worker {  // WorkerWithDpRank
  uint64_t id;
  uint32_t dp;
  bool operator==(const Worker& o) const { return id == o.id && dp == o.dp; }
};
struct WorkerHash {
  size_t operator()(const Worker& w) const { return std::hash<uint64_t>()(w.id * 0x9e3779b97f4a7c15ull + w.dp); }
};
using WorkerSet = std::unordered_set<Worker, WorkerHash>;
using Clock = std::chrono::steady_clock;

struct Block;  // RadixBlock
using BlockPtr = std::shared_ptr<Block>;
struct Block {
  std::unordered_set<Block*>* registry = nullptr;   // the owning tree's set of live blocks (see ~kvr_radix_tree)
  ~Block()
  {
    if (registry) registry->erase(this);
  }
  std::unordered_map<uint64_t, BlockPtr> children;  // LocalBlockHash (tokens hash) -> child
  WorkerSet workers;
  bool has_hash = false;
  uint64_t block_hash = 0;  // ExternalSequenceBlockHash
  std::deque<Clock::time_point> recent_uses;
  void drop_worker(const Worker& w)  // radix_tree.rs:75-81
  {
    workers.erase(w);
    if (workers.empty()) children.clear();
  }
};

// active_set.rs:9-40
template <class F>
void reconcile_active_workers(WorkerSet& active, const WorkerSet& next, F on_drop)
{
  if (next.size() == active.size()) return;
  bool subset = next.size() < active.size();
  if (subset)
    for (const Worker& w : next)
      if (!active.count(w)) {
        subset = false;
        break;
      }
  if (subset) {
    for (const Worker& w : active)
      if (!next.count(w)) on_drop(w);
    active = next;
    return;
  }
  for (auto it = active.begin(); it != active.end();) {
    if (next.count(*it)) {
      ++it;
    } else {
      on_drop(*it);
      it = active.erase(it);
    }
  }
}

}  // namespace

struct kvr_radix_tree {
  BlockPtr root = std::make_shared<Block>();
  std::unordered_map<Worker, std::unordered_map<uint64_t, BlockPtr>, WorkerHash> lookup;  // worker -> block_hash -> block
  bool track_frequency = false;
  std::chrono::milliseconds expiration{0};

  std::unordered_set<Block*> live;  // every block this tree created that is still alive

  // Tear down without recursion (chains are as deep as a sequence has blocks; the reference's Drop is iterative too,
  // radix_tree.rs:104-137) and without trusting the graph to be a tree: hand-made hashes can link blocks into cycles, also
  // into cycles no longer reachable from the root.  Every live block is pinned, every edge is cut, then the pins go.
  ~kvr_radix_tree()
  {
    std::vector<BlockPtr> pins;
    pins.reserve(live.size());
    auto pin = [&](const BlockPtr& p) { pins.push_back(p); };
    for (auto& kv : root->children) pin(kv.second);
    for (auto& wl : lookup)
      for (auto& kv : wl.second) pin(kv.second);
    root->children.clear();
    lookup.clear();
    // blocks only other blocks still own: pin them through their owners' child maps before the maps are cleared
    std::vector<Block*> todo(live.begin(), live.end());
    for (Block* b : todo)
      for (auto& kv : b->children) pin(kv.second);
    for (Block* b : todo) b->children.clear();
    for (Block* b : todo) b->registry = nullptr;   // `live` dies with the tree, before or after the last pin
  }

  void remove_or_clear(uint64_t worker_id, bool keep_worker)  // radix_tree.rs:456-480
  {
    std::vector<Worker> keys;
    for (auto& kv : lookup)
      if (kv.first.id == worker_id) keys.push_back(kv.first);
    for (const Worker& w : keys) {
      auto it = lookup.find(w);
      if (it == lookup.end()) continue;
      for (auto& kv : it->second) kv.second->drop_worker(w);
      lookup.erase(it);
      if (keep_worker) lookup[w];
    }
  }
};

extern "C" uint64_t kvr_compute_hash(const void* data, size_t len) { return XXH3_64bits_withSeed(data, len, KVR_XXH3_SEED); }

extern "C" size_t kvr_compute_block_hash_for_seq(const uint32_t* tokens, size_t n_tokens, uint32_t kv_block_size, const char* lora_name,
                                                 int is_eagle, uint64_t* out, size_t cap)
{
  if (kv_block_size == 0 || (!tokens && n_tokens)) return 0;  // protocols.rs:79-81
  uint64_t seed = KVR_XXH3_SEED;
  if (lora_name && *lora_name) seed += XXH3_64bits(lora_name, std::strlen(lora_name));  // wrapping add, protocols.rs:83-86
  const size_t stride = kv_block_size;
  const size_t window = is_eagle ? stride + 1 : stride;
  size_t n = 0;
  for (size_t start = 0; start + window <= n_tokens; start += stride) {  // only full windows are hashed
    const uint64_t h = XXH3_64bits_withSeed(tokens + start, window * sizeof(uint32_t), seed);  // little-endian u32 bytes
    if (n < cap) out[n] = h;
    ++n;
  }
  return n;
}

extern "C" void kvr_compute_seq_hash_for_block(const uint64_t* block_hashes, size_t n, uint64_t* out)
{
  if (n == 0) return;
  out[0] = block_hashes[0];
  for (size_t i = 1; i < n; ++i) {
    const uint64_t combined[2] = {out[i - 1], block_hashes[i]};
    out[i] = kvr_compute_hash(combined, sizeof(combined));
  }
}

extern "C" kvr_radix_tree* kvr_tree_create(int64_t expiration_ms)
{
  auto* t = new kvr_radix_tree();
  if (expiration_ms >= 0) {
    t->track_frequency = true;
    t->expiration = std::chrono::milliseconds(expiration_ms);
  }
  return t;
}
extern "C" void kvr_tree_destroy(kvr_radix_tree* t) { delete t; }

extern "C" int kvr_tree_apply_stored(kvr_radix_tree* t, uint64_t worker_id, uint32_t dp_rank, uint64_t /*event_id*/, int has_parent,
                                     uint64_t parent_hash, size_t n_blocks, const uint64_t* block_hashes, const uint64_t* tokens_hashes)
{
  if (!t || (n_blocks && (!block_hashes || !tokens_hashes))) return KVR_ERR_ARGUMENT;
  const Worker w{worker_id, dp_rank};
  auto& wl = t->lookup[w];  // lookup.entry(worker).or_default()
  BlockPtr current;
  if (has_parent) {
    auto it = wl.find(parent_hash);
    if (it == wl.end()) return KVR_ERR_PARENT_BLOCK_NOT_FOUND;  // radix_tree.rs:331-343
    current = it->second;
  } else {
    current = t->root;
  }
  bool needs_worker_insert = false;
  for (size_t i = 0; i < n_blocks; ++i) {
    if (needs_worker_insert) current->workers.insert(w);
    needs_worker_insert = true;
    BlockPtr child;
    auto it = current->children.find(tokens_hashes[i]);
    if (it != current->children.end()) {
      child = it->second;  // (block_hash mismatch is only logged by the reference)
    } else {
      auto known = wl.find(block_hashes[i]);
      if (known != wl.end()) {
        child = known->second;
      } else {
        child = std::make_shared<Block>();
        child->registry = &t->live;
        t->live.insert(child.get());
        child->has_hash = true;
        child->block_hash = block_hashes[i];
      }
      current->children[tokens_hashes[i]] = child;
    }
    if (child.get() == current.get()) return KVR_ERR_INVALID_BLOCK_SEQUENCE;  // self reference, radix_tree.rs:391-402
    wl[block_hashes[i]] = child;
    current = child;
  }
  if (needs_worker_insert) current->workers.insert(w);
  return KVR_OK;
}

extern "C" int kvr_tree_apply_removed(kvr_radix_tree* t, uint64_t worker_id, uint32_t dp_rank, uint64_t /*event_id*/, size_t n_blocks,
                                      const uint64_t* block_hashes)
{
  if (!t || (n_blocks && !block_hashes)) return KVR_ERR_ARGUMENT;
  const Worker w{worker_id, dp_rank};
  auto& wl = t->lookup[w];
  int err = KVR_OK;
  for (size_t i = 0; i < n_blocks; ++i) {  // apply the whole batch, return the first error (radix_tree.rs:417-443)
    auto it = wl.find(block_hashes[i]);
    if (it == wl.end()) {
      if (err == KVR_OK) err = KVR_ERR_BLOCK_NOT_FOUND;
      continue;
    }
    it->second->drop_worker(w);
    wl.erase(it);
  }
  return err;
}

extern "C" int kvr_tree_apply_cleared(kvr_radix_tree* t, uint64_t worker_id, uint32_t dp_rank)
{
  if (!t) return KVR_ERR_ARGUMENT;
  t->lookup[Worker{worker_id, dp_rank}];  // apply_event: lookup.entry(worker).or_default() (radix_tree.rs:325)
  t->remove_or_clear(worker_id, true);
  return KVR_OK;
}

extern "C" int64_t kvr_tree_lookup_size(kvr_radix_tree* t, uint64_t worker_id, uint32_t dp_rank)
{
  if (!t) return -1;
  auto it = t->lookup.find(Worker{worker_id, dp_rank});
  return it == t->lookup.end() ? -1 : static_cast<int64_t>(it->second.size());
}
extern "C" size_t kvr_tree_lookup_len(kvr_radix_tree* t) { return t ? t->lookup.size() : 0; }
extern "C" int kvr_tree_node_info(kvr_radix_tree* t, const uint64_t* path, size_t n, size_t* n_workers, size_t* n_children)
{
  if (!t) return -1;
  BlockPtr cur = t->root;
  for (size_t i = 0; i < n; ++i) {
    auto it = cur->children.find(path[i]);
    if (it == cur->children.end()) return -1;
    cur = it->second;
  }
  if (n_workers) *n_workers = cur->workers.size();
  if (n_children) *n_children = cur->children.size();
  return 0;
}
extern "C" void kvr_tree_remove_worker(kvr_radix_tree* t, uint64_t worker_id)
{
  if (t) t->remove_or_clear(worker_id, false);
}
extern "C" void kvr_tree_clear_all_blocks(kvr_radix_tree* t, uint64_t worker_id)
{
  if (t) t->remove_or_clear(worker_id, true);
}
extern "C" void kvr_tree_remove_worker_dp_rank(kvr_radix_tree* t, uint64_t worker_id, uint32_t dp_rank)
{
  if (!t) return;
  const Worker w{worker_id, dp_rank};
  auto it = t->lookup.find(w);
  if (it == t->lookup.end()) return;
  for (auto& kv : it->second) kv.second->drop_worker(w);
  t->lookup.erase(it);
}

extern "C" size_t kvr_tree_get_workers(kvr_radix_tree* t, uint64_t* out, size_t cap)
{
  if (!t) return 0;
  std::vector<uint64_t> ids;
  for (auto& kv : t->lookup) ids.push_back(kv.first.id);
  std::sort(ids.begin(), ids.end());
  ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
  for (size_t i = 0; i < ids.size() && i < cap; ++i) out[i] = ids[i];
  return ids.size();
}

// RadixTree::current_size (radix_tree.rs:567-569): blocks held, summed over (worker, dp_rank)
extern "C" size_t kvr_tree_current_size(kvr_radix_tree* t)
{
  if (!t) return 0;
  size_t n = 0;
  for (auto& kv : t->lookup) n += kv.second.size();
  return n;
}

// RadixTree::dump_tree_as_events (radix_tree.rs:505-565): breadth-first, one single-block Stored event per (block, worker),
// event ids 0, 1, 2, ... -- replaying them into an empty tree rebuilds this one (the order among siblings and among a
// block's workers is the hash maps' there as well).
extern "C" size_t kvr_tree_dump_events(kvr_radix_tree* t, kvr_dump_event* out, size_t cap)
{
  if (!t) return 0;
  struct Item {
    BlockPtr block;
    bool has_parent;
    uint64_t parent_hash;
    uint64_t tokens_hash;
  };
  std::deque<Item> queue;
  for (auto& kv : t->root->children) queue.push_back({kv.second, false, 0, kv.first});
  // With real sequence hashes the structure is a tree.  Hand-made hashes can make a block reachable along several edges
  // or even close a cycle (a block is re-used by hash, radix_tree.rs:366-395); every EDGE is walked once so the dump ends.
  struct EdgeHash {
    size_t operator()(const std::pair<const Block*, const Block*>& e) const
    {
      return std::hash<const void*>()(e.first) * 0x9e3779b97f4a7c15ull ^ std::hash<const void*>()(e.second);
    }
  };
  std::unordered_set<std::pair<const Block*, const Block*>, EdgeHash> walked;
  size_t n = 0;
  while (!queue.empty()) {
    Item it = queue.front();
    queue.pop_front();
    for (const Worker& w : it.block->workers) {
      if (out && n < cap) out[n] = kvr_dump_event{w.id, w.dp, it.has_parent ? 1u : 0u, static_cast<uint64_t>(n), it.parent_hash, it.block->block_hash, it.tokens_hash};
      ++n;
    }
    for (auto& kv : it.block->children)
      if (walked.insert({it.block.get(), kv.second.get()}).second) queue.push_back({kv.second, true, it.block->block_hash, kv.first});
  }
  return n;
}

extern "C" size_t kvr_tree_find_matches(kvr_radix_tree* t, const uint64_t* sequence, size_t n, int early_exit, uint64_t* worker_ids,
                                        uint32_t* dp_ranks, uint32_t* scores_out, uint64_t* tree_sizes, size_t cap, uint64_t* frequencies,
                                        size_t freq_cap, size_t* n_freq)
{
  if (n_freq) *n_freq = 0;
  if (!t || n == 0 || !sequence) return 0;
  std::unordered_map<Worker, uint32_t, WorkerHash> scores;
  std::vector<uint64_t> freqs;
  const auto now = Clock::now();
  auto track = [&](Block& b) {  // radix_tree.rs:198-210 / :262-274
    if (!t->track_frequency) return;
    while (!b.recent_uses.empty() && now - b.recent_uses.front() > t->expiration) b.recent_uses.pop_front();
    if (!b.recent_uses.empty()) freqs.push_back(b.recent_uses.size());  // add_frequency omits zeros
    b.recent_uses.push_back(now);
  };
  auto finish = [&]() -> size_t {
    size_t k = 0;
    for (auto& kv : scores) {
      if (k < cap) {
        worker_ids[k] = kv.first.id;
        dp_ranks[k] = kv.first.dp;
        scores_out[k] = kv.second;
        auto it = t->lookup.find(kv.first);
        tree_sizes[k] = it == t->lookup.end() ? 0 : it->second.size();
      }
      ++k;
    }
    size_t f = 0;
    for (; f < freqs.size() && f < freq_cap; ++f) frequencies[f] = freqs[f];
    if (n_freq) *n_freq = freqs.size();
    return k;
  };

  auto fit = t->root->children.find(sequence[0]);
  if (fit == t->root->children.end()) return finish();
  BlockPtr current = fit->second;
  WorkerSet active = current->workers;
  size_t active_count = active.size();
  track(*current);
  if (active.empty()) return finish();
  if (early_exit && active_count == 1) {
    for (const Worker& w : active) scores[w] = 1;
    return finish();
  }
  uint32_t matched_depth = 1;
  for (size_t idx = 1; idx < n; ++idx) {
    auto it = current->children.find(sequence[idx]);
    if (it == current->children.end()) break;
    BlockPtr block = it->second;
    if (block->workers.size() != active_count) {
      reconcile_active_workers(active, block->workers, [&](const Worker& w) { scores[w] = matched_depth; });
      active_count = active.size();
    }
    track(*block);
    if (active_count == 0) break;
    if (early_exit && active_count == 1) {
      matched_depth = static_cast<uint32_t>(idx + 1);
      break;
    }
    current = block;
    matched_depth = static_cast<uint32_t>(idx + 1);
  }
  for (const Worker& w : active) scores[w] = matched_depth;
  return finish();
}

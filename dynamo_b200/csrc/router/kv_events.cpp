// KV-cache event publisher behind the reference's C ABI:
//   /root/reference/lib/bindings/c/src/lib.rs:74-78     DynamoLlmResult { OK = 0, ERR = 1 }
//   :112-175  dynamo_llm_init(namespace, component, kv_block_size)   (component NULL -> "backend")
//   :177-194  dynamo_llm_shutdown, dynamo_llm_load_publisher_create
//   :228-246  one stored block: tokens_hash = compute_block_hash_for_seq(tokens, kv_block_size, lora)[0]
//   :249-301  kv_event_create_stored_from_parts: blocks are taken while num_block_tokens[i] == kv_block_size; the first
//             partial block ends the event (warned at most 3 times), parent_hash NULL -> None, dp_rank 0
//   :303-319  removed event, :328-391 the two publish entry points
// What is published is a RouterEvent (lib/kv-router/src/protocols.rs:695-735) -- {worker_id, storage_tier, event} -- in
// the JSON serde_json writes for it (snake_case enum tags, bare u64 hashes, mm_extra_info null).  In Dynamo the events
// travel over the DistributedRuntime's "kv-events" subject to the router's indexer; that runtime is Dynamo's and stays
// Dynamo's.  Here the publisher hands every event to the sinks the host registered: an in-process RadixTree (the
// indexer of this library) and/or a callback that receives the JSON (to forward over whatever transport the host has).
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/kvbm_router.h"

namespace {

struct Publisher {
  std::string ns, component;
  uint32_t kv_block_size = 0;
};

std::mutex g_mu;
Publisher* g_pub = nullptr;
uint64_t g_worker_id = 0;
kvr_radix_tree* g_tree = nullptr;
dynamo_kv_event_callback g_cb = nullptr;
void* g_cb_user = nullptr;
std::atomic<uint32_t> g_warns{0};
uint64_t g_published = 0;

void append_u64(std::string& s, uint64_t v) { s += std::to_string(v); }

std::string router_event_json(uint64_t worker_id, uint64_t event_id, const std::string& data)
{
  std::string s = "{\"worker_id\":";
  append_u64(s, worker_id);
  s += ",\"storage_tier\":\"device\",\"event\":{\"event_id\":";
  append_u64(s, event_id);
  s += ",\"data\":" + data + ",\"dp_rank\":0}}";
  return s;
}

// deliver to the sinks; the tree result is the publish result (an event the indexer rejects is an error to the caller)
uint32_t deliver(const std::string& json, int tree_rc)
{
  ++g_published;
  if (g_cb) g_cb(json.data(), json.size(), g_cb_user);
  return tree_rc == KVR_OK ? DYNAMO_LLM_OK : DYNAMO_LLM_ERR;
}

}  // namespace

extern "C" uint32_t dynamo_llm_init(const char* namespace_c_str, const char* component_c_str, uint32_t kv_block_size)
{
  if (!namespace_c_str) return DYNAMO_LLM_ERR;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_pub) return DYNAMO_LLM_OK;  // KV_PUB.get_or_try_init: the first publisher stays
  if (kv_block_size == 0) return DYNAMO_LLM_ERR;
  g_pub = new Publisher{namespace_c_str, component_c_str && *component_c_str ? component_c_str : "backend", kv_block_size};
  return DYNAMO_LLM_OK;
}

extern "C" uint32_t dynamo_llm_shutdown(void)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_pub) return DYNAMO_LLM_ERR;  // "Runtime not initialized"
  delete g_pub;
  g_pub = nullptr;
  g_tree = nullptr;
  g_cb = nullptr;
  g_cb_user = nullptr;
  g_warns.store(0);
  return DYNAMO_LLM_OK;
}

extern "C" uint32_t dynamo_llm_load_publisher_create(void) { return DYNAMO_LLM_OK; }

extern "C" uint32_t dynamo_kv_event_set_worker_id(uint64_t worker_id)
{
  std::lock_guard<std::mutex> lk(g_mu);
  g_worker_id = worker_id;
  return DYNAMO_LLM_OK;
}

extern "C" uint32_t dynamo_kv_event_attach_tree(kvr_radix_tree* tree)
{
  std::lock_guard<std::mutex> lk(g_mu);
  g_tree = tree;
  return DYNAMO_LLM_OK;
}

extern "C" uint32_t dynamo_kv_event_subscribe(dynamo_kv_event_callback cb, void* user)
{
  std::lock_guard<std::mutex> lk(g_mu);
  g_cb = cb;
  g_cb_user = user;
  return DYNAMO_LLM_OK;
}

extern "C" uint64_t dynamo_kv_event_published_count(void)
{
  std::lock_guard<std::mutex> lk(g_mu);
  return g_published;
}

extern "C" uint32_t dynamo_kv_event_publish_stored(uint64_t event_id, const uint32_t* token_ids, const size_t* num_block_tokens,
                                                   const uint64_t* block_ids, size_t num_blocks, const uint64_t* parent_hash,
                                                   const char* lora_name)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_pub) return DYNAMO_LLM_ERR;  // the reference unwraps KV_PUB here: publishing before init is a caller bug
  if (num_blocks && (!token_ids || !num_block_tokens || !block_ids)) return DYNAMO_LLM_ERR;
  std::vector<uint64_t> block_hashes, tokens_hashes;
  size_t token_offset = 0;
  for (size_t i = 0; i < num_blocks; ++i) {
    const size_t num_toks = num_block_tokens[i];
    if (num_toks != g_pub->kv_block_size) {
      if (g_warns.load() < 3) {
        g_warns.fetch_add(1);
        std::fprintf(stderr, "Block not published. Block size must be %u tokens to be published. Block size is: %zu\n",
                     g_pub->kv_block_size, num_toks);
      }
      break;
    }
    uint64_t th = 0;
    if (kvr_compute_block_hash_for_seq(token_ids + token_offset, num_toks, g_pub->kv_block_size, lora_name, 0, &th, 1) != 1)
      return DYNAMO_LLM_ERR;
    token_offset += num_toks;
    block_hashes.push_back(block_ids[i]);
    tokens_hashes.push_back(th);
  }
  std::string data = "{\"stored\":{\"parent_hash\":";
  if (parent_hash)
    append_u64(data, *parent_hash);
  else
    data += "null";
  data += ",\"blocks\":[";
  for (size_t i = 0; i < block_hashes.size(); ++i) {
    if (i) data += ",";
    data += "{\"block_hash\":";
    append_u64(data, block_hashes[i]);
    data += ",\"tokens_hash\":";
    append_u64(data, tokens_hashes[i]);
    data += ",\"mm_extra_info\":null}";
  }
  data += "]}}";
  int rc = KVR_OK;
  if (g_tree)
    rc = kvr_tree_apply_stored(g_tree, g_worker_id, 0, event_id, parent_hash != nullptr, parent_hash ? *parent_hash : 0,
                               block_hashes.size(), block_hashes.data(), tokens_hashes.data());
  return deliver(router_event_json(g_worker_id, event_id, data), rc);
}

extern "C" uint32_t dynamo_kv_event_publish_removed(uint64_t event_id, const uint64_t* block_ids, size_t num_blocks)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_pub) return DYNAMO_LLM_ERR;
  if (num_blocks && !block_ids) return DYNAMO_LLM_ERR;
  std::string data = "{\"removed\":{\"block_hashes\":[";
  for (size_t i = 0; i < num_blocks; ++i) {
    if (i) data += ",";
    append_u64(data, block_ids[i]);
  }
  data += "]}}";
  int rc = KVR_OK;
  if (g_tree) rc = kvr_tree_apply_removed(g_tree, g_worker_id, 0, event_id, num_blocks, block_ids);
  return deliver(router_event_json(g_worker_id, event_id, data), rc);
}

// Race check of the host libraries (race detection).  Built from the library SOURCES with -fsanitize=thread by
// tests/test_host_logic.py and run without a GPU: host-only TransferManager (Memcpy strategy), the KV event publisher and
// its RadixTree sink, driven from many threads the way tokio workers / a TRT-LLM executor thread would.  Any data race makes
// ThreadSanitizer print a report and exit non-zero (halt_on_error).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "kvbm_physical.h"
#include "kvbm_router.h"

#define CHECK(c)                                                                  \
  do {                                                                            \
    if (!(c)) {                                                                   \
      std::fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #c, kvbm_last_error()); \
      std::exit(2);                                                               \
    }                                                                             \
  } while (0)

static std::atomic<unsigned long> g_json_bytes{0};
static void on_event(const char* json, size_t len, void*) { g_json_bytes.fetch_add(len + (json[0] == '{'), std::memory_order_relaxed); }

int main()
{
  // ---------------- TransferManager, host only ----------------
  kvbm_layout_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.num_blocks = 64;
  cfg.num_layers = 2;
  cfg.outer_dim = 2;
  cfg.page_size = 16;
  cfg.inner_dim = 64;
  cfg.alignment = 1;
  cfg.dtype_width_bytes = 2;
  const size_t bytes = kvbm_layout_required_bytes(&cfg);
  std::vector<unsigned char> src(bytes), dst(bytes, 0);
  for (size_t i = 0; i < bytes; ++i) src[i] = static_cast<unsigned char>(i * 131u >> 3);
  kvbm_transfer_manager* m = nullptr;
  CHECK(kvbm_manager_create(-1, 3, &m) == KVBM_OK);
  kvbm_layout_handle hs = 0, hd = 0;
  CHECK(kvbm_manager_register_fully_contiguous(m, &cfg, src.data(), bytes, KVBM_STORAGE_SYSTEM, 0, &hs) == KVBM_OK);
  CHECK(kvbm_manager_register_fully_contiguous(m, &cfg, dst.data(), bytes, KVBM_STORAGE_PINNED, 0, &hd) == KVBM_OK);
  const int T = 8, ROUNDS = 200, PER = 64 / T;
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t)
    th.emplace_back([&, t] {
      std::vector<unsigned char> scratch(bytes);
      for (int r = 0; r < ROUNDS; ++r) {
        size_t sid[PER], did[PER];
        for (int i = 0; i < PER; ++i) {
          sid[i] = static_cast<size_t>((t * 7 + r * 13 + i * 5) % 64);
          did[i] = static_cast<size_t>(t * PER + (i + r) % PER);     // thread t only writes its own destination blocks
        }
        kvbm_notification note = 1;
        CHECK(kvbm_manager_execute_transfer(m, hs, sid, hd, did, PER, nullptr, &note) == KVBM_OK);
        CHECK(kvbm_notification_is_complete(m, note) == 1);
        kvbm_layout_handle h = 0;
        CHECK(kvbm_manager_register_fully_contiguous(m, &cfg, scratch.data(), bytes, KVBM_STORAGE_SYSTEM, 0, &h) == KVBM_OK);
        size_t a = 0, s = 0;
        CHECK(kvbm_layout_memory_region(m, h, 1, 1, 1, &a, &s) == KVBM_OK && s == 16 * 64 * 2);
        kvbm_transfer_plan plan;
        CHECK(kvbm_manager_select_strategy(m, hs, h, &plan) == KVBM_OK);
        CHECK(kvbm_manager_unregister(m, h) == KVBM_OK);
        (void)kvbm_manager_bytes_moved(m);
      }
    });
  for (auto& x : th) x.join();
  th.clear();
  const size_t per_block = kvbm_layout_bytes_per_block(&cfg);
  for (int t = 0; t < T; ++t)                                       // the last round's sources are where they belong
    for (int i = 0; i < PER; ++i) {
      const int r = ROUNDS - 1;
      const size_t s = static_cast<size_t>((t * 7 + r * 13 + i * 5) % 64), d = static_cast<size_t>(t * PER + (i + r) % PER);
      CHECK(std::memcmp(dst.data() + d * per_block, src.data() + s * per_block, per_block) == 0);
    }
  kvbm_manager_destroy(m);

  // ---------------- KV event publisher -> RadixTree + subscriber ----------------
  kvr_radix_tree* tree = kvr_tree_create(-1);
  CHECK(tree != nullptr);
  (void)dynamo_llm_shutdown();
  CHECK(dynamo_llm_init("ns", "backend", 4) == 0);
  CHECK(dynamo_kv_event_set_worker_id(11) == 0);
  CHECK(dynamo_kv_event_attach_tree(tree) == 0);
  CHECK(dynamo_kv_event_subscribe(on_event, nullptr) == 0);
  const unsigned long before = dynamo_kv_event_published_count();
  for (int t = 0; t < T; ++t)
    th.emplace_back([t] {
      for (int i = 0; i < 300; ++i) {
        const uint32_t base = static_cast<uint32_t>((t * 300 + i) * 16);
        uint32_t toks[8];
        for (int k = 0; k < 8; ++k) toks[k] = base + k;
        const size_t nbt[2] = {4, 4};
        const uint64_t ids[2] = {base + 1u, base + 2u};
        CHECK(dynamo_kv_event_publish_stored(static_cast<uint64_t>(t) * 100000 + i, toks, nbt, ids, 2, nullptr, nullptr) == 0);
        if (i % 2) CHECK(dynamo_kv_event_publish_removed(static_cast<uint64_t>(t) * 100000 + 50000 + i, ids + 1, 1) == 0);
        (void)dynamo_kv_event_published_count();
      }
    });
  for (auto& x : th) x.join();
  CHECK(dynamo_kv_event_published_count() - before == static_cast<unsigned long>(T) * (300 + 150));
  CHECK(kvr_tree_current_size(tree) == static_cast<size_t>(T) * (600 - 150));
  CHECK(g_json_bytes.load() > 0);
  CHECK(dynamo_llm_shutdown() == 0);
  kvr_tree_destroy(tree);
  std::puts("race check ok");
  return 0;
}

// Robustness of the three metadata importers against corrupted input (they parse bytes that arrive from other processes):
// kvbm_manager_import_metadata (KVBMLAY1 blob), kvbm_manager_import_serialized_layout (bincode 2 SerializedLayout) and
// kvbm_manager_import_descriptor_json (LayoutDescriptor JSON).  Valid exports are mutated (bit flips, byte overwrites,
// truncation, splices, length-field inflation) and imported into fresh host-only managers.  Built with
// -fsanitize=address,undefined by tests/test_host_logic.py: any out-of-bounds read, overflow or leak fails the run; an
// import may only succeed or return an error code.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "kvbm_physical.h"

#define CHECK(c)                                                                                   \
  do {                                                                                             \
    if (!(c)) {                                                                                    \
      std::fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #c, kvbm_last_error()); \
      std::exit(2);                                                                                \
    }                                                                                              \
  } while (0)

static std::vector<unsigned char> mutate(const std::vector<unsigned char>& in, std::mt19937_64& rng)
{
  std::vector<unsigned char> b = in;
  const int kind = static_cast<int>(rng() % 7);
  auto pos = [&]() { return b.empty() ? 0 : static_cast<size_t>(rng() % b.size()); };
  switch (kind) {
    case 0:  // bit flips
      for (int k = 0, n = 1 + static_cast<int>(rng() % 8); k < n && !b.empty(); ++k) b[pos()] ^= static_cast<unsigned char>(1u << (rng() % 8));
      break;
    case 1:  // random bytes
      for (int k = 0, n = 1 + static_cast<int>(rng() % 16); k < n && !b.empty(); ++k) b[pos()] = static_cast<unsigned char>(rng());
      break;
    case 2:  // truncate
      b.resize(b.empty() ? 0 : rng() % b.size());
      break;
    case 3: {  // splice a run of 0xff (huge varints / lengths)
      const size_t p = pos(), n = 1 + rng() % 9;
      for (size_t k = 0; k < n && p + k < b.size(); ++k) b[p + k] = 0xff;
      break;
    }
    case 4: {  // duplicate a slice
      if (b.size() > 4) {
        const size_t p = pos(), n = 1 + rng() % std::min<size_t>(64, b.size() - p);
        b.insert(b.begin() + static_cast<long>(p), in.begin() + static_cast<long>(p), in.begin() + static_cast<long>(p + n));
      }
      break;
    }
    case 5: {  // replace a run of decimal digits (JSON numbers) by an extreme value
      size_t p = pos();
      while (p < b.size() && !(b[p] >= '0' && b[p] <= '9')) ++p;
      size_t e = p;
      while (e < b.size() && b[e] >= '0' && b[e] <= '9') ++e;
      if (p < b.size()) {
        static const char* vals[] = {"18446744073709551615", "18446744073709551616", "-1", "0", "4294967296", "1e308", "9223372036854775807"};
        const char* v = vals[rng() % 7];
        b.erase(b.begin() + static_cast<long>(p), b.begin() + static_cast<long>(e));
        b.insert(b.begin() + static_cast<long>(p), v, v + std::strlen(v));
      }
      break;
    }
    default:  // append garbage
      for (int k = 0, n = 1 + static_cast<int>(rng() % 32); k < n; ++k) b.push_back(static_cast<unsigned char>(rng()));
  }
  return b;
}

int main(int argc, char** argv)
{
  const int iters = argc > 1 ? std::atoi(argv[1]) : 4000;
  kvbm_layout_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.num_blocks = 8;
  cfg.num_layers = 3;
  cfg.outer_dim = 2;
  cfg.page_size = 16;
  cfg.inner_dim = 64;
  cfg.alignment = 1;
  cfg.dtype_width_bytes = 2;
  cfg.num_heads = 4;
  const size_t bytes = kvbm_layout_required_bytes(&cfg);
  std::vector<unsigned char> pool(bytes), layers(bytes);
  kvbm_transfer_manager* m = nullptr;
  CHECK(kvbm_manager_create(-1, 77, &m) == KVBM_OK);
  kvbm_layout_handle fc = 0, lw = 0;
  CHECK(kvbm_manager_register_fully_contiguous(m, &cfg, pool.data(), bytes, KVBM_STORAGE_SYSTEM, 0, &fc) == KVBM_OK);
  CHECK(kvbm_manager_set_kv_block_layout(m, fc, KVBM_KV_UNIVERSAL_TP) == KVBM_OK);
  void* bases[3];
  size_t sizes[3];
  for (int l = 0; l < 3; ++l) {
    bases[l] = layers.data() + l * (bytes / 3);
    sizes[l] = bytes / 3;
  }
  CHECK(kvbm_manager_register_layer_separate(m, &cfg, bases, sizes, KVBM_BLOCK_IS_SECOND_DIM, KVBM_STORAGE_PINNED, 0, &lw) == KVBM_OK);

  std::vector<std::vector<unsigned char>> seeds[3];
  for (kvbm_layout_handle h : {fc, lw}) {
    size_t n = 0;
    CHECK(kvbm_manager_export_metadata(m, h, nullptr, 0, &n) == KVBM_OK && n > 0);
    std::vector<unsigned char> blob(n);
    CHECK(kvbm_manager_export_metadata(m, h, blob.data(), n, &n) == KVBM_OK);
    seeds[0].push_back(blob);
    char json[8192];
    size_t jl = 0;
    CHECK(kvbm_layout_descriptor_json(m, h, json, sizeof json, &jl) == KVBM_OK);
    seeds[2].push_back(std::vector<unsigned char>(json, json + jl));
  }
  {
    size_t n = 0;
    CHECK(kvbm_manager_export_serialized_layout(m, nullptr, 0, &n) == KVBM_OK && n > 0);
    std::vector<unsigned char> blob(n);
    CHECK(kvbm_manager_export_serialized_layout(m, blob.data(), n, &n) == KVBM_OK);
    seeds[1].push_back(blob);
  }

  std::mt19937_64 rng(20260921);
  unsigned long accepted[3] = {0, 0, 0}, rejected[3] = {0, 0, 0};
  for (int it = 0; it < iters; ++it) {
    kvbm_transfer_manager* peer = nullptr;
    CHECK(kvbm_manager_create(-1, 1000 + static_cast<uint64_t>(it), &peer) == KVBM_OK);
    for (int which = 0; which < 3; ++which) {
      const auto& seed = seeds[which][rng() % seeds[which].size()];
      std::vector<unsigned char> b = it == 0 ? seed : mutate(seed, rng);   // iteration 0: the unmodified exports must import
      if (it % 7 == 3) b = mutate(b, rng);
      int rc;
      kvbm_layout_handle out[8];
      size_t n = 0;
      if (which == 0)
        rc = kvbm_manager_import_metadata(peer, b.data(), b.size(), out);
      else if (which == 1)
        rc = kvbm_manager_import_serialized_layout(peer, b.data(), b.size(), out, 8, &n);
      else
        rc = kvbm_manager_import_descriptor_json(peer, reinterpret_cast<const char*>(b.data()), b.size(), out);
      if (it == 0) CHECK(rc == KVBM_OK);
      if (rc == KVBM_OK) {
        ++accepted[which];
        size_t a = 0, s = 0;                                      // an accepted layout answers geometry queries without touching memory
        (void)kvbm_layout_memory_region(peer, out[0], 0, 0, 0, &a, &s);
        (void)kvbm_layout_is_fully_contiguous(peer, out[0]);
        (void)kvbm_manager_kv_block_layout(peer, out[0]);
      } else {
        ++rejected[which];
        CHECK(kvbm_last_error() != nullptr && kvbm_last_error()[0] != 0);
      }
    }
    kvbm_manager_destroy(peer);
  }
  kvbm_manager_destroy(m);
  for (int w = 0; w < 3; ++w) CHECK(rejected[w] > 0 && accepted[w] > 0);
  std::printf("fuzz ok: blob %lu/%lu, serialized %lu/%lu, json %lu/%lu (accepted/rejected)\n", accepted[0], rejected[0], accepted[1], rejected[1],
              accepted[2], rejected[2]);
  return 0;
}

// Memory-safety fuzz of libkvbm_router's RadixTree under ASan/UBSan: random Stored / Removed / Cleared events, worker removal,
// queries and dumps over a SMALL hash space, so blocks are shared, re-parented, emptied and (with hand-made hashes) even
// linked into cycles -- the shapes that stress the shared-ownership graph (radix_tree.rs:34-137).  Error returns are fine;
// any use-after-free, overflow or leak fails the run.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "kvbm_router.h"

int main(int argc, char** argv)
{
  const int trees = argc > 1 ? std::atoi(argv[1]) : 300;
  // mode "tree" (default): block hash = hash of the whole prefix, the structure is a tree (leak check meaningful);
  // mode "graph": block hash tied to the tokens hash only -> shared blocks, DAGs and cycles (run with detect_leaks=0: a
  // detached cycle of reference-counted blocks is unreachable by construction, in the reference's Rc graph as well)
  const bool graph = argc > 2 && std::string(argv[2]) == "graph";
  std::mt19937_64 rng(7);
  unsigned long ok = 0, err = 0, scored = 0, dumped = 0;
  for (int t = 0; t < trees; ++t) {
    kvr_radix_tree* tree = kvr_tree_create(t % 3 == 0 ? 50 : -1);   // every third tree tracks frequencies
    if (!tree) return 2;
    const int steps = 50 + static_cast<int>(rng() % 200);
    for (int s = 0; s < steps; ++s) {
      const uint64_t w = rng() % 4;
      const uint32_t dp = static_cast<uint32_t>(rng() % 2);
      const int kind = static_cast<int>(rng() % 12);
      if (kind < 6) {
        const size_t n = 1 + rng() % 5;
        std::vector<uint64_t> bh(n), th(n);
        for (size_t i = 0; i < n; ++i) th[i] = 1 + rng() % (graph ? 9 : 4);
        if (graph) {
          for (size_t i = 0; i < n; ++i) bh[i] = th[i] * 100;     // collisions across depths
        } else {
          kvr_compute_seq_hash_for_block(th.data(), n, bh.data()); // real sequence hashes: one block per distinct prefix
        }
        const int has_parent = graph ? static_cast<int>(rng() % 3 == 0) : 0;
        const uint64_t parent = (1 + rng() % 9) * 100;
        (kvr_tree_apply_stored(tree, w, dp, static_cast<uint64_t>(s), has_parent, parent, n, bh.data(), th.data()) == 0 ? ok : err)++;
      } else if (kind < 8) {
        const size_t n = 1 + rng() % 3;
        std::vector<uint64_t> bh(n);
        for (auto& x : bh) x = (1 + rng() % 9) * 100;
        if (!graph) {                                             // evict a real block: the hash of a short random prefix
          std::vector<uint64_t> th(1 + rng() % 4), sh(4);
          for (auto& x : th) x = 1 + rng() % 4;
          kvr_compute_seq_hash_for_block(th.data(), th.size(), sh.data());
          for (size_t i = 0; i < n; ++i) bh[i] = sh[th.size() - 1 - (i < th.size() ? i : 0)];
        }
        (kvr_tree_apply_removed(tree, w, dp, static_cast<uint64_t>(s), n, bh.data()) == 0 ? ok : err)++;
      } else if (kind == 8) {
        kvr_tree_apply_cleared(tree, w, dp);
      } else if (kind == 9) {
        if (rng() % 2) kvr_tree_remove_worker(tree, w); else kvr_tree_remove_worker_dp_rank(tree, w, dp);
      } else if (kind == 10) {
        kvr_tree_clear_all_blocks(tree, w);
      } else {
        std::vector<kvr_dump_event> ev(4096);
        dumped += kvr_tree_dump_events(tree, ev.data(), ev.size());
      }
      uint64_t q[6], wid[16], sizes[16], freq[8];
      uint32_t dps[16], sc[16];
      const size_t qn = 1 + rng() % 6;
      for (size_t i = 0; i < qn; ++i) q[i] = 1 + rng() % (graph ? 9 : 4);
      size_t nf = 0;
      scored += kvr_tree_find_matches(tree, q, qn, static_cast<int>(rng() % 2), wid, dps, sc, sizes, 16, freq, 8, &nf);
      (void)kvr_tree_current_size(tree);
      size_t nw = 0, nc = 0;
      (void)kvr_tree_node_info(tree, q, qn, &nw, &nc);
      uint64_t ws[8];
      (void)kvr_tree_get_workers(tree, ws, 8);
    }
    kvr_tree_destroy(tree);
  }
  if (!ok || !err || !scored || !dumped) return 3;
  std::printf("router fuzz ok: %lu events applied, %lu rejected, %lu scores, %lu dumped events\n", ok, err, scored, dumped);
  return 0;
}

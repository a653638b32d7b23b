/* A plain C99 caller of the host library -- what a cgo / JNI / Rust FFI binding would do, without Python in the loop.
 * Host-only manager (no GPU needed): register two pools, move blocks with the Memcpy strategy, hand the layout
 * metadata to a second manager in the reference's SerializedLayout format, check errors come back as codes + text.
 * Built and run by tests/test_host_logic.py::test_c_program_drives_the_host_abi. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kvbm_physical.h"

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      fprintf(stderr, "FAILED %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #cond, kvbm_last_error()); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(void)
{
  kvbm_layout_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.num_blocks = 8;
  cfg.num_layers = 3;
  cfg.outer_dim = 2;
  cfg.page_size = 16;
  cfg.inner_dim = 64;
  cfg.alignment = 1;
  cfg.dtype_width_bytes = 2;
  CHECK(kvbm_layout_config_validate(&cfg) == KVBM_OK);
  const size_t bytes = kvbm_layout_required_bytes(&cfg);
  const size_t per_block = kvbm_layout_bytes_per_block(&cfg);
  CHECK(bytes == 8 * per_block && per_block == 3 * 2 * 16 * 64 * 2);

  unsigned char* src = malloc(bytes);
  unsigned char* dst = calloc(bytes, 1);
  CHECK(src && dst);
  for (size_t i = 0; i < bytes; ++i) src[i] = (unsigned char)((i * 31 + 7) % 251);

  kvbm_transfer_manager* m = NULL;
  CHECK(kvbm_manager_create(-1, 5, &m) == KVBM_OK);
  kvbm_layout_handle hs = 0, hd = 0;
  CHECK(kvbm_manager_register_fully_contiguous(m, &cfg, src, bytes, KVBM_STORAGE_SYSTEM, 0, &hs) == KVBM_OK);
  CHECK(kvbm_manager_register_fully_contiguous(m, &cfg, dst, bytes, KVBM_STORAGE_SYSTEM, 0, &hd) == KVBM_OK);

  const size_t sids[3] = {6, 0, 3}, dids[3] = {1, 7, 2};
  kvbm_notification note = 99;
  CHECK(kvbm_manager_execute_transfer(m, hs, sids, hd, dids, 3, NULL, &note) == KVBM_OK);
  CHECK(note == 0 && kvbm_notification_is_complete(m, note) == 1);   /* Memcpy strategy: already complete */
  for (int i = 0; i < 3; ++i) CHECK(memcmp(dst + dids[i] * per_block, src + sids[i] * per_block, per_block) == 0);
  for (size_t i = 0; i < per_block; ++i) CHECK(dst[0 * per_block + i] == 0);   /* untouched block stays untouched */

  /* errors are codes + a message, never an abort */
  const size_t dup[3] = {1, 1, 2};
  CHECK(kvbm_manager_execute_transfer(m, hs, sids, hd, dup, 3, NULL, &note) == KVBM_ERR_DUPLICATE_DST);
  CHECK(strstr(kvbm_last_error(), "not unique") != NULL);
  const size_t oob[3] = {1, 8, 2};
  CHECK(kvbm_manager_execute_transfer(m, hs, sids, hd, oob, 3, NULL, &note) == KVBM_ERR_RANGE);

  /* the handshake in the reference's SerializedLayout format */
  size_t len = 0, n = 0;
  CHECK(kvbm_manager_export_serialized_layout(m, NULL, 0, &len) == KVBM_OK && len > 0);
  unsigned char* blob = malloc(len);
  CHECK(kvbm_manager_export_serialized_layout(m, blob, len, &len) == KVBM_OK);
  kvbm_transfer_manager* peer = NULL;
  CHECK(kvbm_manager_create(-1, 6, &peer) == KVBM_OK);
  kvbm_layout_handle imported[4];
  CHECK(kvbm_manager_import_serialized_layout(peer, blob, len, imported, 4, &n) == KVBM_OK && n == 2);
  uintptr_t a = 0, b = 0;
  size_t sz = 0;
  CHECK(kvbm_layout_memory_region(m, hs, 6, 2, 1, &a, &sz) == KVBM_OK);
  CHECK(kvbm_layout_memory_region(peer, imported[0], 6, 2, 1, &b, NULL) == KVBM_OK);
  CHECK(a == b && sz == 16 * 64 * 2);
  CHECK(kvbm_manager_import_serialized_layout(peer, blob, len, imported, 4, &n) != KVBM_OK);
  CHECK(strstr(kvbm_last_error(), "already loaded") != NULL);

  char json[2048];
  size_t jl = 0;
  CHECK(kvbm_layout_descriptor_json(m, hd, json, sizeof json, &jl) == KVBM_OK && jl > 0 && jl < sizeof json);
  json[jl] = 0;
  CHECK(strstr(json, "\"FullyContiguous\"") != NULL && strstr(json, "\"num_blocks\":8") != NULL);

  /* a CUDA strategy on a host-only manager fails loudly: no CPU fallback */
  kvbm_layout_handle hdev = 0, hpin = 0;
  CHECK(kvbm_manager_register_fully_contiguous(m, &cfg, dst, bytes, KVBM_STORAGE_DEVICE, 0, &hdev) == KVBM_OK);
  CHECK(kvbm_manager_register_fully_contiguous(m, &cfg, src, bytes, KVBM_STORAGE_PINNED, 0, &hpin) == KVBM_OK);
  CHECK(kvbm_manager_execute_transfer(m, hpin, sids, hdev, dids, 3, NULL, &note) == KVBM_ERR_CUDA);   /* Pinned -> Device = CudaAsyncH2D */
  CHECK(strstr(kvbm_last_error(), "no CPU fallback") != NULL);
  /* System <-> Device has no direct strategy, exactly as the reference's table (strategy.rs:138-210) */
  CHECK(kvbm_manager_execute_transfer(m, hs, sids, hdev, dids, 3, NULL, &note) != KVBM_OK);

  /* strategy and transform selection are pure functions of plain ints (strategy.rs:78-108,245-281; executor/mod.rs:46-100) */
  {
    kvbm_transfer_capabilities caps = {0, 1};
    kvbm_transfer_plan plan;
    CHECK(kvbm_select_strategy(KVBM_STORAGE_DEVICE, 1, KVBM_STORAGE_DEVICE, 0, &caps, &plan) == KVBM_OK && !plan.two_hop && plan.first == KVBM_STRATEGY_NIXL_WRITE);
    CHECK(kvbm_select_strategy(KVBM_STORAGE_DEVICE, 0, KVBM_STORAGE_DEVICE, 1, &caps, &plan) == KVBM_OK && plan.first == KVBM_STRATEGY_NIXL_READ_FLIPPED);
    CHECK(kvbm_select_strategy(KVBM_STORAGE_DEVICE, 0, KVBM_STORAGE_DEVICE, 0, &caps, &plan) == KVBM_ERR_UNSUPPORTED);
    CHECK(strstr(kvbm_last_error(), "Both src and dst are remote") != NULL);
    caps.allow_gpu_rdma = 0;
    CHECK(kvbm_select_direct_strategy_remote(KVBM_STORAGE_DEVICE, KVBM_STORAGE_SYSTEM, 1, &caps, &plan) == KVBM_OK && plan.two_hop &&
          plan.first == KVBM_STRATEGY_CUDA_ASYNC_D2H && plan.bounce_location == KVBM_STORAGE_PINNED && plan.second == KVBM_STRATEGY_NIXL_WRITE);
    CHECK(kvbm_manager_select_strategy(m, hs, hd, &plan) == KVBM_OK && plan.first == KVBM_STRATEGY_MEMCPY);
    CHECK(kvbm_select_transform_kernel(KVBM_KV_OPERATIONAL_NHD, KVBM_KV_UNIVERSAL_TP) == KVBM_TRANSFORM_BLOCK_TO_UNIVERSAL);
    CHECK(kvbm_select_transform_kernel(KVBM_KV_UNIVERSAL_TP, KVBM_KV_OPERATIONAL_HND) == KVBM_TRANSFORM_UNIVERSAL_TO_BLOCK);
    CHECK(kvbm_select_transform_kernel(KVBM_KV_OPERATIONAL_NHD, KVBM_KV_OPERATIONAL_HND) == KVBM_TRANSFORM_OPERATIONAL_TRANSPOSE);
    CHECK(kvbm_select_transform_kernel(KVBM_KV_UNKNOWN, KVBM_KV_UNKNOWN) == KVBM_TRANSFORM_NONE);
    CHECK(kvbm_select_transform_kernel(KVBM_KV_UNKNOWN, KVBM_KV_UNIVERSAL_TP) == KVBM_TRANSFORM_UNSUPPORTED);
    CHECK(kvbm_kv_layout_requires_transform(KVBM_KV_OPERATIONAL_NHD, KVBM_KV_OPERATIONAL_NHD) == 0);
    CHECK(kvbm_manager_set_kv_block_layout(m, hs, KVBM_KV_OPERATIONAL_NHD) == KVBM_ERR_CONFIG);   /* this config has no num_heads */
    CHECK(strstr(kvbm_last_error(), "num_heads_required_for_kv_block_layout") != NULL);
    CHECK(kvbm_manager_kv_block_layout(m, hs) == KVBM_KV_UNKNOWN);
  }

  kvbm_manager_destroy(peer);
  kvbm_manager_destroy(m);
  free(blob);
  free(src);
  free(dst);
  puts("host ABI ok");
  return 0;
}

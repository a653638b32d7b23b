/*
 * kvbm_oracle.h -- CPU restatement of the reference's KV-block transfer path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dynamo_b200/ may include, link or
 * load this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / reported baseline.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  Parity status:
 *   - layout addressing, memcpy transfer, fill, validation, permutation:
 *     pinned by the reference's own known-answer tests (see tests/test_oracle_*.py).
 *   - fp8(e4m3) <-> bf16 cast: the reference has no implementation; PARITY UNPINNED
 *     against the reference, pinned instead against torch CPU (tests/golden/).
 */
#ifndef KVBM_ORACLE_H
#define KVBM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib/kvbm-physical/src/layout/config.rs:151-163 (BlockDimension) */
enum { ORACLE_BLOCK_IS_FIRST_DIM = 0, ORACLE_BLOCK_IS_SECOND_DIM = 1 };
enum { ORACLE_LAYOUT_FC = 0, ORACLE_LAYOUT_LW = 1 };

#define ORACLE_MAX_LAYERS 256

/* LayoutConfig (layout/config.rs:14-56) + the derived strides of
 * FullyContiguousLayout (fully_contiguous.rs:158-162) /
 * LayerSeparateLayout (layer_separate.rs:171-184). */
typedef struct {
  int kind;      /* ORACLE_LAYOUT_FC | ORACLE_LAYOUT_LW */
  int block_dim; /* LW only */
  size_t num_blocks, num_layers, outer_dim, page_size, inner_dim, dtype_width_bytes;
  /* derived */
  size_t region_size, block_stride, layer_stride, outer_stride;
  uintptr_t layer_base[ORACLE_MAX_LAYERS]; /* FC: layer_base[0] is the allocation base */
} oracle_layout;

/* Error codes (0 = ok).  Mirror the error sites of the reference. */
enum {
  ORACLE_OK = 0,
  ORACLE_ERR_CONFIG = 1,          /* LayoutConfig validation (config.rs:16-44,165-180) */
  ORACLE_ERR_RANGE = 2,           /* block/layer/outer id out of range */
  ORACLE_ERR_LENGTH_MISMATCH = 3, /* validation.rs:27-35 */
  ORACLE_ERR_DUP_DST = 4,         /* validation.rs:17-19 */
  ORACLE_ERR_OVERLAP = 5,         /* validation.rs:21-23 */
  ORACLE_ERR_INCOMPATIBLE = 6,    /* memcpy.rs:49-63 layer/outer mismatch */
  ORACLE_ERR_SIZE_MISMATCH = 7,   /* memcpy.rs:143-153 */
};

int oracle_layout_init_fc(oracle_layout* L, uintptr_t base, size_t num_blocks, size_t num_layers,
                          size_t outer_dim, size_t page_size, size_t inner_dim, size_t dtype_width);
int oracle_layout_init_lw(oracle_layout* L, const uintptr_t* layer_bases, int block_dim,
                          size_t num_blocks, size_t num_layers, size_t outer_dim, size_t page_size,
                          size_t inner_dim, size_t dtype_width);
/* like the two above, but dtype_width==1 (fp8) is accepted: the extension config 3 needs. */
int oracle_layout_init_fc_ext(oracle_layout* L, uintptr_t base, size_t num_blocks, size_t num_layers,
                              size_t outer_dim, size_t page_size, size_t inner_dim, size_t dtype_width);
int oracle_layout_init_lw_ext(oracle_layout* L, const uintptr_t* layer_bases, int block_dim,
                              size_t num_blocks, size_t num_layers, size_t outer_dim,
                              size_t page_size, size_t inner_dim, size_t dtype_width);

size_t oracle_required_bytes(const oracle_layout* L);     /* config.rs:64-71 */
size_t oracle_bytes_per_block(const oracle_layout* L);    /* config.rs:76-82 */
size_t oracle_required_allocation(const oracle_layout* L, size_t idx); /* required_allocations */

/* Layout::memory_region (layout/mod.rs:73-78). Returns ORACLE_ERR_RANGE as the reference errs. */
int oracle_memory_region(const oracle_layout* L, size_t block, size_t layer, size_t outer,
                         uintptr_t* addr, size_t* size);

/* validate_block_transfer, debug flavour (transfer/validation.rs:168-225). same_layout is the
 * Arc-pointer identity test of are_same_layout (validation.rs:96-102). */
int oracle_validate_block_transfer(const size_t* src_ids, size_t n_src, const size_t* dst_ids,
                                   size_t n_dst, const oracle_layout* src, const oracle_layout* dst,
                                   int same_layout);

/* can_use_whole_block_transfer (transfer/mod.rs:150-173) */
int oracle_can_use_whole_block(const oracle_layout* src, const oracle_layout* dst, int has_range,
                               size_t layer_begin, size_t layer_end);

/* execute_memcpy_transfer (transfer/executor/memcpy.rs:30-165): single thread, synchronous.
 * has_range==0 means layer_range None. */
int oracle_execute_memcpy_transfer(const oracle_layout* src, const oracle_layout* dst,
                                   const size_t* src_ids, const size_t* dst_ids, size_t n,
                                   int has_range, size_t layer_begin, size_t layer_end);

/* Same chunk list partitioned over nthreads host threads (BASELINE.md B0'). */
int oracle_execute_memcpy_transfer_mt(const oracle_layout* src, const oracle_layout* dst,
                                      const size_t* src_ids, const size_t* dst_ids, size_t n,
                                      int has_range, size_t layer_begin, size_t layer_end,
                                      int nthreads);

/* fill_blocks / fill_layers + fill_memory_region (transfer/fill.rs:51-213).
 * pattern: -1 = Sequential ((block+layer+offset)%256), else Constant(pattern&0xff). */
int oracle_fill_blocks(const oracle_layout* L, const size_t* ids, size_t n, int pattern);
int oracle_fill_layers(const oracle_layout* L, const size_t* ids, size_t n, size_t layer_begin,
                       size_t layer_end, int pattern);

/* K1 semantics: kvbm_kernels_vectorized_copy_kernel (lib/kvbm-kernels/cuda/tensor_kernels.cu:494-541):
 * for every pair copy copy_size bytes, any alignment. */
void oracle_vectorized_copy(void* const* src_ptrs, void* const* dst_ptrs, size_t copy_size,
                            size_t num_pairs);

/* K2 / K3: block<->universal permutation (tensor_kernels.cu:109-118,150-228). elem = bytes per
 * element (2,2,4,8 for F16,BF16,F32,F64); layout 0 = NHD, 1 = HND. block_ptrs has
 * num_blocks*nl*no entries, universal_ptrs num_blocks. */
void oracle_universal_from_block(void* const* universal_ptrs, const void* const* block_ptrs,
                                 size_t num_blocks, size_t nh, size_t nl, size_t no, size_t nt,
                                 size_t hd, size_t elem, int layout);
void oracle_block_from_universal(const void* const* universal_ptrs, void* const* block_ptrs,
                                 size_t num_blocks, size_t nh, size_t nl, size_t no, size_t nt,
                                 size_t hd, size_t elem, int layout);

/* fp8 e4m3fn <-> bf16 (PARITY UNPINNED vs reference; pinned vs torch CPU in tests/golden).
 * up:   exact; NaN codes 0x7f/0xff -> 0x7fc0 (torch's canonical quiet NaN).
 * down: round-to-nearest-even, saturate-to-finite (|x|>=464 -> +-448), NaN -> 0x7f|sign. */
uint16_t oracle_e4m3_to_bf16(uint8_t v);
uint8_t oracle_bf16_to_e4m3_satfinite(uint16_t v);
void oracle_cast_e4m3_to_bf16(const uint8_t* src, uint16_t* dst, size_t n);
void oracle_cast_bf16_to_e4m3(const uint16_t* src, uint8_t* dst, size_t n);

/* Transfer with element-width change: every (block,layer,outer) region is converted elementwise.
 * cast_mode: 0 none (plain oracle_execute_memcpy_transfer layer-wise path), 1 e4m3->bf16,
 * 2 bf16->e4m3.  Region element counts must match (page_size*inner_dim equal). */
int oracle_execute_cast_transfer(const oracle_layout* src, const oracle_layout* dst,
                                 const size_t* src_ids, const size_t* dst_ids, size_t n,
                                 int has_range, size_t layer_begin, size_t layer_end, int cast_mode);

#ifdef __cplusplus
}
#endif
#endif

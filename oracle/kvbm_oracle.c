#define _GNU_SOURCE
/*
 * kvbm_oracle.c -- CPU restatement of the reference KV-block transfer path (see kvbm_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into, loaded by, or called from the product path.
 */
#include "kvbm_oracle.h"

#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

/* ---- LayoutConfig validation: layout/config.rs:16-44 (ranges), :165-180 (custom) ---- */
static int is_pow2(size_t x) { return x != 0 && (x & (x - 1)) == 0; }

static int validate_config(size_t nb, size_t nl, size_t no, size_t page, size_t inner, size_t dtw,
                           int allow_fp8)
{
  if (nb < 1 || nl < 1 || page < 1 || inner < 1) return ORACLE_ERR_CONFIG;
  if (no < 1 || no > 2) return ORACLE_ERR_CONFIG; /* config.rs:24 range(min=1,max=2) */
  /* config.rs:173-180: power of two and within 2..=8 */
  if (!is_pow2(dtw)) return ORACLE_ERR_CONFIG;
  if (dtw > 8) return ORACLE_ERR_CONFIG;
  if (dtw < 2 && !(allow_fp8 && dtw == 1)) return ORACLE_ERR_CONFIG;
  if (nl > ORACLE_MAX_LAYERS) return ORACLE_ERR_CONFIG;
  return ORACLE_OK;
}

static void set_common(oracle_layout* L, size_t nb, size_t nl, size_t no, size_t page, size_t inner,
                       size_t dtw)
{
  memset(L, 0, sizeof(*L));
  L->num_blocks = nb;
  L->num_layers = nl;
  L->outer_dim = no;
  L->page_size = page;
  L->inner_dim = inner;
  L->dtype_width_bytes = dtw;
  L->region_size = page * inner * dtw; /* fully_contiguous.rs:158, layer_separate.rs:171 */
}

static int init_fc(oracle_layout* L, uintptr_t base, size_t nb, size_t nl, size_t no, size_t page,
                   size_t inner, size_t dtw, int allow_fp8)
{
  int rc = validate_config(nb, nl, no, page, inner, dtw, allow_fp8);
  if (rc) return rc;
  set_common(L, nb, nl, no, page, inner, dtw);
  L->kind = ORACLE_LAYOUT_FC;
  /* fully_contiguous.rs:159-162 */
  L->outer_stride = L->region_size;
  L->layer_stride = L->outer_stride * no;
  L->block_stride = L->layer_stride * nl;
  L->layer_base[0] = base;
  return ORACLE_OK;
}

static int init_lw(oracle_layout* L, const uintptr_t* bases, int block_dim, size_t nb, size_t nl,
                   size_t no, size_t page, size_t inner, size_t dtw, int allow_fp8)
{
  int rc = validate_config(nb, nl, no, page, inner, dtw, allow_fp8);
  if (rc) return rc;
  set_common(L, nb, nl, no, page, inner, dtw);
  L->kind = ORACLE_LAYOUT_LW;
  L->block_dim = block_dim;
  /* layer_separate.rs:173-184 */
  if (block_dim == ORACLE_BLOCK_IS_SECOND_DIM) {
    L->block_stride = L->region_size;
    L->outer_stride = L->block_stride * nb;
  } else {
    L->outer_stride = L->region_size;
    L->block_stride = L->outer_stride * no;
  }
  for (size_t i = 0; i < nl; ++i) L->layer_base[i] = bases[i];
  return ORACLE_OK;
}

int oracle_layout_init_fc(oracle_layout* L, uintptr_t base, size_t nb, size_t nl, size_t no,
                          size_t page, size_t inner, size_t dtw)
{
  return init_fc(L, base, nb, nl, no, page, inner, dtw, 0);
}
int oracle_layout_init_lw(oracle_layout* L, const uintptr_t* bases, int block_dim, size_t nb,
                          size_t nl, size_t no, size_t page, size_t inner, size_t dtw)
{
  return init_lw(L, bases, block_dim, nb, nl, no, page, inner, dtw, 0);
}
int oracle_layout_init_fc_ext(oracle_layout* L, uintptr_t base, size_t nb, size_t nl, size_t no,
                              size_t page, size_t inner, size_t dtw)
{
  return init_fc(L, base, nb, nl, no, page, inner, dtw, 1);
}
int oracle_layout_init_lw_ext(oracle_layout* L, const uintptr_t* bases, int block_dim, size_t nb,
                              size_t nl, size_t no, size_t page, size_t inner, size_t dtw)
{
  return init_lw(L, bases, block_dim, nb, nl, no, page, inner, dtw, 1);
}

size_t oracle_required_bytes(const oracle_layout* L)
{
  return L->num_blocks * L->num_layers * L->outer_dim * L->page_size * L->inner_dim *
         L->dtype_width_bytes;
}
size_t oracle_bytes_per_block(const oracle_layout* L)
{
  return L->num_layers * L->outer_dim * L->page_size * L->inner_dim * L->dtype_width_bytes;
}
size_t oracle_required_allocation(const oracle_layout* L, size_t idx)
{
  (void)idx;
  if (L->kind == ORACLE_LAYOUT_FC) return L->block_stride * L->num_blocks; /* fully_contiguous.rs:283 */
  return L->num_blocks * L->outer_dim * L->region_size;                     /* layer_separate.rs:294 */
}

/* fully_contiguous.rs:225-257 and layer_separate.rs:215-247 */
int oracle_memory_region(const oracle_layout* L, size_t b, size_t l, size_t o, uintptr_t* addr,
                         size_t* size)
{
  if (b >= L->num_blocks || l >= L->num_layers || o >= L->outer_dim) return ORACLE_ERR_RANGE;
  if (L->kind == ORACLE_LAYOUT_FC)
    *addr = L->layer_base[0] + b * L->block_stride + l * L->layer_stride + o * L->outer_stride;
  else
    *addr = L->layer_base[l] + b * L->block_stride + o * L->outer_stride;
  if (size) *size = L->region_size;
  return ORACLE_OK;
}

/* ---- validation.rs ---- */
static int cmp_size(const void* a, const void* b)
{
  size_t x = *(const size_t*)a, y = *(const size_t*)b;
  return (x > y) - (x < y);
}

int oracle_validate_block_transfer(const size_t* src_ids, size_t n_src, const size_t* dst_ids,
                                   size_t n_dst, const oracle_layout* src, const oracle_layout* dst,
                                   int same_layout)
{
  if (n_src != n_dst) return ORACLE_ERR_LENGTH_MISMATCH; /* validation.rs:177-183 */
  size_t n = n_dst;
  /* validate_dst_unique: validation.rs:55-70 */
  if (n > 1) {
    size_t* tmp = (size_t*)malloc(n * sizeof(size_t));
    memcpy(tmp, dst_ids, n * sizeof(size_t));
    qsort(tmp, n, sizeof(size_t), cmp_size);
    for (size_t i = 1; i < n; ++i)
      if (tmp[i] == tmp[i - 1]) {
        free(tmp);
        return ORACLE_ERR_DUP_DST;
      }
    free(tmp);
  }
  /* validate_disjoint_same_layout: validation.rs:112-136 */
  if (same_layout && n > 0) {
    size_t* tmp = (size_t*)malloc(n * sizeof(size_t));
    memcpy(tmp, src_ids, n * sizeof(size_t));
    qsort(tmp, n, sizeof(size_t), cmp_size);
    for (size_t i = 0; i < n; ++i)
      if (bsearch(&dst_ids[i], tmp, n, sizeof(size_t), cmp_size)) {
        free(tmp);
        return ORACLE_ERR_OVERLAP;
      }
    free(tmp);
  }
  /* validate_block_ids_in_range: validation.rs:139-158 */
  for (size_t i = 0; i < n; ++i) {
    if (src_ids[i] >= src->num_blocks) return ORACLE_ERR_RANGE;
    if (dst_ids[i] >= dst->num_blocks) return ORACLE_ERR_RANGE;
  }
  return ORACLE_OK;
}

/* transfer/mod.rs:150-173 */
int oracle_can_use_whole_block(const oracle_layout* src, const oracle_layout* dst, int has_range,
                               size_t lb, size_t le)
{
  int full = !has_range || (lb == 0 && le == src->num_layers);
  if (!full) return 0;
  return src->kind == ORACLE_LAYOUT_FC && dst->kind == ORACLE_LAYOUT_FC;
}

/* memcpy.rs:30-63 precondition checks */
static int check_compat(const oracle_layout* src, const oracle_layout* dst)
{
  if (src->num_layers != dst->num_layers) return ORACLE_ERR_INCOMPATIBLE;
  if (src->outer_dim != dst->outer_dim) return ORACLE_ERR_INCOMPATIBLE;
  return ORACLE_OK;
}

int oracle_execute_memcpy_transfer(const oracle_layout* src, const oracle_layout* dst,
                                   const size_t* src_ids, const size_t* dst_ids, size_t n,
                                   int has_range, size_t lb, size_t le)
{
  int rc = check_compat(src, dst);
  if (rc) return rc;
  if (!has_range) {
    lb = 0;
    le = src->num_layers; /* memcpy.rs:68 */
  }
  if (oracle_can_use_whole_block(src, dst, has_range, lb, le)) {
    /* execute_whole_block_memcpy: memcpy.rs:98-120 */
    size_t bytes = oracle_bytes_per_block(src);
    for (size_t i = 0; i < n; ++i) {
      uintptr_t s, d;
      if ((rc = oracle_memory_region(src, src_ids[i], 0, 0, &s, NULL))) return rc;
      if ((rc = oracle_memory_region(dst, dst_ids[i], 0, 0, &d, NULL))) return rc;
      memcpy((void*)d, (const void*)s, bytes);
    }
    return ORACLE_OK;
  }
  /* execute_layer_wise_memcpy: memcpy.rs:126-165 */
  for (size_t i = 0; i < n; ++i)
    for (size_t l = lb; l < le; ++l)
      for (size_t o = 0; o < src->outer_dim; ++o) {
        uintptr_t s, d;
        size_t ss, ds;
        if ((rc = oracle_memory_region(src, src_ids[i], l, o, &s, &ss))) return rc;
        if ((rc = oracle_memory_region(dst, dst_ids[i], l, o, &d, &ds))) return rc;
        if (ss != ds) return ORACLE_ERR_SIZE_MISMATCH;
        memcpy((void*)d, (const void*)s, ss);
      }
  return ORACLE_OK;
}

/* ---- multi-threaded variant: same chunk list, partitioned by block index ---- */
typedef struct {
  const oracle_layout *src, *dst;
  const size_t *src_ids, *dst_ids;
  size_t begin, end;
  int has_range;
  size_t lb, le;
  int rc;
  int cpu; /* >= 0: pin this worker to that CPU (bench hygiene: no migration between NUMA nodes mid-copy) */
} mt_job;

static void* mt_worker(void* p)
{
  mt_job* j = (mt_job*)p;
  if (j->cpu >= 0) {
    cpu_set_t one;
    CPU_ZERO(&one);
    CPU_SET(j->cpu, &one);
    (void)pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
  }
  j->rc = oracle_execute_memcpy_transfer(j->src, j->dst, j->src_ids + j->begin,
                                         j->dst_ids + j->begin, j->end - j->begin, j->has_range,
                                         j->lb, j->le);
  return NULL;
}

int oracle_execute_memcpy_transfer_mt(const oracle_layout* src, const oracle_layout* dst,
                                      const size_t* src_ids, const size_t* dst_ids, size_t n,
                                      int has_range, size_t lb, size_t le, int nthreads)
{
  if (nthreads <= 1 || n < 2)
    return oracle_execute_memcpy_transfer(src, dst, src_ids, dst_ids, n, has_range, lb, le);
  if ((size_t)nthreads > n) nthreads = (int)n;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  mt_job* jobs = (mt_job*)malloc(sizeof(mt_job) * nthreads);
  /* worker t runs on the t-th CPU this process is allowed to use */
  cpu_set_t allowed;
  int cpus[CPU_SETSIZE], ncpu = 0;
  if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
    for (int c = 0; c < CPU_SETSIZE; ++c)
      if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
  for (int t = 0; t < nthreads; ++t) {
    mt_job j = {src, dst, src_ids, dst_ids, n * t / nthreads, n * (t + 1) / nthreads,
                has_range, lb, le, 0, ncpu > 0 ? cpus[t % ncpu] : -1};
    jobs[t] = j;
    pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
  }
  int rc = ORACLE_OK;
  for (int t = 0; t < nthreads; ++t) {
    pthread_join(th[t], NULL);
    if (jobs[t].rc && !rc) rc = jobs[t].rc;
  }
  free(th);
  free(jobs);
  return rc;
}

/* ---- fill.rs:191-213 ---- */
static void fill_region(uintptr_t addr, size_t size, size_t block, size_t layer, int pattern)
{
  uint8_t* p = (uint8_t*)addr;
  if (pattern >= 0) {
    memset(p, pattern & 0xff, size);
  } else {
    for (size_t off = 0; off < size; ++off) p[off] = (uint8_t)((block + layer + off) % 256);
  }
}

int oracle_fill_layers(const oracle_layout* L, const size_t* ids, size_t n, size_t lb, size_t le,
                       int pattern)
{
  if (le > L->num_layers) return ORACLE_ERR_RANGE; /* fill.rs:146-152 */
  for (size_t i = 0; i < n; ++i) {
    if (ids[i] >= L->num_blocks) return ORACLE_ERR_RANGE;
    for (size_t l = lb; l < le; ++l)
      for (size_t o = 0; o < L->outer_dim; ++o) {
        uintptr_t a;
        size_t s;
        int rc = oracle_memory_region(L, ids[i], l, o, &a, &s);
        if (rc) return rc;
        fill_region(a, s, ids[i], l, pattern);
      }
  }
  return ORACLE_OK;
}

int oracle_fill_blocks(const oracle_layout* L, const size_t* ids, size_t n, int pattern)
{
  return oracle_fill_layers(L, ids, n, 0, L->num_layers, pattern); /* fill.rs:51-123 */
}

/* ---- K1: tensor_kernels.cu:494-541 — byte-exact copy of each pair, any alignment ---- */
void oracle_vectorized_copy(void* const* src_ptrs, void* const* dst_ptrs, size_t copy_size,
                            size_t num_pairs)
{
  for (size_t i = 0; i < num_pairs; ++i) memcpy(dst_ptrs[i], src_ptrs[i], copy_size);
}

/* ---- K2/K3: tensor_kernels.cu:109-118 (inner offset), :150-187 / :191-228 (index peel) ---- */
static size_t inner_offset(int layout, size_t nt_i, size_t nh_i, size_t hd_i, size_t nt, size_t nh,
                           size_t hd)
{
  if (layout == 0) return ((nt_i * nh) + nh_i) * hd + hd_i; /* NHD */
  return ((nh_i * nt) + nt_i) * hd + hd_i;                   /* HND */
}

static void permute(void* const* universal_ptrs, void* const* block_ptrs, size_t num_blocks,
                    size_t nh, size_t nl, size_t no, size_t nt, size_t hd, size_t elem, int layout,
                    int to_universal)
{
  size_t block_stride = nl * no;
  size_t total_per_block = nh * nl * no * nt * hd;
  for (size_t b = 0; b < num_blocks; ++b) {
    uint8_t* uni = (uint8_t*)universal_ptrs[b];
    for (size_t residual = 0; residual < total_per_block; ++residual) {
      size_t tmp = residual;
      size_t hd_i = tmp % hd;
      tmp /= hd;
      size_t nt_i = tmp % nt;
      tmp /= nt;
      size_t no_i = tmp % no;
      tmp /= no;
      size_t nl_i = tmp % nl;
      tmp /= nl;
      size_t nh_i = tmp;
      uint8_t* chunk = (uint8_t*)block_ptrs[b * block_stride + nl_i * no + no_i];
      size_t off = inner_offset(layout, nt_i, nh_i, hd_i, nt, nh, hd);
      if (to_universal)
        memcpy(uni + residual * elem, chunk + off * elem, elem);
      else
        memcpy(chunk + off * elem, uni + residual * elem, elem);
    }
  }
}

void oracle_universal_from_block(void* const* universal_ptrs, const void* const* block_ptrs,
                                 size_t num_blocks, size_t nh, size_t nl, size_t no, size_t nt,
                                 size_t hd, size_t elem, int layout)
{
  permute(universal_ptrs, (void* const*)block_ptrs, num_blocks, nh, nl, no, nt, hd, elem, layout, 1);
}

void oracle_block_from_universal(const void* const* universal_ptrs, void* const* block_ptrs,
                                 size_t num_blocks, size_t nh, size_t nl, size_t no, size_t nt,
                                 size_t hd, size_t elem, int layout)
{
  permute((void* const*)universal_ptrs, block_ptrs, num_blocks, nh, nl, no, nt, hd, elem, layout, 0);
}

/* ---- fp8 e4m3fn <-> bf16 (no reference implementation: parity unpinned) ---- */
uint16_t oracle_e4m3_to_bf16(uint8_t v)
{
  uint16_t sign = (uint16_t)(v & 0x80) << 8;
  unsigned e = (v >> 3) & 0xf, m = v & 7;
  if (e == 0xf && m == 7) return 0x7fc0; /* NaN (sign dropped, as torch does) */
  if (e == 0) {
    if (m == 0) return sign; /* +-0 */
    /* subnormal: m * 2^-9; normalise */
    int shift = 0;
    while (!(m & 8)) {
      m <<= 1;
      ++shift;
    }
    /* value = (m/8) * 2^(-6 - shift), m in [8,16) */
    unsigned be = (unsigned)(127 - 6 - shift);
    return (uint16_t)(sign | (be << 7) | ((m & 7) << 4));
  }
  return (uint16_t)(sign | ((e + 120) << 7) | (m << 4)); /* 2^(e-7) -> bias 127 */
}

uint8_t oracle_bf16_to_e4m3_satfinite(uint16_t v)
{
  uint8_t sign = (uint8_t)((v >> 8) & 0x80);
  unsigned be = (v >> 7) & 0xff, bm = v & 0x7f;
  if (be == 0xff && bm != 0) return (uint8_t)(sign | 0x7f); /* NaN */
  uint32_t bits = (uint32_t)(v & 0x7fff) << 16;
  float x;
  memcpy(&x, &bits, 4);
  if (!(x < 464.0f)) return (uint8_t)(sign | 0x7e); /* inf and |x|>=464 saturate to 448 (464 ties to even = 448) */
  if (x < 0.015625f) {                               /* below min normal 2^-6: subnormal grid 2^-9 */
    unsigned q = (unsigned)nearbyintf(x * 512.0f);   /* RNE; q==8 lands on code 0x08 = 2^-6 */
    return (uint8_t)(sign | q);
  }
  int ex;
  (void)frexpf(x, &ex); /* x = f * 2^ex, f in [0.5,1) -> floor(log2 x) = ex-1 */
  int e = ex - 1;
  unsigned q = (unsigned)nearbyintf(ldexpf(x, 3 - e)); /* in [8,16] */
  if (q == 16) {
    q = 8;
    ++e;
  }
  unsigned code = ((unsigned)(e + 7) << 3) | (q & 7);
  if (code > 0x7e) code = 0x7e;
  return (uint8_t)(sign | code);
}

void oracle_cast_e4m3_to_bf16(const uint8_t* src, uint16_t* dst, size_t n)
{
  for (size_t i = 0; i < n; ++i) dst[i] = oracle_e4m3_to_bf16(src[i]);
}
void oracle_cast_bf16_to_e4m3(const uint16_t* src, uint8_t* dst, size_t n)
{
  for (size_t i = 0; i < n; ++i) dst[i] = oracle_bf16_to_e4m3_satfinite(src[i]);
}

int oracle_execute_cast_transfer(const oracle_layout* src, const oracle_layout* dst,
                                 const size_t* src_ids, const size_t* dst_ids, size_t n,
                                 int has_range, size_t lb, size_t le, int cast_mode)
{
  int rc = check_compat(src, dst);
  if (rc) return rc;
  if (!has_range) {
    lb = 0;
    le = src->num_layers;
  }
  if (cast_mode == 0)
    return oracle_execute_memcpy_transfer(src, dst, src_ids, dst_ids, n, 1, lb, le);
  size_t src_w = cast_mode == 1 ? 1 : 2, dst_w = cast_mode == 1 ? 2 : 1;
  if (src->dtype_width_bytes != src_w || dst->dtype_width_bytes != dst_w) return ORACLE_ERR_SIZE_MISMATCH;
  size_t elems = src->page_size * src->inner_dim;
  if (elems != dst->page_size * dst->inner_dim) return ORACLE_ERR_SIZE_MISMATCH;
  for (size_t i = 0; i < n; ++i)
    for (size_t l = lb; l < le; ++l)
      for (size_t o = 0; o < src->outer_dim; ++o) {
        uintptr_t s, d;
        if ((rc = oracle_memory_region(src, src_ids[i], l, o, &s, NULL))) return rc;
        if ((rc = oracle_memory_region(dst, dst_ids[i], l, o, &d, NULL))) return rc;
        if (cast_mode == 1)
          oracle_cast_e4m3_to_bf16((const uint8_t*)s, (uint16_t*)d, elems);
        else
          oracle_cast_bf16_to_e4m3((const uint16_t*)s, (uint8_t*)d, elems);
      }
  return ORACLE_OK;
}

"""Known-answer generators transcribed from the reference's kernel tests (shared by CPU + GPU tests).

  lib/kvbm-kernels/tests/memcpy_batch.rs:350-431   byte patterns + (copy_size, num_pairs) cases
  lib/kvbm-kernels/tests/kernel_roundtrip.rs:418-493  position-encoded universal tensors, dims (3,2,2,4,5)
  lib/kvbm-kernels/tests/kernel_roundtrip.rs:241-252  roundtrip dims nh=3 nl=2 no=2 nt=4 hd=5 nb=3
"""
import numpy as np

# (name, copy_size, num_pairs, generator(i, j) -> byte)
COPY_KATS = [
    ("single_copy", 256, 1, lambda i, j: j % 256),                    # memcpy_batch.rs:350-359
    ("multiple_copies", 512, 8, lambda i, j: (i * 31 + j * 7) % 256),  # :363-380
    ("large_copy", 1 << 20, 3, lambda i, j: (i + j) % 251),            # :384-397
    ("odd_size", 999, 4, lambda i, j: (i * 13 + j) % 256),             # :401-414
    ("many_pairs", 64, 256, lambda i, j: (i + j) % 256),               # :418-431
]


def copy_kat_data(copy_size, num_pairs, gen):
    j = np.arange(copy_size, dtype=np.int64)
    return [np.asarray(gen(i, j) if callable(gen) else gen, dtype=np.int64).astype(np.uint8) if True else None
            for i in range(num_pairs)]


PERMUTE_DIMS = dict(nh=3, nl=2, no=2, nt=4, hd=5)
PERMUTE_NB = 3
DTYPES = {0: np.float16, 1: np.uint16, 2: np.float32, 3: np.float64}  # F16, BF16(bits), F32, F64
ELEM = {0: 2, 1: 2, 2: 4, 3: 8}
NHD, HND = 0, 1


def position_encoded_universal(nh, nl, no, nt, hd, dtype=np.float32):
    """kernel_roundtrip.rs:427-430: value = ((((nh_i*nl+nl_i)*no+no_i)*nt+nt_i)*hd+hd_i)."""
    return np.arange(nh * nl * no * nt * hd, dtype=np.int64).reshape(nh, nl, no, nt, hd).astype(dtype)


def make_blocks(universal: np.ndarray, layout: int):
    """Reference permutation (kernel_roundtrip.rs:103-125): [nh,nl,no,nt,hd] -> nl*no chunks."""
    nh, nl, no, nt, hd = universal.shape
    out = []
    for l in range(nl):
        for o in range(no):
            chunk = universal[:, l, o, :, :]          # [nh, nt, hd]
            if layout == NHD:
                chunk = chunk.transpose(1, 0, 2)       # [nt, nh, hd]
            out.append(np.ascontiguousarray(chunk).reshape(-1))
    return out

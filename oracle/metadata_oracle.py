"""TEST INFRASTRUCTURE ONLY -- independent Python restatement of the reference's layout-metadata wire formats
(SURVEY.md §8 f3).  Used by tests/ to check the C++ codec in dynamo_b200/csrc/host/serialized_layout.hpp; the product
never imports this file.

Follows /root/reference:
  LayoutDescriptor (serde_json)        lib/kvbm-physical/src/layout/serialize.rs:66-137
  LayoutConfig                         lib/kvbm-physical/src/layout/config.rs:14-56
  StorageKind / MemoryRegion           lib/memory/src/lib.rs:111-126, 235-242
  NixlMetadata                         lib/kvbm-physical/src/layout/physical.rs:41-46
  KvBlockLayout / BlockDim             lib/kvbm-physical/src/layout/kv_block_layout.rs:20-70
  LayoutHandle (u128)                  lib/kvbm-physical/src/manager/handle.rs:16-50
  RdmaLayoutDescriptors / SerializedLayout::pack  lib/kvbm-physical/src/manager/metadata.rs:14-134 (bincode 2.0.0,
                                       config::standard(): little endian + variable-length integers)

PARITY UNPINNED: bincode 2.0.0 and nixl-sys 0.10.1 (MemType) are third-party crates absent from /root/reference, no
Rust toolchain exists here, and the reference's tests only round-trip (no golden bytes).  The encoding below restates
bincode's published spec; tests/test_serialized_layout.py additionally holds byte strings derived by hand from it.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

STORAGE = ["System", "Pinned", "Device", "Disk"]
MEM_TYPE = ["Dram", "Vram", "Block", "Object", "File", "Unknown"]
KV_LAYOUT = ["UniversalTP", "UniversalPP", "OperationalHND", "OperationalNHD", "Custom", "Unknown"]
BLOCK_DIM = ["Layer", "Outer", "Page", "Head"]
LOGICAL = ["G1", "G2", "G3", "G4"]
LS_BLOCK_DIM = ["BlockIsFirstDim", "BlockIsSecondDim"]


@dataclass
class Descriptor:
    num_blocks: int
    num_layers: int
    outer_dim: int
    page_size: int
    inner_dim: int
    alignment: int = 1
    dtype_width_bytes: int = 2
    num_heads: Optional[int] = None
    location: str = "System"
    location_arg: int = 0
    agent_name: str = ""
    mem_type: str = "Dram"
    device_id: int = 0
    regions: List[Tuple[int, int]] = field(default_factory=list)
    fully_contiguous: bool = True
    block_dim: str = "BlockIsFirstDim"
    kv_block_layout: str = "Unknown"
    custom: Tuple[str, str, str, str] = ("Layer", "Outer", "Page", "Head")
    version: int = 1


@dataclass
class Logical:
    worker_id: int
    layout_id: int
    logical_type: str
    layout: Descriptor

    @property
    def handle(self) -> int:   # LayoutHandle::new (handle.rs:19-22)
        return self.worker_id | (self.layout_id << 64)


@dataclass
class Bundle:
    worker_id: int
    agent_name: str
    nixl_metadata: bytes
    layouts: List[Logical]


# ---------------------------------------------------------------------------------------------- bincode 2 (standard)
def varint(v: int) -> bytes:
    if v < 0:
        raise ValueError("unsigned only")
    if v < 251:
        return bytes([v])
    if v < 1 << 16:
        return b"\xfb" + v.to_bytes(2, "little")
    if v < 1 << 32:
        return b"\xfc" + v.to_bytes(4, "little")
    if v < 1 << 64:
        return b"\xfd" + v.to_bytes(8, "little")
    return b"\xfe" + v.to_bytes(16, "little")


def _str(s: str) -> bytes:
    b = s.encode()
    return varint(len(b)) + b


def _kv(d: Descriptor) -> bytes:
    out = varint(KV_LAYOUT.index(d.kv_block_layout))
    if d.kv_block_layout == "Custom":
        out += b"".join(varint(BLOCK_DIM.index(x)) for x in d.custom)
    return out


def encode_descriptor(d: Descriptor) -> bytes:
    out = varint(d.version)
    for v in (d.num_blocks, d.num_layers, d.outer_dim, d.page_size, d.inner_dim, d.alignment, d.dtype_width_bytes):
        out += varint(v)
    out += b"\x00" if d.num_heads is None else b"\x01" + varint(d.num_heads)
    out += varint(STORAGE.index(d.location))
    if d.location in ("Device", "Disk"):
        out += varint(d.location_arg)
    out += _str(d.agent_name) + varint(MEM_TYPE.index(d.mem_type)) + varint(d.device_id)
    out += varint(len(d.regions))
    for a, s in d.regions:
        out += varint(a) + varint(s)
    if d.fully_contiguous:
        out += varint(0) + varint(0) + _kv(d)
    else:
        out += varint(1) + varint(LS_BLOCK_DIM.index(d.block_dim)) + _kv(d)
    return out


def pack(b: Bundle) -> bytes:
    """SerializedLayout::pack (metadata.rs:120-134)."""
    out = varint(b.worker_id) + _str(b.agent_name) + varint(len(b.nixl_metadata)) + bytes(b.nixl_metadata)
    out += varint(len(b.layouts))
    for l in b.layouts:
        out += varint(l.handle) + varint(LOGICAL.index(l.logical_type)) + encode_descriptor(l.layout)
    return out


class _R:
    def __init__(self, b: bytes):
        self.b, self.i = b, 0

    def take(self, n: int) -> bytes:
        if self.i + n > len(self.b):
            raise ValueError("UnexpectedEnd")
        out = self.b[self.i:self.i + n]
        self.i += n
        return out

    def varint(self) -> int:
        t = self.take(1)[0]
        if t < 251:
            return t
        return int.from_bytes(self.take({251: 2, 252: 4, 253: 8, 254: 16}[t]), "little")

    def str(self) -> str:
        return self.take(self.varint()).decode()


def _dkv(r: _R, d: Descriptor) -> None:
    d.kv_block_layout = KV_LAYOUT[r.varint()]
    if d.kv_block_layout == "Custom":
        d.custom = tuple(BLOCK_DIM[r.varint()] for _ in range(4))


def decode_descriptor(r: _R) -> Descriptor:
    version = r.varint()
    cfg = [r.varint() for _ in range(7)]
    nh = r.varint() if r.take(1)[0] else None
    loc = STORAGE[r.varint()]
    arg = r.varint() if loc in ("Device", "Disk") else 0
    agent = r.str()
    mt = MEM_TYPE[r.varint()]
    dev = r.varint()
    regions = [(r.varint(), r.varint()) for _ in range(r.varint())]
    d = Descriptor(*cfg, num_heads=nh, location=loc, location_arg=arg, agent_name=agent, mem_type=mt, device_id=dev,
                   regions=regions, version=version)
    kind = r.varint()
    if kind == 0:
        d.fully_contiguous = True
        if r.varint() != 0:
            raise ValueError("UnexpectedVariant")
    elif kind == 1:
        d.fully_contiguous = False
        d.block_dim = LS_BLOCK_DIM[r.varint()]
    else:
        raise ValueError("UnexpectedVariant")
    _dkv(r, d)
    return d


def unpack(blob: bytes) -> Bundle:
    """SerializedLayout::unpack (metadata.rs:140-144)."""
    r = _R(blob)
    wid, agent = r.varint(), r.str()
    md = r.take(r.varint())
    out = Bundle(wid, agent, md, [])
    for _ in range(r.varint()):
        h = r.varint()
        lt = LOGICAL[r.varint()]
        out.layouts.append(Logical(h & ((1 << 64) - 1), (h >> 64) & 0xFFFF, lt, decode_descriptor(r)))
    return out


# ---------------------------------------------------------------------------------------------- serde_json
def to_json_obj(d: Descriptor) -> dict:
    kv = {"Custom": list(d.custom)} if d.kv_block_layout == "Custom" else d.kv_block_layout
    details = ({"FullyContiguous": {"block_format": "Operational", "kv_block_layout": kv}} if d.fully_contiguous else
               {"LayerSeparate": {"block_dim": d.block_dim, "kv_block_layout": kv}})
    return {
        "version": d.version,
        "layout_config": {"num_blocks": d.num_blocks, "num_layers": d.num_layers, "outer_dim": d.outer_dim, "page_size": d.page_size,
                          "inner_dim": d.inner_dim, "alignment": d.alignment, "dtype_width_bytes": d.dtype_width_bytes,
                          "num_heads": d.num_heads},
        "location": {d.location: d.location_arg} if d.location in ("Device", "Disk") else d.location,
        "nixl_metadata": {"agent_name": d.agent_name, "mem_type": d.mem_type, "device_id": d.device_id},
        "memory_descriptors": [{"addr": a, "size": s} for a, s in d.regions],
        "layout_type_details": details,
    }


def to_json(d: Descriptor) -> str:
    """serde_json::to_string: compact, struct fields in declaration order."""
    return json.dumps(to_json_obj(d), separators=(",", ":"), ensure_ascii=False)


def from_json(text: str) -> Descriptor:
    o = json.loads(text)
    c = o["layout_config"]
    loc = o["location"]
    if isinstance(loc, dict):
        (lname, larg), = loc.items()
    else:
        lname, larg = loc, 0
    (kind, det), = o["layout_type_details"].items()
    kv = det.get("kv_block_layout", "Unknown")
    d = Descriptor(c["num_blocks"], c["num_layers"], c["outer_dim"], c["page_size"], c["inner_dim"], c["alignment"],
                   c["dtype_width_bytes"], c.get("num_heads"), lname, larg, o["nixl_metadata"]["agent_name"],
                   o["nixl_metadata"]["mem_type"], o["nixl_metadata"]["device_id"],
                   [(m["addr"], m["size"]) for m in o["memory_descriptors"]], kind == "FullyContiguous",
                   det.get("block_dim", "BlockIsFirstDim"), version=o["version"])
    if isinstance(kv, dict):
        d.kv_block_layout, d.custom = "Custom", tuple(kv["Custom"])
    else:
        d.kv_block_layout = kv
    return d

"""The BASELINE.json configurations as reproducible synthetic workloads.

Each workload names: element width, pattern, blob size, generator seed / byte mask, planted matches and
the searches that make up one "step".  Blobs come from the counter-based generator in synth.py, so any
byte range can be produced independently on the CPU (oracle) and on a GPU (each rank fills only its
slice); planted matches are a short, deterministic list of (offset, bytes) patches applied on top.
"""
import dataclasses
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .synth import mt19937_bytes, synth_bytes

MiB = 1 << 20
GiB = 1 << 30

HIRAGANA = "あいうえおかきくけこさしすせそたちつてとなにぬねのはひふへほまみむめもやゆよらりるれろわをゃっゅょ"   # src/gui/constants.hpp:48


@dataclasses.dataclass
class Search:
    """One search of a step: pattern + endianness."""
    name: str
    pattern: dict
    big_endian: bool = False


@dataclasses.dataclass
class Workload:
    key: str
    description: str
    bits: int
    size: int                      # bytes per GPU (weak scaling: the file is n_gpus * size)
    seed: int
    searches: List[Search]
    byte_mask: int = 0xFF
    n_planted: int = 0
    block_size: int = 524288       # SearchConfig default (include/mmoore/search_engine.hpp:36)
    plant_pattern: Optional[dict] = None
    generator: str = "splitmix"    # "splitmix": counter based (synth.py);  "mt19937": the reference benchmark's generator
    scaling: str = "weak"          # "weak": `size` bytes per GPU;  "strong": `size` bytes in total, split over the GPUs
    single_gpu_size: int = 0       # bytes scanned when a multi-GPU configuration is measured on ONE GPU (its per-GPU slice)

    def scaled(self, size):
        return dataclasses.replace(self, size=int(size))


WORKLOADS = {
    # configs[0]: the reference's own bench_search case (benchmarks/bench_search.cpp), 6-char keyword
    # data: std::mt19937(42) + uniform_int_distribution<unsigned>(0, 255), exactly benchmarks/bench_search.cpp:11-22
    "cfg1": Workload("cfg1", "8-bit relative search, 6-char ASCII keyword 'monkey', 16 MiB (mt19937(42) data of the reference benchmark)",
                     8, 16 * MiB, 0x5EED0001, [Search("8bit-monkey", dict(keyword="monkey", wildcard=0))], n_planted=64,
                     generator="mt19937"),
    # configs[1]: the configuration the metric is quoted on
    "cfg2": Workload("cfg2", "16-bit LE+BE relative search, 8-char keyword 'mo*key*s' (2 wildcards), 512 MiB", 16,
                     512 * MiB, 0x5EED0002,
                     [Search("16le-mo*key*s", dict(keyword="mo*key*s", wildcard=ord("*")), False),
                      Search("16be-mo*key*s", dict(keyword="mo*key*s", wildcard=ord("*")), True)], n_planted=256),
    # configs[2]
    "cfg3": Workload("cfg3", "8-bit value scan, 10 values, 4 GiB", 8, 4 * GiB, 0x5EED0003,
                     [Search("8bit-values10", dict(values=[10, 12, 15, 11, 30, 31, 29, 40, 41, 45]))], n_planted=1024),
    # configs[3]
    "cfg4": Workload("cfg4", "16-bit LE custom character sequence (49-char Hiragana table), 6-char keyword, 16 GiB",
                     16, 16 * GiB, 0x5EED0004,
                     [Search("16le-kana", dict(keyword="わたしたちは", wildcard=0, char_seq=HIRAGANA))], n_planted=4096,
                     scaling="strong", single_gpu_size=2 * GiB),
    # configs[4]
    "cfg5": Workload("cfg5", "8-bit 3-char keyword 'abc' over a 16-symbol low-entropy blob, 64 GiB (8 x 8 GiB)", 8,
                     8 * GiB, 0x5EED0005, [Search("8bit-abc-dense", dict(keyword="abc", wildcard=0))], byte_mask=0x0F),
}


def _pattern_values(pat) -> List[Optional[int]]:
    if pat.get("values") is not None:
        return [int(v) for v in pat["values"]]
    kw = pat["keyword"]
    cps = [c if isinstance(c, int) else ord(c) for c in kw]
    wc = pat.get("wildcard", 0)
    seq = pat.get("char_seq") or ""
    idx = {ord(ch): i for i, ch in enumerate(seq)}
    return [None if c == wc else (idx.get(c, 0) if seq else c) for c in cps]


_PATCH_CACHE = {}


def planted_patches(w: Workload, total_size: int) -> List[Tuple[int, bytes]]:
    """Cached front of :func:`_planted_patches` (the list depends only on the workload's constants and the file size)."""
    key = (w.key, w.size, w.seed, w.n_planted, w.block_size, w.bits, total_size)
    if key not in _PATCH_CACHE:
        _PATCH_CACHE[key] = _planted_patches(w, total_size)
    return _PATCH_CACHE[key]


def _planted_patches(w: Workload, total_size: int) -> List[Tuple[int, bytes]]:
    """Deterministic (file offset, bytes) patches that plant shifted copies of the workload's pattern.
    Half of the 16-bit plants sit at odd byte offsets; big-endian searches get big-endian plants; some
    plants sit right before a block edge so the overlap logic is exercised."""
    if w.n_planted == 0 or total_size < 4096:
        return []
    rng = np.random.default_rng(w.seed ^ 0xA5A5)
    W = w.bits // 8
    vmax = (1 << w.bits) - 1
    patches = []
    n = max(1, int(w.n_planted * (total_size / w.size))) if total_size != w.size else w.n_planted
    slots = np.sort(rng.choice(total_size // 256 - 2, size=min(n, total_size // 256 - 2), replace=False)) + 1
    for k, slot in enumerate(slots):
        srch = w.searches[k % len(w.searches)]
        vals = _pattern_values(srch.pattern)
        L = len(vals)
        off = int(slot) * 256 + int(rng.integers(0, 64)) * W
        if k % 8 == 7:   # hug the next block edge: the match straddles it
            edge = ((off // w.block_size) + 1) * w.block_size
            if edge + 64 < total_size:
                off = edge - (L // 2) * W
        if W == 2 and k % 2 == 1:
            off += 1
        if off + L * W + 2 > total_size:
            continue
        base = int(rng.integers(0, vmax + 1))
        first = next(v for v in vals if v is not None)
        raw = bytearray()
        filler = rng.integers(0, 256, size=L * W, dtype=np.uint8)
        for i, v in enumerate(vals):
            if v is None:
                raw += bytes(filler[i * W:(i + 1) * W])
                continue
            x = (base + v - first) & vmax
            raw += x.to_bytes(W, "big" if srch.big_endian else "little")
        patches.append((off, bytes(raw)))
    return patches


def host_blob(w: Workload, first_byte=0, nbytes=None, total_size=None) -> np.ndarray:
    """The bytes [first_byte, first_byte + nbytes) of the workload's file, on the host."""
    total_size = total_size or w.size
    nbytes = total_size - first_byte if nbytes is None else nbytes
    if w.generator == "mt19937":
        b = mt19937_bytes(first_byte + nbytes, 42)[first_byte:]
        if w.byte_mask != 0xFF:
            b = b & np.uint8(w.byte_mask)
    else:
        b = synth_bytes(nbytes, w.seed, first_byte=first_byte, byte_mask=w.byte_mask)
    for off, raw in planted_patches(w, total_size):
        lo, hi = max(off, first_byte), min(off + len(raw), first_byte + nbytes)
        if lo < hi:
            b[lo - first_byte:hi - first_byte] = np.frombuffer(raw[lo - off:hi - off], dtype=np.uint8)
    return b


def device_blob(w: Workload, first_byte=0, nbytes=None, total_size=None, device="cuda"):
    """Same bytes generated directly in HBM (torch uint8 tensor); nothing of size O(nbytes) touches the host."""
    import torch

    from .synth import synth_fill_device
    total_size = total_size or w.size
    nbytes = total_size - first_byte if nbytes is None else nbytes
    if w.generator != "splitmix":          # small host-generated blobs (cfg1: 16 MiB) are simply uploaded
        host = host_blob(w, first_byte, nbytes, total_size)
        buf = torch.empty(nbytes + 64, dtype=torch.uint8, device=device)
        buf[:nbytes] = torch.from_numpy(host).to(device)
        return buf[:nbytes]
    lo8 = first_byte & ~7
    hi8 = (first_byte + nbytes + 7) & ~7
    buf = torch.empty(hi8 - lo8 + 64, dtype=torch.uint8, device=device)   # 64 spare bytes keep 16-byte reads in bounds
    synth_fill_device(buf[: hi8 - lo8], w.seed, first_byte=lo8, byte_mask=w.byte_mask)
    view = buf[first_byte - lo8: first_byte - lo8 + nbytes]
    idx, val = [], []
    for off, raw in planted_patches(w, total_size):
        lo, hi = max(off, first_byte), min(off + len(raw), first_byte + nbytes)
        if lo < hi:
            idx.append(np.arange(lo - first_byte, hi - first_byte, dtype=np.int64))
            val.append(np.frombuffer(raw[lo - off:hi - off], dtype=np.uint8))
    if idx:
        view[torch.from_numpy(np.concatenate(idx)).to(device)] = torch.from_numpy(np.concatenate(val).copy()).to(device)
    return view

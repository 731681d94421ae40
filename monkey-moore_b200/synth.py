"""Synthetic ROM blobs: a counter-based generator that produces the same bytes on
the CPU (numpy, here) and on the GPU (``mmg_synth_fill`` in csrc/), addressable by byte range so that a
64 GiB blob never has to exist on the host.

    byte i = byte (i mod 8), little endian, of splitmix64(seed ^ (i // 8)), AND byte_mask
"""
import numpy as np

_M = (1 << 64) - 1


def _splitmix64(x):
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_bytes(nbytes, seed, first_byte=0, byte_mask=0xFF):
    """Stream bytes [first_byte, first_byte + nbytes) as a numpy uint8 array."""
    w0 = first_byte // 8
    w1 = (first_byte + nbytes + 7) // 8
    idx = np.arange(w0, w1, dtype=np.uint64)
    words = _splitmix64(np.uint64(seed & _M) ^ idx)
    b = words.view(np.uint8) if words.dtype.byteorder in ("=", "<", "|") else words.byteswap().view(np.uint8)
    b = b[first_byte - w0 * 8: first_byte - w0 * 8 + nbytes]
    if byte_mask != 0xFF:
        b = b & np.uint8(byte_mask)
    return np.ascontiguousarray(b)


def mt19937_bytes(nbytes, seed=42):
    """The reference benchmark's data (benchmarks/bench_search.cpp:11-22 of the reference): ``std::mt19937 rng(seed)``
    drawn through ``std::uniform_int_distribution<unsigned>(0, 255)``.  libstdc++ (GCC >= 11) maps a 32-bit engine
    onto a power-of-two range with Lemire's multiply-shift, which for 256 values is the TOP byte of every 32-bit
    draw and never rejects; numpy's legacy seeding is init_genrand(seed), i.e. the same engine state."""
    bg = np.random.MT19937()
    bg._legacy_seeding(int(seed))
    return np.ascontiguousarray((bg.random_raw(int(nbytes)) >> np.uint64(24)).astype(np.uint8))


def synth_fill_device(tensor, seed, first_byte=0, byte_mask=0xFF):
    """Fills a CUDA uint8 torch tensor (size and first_byte multiples of 8) with stream bytes."""
    from . import _check, lib
    assert tensor.is_cuda and tensor.is_contiguous()
    import torch
    torch.cuda.synchronize(tensor.device)
    _check(lib().mmg_synth_fill(tensor.data_ptr(), tensor.numel() * tensor.element_size(), int(seed) & _M,
                                int(first_byte), int(byte_mask)))
    return tensor

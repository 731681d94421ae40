"""ctypes bindings for the two CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  -- oracle/_ref/libmmoracle.so, the C restatement (oracle/mm_oracle.c).
* ``Ref``     -- oracle/_ref/libmmref.so, the UNMODIFIED reference sources compiled in
                 place (oracle/Makefile); absent => ``Ref.available()`` is False.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
OUT_DIR = os.path.join(ORACLE_DIR, "_ref")

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i16p = C.POINTER(C.c_int16)
intp = C.POINTER(C.c_int)


def build(force=False):
    """Compiles oracle/ (and oracle/_ref from /root/reference when it is present)."""
    need = force or not os.path.exists(os.path.join(OUT_DIR, "libmmoracle.so"))
    if os.path.isdir("/root/reference/src/core") and not os.path.exists(os.path.join(OUT_DIR, "libmmref.so")):
        need = True
    if need:
        subprocess.run(["make", "-C", ORACLE_DIR] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)


def _u32(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.uint32))
    return a, a.ctypes.data_as(u32p)


def codepoints(s):
    """str / list of ints -> list of code points."""
    if isinstance(s, str):
        return [ord(c) for c in s]
    return [int(c) for c in s]


class MMError(Exception):
    pass


class Oracle:
    """The C restatement.  One instance == one compiled pattern."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build()
            lib = C.CDLL(os.path.join(OUT_DIR, "libmmoracle.so"))
            lib.mmo_compile_keyword.restype = C.c_void_p
            lib.mmo_compile_keyword.argtypes = [u32p, C.c_int, C.c_uint32, u32p, C.c_int, C.c_int, intp]
            lib.mmo_compile_values.restype = C.c_void_p
            lib.mmo_compile_values.argtypes = [i16p, C.c_int, C.c_int, intp]
            lib.mmo_free.argtypes = [C.c_void_p]
            lib.mmo_keyword_len.argtypes = [C.c_void_p]
            lib.mmo_mode.argtypes = [C.c_void_p]
            lib.mmo_search.restype = C.c_int64
            lib.mmo_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, u64p, u32p, C.c_uint64]
            lib.mmo_search_slice.restype = C.c_int64
            lib.mmo_search_slice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p, u32p,
                                             C.c_uint64]
            lib.mmo_set_complete.argtypes = [C.c_int]
            lib.mmo_table_size.argtypes = [C.c_void_p]
            lib.mmo_table.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, u32p, u32p]
            lib.mmo_engine.restype = C.c_int64
            lib.mmo_engine.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int,
                                       u64p, u32p, C.c_uint64]
            lib.mmo_num_blocks.restype = C.c_uint64
            lib.mmo_num_blocks.argtypes = [C.c_uint64, C.c_uint32]
            lib.mmo_engine_synth.restype = C.c_int64
            lib.mmo_engine_synth.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64,
                                             C.c_uint64, C.c_int, u64p, u32p, C.c_void_p, C.c_uint64, C.c_int, u64p,
                                             u64p, u32p, C.c_uint64]
            cls._lib = lib
        return cls._lib

    def __init__(self, bits, keyword=None, wildcard=0, char_seq=(), values=None):
        lib = self.lib()
        self.bits = bits
        err = C.c_int(0)
        if values is not None:
            v = np.ascontiguousarray(np.asarray(values, dtype=np.int16))
            self.h = lib.mmo_compile_values(v.ctypes.data_as(i16p), len(v), bits, C.byref(err))
        else:
            kw, kwp = _u32(codepoints(keyword))
            sq, sqp = _u32(codepoints(char_seq))
            self.h = lib.mmo_compile_keyword(kwp, len(kw), int(wildcard), sqp, len(sq), bits, C.byref(err))
        if not self.h:
            raise MMError({1: "Skip table index out of bounds", 2: "empty keyword", 3: "non-terminating pattern",
                           4: "bad argument"}.get(err.value, "error %d" % err.value))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().mmo_free(self.h)
            self.h = None

    @property
    def mode(self):
        return self.lib().mmo_mode(self.h)

    def table(self, v0, v1):
        n = self.lib().mmo_table_size(self.h)
        k = np.zeros(max(n, 1), np.uint32)
        v = np.zeros(max(n, 1), np.uint32)
        self.lib().mmo_table(self.h, int(v0), int(v1), k.ctypes.data_as(u32p), v.ctypes.data_as(u32p))
        return {int(k[i]): int(v[i]) for i in range(n)}

    def search(self, data):
        """data: numpy array of uint8 / uint16 elements (host order).  -> (positions, vals[n,2])"""
        data = np.ascontiguousarray(data)
        assert data.dtype == (np.uint8 if self.bits == 8 else np.uint16)
        n = self.lib().mmo_search(self.h, data.ctypes.data, data.size, None, None, 0)
        pos = np.zeros(max(n, 1), np.uint64)
        vals = np.zeros((max(n, 1), 2), np.uint32)
        self.lib().mmo_search(self.h, data.ctypes.data, data.size, pos.ctypes.data_as(u64p),
                              vals.ctypes.data_as(u32p), n)
        return pos[:n], vals[:n]

    @classmethod
    def set_complete(cls, on):
        """complete-match mode of the product (every window is evaluated); returns the previous setting"""
        return bool(cls.lib().mmo_set_complete(int(bool(on))))

    def search_slice(self, data, start, owned):
        """The search loop entered at element ``start`` of ``data`` and left when the chain reaches element ``owned``
        (or the data ends).  -> (positions, vals[n,2], exit position)"""
        data = np.ascontiguousarray(data)
        assert data.dtype == (np.uint8 if self.bits == 8 else np.uint16)
        ex = C.c_uint64(0)
        n = self.lib().mmo_search_slice(self.h, data.ctypes.data, data.size, int(start), int(owned), C.byref(ex), None, None, 0)
        pos = np.zeros(max(n, 1), np.uint64)
        vals = np.zeros((max(n, 1), 2), np.uint32)
        self.lib().mmo_search_slice(self.h, data.ctypes.data, data.size, int(start), int(owned), C.byref(ex),
                                    pos.ctypes.data_as(u64p), vals.ctypes.data_as(u32p), n)
        return pos[:n], vals[:n], int(ex.value)

    def engine(self, file_bytes, block_size, big_endian=False, wrap32=True):
        """file_bytes: numpy uint8 image of the file.  -> (offsets, vals[n,2])"""
        fb = np.ascontiguousarray(file_bytes, dtype=np.uint8)
        n = self.lib().mmo_engine(self.h, fb.ctypes.data, fb.size, block_size, int(big_endian), int(wrap32),
                                  None, None, 0)
        off = np.zeros(max(n, 1), np.uint64)
        vals = np.zeros((max(n, 1), 2), np.uint32)
        self.lib().mmo_engine(self.h, fb.ctypes.data, fb.size, block_size, int(big_endian), int(wrap32),
                              off.ctypes.data_as(u64p), vals.ctypes.data_as(u32p), n)
        return off[:n], vals[:n]


    def engine_synth(self, seed, byte_mask, total_size, block_size, first_block=0, nblocks=0, big_endian=False,
                     patches=(), threads=1, want_list=False, list_cap=1 << 24):
        """The engine (64-bit offsets) over blocks of a SYNTHETIC file regenerated block by block on the host
        (oracle/mm_oracle_stream.c).  patches: [(offset, bytes)] in application order.
        -> dict(count, s0, s1[, offsets, values])   (s0/s1: order-sensitive digest, see :func:`digest`)"""
        order = sorted(range(len(patches)), key=lambda k: patches[k][0])      # stable: equal offsets keep their order
        po = np.array([patches[k][0] for k in order], dtype=np.uint64)
        pl = np.array([len(patches[k][1]) for k in order], dtype=np.uint32)
        pb = np.frombuffer(b"".join(patches[k][1] for k in order) or b"\0", dtype=np.uint8).copy()
        out3 = np.zeros(3, np.uint64)
        off = vals = None
        if want_list:
            off = np.zeros(list_cap, np.uint64)
            vals = np.zeros((list_cap, 2), np.uint32)
        n = self.lib().mmo_engine_synth(self.h, int(seed) & ((1 << 64) - 1), int(byte_mask), int(total_size),
                                        int(block_size), int(first_block), int(nblocks), int(bool(big_endian)),
                                        po.ctypes.data_as(u64p), pl.ctypes.data_as(u32p), pb.ctypes.data, len(order),
                                        int(threads), out3.ctypes.data_as(u64p),
                                        off.ctypes.data_as(u64p) if want_list else None,
                                        vals.ctypes.data_as(u32p) if want_list else None, list_cap if want_list else 0)
        if n < 0:
            raise MMError("mmo_engine_synth: bad arguments")
        r = dict(count=int(out3[0]), s0=int(out3[1]), s1=int(out3[2]))
        if want_list:
            k = min(n, list_cap)
            r["offsets"], r["values"] = off[:k], vals[:k]
        return r


_M64 = (1 << 64) - 1


def digest(offsets, values, first_index=0):
    """Order-sensitive digest of a match list, the numpy twin of oracle/mm_oracle_stream.c:
    h_i = mix64(off_i + G * ((v0 | v1 << 16) + 1));  -> (count, S0 = sum h_i, S1 = sum h_i * (2 (first_index + i) + 1))."""
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(off)
    if n == 0:
        return 0, 0, 0
    v = np.asarray(values)
    if v.ndim == 2:
        word = (v[:, 0].astype(np.uint64) & np.uint64(0xFFFF)) | (v[:, 1].astype(np.uint64) << np.uint64(16))
    else:
        word = v.astype(np.uint64)
    s0 = s1 = 0
    with np.errstate(over="ignore"):
        for a in range(0, n, 1 << 22):          # bounded temporaries for 10^8-entry lists
            b = min(n, a + (1 << 22))
            x = off[a:b] + np.uint64(0x9E3779B97F4A7C15) * (word[a:b] + np.uint64(1))
            z = x + np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            h = z ^ (z >> np.uint64(31))
            idx = np.arange(first_index + a, first_index + b, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
            s0 = (s0 + int(h.sum(dtype=np.uint64))) & _M64
            s1 = (s1 + int((h * idx).sum(dtype=np.uint64))) & _M64
    return n, s0, s1


def compose(parts):
    """(count, S0, S1) of the concatenation of adjacent ranges, each digested with first_index = 0."""
    n = s0 = s1 = 0
    for c, a, b in parts:
        s1 = (s1 + b + 2 * n * a) & _M64
        s0 = (s0 + a) & _M64
        n += c
    return n, s0, s1


class Ref:
    """The unmodified reference (oracle/_ref/libmmref.so)."""

    _lib = None

    @classmethod
    def available(cls):
        try:
            build()
        except Exception:
            pass
        return os.path.exists(os.path.join(OUT_DIR, "libmmref.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build()
            lib = C.CDLL(os.path.join(OUT_DIR, "libmmref.so"))
            lib.ref_last_error.restype = C.c_char_p
            sig = [C.c_int, u32p, C.c_int, C.c_uint32, u32p, C.c_int, i16p, C.c_int]
            lib.ref_search.restype = C.c_int64
            lib.ref_search.argtypes = sig + [C.c_void_p, C.c_uint64, u64p, u32p, u32p, u32p, C.c_uint64, C.c_uint64]
            lib.ref_compile.restype = C.c_int
            lib.ref_compile.argtypes = sig
            lib.ref_time_search.restype = C.c_double
            lib.ref_time_search.argtypes = sig + [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_int64)]
            lib.ref_engine_run.restype = C.c_void_p
            lib.ref_engine_run.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, u32p, C.c_int, C.c_uint32, u32p,
                                           C.c_int, i16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
            for f in ("ref_engine_count", "ref_engine_entries", "ref_engine_progress_count"):
                getattr(lib, f).restype = C.c_uint64
                getattr(lib, f).argtypes = [C.c_void_p]
            lib.ref_engine_get.argtypes = [C.c_void_p, u64p, u32p, u32p, u32p, intp, intp]
            lib.ref_engine_preview.restype = C.c_char_p
            lib.ref_engine_preview.argtypes = [C.c_void_p, C.c_uint64]
            lib.ref_engine_free.argtypes = [C.c_void_p]
            cls._lib = lib
        return cls._lib

    @staticmethod
    def _pattern_args(bits, keyword, wildcard, char_seq, values):
        kw, kwp = _u32(codepoints(keyword or []))
        sq, sqp = _u32(codepoints(char_seq or []))
        if values is not None:
            v = np.ascontiguousarray(np.asarray(values, dtype=np.int16))
            vp, nv = v.ctypes.data_as(i16p), len(v)
        else:
            v, vp, nv = None, None, 0
        keep = (kw, sq, v)
        return keep, [bits, kwp, len(kw), int(wildcard), sqp, len(sq), vp, nv]

    @classmethod
    def compile_ok(cls, bits, keyword=None, wildcard=0, char_seq=(), values=None):
        keep, args = cls._pattern_args(bits, keyword, wildcard, char_seq, values)
        return cls.lib().ref_compile(*args) == 0

    @classmethod
    def search(cls, bits, data, keyword=None, wildcard=0, char_seq=(), values=None):
        """-> (positions, [dict per match])"""
        lib = cls.lib()
        data = np.ascontiguousarray(data)
        keep, args = cls._pattern_args(bits, keyword, wildcard, char_seq, values)
        n = lib.ref_search(*args, data.ctypes.data, data.size, None, None, None, None, 0, 0)
        if n < 0:
            raise MMError(lib.ref_last_error().decode())
        nseq = max(len(keep[1]), 2)
        pos = np.zeros(max(n, 1), np.uint64)
        sizes = np.zeros(max(n, 1), np.uint32)
        keys = np.zeros(max(n * nseq, 1), np.uint32)
        vals = np.zeros(max(n * nseq, 1), np.uint32)
        lib.ref_search(*args, data.ctypes.data, data.size, pos.ctypes.data_as(u64p), sizes.ctypes.data_as(u32p),
                       keys.ctypes.data_as(u32p), vals.ctypes.data_as(u32p), n, n * nseq)
        maps, e = [], 0
        for i in range(n):
            maps.append({int(keys[e + j]): int(vals[e + j]) for j in range(sizes[i])})
            e += int(sizes[i])
        return pos[:n], maps

    @classmethod
    def time_search(cls, bits, data, iters=3, keyword=None, wildcard=0, char_seq=(), values=None):
        lib = cls.lib()
        data = np.ascontiguousarray(data)
        keep, args = cls._pattern_args(bits, keyword, wildcard, char_seq, values)
        m = C.c_int64(0)
        t = lib.ref_time_search(*args, data.ctypes.data, data.size, iters, C.byref(m))
        if t < 0:
            raise MMError(lib.ref_last_error().decode())
        return t, m.value

    @classmethod
    def engine(cls, bits, path, keyword=None, wildcard=ord("*"), char_seq=(), values=None, big_endian=False,
               threads=1, block=524288, preview_width=50, previews=False, abort_after=0, count_only=False):
        """-> dict(offsets, maps, previews, progress=[(pct, step)]); count_only=True -> dict(count, seconds) where
        seconds covers SearchEngine::run alone (not the conversion of its result maps to Python objects)."""
        lib = cls.lib()
        keep, args = cls._pattern_args(bits, keyword, wildcard, char_seq, values)
        bits_, kwp, L, wc, sqp, nseq, vp, nv = args
        if vp is None:
            dummy = np.zeros(1, np.int16)
            vp = dummy.ctypes.data_as(i16p)
        import time as _time
        t0 = _time.perf_counter()
        h = lib.ref_engine_run(bits_, os.fsencode(path), 0 if values is not None else 1, int(big_endian), kwp, L,
                               wc, sqp, nseq, vp, nv, threads, block, preview_width, int(previews), abort_after)
        seconds = _time.perf_counter() - t0
        if not h:
            raise MMError(lib.ref_last_error().decode())
        try:
            n = lib.ref_engine_count(h)
            if count_only:
                return dict(count=int(n), seconds=seconds)
            ne = lib.ref_engine_entries(h)
            npg = lib.ref_engine_progress_count(h)
            off = np.zeros(max(n, 1), np.uint64)
            sizes = np.zeros(max(n, 1), np.uint32)
            keys = np.zeros(max(ne, 1), np.uint32)
            vals = np.zeros(max(ne, 1), np.uint32)
            pct = np.zeros(max(npg, 1), np.int32)
            step = np.zeros(max(npg, 1), np.int32)
            lib.ref_engine_get(h, off.ctypes.data_as(u64p), sizes.ctypes.data_as(u32p), keys.ctypes.data_as(u32p),
                               vals.ctypes.data_as(u32p), pct.ctypes.data_as(intp), step.ctypes.data_as(intp))
            maps, e = [], 0
            for i in range(n):
                maps.append({int(keys[e + j]): int(vals[e + j]) for j in range(sizes[i])})
                e += int(sizes[i])
            previews_out = [lib.ref_engine_preview(h, i).decode("utf-8", "replace") for i in range(n)]
            return dict(offsets=off[:n].copy(), maps=maps, previews=previews_out,
                        progress=[(int(pct[i]), int(step[i])) for i in range(npg)])
        finally:
            lib.ref_engine_free(h)

"""monkey-moore_b200 -- B200-native relative search (Monkey-Moore's hot path) behind the reference's API.

This package is the Python host side of the C-ABI in ``include/mmoore_b200.h`` (implemented by
``libmmoore_b200.so``, built in-tree from ``csrc/``).  It mirrors the reference's public surface:

* :class:`MonkeyMoore`   <-> ``MonkeyMoore<Ty>``            (/root/reference/include/mmoore/monkey_moore.hpp:18-51)
* :class:`SearchConfig`  <-> ``mmoore::SearchConfig``       (/root/reference/include/mmoore/search_engine.hpp:23-38)
* :class:`SearchEngine`  <-> ``mmoore::SearchEngine<T>``    (/root/reference/include/mmoore/search_engine.hpp:47-58)

The directory name contains a hyphen, so import it as ``monkey_moore_b200`` (a shim package at the
repo root) -- both names resolve to this module.

There is no CPU fallback: importing works anywhere (the library loads without a GPU so that its
exported symbols can be checked), but every scan raises :class:`MMError` when no CUDA device exists.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMG_LIB", os.path.join(_HERE, "libmmoore_b200.so"))   # MMG_LIB: development builds

MMG_OK = 0
ERROR_NAMES = {1: "Skip table index out of bounds", 2: "empty keyword", 3: "pattern never advances",
               4: "bad argument", 5: "CUDA error", 6: "keyword too long", 7: "out of memory"}
MEM_HOST, MEM_DEVICE = 0, 1

# every symbol include/mmoore_b200.h declares (tests check the library exports all of them)
C_ABI_SYMBOLS = [
    "mmg_last_error", "mmg_device_count", "mmg_program_create_keyword", "mmg_program_create_values",
    "mmg_program_free", "mmg_program_keyword_len", "mmg_program_mode", "mmg_program_table_size",
    "mmg_program_table", "mmg_search", "mmg_engine_scan", "mmg_engine_scan_async", "mmg_results_wait", "mmg_num_blocks", "mmg_results_count",
    "mmg_results_copy", "mmg_results_unique", "mmg_results_device_offsets", "mmg_results_device_values", "mmg_results_free",
    "mmg_results_stats", "mmg_set_path_override", "mmg_synth_fill", "mmg_set_stream", "mmg_host_alloc", "mmg_host_free", "mmg_host_copy",
    "mmg_comm_unique_id", "mmg_comm_create", "mmg_comm_destroy", "mmg_comm_gather", "mmg_comm_wait",
    "mmg_gathered_count", "mmg_gathered_copy", "mmg_gathered_pieces", "mmg_gathered_free",
    "mmg_chain_begin", "mmg_chain_map", "mmg_chain_entry", "mmg_chain_finish", "mmg_program_max_jump", "mmg_comm_search",
    "mmg_set_complete_matches",
]

_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i16p = C.POINTER(C.c_int16)


class ScanStats(C.Structure):
    _fields_ = [("ms_total", C.c_float), ("ms_filter", C.c_float), ("ms_h2d", C.c_float),
                ("launches", C.c_uint32), ("fast_path", C.c_uint32), ("events", C.c_uint64),
                ("bytes_scanned", C.c_uint64), ("resolve_kind", C.c_uint32), ("chain_entry", C.c_uint32)]


class MMError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def build(force=False):
    """Compiles csrc/ for sm_100a into libmmoore_b200.so (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no fallback implementation)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.mmg_last_error.restype = C.c_char_p
        l.mmg_device_count.restype = C.c_int
        l.mmg_program_create_keyword.argtypes = [_u32p, C.c_int, C.c_uint32, _u32p, C.c_int, C.c_int,
                                                 C.POINTER(C.c_void_p)]
        l.mmg_program_create_values.argtypes = [_i16p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        l.mmg_program_free.argtypes = [C.c_void_p]
        l.mmg_program_keyword_len.argtypes = [C.c_void_p]
        l.mmg_program_mode.argtypes = [C.c_void_p]
        l.mmg_program_table_size.argtypes = [C.c_void_p]
        l.mmg_program_table.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, _u32p, _u32p]
        l.mmg_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
        l.mmg_engine_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint32,
                                      C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
        l.mmg_engine_scan_async.argtypes = l.mmg_engine_scan.argtypes
        l.mmg_results_wait.argtypes = [C.c_void_p]
        l.mmg_num_blocks.restype = C.c_uint64
        l.mmg_num_blocks.argtypes = [C.c_uint64, C.c_uint32]
        l.mmg_results_count.restype = C.c_uint64
        l.mmg_results_count.argtypes = [C.c_void_p]
        l.mmg_results_copy.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _u64p, _u32p]
        l.mmg_results_unique.argtypes = [C.c_void_p, C.c_void_p, _u64p, C.c_uint64, _u64p]
        l.mmg_results_device_offsets.restype = C.c_void_p
        l.mmg_results_device_offsets.argtypes = [C.c_void_p]
        l.mmg_results_device_values.restype = C.c_void_p
        l.mmg_results_device_values.argtypes = [C.c_void_p]
        l.mmg_results_free.argtypes = [C.c_void_p]
        l.mmg_results_stats.argtypes = [C.c_void_p, C.POINTER(ScanStats)]
        l.mmg_set_path_override.argtypes = [C.c_int]
        l.mmg_set_stream.argtypes = [C.c_void_p, C.c_int]
        l.mmg_comm_unique_id.argtypes = [C.c_void_p]
        l.mmg_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]
        l.mmg_comm_destroy.argtypes = [C.c_void_p]
        l.mmg_comm_gather.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
        l.mmg_comm_wait.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        l.mmg_gathered_pieces.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _u64p, C.c_int]
        l.mmg_gathered_count.restype = C.c_uint64
        l.mmg_gathered_count.argtypes = [C.c_void_p, C.c_int]
        l.mmg_gathered_copy.argtypes = [C.c_void_p, C.c_int, _u64p, _u32p]
        l.mmg_gathered_free.argtypes = [C.c_void_p]
        l.mmg_synth_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
        l.mmg_chain_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]
        l.mmg_chain_map.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        l.mmg_chain_entry.restype = C.c_uint32
        l.mmg_chain_entry.argtypes = [C.c_char_p, C.c_int, C.c_int]
        l.mmg_chain_finish.argtypes = [C.c_void_p, C.c_uint32]
        l.mmg_program_max_jump.argtypes = [C.c_void_p]
        l.mmg_comm_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64,
                                      C.POINTER(C.c_void_p)]
        _lib = l
    return _lib


def _check(rc):
    if rc != MMG_OK:
        msg = lib().mmg_last_error().decode() or ERROR_NAMES.get(rc, "error %d" % rc)
        if rc == 1:
            msg = "Skip table index out of bounds"   # the reference's std::runtime_error text
        raise MMError(rc, msg)


def _codepoints(s):
    if isinstance(s, str):
        return [ord(c) for c in s]
    return [int(c) for c in (s or [])]


def _u32(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.uint32))
    return a, a.ctypes.data_as(_u32p)


def set_path_override(mode):
    return lib().mmg_set_path_override(int(mode))


def set_complete_matches(on):
    """Opt-in superset of the reference's result: report every matching window, not only those its skip chain visits
    (``mmg_set_complete_matches``).  Returns the previous setting."""
    return bool(lib().mmg_set_complete_matches(int(bool(on))))


def set_stream(cuda_stream_handle, use_it=True):
    """Run later scans of this thread on the given cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""
    return lib().mmg_set_stream(C.c_void_p(cuda_stream_handle or 0), int(bool(use_it)))


def device_count():
    return lib().mmg_device_count()


class Results:
    """Match list of one scan (device resident until freed)."""

    def __init__(self, handle, program, keep=None):
        self._h = handle
        self._program = program
        self._keep = keep          # input buffer of an asynchronous scan: must outlive it
        self._count = None

    def wait(self):
        """Completes an asynchronous scan (no-op otherwise)."""
        _check(lib().mmg_results_wait(self._h))
        self._keep = None
        return self

    @property
    def count(self):
        if self._count is None:
            self.wait()
            self._count = int(lib().mmg_results_count(self._h))
        return self._count

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.mmg_results_free(self._h)
            self._h = None

    __del__ = close

    def __len__(self):
        return self.count

    def arrays(self):
        """-> (offsets uint64[n], values uint32[n, 2])"""
        off = np.zeros(self.count, np.uint64)
        val = np.zeros((self.count, 2), np.uint32)
        if self.count:
            _check(lib().mmg_results_copy(self._h, 0, self.count, off.ctypes.data_as(_u64p),
                                          val.ctypes.data_as(_u32p)))
        return off, val

    @property
    def offsets(self):
        return self.arrays()[0]

    def unique_indices(self):
        """Indices (ascending) of the first match of every distinct inferred table -- the reference GUI's default
        "one row per table" view (src/gui/monkey_frame.cpp:1236-1245), reduced on the device."""
        n = C.c_uint64(0)
        _check(lib().mmg_results_unique(self._program._h, self._h, None, 0, C.byref(n)))
        idx = np.zeros(n.value, np.uint64)
        if n.value:
            _check(lib().mmg_results_unique(self._program._h, self._h, idx.ctypes.data_as(_u64p), n.value, C.byref(n)))
        return idx

    def device_pointers(self):
        return lib().mmg_results_device_offsets(self._h), lib().mmg_results_device_values(self._h)

    def torch_offsets(self):
        """Match offsets as a torch int64 CUDA tensor (a copy that outlives this object); no host round trip."""
        import torch
        if self.count == 0:
            return torch.empty(0, dtype=torch.int64, device="cuda")
        ptr, _ = self.device_pointers()

        class _View:
            __cuda_array_interface__ = {"shape": (self.count,), "typestr": "<i8", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_View(), device="cuda").clone()

    def stats(self):
        s = ScanStats()
        _check(lib().mmg_results_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in ScanStats._fields_}

    def tables(self):
        """The equivalency_map of every match, as the reference's ``result_type::second``."""
        _, val = self.arrays()
        return [self._program.table(int(v[0]), int(v[1])) for v in val]


def chain_entry(maps):
    """Entry phase of the slice that follows the slices whose maps are given (``mmg_chain_entry``)."""
    if not maps:
        return 0
    stride = max(len(m) for m in maps)
    raw = b"".join(bytes(m) + bytes(stride - len(m)) for m in maps)
    return int(lib().mmg_chain_entry(raw, stride, len(maps)))


class ChainSlice:
    """One slice of a longer chain between ``mmg_chain_begin`` and ``mmg_chain_finish``."""

    def __init__(self, handle, program, keep=None):
        self._h, self.program, self._keep = handle, program, keep

    def map(self):
        """Exit phase for every entry phase (waits for the first half of the scan)."""
        buf = C.create_string_buffer(128)
        n = C.c_int(0)
        _check(lib().mmg_chain_map(self._h, buf, 128, C.byref(n)))
        return bytes(buf.raw[:n.value])

    def finish(self, entry_phase):
        """Enqueues the second half -> :class:`Results` (element indices of the whole buffer)."""
        h, self._h = self._h, None
        res = Results(h, self.program, self._keep)
        _check(lib().mmg_chain_finish(h, int(entry_phase)))
        return res

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.mmg_results_free(self._h)
            self._h = None

    __del__ = close


class Comm:
    """NCCL communicator of the C-ABI for the result gather (one process per GPU).

    ``Comm(rank, world, broadcast_bytes)``: ``broadcast_bytes(b)`` must return rank 0's 128-byte id on every
    rank (e.g. via torch.distributed).  ``gather(results_list)`` -> on rank 0 a list of (offsets, values)
    numpy pairs per search (``fetch=True``) or the per-search counts; ``None`` elsewhere."""

    def __init__(self, rank, world, broadcast_bytes, capacity=8192):
        self.rank, self.world = rank, world
        ident = C.create_string_buffer(128)
        if rank == 0:
            _check(lib().mmg_comm_unique_id(ident))
        raw = broadcast_bytes(bytes(ident.raw))
        h = C.c_void_p()
        _check(lib().mmg_comm_create(C.create_string_buffer(raw, 128), rank, world, int(capacity), C.byref(h)))
        self._h = h

    def gather(self, results, fetch=False, lazy=False):
        """Collective.  Rank 0 gets the per-search counts (or the lists with ``fetch=True``, or a
        :class:`Gathered` with ``lazy=True``); the other ranks get ``None``.  The result lists may be closed
        right after the call on every rank (the library releases them behind the gather's reads)."""
        n = len(results)
        arr = (C.c_void_p * n)(*[r._h for r in results])
        g = C.c_void_p()
        _check(lib().mmg_comm_gather(self._h, arr, n, C.byref(g)))
        if not g:
            return None
        out = Gathered(g, n)
        if lazy:
            return out
        try:
            return out.fetch() if fetch else out.counts()
        finally:
            out.close()

    def search(self, program, data, owned_len, first_element):
        """Collective: ``MonkeyMoore<Ty>::search`` over ONE buffer that is spread over the ranks.  ``data`` is this
        rank's slice: ``owned_len`` elements whose windows it owns, followed -- on every rank but the last -- by at
        least ``keyword_len - 1`` elements of the next rank's slice; its first element is element
        ``first_element`` of the whole buffer.  Returns this rank's part of the match list (element indices of
        the whole buffer); ``gather`` concatenates the parts on rank 0."""
        ptr, nbytes, mem, keep = program._pointer(data)
        h = C.c_void_p()
        _check(lib().mmg_comm_search(self._h, program._h, ptr, int(owned_len), nbytes // (program.bits // 8), mem,
                                     int(first_element), C.byref(h)))
        return Results(h, program, keep)

    def wait(self):
        """Blocks until this rank's part of every gather so far has executed -> device ms of the last gather."""
        ms = C.c_float(0)
        _check(lib().mmg_comm_wait(self._h, C.byref(ms)))
        return float(ms.value)

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.mmg_comm_destroy(self._h)
            self._h = None

    __del__ = close


class Gathered:
    """Rank 0's view of one gather (valid until the next gather on the same communicator)."""

    def __init__(self, handle, nlists):
        self._h, self.nlists = handle, nlists

    def counts(self):
        return [int(lib().mmg_gathered_count(self._h, k)) for k in range(self.nlists)]

    def fetch(self):
        out = []
        for k, n in enumerate(self.counts()):
            off = np.zeros(n, np.uint64)
            val = np.zeros((n, 2), np.uint32)
            if n:
                _check(lib().mmg_gathered_copy(self._h, k, off.ctypes.data_as(_u64p), val.ctypes.data_as(_u32p)))
            out.append((off, val))
        return out

    def pieces(self, k):
        """Device-resident pieces of list k in file order: [(offsets_ptr, values_ptr, n)]."""
        cap = 1024
        offs, vals, ns = (C.c_void_p * cap)(), (C.c_void_p * cap)(), (C.c_uint64 * cap)()
        n = lib().mmg_gathered_pieces(self._h, k, offs, vals, ns, cap)
        return [(offs[i], vals[i], int(ns[i])) for i in range(min(n, cap))]

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.mmg_gathered_free(self._h)
            self._h = None

    __del__ = close


class Program:
    """A compiled pattern: what one ``MonkeyMoore<Ty>`` instance holds after construction."""

    def __init__(self, bits, keyword=None, wildcard=0, char_seq=(), values=None):
        self.bits = int(bits)
        h = C.c_void_p()
        if values is not None:
            v = np.ascontiguousarray(np.asarray(values, dtype=np.int16))
            _check(lib().mmg_program_create_values(v.ctypes.data_as(_i16p), len(v), self.bits, C.byref(h)))
        else:
            kw, kwp = _u32(_codepoints(keyword))
            sq, sqp = _u32(_codepoints(char_seq))
            _check(lib().mmg_program_create_keyword(kwp, len(kw), int(wildcard), sqp, len(sq), self.bits,
                                                    C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.mmg_program_free(self._h)
            self._h = None

    @property
    def keyword_len(self):
        return lib().mmg_program_keyword_len(self._h)

    @property
    def mode(self):
        return lib().mmg_program_mode(self._h)

    def table(self, v0, v1):
        n = lib().mmg_program_table_size(self._h)
        k = np.zeros(max(n, 1), np.uint32)
        v = np.zeros(max(n, 1), np.uint32)
        lib().mmg_program_table(self._h, int(v0), int(v1), k.ctypes.data_as(_u32p), v.ctypes.data_as(_u32p))
        return {int(k[i]): int(v[i]) for i in range(n)}

    # -- scans -------------------------------------------------------------------------------
    def _pointer(self, data):
        """numpy array (host) or torch CUDA tensor / (ptr, nbytes) tuple (device) -> (ptr, nbytes, mem)"""
        if isinstance(data, np.ndarray):
            data = np.ascontiguousarray(data)
            return data.ctypes.data, data.nbytes, MEM_HOST, data
        if isinstance(data, tuple):
            return int(data[0]), int(data[1]), MEM_DEVICE, data
        # torch tensor
        t = data.contiguous()
        nbytes = t.numel() * t.element_size()
        return t.data_ptr(), nbytes, (MEM_DEVICE if t.is_cuda else MEM_HOST), t

    def search(self, data):
        """``MonkeyMoore<Ty>::search(data, data_len)``: one chain over the element buffer."""
        ptr, nbytes, mem, keep = self._pointer(data)
        h = C.c_void_p()
        _check(lib().mmg_search(self._h, ptr, nbytes // (self.bits // 8), mem, C.byref(h)))
        return Results(h, self)

    @property
    def max_jump(self):
        """Number of entry phases of a chain slice (== largest advance of the pattern)."""
        return lib().mmg_program_max_jump(self._h)

    def chain_begin(self, data, owned_len, first_element):
        """First half of a slice of a longer chain (``mmg_chain_begin``): ``data`` = ``owned_len`` owned elements plus,
        unless it is the last slice, at least ``keyword_len - 1`` elements of the next one."""
        ptr, nbytes, mem, keep = self._pointer(data)
        h = C.c_void_p()
        _check(lib().mmg_chain_begin(self._h, ptr, int(owned_len), nbytes // (self.bits // 8), mem, int(first_element), C.byref(h)))
        return ChainSlice(h, self, keep)

    def search_sliced(self, data, slice_len):
        """``search(data)`` computed slice by slice on this GPU (all slices in flight at once): the single-GPU
        rehearsal of ``Comm.search``, and the way to search one chain through a buffer in pieces.
        ``slice_len`` elements per slice, a multiple of ``4096 / sizeof(Ty)``."""
        W = self.bits // 8
        ptr, nbytes, mem, keep = self._pointer(data)
        n = nbytes // W
        tail = self.keyword_len - 1
        slices = []
        first = 0
        while first < n:
            owned = min(slice_len, n - first)
            if n - (first + owned) <= tail:          # no window fits into what would remain: this is the last slice
                owned = n - first
            avail = owned if first + owned >= n else owned + tail
            h = C.c_void_p()
            _check(lib().mmg_chain_begin(self._h, ptr + first * W, owned, avail, mem, first, C.byref(h)))
            slices.append(ChainSlice(h, self, keep))
            first += owned
        maps = [s.map() for s in slices]
        return [s.finish(chain_entry(maps[:k])) for k, s in enumerate(slices)]

    def engine_scan(self, data, block_size, big_endian=False, file_size=None, first_block=0, num_blocks=0,
                    asynchronous=False):
        """The chunk engine of ``SearchEngine<T>::run`` over a file image (or one rank's slice of it).
        ``asynchronous=True`` only enqueues the scan; the returned Results completes on first use / wait()."""
        ptr, nbytes, mem, keep = self._pointer(data)
        if file_size is None:
            file_size = nbytes
        h = C.c_void_p()
        fn = lib().mmg_engine_scan_async if asynchronous else lib().mmg_engine_scan
        _check(fn(self._h, ptr, nbytes, mem, int(file_size), int(block_size), int(first_block),
                  int(num_blocks), int(bool(big_endian)), C.byref(h)))
        return Results(h, self, keep if asynchronous else None)


class MonkeyMoore:
    """Mirror of ``MonkeyMoore<Ty>`` (include/mmoore/monkey_moore.hpp:18-51 of the reference).

    ``MonkeyMoore(bits, keyword, wildcard=0, char_seq=())`` or ``MonkeyMoore(bits, values=[...])``.
    ``search(data)`` returns ``[(position, {char: value})]`` exactly like ``std::vector<result_type>``.
    """

    def __init__(self, bits, keyword=None, wildcard=0, char_seq=(), values=None):
        self.program = Program(bits, keyword=keyword, wildcard=wildcard, char_seq=char_seq, values=values)

    def search(self, data):
        res = self.program.search(data)
        off, val = res.arrays()
        out = [(int(o), self.program.table(int(v[0]), int(v[1]))) for o, v in zip(off, val)]
        res.close()
        return out


from .synth import synth_bytes, synth_fill_device  # noqa: E402,F401
from .engine import SearchConfig, SearchEngine, SearchResult, SearchStep  # noqa: E402,F401

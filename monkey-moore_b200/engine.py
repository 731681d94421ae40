"""Python mirror of ``mmoore::SearchEngine<T>`` over the C-ABI.

Reference: /root/reference/include/mmoore/search_engine.hpp:16-58 (types) and
/root/reference/src/core/search_engine.cpp:23-216 (run), :256-348 (previews).

What changes relative to the reference: the per-block ``ifstream`` reads, the per-alignment buffer
copies, the ``std::async`` pool and its 5 ms polling sleep are gone -- the file image is handed to
``mmg_engine_scan`` which runs every (block, alignment) chain on the GPU in one pass.  What does not
change: block geometry, result order, the progress-callback protocol and the abort contract.
"""
import dataclasses
import enum
import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


class SearchStep(enum.IntEnum):      # search_engine.hpp:40-45
    Initializing = 0
    Searching = 1
    GeneratingPreviews = 2
    Aborting = 3


@dataclasses.dataclass
class SearchConfig:                  # search_engine.hpp:23-38 (same fields, same defaults)
    file_path: str = ""
    is_relative_search: bool = True
    big_endian: bool = False         # endianness == Endianness::Big
    keyword: Sequence = ()
    custom_char_seq: Sequence = ()
    wildcard: int = ord("*")
    reference_values: Sequence[int] = ()
    preferred_num_threads: int = os.cpu_count() or 1   # advisory on the GPU
    preferred_search_block_size: int = 524288          # binding: chains restart at every block
    preferred_preview_width: int = 50


@dataclasses.dataclass
class SearchResult:                  # search_engine.hpp:16-21
    offset: int
    values_map: Dict[int, int]
    preview: str = ""


def _is_set(flag) -> bool:
    if flag is None:
        return False
    if callable(flag):
        return bool(flag())
    if hasattr(flag, "is_set"):
        return bool(flag.is_set())
    if isinstance(flag, (list, tuple)):
        return bool(flag[0])
    return bool(flag)


class SearchEngine:
    """``SearchEngine<T>``: ``SearchEngine(bits, config).run(on_progress, abort_flag, generate_previews)``."""

    # blocks handed to the GPU per scan call; progress callbacks and the abort check run between calls
    BLOCKS_PER_CALL_BYTES = 1 << 30

    def __init__(self, bits: int, config: SearchConfig):
        self.bits = int(bits)
        self.config = dataclasses.replace(config)

    def run(self, on_progress: Optional[Callable[[int, SearchStep], None]] = None, abort_flag=None,
            generate_previews: bool = False) -> List[SearchResult]:
        from . import Program

        cfg = self.config
        cb = on_progress or (lambda pct, step: None)
        if not os.path.exists(cfg.file_path):
            raise RuntimeError("File not found")                       # search_engine.cpp:43-45
        cb(0, SearchStep.Initializing)                                 # :47
        file_size = os.path.getsize(cfg.file_path)
        if cfg.is_relative_search:                                     # :53-62
            program = Program(self.bits, keyword=cfg.keyword, wildcard=cfg.wildcard, char_seq=cfg.custom_char_seq)
        else:
            program = Program(self.bits, values=list(cfg.reference_values))

        block = int(cfg.preferred_search_block_size)
        nblocks = (file_size + block - 1) // block if block > 0 else 0
        W = self.bits // 8
        overlap = (program.keyword_len - 1) * W
        cb(0, SearchStep.Searching)                                    # :80

        image = np.memmap(cfg.file_path, dtype=np.uint8, mode="r") if file_size else np.zeros(0, np.uint8)
        per_call = max(1, self.BLOCKS_PER_CALL_BYTES // max(block, 1))
        total_progress = np.float32(0.0)
        increment = np.float32(100.0) / np.float32(nblocks) if nblocks else np.float32(0.0)   # :75-76
        offsets, values = [], []
        first = 0
        while first < nblocks:
            n = min(per_call, nblocks - first)
            lo = first * block
            hi = min(file_size, (first + n) * block + overlap)
            res = program.engine_scan(np.ascontiguousarray(image[lo:hi]), block, big_endian=cfg.big_endian,
                                      file_size=file_size, first_block=first, num_blocks=n)
            off, val = res.arrays()
            res.close()
            offsets.append(off)
            values.append(val)
            for _ in range(n):                                         # one callback per block, :161-165
                total_progress = np.float32(total_progress + increment)
                cb(int(total_progress), SearchStep.Searching)
                if _is_set(abort_flag):                                # :177-187
                    return []
            first += n
        if _is_set(abort_flag):
            return []
        cb(100, SearchStep.GeneratingPreviews)                         # :191

        off = np.concatenate(offsets) if offsets else np.zeros(0, np.uint64)
        val = np.concatenate(values) if values else np.zeros((0, 2), np.uint32)
        results = [SearchResult(int(o), program.table(int(v[0]), int(v[1]))) for o, v in zip(off, val)]
        if generate_previews and results:                              # :199-213
            reader = _PreviewReader(cfg.file_path, file_size)
            for r in results:
                r.preview = self._generate_preview(reader, r.offset, r.values_map)
        return results

    # -- previews (host side; SURVEY.md section 8f "next" row 1) ------------------------------
    def _generate_preview(self, reader, match_offset, values_map):
        """search_engine.cpp:256-300"""
        cfg = self.config
        W = self.bits // 8
        keyword_len = len(cfg.keyword)
        width = int(cfg.preferred_preview_width)
        kw_half = keyword_len // 2
        window_half = int(width / 2)
        bytes_to_backup = (window_half - kw_half) * W
        bytes_to_backup = (bytes_to_backup + (W - 1)) & ~(W - 1)      # align_up<sizeof(T)>, memory_utils.hpp:13-23
        start = match_offset - bytes_to_backup
        end = start + width * W
        if end > reader.file_size:
            start -= end - reader.file_size
        raw = reader.read(max(0, start), width * W)
        items = len(raw) // W
        if W == 1:
            data = np.frombuffer(raw[:items], dtype=np.uint8)
        else:
            data = np.frombuffer(raw[: items * 2], dtype=">u2" if cfg.big_endian else "<u2")
        return self._decode(values_map, [int(x) for x in data])

    def _decode(self, values_map, raw):
        """search_engine.cpp:302-348"""
        cfg = self.config
        mask = (1 << self.bits) - 1
        ascii_search = len(cfg.custom_char_seq) == 0
        decoding = {}
        for ch in sorted(values_map):                                  # std::map order
            value = values_map[ch]
            if ascii_search and ch in (ord("a"), ord("A")):
                for k in range(26):
                    decoding[(value + k) & mask] = chr(ch + k)
            else:
                decoding[value] = chr(ch)
        if cfg.is_relative_search:
            return "".join(decoding.get(v, "#") for v in raw)
        return " ".join("%0*X" % (2 * (self.bits // 8), v) for v in raw)


class _PreviewReader:
    """An ``std::ifstream`` stand-in that keeps the reference's sticky fail state: once a read comes up
    short (EOF), failbit stays set and every later seek+read yields nothing (search_engine.cpp:202-212)."""

    def __init__(self, path, file_size):
        self.f = open(path, "rb")
        self.file_size = file_size
        self.failed = False

    def read(self, pos, n):
        if self.failed:
            return b""
        self.f.seek(pos)
        raw = self.f.read(n)
        if len(raw) < n:
            self.failed = True
        return raw

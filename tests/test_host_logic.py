"""CPU-side checks of the product: the C-ABI library loads and exports every declared symbol, the pattern
compiler agrees with the oracle (errors and inferred tables), the synthetic generator is reproducible.
No compute calls: scans must FAIL without a GPU (there is no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from _cases import random_pattern, ref_kats
from _oracle import MMError as OracleError
from _oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kw(p):
    return dict(keyword=p.get("keyword"), wildcard=p.get("wildcard", 0), char_seq=p.get("char_seq", ()),
                values=p.get("values"))


def test_library_exports_every_declared_symbol(mm):
    header = open(os.path.join(ROOT, "include", "mmoore_b200.h")).read()
    declared = set(re.findall(r"\b(mmg_[a-z0-9_]+)\s*\(", header))
    assert declared == set(mm.C_ABI_SYMBOLS), declared ^ set(mm.C_ABI_SYMBOLS)
    lib = mm.lib()
    for s in declared:
        assert hasattr(lib, s), s


def test_product_does_not_reference_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "monkey-moore_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh", "Makefile")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "mm_oracle" not in text and "libmmref" not in text and "_oracle" not in text, f


def test_scan_fails_loudly_without_gpu(mm):
    if mm.device_count() > 0:
        pytest.skip("a GPU is present")
    prog = mm.Program(8, keyword="monkey")
    with pytest.raises(mm.MMError) as e:
        prog.search(np.zeros(64, np.uint8))
    assert e.value.code == 5


def test_pattern_compiler_matches_oracle_errors_and_tables(mm):
    rng = np.random.default_rng(5)
    for _ in range(1500):
        bits = int(rng.choice([8, 16]))
        pat = random_pattern(rng, bits)
        try:
            o = Oracle(bits, **kw(pat))
            oerr = None
        except OracleError as e:
            oerr = str(e)
        try:
            p = mm.Program(bits, **kw(pat))
            perr = None
        except mm.MMError as e:
            perr = e
        assert (oerr is None) == (perr is None), (pat, oerr, perr)
        if oerr:
            if "Skip table" in oerr:
                assert perr.code == 1 and str(perr) == "Skip table index out of bounds"
            continue
        assert p.mode == o.mode and p.keyword_len == len(pat.get("values") or pat["keyword"])
        for _ in range(3):
            v0, v1 = int(rng.integers(0, 1 << bits)), int(rng.integers(0, 1 << bits))
            assert p.table(v0, v1) == o.table(v0, v1), (pat, v0, v1)


def test_known_answer_tables_from_reference_values(mm):
    """The inferred tables of the reference's own tests, rebuilt from the raw element values."""
    for k in ref_kats():
        if not k["maps"]:
            continue
        o = Oracle(k["bits"], **kw(k))
        p = mm.Program(k["bits"], **kw(k))
        _, vals = o.search(k["data"])
        assert [p.table(int(v[0]), int(v[1])) for v in vals] == k["maps"], k["name"]


def test_rejects_what_the_reference_cannot_finish(mm):
    for bad in (dict(keyword="a"), dict(keyword="**", wildcard=ord("*")), dict(keyword="*a", wildcard=ord("*"))):
        with pytest.raises(mm.MMError) as e:
            mm.Program(8, **bad)
        assert e.value.code == 3
    assert mm.Program(8, keyword="xy" * 65).keyword_len == 130       # > 128: accepted (per-chain kernels), like the reference
    with pytest.raises(mm.MMError) as e:
        mm.Program(8, keyword="x" * 40000)
    assert e.value.code == 6
    with pytest.raises(mm.MMError):
        mm.Program(8, keyword="")


def test_num_blocks_and_progress_protocol_shape(mm):
    assert mm.lib().mmg_num_blocks(128, 16) == 8          # tests/test_search_engine.cpp:376-396
    assert mm.lib().mmg_num_blocks(0, 16) == 0
    assert mm.lib().mmg_num_blocks(129, 16) == 9


def test_synthetic_generator_is_chunk_addressable(mm):
    whole = mm.synth_bytes(4096, 0x5EED0002)
    assert np.array_equal(mm.synth_bytes(100, 0x5EED0002, first_byte=1003), whole[1003:1103])
    assert mm.synth_bytes(64, 1, byte_mask=0x0F).max() <= 15
    # splitmix64 known answer (seed 0, counter 0 -> state 0): first output 0xE220A8397B1DCDAF
    assert mm.synth_bytes(8, 0).view("<u8")[0] == 0xE220A8397B1DCDAF


def test_workload_blobs_are_slice_consistent(mm):
    import monkey_moore_b200.workloads as wl
    w = wl.WORKLOADS["cfg2"].scaled(1 << 20)
    whole = wl.host_blob(w)
    part = wl.host_blob(w, first_byte=300000, nbytes=50000)
    assert np.array_equal(part, whole[300000:350000])
    o = Oracle(16, keyword="mo*key*s", wildcard=ord("*"))
    le, _ = o.engine(whole, w.block_size, big_endian=False)
    be, _ = o.engine(whole, w.block_size, big_endian=True)
    assert len(le) > 0 and len(be) > 0      # planted matches are found in both byte orders


def test_committed_bench_line_has_the_contract_keys():
    """The bench line the last GPU run produced (profiles/r1_bench_cfg2_n1.json) carries every key of the driver's
    contract: base keys, roofline, cpu_baseline, e2e, gpu_launches, clocks."""
    import json
    path = os.path.join(ROOT, "profiles", "r1_bench_cfg2_n1.json")
    line = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in line, k
    assert line["config"]["workload"] == "cfg2" and "model" not in line["config"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in line["roofline"], k
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in line["cpu_baseline"], k
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in line["e2e"], k
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["gpu_launches"] > 0
    assert line["parity"]["bit_exact"] is True

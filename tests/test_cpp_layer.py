"""The C++17 host layer (include/mmoore/*.hpp + monkey-moore_b200/cpp) -- the reference's public API.

* CPU: the minimal Catch2 stand-in enumerates SECTION / GENERATE leaf paths correctly, and the reference's
  own tests/test_text_utils.cpp passes against this repo's text_utils.hpp.
* GPU: the reference's UNMODIFIED Catch2 sources (tests/test_monkey_moore.cpp, tests/test_search_engine.cpp),
  compiled in place against this repo's headers by monkey-moore_b200/cpp/Makefile, pass against the GPU
  library: known-answer offsets + full value tables, block/threads matrix incl. big-endian, previews,
  progress protocol (exactly 11 callbacks), abort contract, wildcard pass-through, missing-file error.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "ref_unit_tests")


def test_catch2_standin_enumerates_paths(tmp_path):
    exe = str(tmp_path / "shimtest")
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "tests", "catch2_shim"), "-o", exe,
                    os.path.join(ROOT, "tests", "catch2_shim", "main.cpp"),
                    os.path.join(ROOT, "tests", "cpp", "shim_selftest.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "runs: 10" in out.stdout


@pytest.mark.skipif(not os.path.exists(BIN), reason="ref_unit_tests not built (needs /root/reference at build time)")
def test_reference_text_utils_tests_pass():
    out = subprocess.run([BIN, "Text utilities"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_reference_catch2_suite_passes_on_gpu(gpu):
    assert os.path.exists(BIN), "tests/cpp/_build/ref_unit_tests missing: run __graft_entry__.build() where /root/reference exists"
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:])
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    assert "failed: 0" in out.stdout and "test cases: 14" in out.stdout

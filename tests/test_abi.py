"""The C-ABI shared library: it loads on a CPU-only box, exports every symbol include/pcpx.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcpx.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcpx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(pcpx):
    if not os.path.exists(pcpx.LIB_PATH):
        import importlib.util

        spec = importlib.util.spec_from_file_location(
            "pcpx_build", os.path.join(ROOT, "point-cloud-processing_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = ctypes.CDLL(pcpx.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), "libpcpx.so does not export " + n
    # and the binding knows every one of them
    assert set(pcpx.exported_symbols()) == set(names)


def test_header_compiles_as_c_and_cpp(tmp_path):
    src_c = tmp_path / "t.c"
    src_c.write_text('#include "pcpx.h"\nint main(void){pcpx_index_params p = {0}; return p.device;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src_c), "-o", str(tmp_path / "t.o")], check=True)
    src_cpp = tmp_path / "t.cpp"
    src_cpp.write_text('#include "pcpx.h"\nint main(){pcpx_index_params p{}; return p.device;}\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src_cpp), "-o", str(tmp_path / "t2.o")], check=True)


def test_no_cpu_fallback(pcpx):
    if pcpx.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pcpx.PcpxError) as e:
        pcpx.Index(np.zeros((8, 3), np.float32))
    assert e.value.code == -3  # PCPX_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)


def test_product_does_not_link_the_oracle(pcpx):
    out = subprocess.run(["nm", "-D", pcpx.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "ref_cloud" not in out and "emu_" not in out

"""The C-ABI boundary: libjdeflate.so loads, exports every symbol the headers under include/
declare, keeps the reference's public struct layouts, and has no CPU fallback."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from jdeflate_b200 import api
from jdeflate_b200.build import build_cuda, lib_path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    names = []
    for h in sorted((ROOT / "include" / "jdeflate").rglob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
        for m in re.finditer(r"JDEFLATE_API\s+[^;{]*?\b([a-z_0-9]+)\s*\(", text, re.S):
            names.append(m.group(1))
    return sorted(set(names))


@pytest.fixture(scope="module")
def product():
    return api.JDeflateLib(build_cuda())


def test_every_declared_symbol_is_exported(product):
    names = declared_symbols()
    # 5 deflator + 5 inflator + 13 zstrm (+ crc32_ncombine) + version + 4 additive b200 calls
    assert len(names) >= 28, names
    for n in names:
        assert product.has(n), f"{n} declared in include/ but not exported"
    assert not product.missing


def test_exports_are_only_the_api(product):
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib_path())], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert exported == set(declared_symbols())
    # the product never links or names the oracle
    assert "jdo_" not in out


def test_struct_layouts_match_reference_abi():
    # jdeflate/deflator.h:81-99, jdeflate/inflator.h:71-89 (72 B), jdeflate/zstrm.h:105-132 (56 B)
    assert C.sizeof(api.TCodec) == 72
    assert api.TCodec.flush.offset == 12 and api.TCodec.status.offset == 16 and api.TCodec.source.offset == 24
    assert C.sizeof(api.TZStrm) == 56
    assert api.TZStrm.total.offset == 24 and api.TZStrm.usedinput.offset == 48


def test_headers_compile_as_c99_and_cxx(tmp_path):
    src = tmp_path / "t.c"
    src.write_text("#include <jdeflate/deflator.h>\n#include <jdeflate/inflator.h>\n#include <jdeflate/zstrm.h>\n"
                   "#include <jdeflate/b200.h>\n"
                   "typedef char a[sizeof(TDeflator) == 72 ? 1 : -1]; typedef char b[sizeof(TInflator) == 72 ? 1 : -1];\n"
                   "typedef char c[sizeof(TZStrm) == 56 ? 1 : -1];\n"
                   "int main(void) { return DEFLT_OK + INFLT_OK + ZSTRM_OK; }\n")
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++11")):
        lang = ["-x", "c++"] if cc == "g++" else []
        subprocess.run([cc, std, "-Wall", "-Werror", "-I", str(ROOT / "include"), *lang, "-c", str(src),
                        "-o", str(tmp_path / "t.o")], check=True)


def test_version(product):
    assert product.version().startswith("0.4.0")


def test_no_cpu_fallback_without_a_device(product):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    # constructors refuse to build an instance when there is no GPU to run on
    assert not product.lib.deflator_create(0, 6, None)
    assert not product.lib.inflator_create(0, None)
    assert not product.lib.zstrm_create(api.ZSTRM_DEFLATE | api.ZSTRM_GZIP, 6, None)
    assert product.lib.jdb200_device_count() == 0


def _build_example(tmp_path):
    exe = tmp_path / "roundtrip"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(ROOT / "examples" / "roundtrip.c"),
                    "-I", str(ROOT / "include"), "-L", str(lib_path().parent), "-ljdeflate",
                    f"-Wl,-rpath,{lib_path().parent}", "-o", str(exe)], check=True)
    return exe


def test_c_program_written_for_the_reference_links(product, tmp_path):
    """examples/roundtrip.c uses only the reference's C API (the README loops); it must compile
    against include/ and link against libjdeflate.so unchanged."""
    assert _build_example(tmp_path).exists()


@pytest.mark.gpu
def test_c_program_runs_on_the_gpu(product, tmp_path):
    res = subprocess.run([str(_build_example(tmp_path))], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "raw:" in res.stdout and "gzip:" in res.stdout and "round trip ok" in res.stdout

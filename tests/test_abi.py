"""The drop-in boundary without a GPU: libsfh.so loads, exports every symbol include/sfh.h declares, its structs have
the layout the Python mirror assumes, and it fails loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes as C
import os
import re
import subprocess

import pytest

from stonefish_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "sfh.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sfh_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sfh.h but not exported by libsfh.so"
    assert sorted(_abi.PROTOTYPES) == names, "stonefish_b200/_abi.py PROTOTYPES out of sync with include/sfh.h"
    assert lib.sfh_abi_version() == 2


def test_struct_layouts_match_the_c_headers(tmp_path):
    src = tmp_path / "layout.c"
    fields = []
    for s in (_abi.sfh_current, _abi.sfh_ocean, _abi.sfh_waves, _abi.sfh_body_state, _abi.sfh_info, _abi.sfh_mesh_stats, _abi.sfh_timings):
        fields.append(f'printf("{s.__name__} %zu\\n", sizeof({s.__name__}));')
        for name, _ in s._fields_:
            fields.append(f'printf("{s.__name__}.{name} %zu\\n", offsetof({s.__name__}, {name}));')
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sfh.h"\nint main(void){' + "".join(fields) + "return 0;}")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for s in (_abi.sfh_current, _abi.sfh_ocean, _abi.sfh_waves, _abi.sfh_body_state, _abi.sfh_info, _abi.sfh_mesh_stats, _abi.sfh_timings):
        assert int(got[s.__name__]) == C.sizeof(s), s.__name__
        for name, _ in s._fields_:
            assert int(got[f"{s.__name__}.{name}"]) == getattr(s, name).offset, f"{s.__name__}.{name}"
    assert C.sizeof(_abi.sfh_body_state) == 464 and C.sizeof(_abi.sfh_info) == 224


def test_headers_compile_as_c_and_cpp(tmp_path):
    for comp, std, ext in (("gcc", "-std=c99", "c"), ("g++", "-std=c++17", "cpp")):
        src = tmp_path / f"inc.{ext}"
        src.write_text('#include "sfh.h"\nint main(void){ sfh_ocean o; (void)o; return SFH_OK; }\n')
        subprocess.check_call([comp, std, "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", "-o",
                               str(tmp_path / f"inc_{ext}.o"), str(src)])


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a CUDA device")
def test_no_cpu_fallback_without_a_device():
    from stonefish_b200 import HydroContext, SfhError

    with pytest.raises(SfhError) as e:
        HydroContext(0)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)
    lib = _abi.load_library()
    # and nothing computes through a null context
    assert lib.sfh_sync(None) != 0
    assert lib.sfh_step(None, None, None, 0, None) != 0


def test_missing_library_is_a_loud_error(tmp_path, monkeypatch):
    monkeypatch.setattr(_abi, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _abi.load_library(str(tmp_path / "nope.so"))


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under stonefish_b200/ may import, link or call it."""
    pkg = os.path.join(ROOT, "stonefish_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|\borc_[a-z_]+\s*\(|hydro_oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(text), os.path.join(dirpath, f)

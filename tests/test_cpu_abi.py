"""CPU suite: the C-ABI library loads and exports every symbol include/arseg.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from arseg_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "arseg.h")).read()
    return sorted(set(re.findall(r"\b(arseg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "symbol %s declared in include/arseg.h but not exported" % s


def test_binding_covers_header(lib_path):
    assert sorted(_lib.EXPORTED_SYMBOLS) == header_symbols()
    lib = _lib.load()
    assert lib.arseg_abi_version() == _lib.ABI_VERSION == 6


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of the argument structs must have the C compiler's layout of include/arseg.h (size and the offset of
    every field): a plain-C probe compiled with gcc prints them."""
    import subprocess
    structs = {"arseg_creff_args": _lib.CreffArgs, "arseg_conv_desc": getattr(_lib, "ConvDesc", None)}
    text = open(os.path.join(ROOT, "include", "arseg.h")).read()
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "arseg.h"', "int main(void) {"]
    checked = {}
    for cname, cls in structs.items():
        if cls is None or ("} %s;" % cname) not in text:
            continue
        checked[cname] = cls
        src.append('printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            cfield = {"inp": "in"}.get(fname, fname)          # `in` is a Python keyword
            src.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, cfield))
    src.append("return 0; }")
    assert "arseg_creff_args" in checked
    c = tmp_path / "probe.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    seen = 0
    for line in out:
        if not line:
            continue
        cname, field, val = line.split()
        cls = checked[cname]
        if field == "size":
            assert ctypes.sizeof(cls) == int(val), (cname, ctypes.sizeof(cls), val)
        else:
            assert getattr(cls, field).offset == int(val), (cname, field, getattr(cls, field).offset, val)
        seen += 1
    assert seen > 40


def test_error_reporting_without_gpu(lib_path):
    lib = _lib.load()
    # null pointers are rejected before any CUDA call; message is retrievable
    rc = lib.arseg_local_similar_fwd(None, None, None, 1, 1, 1, 1, 3, 3, None)
    assert rc == _lib.E_BADARG
    assert b"similar_forward" in lib.arseg_last_error()
    with pytest.raises(_lib.ArsegError):
        _lib.check(rc, "similar_forward")


def test_product_path_has_no_cpu_fallback():
    import torch
    from arseg_b200 import ops
    with pytest.raises(RuntimeError):
        ops.similar_forward(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4), 3, 3)
    with pytest.raises(RuntimeError):
        ops.warp_feature(torch.zeros(1, 1, 4, 4), torch.zeros(1, 4, 4, 2))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "arseg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, f

"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/b200zstd.h declares, shares its error
numbering with the oracle, and refuses to work without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared_functions():
    txt = open(os.path.join(ROOT, "include", "b200zstd.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200z_[a-z0-9_]+)\s*\(", txt)) - {"b200z_read_fn", "b200z_write_fn"})


def test_header_is_plain_c():
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "b200zstd.h")])
    txt = open(os.path.join(ROOT, "include", "b200zstd.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", txt, flags=re.S) and "cudaStream_t" not in re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def test_library_exports_every_declared_symbol(pkg):
    L = C.CDLL(pkg.lib_path())
    names = _declared_functions()
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/b200zstd.h but not exported"
    assert pkg.lib().b200z_abi_version() == 1


def test_error_numbering_shared_with_oracle(pkg, oracle):
    a = {k.replace("B200Z_", ""): v for v, k in pkg.error_names().items()}
    b = {k.replace("ZO_", ""): v for v, k in oracle.error_names().items()}
    for k, v in b.items():
        assert a.get(k) == v, k
    for code, name in pkg.error_names().items():
        assert pkg.lib().b200z_error_name(code).decode() == name.replace("B200Z_", "")


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the library must fail loudly, never decode on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.B200ZError) as e:
        pkg.Context(0)
    assert pkg.error_names()[e.value.code] == "B200Z_ERR_NO_DEVICE"


def test_product_never_references_oracle():
    """The product path must not import, link or call anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zstd-rs_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "ruzstd_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dirpath, f)
    out = subprocess.check_output(["ldd", os.path.join(ROOT, "zstd-rs_b200", "libb200zstd.so")]).decode()
    assert "oracle" not in out and "libzstd" not in out


def test_xxh64_host(pkg):
    assert pkg.xxh64(b"") == 0xEF46DB3751D8E999
    assert pkg.xxh64(b"Nobody inspects the spammish repetition") == 0xFBCEA83C8A378BF1


def test_struct_layouts_match_the_binding(pkg, tmp_path):
    """Every struct that crosses the C-ABI as an array: the header's layout (as gcc sees it) == the numpy dtype the binding
    fills, field by field (name, offset, size)."""
    B = pkg.binding
    pairs = {"b200z_frame_io": B.FRAME_IO_DTYPE, "b200z_frame_result": B.FRAME_RESULT_DTYPE, "b200z_block_desc": B.BLOCK_DESC_DTYPE,
             "b200z_block_frame": B.BLOCK_FRAME_DTYPE, "b200z_block_status": B.BLOCK_STATUS_DTYPE}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200zstd.h"', 'int main(void) {']
    for s, dt in pairs.items():
        lines.append(f'  printf("{s} * %zu 0\\n", sizeof({s}));')
        for f in dt.names:
            lines.append(f'  printf("{s} {f} %zu %zu\\n", offsetof({s}, {f}), sizeof((({s} *)0)->{f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    seen = 0
    for ln in subprocess.check_output([str(exe)]).decode().split("\n"):
        if not ln:
            continue
        s, f, a, b = ln.split()
        dt = pairs[s]
        if f == "*":
            assert dt.itemsize == int(a), s
        else:
            assert dt.fields[f][1] == int(a) and dt.fields[f][0].itemsize == int(b), (s, f)
            seen += 1
    assert seen == sum(len(dt.names) for dt in pairs.values())

"""CPU-only: the C-ABI shared library builds, loads, and exports every symbol include/icepy4d_b200.h declares."""
import ctypes
import os
import subprocess

import pytest

from icepy4d_b200 import _native, build


@pytest.fixture(scope="module")
def libpath():
    return build.build()


def test_header_parses_and_library_exports_every_symbol(libpath):
    protos = _native.parse_header()
    assert len(protos) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = sorted(set(protos) - exported)
    assert not missing, f"declared in the header but not exported: {missing}"
    extra = sorted(n for n in exported if n.startswith("i4d_") and n not in protos)
    assert not extra, f"exported but not declared in the header: {extra}"


def test_library_loads_without_gpu_and_reports_version(libpath):
    l = _native.lib()
    assert l.i4d_version() == 100
    assert isinstance(_native.last_error(), str)
    assert l.i4d_assignment_workspace_bytes(8192, 8192) > 3 * 256 * 8192 * 4
    assert l.i4d_assignment_workspace_bytes(0, 5) == 0


def test_built_for_sm100a(libpath):
    out = subprocess.run(["cuobjdump", "--list-elf", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in out

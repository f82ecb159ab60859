"""The C-ABI library loads without a GPU and exports every symbol include/rt580.h declares;
compute entry points fail loudly (RT580_FAILURE + message) when no CUDA device exists."""
import ctypes
import os
import re
import subprocess

import numpy as np

import pytest

from conftest import ROOT


def header_functions():
    with open(os.path.join(ROOT, "include", "rt580.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt580_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(pkg):
    names = header_functions()
    assert len(names) >= 25
    out = subprocess.check_output(["nm", "-D", "--defined-only", pkg.LIB_PATH], text=True)
    exported = set(re.findall(r" T (rt580_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, "declared in include/rt580.h but not exported: %s" % missing
    lib = pkg.lib()
    for n in names:
        assert getattr(lib, n) is not None
    # the Python mirror binds the same list
    assert sorted(pkg.EXPORTS) == names


def test_struct_layouts_match_the_header(pkg):
    """ctypes mirrors vs the C structs, via a tiny C program compiled against the header."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "rt580.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(rt580_flat_scene), sizeof(rt580_render_params), sizeof(rt580_stats), sizeof(rt580_scene_info));
  printf("%zu %zu %zu\n", offsetof(rt580_flat_scene, origin_hint), offsetof(rt580_render_params, farfield), offsetof(rt580_stats, kernel_launches));
  return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        with open(c, "w") as f:
            f.write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        a, b = subprocess.check_output([exe], text=True).strip().split("\n")
    sizes = [int(x) for x in a.split()]
    offs = [int(x) for x in b.split()]
    assert sizes == [ctypes.sizeof(pkg.FlatScene), ctypes.sizeof(pkg.RenderParams), ctypes.sizeof(pkg.Stats), ctypes.sizeof(pkg.SceneInfo)]
    assert offs == [pkg.FlatScene.origin_hint.offset, pkg.RenderParams.farfield.offset, pkg.Stats.kernel_launches.offset]


def test_no_gpu_means_loud_failure_not_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    st = pkg.lib().rt580_create(0, ctypes.byref(h))
    assert st == pkg.RT_FAILURE
    msg = pkg.lib().rt580_last_error().decode()
    assert "no CUDA device" in msg and "no CPU fallback" in msg
    with pytest.raises(pkg.Rt580Error):
        pkg.Context(0)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under 580-raytracer_b200/ may import, link or open it."""
    pkg_dir = os.path.join(ROOT, "580-raytracer_b200")
    for base, _, files in os.walk(pkg_dir):
        if os.path.basename(base) == "build":
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                with open(os.path.join(base, fn), errors="replace") as f:
                    txt = f.read()
                assert "oracle580" not in txt and "liboracle" not in txt and "libref580" not in txt, fn
                assert not re.search(r"^\s*(import|from)\s+oracle", txt, flags=re.M), fn


def test_host_alloc_without_a_gpu_is_plain_memory(pkg):
    """rt580_host_alloc / rt580_host_free work on a box without a GPU (malloc), and HostArray wraps them."""
    ha = pkg.HostArray((7, 5, 3), np.int16)
    ha.array[:] = 3
    assert ha.array.sum() == 7 * 5 * 3 * 3
    ha.close()
    pkg.lib().rt580_host_free(None)

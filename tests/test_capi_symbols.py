"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/gsgp_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from go_gsgp_b200 import build, capi
    build.build()
    return capi.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gsgp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsgp_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from go_gsgp_b200 import capi
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gsgp_b200.h but not exported"
    assert set(names) == set(capi.SIGNATURES), set(names) ^ set(capi.SIGNATURES)


def test_struct_layouts_match_header(lib):
    from go_gsgp_b200 import capi
    assert capi.PLAN_DTYPE.itemsize == 40
    assert C.sizeof(capi.GsgpConfig) == 4 * 10 + 8 * 2 + 8 + 4 * 4 + 128 + 8 + 4 * 2
    cfg = capi.GsgpConfig()
    lib.gsgp_config_defaults(C.byref(cfg))
    assert (cfg.population_size, cfg.max_depth_creation, cfg.tournament_size) == (200, 6, 4)   # main_cpu.go:157-163
    assert (cfg.p_crossover, cfg.p_mutation, cfg.minimization_problem) == (0.6, 0.3, 1)


def test_shard_range_partitions_cases(lib):
    for n in (0, 1, 7, 1000, 10_000_000):
        for w in (1, 2, 4, 8):
            prev = 0
            for r in range(w):
                lo, hi = C.c_int32(), C.c_int32()
                lib.gsgp_shard_range(n, r, w, C.byref(lo), C.byref(hi))
                assert lo.value == prev and hi.value >= lo.value
                prev = hi.value
            assert prev == n


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from go_gsgp_b200 import Engine, GsgpError
    with pytest.raises(GsgpError, match="no CPU fallback"):
        Engine(population_size=4, nvar=2, nrow=2, nrow_test=2)


def test_product_never_imports_oracle():
    """The product path must not route through the CPU oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "go_gsgp_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle|oracle[./]binding|gsgp_oracle|libgsgp_oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not pat.search(txt), f"{f} references the oracle"


def test_host_header_symbols_are_exported_and_bound(lib):
    """include/gsgp_host.h: every declared gsgph_* entry point is exported by the same library and bound."""
    from go_gsgp_b200 import host
    src = open(os.path.join(ROOT, "include", "gsgp_host.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(gsgph_[a-z0-9_A-Z]+)\s*\(", src)))
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gsgp_host.h but not exported"
    assert set(names) == set(host.HOST_SIGNATURES), set(names) ^ set(host.HOST_SIGNATURES)


def test_headers_are_plain_c_and_link_from_c(tmp_path):
    """cgo compiles the preamble as C: both headers must be valid C99 (no C++-isms), and a C program must link against
    the library and call it (here: ABI version, config defaults, two host-side formatters -- nothing that needs a GPU)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "cabi.c"
    src.write_text('''
#include <stdio.h>
#include <string.h>
#include "gsgp_b200.h"
#include "gsgp_host.h"
int main(void)
{
    gsgp_config cfg;
    char buf[64];
    gsgp_config_defaults(&cfg);
    if (gsgp_abi_version() != 1 || sizeof(gsgp_plan_entry) != 40) return 1;
    if (gsgph_format_duration(1500000, buf, (int32_t)sizeof buf) < 0 || strcmp(buf, "1.5ms") != 0) return 2;
    if (gsgph_format_float(1e6, buf, (int32_t)sizeof buf) != 5 || strcmp(buf, "1e+06") != 0) return 3;
    printf("%d %s\\n", (int)cfg.tournament_size, buf);
    return 0;
}
''')
    exe = tmp_path / "cabi"
    libdir = os.path.join(root, "go_gsgp_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                           "-o", str(exe), str(src), "-L" + libdir, "-lgsgp_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stderr)
    assert out.stdout.split() == ["4", "1e+06"]                # main_cpu.go:161 tournament_size default, %v of 1e6

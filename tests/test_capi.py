"""Host-side checks that need no GPU: the C-ABI library builds, loads, exports every symbol the header declares, its
ctypes mirror has the C layout, and it fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import importlib
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("dune-multidomain_b200")
capi = pkg.capi


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(pkg.library_path()):
        subprocess.check_call(["make", "-j8", "-C", os.path.join(ROOT, "dune-multidomain_b200", "csrc")])
    return pkg.load_library()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mdb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    syms = _declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "libmdb200.so does not export %s" % s


def test_every_entry_point_is_mapped_in_integration_md():
    """INTEGRATION.md section 1 names, for every function the header declares, the reference interface it replaces (or says that
    it has none).  Families are written with a wildcard (`mdb_halo_*`) or a slash (`mdb_create / mdb_destroy`)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    table = text[text.index("## 1."):text.index("## 2.")]
    names = set(re.findall(r"\bmdb_[a-z0-9_]+\b", table))
    stems = [w[:-1] for w in re.findall(r"\bmdb_[a-z0-9_]+_\*", table)]
    missing = [s for s in _declared_symbols() if s not in names and not any(s.startswith(t) for t in stems)]
    assert not missing, missing


def test_struct_layouts_match_c():
    src = r"""
#include <stdio.h>
#include "mdb200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(mdb_function), sizeof(mdb_bctype), sizeof(mdb_poisson_params),
         sizeof(mdb_l2_params), sizeof(mdb_cvcf_params), sizeof(mdb_propflow_params), sizeof(mdb_solve_stats),
         sizeof(mdb_mortar_params), sizeof(mdb_ndcoupling_params));
  return 0; }
"""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    sizes = [int(x) for x in out]
    mine = [C.sizeof(t) for t in (capi.Function, capi.BCType, capi.PoissonParams, capi.L2Params, capi.CVCFParams,
                                  capi.PropFlowParams, capi.SolveStats, capi.MortarParams, capi.NDCouplingParams)]
    assert sizes == mine


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C (no C++/torch types in the signatures)."""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write('#include "mdb200.h"\nint main(void){return 0;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                               "-c", os.path.join(d, "t.c"), "-o", os.path.join(d, "t.o")])


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.MdbError, match="no CUDA device|CPU fallback"):
        pkg.Mdb(0)


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may touch oracle/."""
    bad = []
    for base in ("dune-multidomain_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cpp")) or f == "Makefile":
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle/|liboracle|orc_[a-z]+\(|import orc", t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def _build_example(outdir, name="testpoisson"):
    exe = os.path.join(outdir, name)
    libdir = os.path.dirname(pkg.library_path())
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", name + ".cc"), "-o", exe, "-L", libdir, "-lmdb200", "-Wl,-rpath," + libdir])
    return exe


def test_cxx_facade_compiles_and_fails_loudly_without_gpu(lib):
    """include/mdb200/multidomain.hh (the reference's spellings over the C ABI) builds as C++11 and links against
    libmdb200.so; without a GPU the driver must stop with the "no CPU fallback" error, not compute anything."""
    import torch
    with tempfile.TemporaryDirectory() as d:
        exe = _build_example(d)
        _build_example(d, "testinstationarypoisson")          # the instationary driver must at least compile and link
        _build_example(d, "testmortarpoisson")                # ... and so must the mortar (enriched coupling) driver
        _build_example(d, "testpoisson-dirichlet-neumann")    # ... and the Dirichlet-Neumann driver
        _build_example(d, "testpowermortarpoisson")           # ... and the power mortar driver
        if torch.cuda.is_available():
            pytest.skip("a GPU is present (the run is covered by the gpu tests)")
        p = subprocess.run([exe, "3", "2.0"], capture_output=True, text=True)
        assert p.returncode == 1 and "no CPU fallback" in p.stderr

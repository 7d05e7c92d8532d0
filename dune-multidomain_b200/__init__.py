"""dune-multidomain_b200 — B200-native backend for dune-multidomain's assembly-and-solve hot path.

The product is ``libmdb200.so`` (CUDA, sm_100a) behind the C ABI of ``include/mdb200.h``; the C++ facade with the
reference's spellings lives in ``include/mdb200/multidomain.hh``.  This Python package is only the plumbing used
by the tests and ``bench.py`` (ctypes binding, torch.distributed launch glue).  The directory name contains a
hyphen, so import it with ``importlib.import_module("dune-multidomain_b200")``.
"""
from . import capi, gmsh, parallel  # noqa: F401
from .capi import Mdb, MdbError, load_library, library_path  # noqa: F401

"""CPU-side checks: the C-ABI library builds, loads and exports every symbol that
include/mdbn_b200.h declares; ctypes mirrors match; the product never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from mdbn_b200 import _lib
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mdbn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mdbn_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), "libmdbn_b200.so does not export %s" % s
    from mdbn_b200 import _lib
    assert sorted(_lib.EXPORTS) == syms


def test_abi_version_and_error_string(lib):
    assert lib.mdbn_abi_version() == 2
    assert isinstance(lib.mdbn_last_error(), bytes)
    assert lib.mdbn_stats_size(784, 500) == 784 * 500 + 500 + 784 + 2


def test_null_arguments_fail_cleanly(lib):
    from mdbn_b200 import _lib
    assert lib.mdbn_cd_step(None, None, None) != 0
    assert b"NULL" in lib.mdbn_last_error()
    assert lib.mdbn_create(None, 0) != 0


def test_struct_layout_matches_header():
    """sizeof(mdbn_cd_args) / sizeof(mdbn_rng) as gcc sees the header == the ctypes mirror."""
    from mdbn_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "mdbn_b200.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(mdbn_rng), sizeof(mdbn_cd_args),
  offsetof(mdbn_cd_args, rng), offsetof(mdbn_cd_args, cost_out), offsetof(mdbn_cd_args, B_total),
  offsetof(mdbn_cd_args, comm)); return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o",
                               os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    got = [ctypes.sizeof(_lib.Rng), ctypes.sizeof(_lib.CdArgs), _lib.CdArgs.rng.offset,
           _lib.CdArgs.cost_out.offset, _lib.CdArgs.B_total.offset, _lib.CdArgs.comm.offset]
    assert [int(x) for x in out] == got


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mdbn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", s, flags=re.M), f
                assert "rbm_oracle" not in s, f


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mdbn_b200
    with pytest.raises(RuntimeError):
        mdbn_b200.RBM(n_visible=4, n_hidden=3)


def test_host_minibatches_match_reference_golden():
    import mdbn_b200
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "minibatches.npz")))
    for n, b in ((170, 20), (23, 5), (20, 20), (7, 10)):
        np.random.seed(n * 100 + b)
        idx, mbs = mdbn_b200.get_minibatches_idx(n, b, shuffle=True)
        assert [len(m) for m in mbs] == list(g["n%d_b%d_lens" % (n, b)])
        np.testing.assert_array_equal(np.concatenate(mbs), g["n%d_b%d" % (n, b)])
        assert len(idx) == len(mbs)

"""CPU: the C-ABI library loads and exports every symbol include/riser_b200.h declares;
the host mirrors keep the reference's constants and error behaviour; the product path
refuses to run without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from riser_b200 import _lib, Kit, SignalProcessor, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def test_header_symbols_all_exported(lib):
    header = open(os.path.join(ROOT, "include", "riser_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(riser_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.riser_version() >= 100
    assert lib.riser_normalise_max_len() >= 16000


def test_host_mirror_constants_and_errors():
    p2 = SignalProcessor(Kit.create_from_version("RNA002"))
    p4 = SignalProcessor(Kit.create_from_version("RNA004"))
    assert (p2.get_min_length(), p2.get_max_length(), p2.get_fixed_trim_length()) == (4096, 12048, 6480)
    assert (p4.get_min_length(), p4.get_max_length(), p4.get_fixed_trim_length()) == (4096, 8615, 4633)
    assert p2.should_trim_fixed_length(np.zeros(18529)) and not p2.should_trim_fixed_length(np.zeros(18528))
    assert p2.is_max_length(np.zeros(12048)) and not p2.is_max_length(np.zeros(12047))
    assert len(p2.trim_polyA_fixed_length(np.zeros(7000))) == 520
    with pytest.raises(Exception):
        Kit.create_from_version("RNA003")
    with pytest.raises(ValueError):
        p2.mad_normalise(np.array([], dtype=np.int16))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    p = SignalProcessor(Kit.create_from_version("RNA002"))
    with pytest.raises(_lib.RiserError):
        p.mad_normalise(np.arange(5000, dtype=np.int16))
    with pytest.raises(_lib.RiserError):
        p.get_polyA_end(np.arange(5000, dtype=np.int16))


def test_product_package_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "riser_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f

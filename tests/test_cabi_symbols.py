"""CPU: the C-ABI library builds, loads, and exports every symbol include/seir_b200.h declares."""
import os
import re

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import build

    build.build()
    lib = nat.load()
    header = open(os.path.join(ROOT, "include", "seir_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(seir_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in seir_b200.h but not exported"
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
    assert lib.seir_abi_version() == nat.ABI_VERSION


def test_bad_arguments_return_errors_without_gpu():
    from covid19uk_b200 import _native as nat

    lib = nat.load()
    assert lib.seir_model_create(None, 0, None) == -1
    assert b"NULL" in lib.seir_last_error()
    assert lib.seir_model_dims(None, None, None, None, None) == -1

"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/tf21.h declares; the ctypes table covers exactly the header; no compute is attempted."""
import ctypes
import importlib
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tf21.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tf21_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    tf = importlib.import_module("twenty-first_b200")
    lib = ctypes.CDLL(tf.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 35
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/tf21.h but not exported by libtf21.so"


def test_ctypes_table_matches_header():
    tf = importlib.import_module("twenty-first_b200")
    assert sorted(tf.SIGNATURES) == header_symbols()


def test_strerror_and_error_codes_need_no_gpu():
    tf = importlib.import_module("twenty-first_b200")
    header = open(os.path.join(ROOT, "include", "tf21.h")).read()
    for name, code in (("TF21_E_LEN_NOT_POW2", tf.E_LEN_NOT_POW2), ("TF21_E_LEN_TOO_LARGE", tf.E_LEN_TOO_LARGE),
                       ("TF21_E_TOO_FEW_LEAFS", tf.E_TOO_FEW_LEAFS),
                       ("TF21_E_INCORRECT_NUMBER_OF_LEAFS", tf.E_INCORRECT_NUMBER_OF_LEAFS),
                       ("TF21_E_ORDER_LE_DEGREE", tf.E_ORDER_LE_DEGREE), ("TF21_E_ALLOC", tf.E_ALLOC),
                       ("TF21_E_CUDA", tf.E_CUDA), ("TF21_E_BAD_ARG", tf.E_BAD_ARG)):
        assert re.search(rf"{name}\s*=\s*{code}\b", header), name
        assert tf.lib.tf21_strerror(code)
    assert b"power of two" in tf.lib.tf21_strerror(tf.E_LEN_NOT_POW2)
    # argument validation happens before any CUDA call
    assert tf.lib.tf21_ntt(None, 12, 1, 1) == tf.E_LEN_NOT_POW2
    assert tf.lib.tf21_ntt(None, 1 << 33, 1, 1) == tf.E_LEN_TOO_LARGE
    assert tf.lib.tf21_ntt(None, 8, 2, 1) == tf.E_BAD_ARG
    assert tf.lib.tf21_merkle_build(None, 0, None) == tf.E_TOO_FEW_LEAFS
    assert tf.lib.tf21_merkle_build(None, 12, None) == tf.E_INCORRECT_NUMBER_OF_LEAFS


def test_no_cpu_fallback_without_a_device():
    """The product path must fail loudly, not fall back: on a box without a GPU a compute call
    returns TF21_E_CUDA (this test is skipped where a GPU exists)."""
    import numpy as np
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    tf = importlib.import_module("twenty-first_b200")
    x = np.arange(8, dtype=np.uint64)
    rc = tf.lib.tf21_ntt(x.ctypes.data, 8, 1, 1)
    assert rc in (tf.E_CUDA, tf.E_ALLOC)
    assert (x == np.arange(8, dtype=np.uint64)).all()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "twenty-first_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "oracle/" not in src, f


def test_header_is_plain_c(tmp_path):
    """include/tf21.h must be consumable by a C compiler (the Rust/cgo/JNI side binds plain C)."""
    import subprocess

    src = tmp_path / "t.c"
    src.write_text('#include "tf21.h"\nint main(void) { return tf21_strerror(TF21_E_CUDA) == 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", f"-I{inc}", str(src)],
                   check=True)


def test_rust_sys_crate_declares_exactly_the_header():
    """host/rust/tf21-sys/src/lib.rs is generated from include/tf21.h + the ctypes table (tools/gen_rust_sys.py);
    the committed file must be up to date, so the Rust shim binds what the library exports"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(ROOT, "tools", "gen_rust_sys.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(ROOT, "host", "rust", "tf21-sys", "src", "lib.rs")) as f:
        assert f.read() == mod.generate()

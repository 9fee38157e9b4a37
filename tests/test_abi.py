"""The C-ABI shared library: loads without a GPU, exports every symbol include/vitk.h declares, its argument
struct matches the ctypes mirror, and argument validation fails loudly through the error slot (no compute
call is made here: every call below is rejected before any CUDA work)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vitk.h")


@pytest.fixture(scope="module")
def lib():
    from vision_transformers_torch_xla_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return _lib.load()


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vitk_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from vision_transformers_torch_xla_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 19
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/vitk.h but not exported by libvitk.so"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "python binding and header disagree on the symbol list"


def test_abi_version_and_arch(lib):
    from vision_transformers_torch_xla_b200 import _lib

    src = open(HEADER).read()
    assert lib.vitk_abi_version() == _lib.ABI_VERSION == int(re.search(r"#define VITK_ABI_VERSION (\d+)", src).group(1))
    assert lib.vitk_arch() == b"sm_100a"


def test_build_id_ties_the_binary_to_the_sources(lib, monkeypatch, tmp_path):
    """libvitk.so carries the sha256 of the sources it was built from; load() refuses a binary that does not match the
    sources next to it (a stale .so shipped with newer sources would otherwise go unnoticed)."""
    from vision_transformers_torch_xla_b200 import _lib

    lib.vitk_build_id.restype = ctypes.c_char_p
    assert lib.vitk_build_id().decode() == _lib.source_build_id()
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "source_build_id", lambda: "0" * 16)
    with pytest.raises(_lib.VitkError, match="stale"):
        _lib.load()


def test_header_is_plain_c_and_struct_layout_matches_ctypes(tmp_path):
    """include/vitk.h must compile as C (it is the FFI contract) and agree with _lib.GemmArgs."""
    from vision_transformers_torch_xla_b200._lib import GemmArgs

    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vitk.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(vitk_gemm_args), offsetof(vitk_gemm_args, M),'
                   ' offsetof(vitk_gemm_args, out), offsetof(vitk_gemm_args, rowscale), offsetof(vitk_gemm_args, block_n));'
                   'return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    size, off_m, off_out, off_rs, off_bn = map(int, subprocess.run([str(exe)], check=True, capture_output=True,
                                                                   text=True).stdout.split())
    assert size == ctypes.sizeof(GemmArgs)
    assert off_m == GemmArgs.M.offset and off_out == GemmArgs.out.offset
    assert off_rs == GemmArgs.rowscale.offset and off_bn == GemmArgs.block_n.offset


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG must be in the shipped binary."""
    from vision_transformers_torch_xla_b200 import _lib

    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR"):
        assert mnemonic in out.stdout, mnemonic
    assert "HMMA.16816" not in out.stdout  # no legacy mma.sync tensor path


def test_argument_validation_reports_through_error_slot(lib):
    from vision_transformers_torch_xla_b200._lib import GemmArgs

    args = GemmArgs()
    args.M, args.N, args.K = 0, 16, 64
    rc = lib.vitk_gemm_bf16(ctypes.byref(args), None)
    assert rc < 0 and b"empty problem" in lib.vitk_last_error()
    assert lib.vitk_gemm_bf16(None, None) < 0
    rc = lib.vitk_attn_fwd(None, None, None, 1, 16, 1, 96, ctypes.c_float(1.0), None)
    assert rc == -6 and b"head_dim" in lib.vitk_last_error()  # VITK_STATUS_UNSUPPORTED (wider than a head tile + its tail)
    rc = lib.vitk_attn_fwd(None, None, None, 1, 300, 1, 72, ctypes.c_float(1.0), None)
    assert rc == -6 and b"N <= 256" in lib.vitk_last_error()   # heads of 72 / 80: sequences of at most 256 tokens
    rc = lib.vitk_layernorm_fwd(None, 770, None, None, None, 770, None, None, 4, 770, ctypes.c_float(1e-6), None)
    assert rc < 0 and b"multiple of 4" in lib.vitk_last_error()
    rc = lib.vitk_adamw_flat(None, None, None, None, None, None, None, None, 6, None, 64, 1, None, None, ctypes.c_float(0.9),
                             ctypes.c_float(0.999), ctypes.c_float(1e-8), 1, ctypes.c_float(1.0), ctypes.c_float(0.0), 0, None)
    assert rc < 0
    probs = (ctypes.c_float * 2)(0.1, 1.5)
    rc = lib.vitk_droppath_masks(ctypes.c_void_p(16), probs, 2, 4, 0, 0, None)
    assert rc < 0 and b"drop_probs" in lib.vitk_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from vision_transformers_torch_xla_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.VitkError, match="no CPU/PyTorch fallback"):
        _lib.load()

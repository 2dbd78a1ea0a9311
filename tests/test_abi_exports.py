"""CPU: the C-ABI library loads and exports every symbol include/vi_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import vectorindex as vi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(vi_[a-z0-9_]+)\s*\(", text))
    names -= {"vi_allreduce_u64_fn"}
    return sorted(names)


def test_header_and_binding_agree():
    assert _declared() == sorted(vi.EXPORTS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(vi.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(vi.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.vi_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        return
    try:
        vi.Context(0)
    except vi.VectorIndexError as e:
        assert e.code == vi.VI_ERR_CUDA
    else:
        raise AssertionError("vi_create must fail without a CUDA device")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vector-database_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".cs")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                # comments may cite the oracle as the CPU statement of a rule; code must not load, link or include it
                assert "import oracle" not in src and "from oracle" not in src and "libvi_oracle" not in src, f
                assert '#include "vi_oracle' not in src, f
                if f == "vi_comm.cu":
                    # the one place that loads a library at run time: NCCL, and nothing else
                    import re
                    names = re.findall(r'"([^"]*\.so[^"]*)"', src)
                    assert names and all("nccl" in n for n in names), names
                else:
                    assert "dlopen" not in src, f

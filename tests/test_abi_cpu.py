"""CPU checks of the drop-in boundary: the library loads without a GPU, exports every symbol the
header declares, and its struct layouts match the ctypes mirror.  No compute call is made."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    import offsim4rl_b200
    return offsim4rl_b200._abi.load()


def _header_functions():
    text = open(os.path.join(ROOT, "include", "ols_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ols_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    import offsim4rl_b200
    declared = _header_functions()
    assert len(declared) >= 19
    assert sorted(offsim4rl_b200._abi.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layouts_match(lib):
    import offsim4rl_b200
    for k, st in enumerate(offsim4rl_b200._abi.STRUCTS):
        assert lib.ols_struct_bytes(k) == ctypes.sizeof(st)
    assert lib.ols_abi_version() == offsim4rl_b200._abi.ABI_VERSION


def test_pure_host_queries(lib):
    assert lib.ols_record_bytes(4) == 32 and lib.ols_record_bytes(2) == 32
    assert lib.ols_record_bytes(5) == 48 and lib.ols_record_bytes(16) == 80
    assert lib.ols_record_bytes(17) == 0
    assert lib.ols_bucket_scratch_bytes(10**8, 10**4) > 0
    assert lib.ols_bucket_scratch_bytes(1000, 10**6) == 0          # one histogram per warp must fit in smem
    assert lib.ols_curve_scratch_bytes(1000, 10) >= 4 * 3 * 10 * 8


def test_argument_errors_are_reported_not_thrown(lib):
    import offsim4rl_b200
    abi = offsim4rl_b200._abi
    rc = lib.ols_encode_grid(None, 10, 2, 5, None, None)
    assert rc == abi.ERR_INVALID
    assert b"null" in lib.ols_last_error()
    with pytest.raises(abi.OlsError):
        abi.check(rc)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import offsim4rl_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        offsim4rl_b200.BucketedLog(z=np.zeros(4), z_next=np.zeros(4), action=np.zeros(4), reward=np.zeros(4),
                                   terminal=np.zeros(4), p_log=np.full((4, 2), 0.5))

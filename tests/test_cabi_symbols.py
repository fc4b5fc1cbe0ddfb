"""CPU: the C-ABI library loads and exports every symbol include/mogp_b200.h declares
(no compute call is made without a GPU)."""
import ctypes as C

from mogptk_b200 import _cabi


def test_exports_match_header(lib):
    names = _cabi.header_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert set(_cabi._SIGNATURES) == set(names)


def test_version_and_param_counts(lib):
    assert lib.mogp_version() >= 1000
    assert lib.mogp_num_params(0, 4, 5, 1) == 4 * 5 * 5          # MOSM cfg2: 100 kernel parameters
    assert lib.mogp_num_params(1, 1, 3, 1) == 9                  # SM cfg1
    assert lib.mogp_num_params(2, 4, 1, 1) == 9                  # CONV cfg4
    assert lib.mogp_num_params(0, 0, 1, 1) < 0
    assert lib.mogp_num_params(7, 1, 1, 1) < 0


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    assert lib.mogp_create(0, 128, C.byref(h)) != 0          # fails loudly: no device, no fallback
    import pytest
    from mogptk_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(device=0, max_n=128)

"""CPU tier: the C-ABI library loads and exports every symbol include/lmpc_b200.h declares, and the
product path fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import pytest
import torch
from racinglmpc_b200 import _native as nat, reference_params as rp


def test_library_exports_every_declared_symbol():
    L = nat.lib()
    names = nat.exported_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), n


def test_struct_layout_matches_header():
    # sizeof(lmpc_params) as compiled == ctypes mirror (guards against silent ABI drift)
    L = nat.lib()
    if hasattr(L, "lmpc_sizeof_params"):
        L.lmpc_sizeof_params.restype = C.c_int
        assert L.lmpc_sizeof_params() == C.sizeof(nat.Params)
        L.lmpc_sizeof_model_params.restype = C.c_int
        assert L.lmpc_sizeof_model_params() == C.sizeof(nat.ModelParams)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback():
    from racinglmpc_b200 import BatchedFTOCP
    with pytest.raises(nat.NativeError):
        BatchedFTOCP(rp.mpc_params(12), batch=4)

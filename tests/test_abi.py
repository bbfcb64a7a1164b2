"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/sphe.h
declares, refuses to compute without a GPU (no CPU fallback), and host-only calls work."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, product


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sphe.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sphe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    m = product()
    L = m.capi.lib()
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert sorted(m.capi.SYMBOLS) == names, "capi.SYMBOLS out of sync with include/sphe.h"
    assert L.sphe_abi_version() == 1


def test_struct_layouts_match_reference_particle():
    m = product()
    assert C.sizeof(m.capi.Particle) == 112  # sizeof(FluidParticle), SURVEY a1
    assert C.sizeof(m.capi.Params) == 12 * 4


def test_host_only_calls_need_no_gpu():
    m = product()
    s = m.FluidSystemSPH()  # constructor does no CUDA work (global object in main.cpp:50)
    assert s.GetDeltaTime() == 0.0
    s.SetDeltaTime(0.01)
    assert abs(s.GetDeltaTime() - 0.01) < 1e-9
    s.SetOrigin((1.0, 2.0, 3.0))
    assert s.GetOrigin().tolist() == [1.0, 2.0, 3.0]
    p = s.params
    assert (round(p.mass, 4), round(p.visc, 3), round(p.p0, 2), round(p.h, 4), round(p.len, 3)) == (0.02, 3.5, 998.29, 0.0457, 0.2)
    p.visc = 4.0  # ImGui-style write-through (main.cpp:278-290)
    assert s.params.visc == 4.0
    assert s.count() == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = product()
    s = m.FluidSystemSPH()
    with pytest.raises(m.capi.SpheError, match="no CPU fallback"):
        s.Initialize(1000)
    with pytest.raises(m.capi.SpheError):
        s.upload_state([[0.0, 0.0, 0.0]], [[0.0, 0.0, 0.0]])

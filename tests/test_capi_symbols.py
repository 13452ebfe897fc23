"""The C-ABI library loads and exports every symbol include/fluxb200.h declares (no compute calls:
this runs without a GPU).  Also checks the no-CPU-fallback rule."""
import ctypes
import os
import re

import pytest

from flux_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "fluxb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flux_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_capi.PROTOTYPES)


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()  # raises FluxLibraryMissing if the .so was not built
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.flux_version().startswith(b"fluxb200")


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_capi.flux_material) == 48
    assert ctypes.sizeof(_capi.flux_job_config) == 12
    assert ctypes.sizeof(_capi.flux_counters) == 8 * len(_capi.COUNTER_FIELDS) == 160
    # 2 u32 + f64 + 3*4 f64-triples... : compare with the C compiler's view through the shard helper instead
    n = ctypes.c_uint32()
    assert _capi.lib().flux_shard_rows(600, 4, 1, 8, None, ctypes.byref(n)) == 0 and n.value == 76
    rows = (ctypes.c_uint32 * n.value)()
    assert _capi.lib().flux_shard_rows(600, 4, 1, 8, rows, ctypes.byref(n)) == 0
    assert list(rows[:6]) == [4, 5, 6, 7, 36, 37]
    assert _capi.lib().flux_shard_rows(600, 0, 0, 8, None, ctypes.byref(n)) == _capi.FLUX_ERR_INVALID


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product refuses to create a context (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from flux_b200.worker import FluxError, GpuContext
    with pytest.raises(FluxError) as e:
        GpuContext(0)
    assert e.value.code == _capi.FLUX_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flux_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in src and "liboracle" not in src and "from oracle" not in src, f


def test_ppm_writer_matches_image_rs(tmp_path):
    import numpy as np
    rgb = np.array([[[0.0, 0.5, 1.0], [2.0, float("nan"), -1.0]]], np.float64)
    p = tmp_path / "x.ppm"
    assert _capi.lib().flux_write_ppm(str(p).encode(), 2, 1, _capi.as_dp(rgb)) == 0
    assert p.read_text() == "P3\n2 1\n65535\n0 32767 65535\n65535 0 0\n"  # image.rs:45-52


def test_frame_entry_points_refuse_null_arguments_without_a_device():
    """flux_frame_* / flux_ctx_sync check their arguments before they touch CUDA (no GPU here)."""
    lib = _capi.lib()
    out = ctypes.c_void_p()
    assert lib.flux_frame_create(None, 8, 8, ctypes.byref(out)) == _capi.FLUX_ERR_INVALID
    assert lib.flux_frame_open_ipc(None, bytes(64), 8, 8, ctypes.byref(out)) == _capi.FLUX_ERR_INVALID
    assert lib.flux_frame_open_peer(None, None, ctypes.byref(out)) == _capi.FLUX_ERR_INVALID
    assert lib.flux_frame_export(None, ctypes.create_string_buffer(64)) == _capi.FLUX_ERR_INVALID
    assert lib.flux_frame_read(None, None) == _capi.FLUX_ERR_INVALID
    assert lib.flux_frame_device_ptr(None, ctypes.byref(out)) == _capi.FLUX_ERR_INVALID
    assert lib.flux_frame_close(None) == _capi.FLUX_ERR_INVALID
    assert lib.flux_ctx_sync(None) == _capi.FLUX_ERR_INVALID
    assert lib.flux_render_row_list_into_frame(None, None, 0, None, None) == _capi.FLUX_ERR_INVALID

"""The C-ABI library loads without a GPU and exports every symbol include/nbody_b200.h declares; the
host-arithmetic entry points work; the product path fails loudly (no CPU fallback) and never
touches oracle/."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "nbody_b200.h"
PKG = ROOT / "nbody-gnn-hpc_b200"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(nbh?_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for required in ("nb_accel_f32", "nb_accel_f64", "nb_step_f32", "nb_step_f64", "nb_run_f64", "nb_ensemble_f64",
                     "nb_ensemble_f32", "nb_energy_f64", "nb_pack_f32", "nbh_accel_direct", "nbh_run",
                     "nbh_ensemble_run", "nbh_total_energy", "nb_last_error", "nb_window_gather_f32", "nb_window_count",
                     "nb_step_peer_f32", "nb_step_peer_f64"):
        assert required in names
    text = HEADER.read_text()
    assert "torch" not in text.lower().replace("pytorch", "")     # plain pointers and sizes only
    assert 'extern "C"' in text


def test_library_exports_every_declared_symbol():
    from hpc import _cuda
    lib = _cuda.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    # and the Python binding declares a signature for each of them
    assert set(declared_functions()) == set(_cuda.exported_symbols())
    assert lib.nb_abi_version() == 1


def test_planning_entry_points_run_without_gpu():
    from hpc import _cuda
    lib = _cuda.load_library()
    assert lib.nb_padded_bodies(1) == 32 and lib.nb_padded_bodies(200) == 224 and lib.nb_padded_bodies(65536) == 65536
    for n in (1, 200, 1023, 16384, 65536, 262144, 1048576, 1000003):
        seg, nseg = ctypes.c_int(), ctypes.c_int()
        assert lib.nb_segment_plan(n, seg, nseg) == 0
        n_pad = lib.nb_padded_bodies(n)
        assert seg.value % 32 == 0 and seg.value * nseg.value >= n_pad > seg.value * (nseg.value - 1)
        assert nseg.value <= 64
        assert lib.nb_workspace_bytes(n, n, 1) >= nseg.value * 3 * n * 8
        assert lib.nb_workspace_bytes(n, n, 0) >= nseg.value * 3 * n * 4
    assert lib.nb_ensemble_max_bodies() >= 512
    assert lib.nb_ensemble_workspace_bytes(300) >= 301 * 4


def test_bad_arguments_return_codes_not_crashes():
    from hpc import _cuda
    lib = _cuda.load_library()
    rc = lib.nb_accel_f64(None, 0, 0, 0, 1e-9, None, None, 0, None)
    assert rc == 1 and lib.nb_last_error()
    rc = lib.nb_ensemble_f64(None, None, None, None, 0, 0, 1, 1, 1e-3, 1e-9, 1, 1, 1, 1, None, None, None, 1, 0,
                             None, 0, None)
    assert rc == 1 and b"null" in lib.nb_last_error()
    rc = lib.nb_window_gather_f32(None, None, 1, 1, 1, 1, 1, 1, None, None, None)
    assert rc == 1 and b"null" in lib.nb_last_error()
    import ctypes
    buf = ctypes.create_string_buffer(64)
    rc = lib.nb_window_gather_f32(buf, buf, 1, 4, 2, 9, 2, 1, buf, buf, None)      # n_states > rows
    assert rc == 1 and b"n_states" in lib.nb_last_error()
    rc = lib.nb_window_gather_f32(buf, buf, 1, 4, 2, 4, 0, 1, buf, buf, None)      # sequence_length < 1
    assert rc == 1 and b"sequence_length" in lib.nb_last_error()


def test_no_cpu_fallback_and_no_oracle_in_product():
    """Without a CUDA device the engine raises; nothing under the package imports the oracle."""
    import torch
    from hpc import _cuda, nbody
    for path in PKG.rglob("*.py"):
        src = path.read_text()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f"{path} imports the oracle"
    for path in list(PKG.rglob("*.cu")) + list(PKG.rglob("*.cuh")):
        assert "oracle" not in path.read_text().lower(), f"{path} mentions the oracle"
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the failure path cannot be shown here")
    import numpy as np
    with pytest.raises(_cuda.EngineUnavailable):
        nbody.compute_accelerations_direct(np.zeros((4, 3)), np.ones(4))
    with pytest.raises(_cuda.EngineUnavailable):
        nbody.NBodySimulator(n_particles=8, seed=1)


def test_missing_library_is_a_loud_error(monkeypatch, tmp_path):
    from hpc import _cuda
    monkeypatch.setattr(_cuda, "_lib", None)
    monkeypatch.setenv("NBODY_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_cuda.EngineUnavailable, match="no CPU fallback"):
        _cuda.load_library()


def test_bad_arguments_of_the_round2_entry_points():
    """nb_snapshot_energy_f64, nb_step_peer_*, nb_step_status, nb_probe_occupy: argument errors are return codes with a
    message (checked before any CUDA call, so this runs without a GPU)."""
    import ctypes
    from hpc import _cuda
    lib = _cuda.load_library()
    buf = ctypes.create_string_buffer(4096)
    assert lib.nb_snapshot_energy_max_bodies() >= 1024 and lib.nb_persist_max_bodies() >= 16384
    rc = lib.nb_snapshot_energy_f64(buf, None, buf, 0, 0, 1, 1, 4, 6.6743e-11, 1e-9, buf, None)      # vel null
    assert rc == 1 and b"null" in lib.nb_last_error()
    rc = lib.nb_snapshot_energy_f64(buf, buf, buf, 0, 0, 1, 1, 10 ** 6, 6.6743e-11, 1e-9, buf, None)  # N too large
    assert rc == 1 and b"N <=" in lib.nb_last_error()
    rc = lib.nb_snapshot_energy_f64(buf, buf, buf, 0, 3, 2, 2, 4, 6.6743e-11, 1e-9, buf, None)         # stride not 0 / N
    assert rc == 1 and b"mass_stride" in lib.nb_last_error()
    ptrs = (ctypes.c_void_p * 2)(ctypes.addressof(buf), ctypes.addressof(buf))
    ws_bytes = lib.nb_workspace_bytes(64, 32, 1)
    rc = lib.nb_step_peer_f64(buf, ptrs, ptrs, 2, 5, 0, 1, buf, buf, 64, 0, 32, 1e-3, 1e-9, 0, None, None, None, buf,
                              ws_bytes, None)                                                          # my_rank >= n_ranks
    assert rc == 1 and b"my_rank" in lib.nb_last_error()
    rc = lib.nb_step_peer_f64(buf, ptrs, ptrs, 2, 1, 0, 1, buf, buf, 64, 16, 32, 1e-3, 1e-9, 0, None, None, None, buf,
                              ws_bytes, None)                                                          # slab start not a chunk
    assert rc == 1 and b"multiple of 32" in lib.nb_last_error()
    rc = lib.nb_step_peer_f64(buf, ptrs, ptrs, 2, 1, 0, 1, buf, buf, 64, 32, 32, 1e-3, 1e-9, 0, None, None, None, buf,
                              16, None)                                                                # workspace too small
    assert rc == 1 and b"workspace too small" in lib.nb_last_error()
    assert lib.nb_step_status(None, 64, None) == 1 and lib.nb_probe_occupy(0, 0, 1.0, None) == 1
    assert lib.nb_workspace_bytes(4096, 4096, 0) > lib.nb_workspace_bytes(4096, 2048, 0)   # whole-system workspaces carry K2p's words
    # batched mid-size ensembles
    assert lib.nb_batched_max_bodies() > lib.nb_ensemble_max_bodies()
    assert lib.nb_accel_batched_f64(None, 2, 2000, 1e-9, buf, None) == 1 and b"nb_accel_batched" in lib.nb_last_error()
    assert lib.nb_accel_batched_f32(buf, 2, 10 ** 6, 1e-9, buf, None) == 1                  # n beyond the batched limit
    rc = lib.nb_run_batched_f64(buf, buf, buf, buf, 2, 2000, 1e-3, 1e-9, 4, 0, None, None, None, None, None)
    assert rc == 1 and b"save_interval" in lib.nb_last_error()
    rc = lib.nb_run_batched_f32(buf, buf, buf, buf, 2, 2000, 1e-3, 1e-9, 4, 1, buf, None, None, None, None)
    assert rc == 1 and b"all set or all null" in lib.nb_last_error()

"""bench.py pieces that can be checked without a GPU: both arms print the same `config`, the source hash the ncu
digest is stamped with, the worker of the reference arm (the reference's own Numba code from baseline/_ref, when this
container has it), and the fallback to the C port."""
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_config_is_one_function_for_both_arms(bench):
    cfg = bench.ensemble_config()
    assert cfg == bench.ensemble_config() and "configs[1]" in cfg["workload"] and "l2" in cfg
    assert cfg["simulations_per_gpu"] == 300 and cfg["bodies"] == 200 and cfg["sim_steps"] == 400
    src = (ROOT / "bench.py").read_text()
    assert src.count('"config": ensemble_config()') == 1 and "cfg = ensemble_config()" in src   # b200 arm / reference arm


def test_source_hash_matches_the_digest_tools(bench):
    h = bench.source_hash()
    assert len(h) == 16 and int(h, 16) >= 0
    summary = json.loads((ROOT / "profiles" / "summary.json").read_text())
    assert "_source_hash" in summary and len(summary["_source_hash"]) == 16
    # the SASS histogram is headed by the hash of the sources it was taken from
    assert "CUDA sources hash" in (ROOT / "profiles" / "r02_sass_opcodes.md").read_text()


def test_reference_worker_runs_the_vendored_numba_module(bench):
    """One tiny simulation through the reference arm's worker (generate_data.py:32-58 call for call) in this process."""
    pytest.importorskip("numba")
    if not bench.REF_FILE.exists():
        pytest.skip("baseline/_ref/nbody.py not vendored (built where /root/reference is mounted)")
    bench._numba_worker_init(str(bench.REF_FILE))
    m32 = np.random.RandomState(42).uniform(1e10, 1e12, 12).astype(np.float32)
    out = bench._numba_single_simulation((0, 12, 5, 1, 10.0, 42, m32))
    assert out["positions"].shape == (6, 12, 3) and out["n_steps"] == 6 and out["masses"].dtype == np.float32
    sys.path.insert(0, str(ROOT))
    import oracle
    from hpc import ics
    x0, v0, _ = ics.reference_default_ic(12, 42)
    chk = oracle.run(x0, v0, oracle.accel_direct(x0, m32), m32, 1e-3, 1e-9, 5, 1)
    assert np.abs(out["positions"] - chk["positions"]).max() < 1e-9       # the port the fallback arm times agrees with it


def test_reference_engine_falls_back_to_the_port(bench, monkeypatch, tmp_path, capsys):
    monkeypatch.setattr(bench, "REF_FILE", tmp_path / "missing.py")
    kind, ref = bench.reference_engine()
    assert kind == "port" and ref is None
    assert "Numba reference unavailable" in capsys.readouterr().err

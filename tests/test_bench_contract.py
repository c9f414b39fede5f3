"""bench.py's reference arm and JSON contract, on a reduced row count (CPU only)."""
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line(monkeypatch):
    import bench
    monkeypatch.setattr(bench, "N_ROWS", 20_000)
    monkeypatch.setattr(bench, "make_data", lambda n=20_000, seed=0: bench.__dict__["_orig_make_data"](20_000, seed))
    args = types.SimpleNamespace(gpus=1, steps=1, warmup=0)
    buf = io.StringIO()
    with redirect_stdout(buf):
        assert bench.run_reference(args) == 0
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "estimates/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["metric"].startswith("KSG MI estimates/sec at N=10^6")
    assert d["config"] == bench.CONFIG            # the GPU arm prints the same dict: the driver compares the two
    assert "every" not in d["cpu_baseline"]["sample"].split("rows per step")[0] or "ALL" in d["cpu_baseline"]["sample"]


def test_other_ranks_of_the_reference_arm_do_nothing(monkeypatch):
    import bench
    monkeypatch.setenv("RANK", "3")
    buf = io.StringIO()
    with redirect_stdout(buf):
        assert bench.run_reference(types.SimpleNamespace(gpus=8, steps=1, warmup=0)) == 0
    assert buf.getvalue() == ""

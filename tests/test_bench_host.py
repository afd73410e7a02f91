"""Host logic of bench.py (no GPU): configuration table and the algorithmic-FLOP bookkeeping of BASELINE.md section 2."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _cfg(name, **over):
    ns = argparse.Namespace(config=name, batch=over.get("batch", 0), timesteps=over.get("timesteps", 0))
    return bench.resolve_config(ns)


def test_config_flops_match_baseline_md():
    c2 = _cfg("2")
    assert abs(c2["B"] * bench.alg_flops_per_image(c2) / 1e15 - 3.281) < 1e-3          # 3.281 PFLOP per GPU
    c4 = _cfg("4")
    assert abs(4 * c4["B"] * bench.alg_flops_per_image(c4) / 1e15 - 8.436) < 2e-3      # 128 images over 4 GPUs
    c5 = _cfg("5")
    assert abs(8 * c5["B"] * bench.alg_flops_per_image(c5) / 1e12 - 64.43) < 2e-2      # 1024 decodes over 8 GPUs
    assert _cfg("3cfg8")["passes"] == 2 and _cfg("3")["passes"] == 1 and _cfg("2")["passes"] == 1


def test_same_workload_at_every_gpu_count():
    """The per-GPU workload of a configuration does not depend on --gpus (N=1 and N>1 ran different
    workloads, so the scaling efficiency was not like-for-like)."""
    for name in bench.CONFIGS:
        c = _cfg(name)
        assert "conditional" in c and isinstance(c["B"], int)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "n_gpus > 1" not in src.split("def main()")[1].split("shard_parity")[0]


def test_overrides():
    c = _cfg("2", batch=8, timesteps=10)
    assert c["B"] == 8 and c["timesteps"] == 10
    assert _cfg("5", timesteps=10)["timesteps"] == 0     # decode-only ignores --timesteps

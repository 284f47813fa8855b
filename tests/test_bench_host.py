"""bench.py's host-side logic that needs no GPU: the queue depth rule, the two arms' shared workload description, and that
the product arm refuses to run without a CUDA device (there is no CPU fallback to time by accident)."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_queue_depth_rule():
    b = _bench()
    # 100k columns = 391 strips; per rank 391 / 196 / 98 / 49 strips = 131 / 66 / 33 / 17 blocks of three strips on 148 SMs
    assert [b.queue_depth(100000, w, 148) for w in (1, 2, 4, 8)] == [3, 3, 5, 9]
    assert b.queue_depth(10000, 1, 148) == 11           # config 2: 14 blocks per fill
    assert b.queue_depth(256, 8, 148) == 12             # never more than 12 plans, never a division by zero
    assert b.queue_depth(10 ** 6, 1, 148) == 3          # more blocks than SMs: still a queue


def test_both_arms_describe_the_same_workload():
    b = _bench()
    for n in (1, 2, 8):
        c = b.workload_config(n)
        assert (c["top_len"], c["side_len"], c["m"], c["k"], c["d"]) == (100000, 100000, 1, 1, 1)
        assert c["cells_per_step"] == 10 ** 10
        json.dumps(c)


def test_product_arm_needs_a_gpu(nwb):
    if nwb.device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-extras", "--no-cpu"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not r.stdout.strip().startswith("{")          # no bench line from a machine without a GPU

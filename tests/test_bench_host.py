"""CPU: host-side logic of bench.py (no GPU, no oracle run): the GPU arm and the reference arm of one workload must
print the SAME ``config`` object, input sets must exceed L2, every BASELINE config is a named workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_workload_table_covers_baseline_configs():
    import bench
    idx = sorted(v[6] for v in bench.WORKLOADS.values())
    assert idx == [1, 2, 3, 4]                      # configs[0] is the CPU reference arm itself
    for name, (_, H, W, T, l_t, B, _) in bench.WORKLOADS.items():
        assert H % 60 == 0 and W % 108 == 0 and 0 < l_t < T and B >= 1
        set_bytes = B * T * 3 * H * W * 4
        assert bench.n_input_sets(set_bytes) * set_bytes > bench.L2_BYTES, name
        cfg = bench.workload_config(name, B, 1)
        assert json.loads(json.dumps(cfg)) == cfg and "workload" in cfg and "model" not in cfg


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU arm; the others exit 0 without output or work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""

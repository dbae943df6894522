"""CPU: the reference arm of bench.py (`--impl reference`) prints the JSON line the driver expects, times the oracle
port on the host cores, and stays silent on ranks other than 0 (it is launched under torch.distributed.run for N > 1).
Run on a small workload so that the test takes seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, args=()):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(env_extra or {})
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2_graphene",
           "--steps", "1", "--warmup", "1"] + list(args)
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)


def test_reference_arm_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                     # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "get_emb_eri_fp64_tflops" and d["unit"] == "TFLOP/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("c2_graphene")
    for k in ("kmesh", "nao", "naux", "neo", "nspin", "symmetry", "t_reversal_symm"):
        assert k in d["config"], k                    # the same config keys as our arm (driver: same_config)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "blocks" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, ["--gpus", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""

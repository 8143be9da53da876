"""Host-side logic of bench.py that needs no GPU: the reference arm (oracle on the host cores, measured, not
extrapolated, when the matrix fits), the headline detection and the committed golden files the parity checks use."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    return lines


def test_reference_arm_measures_the_workload_itself():
    (line,) = _run(["--impl", "reference", "--n", "3000", "--lowest", "4", "--gpus", "1", "--steps", "7", "--warmup", "2"])
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "s" and d["higher_is_better"] is False
    assert d["steps"] == 1 and d["warmup"] == 0 and d["steps_requested"] == 7       # honest about what was timed
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["sample_n"] == 3000 and cb["extrapolation_factor"] == 1.0
    assert abs(cb["value"] - d["value"]) < 1e-12 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert d["config"]["n"] == 3000 and d["config"]["parallelism"] == "row-block x1"


def test_reference_arm_other_ranks_stay_silent_and_config_follows_gpus():
    assert _run(["--impl", "reference", "--n", "2000", "--lowest", "3", "--gpus", "8"], {"RANK": "3", "WORLD_SIZE": "8"}) == []
    (line,) = _run(["--impl", "reference", "--n", "2000", "--lowest", "3", "--gpus", "8"], {"RANK": "0", "WORLD_SIZE": "8"})
    d = json.loads(line)
    assert d["n_gpus"] == 8 and d["config"]["parallelism"] == "row-block x8"        # same config as the GPU arm


def test_golden_files_of_the_parity_checks():
    sys.path.insert(0, ROOT)
    import bench
    g = bench.golden_full_size()
    assert g["n"] == 100000 and g["lowest"] == 16 and g["iters"] == 3 and g["trace_k"] == [32, 64, 128]
    ev = np.asarray(g["eigenvalues"])
    assert len(ev) == 16 and np.all(np.diff(ev) > 0) and abs(ev[0] - 1.0) < 1e-6
    assert len(g["eigenvectors"]) == 16 and max(g["residual_norms"]) < g["tolerance"]
    lr = json.load(open(os.path.join(ROOT, "tests", "golden", "longrun_n20k_oracle.json")))
    assert lr["iters"] == 20 and lr["trace_k"][:4] == [20, 40, 20, 40]
    w = bench.W()
    assert bench.is_headline(w) and not bench.is_headline(bench.W(n=20000)) and not bench.is_headline(bench.W(gev=True))
    assert "configs[2]" in bench.workload_config(w, 8)["workload"] and "8 GPU" in bench.workload_config(w, 8)["workload"]


def test_bench_parity_checker_accepts_the_golden_and_rejects_deviations():
    """bench.py's `parity_check` is what turns a wrong result at any N into a non-zero exit: feed it a synthetic
    result built from the golden file (pass, also with flipped eigenvector signs), then break one thing at a time."""
    import types
    sys.path.insert(0, ROOT)
    import bench
    g = bench.golden_full_size()
    n, L = g["n"], g["lowest"]
    probes = np.asarray(g["probe_rows"])
    vec = np.zeros((n, L))
    for j, fp in enumerate(g["eigenvectors"]):
        vec[probes, j] = fp["probes"]
        vec[fp["imax"], j] = fp["vmax"]
        spare = next(i for i in range(n) if i != fp["imax"] and i not in set(probes.tolist()))
        rest = fp["norm"] ** 2 - float((vec[:, j] ** 2).sum())
        assert rest >= 0
        vec[spare, j] = np.sqrt(rest)
        if j % 2:
            vec[:, j] *= -1.0                      # eigenvectors are defined up to sign
    trace = g["trace_k"]
    st = types.SimpleNamespace(trace_k=trace + [0] * 8, trace_len=len(trace))
    cx = types.SimpleNamespace(np=np, world=4)
    w = bench.W()

    def check(ev=None, iters=None, vec_=None, max_res=1e-9, trace_=None):
        s = st if trace_ is None else types.SimpleNamespace(trace_k=trace_ + [0] * 8, trace_len=len(trace_))
        res = {"ev": np.asarray(g["eigenvalues"]) if ev is None else ev, "iters": g["iters"] if iters is None else iters,
               "st": s}
        return bench.golden_parity(cx, w, res, max_res, vec if vec_ is None else vec_)

    ok = check()
    assert ok["pass"] and ok["n_gpus"] == 4 and ok["eigenvector_probe_max_abs_err"] < 1e-12
    ev = np.asarray(g["eigenvalues"]).copy(); ev[7] *= 1 + 1e-9
    assert not check(ev=ev)["pass"]                                   # eigenvalue off by 1e-9 relative
    assert not check(iters=g["iters"] + 1)["pass"]                    # full-size run: iteration count must be equal
    assert not check(trace_=[32, 64, 32])["pass"]                     # another basis schedule
    assert not check(max_res=2e-8)["pass"]                            # residual above the tolerance
    bad = vec.copy(); bad[probes[3], 5] += 1e-7
    assert not check(vec_=bad)["pass"]                                # one eigenvector entry off by 1e-7

"""Generates tests/golden/config2_n100k_oracle.json: ONE full-size run of the CPU oracle on BASELINE.json configs[2]
(generate_diagonal_dominant(100000, 1e-4), seed 0, lowest 16, DPR, tol 1e-8, default max_dim_sub = 160).

The matrix takes 80 GB of host RAM and the solve ~2-3 minutes on the GPU box's host cores (k DGEMVs per iteration
stream the matrix 224 times), so this runs once on a GPU box host (it needs no GPU):

    python tests/golden/make_golden_n100k.py gpurun_out/config2_n100k_oracle.json

and the result is committed; bench.py and tests/test_gpu_r02.py compare the CUDA path against it at every N.
Stored per eigenvector: the 2-norm, the entry of largest magnitude with its index, a handful of fixed probe entries
and a position-weighted checksum -- enough to pin the vectors up to sign without committing 12.8 MB."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402

PROBES = [0, 1, 2, 3, 7, 15, 16, 31, 63, 100, 1000, 4095, 4096, 12543, 12544, 25000, 49999, 50000, 75000, 99999]


def fingerprint(vec):
    n, L = vec.shape
    w = np.cos(np.arange(n, dtype=np.float64) * 0.001) + 2.0  # position weights (positive, non-constant)
    out = []
    for j in range(L):
        v = vec[:, j]
        imax = int(np.argmax(np.abs(v)))
        sg = 1.0 if v[imax] >= 0 else -1.0  # sign convention: largest entry positive
        v = sg * v
        out.append({"norm": float(np.linalg.norm(v)), "imax": imax, "vmax": float(v[imax]),
                    "probes": [float(v[i]) for i in PROBES if i < n], "checksum": float(w @ v),
                    "abs_checksum": float(w @ np.abs(v))})
    return out


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "config2_n100k_oracle.json")
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    L, tol = 16, 1e-8
    avail_gb = 0.0
    for ln in open("/proc/meminfo"):
        if ln.startswith("MemAvailable"):
            avail_gb = float(ln.split()[1]) * 1e-6
    need_gb = 8e-9 * n * n * 1.05 + 2.0
    if avail_gb < need_gb:
        raise SystemExit("host RAM: need %.0f GB, %.0f GB available" % (need_gb, avail_gb))
    cores = os.cpu_count() or 1
    orc.set_num_threads(cores)
    t0 = time.perf_counter()
    A = orc.generate_diagonal_dominant(n, 1e-4, None, 0)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = orc.generalized_eigensolver(A, L, "DPR", 1000, tol, None)
    t_solve = time.perf_counter() - t0
    res = [float(np.linalg.norm(A @ r.eigenvectors[:, j] - r.eigenvalues[j] * r.eigenvectors[:, j])) for j in range(L)]
    doc = {"what": "oracle (oracle/davidson_oracle.cpp + scipy OpenBLAS LAPACK) on BASELINE.json configs[2]",
           "n": n, "lowest": L, "method": "DPR", "tolerance": tol, "max_dim_sub": 10 * L, "sparsity": 1e-4, "seed": 0,
           "iters": int(r.iters), "trace_k": [int(k) for k in r.trace_k], "trace_err": [float(e) for e in r.trace_err],
           "eigenvalues": [float(x) for x in r.eigenvalues], "residual_norms": res, "probe_rows": PROBES,
           "eigenvectors": fingerprint(r.eigenvectors),
           "host": {"cores": cores, "generate_s": t_gen, "solve_s": t_solve, "mem_available_gb": avail_gb}}
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    json.dump(doc, open(out_path, "w"), indent=1)
    print(json.dumps({k: doc[k] for k in ("n", "iters", "trace_k", "trace_err", "host")}))


if __name__ == "__main__":
    main()

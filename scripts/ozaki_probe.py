"""VERDICT r1 item 9 (stretch, gated): can the b = 32 / 64 block matvec go BELOW the FP64 pipe floor by an Ozaki-style
error-free split into INT8 slices (tcgen05 kind::i8, int32 accumulation in TMEM)?  This probe answers the numerical
half of the gate on the CPU, exactly (integer arithmetic emulated with int64, every partial sum checked against the
int32 range): relative error of W = A X against an 80-bit reference, on generate_diagonal_dominant data, as a function
of the slice width and of how many slice-pair levels are multiplied.

    python scripts/ozaki_probe.py [n] [b]          (n = 2048, b = 32 by default; ~1 minute)

Scheme.  A = D + E (the diagonal 1..n would cost ~5 leading slices under row scaling; D X is an exact-to-rounding
elementwise FP64 product, E = 1e-4 U(0,1) is what gets split).  Row i of E:  E_ij = 2^e_i sum_s q_s(i,j) 2^(-w(s+1)),
|q_s| < 2^w, column j of X likewise with 2^f_j.  Then  E X = sum_{s,t} 2^(e_i + f_j - w(s+t+2)) (Q_s^E Q_t^X)  with
every Q_s^E Q_t^X an exact integer GEMM (|sum| <= n 2^(2w) must stay below 2^31: n <= 131,072 for w = 7).  Keeping
the pairs with s + t <= T truncates at relative 2^(-w(T+1)).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402  (test infrastructure: only the generator is used)


def split(M, w, nslices, axis):
    """Error-free slices of M scaled per row (axis=1) or per column (axis=0): returns (exponents, [int64 slices])."""
    amax = np.abs(M).max(axis=axis, keepdims=True)
    e = np.where(amax > 0, np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0.0)  # |M| 2^-e < 1/2: q fits w bits signed
    r = M * np.exp2(-e)
    out = []
    for _ in range(nslices):
        r = r * (1 << w)
        q = np.trunc(r)
        out.append(q.astype(np.int64))
        r = r - q
    return e, out, np.abs(r).max()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    A = orc.generate_diagonal_dominant(n, 1e-4, None, 0)
    rng = np.random.default_rng(0)
    X, _ = np.linalg.qr(rng.standard_normal((n, b)))
    d = np.diag(A).copy()
    E = A - np.diag(d)
    ref = (A.astype(np.longdouble) @ X.astype(np.longdouble))          # 80-bit reference
    scale = np.abs(ref).max(axis=0)
    plain = A @ X
    rows = []
    doc = {"n": n, "b": b, "fp64_matmul_max_rel_err": float((np.abs(plain - ref) / scale).max())}
    for w in (6, 7):
        ns = -(-53 // w) + 1
        eE, QE, remE = split(E, w, ns, axis=1)
        eX, QX, remX = split(X, w, ns, axis=0)
        assert max(np.abs(q).max() for q in QE + QX) < (1 << w)
        worst_int = 0
        for T in range(3, 2 * ns - 1):
            acc = np.zeros((n, b), dtype=np.longdouble)
            pairs = 0
            for lvl in range(T, -1, -1):                      # smallest contributions first
                P = np.zeros((n, b), dtype=np.int64)
                for s in range(max(0, lvl - ns + 1), min(lvl, ns - 1) + 1):
                    t = lvl - s
                    prod = QE[s] @ QX[t]
                    worst_int = max(worst_int, int(np.abs(prod).max()))
                    P += prod
                    pairs += 1
                acc += P.astype(np.longdouble) * np.exp2(-float(w) * (lvl + 2))
            W = (acc * np.exp2(eE) * np.exp2(eX)).astype(np.float64) + d[:, None] * X
            err = float((np.abs(W.astype(np.longdouble) - ref) / scale).max())
            rows.append({"w": w, "levels_T": T, "int8_gemms": pairs, "max_rel_err": err})
            print("w=%d  s+t<=%2d  %3d INT8 GEMMs  max rel err %.2e" % (w, T, pairs, err), flush=True)
            if err < 2e-16:
                break
        doc["w%d_max_abs_int32_partial" % w] = worst_int
        doc["w%d_int32_ok_up_to_n" % w] = int((2 ** 31 - 1) // ((1 << w) - 1) ** 2)
    doc["rows"] = rows
    need = {w: min((r["int8_gemms"] for r in rows if r["w"] == w and r["max_rel_err"] <= 1e-13), default=None) for w in (6, 7)}
    doc["int8_gemms_for_1e-13"] = need
    print(json.dumps(doc)[:600])
    out = os.path.join(ROOT, "profiles", "r02_ozaki_probe_cpu.json")
    json.dump(doc, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()

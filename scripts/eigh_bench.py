"""Times the Rayleigh-Ritz eigensolver (csrc/trideig.cu vs the one-CTA Jacobi) on random symmetric matrices."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortran_davidson_b200 import lapack_wrapper as lw
rng = np.random.default_rng(0)
for k in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "32,64,128,160,256,512").split(",")]:
    a = rng.standard_normal((k, k)); S = (a + a.T) / 2
    w, v, info, ms = lw.sym_eigh_info(S, reps=5)
    ref = np.linalg.eigvalsh(S)
    print("k %4d  ms %s  accepted %d  orth_defect %.2e  resid %.2e  ev_err %.2e  orth %.2e" % (
        k, " ".join("%.3f" % m for m in ms), info["accepted"], info["orth_defect"], info["residual"],
        np.abs(w - ref).max(), np.abs(v.T @ v - np.eye(k)).max()), info["prof"] if any(info["prof"]) else "", flush=True)

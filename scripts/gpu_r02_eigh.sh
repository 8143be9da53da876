#!/bin/bash
# r02: Rayleigh-Ritz eigensolver A/B (register-resident tridiagonalisation, reflectors staged in shared memory)
O=gpurun_out; mkdir -p $O
K=${KS:-24,32,48,64,97,128,160,256}
{
echo "== new (register tridiagonalisation k<=128, reflectors in smem)"; python scripts/eigh_bench.py $K
echo "== new, tridiag path from k>=24"; DAV_EIGH_TRIDIAG_MIN_K=24 python scripts/eigh_bench.py 24,32,40
echo "== shared-memory tridiagonalisation (r01)"; DAV_TRIDIAG_REG=0 python scripts/eigh_bench.py $K
echo "== r01 kernels (DAV_TRIDIAG_REG=0 DAV_EIGVEC_VH_SMEM=0)"; DAV_TRIDIAG_REG=0 DAV_EIGVEC_VH_SMEM=0 python scripts/eigh_bench.py $K
} 2>&1 | tee $O/r02_eigh_bench_${TAG:-v6}.txt
timeout 900 python -m pytest tests -x -q -m gpu -k "sym_eigh or lapack_wrapper or dense_dropin or lowest16" 2>&1 | tail -3

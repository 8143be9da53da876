"""Static checks on the SASS of the built library (no GPU needed: cuobjdump reads the cubin).  They pin what the
design claims about the hot kernel -- TMA loads (UTMALDG / UBLKCP), mbarrier pipeline (SYNCS), FP64 tensor-core math
(DMMA.8x8x4) -- and guard the two ordering fixes of the generic/async proxy hand-back
(profiles/r01_matvec_war_hazard_diagnosis_v6.txt): a block fence right before every consumer `mbarrier.arrive`, a
`fence.proxy.async` between the producer's wait and its `expect_tx` arrive + TMA issue."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "fortran_davidson_b200", "_build")
LIB = os.path.join(ROOT, "fortran_davidson_b200", "libdavidson_b200.so")


def _sass(name):
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    obj = os.path.join(OBJ, name + ".o")
    src = obj if os.path.exists(obj) else LIB
    if not os.path.exists(src):
        pytest.skip("library not built")
    out = subprocess.run([tool, "-sass", src], capture_output=True, text=True, timeout=300).stdout
    funcs, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and cur is not None:
            funcs[cur].append(m.group(1).strip())
    return funcs


def test_matvec_kernel_sass():
    funcs = {k: v for k, v in _sass("matvec_dmma").items() if "matvec_kernel" in k}
    assert len(funcs) == 16, sorted(funcs)  # 8 tile shapes x 2 stage depths
    for name, ins in funcs.items():
        text = "\n".join(ins)
        for mnemonic in ("DMMA.8x8x4", "UTMALDG.2D", "UBLKCP", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "LDS.128",
                         "FENCE.VIEW.ASYNC"):
            assert mnemonic in text, (name, mnemonic)
        assert "WGMMA" not in text and "HMMA" not in text
        # no 64-bit software division (CALL.REL to the div routine) per pipeline stage: only the per-segment index maths
        assert sum("CALL.REL" in i for i in ins) <= 8, name
        # consumer release: MEMBAR between the last fragment load / DMMA and the arrive
        arrives = [i for i, s in enumerate(ins) if "SYNCS.ARRIVE" in s and "A1T0" in s]
        assert arrives, name
        for a in arrives:
            back = ins[max(0, a - 40):a][::-1]
            k = next((j for j, s in enumerate(back) if "MEMBAR" in s), None)
            assert k is not None, (name, "no MEMBAR before the consumer arrive")
            assert not any(("LDS" in s or "DMMA" in s) for s in back[:k]), (name, "loads/MMAs between MEMBAR and arrive")
        # producer: proxy fence between the wait on the empty barrier and the expect_tx arrive
        expects = [i for i, s in enumerate(ins) if "SYNCS.ARRIVE.TRANS64 " in s and "A1T0" not in s]
        assert expects, name
        for e in expects:
            back = ins[max(0, e - 16):e]
            assert any("FENCE.VIEW.ASYNC" in s for s in back), (name, "no proxy fence before the TMA refill")


def test_tall_skinny_gemm_uses_the_tensor_pipe():
    funcs = _sass("dgemm")
    dm = [v for k, v in funcs.items() if "gemm_dmma_kernel" in k]
    # {TN, TN with 16-byte loads, NN} x {1, 2, 4 warp groups along K} x {128x32, 64x64 CTA tiles}
    assert len(dm) == 18
    for ins in dm:
        assert sum("DMMA.8x8x4" in i for i in ins) >= 32
        # latency tolerance: one chunk of loads (16 or 32 per lane) is issued before the first DMMA consumes it
        runs, cur = [], 0
        for i in ins:
            if "LDG" in i:
                cur += 1
            elif "DMMA" in i:
                runs.append(cur)
                cur = 0
        assert max(runs) >= 12, max(runs)


def _resource_usage():
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(LIB):
        pytest.skip("library not built")
    out = subprocess.run([tool, "--dump-resource-usage", LIB], capture_output=True, text=True, timeout=300).stdout
    usage, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
        if m and cur:
            usage[cur] = dict(zip(("reg", "stack", "shared", "local"), map(int, m.groups())))
            cur = None
    return usage


def test_hot_kernels_do_not_spill():
    """Register budgets of the kernels the solve time is made of (ptxas resource usage of the shipped cubin): the block
    matvec and the fused residual keep everything in registers; the tall-skinny GEMM variants do except the 4-warp-group
    ones, which trade a few spilled words for 512 threads at 128 registers; the single-CTA small-matrix kernels stay
    within a few words."""
    usage = _resource_usage()
    pick = lambda frag: {k: v for k, v in usage.items() if frag in k}
    mv = pick("matvec_kernel")
    assert len(mv) == 16
    assert all(v["stack"] == 0 and v["local"] == 0 and v["reg"] <= 168 for v in mv.values()), mv
    rs = pick("resid_dmma_kernel")
    assert len(rs) == 2 and all(v["stack"] == 0 for v in rs.values()), rs
    gm = pick("gemm_dmma_kernel")
    assert len(gm) == 18
    spilled = {k: v for k, v in gm.items() if v["stack"] > 0}
    assert len(spilled) == 6 and all(v["reg"] == 128 and v["stack"] <= 96 for v in spilled.values()), spilled
    tri = pick("tridiag_reg_kernel")
    assert tri and all(v["stack"] <= 32 for v in tri.values()), tri
    ev = pick("tri_eigvec_kernel")
    assert ev and all(v["stack"] == 0 for v in ev.values()), ev
    fr = pick("free_dmma_kernel")
    assert fr and all(v["reg"] <= 128 and v["stack"] <= 40 for v in fr.values()), fr   # 2 CTAs of 256 threads per SM
    assert all(v["local"] == 0 for v in usage.values())

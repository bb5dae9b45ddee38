"""CPU tier: properties of the BUILT device code that the measured performance rests on, read from the SASS of libmpcc_b200.so
(cuobjdump; no GPU needed).  They are code-generation facts, not arithmetic -- a compiler or source change can lose them silently while
every parity test stays green (the noinline-member trap of the SQP kernel, the convergence-barrier trap of the MMA issuer)."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "mpcc_manipulator_b200" / "libmpcc_b200.so"


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    if not LIB.exists():
        pytest.skip("library not built")
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); per[cur] = []
        elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            per[cur].append(line)
    return txt, per


def kernel(per, key):
    names = [n for n in per if key in n]
    assert len(names) == 1, (key, names)
    return per[names[0]]


def test_sm_100a_only(sass):
    txt, _ = sass
    archs = set(re.findall(r"arch = (sm_\w+)", txt))
    assert archs == {"sm_100a"}, archs


def test_mlp_kernel_uses_tcgen05_and_tmem(sass):
    """The default network kernel issues its three 256 x 256 layers as tcgen05.mma kind::i8 with the A operand copied to TMEM and the
    accumulators read back by tcgen05.ld: 2 passes' worth of unrolled issue code = S (S + 1) / 2 x 8 = 224 MMAs, 56 copies (S = 7)."""
    _, per = sass
    k = kernel(per, "k_mlp_oz")
    n = lambda op: sum(op in l for l in k)
    assert n("UTCIMMA") == 224 and n("UTCCP") == 56 and n("LDTM") >= 7 and n("DMMA") > 0 and n("LDGSTS") > 0


def test_no_convergence_barriers_inside_the_mma_issue(sass):
    """Between the first and the last tcgen05.mma of k_mlp_oz there must be no BSSY / BSYNC pair: with a warp index the compiler cannot prove
    uniform, ptxas wraps each per-chunk single-thread region of the issuer in one (2.5 k cycles per pass, 9 % of the kernel)."""
    _, per = sass
    k = kernel(per, "k_mlp_oz")
    idx = [i for i, l in enumerate(k) if "UTCIMMA" in l]
    span = k[idx[0]:idx[-1] + 1]
    assert not any("BSSY" in l or "BSYNC" in l for l in span), sum("BSSY" in l for l in span)


def test_default_sqp_kernels_do_not_carry_the_correction(sass):
    """The second-order correction is a template flag so that the default SQP kernels stay the code they were tuned as: the instantiations
    with it (two inlined copies of the interior point) are visibly larger."""
    _, per = sass
    plain = len(kernel(per, "10k_sqp_warpE"))
    soc = len(kernel(per, "k_sqp_warp_soc"))
    assert soc > 1.2 * plain, (plain, soc)
    # the steady-state build keeps five CTAs per SM: its register count is capped by the launch bounds (checked through the spills it accepts:
    # local-memory instructions exist) while the solver object itself must not live in local memory (the noinline-member trap: +514 LDL)
    ldl = sum("LDL" in l for l in kernel(per, "10k_sqp_warpE"))
    assert ldl < 1400, ldl

"""The fp64 Gaussian of the DP kernel, emulated on the host with the kernel's own constants (no GPU needed)."""
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_table_polynomial_gaussian_error_bound(tmp_path):
    src = open(os.path.join(REPO, "kamphir_b200", "csrc", "phylok_dev.cu")).read()
    m = re.search(r"pk_poly\[4\]\s*=\s*\{([^}]*)\}", src)
    coef = [c.strip() for c in m.group(1).split(",")]
    assert len(coef) == 4
    exe = str(tmp_path / "gauss_check")
    cmd = ["gcc", "-O2", "-ffp-contract=off"] + ["-DPK_C%d=%s" % (k + 1, c) for k, c in enumerate(coef)]
    subprocess.check_call(cmd + ["-o", exe, os.path.join(REPO, "scripts", "gauss_check.c"), "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True)
    print(out.stdout)
    assert out.returncode == 0, out.stdout
    errs = [float(x) for x in re.findall(r"max rel err ([0-9.e+-]+)", out.stdout)]
    assert len(errs) == 7 and max(errs) < 1e-14            # 5e-15
    assert "E=0 -> 0.2000000000000000" in out.stdout         # exp(0) = 1 exactly: self-kernels keep their exact cells

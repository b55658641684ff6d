#!/usr/bin/env python
"""
Copy what the GPU box needs of the reference into baseline/_ref/ (git-ignored, NOT
gpurun-ignored: it travels with the snapshot, it never enters history):

    baseline/_ref/src/optpricing/{atoms,models,techniques,calibration/technique_selector.py}
    baseline/_ref/tests/techniques/{test_monte_carlo,test_american_monte_carlo}.py
    baseline/_ref/tests/techniques/{kernels/test_mc_kernels,kernels/test_path_kernels}.py
    baseline/_ref/tests/techniques/base/*.py

Used by (a) tests/test_gpu_reference_suite.py: the reference's own 47 hot-path tests run
unmodified against the drop-in on the B200, and (b) bench.py's cpu_baseline.reference_numba:
the reference's numba path timed on the GPU box's host cores.
/root/reference exists only in the build container; __graft_entry__.build() calls this there.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OPTPRICING_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def populate(verbose=False) -> bool:
    if not os.path.isdir(os.path.join(REF, "src", "optpricing")):
        return os.path.isdir(os.path.join(DST, "src", "optpricing"))
    shutil.rmtree(DST, ignore_errors=True)
    ign = shutil.ignore_patterns("__pycache__", "*.pyc", "*.nbi", "*.nbc")
    for sub in ("atoms", "models", "techniques"):
        shutil.copytree(os.path.join(REF, "src", "optpricing", sub),
                        os.path.join(DST, "src", "optpricing", sub), ignore=ign)
    os.makedirs(os.path.join(DST, "src", "optpricing", "calibration"))
    shutil.copy(os.path.join(REF, "src", "optpricing", "calibration", "technique_selector.py"),
                os.path.join(DST, "src", "optpricing", "calibration", "technique_selector.py"))
    open(os.path.join(DST, "src", "optpricing", "calibration", "__init__.py"), "w").close()
    t = os.path.join(REF, "tests", "techniques")
    for rel in ("test_monte_carlo.py", "test_american_monte_carlo.py",
                "kernels/test_mc_kernels.py", "kernels/test_path_kernels.py"):
        dst = os.path.join(DST, "tests", "techniques", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copy(os.path.join(t, rel), dst)
    shutil.copytree(os.path.join(t, "base"), os.path.join(DST, "tests", "techniques", "base"),
                    ignore=ign)
    for extra in ("LICENSE", "pyproject.toml"):
        if os.path.exists(os.path.join(REF, extra)):
            shutil.copy(os.path.join(REF, extra), os.path.join(DST, extra))
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print(f"baseline/_ref populated from {REF}: {n} files")
    return True


if __name__ == "__main__":
    ok = populate(verbose=True)
    sys.exit(0 if ok else 1)

"""
The reference's OWN hot-path tests, unmodified, against the drop-in on the B200
(tests/techniques/{test_monte_carlo,test_american_monte_carlo}.py,
tests/techniques/kernels/{test_mc_kernels,test_path_kernels}.py, tests/techniques/base/*).

They live, with the reference's atoms / models / techniques they import, in the git-ignored
baseline/_ref/ (tools/populate_baseline_ref.py, run by __graft_entry__.build() where
/root/reference exists); the overlay (optpricing_b200/overlay.py) redirects the five Monte
Carlo modules to this engine and everything else resolves to the reference's files.

Two runs:
  * rng_mode="numpy": the engine replays the reference's host draws, so ALL of them must pass,
    including the lucky-seed price test (test_monte_carlo.py:111-126, abs=1e-2 at 20 000 x 10
    where one standard error is 0.07);
  * rng_mode="philox" (the default, throughput mode): all but that one, which the device RNG
    meets statistically only (tests/test_gpu_parity.py::test_philox_bsm_against_closed_form).
"""
import os
import re
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
SUITE = ["tests/techniques/test_monte_carlo.py", "tests/techniques/test_american_monte_carlo.py",
         "tests/techniques/kernels/test_mc_kernels.py", "tests/techniques/kernels/test_path_kernels.py",
         "tests/techniques/base"]


def _run(rng_mode, extra):
    if not os.path.isdir(os.path.join(REF, "src", "optpricing")):
        pytest.skip("baseline/_ref not populated (tools/populate_baseline_ref.py needs /root/reference)")
    files = ", ".join(repr(os.path.join(REF, f)) for f in SUITE)
    code = textwrap.dedent(f"""
        import sys, pytest
        import optpricing_b200.overlay as overlay
        overlay.install(reference_src={os.path.join(REF, "src")!r}, rng_mode={rng_mode!r})
        import optpricing_b200 as ob
        from optpricing.techniques import MonteCarloTechnique, AmericanMonteCarloTechnique
        assert MonteCarloTechnique is ob.MonteCarloTechnique
        assert AmericanMonteCarloTechnique is ob.AmericanMonteCarloTechnique
        from optpricing.techniques.kernels import mc_kernels
        assert mc_kernels.__name__ == "optpricing_b200.techniques.kernels.mc_kernels"
        from optpricing.models import HestonModel           # the reference's own models
        assert "baseline/_ref" in sys.modules["optpricing.models"].__file__
        sys.exit(pytest.main(["-c", "/dev/null", "--rootdir=/tmp", "-p", "no:cacheprovider", "-q",
                              {files}] + {extra!r}))
    """)
    env = dict(os.environ, PYTHONPATH=ROOT, NUMBA_CACHE_DIR="/tmp/numba_cache",
               PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                       timeout=1200)
    tail = r.stdout[-4000:] + r.stderr[-2000:]
    m = re.search(r"(\d+) passed", r.stdout)
    return r.returncode, int(m.group(1)) if m else 0, tail


def test_reference_hot_path_tests_pass_in_numpy_mode():
    rc, passed, tail = _run("numpy", [])
    assert rc == 0 and passed >= 47, tail


def test_reference_hot_path_tests_pass_on_the_device_rng():
    rc, passed, tail = _run("philox", ["-k", "not test_mc_price_correctness_for_bsm"])
    assert rc == 0 and passed >= 46, tail

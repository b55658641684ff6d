"""
Drop-in check against the reference's OWN test files (build container only:
needs /root/reference).  With the overlay installed, the reference's dispatch /
selector / error-path tests for the two Monte Carlo techniques must pass
unmodified against this engine's classes.  Tests that price (need a GPU) are
deselected here and covered on the B200 by tests/test_gpu_parity.py.
"""
import os
import subprocess
import sys
import textwrap

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def _run(code):
    env = dict(os.environ, PYTHONPATH=ROOT, NUMBA_CACHE_DIR="/tmp/numba_cache",
               PYTHONDONTWRITEBYTECODE="1")
    return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env,
                          capture_output=True, text=True, timeout=600)


def test_reference_technique_tests_pass_on_the_overlay():
    r = _run(f"""
        import sys, pytest
        import optpricing_b200.overlay as overlay
        overlay.install(reference_src="{REF}/src")
        from optpricing.techniques import MonteCarloTechnique, AmericanMonteCarloTechnique
        import optpricing_b200 as ob
        assert MonteCarloTechnique is ob.MonteCarloTechnique
        assert AmericanMonteCarloTechnique is ob.AmericanMonteCarloTechnique
        # the reference's own selector still resolves to the engine
        from optpricing.calibration.technique_selector import select_fastest_technique
        from optpricing.models import SABRModel
        assert type(select_fastest_technique(SABRModel())) is ob.MonteCarloTechnique
        sys.exit(pytest.main(["-c", "/dev/null", "--rootdir=/tmp", "-p", "no:cacheprovider", "-q",
                              "{REF}/tests/techniques/test_monte_carlo.py",
                              "{REF}/tests/techniques/test_american_monte_carlo.py",
                              "{REF}/tests/techniques/base",
                              "-k", "not correctness_for_bsm and not against_lattice"]))
    """)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_reference_models_and_atoms_drive_the_engine_classes():
    """Reference atoms/models are accepted by duck typing (no GPU: stops at the engine)."""
    r = _run(f"""
        import optpricing_b200.overlay as overlay
        overlay.install(reference_src="{REF}/src")
        from optpricing.atoms import Option, OptionType, Rate, Stock
        from optpricing.models import HestonModel, KouModel, SABRJumpModel, BSMModel
        from optpricing.techniques import MonteCarloTechnique
        from optpricing.techniques.kernels import mc_kernels, path_kernels
        import optpricing_b200 as ob
        from optpricing_b200 import _engine
        t = MonteCarloTechnique(n_paths=101, n_steps=3, seed=1)
        assert t.n_paths == 102
        for m, kind in ((HestonModel(), "heston"), (KouModel(), "kou"),
                        (SABRJumpModel(), "sabr_jump"), (BSMModel(), "bsm")):
            assert ob.model_kind(m) == kind
            fn, kp = t._get_sde_kernel_and_params(m, 0.05, 0.0, 0.01)
            assert fn is getattr(mc_kernels, kind + "_kernel")
            _engine.make_model(kind, m.params, 0.05, 0.0)
        import torch
        if not torch.cuda.is_available():
            try:
                t.price(Option(100, 1.0, OptionType.CALL), Stock(100.0), HestonModel(), Rate(0.05))
            except _engine.EngineError as e:
                assert "no CPU fallback" in str(e)
            else:
                raise SystemExit("priced without a GPU")
    """)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]

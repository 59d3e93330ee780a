"""-m gpu: the input pipeline and the evaluate / predict tails through the C ABI,
bit-exact against the oracle and the reference fixtures.  See tests/gpu_io_checks.py."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", ["io_pipeline", "io_pipeline_full", "eval_tail", "predict_tail", "tail_entry_points"])
def test_io_and_tails(group):
    import gpu_io_checks as G
    results = G.GROUPS[group]()
    assert results
    bad = [(label, err, tol) for label, err, tol in results if not (err <= tol)]
    assert not bad, "parity failures: " + "; ".join(f"{l}: err={e:.3e} > tol={t:.1e}" for l, e, t in bad)

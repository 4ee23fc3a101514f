"""CPU, build container only: the C oracle step-for-step against the LIVE unmodified reference
(skipped where /root/reference is absent, e.g. on the GPU box)."""
import pytest

from oracle import _refload


@pytest.mark.skipif(not _refload.available(), reason="reference tree not present")
def test_oracle_matches_live_reference():
    from oracle.validate_against_reference import run

    n, g = run(scale=1)
    assert n > 1000 and g > 300

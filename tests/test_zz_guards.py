"""Runs last (alphabetical order): what the guard zones around the device buffers saw during this session.
compute-sanitizer is closed on the GPU pool this repository is measured on (profiles/README.md, round 2), so the
library brackets every device buffer with 256-byte pattern zones itself and every context opened by the tests
(SB200_GUARD=1, tests/conftest.py) compares them when it is closed."""
import pytest

pytestmark = pytest.mark.gpu


def test_no_kernel_wrote_outside_its_buffers():
    from simplex_b200 import _abi
    st = _abi.GUARD_STATS
    print("guard zones: %d contexts, %d buffers, %d damaged" % (st["contexts"], st["buffers"], st["damaged"]))
    assert _abi.GUARD_CHECK_AT_CLOSE
    assert st["damaged"] == 0, st
    # the whole GPU suite opens hundreds of contexts; a lone run of this file opens none
    assert st["contexts"] == 0 or st["buffers"] >= 10 * st["contexts"]

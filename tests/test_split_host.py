"""Host logic of the speculative chain splitting (abi_gw.inc), without a GPU: the library's own
self-test cuts a toy search - whose state after a change is a function of ``start`` alone, the
property the split relies on - into sub-chains with give-up limits, stitches them round by round
with the real ``gw_walk`` / ``gw_plan_continuations`` and compares the result with the toy's
sequential search, record for record.  ``p_sync`` is the probability that searches from
different starts agree on a change: 1.0 = every sub-chain meets the next one at once,
0.0 = never (everything is continued until the unlimited rounds)."""

import pytest

import spkdiar                              # noqa: F401
from spkdiar import _abi


@pytest.mark.parametrize('p_sync', [1.0, 0.7, 0.3, 0.05, 0.0])
def test_stitch_equals_sequential(p_sync):
    lib = _abi.load_library()
    for seed in range(25):
        for nframes, target in ((360000, 148), (60000, 148), (60000, 17), (12345, 8), (3999, 148)):
            rc = lib.spkdiar_selftest_stitch(seed, p_sync, nframes, target)
            assert rc == 0, (seed, p_sync, nframes, target, rc)

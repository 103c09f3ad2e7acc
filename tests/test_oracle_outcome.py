"""CPU: the per-leaf legal-outcome checker (tests/common.check_frame_outcome) accepts the oracle's canonical frames
and rejects a tampered pool -- the checker itself is test infrastructure the GPU parity tests lean on."""
import numpy as np
import pytest

from common import check_frame_outcome, descend, pkg
from oracle import oracle as orc


@pytest.mark.parametrize("D", [6, 8])
def test_checker_accepts_oracle_frames_and_rejects_tampering(D):
    P = pkg()
    w, h = 160, 120
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    t = orc.OracleSVO(center, half, D)
    before = t.pool()
    seen = 0
    for k in range(3):
        pose = P.synth.orbit_pose(25 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        keys = orc.compute_keys(orc.transform(orc.vertex_map(depth, fx, fy), pose), center, half, D)
        t.integrate_depth(depth, rgb, fx, fy, pose)
        after = t.pool()
        n = check_frame_outcome(before, after, keys, rgb, D, canonical=True)
        assert n == t.counters().n_unique - seen  # the oracle's counters are running totals
        seen = t.counters().n_unique
        check_frame_outcome(before, after, keys, rgb, D, canonical=False)  # the canon is one of the legal outcomes
        bad = after.copy()
        uk = np.unique(keys[keys != 1])
        leaf = int(descend(after, uk[uk.size // 2:uk.size // 2 + 1], D)[D - 1, 0])
        bad[2 * leaf + 1] ^= 0x00010000
        with pytest.raises(AssertionError):
            check_frame_outcome(before, bad, keys, rgb, D, canonical=False)
        before = after

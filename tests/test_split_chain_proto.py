"""CPU check of the lane-accurate prototype of the four-lanes-per-instance reduced-KKT solve (tools/split_chain_proto.py,
DESIGN.md section 5): upper lane top-down, lower lane bottom-up with the u_0 border, dense 6x6 on the separators and u_0,
against a dense solve of the same reduced KKT matrix taken from the oracle's MPC."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _proto():
    spec = importlib.util.spec_from_file_location("split_chain_proto", os.path.join(ROOT, "tools", "split_chain_proto.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("N", [10, 25])
def test_split_solve_matches_dense(N):
    P = _proto()
    rng = np.random.default_rng(N)
    for K in P.reduced_kkt_from_oracle(N, tick=12):
        cx, cy = P.chain_order(N, 0), P.chain_order(N, 1)
        n = len(cx)
        Tx, Ty, Cxy = K[np.ix_(cx, cx)], K[np.ix_(cy, cy)], K[np.ix_(cx, cy)]
        kxy = Cxy[n - 2, n - 2]
        assert np.count_nonzero(Cxy) <= 1                      # the axes couple through u_0 only
        assert np.abs(np.triu(Tx, 3)).max() == 0.0             # pentadiagonal in chain order
        r = rng.standard_normal(2 * n)
        for s in ((n - 2) // 2 - 1, 4, n - 8):                 # separator in the middle and near either end
            zx, zy = P.solve_split(Tx, Ty, kxy, r[:n], r[n:], s)
            zref = np.linalg.solve(np.block([[Tx, Cxy], [Cxy.T, Ty]]), r)
            err = np.abs(np.concatenate([zx, zy]) - zref).max() / np.abs(zref).max()
            assert err < 1e-10, (N, s, err)

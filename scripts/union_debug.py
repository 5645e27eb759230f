"""Debug helper: device-emitted rows of every union child vs the oracle's cursors (first mismatch per child)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solverforge_b200 import ForageParams, GpuScoreDirector, instances, models  # noqa: E402
from solverforge_b200 import _lib as L  # noqa: E402
from tests import oracle_lib  # noqa: E402
from tests.oracle_lib import Oracle  # noqa: E402

D = [(0, 20), (1, 20), (2, 1, 3), (3, 1, 3), (4,)]
c = instances.cvrp(46, 7, seed=31)
c.matrix = (c.matrix // 40) * 40
R = 3
starts = [instances.perturb_routes(c, 40 + r, 30 + 10 * r) for r in range(R)]
d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
lib = L.load()
W = 1 << 14
bad = 0
for order in (0, 1, 2):
    seeds, steps = [5, 77, 0xDEADBEEF], [0, 3, 900]
    desc = GpuScoreDirector.union_desc(D, L.UNION_SEQUENTIAL, order, W, W)
    d.step_union(desc, ForageParams(0, 0, 0), step_seeds=seeds, step_indices=steps)
    for r in range(R):
        kids = oracle_lib.union_children(oracles[r], D, steps[r], seeds[r], order)
        for ci, k in enumerate(kids):
            rows = np.zeros((W, 4), dtype=np.uint32)
            n, ended = C.c_uint32(), C.c_uint32()
            lib.sfgpu_debug_union_rows(C.c_void_p(d.h.value if hasattr(d.h, "value") else d.h), r, ci, W, rows.ctypes.data_as(C.c_void_p), W, C.byref(n), C.byref(ended))
            want = k[1][:W]
            got = rows[:n.value].astype(np.int64)
            ok = len(got) == min(len(k[1]), W) and (got == want).all() and ended.value == (1 if len(k[1]) <= W else 0)
            if not ok:
                bad += 1
                m = min(len(got), len(want))
                diff = np.flatnonzero((got[:m] != want[:m]).any(axis=1))
                at = int(diff[0]) if len(diff) else m
                print(f"order={order} r={r} child={ci} fam={D[ci][0]} n_dev={len(got)} n_ref={len(k[1])} ended={ended.value} first_diff={at}")
                print("  dev", got[max(at - 1, 0):at + 3].tolist())
                print("  ref", want[max(at - 1, 0):at + 3].tolist())
print("mismatching children:", bad)

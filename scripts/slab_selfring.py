#!/usr/bin/env python
"""One slab on one GPU whose ring neighbours are itself (periodic in z): runs the slab launch structure -- wait, boundary tiles,
interior tiles, side-stream halo exchange of the moment planes -- in a single stream order, so that it can run under
compute-sanitizer (which makes launches synchronous); compared with the undecomposed supercell.  Development tool."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ncell = (64, 8, 16)
for solver in (1, 5):
    ref, n = bench.bcc_engine(ncell, solver, 300.0, 0.5, 1, 0, 0)
    sl, _ = bench.bcc_engine(ncell, solver, 300.0, 0.5, 1, 0, 0, slab=(1, 0, None))
    assert sl.layout_info()['planes'] == 1, sl.layout_info()
    ref.sd_steps(6)
    sl.sd_steps(6)
    a, b = ref.get_moments()[0], sl.get_moments()[0]
    print('solver', solver, 'slab (self-ring) vs undecomposed: max diff', float(np.abs(a - b).max()), 'slab status', sl.slab_status())
    assert np.array_equal(a, b)
    ref.close(); sl.close()

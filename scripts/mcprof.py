#!/usr/bin/env python
"""Phase timing of the Monte Carlo run-form block kernel (library built with ASD_MC_PROF=1): clock64 at the phase barriers of
thread 0 of every CTA, summed per phase.  Development tool."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ncell = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 128, 128))]
e, n = bench.bcc_engine(ncell, 1, 300.0, 0.5, 1, 0, 0)
lib = e.lib
for mode in ('M', 'H'):
    e.mc_sweeps(mode, 3, 300.0)
    e.synchronize()
    buf = (C.c_ulonglong * 8)()
    lib.asd_debug_mc_prof(None, 1)
    ms = e.time_mc_sweeps(mode, 10, 300.0)
    lib.asd_debug_mc_prof(buf, 0)
    tiles = buf[3] / 2.0          # two batches per tile
    print('   per tile: neighbour loops %.0f, accept %.0f, barrier wait %.0f (warp 0)' % (buf[5] / tiles, buf[6] / tiles, buf[7] / tiles))
    print('%s %dx%dx%d: %.3f ms/sweep; per tile-CTA cycles: stage %.0f, draws %.0f, colours %.0f, wait %.0f (tile visits %d)'
          % (mode, *ncell, ms / 10, buf[0] / tiles, buf[1] / tiles, buf[2] / tiles, buf[4] / tiles, tiles))

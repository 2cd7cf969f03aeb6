#!/usr/bin/env python
"""Size / ensemble table of SURVEY 8(d): bcc Fe, midpoint, atom-steps/s for supercells from 6^3 to 256^3 cells at T = 0 and
300 K (Mensemble 1) and for Mensemble 8 on one GPU, device-built tables, CUDA-event timing."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

for ncell, mens, steps in (((6, 6, 6), 1, 20000), ((6, 6, 6), 8, 20000), ((32, 32, 32), 1, 2000), ((64, 64, 64), 1, 500),
                           ((64, 64, 64), 8, 100), ((128, 128, 128), 1, 200), ((128, 128, 128), 8, 30), ((256, 256, 256), 1, 30)):
    for temp in (0.0, 300.0):
        e, n = bench.bcc_engine(ncell, 1, temp, 0.5, mens, 0, 0)
        e.sd_steps(5)
        ms = e.time_sd_steps(steps, first_step=6)
        lay = e.layout_info()
        rate = n * mens * steps / (ms * 1e-3)
        print('bcc %3dx%3dx%3d M=%d T=%3.0f | %9d atom-ensembles | %10.4f ms/step | %.3e atom-steps/s | %5.1f %% of 536 B roofline | runs=%d tile=%d'
              % (*ncell, mens, temp, n * mens, ms / steps, rate, 100 * 536 * rate / 6550.1e9, lay['runs'], lay['tile_slots']), flush=True)
        e.close()

#!/usr/bin/env python
"""Kernel micro-benchmark: times the two fused stage kernels on the bcc Fe supercell for the current build /
environment knobs (ASD_VARIANT, ASD_PF; compile-time ASD_MINB, ASD_CHUNK).  Development tool."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ncell = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 128, 128))]
    tag = ' '.join('%s=%s' % (k, os.environ.get(k, '-')) for k in ('ASD_STAGED', 'ASD_RUNS', 'ASD_PF', 'ASD_KEYWRAP'))
    for solver in (1, 5):
        for temp in (0.0, 300.0):
            e, n = bench.bcc_engine(ncell, solver, temp, 0.5, 1, 0, 0)
            e.sd_steps(5)
            if solver == 1 and temp == 0.0:
                print('KB layout', e.layout_info(), flush=True)
            ms = e.time_sd_steps(40, first_step=6)
            s1, s2 = [], []
            for r in range(5):
                _, (a, b) = e.time_sd_steps(0, first_step=100 + r, stages=True)
                s1.append(a); s2.append(b)
            print('KB %s | solver %d T=%3.0f | step %.4f ms  stage1 %.4f  stage2 %.4f | %.3e atom-steps/s | roof(536B) %.3f'
                  % (tag, solver, temp, ms / 40, np.median(s1), np.median(s2), n * 40 / (ms * 1e-3),
                     (536 if solver == 1 else 584) * n * 40 / (ms * 1e-3) / 6550.1e9), flush=True)
            e.close()
            if solver == 5 and temp > 0:
                break


if __name__ == '__main__':
    main()

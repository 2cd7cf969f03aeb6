import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for ncell, mens, steps in (((32, 32, 32), 1, 3000), ((48, 48, 48), 1, 1000), ((64, 64, 64), 1, 800), ((128, 128, 128), 1, 300)):
    for temp in (0.0, 300.0):
        e, n = bench.bcc_engine(ncell, 1, temp, 0.5, mens, 0, 0)
        e.sd_steps(5)
        ms = e.time_sd_steps(steps, first_step=6)
        rate = n * mens * steps / (ms * 1e-3)
        print('bcc %3dx%3dx%3d M=%d T=%3.0f | %10.4f ms/step | %.3e atom-steps/s | %5.1f %%' % (*ncell, mens, temp, ms / steps, rate, 100 * 536 * rate / 6550.1e9), flush=True)
        e.close()

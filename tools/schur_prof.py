"""Reads gpurun_out/schur_prof.bin (G2OCU_SCHUR_DEBUG=3 instrumentation of schur_mma_kernel) and prints the per-warp cycle breakdown."""
import numpy as np, sys
a = np.fromfile(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/schur_prof.bin', dtype=np.int64).reshape(-1, 8, 8)
loop, wait, prod, slots, epi, ne = a[:, :, 0], a[:, :, 1], a[:, :, 2], a[:, :, 3], a[:, :, 4], a[:, 0, 5]
print("chunks", len(a), "entries", ne.sum(), "slots", slots.sum(), "slots/entry", slots.sum() / ne.sum())
print("ms if SMs perfectly packed: loop %.2f  epilogue %.2f" % (loop.max(axis=1).sum() / 148 / 1.965e6, epi.max(axis=1).sum() / 148 / 1.965e6))
print("wait fraction %.3f, warp0 produce fraction %.3f" % (wait.sum() / loop.sum(), prod[:, 0].sum() / loop[:, 0].sum()))
comp = loop - wait - prod
print("compute cycles per slot per warp %.1f;  pipe-bound ms at 16.5 cyc/slot/SMSP: %.2f" % (comp.sum() / slots.sum(), slots.sum() * 16.5 / 592 / 1.965e6))
print("busiest-warp slots * 8 / total: %.3f" % (slots.max(axis=1).sum() * 8 / slots.sum()))
sm = slots[:, :4] + slots[:, 4:]
print("busiest-SMSP slots * 4 / total: %.3f" % (sm.max(axis=1).sum() * 4 / slots.sum()))
for w in range(8):
    print("warp", w, "loop %.0fM wait %.0fM slots %.1fM cyc/slot %.1f" % (loop[:, w].sum() / 1e6, wait[:, w].sum() / 1e6, slots[:, w].sum() / 1e6, comp[:, w].sum() / max(1, slots[:, w].sum())))

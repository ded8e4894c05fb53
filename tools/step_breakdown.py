import sys, time, torch
sys.path.insert(0, "/root/repo")
from madflow_b200 import integrand as mfi, matrix as mfm, vegas as mfv
MT = 173.0
m, model = mfm.get_process("1_gg_ttxgg")
fi = mfi.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT, 0, 0], pt_cut=30.0, lab_frame=True, running=True)
v = mfv.VegasFlow(fi.n_dim, 100_000_000, seed=4)
v.compile(fi)
for _ in range(3): v.run_iteration()
evs = []
orig = v._run_chunk_fused
def timed(first, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(first, n); e1.record(); evs.append((n, e0, e1))
v._run_chunk_fused = timed
torch.cuda.synchronize(); t0 = time.perf_counter()
E0, E1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
E0.record(); v.run_iteration(); E1.record(); torch.cuda.synchronize()
print("iteration wall %.1f ms, device %.1f ms, ME events %d" % (1e3 * (time.perf_counter() - t0), E0.elapsed_time(E1), v.last_me_events))
print("chunks:", ["%d: %.1f ms" % (n, a.elapsed_time(b)) for n, a, b in evs])
# the same chunk launched alone, as bench.time_kernel_alone does
nblocks = fi.nblocks(); partial = v._partial_buf(nblocks)
for it in (9999, 10000, 10001, 3):
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(); fi.launch(v.divisions, 4, it, 0, 8388608, 1e-8, partial, nblocks, True); k1.record(); torch.cuda.synchronize()
    print("alone, iteration index", it, "%.1f ms" % k0.elapsed_time(k1))

#!/usr/bin/env python
"""Where does the time of the reference's generate_max_style_image go with either layer?  (measurement infrastructure under tests/: it imports oracle/)
cProfile of the host side + torch.profiler totals of the device side, FCN_16 on the notebook batch."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ref_shims, ref_loop
from maxstyle_b200 import MaxStyle

ref = ref_shims.load()
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
solver = ref_loop.build_solver(ref, "FCN_16_standard_no_STN", use_gpu=True, pretrained=True)
img, lab = ref_loop.load_fixture(ref, dev)
for name, cls in (("reference", ref.MaxStyle), ("replacement", MaxStyle)):
    for _ in range(2):
        ref_loop.run_loop(ref, solver, img, lab, cls, n_iter=5)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); ref_loop.run_loop(ref, solver, img, lab, cls, n_iter=5); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    print(name, "host-return ms / synced ms:", [(round(a, 1), round(b, 1)) for a, b in ts])
    pr = cProfile.Profile(); pr.enable()
    ref_loop.run_loop(ref, solver, img, lab, cls, n_iter=5)
    pr.disable(); torch.cuda.synchronize()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:6000])
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        ref_loop.run_loop(ref, solver, img, lab, cls, n_iter=5); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60)[:7000])

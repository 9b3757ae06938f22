import sys, os, io, cProfile, pstats, time
sys.path.insert(0, '/root/repo')
import torch
import bench
import spkdiar
from spkdiar import _abi, synth
rec = bench.make_recording(0)
host = torch.from_numpy(rec.frames).pin_memory()
lines = synth.one_line_recipe('/syn/c2_0.wav', rec)
ctx = _abi.Context(0, stream=torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    bench.e2e_step(ctx, host, lines)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    bench.e2e_step(ctx, host, lines)
pr.disable()
torch.cuda.synchronize()
print('ms per step', (time.perf_counter() - t0) / 3 * 1e3)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28)
print(s.getvalue()[:6000])

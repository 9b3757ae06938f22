"""Timing driver (not a test): sharded clustering of a config-5-like recording under torchrun
(one rank per GPU, NCCL all-gather of the 16-byte candidates) or as a single process.
Usage: [torchrun --nproc-per-node N] python profiles/drivers/sharded_time.py <nsegments>"""
import hashlib
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import spkdiar                                   # noqa: F401
from spkdiar import synth, _abi, sharded

nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
ex = None
nid = None
mbx = None
mode = os.environ.get('SPKDIAR_SHARD_MODE', 'nccl')
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ex = sharded.dist_exchange()
rec = synth.config5(n_frames=nseg * 173)
a = [t[0] for t in rec.turns]
b = [t[1] for t in rec.turns]
ctx = _abi.Context(local)
f = ctx.upload(rec.frames)
if world > 1 and mode == 'p2p':
    mbx = sharded.Mailboxes(ctx)
for rep in range(2):
    if world > 1:
        if mode == 'nccl':
            nid = sharded.broadcast_nccl_id()      # an id makes ONE communicator
        dist.barrier()
    torch.cuda.synchronize()
    ctx.profile(True)
    t0 = time.perf_counter()
    with f.cluster(a, b, _abi.BIC, 1.3) as cl:
        if mode == 'p2p' and world > 1:
            merges, stats = cl.run_sharded_p2p(0.0, 0, rank, world, mbx.ptrs, mbx.next_base(len(a)))
        elif mode == 'nccl' or world == 1:
            merges, stats = cl.run_sharded_nccl(0.0, 0, rank, world, nid)
        else:
            merges, stats = cl.run_sharded(0.0, 0, rank, world, ex)
    dt = time.perf_counter() - t0
    prof = ctx.profile_read()
    ctx.profile(False)
hours = rec.frames.shape[0] / 100.0 / 3600.0
print(mode, 'rank %d/%d: segments %d merges %d speakers %d  wall %.1f ms (fill %.1f ms, merge kernels %.1f ms)  %.2f audio-h/s  sha %s'
      % (rank, world, len(a), len(merges), len(a) - len(merges), dt * 1e3, prof['score'][0], prof['merge'][0],
         hours / dt, hashlib.sha256(merges.tobytes()).hexdigest()[:12]), flush=True)
if mbx is not None:
    dist.barrier()
    mbx.close()
if world > 1:
    dist.destroy_process_group()

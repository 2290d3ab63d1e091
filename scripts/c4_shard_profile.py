"""One rank's share of config 4 at W ranks, on one GPU: 1M queries x (62500 / W) vocabulary rows through slm_knn2_exchange
with world = 1 (the rank's own buffers only), so that the search / refine / merge kernels of the sharded step can be timed
and captured with ncu without a multi-rank launch.  usage: c4_shard_profile.py [W] [steps]"""
import ctypes
import sys

import torch

from slammatch import _lib, synth

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nq, total = 1 << 20, 62500
nt = total // W
q = torch.from_numpy(synth.uniform(nq, 1)).cuda()
t = torch.from_numpy(synth.uniform(nt, 2)).cuda()
ctx = _lib.Context(0)
cap = nq
keys = torch.zeros((4, 1, cap, 2), dtype=torch.int64, device="cuda")
flags = torch.zeros((4,), dtype=torch.int32, device="cuda")
kp, fp = (ctypes.c_uint64 * 1)(keys.data_ptr()), (ctypes.c_uint64 * 1)(flags.data_ptr())
idx = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
dist = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
acc = torch.empty((nq,), dtype=torch.uint8, device="cuda")


def step(i):
    _lib.check(ctx.lib.slm_knn2_exchange(ctx.handle, q.data_ptr(), nq, t.data_ptr(), nt, 0, cap, total,
                                         ctypes.cast(kp, ctypes.c_void_p), ctypes.cast(fp, ctypes.c_void_p), 0, 1, i + 1, 7, 10,
                                         idx.data_ptr(), dist.data_ptr(), acc.data_ptr(), None))


for i in range(3):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    step(3 + i)
e1.record()
torch.cuda.synchronize()
_lib.check(ctx.lib.slm_exchange_status(ctx.handle))
print(f"W={W} nt={nt} step {e0.elapsed_time(e1) / steps:.4f} ms  kernel {ctx.last_kernel()}  accepted {int(acc.sum())}")

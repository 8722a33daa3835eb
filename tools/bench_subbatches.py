"""One handle, K split steps back to back (no L2 flush): effect of SG_SUBBATCHES (concurrent sub-batches of the
handle on internal streams, replayed as a CUDA graph when SG_GRAPHS=1).  Run once per setting (the env var is
read at sg_create): SG_SUBBATCHES=2 python tools/bench_subbatches.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spacegym_b200 as sg  # noqa: E402

for n in (262144, 1048576):
    env = sg.make("GoalContinuous3P-v0", n, seed=1)
    env.reset()
    env.rollout_random(64)
    torch.cuda.synchronize()
    act = (torch.rand((n, 2), device=env.device) * 2 - 1).float()
    for name, fn in (("rollout_random(64)", lambda: env.rollout_random(64)),
                     ("64 x step(actions)", lambda: [env.step(act) for _ in range(64)])):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        t0.record()
        for _ in range(3):
            fn()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 3
        print(f"SG_SUBBATCHES={os.environ.get('SG_SUBBATCHES', '1')} SG_GRAPHS={os.environ.get('SG_GRAPHS', '-')} "
              f"n={n} {name}: {n * 64 / ms * 1e-6:.3f} G env-steps/s ({ms / 64 * 1e3:.1f} us/step)")
    env.close()

import time, torch, sys
sys.path.insert(0, '.')
import spacegym_b200 as sg
for n in (262144, 1048576):
    for mode in ("fused", "split"):
        env = sg.make("GoalContinuous3P-v0", n, seed=1, step_mode=mode)
        env.reset(); env.rollout_random(64); torch.cuda.synchronize()
        for K in (8, 64):
            for want in (False, True):
                t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
                reps = 5
                t0.record()
                for _ in range(reps):
                    if want: env.step_many(K=K)
                    else: env.rollout_random(K)
                t1.record(); torch.cuda.synchronize()
                ms = t0.elapsed_time(t1) / reps
                print(f"n={n} mode={mode} K={K} outputs={want}: {n*K/ms*1e-6:.3f} G env-steps/s ({ms/K*1e3:.1f} us/step)")
        env.close()

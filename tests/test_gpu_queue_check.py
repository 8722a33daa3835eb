"""Sanitizer substitute: compute-sanitizer is not available on the GPU pool, so the work queues of the
split step (64-bit packed atomics, cross-kernel counters emptied by the tail kernel's last block) are checked by a
-DSG_CHECK_QUEUES build of the same engine that counts invariant violations on the device
(include/spacegym_b200.h: sg_debug_queue_check)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK_LIB = os.path.join(ROOT, "spacegym_b200", "libspacegym_b200_check.so")

SCRIPT = r"""
import ctypes as C, json, sys
import numpy as np, torch
sys.path.insert(0, %r)
import spacegym_b200 as sg
out = {}
for n, mes in ((20000, 40), (33, 25), (262144, 500)):
    env = sg.make("GoalContinuous3P-v0", n, seed=5, max_episode_steps=mes, copy=False)
    env.reset()
    gen = torch.Generator(device="cuda:0").manual_seed(1)
    steps = 120 if n < 100000 else 40
    for k in range(steps):
        a = (torch.rand((n, 2), device="cuda:0", generator=gen) * 2 - 1).float()
        env.step(a)
    env.rollout_random(30)                      # synthetic actions, K steps per call
    env.reset(mask=torch.arange(n, device="cuda:0") %% 3 == 0)
    env.step_many(K=10)
    buf = (C.c_uint64 * 8)()
    rc = env._L.sg_debug_queue_check(env._h, buf)
    st = env.stats()
    out[str(n)] = dict(rc=rc, counters=list(buf), steps=steps + 40, episodes=st["episodes"],
                       faults=int((env.state_views()["status"] != 0).sum()))
    env.close()
print("RESULT " + json.dumps(out))
"""


@pytest.mark.gpu
def test_queue_invariants_hold_in_the_check_build():
    assert os.path.exists(CHECK_LIB), "build it with `python -m spacegym_b200.build` (or __graft_entry__.build())"
    env = dict(os.environ, SPACEGYM_B200_LIB=CHECK_LIB)
    p = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    for n, r in res.items():
        assert r["rc"] == 0, f"{n} envs: sg_debug_queue_check rc {r['rc']}"
        assert r["counters"][:6] == [0] * 6, f"{n} envs: queue invariant violations {r['counters'][:6]}"
        assert r["counters"][6] == r["steps"] and r["counters"][7] == 1  # every split step was checked
        assert r["episodes"] > 0 and r["faults"] == 0


@pytest.mark.gpu
def test_normal_build_reports_the_check_as_not_compiled_in():
    import ctypes as C

    import spacegym_b200 as sg
    env = sg.make("GoalContinuous3P-v0", 64, seed=1)
    buf = (C.c_uint64 * 8)()
    assert env._L.sg_debug_queue_check(env._h, buf) != 0
    env.close()

"""Amortised env-steps/s of FeedingJaco-v1 INCLUDING the resets at the end of every 200-step episode:
the plain vector env (reset inside step, the batch waits) against the double-buffered one (a standby copy of the batch is
re-randomised by a background thread while the live one is stepped).  `python tools/gpu_amortised.py [N] [episodes]`"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from assistive_gym_b200.vec_env import AssistiveVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
episodes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
out = {}
for mode in (False, True):
    v = AssistiveVecEnv('assistive_gym:FeedingJaco-v1', n_envs=n, seed=3, double_buffer=mode)
    t0 = time.perf_counter()
    v.reset()
    first_reset = time.perf_counter() - t0
    g = torch.Generator(device='cuda').manual_seed(0)
    acts = torch.rand((200, n, 7), generator=g, device='cuda') * 2 - 1
    for i in range(5):
        v.step(acts[i])
    torch.cuda.synchronize()
    v._t = 0
    t0 = time.perf_counter()
    for ep in range(episodes):
        for i in range(200):
            v.step(acts[i])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out['double_buffered' if mode else 'plain'] = {'env_steps_per_s': n * 200 * episodes / dt, 'seconds': dt, 'first_reset_s': first_reset}
    v.close()
print(json.dumps({'n_envs': n, 'episodes': episodes, **out}))

import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import peg_in_hole_gym_b200 as gym
env = gym.make('peg-in-hole-mp-v0', client=gym.DIRECT, task='peg-in-hole', mp_num=4, sub_num=1024, offset=[2., 3., 0.],
               max_steps=200, auto_reset=True, shard=False)
env.reset()
a_np = env.action_space.sample_array()
a_dev = torch.from_numpy(a_np).cuda()
def timeit(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e6
print("env.step(device tensor)        %7.1f us" % timeit(lambda: env.step(a_dev)))
print("env.step(numpy [mp,sub,A])     %7.1f us" % timeit(lambda: env.step(a_np)))
def full():
    o, r, d, i = env.step(a_np); return o.cpu(), r.cpu(), d.cpu()
print("env.step(numpy) + .cpu() x3    %7.1f us" % timeit(full))
lst = env.action_space.sample()
print("env.step(nested lists)         %7.1f us" % timeit(lambda: env.step(lst), 50))

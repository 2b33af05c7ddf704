"""Rollout bookkeeping on the device: GAE (the numpy tail of Runner.run, ppo_multi_agent_new.py:205-218)."""
import ctypes as C

import torch

from . import _lib


def gae(rewards, values, dones, last_values, last_dones, gamma=0.99, lam=0.95):
    """Generalised advantage estimation for a [T, N] rollout held on the GPU.

    rewards, values: float32 [T, N]; dones: bool/uint8 [T, N], dones[t] = flag BEFORE step t
    (the reference's mb_dones); last_values float32 [N], last_dones bool [N] (after the last step).
    Returns (advantages, returns), float32 [T, N], bit-exact with the reference's numpy loop."""
    dev = rewards.device
    T, N = rewards.shape
    r = rewards.contiguous().float()
    v = values.contiguous().float()
    d = dones.to(torch.uint8).contiguous()
    lv = last_values.contiguous().float()
    ld = last_dones.to(torch.uint8).contiguous()
    advs = torch.empty_like(r)
    rets = torch.empty_like(r)
    p = lambda t: C.c_void_p(t.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(_lib.lib().snk_gae(p(r), p(v), p(d), p(lv), p(ld), float(gamma), float(lam), int(T), int(N), p(advs), p(rets),
                                  int(dev.index or 0), stream))
    return advs, rets

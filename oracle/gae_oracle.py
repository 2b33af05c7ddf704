"""CPU oracle of the rollout's advantage estimation -- TEST INFRASTRUCTURE.

Restates the reversed loop at the end of the reference's Runner.run
(/root/reference/src/ppo_multi_agent_new.py:205-218) with the same array dtypes, so that numpy
performs the same mixed-precision arithmetic: rewards / values float32, done flags bool (so
`1.0 - flags` is float64), gamma and lam python floats.

Pinned: the reference module needs TensorFlow to import, but its GAE statements are pure numpy; oracle/ref_gae.py lifts
them out of Runner.run with `ast` and EXECUTES them.  tests/golden/gae_golden.npz holds their outputs (written by
oracle/make_gae_golden.py in the build container); tests/test_gae_oracle.py checks this restatement against the golden
file everywhere and against the lifted reference statements wherever /root/reference exists."""
import numpy as np


def gae(rewards, values, dones, last_values, last_dones, gamma, lam):
    rewards = np.asarray(rewards, dtype=np.float32)
    values = np.asarray(values, dtype=np.float32)
    dones = np.asarray(dones, dtype=bool)
    last_values = np.asarray(last_values, dtype=np.float32)
    last_dones = np.asarray(last_dones, dtype=bool)
    T = rewards.shape[0]
    advs = np.zeros_like(rewards)
    running = 0
    for t in range(T - 1, -1, -1):
        if t == T - 1:
            alive_next, value_next = 1.0 - last_dones, last_values
        else:
            alive_next, value_next = 1.0 - dones[t + 1], values[t + 1]
        td_error = rewards[t] + gamma * value_next * alive_next - values[t]
        running = td_error + gamma * lam * alive_next * running
        advs[t] = running
    return advs, advs + values

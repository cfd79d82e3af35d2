"""TEST INFRASTRUCTURE ONLY -- restatement of the reference A2C/A3C return + GAE loop.

Follows machin/frame/algorithms/a2c.py:275-313.  Python floats (float64); the critic
value of each step is what ``critic(state).item()`` yields (a float32 widened to double).
"""


def discounted_returns(reward, discount):
    # a2c.py:275-281
    n = len(reward)
    out = [0.0] * n
    out[-1] = reward[-1]
    for i in range(n - 1, 0, -1):
        out[i - 1] = out[i] * discount + reward[i - 1]
    return out


def gae(reward, terminal, critic_value, next_critic_value, returns, discount, lam):
    """reward/terminal/critic_value/next_critic_value/returns: python sequences (one episode)."""
    n = len(reward)
    if lam == 1.0:                                  # a2c.py:284-286
        return [returns[i] - critic_value[i] for i in range(n)]
    if lam == 0.0:                                  # a2c.py:287-295
        return [reward[i] + discount * (1 - float(terminal[i])) * next_critic_value[i]
                - critic_value[i] for i in range(n)]
    out = [0.0] * n                                 # a2c.py:296-313
    last_v = 0
    last_a = 0
    for i in range(n - 1, -1, -1):
        delta = reward[i] + discount * last_v * (1 - float(terminal[i])) - critic_value[i]
        last_v = critic_value[i]
        last_a = out[i] = last_a * discount * (1 - float(terminal[i])) * lam + delta
    return out

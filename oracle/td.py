"""TEST INFRASTRUCTURE ONLY -- restatement of the reference TD target / priority math.

Follows machin/frame/algorithms/dqn_per.py:105-139 + dqn.py:497-501 (double DQN) and
ddpg_per.py:117-160 + ddpg.py:506-510 (DDPG), as CPU torch eager ops in the same order.
"""
import torch as t


def dqn_per_target(q, q_next, q_tgt_next, action, reward, terminal, discount):
    """q, q_next, q_tgt_next: f32[B,A]; action i64[B,1]; reward f32[B,1]; terminal bool[B,1].
    Returns (action_value[B,1], y[B,1], abs_error[B])."""
    a_star = t.max(q_next, dim=1)[1].unsqueeze(1)           # dqn_per.py:108-110
    tgt = q_tgt_next.gather(dim=1, index=a_star)
    action_value = q.gather(dim=1, index=action.long())    # :115-120
    y = reward + discount * ~terminal * tgt                 # dqn.py:501
    abs_error = t.sum(t.abs(action_value - y), dim=1).flatten()  # dqn_per.py:132-134
    return action_value, y, abs_error


def per_weighted_loss(elementwise_loss, is_weight_f64):
    # dqn_per.py:126-130 -- float64 IS weights promote the product before the mean
    return (elementwise_loss * is_weight_f64.view(-1, 1)).mean()


def ddpg_target(next_value, reward, terminal, discount):
    # ddpg_per.py:124-128 + ddpg.py:510
    next_value = next_value.view(reward.shape[0], -1)
    return reward + discount * ~terminal * next_value


def abs_error(value, y):
    # ddpg_per.py:153-159 / dqn_per.py:132-138
    return t.sum(t.abs(value - y.type_as(value)), dim=1).flatten()


def soft_update_(target_params, source_params, rate):
    # algorithms/utils.py:20-27
    with t.no_grad():
        for tp, sp in zip(target_params, source_params):
            tp.copy_(tp * (1.0 - rate) + sp * rate)

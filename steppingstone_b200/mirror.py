"""Left/right mirror augmentation for PPO minibatches.

``get_mirror_indices`` restates walker3d_envs.py:1358-1411 (= mike.py:868-921) with the joint index sets of
robots.py:514-524; ``get_mirror_function`` restates common/envs_utils.py:687-740 as ONE gather + sign
multiply per tensor (the reference does clone + fancy-index negate + fancy-index swap).
"""
import numpy as np

ACTION_DIM = 21
RIGHT_JOINTS = np.array([3, 4, 5, 6, 7, 13, 14, 15, 16], dtype=np.int64)      # robots.py:514-517
LEFT_JOINTS = np.array([8, 9, 10, 11, 12, 17, 18, 19, 20], dtype=np.int64)    # robots.py:518-521
NEGATION_JOINTS = np.array([0, 2], dtype=np.int64)                            # robots.py:514 (abdomen_z, abdomen_x)


def get_mirror_indices():
    right_obs = np.concatenate((6 + RIGHT_JOINTS, 6 + RIGHT_JOINTS + ACTION_DIM, [6 + 2 * ACTION_DIM]))
    left_obs = np.concatenate((6 + LEFT_JOINTS, 6 + LEFT_JOINTS + ACTION_DIM, [6 + 2 * ACTION_DIM + 1]))
    negation_obs = np.array([2, 4, 6, 6, 8, 27, 29, 50, 53, 55, 58], dtype=np.int64)
    return negation_obs, right_obs, left_obs, NEGATION_JOINTS.copy(), RIGHT_JOINTS.copy(), LEFT_JOINTS.copy()


def _perm_sign(dim, neg, right, left):
    perm = np.arange(dim)
    perm[np.concatenate((right, left))] = np.concatenate((left, right))
    sign = np.ones(dim, dtype=np.float32)
    sign[neg] = -1.0
    # reference order: negate first, then swap  ->  out[:, i] = (sign * x)[:, perm[i]]
    return perm, sign[perm]


def get_mirror_function(indices, device=None):
    import torch
    neg_o, r_o, l_o, neg_a, r_a, l_a = indices
    po, so = _perm_sign(60, neg_o, r_o, l_o)
    pa, sa = _perm_sign(ACTION_DIM, neg_a, r_a, l_a)
    po_t, so_t = torch.as_tensor(po, device=device), torch.as_tensor(so, device=device)
    pa_t, sa_t = torch.as_tensor(pa, device=device), torch.as_tensor(sa, device=device)

    def mirror_function(trajectory_samples):
        obs, states, actions, value_preds, returns, masks, old_logp, adv = trajectory_samples
        obs_m = obs.index_select(1, po_t.to(obs.device)) * so_t.to(obs.device)
        act_m = actions.index_select(1, pa_t.to(actions.device)) * sa_t.to(actions.device)
        rep = lambda t: t.repeat((2, 1))
        return (torch.cat([obs, obs_m]), rep(states), torch.cat([actions, act_m]), rep(value_preds), rep(returns),
                rep(masks), rep(old_logp), rep(adv))

    return mirror_function

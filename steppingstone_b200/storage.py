"""Device-resident ``RolloutStorage`` (algorithms/storage.py:5-103), same attributes and methods.

Differences from the reference, none visible to ``algorithms/ppo.py``:
  * buffers are created on the device (the reference builds them on the host and calls ``.cuda()``,
    storage.py:20-29; ``insert`` then copies host->device every step, :31-51);
  * ``compute_returns(use_gae=True)`` is one ``ss_gae`` kernel launch (time scan per env) instead of a
    Python loop over T (:59-71);
  * ``feed_forward_generator`` draws one device permutation instead of a ``SubsetRandomSampler`` (:84-103).
"""
import ctypes

import torch

from . import native


class RolloutStorage(object):
    def __init__(self, num_steps, num_processes, obs_shape, action_dim, state_size, device=None):
        dev = torch.device("cuda") if device is None else torch.device(device)
        z = lambda *s: torch.zeros(*s, device=dev)
        self.observations = z(num_steps + 1, num_processes, *obs_shape)
        self.states = z(num_steps + 1, num_processes, state_size)
        self.rewards = z(num_steps, num_processes, 1)
        self.value_preds = z(num_steps + 1, num_processes, 1)
        self.returns = z(num_steps + 1, num_processes, 1)
        self.action_log_probs = z(num_steps, num_processes, 1)
        self.actions = z(num_steps, num_processes, action_dim)
        self.masks = torch.ones(num_steps + 1, num_processes, 1, device=dev)
        self.bad_masks = torch.ones(num_steps + 1, num_processes, 1, device=dev)
        self.num_steps = num_steps
        self.step = 0
        self.device = dev

    def cuda(self):
        return self          # already resident (reference: storage.py:20-29)

    def insert(self, current_obs, state, action, action_log_prob, value_pred, reward, mask, bad_mask):
        self.observations[self.step + 1].copy_(current_obs, non_blocking=True)
        self.states[self.step + 1].copy_(state, non_blocking=True)
        self.actions[self.step].copy_(action, non_blocking=True)
        self.action_log_probs[self.step].copy_(action_log_prob, non_blocking=True)
        self.value_preds[self.step].copy_(value_pred, non_blocking=True)
        self.rewards[self.step].copy_(reward, non_blocking=True)
        self.masks[self.step + 1].copy_(mask, non_blocking=True)
        self.bad_masks[self.step + 1].copy_(bad_mask, non_blocking=True)
        self.step = (self.step + 1) % self.num_steps

    def after_update(self):
        self.observations[0].copy_(self.observations[-1])
        self.states[0].copy_(self.states[-1])
        self.masks[0].copy_(self.masks[-1])
        self.bad_masks[0].copy_(self.bad_masks[-1])

    def compute_returns(self, next_value, use_gae, gamma, gae_lambda):
        T, N = self.rewards.shape[0], self.rewards.shape[1]
        if use_gae:
            self.value_preds[-1] = next_value
            if self.rewards.is_cuda:
                lib = native.load()
                st = ctypes.c_void_p(torch.cuda.current_stream(self.rewards.device).cuda_stream)
                native.check(lib.ss_gae(self.rewards.data_ptr(), self.value_preds.data_ptr(), self.masks.data_ptr(),
                                        self.bad_masks.data_ptr(), self.returns.data_ptr(), T, N, float(gamma),
                                        float(gae_lambda), st), "ss_gae")
            else:
                raise RuntimeError("RolloutStorage.compute_returns needs CUDA tensors (no CPU path)")
        else:
            self.returns[-1] = next_value
            for step in reversed(range(T)):
                self.returns[step] = ((self.returns[step + 1] * gamma * self.masks[step + 1] + self.rewards[step])
                                      * self.bad_masks[step + 1] + (1 - self.bad_masks[step + 1]) * self.value_preds[step])

    def feed_forward_generator(self, advantages, num_mini_batch):
        num_steps, num_processes = self.rewards.size()[0:2]
        batch_size = num_processes * num_steps
        mini_batch_size = batch_size // num_mini_batch
        perm = torch.randperm(batch_size, device=self.rewards.device)
        obs = self.observations[:-1].view(-1, *self.observations.size()[2:])
        states = self.states[:-1].view(-1, self.states.size(-1))
        actions = self.actions.view(-1, self.actions.size(-1))
        values = self.value_preds[:-1].view(-1, 1)
        returns = self.returns[:-1].view(-1, 1)
        masks = self.masks[:-1].view(-1, 1)
        logp = self.action_log_probs.view(-1, 1)
        adv = advantages.view(-1, 1)
        for start in range(0, batch_size, mini_batch_size):      # BatchSampler(drop_last=False)
            idx = perm[start:start + mini_batch_size]
            yield obs[idx], states[idx], actions[idx], values[idx], returns[idx], masks[idx], logp[idx], adv[idx]

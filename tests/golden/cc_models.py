"""Small actor-critic MLPs shared by the fixture generator (built on the reference's ``parl.Model``) and the GPU
tests (built on ``parl_b200.Model``): the layer shapes follow benchmark/torch/{ddpg,td3,sac}/mujoco_model.py with a
hidden width of 32 instead of 256/400/300 so that the fixtures stay small."""
import torch
import torch.nn as nn
import torch.nn.functional as F

HID = 32
LOG_SIG_MAX, LOG_SIG_MIN = 2.0, -20.0


def make_models(Model):
    class _Actor(Model):
        def __init__(self, obs_dim, act_dim, gaussian):
            super().__init__()
            self.gaussian = gaussian
            self.l1 = nn.Linear(obs_dim, HID)
            self.l2 = nn.Linear(HID, HID)
            self.l3 = nn.Linear(HID, act_dim)
            if gaussian:
                self.std = nn.Linear(HID, act_dim)

        def forward(self, obs):
            x = F.relu(self.l2(F.relu(self.l1(obs))))
            if self.gaussian:                                   # sac/mujoco_model.py:56-63
                return self.l3(x), torch.clamp(self.std(x), min=LOG_SIG_MIN, max=LOG_SIG_MAX)
            return torch.tanh(self.l3(x))                       # td3/mujoco_model.py:57-60 (max_action = 1)

    class _Critic(Model):
        def __init__(self, obs_dim, act_dim, twin):
            super().__init__()
            self.twin = twin
            self.l1 = nn.Linear(obs_dim + act_dim, HID)
            self.l2 = nn.Linear(HID, HID)
            self.l3 = nn.Linear(HID, 1)
            if twin:
                self.l4 = nn.Linear(obs_dim + act_dim, HID)
                self.l5 = nn.Linear(HID, HID)
                self.l6 = nn.Linear(HID, 1)

        def forward(self, obs, action):
            x = torch.cat([obs, action], 1)
            q1 = self.l3(F.relu(self.l2(F.relu(self.l1(x)))))
            if not self.twin:
                return q1
            return q1, self.l6(F.relu(self.l5(F.relu(self.l4(x)))))

        def Q1(self, obs, action):
            x = torch.cat([obs, action], 1)
            return self.l3(F.relu(self.l2(F.relu(self.l1(x)))))

    class ACModel(Model):
        """kind: 'ddpg' (single critic, tanh actor), 'td3' (twin critic, tanh actor), 'sac' (twin critic, Gaussian)."""

        def __init__(self, obs_dim, act_dim, kind):
            super().__init__()
            self.actor_model = _Actor(obs_dim, act_dim, kind == 'sac')
            self.critic_model = _Critic(obs_dim, act_dim, kind != 'ddpg')

        def policy(self, obs):
            return self.actor_model(obs)

        def value(self, obs, action):
            return self.critic_model(obs, action)

        def Q1(self, obs, action):
            return self.critic_model.Q1(obs, action)

        def get_actor_params(self):
            return self.actor_model.parameters()

        def get_critic_params(self):
            return self.critic_model.parameters()

    return ACModel

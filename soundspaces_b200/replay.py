"""Trace-replay environment for BASELINE.json config 5 (DD-PPO rollout with the audio observation fused into the
step), used where habitat-lab / habitat-sim are not installed ("a fake env that replays recorded
(rir_id, action) traces").

What is real here is everything ON the audio path, driven exactly as the reference drives it:

* :class:`ReplaySim` is a simulator object with the attributes and the ``step`` semantics of the discrete
  ``SoundSpacesSim`` (grid graph, MOVE_FORWARD / TURN_LEFT / TURN_RIGHT / STOP, soundspaces/simulator.py:478-566,
  azimuth rule :568-573, memo dicts replaced on scene / sound change :395-397) with the audio slice supplied by
  :class:`~soundspaces_b200.simulator.B200AudioMixin`;
* observations come from this package's ``SpectrogramSensor`` (deferred handles) through ``batch_obs`` into the
  ``RolloutStorage`` slot ``observations["spectrogram"][step + 1]`` (ss_baselines/common/rollout_storage.py:27-35,88-91);
* the consumer is an ``AudioCNN``-shaped encoder (ss_baselines/av_nav/models/audio_cnn.py:31-89: three convolutions,
  flatten, linear) with a 4-way policy head -- ordinary PyTorch policy code, not part of this library's kernels.

What is replayed instead of simulated: the visual sensors, rewards and episode logic (none of which touch audio).
"""
from __future__ import annotations

import os
import time
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from .simulator import AudioRenderService, B200AudioMixin

STOP, MOVE_FORWARD, TURN_LEFT, TURN_RIGHT = 0, 1, 2, 3          # HabitatSimActions


class ReplayScene:
    """A navigation graph on a square grid (node = row * side + col, 1 m spacing in the X-Z plane) and the bank ids
    of its binaural RIRs keyed like the reference's wav tree ``<dir>/<azimuth>/<receiver>_<source>.wav``."""

    def __init__(self, name: str, side: int, rir_root: str = "/replay", dataset: str = "replica"):
        import networkx as nx
        self.name, self.side = name, int(side)
        self.rir_dir = os.path.join(rir_root, dataset, name)
        g = nx.Graph()
        for r in range(side):
            for c in range(side):
                g.add_node(r * side + c, point=(float(c), 0.0, float(r)))
        for r in range(side):
            for c in range(side):
                if c + 1 < side:
                    g.add_edge(r * side + c, r * side + c + 1)
                if r + 1 < side:
                    g.add_edge(r * side + c, (r + 1) * side + c)
        self.graph = g

    @property
    def n_nodes(self) -> int:
        return self.side * self.side

    def register_rirs(self, service: AudioRenderService, source: int, rirs, azimuths=(0, 90, 180, 270)):
        """Make the scene's RIRs device-resident: ``rirs[az_index][receiver]`` -> one batched upload; afterwards
        ``service.rir((rir_dir, az, receiver, source))`` is a hit, as after the reference's first visit + N1 prefetch."""
        flat, keys = [], []
        for ai, az in enumerate(azimuths):
            for recv in range(self.n_nodes):
                flat.append(rirs[ai][recv])
                keys.append((self.rir_dir, az, recv, source))
        ids = service.renderer.add_rirs(flat)
        for k, i in zip(keys, ids):
            service._rir_ids[k] = i
            service._touched.setdefault(k, 0)
        return ids


class ReplaySim(B200AudioMixin):
    """The audio-relevant state machine of ``SoundSpacesSim`` driven by actions (see module docstring)."""

    def __init__(self, service: AudioRenderService, scene: ReplayScene, sound: str, clip: np.ndarray, source_node: int,
                 start_node: int = 0, start_rotation: int = 0, deferred: bool = True, duration: int = 500):
        sr = service.sr
        self._b200_svc = service
        self.b200_deferred = deferred
        self.config = SimpleNamespace(
            USE_RENDERED_OBSERVATIONS=True, SCENE_DATASET="replica",
            AUDIO=SimpleNamespace(RIR_SAMPLING_RATE=sr, HAS_DISTRACTOR_SOUND=False, EVERLASTING=True))
        self._source_sound_dict: Dict[str, np.ndarray] = {}
        self._current_sound = None
        self._current_scene = None
        self._audiogoal_cache, self._spectrogram_cache = dict(), dict()
        self._duration = duration
        self.reconfigure(scene, sound, clip, source_node, start_node, start_rotation)

    # -- what the reference's reconfigure() does to the audio state (simulator.py:340-397) ------------------------
    def reconfigure(self, scene: ReplayScene, sound: str, clip: np.ndarray, source_node: int, start_node: int,
                    start_rotation: int):
        sr = self.config.AUDIO.RIR_SAMPLING_RATE
        self._audio_index = 0
        same_sound = sound == self._current_sound
        if not same_sound:
            self._current_sound = sound
            self._source_sound_dict.setdefault(sound, clip)
        self._audio_length = self._source_sound_dict[self._current_sound].shape[0] // sr
        same_scene = scene is self._current_scene
        if not same_scene:
            self._current_scene = scene
            self.graph = scene.graph
        if not same_scene or not same_sound:
            self._audiogoal_cache = dict()
            self._spectrogram_cache = dict()
        self._episode_step_count = 0
        self._receiver_position_index = int(start_node)
        self._source_position_index = int(source_node)
        self._rotation_angle = int(start_rotation) % 360
        self._is_episode_active = True

    @property
    def binaural_rir_dir(self):
        return self._current_scene.rir_dir

    @property
    def current_source_sound(self):
        return self._source_sound_dict[self._current_sound]

    @property
    def azimuth_angle(self):
        return -(self._rotation_angle + 0) % 360            # simulator.py:568-573

    @property
    def is_silent(self):
        return self._episode_step_count > self._duration

    def get_orientation(self):
        return (270 - self._rotation_angle) % 360            # simulator.py:563-565

    def step(self, action: int):
        """simulator.py:496-516."""
        if action == STOP:
            self._is_episode_active = False
        elif action == MOVE_FORWARD:
            here = self.graph.nodes[self._receiver_position_index]["point"]
            for nb in self.graph[self._receiver_position_index]:
                p2 = self.graph.nodes[nb]["point"]
                direction = int(np.around(np.rad2deg(np.arctan2(p2[2] - here[2], p2[0] - here[0])))) % 360
                if direction == self.get_orientation():
                    self._receiver_position_index = nb
                    break
        elif action == TURN_LEFT:
            self._rotation_angle = (self._rotation_angle + 90) % 360
        elif action == TURN_RIGHT:
            self._rotation_angle = (self._rotation_angle - 90) % 360
        self._episode_step_count += 1


class ReplayVectorEnv:
    """``n`` in-process envs stepped like ``SyncVectorEnv`` (ss_baselines/common/sync_vector_env.py:186-199): every
    env's sensor suite is asked for its observation once per step; the spectrogram sensors return handles."""

    def __init__(self, sims: Sequence[ReplaySim]):
        from .sensors import SpectrogramSensor
        self.sims = list(sims)
        self.sensors = [SpectrogramSensor(sim=s, config=None) for s in self.sims]
        self.num_envs = len(self.sims)

    def observe(self) -> List[dict]:
        return [{"spectrogram": sensor.get_observation(observations=None, episode=None)} for sensor in self.sensors]

    def step(self, actions: Sequence[int]):
        for sim, a in zip(self.sims, actions):
            if a == STOP or not sim._is_episode_active:       # episode over: a new one starts where it stands
                sim._is_episode_active, sim._episode_step_count = True, 0
            else:
                sim.step(int(a))
        return self.observe()


class AudioPolicy(nn.Module):
    """``AudioCNN`` (audio_cnn.py:31-89: Conv 8x8/4 -> 4x4/2 -> 3x3/1, or the 5/3/3 variant for small inputs, flatten,
    linear, ReLU) feeding a 4-action categorical head and a value head.  ``channels_first``: the observation already
    is (N, 2, 65, T') -- what the reference obtains with ``permute(0, 3, 1, 2)`` (audio_cnn.py:86)."""

    def __init__(self, spec_shape, hidden: int = 512, channels_first: bool = False):
        super().__init__()
        h, w, c = spec_shape
        if h < 30 or w < 30:
            ks, st = [(5, 5), (3, 3), (3, 3)], [(2, 2), (2, 2), (1, 1)]
        else:
            ks, st = [(8, 8), (4, 4), (3, 3)], [(4, 4), (2, 2), (1, 1)]
        dims = [h, w]
        for k, s in zip(ks, st):
            dims = [int(np.floor((d - (kk - 1) - 1) / ss + 1)) for d, kk, ss in zip(dims, k, s)]
        self.channels_first = channels_first
        self.cnn = nn.Sequential(
            nn.Conv2d(c, 32, ks[0], st[0]), nn.ReLU(True), nn.Conv2d(32, 64, ks[1], st[1]), nn.ReLU(True),
            nn.Conv2d(64, 64, ks[2], st[2]), nn.Flatten(), nn.Linear(64 * dims[0] * dims[1], hidden), nn.ReLU(True))
        self.actor, self.critic = nn.Linear(hidden, 4), nn.Linear(hidden, 1)

    def forward(self, spectrogram: torch.Tensor):
        x = spectrogram if self.channels_first else spectrogram.permute(0, 3, 1, 2)
        f = self.cnn(x)
        return self.actor(f), self.critic(f)

    @torch.no_grad()
    def forward_fused(self, spectrogram: torch.Tensor):
        """Rollout-time forward with the first layer fused: permute + Conv2d + ReLU of ``cnn[0:2]``
        run as ONE kernel straight from the (N, 65, T', 2) observation (``WaveformOps.audio_conv1``)."""
        from .renderer import WaveformOps
        if self.channels_first or not spectrogram.is_cuda:
            return self.forward(spectrogram)
        f = self.cnn[2:](WaveformOps.get(spectrogram.device).audio_conv1(spectrogram.contiguous(), self.cnn[0], relu=True))
        return self.actor(f), self.critic(f)

    @torch.no_grad()
    def act(self, spectrogram: torch.Tensor, fused: bool = False):
        logits, value = self.forward_fused(spectrogram) if fused else self.forward(spectrogram)
        dist = torch.distributions.Categorical(logits=logits)
        actions = dist.sample()
        return value, actions, dist.log_prob(actions)


def collect_rollout(envs: ReplayVectorEnv, policy: AudioPolicy, storage: torch.Tensor, num_steps: int,
                    forced_actions: Optional[np.ndarray] = None, fused: bool = False):
    """One rollout of ``num_steps`` env steps with the reference's timers (ppo_trainer.py:125-194): ``pth_time`` =
    action sampling + observation batching / insertion, ``env_time`` = ``envs.step``.  ``storage``: the
    ``(num_steps + 1, num_envs, 65, T', 2)`` tensor ``rollouts.observations["spectrogram"]``; the step's batch is
    rendered straight into ``storage[step + 1]`` by ``batch_obs`` (nothing for ``RolloutStorage.insert`` to copy).
    ``forced_actions`` (``(num_steps, num_envs)``): replay a recorded action trace instead of the sampled actions.
    ``fused``: run the policy's first layer with the fused permute + Conv2d + ReLU kernel."""
    from .sensors import batch_obs
    dev = storage.device
    pth_time = env_time = 0.0
    batch_obs(envs.observe(), device=dev, out={"spectrogram": storage[0]})
    for step in range(num_steps):
        t0 = time.time()
        _, actions, _ = policy.act(storage[step], fused=fused)
        acts = actions.tolist() if forced_actions is None else forced_actions[step].tolist()
        pth_time += time.time() - t0
        t0 = time.time()
        observations = envs.step(acts)
        env_time += time.time() - t0
        t0 = time.time()
        batch_obs(observations, device=dev, out={"spectrogram": storage[step + 1]})
        pth_time += time.time() - t0
    torch.cuda.synchronize(dev) if dev.type == "cuda" else None
    return pth_time, env_time, num_steps * envs.num_envs

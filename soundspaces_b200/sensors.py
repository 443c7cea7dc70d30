"""Habitat sensor plugins for the audio observation + the batch boundary.

Mirrors ``soundspaces/tasks/nav.py:37-105`` (``AudioGoalSensor``, ``SpectrogramSensor``: same
class names, uuids ``"audiogoal"`` / ``"spectrogram"``, ``SensorTypes.PATH``, ``spaces.Box`` of
float32 whose shape is obtained by running ``compute_spectrogram`` on ``np.ones((2, sr))``) and
``ss_baselines/common/utils.py:117-153`` (``to_tensor`` / ``batch_obs``).

When habitat-lab is importable the classes subclass ``habitat.core.simulator.Sensor`` and are
registered with ``@registry.register_sensor``; otherwise a minimal stand-in base with the same
protocol is used (habitat is not installable in the build image).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from .renderer import BatchedAudioRenderer, spectrogram_shape
from .simulator import SPECTROGRAM_NATIVE_ATTR

try:  # pragma: no cover - habitat is absent in the build image
    from habitat.core.registry import registry
    from habitat.core.simulator import Sensor, SensorTypes
    from gym import spaces
    HAVE_HABITAT = True
except Exception:  # noqa: BLE001
    HAVE_HABITAT = False

    class _Registry:
        def __init__(self):
            self.sensors = {}

        def register_sensor(self, cls=None, *, name=None):
            def wrap(c):
                self.sensors[name or c.__name__] = c
                return c
            return wrap(cls) if cls is not None else wrap

        def get_sensor(self, name):
            return self.sensors.get(name)

    registry = _Registry()

    class SensorTypes:
        PATH = "PATH"

    class _Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    class spaces:  # noqa: N801
        Box = _Box

    class Sensor:
        """Protocol of habitat.core.simulator.Sensor (v0.2.2)."""

        def __init__(self, *args: Any, **kwargs: Any) -> None:
            self.config = kwargs["config"] if "config" in kwargs else None
            self.uuid = self._get_uuid(*args, **kwargs)
            self.sensor_type = self._get_sensor_type(*args, **kwargs)
            self.observation_space = self._get_observation_space(*args, **kwargs)


def _device_spectrogram(audio_data, pad_mode="reflect", device="cuda:0"):
    """compute_spectrogram on the GPU for a host (2, n) array; result on the host."""
    a = np.asarray(audio_data)
    if a.ndim != 2 or a.shape[0] != 2:
        raise ValueError(f"audio_data must be (2, n), got {a.shape}")
    from .simulator import AudioRenderService
    r = AudioRenderService.get(a.shape[1], device, pad_mode=pad_mode).renderer
    wave = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(r.device)
    out = r.spectrogram(wave[None])[0].cpu().numpy()
    return out.astype(a.dtype) if a.dtype == np.float64 else out


@registry.register_sensor
class AudioGoalSensor(Sensor):
    def __init__(self, *args: Any, sim, config, **kwargs: Any):
        self._sim = sim
        super().__init__(config=config)

    def _get_uuid(self, *args: Any, **kwargs: Any):
        return "audiogoal"

    def _get_sensor_type(self, *args: Any, **kwargs: Any):
        return SensorTypes.PATH

    def _get_observation_space(self, *args: Any, **kwargs: Any):
        sensor_shape = (2, self._sim.config.AUDIO.RIR_SAMPLING_RATE)
        return spaces.Box(low=np.finfo(np.float32).min, high=np.finfo(np.float32).max,
                          shape=sensor_shape, dtype=np.float32)

    def get_observation(self, *args: Any, observations=None, episode=None, **kwargs: Any):
        return self._sim.get_current_audiogoal_observation()


@registry.register_sensor
class SpectrogramSensor(Sensor):
    cls_uuid: str = "spectrogram"
    pad_mode: str = "reflect"
    device: str = "cuda:0"

    def __init__(self, *args: Any, sim, config, **kwargs: Any):
        self._sim = sim
        super().__init__(config=config)

    def _get_uuid(self, *args: Any, **kwargs: Any):
        return "spectrogram"

    def _get_sensor_type(self, *args: Any, **kwargs: Any):
        return SensorTypes.PATH

    def _get_observation_space(self, *args: Any, **kwargs: Any):
        # nav.py:77 runs compute_spectrogram on ones((2, sr)) only for its shape
        shape = spectrogram_shape(self._sim.config.AUDIO.RIR_SAMPLING_RATE)
        return spaces.Box(low=np.finfo(np.float32).min, high=np.finfo(np.float32).max,
                          shape=shape, dtype=np.float32)

    @staticmethod
    def compute_spectrogram(audio_data):
        """nav.py:86-100 on the GPU: host (2, n) in, host (65, T', 2) out.  Imported by
        savi/ppo/ppo_trainer.py:52 and belief_predictor.py:15, so it keeps working on host arrays."""
        return _device_spectrogram(audio_data, SpectrogramSensor.pad_mode, SpectrogramSensor.device)

    def get_observation(self, *args: Any, observations=None, episode=None, **kwargs: Any):
        return self._sim.get_current_spectrogram_observation(self.compute_spectrogram)


setattr(SpectrogramSensor.compute_spectrogram, SPECTROGRAM_NATIVE_ATTR, True)


@registry.register_sensor(name="Intensity")
class Intensity(Sensor):
    """AV-WaN ``Intensity`` sensor (ss_baselines/av_wan/avwan_sensors.py:69-100): mean square of the 150
    samples after the onset of the current waveform, reduced on the device."""

    def __init__(self, sim, config, *args: Any, **kwargs: Any):
        self._sim = sim
        super().__init__(config=config)

    def _get_uuid(self, *args: Any, **kwargs: Any):
        return "intensity"

    def _get_sensor_type(self, *args: Any, **kwargs: Any):
        return getattr(SensorTypes, "COLOR", "COLOR")

    def _get_observation_space(self, *args: Any, **kwargs: Any):
        return spaces.Box(low=0, high=1, shape=(1,), dtype=bool)

    def get_observation(self, *args: Any, observations=None, episode=None, **kwargs: Any):
        from .simulator import AudioRenderService
        audiogoal = np.asarray(self._sim.get_current_audiogoal_observation())
        r = AudioRenderService.get(audiogoal.shape[1], getattr(self._sim, "b200_device", "cuda:0"),
                                   pad_mode=getattr(self._sim, "b200_pad_mode", "reflect")).renderer
        rms = r.intensity(torch.from_numpy(np.ascontiguousarray(audiogoal, dtype=np.float32))[None])
        return [float(rms[0])]


def to_tensor(v):
    # ss_baselines/common/utils.py:117-123
    if torch.is_tensor(v):
        return v
    elif isinstance(v, np.ndarray):
        return torch.from_numpy(v)
    else:
        return torch.tensor(v, dtype=torch.float)


def batch_obs(observations: List[Dict], device: Optional[torch.device] = None, skip_list=[]) -> Dict[str, torch.Tensor]:
    """ss_baselines/common/utils.py:126-153 with one change: sensor values that already are CUDA
    tensors on ``device`` (the renderer's output rows) are stacked on the device, so the
    spectrogram batch never takes the D2H -> H2D round trip."""
    batch = defaultdict(list)
    for obs in observations:
        for sensor in obs:
            if sensor in skip_list:
                continue
            batch[sensor].append(to_tensor(obs[sensor]).float())
    for sensor in batch:
        batch[sensor] = torch.stack(batch[sensor], dim=0).to(device=device, dtype=torch.float)
    return batch


class VectorAudioObservations:
    """Batch the audio observation of many in-process envs into ONE render call.

    ``collect(sims)`` builds one request per env through the same attribute reads as the
    reference's per-env sensors (including the memo caches keyed ``(source, receiver, azimuth)``,
    simulator.py:683-699: only cache-missing envs are rendered) and returns the
    ``(n_envs, 65, T', 2)`` CUDA tensor the policy consumes -- the place where the reference does
    N separate CPU renders, N pickles and one ``batch_obs`` H2D copy.
    """

    def __init__(self, sr: int, device="cuda:0", pad_mode="reflect"):
        from .simulator import AudioRenderService
        self.service = AudioRenderService.get(sr, device, pad_mode=pad_mode)
        self.renderer: BatchedAudioRenderer = self.service.renderer
        self._cache: Dict[tuple, torch.Tensor] = {}

    def clear_cache(self):
        self._cache.clear()

    def collect(self, sims) -> torch.Tensor:
        r = self.renderer
        n = len(sims)
        self.service.maybe_trim()
        out = torch.empty((n,) + r.spec_shape, dtype=torch.float32, device=r.device)
        todo, reqs, keys = [], [], []
        for i, sim in enumerate(sims):
            cacheable = not sim.config.AUDIO.HAS_DISTRACTOR_SOUND
            key = (id(sim), sim._current_sound, sim._source_position_index, sim._receiver_position_index,
                   sim.azimuth_angle) if cacheable else None
            if sim._episode_step_count > sim._duration:
                out[i].zero_()
                continue
            hit = self._cache.get(key) if cacheable else None
            if hit is not None:
                out[i].copy_(hit)
                continue
            sim._b200_svc = self.service
            reqs.append(sim._b200_request())
            todo.append(i)
            keys.append(key)
        if reqs:
            spec = r.render(reqs)
            idx = torch.tensor(todo, device=r.device)
            out.index_copy_(0, idx, spec)
            for j, key in enumerate(keys):
                if key is not None:
                    self._cache[key] = spec[j]
        return out

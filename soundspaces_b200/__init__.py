"""soundspaces_b200 -- B200-native (sm_100a) audio-observation path for SoundSpaces.

Public surface (mirrors the reference names for this path):

* :class:`BatchedAudioRenderer`, :class:`AudioRequest` -- batched device renderer
* :mod:`soundspaces_b200.simulator` -- ``get_current_audiogoal_observation`` /
  ``get_current_spectrogram_observation`` drop-ins (soundspaces/simulator.py:678-701,
  soundspaces/continuous_simulator.py:458-462)
* :mod:`soundspaces_b200.sensors` -- ``AudioGoalSensor`` / ``SpectrogramSensor``
  (soundspaces/tasks/nav.py:37-105) and ``batch_obs`` (ss_baselines/common/utils.py:126-153)

Importing the package does not load CUDA; constructing a renderer does and fails
loudly if ``libssb200.so`` is missing.
"""
from .renderer import AudioRequest, BatchedAudioRenderer, spectrogram_shape  # noqa: F401

__all__ = ["AudioRequest", "BatchedAudioRenderer", "spectrogram_shape"]

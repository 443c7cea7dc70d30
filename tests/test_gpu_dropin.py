"""GPU tests of the drop-in boundary: the simulator mixins and Habitat sensor plugins driven like
the reference's own classes (bare simulator objects with the attributes of SURVEY.md App. D, RIR
wav trees on disk), compared with the oracle / the reference's golden outputs."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import audio_oracle as ao  # noqa: E402
from oracle.ref_harness import AttrDict, write_rir  # noqa: E402
from synth import make_rir, make_source  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def make_sim_class():
    from soundspaces_b200.simulator import B200AudioMixin

    class FakeSoundSpacesSim(B200AudioMixin):
        """Only the non-audio attributes/properties the reference class provides
        (simulator.py:303-335, :568-573)."""

        @property
        def binaural_rir_dir(self):
            return os.path.join(self.config.AUDIO.BINAURAL_RIR_DIR, self.config.SCENE_DATASET, "apartment_0")

        @property
        def current_source_sound(self):
            return self._source_sound_dict[self._current_sound]

        @property
        def azimuth_angle(self):
            return -(self._rotation_angle + 0) % 360

        @property
        def is_silent(self):
            return self._episode_step_count > self._duration

    return FakeSoundSpacesSim


def make_sim(rir_root, sr, sounds, *, rotation=0, step_count=0, audio_index=0, distractor=None,
             distractor_sound=None, receiver=0, source=1):
    sim = make_sim_class()()
    sim.config = AttrDict(USE_RENDERED_OBSERVATIONS=True, SCENE_DATASET="replica",
                          AUDIO=AttrDict(RIR_SAMPLING_RATE=sr, HAS_DISTRACTOR_SOUND=distractor is not None,
                                         BINAURAL_RIR_DIR=rir_root, EVERLASTING=True))
    sim._episode_step_count, sim._duration = step_count, 500
    sim._rotation_angle = rotation
    sim._receiver_position_index, sim._source_position_index = receiver, source
    sim._current_sound = "telephone.wav"
    sim._source_sound_dict = dict(sounds)
    sim._audio_index = audio_index
    sim._audio_length = sounds["telephone.wav"].shape[0] // sr
    sim._audiogoal_cache, sim._spectrogram_cache = {}, {}
    if distractor is not None:
        sim._distractor_position_index, sim._current_distractor_sound = distractor, distractor_sound
    return sim


def check_wave(got, ref, stride=1):
    got, ref = np.asarray(got, np.float64)[:, ::stride], np.asarray(ref, np.float64)
    peak = np.abs(ref).max()
    assert (np.abs(got - ref).max() <= 1e-4 * peak) if peak else (not got.any())


@pytest.mark.parametrize("name", sorted(mg.DISCRETE_CASES))
def test_discrete_simulator_dropin(golden, name, tmp_path):
    from soundspaces_b200.sensors import AudioGoalSensor, SpectrogramSensor
    c = mg.DISCRETE_CASES[name]
    src, rir, dsrc, drir = mg.discrete_inputs(c)
    sr, rot = c["sr"], c.get("rotation_angle", 0)
    az = -rot % 360
    d = str(tmp_path)
    if c.get("rir_mode") == "unreadable":
        write_rir(d, "replica", "apartment_0", az, 0, 1, sr, None)
    elif c.get("rir_mode") == "empty":
        write_rir(d, "replica", "apartment_0", az, 0, 1, sr, np.zeros((0, 2), np.float32))
    else:
        write_rir(d, "replica", "apartment_0", az, 0, 1, sr, rir)
    sounds, kw = {"telephone.wav": src}, {}
    if dsrc is not None:
        sounds["distractor.wav"] = dsrc
        write_rir(d, "replica", "apartment_0", az, 0, 2, sr, drir)
        kw = dict(distractor=2, distractor_sound="distractor.wav")
    sim = make_sim(d, sr, sounds, rotation=rot, step_count=c.get("step_count", 0),
                   audio_index=c.get("audio_index", 0), **kw)
    cfg = AttrDict()
    wave_sensor = AudioGoalSensor(sim=sim, config=cfg)
    spec_sensor = SpectrogramSensor(sim=sim, config=cfg)
    assert wave_sensor.uuid == "audiogoal" and spec_sensor.uuid == "spectrogram"
    assert wave_sensor.observation_space.shape == (2, sr)
    assert spec_sensor.observation_space.shape == ao.spectrogram_shape(sr)

    wave = wave_sensor.get_observation(observations={}, episode=None)
    gw = golden.get(f"{name}/wave")
    stride = 1
    if gw is None:
        gw, stride = golden[f"{name}/wave_stride5"], 5
    check_wave(wave, gw, stride)
    assert wave.shape == (2, sr) and str(wave.dtype) == str(golden[f"{name}/wave_dtype"])
    if dsrc is None:
        assert int(golden[f"{name}/audio_index_after"]) == sim._audio_index
        # second call is a memo hit: same object, index not advanced again (simulator.py:683-686)
        assert wave_sensor.get_observation(observations={}, episode=None) is wave
        assert int(golden[f"{name}/audio_index_after"]) == sim._audio_index
        spec = spec_sensor.get_observation(observations={}, episode=None)
        assert np.allclose(spec, golden[f"{name}/spec_reflect"], rtol=1e-4, atol=1e-5)
        assert spec_sensor.get_observation(observations={}, episode=None) is spec
    # fresh simulator, spectrogram first (fused device path, cache miss)
    sim2 = make_sim(d, sr, sounds, rotation=rot, step_count=c.get("step_count", 0),
                    audio_index=c.get("audio_index", 0), **kw)
    spec2 = SpectrogramSensor(sim=sim2, config=cfg).get_observation(observations={}, episode=None)
    if dsrc is None:
        assert np.allclose(spec2, golden[f"{name}/spec_reflect"], rtol=1e-4, atol=1e-5)
        assert int(golden[f"{name}/audio_index_after"]) == sim2._audio_index
    else:
        ref = ao.compute_spectrogram(ao.compute_audiogoal(src, rir, sr, distractor=dsrc, distractor_rir=drir))
        assert np.allclose(spec2, ref, rtol=1e-4, atol=1e-5)
    if c.get("step_count", 0) > 500:
        assert not np.any(spec2) and not np.any(wave)


def test_compute_spectrogram_static_on_host_array(golden):
    """savi imports SpectrogramSensor.compute_spectrogram and calls it on host arrays."""
    from soundspaces_b200.sensors import SpectrogramSensor
    w = golden["a2_16k/wave"]
    s = SpectrogramSensor.compute_spectrogram(w)
    assert isinstance(s, np.ndarray) and s.shape == (65, 26, 2)
    assert np.allclose(s, golden["a2_16k/spec_reflect"], rtol=1e-4, atol=1e-5)
    assert not SpectrogramSensor.compute_spectrogram(np.zeros((2, 16000))).any()


def test_custom_callable_gets_host_waveform(tmp_path):
    sr = 16000
    src, rir = make_source(1, sr), make_rir(1, 5000)
    write_rir(str(tmp_path), "replica", "apartment_0", 0, 0, 1, sr, rir)
    sim = make_sim(str(tmp_path), sr, {"telephone.wav": src})
    seen = {}

    def fn(audio):
        seen["a"] = audio
        return audio.sum(axis=1)
    out = sim.get_current_spectrogram_observation(fn)
    assert isinstance(seen["a"], np.ndarray) and seen["a"].shape == (2, sr) and out.shape == (2,)


@pytest.mark.parametrize("name", sorted(mg.CONTINUOUS_CASES))
def test_continuous_simulator_dropin(golden, name):
    from soundspaces_b200.sensors import SpectrogramSensor
    from soundspaces_b200.simulator import B200ContinuousAudioMixin
    c = mg.CONTINUOUS_CASES[name]
    src, rir, last = mg.continuous_inputs(c)

    class FakeContinuousSim(B200ContinuousAudioMixin):
        @property
        def current_source_sound(self):
            return self._source_sound_dict[self._current_sound]

    sim = FakeContinuousSim()
    sim.config = AttrDict(STEP_TIME=0.25, AUDIO=AttrDict(RIR_SAMPLING_RATE=c["sr"], CROSSFADE=last is not None))
    sim._episode_step_count, sim._duration = c.get("step_count", 0), 500
    sim._current_sound, sim._source_sound_dict = "s", {"s": src}
    sim._current_sample_index = c["sample_index"]
    sim._prev_sim_obs = {"audio_sensor": np.asarray(rir).T.tolist()}
    sim._last_rir = last
    wave = sim.get_current_audiogoal_observation()
    check_wave(wave, golden[f"{name}/wave"])
    spec = sim.get_current_spectrogram_observation(SpectrogramSensor.compute_spectrogram)
    assert np.allclose(spec, golden[f"{name}/spec_reflect"], rtol=1e-4, atol=1e-5)
    n0 = len(sim._b200_service().renderer._rir_off)
    sim.get_current_audiogoal_observation()
    assert len(sim._b200_service().renderer._rir_off) == n0          # transient RIRs are released


def test_vector_batching_and_batch_obs(tmp_path):
    from soundspaces_b200.sensors import VectorAudioObservations, batch_obs
    sr, n = 16000, 6
    src = make_source(2, sr)
    d = str(tmp_path)
    rirs = [make_rir(10 + i, 3000 + 500 * i) for i in range(n)]
    sims = []
    for i in range(n):
        write_rir(d, "replica", "apartment_0", 0, i, 1, sr, rirs[i])
        sims.append(make_sim(d, sr, {"telephone.wav": src}, receiver=i, step_count=501 if i == 4 else 0))
    vec = VectorAudioObservations(sr)
    out = vec.collect(sims)
    assert out.is_cuda and out.shape == (n, 65, 26, 2)
    launches = vec.renderer.ctx.launch_count
    again = vec.collect(sims)                                  # all memo hits: no kernels
    assert vec.renderer.ctx.launch_count == launches and torch.equal(out, again)
    for i in range(n):
        ref = ao.compute_spectrogram(ao.compute_audiogoal(src, rirs[i], sr, silent=(i == 4)).astype(np.float32))
        assert np.allclose(out[i].cpu().numpy(), ref, rtol=1e-4, atol=1e-5)
    batch = batch_obs([{"spectrogram": out[i], "x": np.float32(i)} for i in range(n)], device=out.device)
    assert batch["spectrogram"].is_cuda and torch.equal(batch["spectrogram"], out)


def test_rir_bank_budget_trim(tmp_path):
    """The file-backed bank is dropped between steps when it outgrows its budget and refilled on demand."""
    from soundspaces_b200.simulator import AudioRenderService
    sr = 16000
    src = make_source(4, sr)
    d = str(tmp_path)
    rirs = [make_rir(30 + i, 4000) for i in range(4)]
    for i in range(4):
        write_rir(d, "replica", "apartment_0", 0, i, 1, sr, rirs[i])
    svc = AudioRenderService(sr, max_bank_bytes=2 * 4000 * 8 + 1)       # room for two RIRs
    outs = []
    for i in range(4):
        sim = make_sim(d, sr, {"telephone.wav": src}, receiver=i)
        sim._b200_svc = svc
        outs.append(sim.get_current_audiogoal_observation())
        assert svc.renderer.bank_bytes <= 3 * 4000 * 8
    for i in range(4):
        ref = ao.compute_audiogoal(src, rirs[i], sr)
        assert np.abs(outs[i] - ref).max() <= 1e-4 * np.abs(ref).max()


# ----------------------------------------------------------------------------------------------
# round 2: deferred handles behind the sensor API, batch_obs replacement, scene-safe memo, device Intensity
# ----------------------------------------------------------------------------------------------
def _fresh_service(sr, **kw):
    from soundspaces_b200.simulator import AudioRenderService
    return AudioRenderService(sr, **kw)


def test_deferred_sensor_path_one_render_per_step(tmp_path):
    """Per-env SpectrogramSensor.get_observation returns handles; batch_obs renders the whole step ONCE and hands
    back the (N, 65, T', 2) CUDA view; values equal the oracle; memo hits launch nothing."""
    from soundspaces_b200.batching import DeferredObservation
    from soundspaces_b200.sensors import SpectrogramSensor, batch_obs
    sr, n = 16000, 7
    src = make_source(5, sr)
    d = str(tmp_path)
    rirs = [make_rir(40 + i, 2000 + 700 * i) for i in range(n)]
    svc = _fresh_service(sr)
    sims, sensors = [], []
    for i in range(n):
        write_rir(d, "replica", "apartment_0", 0, i, 1, sr, rirs[i])
        sim = make_sim(d, sr, {"telephone.wav": src}, receiver=i, step_count=501 if i == 2 else 0)
        sim.b200_deferred, sim._b200_svc = True, svc
        sims.append(sim)
        sensors.append(SpectrogramSensor(sim=sim, config=AttrDict()))
    l0 = svc.renderer.ctx.launch_count
    obs = [{"spectrogram": s.get_observation(observations={}, episode=None), "gps": np.float32([i, 0])}
           for i, s in enumerate(sensors)]
    assert all(isinstance(o["spectrogram"], DeferredObservation) and o["spectrogram"].pending for o in obs)
    assert svc.renderer.ctx.launch_count == l0                       # nothing launched yet
    batch = batch_obs(obs, device=svc.renderer.device)
    spec = batch["spectrogram"]
    assert spec.is_cuda and spec.shape == (n, 65, 26, 2) and svc.batcher.flushes == 1
    assert spec.data_ptr() == svc.batcher.ring.data_ptr()            # a view of the ring: no stack, no copy
    per_step = svc.renderer.ctx.launch_count - l0
    for i in range(n):
        ref = ao.compute_spectrogram(ao.compute_audiogoal(src, rirs[i], sr, silent=(i == 2)).astype(np.float32))
        assert np.allclose(spec[i].cpu().numpy(), ref, rtol=1e-4, atol=1e-5)
    assert not spec[2].any()                                          # silence is exactly 0
    # same positions again: all memo hits (the sims' own dicts), no kernels
    l1 = svc.renderer.ctx.launch_count
    obs2 = [{"spectrogram": s.get_observation(observations={}, episode=None)} for s in sensors]
    assert all(o2["spectrogram"] is o["spectrogram"] for o, o2 in zip(obs, obs2))
    assert torch.equal(batch_obs(obs2, device=svc.renderer.device)["spectrogram"], spec)
    assert svc.renderer.ctx.launch_count == l1
    # rendering straight into a rollout-storage slot
    for s in sims:
        s._spectrogram_cache = dict()
    slot = torch.zeros((n, 65, 26, 2), device=svc.renderer.device)
    obs3 = [{"spectrogram": s.get_observation(observations={}, episode=None)} for s in sensors]
    got = batch_obs(obs3, device=svc.renderer.device, out={"spectrogram": slot})["spectrogram"]
    assert got.data_ptr() == slot.data_ptr() and torch.equal(slot, spec)
    assert per_step <= 12                                             # one launch chain, not one per env
    # host compat of a handle: the array the reference's sensor returns
    assert np.allclose(np.asarray(obs3[1]["spectrogram"]), spec[1].cpu().numpy())


def test_scene_change_with_equal_node_indices_is_not_a_memo_hit(tmp_path):
    """ADVICE r1 (high) / VERDICT weak #2 on the real kernels: same (source, receiver, azimuth), other scene."""
    from soundspaces_b200.sensors import VectorAudioObservations
    sr, n = 16000, 3
    src = make_source(6, sr)
    d = str(tmp_path)
    rirs = {s: [make_rir(seed + i, 3000 + 100 * i) for i in range(n)] for s, seed in (("apartment_0", 100), ("office_3", 200))}
    for scene in rirs:
        for i in range(n):
            write_rir(d, "replica", scene, 0, i, 1, sr, rirs[scene][i])
    Sim = make_sim_class()

    class SceneSim(Sim):
        scene = "apartment_0"

        @property
        def binaural_rir_dir(self):
            return os.path.join(self.config.AUDIO.BINAURAL_RIR_DIR, self.config.SCENE_DATASET, self.scene)

    sims = []
    for i in range(n):
        base = make_sim(d, sr, {"telephone.wav": src}, receiver=i)
        sim = SceneSim()
        sim.__dict__.update(base.__dict__)
        sims.append(sim)
    vec = VectorAudioObservations(sr)
    a = vec.collect(sims).clone()
    for s in sims:                                   # reconfigure() to another scene (simulator.py:395-397)
        s.scene = "office_3"
        s._audiogoal_cache, s._spectrogram_cache = dict(), dict()
    b = vec.collect(sims).clone()
    for i in range(n):
        ra = ao.compute_spectrogram(ao.compute_audiogoal(src, rirs["apartment_0"][i], sr).astype(np.float32))
        rb = ao.compute_spectrogram(ao.compute_audiogoal(src, rirs["office_3"][i], sr).astype(np.float32))
        assert np.allclose(a[i].cpu().numpy(), ra, rtol=1e-4, atol=1e-5)
        assert np.allclose(b[i].cpu().numpy(), rb, rtol=1e-4, atol=1e-5), "stale observation from the previous scene"


def test_intensity_sensor_stays_on_device(tmp_path):
    from soundspaces_b200.sensors import Intensity
    sr = 16000
    src, rir = make_source(8, sr), make_rir(8, 4000)
    write_rir(str(tmp_path), "replica", "apartment_0", 0, 0, 1, sr, rir)
    sim = make_sim(str(tmp_path), sr, {"telephone.wav": src})
    sim._b200_svc = _fresh_service(sr)
    val = Intensity(sim, AttrDict()).get_observation(observations={}, episode=None)
    wave = ao.compute_audiogoal(src, rir, sr).astype(np.float32)
    nonzero = (wave > 0.1 * wave.max()).argmax(axis=1).min()
    ref = float(np.mean(wave[:, nonzero: nonzero + 150] ** 2))
    assert isinstance(val, list) and abs(val[0] - ref) <= 1e-4 * ref
    assert not sim._audiogoal_cache                   # the waveform never visited the host memo
    w = sim.get_current_audiogoal_device()
    assert w.is_cuda and sim.get_current_audiogoal_device() is w
    sim._audiogoal_cache = dict()                     # scene / sound change resets the device memo too
    assert sim.get_current_audiogoal_device() is not w


def test_continuous_early_branch_does_not_wrap(golden):
    """index < len(rir) and index + num_sample > len(clip): zeros past the clip end, not wrapped audio."""
    from soundspaces_b200.simulator import B200ContinuousAudioMixin
    sr = 16000
    src = make_source(9, sr)[:15000].copy()           # short clip so that the window runs past its end
    rir = make_rir(9, 14000)

    class Sim(B200ContinuousAudioMixin):
        @property
        def current_source_sound(self):
            return self._source_sound_dict[self._current_sound]

    sim = Sim()
    sim.config = AttrDict(STEP_TIME=0.25, AUDIO=AttrDict(RIR_SAMPLING_RATE=sr, CROSSFADE=False))
    sim._episode_step_count, sim._duration = 0, 500
    sim._current_sound, sim._source_sound_dict = "s", {"s": src}
    sim._current_sample_index = 12000                 # 12000 - 14000 < 0: early branch; 12000 + 4000 > 15000
    sim._prev_sim_obs = {"audio_sensor": np.asarray(rir).T.tolist()}
    sim._last_rir = None
    wave = sim.get_current_audiogoal_observation()
    ref = ao.continuous_convolve_with_rir(src.astype(np.float64), rir.astype(np.float64), sr, 0.25, 12000)
    check_wave(wave, ref)


def test_render_on_second_renderer_leaves_current_device_alone():
    """constructing / using a renderer must not switch the caller's current CUDA device."""
    from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
    cur = torch.cuda.current_device()
    r = BatchedAudioRenderer(16000, 4000, device=f"cuda:{torch.cuda.device_count() - 1}")
    assert torch.cuda.current_device() == cur
    sid, rid = r.add_source(make_source(1, 16000)), r.add_rirs([make_rir(1, 3000)])[0]
    spec = r.render([AudioRequest(rir=rid, source=sid)])
    torch.cuda.synchronize(r.device)
    assert torch.cuda.current_device() == cur and spec.device == r.device
    ref = ao.compute_spectrogram(ao.compute_audiogoal(make_source(1, 16000), make_rir(1, 3000), 16000).astype(np.float32))
    assert np.allclose(spec[0].cpu().numpy(), ref, rtol=1e-4, atol=1e-5)

"""Config 5 (DD-PPO rollout with the audio observation fused into the step) on the trace-replay env: host logic on
the CPU with a stand-in renderer, and -- on the GPU -- every stored rollout slot against the oracle."""
import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao
from stubs import StubRenderer
from synth import make_rir, make_source

from soundspaces_b200.replay import (MOVE_FORWARD, STOP, TURN_LEFT, TURN_RIGHT, AudioPolicy, ReplayScene, ReplaySim,
                                     ReplayVectorEnv, collect_rollout)
from soundspaces_b200.simulator import AudioRenderService


def test_replay_sim_follows_reference_step_semantics():
    """soundspaces/simulator.py:496-516, :563-573: forward moves to the neighbour in the facing direction, turns change
    the azimuth by 90 degrees, azimuth = (-rotation) % 360."""
    svc = AudioRenderService(16000, renderer=StubRenderer(16000))
    scene = ReplayScene("s", side=3)
    sim = ReplaySim(svc, scene, "t.wav", make_source(0, 16000), source_node=8, start_node=4, start_rotation=0)
    assert sim.azimuth_angle == 0 and sim.get_orientation() == 270          # faces -Z: row - 1
    sim.step(MOVE_FORWARD)
    assert sim._receiver_position_index == 1
    sim.step(MOVE_FORWARD)
    assert sim._receiver_position_index == 1                                # edge of the grid: collision, stays
    sim.step(TURN_LEFT)
    assert sim._rotation_angle == 90 and sim.azimuth_angle == 270 and sim.get_orientation() == 180   # faces -X
    sim.step(MOVE_FORWARD)
    assert sim._receiver_position_index == 0
    sim.step(TURN_RIGHT); sim.step(TURN_RIGHT)
    assert sim._rotation_angle == 270 and sim.get_orientation() == 0        # faces +X
    sim.step(MOVE_FORWARD)
    assert sim._receiver_position_index == 1 and sim._episode_step_count == 7
    sim.step(STOP)
    assert not sim._is_episode_active
    # scene / sound change replaces the memo dicts (simulator.py:395-397); same scene and sound keeps them
    cache = sim._spectrogram_cache
    sim.reconfigure(scene, "t.wav", sim.current_source_sound, 8, 0, 0)
    assert sim._spectrogram_cache is cache
    sim.reconfigure(ReplayScene("other", side=3), "t.wav", sim.current_source_sound, 8, 0, 0)
    assert sim._spectrogram_cache is not cache and sim._episode_step_count == 0


def test_collect_rollout_host_logic_one_render_per_step():
    sr = 16000
    svc = AudioRenderService(sr, renderer=StubRenderer(sr))
    scene = ReplayScene("s", side=4)
    rirs = [[np.zeros((10 + az + 4 * n, 2), np.float32) for n in range(scene.n_nodes)] for az in range(4)]
    scene.register_rirs(svc, source=0, rirs=rirs)
    clip = make_source(0, sr)
    sims = [ReplaySim(svc, scene, "t.wav", clip, source_node=0, start_node=i, start_rotation=90 * (i % 4)) for i in range(5)]
    envs = ReplayVectorEnv(sims)
    policy = AudioPolicy((65, 26, 2))
    steps = 12
    storage = torch.zeros((steps + 1, 5, 65, 26, 2))
    trace = np.random.default_rng(0).choice([1, 2, 3], size=(steps, 5))
    pth, env_t, n = collect_rollout(envs, policy, storage, steps, trace)
    assert n == steps * 5 and pth > 0 and env_t > 0
    assert len(svc.renderer.renders) <= steps + 1                           # ONE render per step at most (memo hits: none)
    assert svc.stats["misses"] == 0                                         # every RIR was resident
    # every stored slot holds the row of the RIR at that env's (node, azimuth) at that step (stub: 1000 * rir id)
    replay = [ReplaySim(svc, scene, "t.wav", clip, source_node=0, start_node=i, start_rotation=90 * (i % 4)) for i in range(5)]
    for step in range(steps + 1):
        for e, s in enumerate(replay):
            rid = svc._rir_ids[(scene.rir_dir, s.azimuth_angle, s._receiver_position_index, 0)]
            assert float(storage[step, e, 0, 0, 0]) == 1000.0 * rid
            if step < steps:
                s.step(int(trace[step, e]))


@pytest.mark.gpu
def test_rollout_slots_equal_the_oracle():
    """the observation stored in rollouts.observations['spectrogram'][step + 1] equals the oracle's
    spectrogram of the RIR at the env's node / heading at that step."""
    sr, taps, n_envs, steps = 16000, 3000, 6, 10
    svc = AudioRenderService(sr, device="cuda:0", max_taps=taps, n_terms=1)
    scene = ReplayScene("gpu_scene", side=3)
    rirs = [[make_rir(100 * az + n, taps - 7 * n) for n in range(scene.n_nodes)] for az in range(4)]
    scene.register_rirs(svc, source=0, rirs=rirs)
    clip = make_source(5, sr)
    mk = lambda: [ReplaySim(svc, scene, "t.wav", clip, source_node=0, start_node=i, start_rotation=90 * (i % 4)) for i in range(n_envs)]
    envs = ReplayVectorEnv(mk())
    policy = AudioPolicy(svc.renderer.spec_shape).cuda()
    storage = torch.zeros((steps + 1, n_envs) + svc.renderer.spec_shape, device="cuda")
    trace = np.random.default_rng(1).choice([1, 1, 2, 3], size=(steps, n_envs))
    f0 = svc.batcher.flushes
    collect_rollout(envs, policy, storage, steps, trace)
    assert svc.batcher.flushes - f0 <= steps + 1
    got = storage.cpu().numpy()
    shadow = mk()
    for s in shadow:
        s.b200_deferred = False
    az_index = {0: 0, 90: 1, 180: 2, 270: 3}
    for step in range(steps + 1):
        for e, s in enumerate(shadow):
            rir = rirs[az_index[s.azimuth_angle]][s._receiver_position_index]
            ref = ao.compute_spectrogram(ao.compute_audiogoal(clip, rir, sr).astype(np.float32))
            assert np.allclose(got[step, e], ref, rtol=1e-4, atol=1e-5), (step, e)
            if step < steps:
                s.step(int(trace[step, e]))
    logits, value = policy(storage[3])
    assert logits.shape == (n_envs, 4) and value.shape == (n_envs, 1) and torch.isfinite(logits).all()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(65, 69, 2), (65, 26, 2)])
def test_fused_first_conv_matches_torch(shape):
    """permute + first Conv2d + ReLU of AudioCNN as one kernel from the (N, 65, T', 2) observation, against
    PyTorch's float64 convolution of the permuted tensor (audio_cnn.py:51-58,86); then the whole fused forward."""
    from soundspaces_b200.renderer import WaveformOps
    torch.manual_seed(0)
    policy = AudioPolicy(shape).cuda()
    conv = policy.cnn[0]
    assert conv.kernel_size == ((8, 8) if shape[1] >= 30 else (5, 5))
    spec = torch.rand((9,) + shape, device="cuda") * 3.0
    got = WaveformOps.get("cuda:0").audio_conv1(spec, conv, relu=True)
    ref = torch.relu(torch.nn.functional.conv2d(spec.permute(0, 3, 1, 2).double().cpu(), conv.weight.double().cpu(),
                                                conv.bias.double().cpu(), stride=conv.stride))
    assert got.shape == ref.shape
    assert torch.allclose(got.cpu().double(), ref, rtol=1e-5, atol=1e-5)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        la, va = policy.forward(spec)
        lb, vb = policy.forward_fused(spec)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert torch.allclose(la, lb, rtol=1e-4, atol=1e-4) and torch.allclose(va, vb, rtol=1e-4, atol=1e-4)

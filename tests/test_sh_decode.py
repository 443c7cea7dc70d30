"""Ambisonic -> binaural decode: the LTI oracle against the outputs recorded from the
reference's closed AmbisonicBinauralizer (tests/golden/sh_golden.npz, made by make_sh_golden.py), and
the CUDA path against the oracle."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import sh_oracle as so

_spec = importlib.util.spec_from_file_location(
    "make_sh_golden", os.path.join(os.path.dirname(__file__), "golden", "make_sh_golden.py"))
msg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(msg)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shg():
    with np.load(os.path.join(ROOT, "tests", "golden", "sh_golden.npz")) as z:
        return {k: z[k] for k in z.files}


def test_bank_matches_packaged_data(shg):
    bank = np.load(os.path.join(ROOT, "soundspaces_b200", "data", "sh_hrtf_bank.npy"))
    assert bank.shape == (9, 2, 256) and np.array_equal(bank, shg["bank"])
    # ears mirror-symmetric, sign flips for ACN 3, 4, 7 (inter-aural axis = X)
    for k in range(9):
        sign = -1.0 if k in (3, 4, 7) else 1.0
        assert np.allclose(bank[k, 1], sign * bank[k, 0], atol=1e-6)
    energy = (bank[:, 0].astype(np.float64) ** 2).sum(axis=1)
    assert np.allclose(energy, [6.289, 0.609, 0.621, 4.007, 0.338, 0.163, 0.106, 0.266, 0.390], atol=2e-3)


def test_rotation_matrix_properties():
    for az in (0.0, 30.0, 90.0, 123.4, 270.0):
        R = so.rotation_matrix(az)
        assert np.allclose(R @ R.T, np.eye(9), atol=1e-12)
        assert np.allclose(so.rotation_matrix(-az), R.T, atol=1e-12)
    R = so.rotation_matrix(90.0)                      # swaps 1<->3, 5<->7, flips 4 and 8
    assert abs(R[1, 3]) == pytest.approx(1) and abs(R[5, 7]) == pytest.approx(1)
    assert R[4, 4] == pytest.approx(-1) and R[8, 8] == pytest.approx(-1) and R[0, 0] == R[2, 2] == R[6, 6] == 1


@pytest.mark.parametrize("name", sorted(msg.CASES))
def test_oracle_vs_tool_output(shg, name):
    az, seed = msg.CASES[name]
    ref = shg[f"{name}/out"]
    got = so.sh_decode(msg.make_amb(seed), az, shg["bank"])
    assert got.shape == ref.shape == (msg.N, 2)
    # the tool is only approximately LTI (block convolver): 5e-3 of peak
    assert np.abs(got - ref).max() <= 5e-3 * np.abs(ref).max()


def test_oracle_exact_on_block_aligned_impulse(shg):
    imp = np.zeros((1024, 9), np.float32)
    imp[384, 3] = 1.0
    assert np.abs(so.sh_decode(imp, 90.0, shg["bank"]) - shg["imp_az90/out"]).max() < 1e-6


@pytest.mark.gpu
def test_cuda_sh_decode_matches_oracle_and_tool(shg):
    import torch
    from soundspaces_b200 import BatchedAudioRenderer
    r = BatchedAudioRenderer(16000, 4096, device="cuda:0")
    names = sorted(msg.CASES)
    amb = np.stack([msg.make_amb(msg.CASES[n][1]) for n in names])
    az = [msg.CASES[n][0] for n in names]
    out = r.sh_decode(torch.from_numpy(amb), az).cpu().numpy()
    assert out.shape == (len(names), msg.N, 2)
    for i, n in enumerate(names):
        lti = so.sh_decode(amb[i], az[i], shg["bank"])
        assert np.abs(out[i] - lti).max() <= 1e-4 * np.abs(lti).max()                     # vs the LTI restatement
        assert np.abs(out[i] - shg[f"{n}/out"]).max() <= 5e-3 * np.abs(shg[f"{n}/out"]).max()   # vs the tool itself
    # long responses take the FFT (overlap-save) path: against the oracle and against the direct-form kernel
    rng = np.random.default_rng(6)
    for L in (7424, 9001, 20000):
        amb_l = (rng.standard_normal((3, L, 9)) * np.exp(-np.arange(L) / (L / 4.0))[None, :, None] * 0.2).astype(np.float32)
        az_l = [0.0, 45.0, 270.0]
        fft_out = r.sh_decode(torch.from_numpy(amb_l), az_l).cpu().numpy()
        r.lib.ssb_set_debug(r.ctx.handle, 32)                       # force the direct form
        fir_out = r.sh_decode(torch.from_numpy(amb_l), az_l).cpu().numpy()
        r.lib.ssb_set_debug(r.ctx.handle, 0)
        for i in range(3):
            lti = so.sh_decode(amb_l[i], az_l[i], shg["bank"])
            peak = np.abs(lti).max()
            assert np.abs(fft_out[i] - lti).max() <= 1e-4 * peak
            assert np.abs(fir_out[i] - lti).max() <= 1e-4 * peak
            assert np.abs(fft_out[i] - fir_out[i]).max() <= 1e-5 * peak
    # ragged length (not a multiple of the tile) and a long IR
    rng = np.random.default_rng(3)
    long_amb = (rng.standard_normal((2, 5003, 9)) * 0.2).astype(np.float32)
    got = r.sh_decode(torch.from_numpy(long_amb), [12.5, 300.0]).cpu().numpy()
    for i, a in enumerate((12.5, 300.0)):
        lti = so.sh_decode(long_amb[i], a, shg["bank"])
        assert np.abs(got[i] - lti).max() <= 1e-4 * np.abs(lti).max()


@pytest.mark.gpu
def test_cuda_config4_chain(shg):
    """BASELINE config 4 shape of work: 9-channel ambisonic RIR -> binaural decode -> convolution with the
    source -> spectrogram, all on the device, against the oracle chain."""
    import torch
    from oracle import audio_oracle as ao
    from synth import make_source
    from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
    sr, L, n = 16000, 6000, 4
    r = BatchedAudioRenderer(sr, L, device="cuda:0")
    rng = np.random.default_rng(9)
    amb = (rng.standard_normal((n, L, 9)) * np.exp(-np.arange(L) / 900.0)[None, :, None] * 0.2).astype(np.float32)
    az = [0.0, 90.0, 180.0, 270.0]
    rirs = r.sh_decode(torch.from_numpy(amb), az)
    ids = r.set_dense_rir_bank(rirs)
    src = make_source(5, sr)
    sid = r.add_source(src)
    spec, wave = r.render([AudioRequest(rir=i, source=sid) for i in ids], want_wave=True)
    torch.cuda.synchronize()
    rirs_h = rirs.cpu().numpy()
    for i in range(n):
        rir_ref = so.sh_decode(amb[i], az[i], shg["bank"])
        assert np.abs(rirs_h[i] - rir_ref).max() <= 1e-4 * np.abs(rir_ref).max()
        w_ref, _ = ao.render_frame(src, rir_ref, sr)                     # whole chain, waveform level
        assert np.abs(wave[i].cpu().numpy() - w_ref).max() <= 1e-4 * np.abs(w_ref).max()
        # spectrogram stage on the SAME decoded RIR (each stage is held to its own tolerance; the log
        # spectrogram of quiet bins amplifies the 1e-6-of-peak difference between two fp32 decodes)
        _, s_ref = ao.render_frame(src, rirs_h[i], sr)
        assert np.allclose(spec[i].cpu().numpy(), s_ref, rtol=1e-4, atol=1e-5)

"""PSNR / SSIM (utils/metrics.py): the oracle against the reference's committed values on the CPU, the device kernel
against both on the GPU, and the eval / video loops end to end on a synthetic scene."""
import os

import numpy as np
import pytest
import torch

from helpers import golden, make_state_dict, oracle

import mipnerf_pl_b200 as mp

CASES = [f"{t}_{n}" for t in "abc" for n in ("near", "far")]


@pytest.mark.parametrize("case", CASES)
def test_oracle_metrics_match_reference(case):
    g = golden("metrics.npz")
    psnr, ssim = oracle.eval_errors(torch.from_numpy(g[f"{case}_pred"])[None], torch.from_numpy(g[f"{case}_target"])[None])
    assert float(psnr) == pytest.approx(float(g[f"{case}_psnr"]), rel=1e-6)
    assert float(ssim) == pytest.approx(float(g[f"{case}_ssim"]), rel=1e-6, abs=1e-7)
    assert np.array_equal(oracle.gaussian_window().numpy(), g["window"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_device_metrics_match_reference(case):
    g = golden("metrics.npz")
    pred, tgt = torch.from_numpy(g[f"{case}_pred"])[None].cuda(), torch.from_numpy(g[f"{case}_target"])[None].cuda()
    psnr, ssim = mp.eval_errors(pred, tgt)
    assert float(psnr) == pytest.approx(float(g[f"{case}_psnr"]), rel=2e-6)
    assert float(ssim) == pytest.approx(float(g[f"{case}_ssim"]), rel=2e-5, abs=2e-6)
    assert float(mp.ssim(pred.permute(0, 3, 1, 2), tgt.permute(0, 3, 1, 2))) == float(ssim)


@pytest.mark.gpu
def test_device_metrics_full_frame_and_identity():
    gen = torch.Generator().manual_seed(0)
    a = torch.rand(1, 800, 800, 3, generator=gen)
    b = (a + 0.05 * torch.randn(1, 800, 800, 3, generator=gen)).clamp(0, 1)
    want_p, want_s = oracle.eval_errors(a, b)
    got_p, got_s = mp.eval_errors(a.cuda(), b.cuda())
    assert float(got_p) == pytest.approx(float(want_p), rel=1e-5) and float(got_s) == pytest.approx(float(want_s), rel=1e-4)
    _, same = mp.eval_errors(a.cuda(), a.cuda())
    assert float(same) == pytest.approx(1.0, abs=1e-6)


@pytest.mark.gpu
def test_evaluate_and_render_path_loops(tmp_path):
    """eval.py / render_video.py loops on a tiny synthetic Blender scene: files written, metrics finite, frames shaped."""
    root = str(tmp_path / "scene")
    mp.write_synthetic_blender_scene(root, 3, 16, 16, seed=5)
    ds = mp.Blender(root, "test", white_bkgd=True, batch_type="single_image")
    system = mp.MipNeRFSystem(mp.default_hparams(**{"val.chunk_size": 128}), precision="bf16")
    system.mip_nerf.load_state_dict(make_state_dict(seed=0))
    system = system.cuda().eval()
    psnrs, ssims = mp.evaluate(system, ds, out_dir=str(tmp_path / "eval"), save_image=True, max_images=2)
    assert len(psnrs) == len(ssims) == 2 and all(np.isfinite(psnrs)) and all(-1 <= s <= 1 for s in ssims)
    assert os.path.exists(tmp_path / "eval" / "psnrs.txt") and os.path.exists(tmp_path / "eval" / "images" / "00001_rgb.png")
    out = mp.render_path(system.mip_nerf, n_poses=3, height=24, width=24, out_dir=str(tmp_path / "video"))
    assert out["frames"] == 3 and all(t > 0 for t in out["ms_per_frame"])
    assert os.path.exists(tmp_path / "video" / "00002_dist.png")
    assert mp.spheric_path(120).shape == (120, 3, 4)

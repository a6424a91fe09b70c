"""the Blender / multi-scale loaders and the converter against tests/golden/datasets.npz, which was
produced by the reference's own datasets/datasets.py and datasets/convert_blender_data.py on the same synthetic
scene (tests/golden/make_golden.py datasets).  Images, origins, scalars: exact.  Directions / viewdirs: 1e-6 (the
3-term dot products run in another order than numpy's matmul).  Radii: 1e-4, the reference's finite difference of
fp32 directions carries ~3e-5 of cancellation noise (same bar as the on-device ray generator)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import golden

import mipnerf_pl_b200 as mp

FIELDS = mp.Rays_keys


@pytest.fixture(scope="module")
def scene_dir(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("scene"))
    mp.write_synthetic_blender_scene(root, 3, 16, 16, seed=5)
    mp.convert_blender_to_multiscale(root, root + "_ms", 3)
    return root


def check_rays(got, g, prefix):
    for k in FIELDS:
        a, b = np.asarray(getattr(got, k), dtype=np.float64), g[f"{prefix}_{k}"].astype(np.float64)
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if k in ("origins", "lossmult", "near", "far"):
            np.testing.assert_array_equal(a, b, err_msg=k)
        elif k == "radii":
            np.testing.assert_allclose(a, b, rtol=1e-4, err_msg=k)
        else:
            np.testing.assert_allclose(a, b, rtol=0, atol=2e-6, err_msg=k)


@pytest.mark.parametrize("white", [True, False])
def test_blender_train_matches_reference_loader(scene_dir, white):
    g = golden("datasets.npz")
    tag = "w" if white else "k"
    ds = mp.Blender(scene_dir, "train", white_bkgd=white)
    assert len(ds) == 3 * 16 * 16
    check_rays(ds.rays, g, f"blender_{tag}_train")
    np.testing.assert_array_equal(ds.images, g[f"blender_{tag}_train_images"])
    rays, rgb = ds[137]
    assert rays.origins.shape == (3,) and rgb.shape == (3,)
    assert ds.focal == pytest.approx(float(g["blender_focal"]), rel=1e-12)


def test_blender_val_cycles_like_reference(scene_dir):
    g = golden("datasets.npz")
    ds = mp.Blender(scene_dir, "val", batch_type="single_image")
    ds[0]
    rays, img = ds[0]                                   # the index is ignored: second call -> second image
    check_rays(rays, g, "blender_val1")
    np.testing.assert_array_equal(img, g["blender_val1_image"])
    with pytest.raises(AssertionError):
        mp.Blender(scene_dir, "val", batch_type="all_images")
    with pytest.raises(ValueError):
        mp.Blender(scene_dir, "train", factor=3)


def test_converter_metadata_and_multicam_match_reference(scene_dir):
    g = golden("datasets.npz")
    meta = json.load(open(os.path.join(scene_dir + "_ms", "metadata.json")))["train"]
    assert list(g["ms_meta_file_path"]) == meta["file_path"]
    for k in ("pix2cam", "cam2world", "width", "height", "focal", "lossmult", "near", "far", "label"):
        np.testing.assert_allclose(np.array(meta[k], dtype=np.float64), g[f"ms_meta_{k}"], rtol=1e-15, err_msg=k)
    ds = mp.Multicam(scene_dir + "_ms", "train")
    assert len(ds) == 3 * (256 + 64 + 16)
    check_rays(ds.rays, g, "ms_train")
    np.testing.assert_array_equal(ds.images, g["ms_train_images"])     # includes the box-filtered PNGs
    test = mp.Multicam(scene_dir + "_ms", "test", batch_type="single_image")
    rays, img = test[4]
    check_rays(rays, g, "ms_test4")
    np.testing.assert_array_equal(img, g["ms_test4_image"])
    assert mp.dataset_dict["multi_blender"] is mp.Multicam and mp.dataset_dict["blender"] is mp.Blender


def test_device_bank_refuses_cpu(scene_dir):
    with pytest.raises(RuntimeError):
        mp.DeviceRayBank(mp.load_blender_scene(scene_dir, "train"), device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("multiscale", [False, True])
def test_device_ray_bank_matches_host_loader(scene_dir, multiscale):
    """Rays + colours produced on the device from pixel ids equal the host loader's arrays at the same flat index."""
    dev = "cuda:0"
    if multiscale:
        scene, host = mp.load_multicam_scene(scene_dir + "_ms", "train"), mp.Multicam(scene_dir + "_ms", "train")
    else:
        scene, host = mp.load_blender_scene(scene_dir, "train"), mp.Blender(scene_dir, "train")
    bank = mp.DeviceRayBank(scene, dev)
    assert bank.num_pixels == len(host)
    ids = torch.arange(bank.num_pixels, device=dev)
    rays, rgb = bank.rays(ids)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(rgb.cpu().numpy(), host.images)
    for k in FIELDS:
        a, b = getattr(rays, k).cpu().numpy().astype(np.float64), np.asarray(getattr(host.rays, k), dtype=np.float64)
        if k in ("origins", "lossmult", "near", "far"):
            np.testing.assert_array_equal(a, b, err_msg=k)
        elif k == "radii":
            np.testing.assert_allclose(a, b, rtol=1e-4, err_msg=k)
        else:
            np.testing.assert_allclose(a, b, rtol=0, atol=2e-6, err_msg=k)
    # a random training batch straight into the training step
    gen = torch.Generator(device=dev).manual_seed(0)
    batch, target = bank.sample(64, generator=gen)
    assert batch.origins.shape == (64, 3) and target.shape == (64, 3) and batch.lossmult.min() >= 1
    model = mp.MipNerf().to(dev)
    out = mp.forward_backward(model, batch, target, True, True)
    assert torch.isfinite(out["loss"])


def test_system_setup_and_dataloaders(scene_dir):
    """MipNeRFSystem.setup / train_dataloader / val_dataloader (models/nerf_system.py:56-93) on the synthetic scene:
    a training batch is (Rays of [B,*] tensors, rgbs [B,3]), a validation batch one whole image."""
    hp = mp.default_hparams(**{"dataset_name": "blender", "data_path": scene_dir, "train.batch_size": 32,
                               "train.num_work": 0})
    system = mp.MipNeRFSystem(hp)
    system.setup("fit")
    rays, rgbs = next(iter(system.train_dataloader()))
    assert isinstance(rays, mp.Rays) and rays.origins.shape == (32, 3) and rays.radii.shape == (32, 1)
    assert rgbs.shape == (32, 3) and rgbs.dtype == torch.float32
    vrays, vimg = system.val_dataset[0]
    assert vrays.directions.shape == (16, 16, 3) and vimg.shape == (16, 16, 3)

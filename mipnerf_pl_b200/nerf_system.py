"""`MipNeRFSystem`: the reference's LightningModule surface for the render path
(models/nerf_system.py:13-177) on top of the B200 `MipNerf`.

Kept: constructor from the flat dotted `hparams` dict, `forward`, `render_image`,
`validation_step`, the `mip_nerf.` state_dict prefix and the checkpoint layout
(`state_dict` + `hyper_parameters`).  pytorch-lightning is not part of this image,
so when it cannot be imported a minimal stand-in base class provides
`save_hyperparameters` / `hparams` / `log` / `load_from_checkpoint`.
`training_step` / `configure_optimizers` run on the library's fp32 backward and Adam kernels
(mipnerf_pl_b200/train.py); `setup` / dataloaders use mipnerf_pl_b200/datasets.py (row N4).
"""
from __future__ import annotations

from typing import Optional

import torch

from .mip_nerf import MipNerf
from .rays import Rays, rearrange_render_image

try:  # pragma: no cover - not installed in the build image
    from pytorch_lightning import LightningModule as _Base
    _HAVE_PL = True
except Exception:  # noqa: BLE001
    _HAVE_PL = False

    class _Base(torch.nn.Module):
        """The few LightningModule features the render path touches."""

        def __init__(self):
            super().__init__()
            self.hparams = {}
            self.logger = None
            self.global_step = 0
            self._logged = {}

        def save_hyperparameters(self, hparams):
            self.hparams = dict(hparams)

        def log(self, name, value, **_):
            self._logged[name] = value

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
            """PL-1.5 checkpoint layout: {'state_dict', 'hyper_parameters', ...}."""
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            model = cls(ckpt["hyper_parameters"], **kwargs)
            model.load_state_dict(ckpt["state_dict"])
            return model


def calc_psnr(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """utils/metrics.py:182-188."""
    mse = torch.mean((x - y) ** 2)
    return -10.0 * torch.log10(mse)


class MipNeRFSystem(_Base):
    def __init__(self, hparams, precision: Optional[str] = None):
        super().__init__()
        self.save_hyperparameters(hparams)
        self.train_randomized = hparams['train.randomized']
        self.val_randomized = hparams['val.randomized']
        self.white_bkgd = hparams['train.white_bkgd']
        self.val_chunk_size = hparams['val.chunk_size']
        self.batch_size = hparams['train.batch_size']
        self.mip_nerf = MipNerf(
            num_samples=hparams['nerf.num_samples'],
            num_levels=hparams['nerf.num_levels'],
            resample_padding=hparams['nerf.resample_padding'],
            stop_resample_grad=hparams['nerf.stop_resample_grad'],
            use_viewdirs=hparams['nerf.use_viewdirs'],
            disparity=hparams['nerf.disparity'],
            ray_shape=hparams['nerf.ray_shape'],
            min_deg_point=hparams['nerf.min_deg_point'],
            max_deg_point=hparams['nerf.max_deg_point'],
            deg_view=hparams['nerf.deg_view'],
            density_activation=hparams['nerf.density_activation'],
            density_noise=hparams['nerf.density_noise'],
            density_bias=hparams['nerf.density_bias'],
            rgb_activation=hparams['nerf.rgb_activation'],
            rgb_padding=hparams['nerf.rgb_padding'],
            disable_integration=hparams['nerf.disable_integration'],
            append_identity=hparams['nerf.append_identity'],
            mlp_net_depth=hparams['nerf.mlp.net_depth'],
            mlp_net_width=hparams['nerf.mlp.net_width'],
            mlp_net_depth_condition=hparams['nerf.mlp.net_depth_condition'],
            mlp_net_width_condition=hparams['nerf.mlp.net_width_condition'],
            mlp_skip_index=hparams['nerf.mlp.skip_index'],
            mlp_num_rgb_channels=hparams['nerf.mlp.num_rgb_channels'],
            mlp_num_density_channels=hparams['nerf.mlp.num_density_channels'],
            mlp_net_activation=hparams['nerf.mlp.net_activation'],
            precision=precision,
        )

    def forward(self, batch_rays: Rays, randomized: bool, white_bkgd: bool):
        return self.mip_nerf(batch_rays, randomized, white_bkgd)  # num_levels results

    def setup(self, stage=None):
        """models/nerf_system.py:56-68."""
        from .datasets import dataset_dict
        dataset = dataset_dict[self.hparams['dataset_name']]
        self.train_dataset = dataset(data_dir=self.hparams['data_path'], split='train',
                                     white_bkgd=self.hparams['train.white_bkgd'],
                                     batch_type=self.hparams['train.batch_type'])
        self.val_dataset = dataset(data_dir=self.hparams['data_path'], split='val',
                                   white_bkgd=self.hparams['val.white_bkgd'],
                                   batch_type=self.hparams['val.batch_type'])

    def train_dataloader(self):
        """models/nerf_system.py:78-83 (the reference's host path; `datasets.DeviceRayBank.sample` is the
        device-resident alternative that needs no loader)."""
        from torch.utils.data import DataLoader
        return DataLoader(self.train_dataset, shuffle=True, num_workers=self.hparams['train.num_work'],
                          batch_size=self.hparams['train.batch_size'], pin_memory=True)

    def val_dataloader(self):
        """models/nerf_system.py:85-93: one image (H*W rays) at a time."""
        from torch.utils.data import DataLoader
        return DataLoader(self.val_dataset, shuffle=False, num_workers=1, batch_size=1, pin_memory=True,
                          persistent_workers=True)

    def configure_optimizers(self):
        """models/nerf_system.py:70-76: Adam(lr_init) + MipLRDecay stepped every optimiser step."""
        from .train import FusedAdam, MipLRDecay
        optimizer = FusedAdam(self.mip_nerf.parameters(), lr=self.hparams['optimizer.lr_init'])
        scheduler = MipLRDecay(optimizer, self.hparams['optimizer.lr_init'], self.hparams['optimizer.lr_final'],
                               self.hparams['optimizer.max_steps'], self.hparams['optimizer.lr_delay_steps'],
                               self.hparams['optimizer.lr_delay_mult'])
        return [optimizer], [{'scheduler': scheduler, 'interval': 'step'}]

    def training_step(self, batch, batch_nb, *, t_rand=None, u_jitter=None):
        """models/nerf_system.py:95-121.  The returned loss carries a grad_fn: forward and backward both ran
        in the library (mipnerf_b200_forward_backward); `loss.backward()` hands the gradients to autograd."""
        from .train import fused_loss
        rays, rgbs = batch
        loss, info = fused_loss(self.mip_nerf, rays, rgbs, self.train_randomized, self.white_bkgd,
                                coarse_loss_mult=self.hparams['loss.coarse_loss_mult'],
                                disable_multiscale_loss=self.hparams['loss.disable_multiscale_loss'],
                                t_rand=t_rand, u_jitter=u_jitter)
        with torch.no_grad():
            psnr_fine = calc_psnr(info["ret"][-1][0], rgbs[..., :3])
        self.log('train/loss', loss.detach())
        self.log('train/psnr', psnr_fine, prog_bar=True)
        return loss

    def validation_step(self, batch, batch_nb):
        """models/nerf_system.py:123-142 (image logging only when a logger is attached)."""
        _, rgbs = batch
        rgb_gt = rgbs[..., :3]
        coarse_rgb, fine_rgb, val_mask = self.render_image(batch)
        val_mse_coarse = (val_mask * (coarse_rgb - rgb_gt) ** 2).sum() / val_mask.sum()
        val_mse_fine = (val_mask * (fine_rgb - rgb_gt) ** 2).sum() / val_mask.sum()
        val_loss = self.hparams['loss.coarse_loss_mult'] * val_mse_coarse + val_mse_fine
        return {'val/loss': val_loss, 'val/psnr': calc_psnr(fine_rgb, rgb_gt)}

    def render_image(self, batch, return_distance: bool = False):
        """models/nerf_system.py:151-177: batch = (Rays with [1,H,W,C] fields, rgbs [1,H,W,3]) ->
        (coarse_rgb [1,H,W,3], fine_rgb [1,H,W,3], val_mask).  The reference's depth visualisation
        goes to the TensorBoard logger; here the raw distance map is returned on request."""
        rays, rgbs = batch
        _, height, width, _ = rgbs.shape
        single_image_rays, val_mask = rearrange_render_image(rays, self.val_chunk_size)
        coarse_rgb, fine_rgb, distances = [], [], []
        with torch.no_grad():
            for batch_rays in single_image_rays:
                ret = self(batch_rays, self.val_randomized, self.white_bkgd)
                (c_rgb, _, _, _, _), (f_rgb, distance, _, _, _) = ret[0], ret[-1]
                coarse_rgb.append(c_rgb)
                fine_rgb.append(f_rgb)
                distances.append(distance)
        coarse_rgb = torch.cat(coarse_rgb, dim=0).reshape(1, height, width, -1)
        fine_rgb = torch.cat(fine_rgb, dim=0).reshape(1, height, width, -1)
        distances = torch.cat(distances, dim=0).reshape(1, height, width)
        if return_distance:
            return coarse_rgb, fine_rgb, val_mask, distances
        return coarse_rgb, fine_rgb, val_mask


def default_hparams(**over) -> dict:
    """configs/lego.yaml as the flat dotted dict configs/config.py produces."""
    hp = {
        'seed': 4, 'num_gpus': 1, 'exp_name': 'lego',
        'train.batch_size': 3072, 'train.batch_type': 'all_images', 'train.num_work': 4,
        'train.randomized': True, 'train.white_bkgd': True,
        'val.batch_size': 1, 'val.batch_type': 'single_image', 'val.num_work': 4, 'val.randomized': False,
        'val.white_bkgd': True, 'val.check_interval': 10000, 'val.chunk_size': 8192, 'val.sample_num': 4,
        'nerf.num_samples': 128, 'nerf.num_levels': 2, 'nerf.resample_padding': 0.01,
        'nerf.stop_resample_grad': True, 'nerf.use_viewdirs': True, 'nerf.disparity': False,
        'nerf.ray_shape': 'cone', 'nerf.min_deg_point': 0, 'nerf.max_deg_point': 16, 'nerf.deg_view': 4,
        'nerf.density_activation': 'softplus', 'nerf.density_noise': 0., 'nerf.density_bias': -1.,
        'nerf.rgb_activation': 'sigmoid', 'nerf.rgb_padding': 0.001, 'nerf.disable_integration': False,
        'nerf.append_identity': 'Ture',  # sic: configs/lego.yaml:36 (a truthy string)
        'nerf.mlp.net_depth': 8, 'nerf.mlp.net_width': 256, 'nerf.mlp.net_depth_condition': 1,
        'nerf.mlp.net_width_condition': 128, 'nerf.mlp.net_activation': 'relu', 'nerf.mlp.skip_index': 4,
        'nerf.mlp.num_rgb_channels': 3, 'nerf.mlp.num_density_channels': 1,
        'optimizer.lr_init': 5e-4, 'optimizer.lr_final': 5e-6, 'optimizer.lr_delay_steps': 2500,
        'optimizer.lr_delay_mult': 0.01, 'optimizer.max_steps': 1000000,
        'loss.disable_multiscale_loss': False, 'loss.coarse_loss_mult': 0.1,
        'checkpoint.resume_path': None,
    }
    hp.update(over)
    return hp

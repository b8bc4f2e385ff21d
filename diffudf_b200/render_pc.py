"""NDF-style oriented point-cloud extraction: drop-in for /root/reference/src/render_pc.py:11-73.
Seeds are drawn on the host with numpy's global RNG exactly like the reference (same seed ->
same seeds); the projection steps, the acceptance test and the eigen-normals stay on the device."""
import warnings

import numpy as np
import torch

from .model import SIREN


class Sampler:
    def __init__(self, n_in_features=3, hidden_layers=[256, 256, 256, 256], w0=30, ww=None, checkpoint=None, device=0,
                 decoder=None):
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if decoder is None:
            decoder = SIREN(n_in_features=n_in_features, n_out_features=1, hidden_layer_config=hidden_layers, w0=w0, ww=ww)
            decoder.to(self.device)
            decoder.load_state_dict(torch.load(checkpoint, map_location=self.device))
        self.decoder = decoder
        self.features = n_in_features
        self.decoder.eval()

    def project(self, samples, gt_mode, alpha, num_steps):
        """Inner loop of generate_point_cloud (:43-53) on a float64 CUDA tensor (P,3).
        Returns (samples, last steps (P,), last gradients (P,3), Hessians of the last step (P,3,3))."""
        eng = self.decoder._engine_synced()
        samples = samples.to(torch.float64).contiguous().clone()
        if num_steps <= 0:
            return samples, None, None, None
        steps, g, H = eng.project_points(samples, num_steps, gt_mode, alpha, self.decoder.precision)
        return samples, steps, g, H

    def generate_point_cloud(self, gt_mode, alpha, num_steps=5, num_points=20000, surf_thresh=0.01, max_iter=1000):
        for p in self.decoder.parameters():
            p.requires_grad = False
        eng = self.decoder._engine_synced()
        surface_points = np.zeros((0, 3))
        normals = np.zeros((0, 3))
        for _ in range(max_iter):
            if len(surface_points) != 0:
                pick = np.random.uniform(0, len(surface_points), num_points // 2).astype(np.uint32)
                samples = surface_points[pick] + np.random.normal(0, 0.1, (num_points // 2, 3))
                samples = np.concatenate([samples, np.random.uniform(-1, 1, (num_points // 2, 3))])
            else:
                samples = np.random.uniform(-1, 1, (num_points, 3))
            s_d = torch.from_numpy(samples).to(self.device)
            s_d, steps, g, H = self.project(s_d, gt_mode, alpha, num_steps)
            on_domain = ((s_d >= -1) & (s_d <= 1)).all(dim=1)
            on_surf = (steps < surf_thresh) & on_domain
            if bool(on_surf.any()):
                surface_points = np.vstack((surface_points, s_d[on_surf].cpu().numpy()))
                if gt_mode == "siren":
                    gsel = g[on_surf]
                    nsel = gsel / torch.linalg.norm(gsel, dim=1, keepdim=True)
                else:
                    nsel, _, _ = eng.eig_normals(H[on_surf].contiguous())
                normals = np.vstack((normals, nsel.to(torch.float64).cpu().numpy()))
            if len(surface_points) >= num_points:
                break
        if len(surface_points) < num_points:
            warnings.warn(f"Max iterations reached. Only sampled {len(surface_points)} surface points.", RuntimeWarning)
        return surface_points, normals

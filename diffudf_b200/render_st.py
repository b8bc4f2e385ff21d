"""Sphere tracing field queries: drop-in for the query side of /root/reference/src/render_st.py
(evaluate :13-36, compute_curvature :42-55, compute_normals_and_cd :57-62, compute_grad :64-65,
propagate_rays :136-161, grad_descent :163-172) and of its shading side (create_projectional_image :67-134,
phong_shading :174-205, ward_reflectance :207-245).  The march keeps rays on the device (one active-count
read-back per iteration instead of a full D2H of the values); hit attributes and shading run on the device too
(dudf_shade_hits, float64 like the reference's numpy arrays)."""
import numpy as np
import torch

from .diff_operators import gradient, hessian, jacobian
from .inverses import inverse_torch


def grid_points(N, idx, device):
    """Coordinates of flat grid indices as render_mc.py:36-49 computes them (fp32)."""
    idx = idx.to(device=device, dtype=torch.int64)
    vs = torch.tensor(2.0 / (N - 1), dtype=torch.float32, device=device)
    i2 = (idx % N).to(torch.float32)
    i1 = ((idx // N) % N).to(torch.float32)
    i0 = ((idx // N // N) % N).to(torch.float32)
    return torch.stack([i0 * vs - 1, i1 * vs - 1, i2 * vs - 1], dim=1).contiguous()


def evaluate(model, samples, max_batch=64 ** 2, device=torch.device(0)):
    """Returns ([model_in], [model_out]) like the reference's chunk lists (one chunk: the kernels tile internally)."""
    if not torch.is_tensor(samples):
        samples = torch.from_numpy(np.asarray(samples)).float()
    x, y = model(samples.to(device).float().unsqueeze(0)).values()
    return [x], [y]


def batched_op(inputs, outputs, op, *args, **kwargs):
    return [op(x, y, *args, **kwargs) for x, y in zip(inputs, outputs)]


def compute_grad(inputs, outputs):
    return gradient(outputs, inputs)


def compute_normals_and_cd(inputs, outputs):
    """Eigen-normal (1,P,3) and principal directions (1,P,3,2) (on CPU like the reference)."""
    H = hessian(outputs, inputs)
    eng = inputs._dudf.model._engine
    n, dirs, _ = eng.eig_normals(H.reshape(-1, 3, 3).contiguous(), want_dirs=True)
    return n.unsqueeze(0), dirs.unsqueeze(0).detach().cpu()


def compute_curvature(inputs, normals, curvature="mean", device=torch.device(0)):
    """Mean: tr(dn/dx)/2, Gaussian: -det [[dn/dx, n],[n^T, 0]] — shapes (1,P,1) on CPU as in the reference."""
    shape_op, status = jacobian(normals, inputs)
    if curvature == "mean":
        return (torch.sum(torch.diagonal(shape_op[0], dim1=1, dim2=2), dim=-1) / 2).detach().cpu()[None, ..., None]
    if curvature == "gaussian":
        n = normals.reshape(-1, 3)
        ext = torch.zeros((shape_op.shape[1], 4, 4), device=shape_op.device)
        ext[:, :3, :3] = shape_op[0]
        ext[:, :3, 3] = n
        ext[:, 3, :3] = n
        return (-1 * torch.linalg.det(ext)[None, ..., None]).detach().cpu()
    return None


def _march(model, rays_d, t0_d, idx, gt_mode, alpha, thr, max_it):
    """The marching loop of propagate_rays (render_st.py:150-166) through dudf_march_rays: rays_d, t0_d (R,3) float64 CUDA
    (t0_d advanced in place), idx the rays to march.  Returns (hit mask (R,) bool, indices still marching, value queries)."""
    eng = model._engine_synced()
    R = t0_d.shape[0]
    active = torch.zeros(R, dtype=torch.uint8, device=t0_d.device)
    active[idx] = 1
    hit = torch.zeros(R, dtype=torch.uint8, device=t0_d.device)
    nq = eng.march_rays(t0_d, rays_d.contiguous(), active, hit, gt_mode, alpha, thr, max_it, model.precision)
    return hit.bool(), torch.nonzero(active).reshape(-1), nq


def propagate_rays(model, rays, t0, mask_rays, network_config, rendering_config, device):
    """Same contract as the reference: t0 (R,3) float64 and mask_rays (R,) bool numpy arrays are updated
    in place, the hit mask is returned; raises ValueError when nothing is hit."""
    dev = torch.device(device)
    rays_d = torch.from_numpy(np.ascontiguousarray(rays, dtype=np.float64)).to(dev)
    t0_d = torch.from_numpy(np.ascontiguousarray(t0, dtype=np.float64)).to(dev)
    idx = torch.nonzero(torch.from_numpy(np.asarray(mask_rays, dtype=bool)).to(dev)).reshape(-1)
    hits, idx, _ = _march(model, rays_d, t0_d, idx, network_config["gt_mode"], network_config["alpha"],
                          rendering_config["surface_threshold"], rendering_config["max_iterations"])
    t0[...] = t0_d.cpu().numpy()
    still = np.zeros(mask_rays.shape, dtype=bool)
    still[idx.cpu().numpy()] = True
    mask_rays[...] = still
    hits_np = hits.cpu().numpy()
    if hits_np.sum() == 0:
        raise ValueError(f"Ray tracing did not converge in {rendering_config['max_iterations']} iterations to any point at "
                         f"distance {rendering_config['surface_threshold']} or lower from surface.")
    return hits_np


def grad_descent(model, t0, mask_rays, network_config, rendering_config, device):
    dev = torch.device(device)
    eng = model._engine_synced()
    sel = np.asarray(mask_rays, dtype=bool)
    pos = torch.from_numpy(np.ascontiguousarray(t0[sel], dtype=np.float64)).to(dev)
    for _ in range(rendering_config["gd_steps"]):
        f, g, _, _ = eng.query(pos.to(torch.float32).contiguous(), 1, model.precision)
        gn = g / torch.linalg.norm(g, dim=1, keepdim=True)
        steps = inverse_torch(network_config["gt_mode"], f.abs(), network_config["alpha"])
        pos = pos - (gn * steps[:, None]).to(torch.float64)
    t0[sel] = pos.cpu().numpy()


def hit_attributes(model, points, ray_dirs=None, curvature="mean"):
    """Device-side replacement of the per-hit block of create_projectional_image (render_st.py:92-108):
    eigen-normals, principal directions and (optionally) curvature at `points` (P,3), sign-fixed
    against the ray directions.  Returns dict of CUDA tensors."""
    eng = model._engine_synced()
    x = points.to(torch.float32).contiguous()
    out = {}
    if curvature == "mean" and model.precision == "tcx3":
        # tensor-core route: Hessian jet + eigen-solve + 10-channel directional third-order jet (engine.mean_curvature)
        n, dirs, out["mean"] = eng.mean_curvature(x)
    elif curvature in ("mean", "gaussian"):
        _, _, H, T = eng.query(x, 3, "fp32")
        n, mean, gauss, _ = eng.curvature(H, T)
        out["mean"], out["gauss"] = mean, gauss
        _, dirs, _ = eng.eig_normals(H, want_dirs=True)
    else:
        _, _, H, _ = eng.query(x, 2, model.precision)
        n, dirs, _ = eng.eig_normals(H, want_dirs=True)
    out["dirs"] = dirs
    if ray_dirs is not None:
        align = -torch.sign((n * ray_dirs.to(n.dtype)).sum(-1, keepdim=True))
        n = n * align
        if "mean" in out:
            out["mean"] = out["mean"] * align[:, 0]
    out["normals"] = n
    return out


# ---- shading (src/render_st.py:174-245) -------------------------------------------------------------------------------------
def _shade(method, light_position, camera_position, hits, samples, normals, shininess=0.0, alpha1=1.0, alpha2=1.0, pc1=None, pc2=None,
           color_map=None, device=None):
    import ctypes

    from . import _lib
    dev = torch.device(device) if device is not None else (samples.device if torch.is_tensor(samples) and samples.is_cuda else torch.device("cuda:0"))

    def f64(a):
        return None if a is None else torch.as_tensor(a).to(device=dev, dtype=torch.float64).contiguous()

    hits_t = torch.as_tensor(hits).to(dev).bool().reshape(-1)
    smp, nrm, p1, p2, cm = f64(samples), f64(normals), f64(pc1), f64(pc2), f64(color_map)
    rows = torch.nonzero(hits_t).reshape(-1).contiguous()
    H = int(rows.numel())
    if nrm.shape[0] != H:
        raise ValueError(f"shading: {H} hits but {nrm.shape[0]} normals")
    colors = torch.ones_like(smp)
    light = (ctypes.c_double * 3)(*[float(v) for v in np.asarray(light_position, dtype=np.float64).reshape(3)])
    cam = None if camera_position is None else (ctypes.c_double * 3)(*[float(v) for v in np.asarray(camera_position, dtype=np.float64).reshape(3)])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dudf_shade_hits(rows.data_ptr(), H, smp.data_ptr(), nrm.data_ptr(), _lib.ptr(p1), _lib.ptr(p2), _lib.ptr(cm), light, cam,
                                              method, float(shininess), float(alpha1), float(alpha2), colors.data_ptr(), _lib.current_stream()),
                   "dudf_shade_hits")
    return colors


def phong_shading(light_position, shininess, hits, samples, normals, color_map=None):
    """src/render_st.py:174-205 with the same arguments (numpy arrays or tensors); returns the (R,3) float64 colours as numpy."""
    return _shade(0, light_position, None, hits, samples, normals, shininess=shininess, color_map=color_map).cpu().numpy()


def ward_reflectance(light_position, camera_position, hits, samples, normals, alpha1, alpha2, pc1, pc2, color_map=None):
    """src/render_st.py:207-245 with the same arguments; returns the (R,3) float64 colours as numpy."""
    return _shade(1, light_position, camera_position, hits, samples, normals, alpha1=alpha1, alpha2=alpha2, pc1=pc1, pc2=pc2,
                  color_map=color_map).cpu().numpy()


# the 11 ColorBrewer RdYlBu anchors matplotlib's 'RdYlBu' map interpolates linearly (used only when matplotlib is not installed;
# matplotlib itself is not part of the reference tree: parity of this fallback is unpinned)
_RDYLBU = np.array([[165, 0, 38], [215, 48, 39], [244, 109, 67], [253, 174, 97], [254, 224, 144], [255, 255, 191], [224, 243, 248],
                    [171, 217, 233], [116, 173, 209], [69, 117, 180], [49, 54, 149]], dtype=np.float64) / 255.0


def _rdylbu(x):
    try:
        from matplotlib import cm
        return cm.get_cmap("RdYlBu")(x)[:, :3]
    except Exception:
        lut_x = np.linspace(0.0, 1.0, 256)
        lut = np.stack([np.interp(lut_x, np.linspace(0.0, 1.0, 11), _RDYLBU[:, c]) for c in range(3)], 1)
        idx = np.clip((np.asarray(x, dtype=np.float64) * 256).astype(np.int64), 0, 255)
        idx[np.asarray(x) == 1.0] = 255
        return lut[idx]


def _percentile(sorted_vals, q):
    """np.percentile(a, q) (method 'linear') on an ascending float64 CUDA vector, with numpy's own interpolation formula
    (a + (b - a) t below t = 0.5, b - (b - a)(1 - t) from there on), so the clip bounds equal the reference's to the last bit."""
    n = sorted_vals.numel()
    pos = (n - 1) * (float(q) / 100.0)
    i = int(np.floor(pos))
    t = pos - i
    a = sorted_vals[i]
    b = sorted_vals[min(i + 1, n - 1)]
    return a + (b - a) * t if t < 0.5 else b - (b - a) * (1.0 - t)


def _curvature_colours(curv, q_lo, q_hi):
    """The colour-map block of create_projectional_image (src/render_st.py:109-114) on the device: 5-95 percentile clip (one sort of
    the (H,) float64 curvature vector), shift / scale to [0, 1], RdYlBu lookup.  Returns (H, 3) float64 on the device.  matplotlib's
    own map is used when it is installed (one small host round trip); otherwise the ColorBrewer anchors interpolated on the device."""
    srt = torch.sort(curv).values
    lo, hi = _percentile(srt, q_lo), _percentile(srt, q_hi)
    c = torch.clamp(curv, min=lo, max=hi)
    c = c - c.min()
    c = c / c.max()
    try:
        from matplotlib import cm
        return torch.from_numpy(np.ascontiguousarray(cm.get_cmap("RdYlBu")(c.cpu().numpy())[:, :3])).to(curv.device)
    except Exception:
        lut_x = np.linspace(0.0, 1.0, 256)
        lut = torch.from_numpy(np.stack([np.interp(lut_x, np.linspace(0.0, 1.0, 11), _RDYLBU[:, k]) for k in range(3)], 1)).to(curv.device)
        idx = torch.clamp((c * 256).to(torch.int64), 0, 255)
        idx[c == 1.0] = 255
        return lut[idx]


def create_projectional_image(model, rays, t0, mask_rays, network_config, rendering_config, device):
    """src/render_st.py:67-134 with the same arguments and in-place effects on t0 / mask_rays; returns the (height, width, 3)
    float64 image.  Marching, the gradient-descent refinement, the hit attributes (eigen-normals, principal directions, curvature)
    and the shading all run on the device, the percentile clip of the curvature colour map (a global statistic over the hits:
    one device sort) and the colour-map lookup included (_curvature_colours)."""
    dev = torch.device(device)
    eng = model._engine_synced()
    # one upload of the ray set; marching, refinement, attributes and shading work on device-resident state, the reference's in-place
    # contract on t0 / mask_rays is honoured with one download each
    rays_d = torch.from_numpy(np.ascontiguousarray(rays, dtype=np.float64)).to(dev)
    t0_d = torch.from_numpy(np.ascontiguousarray(t0, dtype=np.float64)).to(dev)
    idx = torch.nonzero(torch.from_numpy(np.asarray(mask_rays, dtype=bool)).to(dev)).reshape(-1)
    hits_d, idx, _ = _march(model, rays_d, t0_d, idx, network_config["gt_mode"], network_config["alpha"],
                            rendering_config["surface_threshold"], rendering_config["max_iterations"])
    if int(hits_d.sum()) == 0:
        raise ValueError(f"Ray tracing did not converge in {rendering_config['max_iterations']} iterations to any point at "
                         f"distance {rendering_config['surface_threshold']} or lower from surface.")
    pts = t0_d[hits_d]
    for _ in range(rendering_config["gd_steps"]):          # grad_descent (:163-172) on the hit points
        f, g, _, _ = eng.query(pts.to(torch.float32).contiguous(), 1, model.precision)
        gn = g / torch.linalg.norm(g, dim=1, keepdim=True)
        steps = inverse_torch(network_config["gt_mode"], f.abs(), network_config["alpha"])
        pts = pts - (gn * steps[:, None]).to(torch.float64)
    if rendering_config["gd_steps"] > 0:
        t0_d[hits_d] = pts
    t0[...] = t0_d.cpu().numpy()
    still = torch.zeros(t0_d.shape[0], dtype=torch.bool, device=dev)
    still[idx] = True
    mask_rays[...] = still.cpu().numpy()
    shape = (rendering_config["height"], rendering_config["width"], 3)
    if network_config["gt_mode"] == "siren":
        _, g, _, _ = eng.query(pts.to(torch.float32).contiguous(), 1, model.precision)
        normals = (g / torch.linalg.norm(g, dim=1, keepdim=True)).to(torch.float64)
        return _shade(0, rendering_config["light_position"], None, hits_d, t0_d, normals, shininess=rendering_config["shininess"]
                      ).cpu().numpy().reshape(shape)
    kind = rendering_config["plot_curvatures"] if rendering_config["plot_curvatures"] in ("mean", "gaussian") else None
    att = hit_attributes(model, pts, rays_d[hits_d], curvature=kind)
    normals = att["normals"].to(torch.float64)
    colours = None
    if kind is not None:
        colours = _curvature_colours((att["mean"] if kind == "mean" else att["gauss"]).to(torch.float64),
                                     rendering_config["curv_low_bound"], rendering_config["curv_high_bound"])
    if rendering_config["reflection_method"] == "blinn-phong":
        return _shade(0, rendering_config["light_position"], None, hits_d, t0_d, normals, shininess=rendering_config["shininess"],
                      color_map=colours).cpu().numpy().reshape(shape)
    if rendering_config["reflection_method"] == "ward":
        dirs = att["dirs"].to(torch.float64)
        return _shade(1, rendering_config["light_position"], rendering_config["camera_position"], hits_d, t0_d, normals,
                      alpha1=rendering_config["alpha1"], alpha2=rendering_config["alpha2"], pc1=dirs[..., 0].contiguous(),
                      pc2=dirs[..., 1].contiguous(), color_map=colours).cpu().numpy().reshape(shape)
    raise KeyError(rendering_config["reflection_method"])

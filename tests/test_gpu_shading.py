"""dudf_shade_hits behind phong_shading / ward_reflectance against the outputs of the UNMODIFIED reference functions
(src/render_st.py:174-245; tests/golden/make_golden_shading.py -> shading.npz), and create_projectional_image end to end."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cmap", ["plain", "cmap"])
def test_phong_and_ward_match_the_reference(cmap, golden):
    from diffudf_b200.render_st import phong_shading, ward_reflectance
    G = golden("shading.npz")
    cm = None if cmap == "plain" else G["color_map"]
    for sh in (0, 8, 40):
        got = phong_shading(G["light"], sh, G["hits"], G["samples"], G["normals"], color_map=cm)
        want = G[f"phong_{cmap}_{sh}"]
        assert got.shape == want.shape and got.dtype == np.float64
        assert np.abs(got - want).max() <= 1e-12, (sh, np.abs(got - want).max())      # pow() differs from libm by ulps
        assert np.array_equal(got[~G["hits"]], np.ones((int((~G["hits"]).sum()), 3)))
    got = ward_reflectance(G["light"], G["camera"], G["hits"], G["samples"], G["normals"], alpha1=0.2, alpha2=0.5, pc1=G["pc1"], pc2=G["pc2"],
                           color_map=cm)
    assert np.abs(got - G[f"ward_{cmap}"]).max() <= 1e-12
    # every branch of the Ward term is exercised by the fixture: NaN weights (back-facing), clipped highlights, plain diffuse
    assert (G[f"ward_{cmap}"][G["hits"]] == 0.9).any() and (G[f"ward_{cmap}"][G["hits"]] < 0.9).any()


def test_shading_argument_errors(golden):
    from diffudf_b200.render_st import phong_shading
    G = golden("shading.npz")
    with pytest.raises(ValueError):
        phong_shading(G["light"], 8, G["hits"], G["samples"], G["normals"][:-1])
    none = np.zeros_like(G["hits"])
    out = phong_shading(G["light"], 8, none, G["samples"], G["normals"][:0])
    assert np.array_equal(out, np.ones_like(G["samples"]))


@pytest.mark.parametrize("method,curv", [("blinn-phong", "mean"), ("ward", "gaussian"), ("blinn-phong", "none")])
def test_create_projectional_image_runs_end_to_end(method, curv, weights):
    """generate_st.py:127 with a 96 x 96 frame of configs/st_mean_cfg.json's camera: the image equals shading the pieces by hand."""
    from diffudf_b200 import SIREN, render_st
    m = SIREN(3, 1, [256] * 8, w0=30, delay_init=True)
    m.load_state_dict({f"net.{i}.0.{k}": torch.from_numpy(v) for i, (W, b) in enumerate(weights["trained"]) for k, v in (("weight", W), ("bias", b))})
    m = m.cuda()
    R = 96
    cam = np.array([0.8939, 0.7, 2.86]) * 0.45
    u, v = np.meshgrid(np.linspace(-0.6, 0.6, R), np.linspace(-0.6, 0.6, R))
    d = np.stack([u.ravel(), v.ravel(), -np.ones(R * R)], 1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    fwd = -cam / np.linalg.norm(cam)
    right = np.cross(fwd, [0, 1.0, 0]); right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    rays = d[:, :1] * right + d[:, 1:2] * up - d[:, 2:3] * fwd
    t0 = np.tile(cam, (R * R, 1)) + rays * 0.35
    net_cfg = {"gt_mode": "tanh", "alpha": 100.0}
    cfg = {"surface_threshold": 0.004, "max_iterations": 100, "gd_steps": 2, "height": R, "width": R, "light_position": [1.0, 1.0, 2.5],
           "camera_position": cam.tolist(), "shininess": 16, "plot_curvatures": curv, "curv_low_bound": 5, "curv_high_bound": 95,
           "reflection_method": method, "alpha1": 0.2, "alpha2": 0.5}
    mask = np.ones(R * R, dtype=bool)
    img = render_st.create_projectional_image(m, rays, t0, mask, net_cfg, cfg, torch.device("cuda:0"))
    assert img.shape == (R, R, 3) and img.dtype == np.float64 and np.isfinite(img).all()
    flat = img.reshape(-1, 3)
    bg = (flat == 1.0).all(axis=1)
    assert 0.02 < 1.0 - bg.mean() < 0.9                      # some rays hit, some do not
    assert flat[~bg].min() >= 0.0 and flat[~bg].max() <= 0.9

"""CPU port of the reference's OWN way of running the hot path — TEST / BASELINE INFRASTRUCTURE ONLY.

The reference (LIA-DiTella/DiffUDF) is Python + torch and cannot travel to the GPU box, so the CPU
baseline that bench.py times there is this restatement: the same operation sequence the reference
executes — nn.Linear-style matmuls + torch.sin, `torch.autograd.grad(create_graph=True)` for the
gradient, three more double-backward calls for the Hessian rows, `torch.linalg.eigh`, the loss
terms, `backward()` and `torch.optim.Adam.step()` — written functionally in our own words.

Follows: src/model.py:116-135, src/diff_operators.py:187-212, src/loss_functions.py:9-53,106-155,
train.py:204-222, src/evaluate.py:5-37.  Pinned against the unmodified reference in
tests/test_oracle_golden.py::test_autograd_port_matches_reference_fixtures.
Only tests/ and bench.py's cpu_baseline / --impl reference legs may import it.
"""
import numpy as np
import torch


def make_params(params_np, dtype=torch.float32, requires_grad=True, device="cpu"):
    """device="cuda:0" runs the same operation sequence as stock eager PyTorch on the GPU (bench.py's "reference on the same
    B200" figure, SURVEY 8d last row); tests and the CPU baseline use the default."""
    out = []
    for W, b in params_np:
        out.append((torch.tensor(np.asarray(W), dtype=dtype, requires_grad=requires_grad, device=device),
                    torch.tensor(np.asarray(b), dtype=dtype, requires_grad=requires_grad, device=device)))
    return out


def field(params, x, w0=30.0):
    """x (1,P,3) -> (coords leaf, f (1,P,1))."""
    coords = x.clone().detach().requires_grad_(True)
    h = coords
    for W, b in params[:-1]:
        h = torch.sin(w0 * (h @ W.t() + b))
    W, b = params[-1]
    return coords, h @ W.t() + b


def grad_of(y, x):
    return torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), create_graph=True)[0]


def hessian_of(y, x):
    g = grad_of(y, x).squeeze(0)
    rows = [grad_of(g[:, i], x).squeeze(0)[:, None, :] for i in range(3)]
    return torch.cat(rows, dim=1).unsqueeze(0)


def loss_s1(params, x, normals, d, w, alpha):
    coords, f = field(params, x)
    t = torch.tanh(alpha * d)
    tdf = d * t
    zero = torch.zeros_like(f)
    out = {
        "sdf_on_surf": torch.where(d == 0, f.abs(), zero).mean() * w[0],
        "sdf_off_surf": torch.where(d != 0, (tdf - f).abs(), zero).mean() * w[1],
    }
    if w[2] != 0:
        H = hessian_of(f.squeeze(-1), coords)
        _, V = torch.linalg.eigh(H)
        v = V[..., 2]
        cosv = torch.nn.functional.cosine_similarity(normals, v, dim=-1)
        on = (d == 0).flatten()
        out["hessian_constraint"] = torch.where(on, 1 - cosv.abs(), torch.zeros_like(cosv)).mean() * w[2]
    else:
        out["hessian_constraint"] = torch.zeros(1, device=f.device)[0]
    if w[3] != 0:
        g = grad_of(f, coords)
        tgt = (t + d * alpha * (1 - t ** 2)).abs().squeeze(-1)
        out["grad_constraint"] = (torch.linalg.norm(g.squeeze(0), dim=-1) - tgt).abs().mean() * w[3]
    else:
        out["grad_constraint"] = torch.zeros(1, device=f.device)[0]
    return out


def loss_s2(params, x, normals, d, w, alpha=None):
    _, f = field(params, x)
    s = f[d == 0]
    return {"sdf_on_surf": s.mean().abs() * w[0], "std_on_surf": torch.std(s) * w[1]}


def train_step(params, opt, x, normals, d, mode, w, alpha):
    """optim.zero_grad(); loss; backward; optim.step() — returns the loss terms as floats (one host read-back each, as the
    reference's .item() calls, train.py:227-233)."""
    opt.zero_grad()
    loss = loss_s1(params, x, normals, d, w, alpha) if mode == "s1" else loss_s2(params, x, normals, d, w, alpha)
    total = 0
    for v in loss.values():
        total = total + v
    total.backward()
    opt.step()
    return {k: float(v) for k, v in loss.items()}


def make_optimizer(params, lr):
    flat = []
    for W, b in params:
        flat += [W, b]
    return torch.optim.Adam(flat, lr=lr)


def evaluate(params, samples, want_grad=True, want_hess=False, max_batch=64 ** 2):
    """Chunked f / grad / Hessian query (src/evaluate.py) returning float64 numpy arrays."""
    n = samples.shape[0]
    dev = params[0][0].device
    f_out = np.zeros((n, 1))
    g_out = np.zeros((n, 3)) if want_grad else None
    h_out = np.zeros((n, 3, 3)) if want_hess else None
    head = 0
    while head < n:
        xs = torch.from_numpy(samples[head:head + max_batch]).float().unsqueeze(0).to(dev)
        coords, f = field(params, xs)
        if want_grad:
            g_out[head:head + max_batch] = grad_of(f, coords).squeeze(0).detach().cpu().numpy()
        if want_hess:
            h_out[head:head + max_batch] = hessian_of(f.squeeze(-1), coords)[0].detach().cpu().numpy()
        f_out[head:head + max_batch] = f.squeeze(0).detach().cpu().numpy()
        head += max_batch
    return f_out, g_out, h_out

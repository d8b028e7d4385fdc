"""Generate tests/golden/mlp.npz by running the UNMODIFIED reference MLP (models/mlp.py) -- TEST INFRASTRUCTURE.

Run in the build container (where /root/reference is mounted):   python -m oracle.make_golden_mlp
The reference module is loaded from its file; the only thing touched is the ``p`` attribute of its Dropout
modules on the constructed object (0.0 for the deterministic training-mode case: torch's dropout generator
cannot be reproduced outside torch).  Stored: the seeded state_dict, an input batch, the eval-mode output, and a
training-mode (p = 0) forward + backward: output, updated running statistics, parameter and input gradients.
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("GSR_REFERENCE_ROOT", "/root/reference")
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASE = dict(input_size=190, output_size=144, hidden_dim=52, n_layers=3, batch=9, seed=21)


def load_reference_mlp():
    spec = importlib.util.spec_from_file_location("ref_mlp", os.path.join(REFERENCE_ROOT, "models", "mlp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.SuperResMLP


def main():
    c = CASE
    SuperResMLP = load_reference_mlp()
    torch.manual_seed(c["seed"])
    model = SuperResMLP(c["input_size"], c["output_size"], c["hidden_dim"], c["n_layers"])
    g = torch.Generator().manual_seed(c["seed"] + 1)
    with torch.no_grad():                      # non-trivial affine parameters and running statistics
        for m in model.layers:
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(1 + 0.3 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
    out = {f"state.{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
    x = torch.randn(c["batch"], c["input_size"], generator=g)
    out["x"] = x.numpy().copy()
    model.eval()
    with torch.no_grad():
        out["y_eval"] = model(x).numpy().copy()
    model.train()
    for m in model.layers:
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    xr = x.clone().requires_grad_(True)
    y = model(xr)
    w = torch.randn(y.shape, generator=g)
    (y * w).sum().backward()
    out["grad_out"] = w.numpy().copy()
    out["y_train_p0"] = y.detach().numpy().copy()
    out["grad_x"] = xr.grad.numpy().copy()
    for k, p in model.named_parameters():
        out[f"grad.{k}"] = p.grad.numpy().copy()
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            out[f"after.{k}"] = v.detach().numpy().copy()
    for k, v in c.items():
        out[f"case.{k}"] = np.int64(v)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "mlp.npz"), **out)
    print("wrote", os.path.join(GOLDEN_DIR, "mlp.npz"), {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()

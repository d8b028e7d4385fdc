"""Generate tests/golden/mlp_notebook.npz by running the UNMODIFIED notebook MLP -- TEST INFRASTRUCTURE.

Run in the build container (where /root/reference is mounted):   python -m oracle.make_golden_mlp_notebook
The class is ``SuperResMLP`` of training_template.ipynb (cell 5, lines 100-216 of the notebook JSON: the 12 720 -> 35 778
model of BASELINE configs[3]).  Its source is executed as is; torch_geometric (absent here, and outside the path) is
stubbed: ``to_dense_adj`` returns the dense batch the fixture passes in through a fake ``Batch``.  The only thing
touched on the constructed object is the ``p`` of its Dropout modules for the deterministic training-mode case.
Stored: the seeded state_dict (Xavier-on-the-derived-weight quirk included), an input batch, the eval-mode output, a
training-mode (p = 0) forward + backward (output, parameter gradients, updated BatchNorm statistics and
spectral-norm ``weight_u`` / ``weight_v`` buffers).
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("GSR_REFERENCE_ROOT", "/root/reference")
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASE = dict(nodes_in=12, nodes_out=16, hidden=40, n_layers=2, batch=7, seed=33)


def load_notebook_mlp():
    nb = json.load(open(os.path.join(REFERENCE_ROOT, "training_template.ipynb")))
    src = next("".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code" and "class SuperResMLP" in "".join(c["source"]))
    tg, tgu, tgd = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.utils"), types.ModuleType("torch_geometric.data")
    tgu.to_dense_adj = lambda edge_index, edge_attr=None, batch=None: edge_index        # the fake Batch carries the dense matrices
    tgd.Batch = object
    saved = {k: sys.modules.get(k) for k in ("torch_geometric", "torch_geometric.utils", "torch_geometric.data")}
    sys.modules.update({"torch_geometric": tg, "torch_geometric.utils": tgu, "torch_geometric.data": tgd})
    try:
        ns = {}
        exec(compile(src, "training_template.ipynb:cell5", "exec"), ns)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ns["SuperResMLP"]


class FakeBatch:
    def __init__(self, dense):
        self.edge_index, self.edge_attr, self.batch = dense, None, None


def main():
    c = CASE
    SuperResMLP = load_notebook_mlp()
    torch.manual_seed(c["seed"])
    model = SuperResMLP(c["nodes_in"], c["nodes_out"], c["hidden"], c["n_layers"])
    g = torch.Generator().manual_seed(c["seed"] + 1)
    with torch.no_grad():                      # non-trivial affine parameters and running statistics
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(1 + 0.3 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
            if isinstance(m, torch.nn.Linear):
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    out = {f"state.{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
    x = torch.rand(c["batch"], c["nodes_in"], c["nodes_in"], generator=g)
    x = (x + x.transpose(1, 2)) / 2
    out["x"] = x.numpy().copy()
    model.eval()
    with torch.no_grad():
        out["y_eval"] = model(FakeBatch(x)).numpy().copy()
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    y = model(FakeBatch(x))
    w = torch.randn(y.shape, generator=g)
    (y * w).sum().backward()
    out["grad_out"] = w.numpy().copy()
    out["y_train_p0"] = y.detach().numpy().copy()
    for k, p in model.named_parameters():
        out[f"grad.{k}"] = p.grad.numpy().copy()
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k or k.endswith("weight_u") or k.endswith("weight_v"):
            out[f"after.{k}"] = v.detach().numpy().copy()
    for k, v in c.items():
        out[f"case.{k}"] = np.int64(v)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "mlp_notebook.npz"), **out)
    print("wrote mlp_notebook.npz", sorted(k for k in out if k.startswith("state."))[:8], "...")


if __name__ == "__main__":
    main()

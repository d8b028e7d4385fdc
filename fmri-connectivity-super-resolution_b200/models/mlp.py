"""MLP baseline -- drop-in for the reference's models/mlp.py:7-39 (SURVEY 8 f, row N2).

``SuperResMLP(input_size, output_size, hidden_dim, n_layers)`` keeps the reference's constructor, its ``layers``
ModuleList (same module types at the same indices, so the ``state_dict`` keys and the seeded initialisation are
the reference's) and its ``device`` property.  The Linear layers run on the library's GEMM kernels
(``gsr_tcgemm_f32`` where the shape allows the tensor-core path, bias in the epilogue), every
BatchNorm1d -> Dropout -> ReLU run is one fused pass (``gsr_bn_stats_f32`` + ``gsr_bn_act_fwd_f32``).  As in the
reference, ``output_size`` must be a perfect square (``nn.Unflatten`` to (h, h), models/mlp.py:9,27).  There is
no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import functional as F_


class SuperResMLP(nn.Module):
    def __init__(self, input_size, output_size, hidden_dim, n_layers):
        super().__init__()
        h = int(np.sqrt(output_size))                                                       # models/mlp.py:9
        layers = [nn.Flatten(start_dim=1), nn.Linear(input_size, hidden_dim), nn.BatchNorm1d(hidden_dim),
                  nn.Dropout(p=0.1), nn.ReLU()]                                             # :10-18
        for _ in range(n_layers - 1):                                                       # :19-24
            layers += [nn.Linear(hidden_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.Dropout(p=0.1), nn.ReLU()]
        layers += [nn.Linear(hidden_dim, output_size), nn.Unflatten(dim=1, unflattened_size=(h, h))]   # :26-27
        self.layers = nn.ModuleList(layers)
        self.dropout_seed = None          # tests: fixed seed of the counter-based dropout mask (None: torch's generator)

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, samples: torch.Tensor):
        if self.device.type != "cuda":
            raise RuntimeError("SuperResMLP: move the model to the GPU first (model.to('cuda')); there is no CPU path")
        x = samples.to(self.device)                                                         # :36
        layers, i = self.layers, 0
        while i < len(layers):
            layer = layers[i]
            if isinstance(layer, nn.Linear):
                x = F_.LinearFn.apply(x.contiguous(), layer.weight, layer.bias)
                i += 1
            elif isinstance(layer, nn.BatchNorm1d):
                # the run BatchNorm1d [-> Dropout] [-> ReLU] is one pass over x
                p, relu, j = 0.0, False, i + 1
                if j < len(layers) and isinstance(layers[j], nn.Dropout):
                    p, j = layers[j].p, j + 1
                if j < len(layers) and isinstance(layers[j], nn.ReLU):
                    relu, j = True, j + 1
                training = self.training or layer.running_mean is None
                if training and layer.num_batches_tracked is not None:
                    layer.num_batches_tracked += 1
                x = F_.batch_norm_act(x, layer.weight, layer.bias, layer.running_mean, layer.running_var, training,
                                      momentum=layer.momentum, eps=layer.eps, p_drop=p if self.training else 0.0,
                                      relu=relu, seed=self.dropout_seed)
                i = j
            elif isinstance(layer, nn.Dropout):
                if self.training and layer.p > 0:
                    raise NotImplementedError("SuperResMLP: a Dropout that does not follow a BatchNorm1d")
                i += 1
            elif isinstance(layer, nn.ReLU):
                x = torch.clamp_min(x, 0)
                i += 1
            else:                                                                           # Flatten / Unflatten: views
                x = layer(x)
                i += 1
        return x

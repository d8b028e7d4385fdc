"""The notebook MLP -- ``SuperResMLP`` of the reference's training_template.ipynb:100-216, the model BASELINE
configs[3] sizes ("vectorised upper triangles, 12720 -> 35778 features"): 160 (160 - 1) / 2 = 12 720 inputs,
268 (268 - 1) / 2 = 35 778 outputs.  (``models/mlp.py`` cannot be built at that size: its ``nn.Unflatten`` needs a square
output)

Same constructor (``num_nodes_input, num_nodes_output, num_hidden_nodes, n_layers, dropout``), same sub-module tree
(``input_layer`` / ``residual_blocks`` / ``output_layer`` Sequentials of ``spectral_norm(nn.Linear)``, ``BatchNorm1d``,
``Dropout``, ``LeakyReLU``, ``Sigmoid``; ``input_mask`` / ``output_mask`` buffers), hence the same ``state_dict`` keys
(``...weight_orig``, ``weight_u``, ``weight_v``) and the same seeded initialisation (Xavier-uniform weights, zero
biases).  The arithmetic runs on the library's kernels:
  * ``x[:, input_mask]``                      -> ``gsr_triu_flatten``
  * every ``spectral_norm(nn.Linear)``        -> ``gsr_tcgemm_f32`` / ``gsr_bgemm_f32`` with the bias epilogue on the
    spectrally normalised weight W / sigma.  The power iteration on the ``weight_u`` / ``weight_v`` buffers and
    sigma = u . (W v) are torch's own three mat-vecs, statement for statement as in its hook (parameter plumbing, as
    the optimiser is), so the buffers behave exactly as the reference's; the passes over the weight -- W / sigma and,
    backward, G / sigma - (sum(G o W_eff) / sigma) u v^T -- run on ``gsr_scale_div_f32`` / ``gsr_spectral_bwd_f32``
    (torch's hook spends ~12 broadcast elementwise launches per layer on them);
  * ``BatchNorm1d -> Dropout -> LeakyReLU``   -> ``gsr_bn_stats_f32`` + ``gsr_bn_act_fwd_f32`` (negative slope 0.01)
  * ``leaky_relu(block(x) + residual)``       -> ``gsr_add_act_fwd_f32``
  * ``nn.Sigmoid``                            -> ``gsr_sigmoid_fwd_f32``
  * ``matrix[:, output_mask] = x; + ^T``      -> ``gsr_anti_vectorize``.
The reference's ``forward`` takes a torch_geometric ``Batch`` and densifies it with ``to_dense_adj`` (:163);
torch_geometric is outside the path, so ``forward`` takes the dense ``[B, n, n]`` adjacency batch (or
the already vectorised ``[B, n (n - 1) / 2]`` triangles).  There is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as nnF
from torch.nn.utils.spectral_norm import SpectralNorm, spectral_norm

from .. import functional as F_


class SuperResMLP(nn.Module):
    def __init__(self, num_nodes_input: int, num_nodes_output: int, num_hidden_nodes: int, n_layers: int, dropout: float = 0.1):
        super().__init__()
        self.num_nodes_input = num_nodes_input
        self.num_nodes_output = num_nodes_output
        input_size = (num_nodes_input * (num_nodes_input - 1)) // 2                      # :113-114
        output_size = (num_nodes_output * (num_nodes_output - 1)) // 2
        hidden_size = num_hidden_nodes
        self.register_buffer("input_mask", torch.triu(torch.ones(num_nodes_input, num_nodes_input), diagonal=1).bool())
        self.register_buffer("output_mask", torch.triu(torch.ones(num_nodes_output, num_nodes_output), diagonal=1).bool())
        self.input_layer = nn.Sequential(                                                  # :123-129
            nn.Flatten(start_dim=1), spectral_norm(nn.Linear(input_size, hidden_size)), nn.BatchNorm1d(hidden_size),
            nn.Dropout(dropout), nn.LeakyReLU(negative_slope=0.01))
        self.residual_blocks = nn.ModuleList([                                             # :132-139
            nn.Sequential(spectral_norm(nn.Linear(hidden_size, hidden_size)), nn.BatchNorm1d(hidden_size),
                          nn.Dropout(dropout), nn.LeakyReLU(negative_slope=0.01)) for _ in range(n_layers)])
        self.output_layer = nn.Sequential(spectral_norm(nn.Linear(hidden_size, output_size)), nn.Sigmoid())   # :142-145
        self._init_weights()
        self.dropout_seed = None          # tests: fixed seed of the counter-based dropout mask (None: torch's generator)

    def _init_weights(self):
        # :150-155, statement for statement: under spectral_norm ``m.weight`` is the DERIVED tensor (weight_orig / sigma,
        # recomputed by the hook at every forward), not the parameter -- so the reference's Xavier call consumes the
        # generator and is then overwritten; ``weight_orig`` keeps nn.Linear's default initialisation.  Reproduced so
        # that a seeded construction gives the reference's parameters.
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    @property
    def device(self):
        return next(self.parameters()).device

    # ---- the pieces, on the library's kernels
    def _spectral_weight(self, lin: nn.Linear) -> torch.Tensor:
        """``lin.weight = weight_orig / sigma`` as torch's SpectralNorm forward pre-hook computes it (``compute_weight``:
        one power iteration on the ``weight_u`` / ``weight_v`` buffers per TRAINING forward, in place and without autograd,
        then sigma = u . (W v)), with the passes over the weight on the library's kernels (``SpectralScaleFn``)."""
        hook = next((h for h in lin._forward_pre_hooks.values() if isinstance(h, SpectralNorm)), None)
        if hook is None or hook.dim != 0:                   # not the layout this model builds: torch's own hook
            for h in lin._forward_pre_hooks.values():
                h(lin, None)
            return lin.weight
        weight = getattr(lin, hook.name + "_orig")
        u, v = getattr(lin, hook.name + "_u"), getattr(lin, hook.name + "_v")
        if lin.training:
            with torch.no_grad():
                for _ in range(hook.n_power_iterations):
                    v = nnF.normalize(torch.mv(weight.t(), u), dim=0, eps=hook.eps, out=v)
                    u = nnF.normalize(torch.mv(weight, v), dim=0, eps=hook.eps, out=u)
                if hook.n_power_iterations > 0:             # the buffers are updated in place; autograd sees copies
                    u, v = u.clone(memory_format=torch.contiguous_format), v.clone(memory_format=torch.contiguous_format)
        w = F_.SpectralScaleFn.apply(weight, u, v)
        setattr(lin, hook.name, w)                          # as the hook does: ``lin.weight`` is the derived tensor
        return w

    def _linear(self, lin: nn.Linear, x: torch.Tensor) -> torch.Tensor:
        return F_.LinearFn.apply(x.contiguous(), self._spectral_weight(lin), lin.bias)

    def _bn_drop_act(self, bn: nn.BatchNorm1d, drop: nn.Dropout, act: nn.LeakyReLU, x: torch.Tensor) -> torch.Tensor:
        training = self.training or bn.running_mean is None
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        return F_.batch_norm_act(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, momentum=bn.momentum,
                                 eps=bn.eps, p_drop=drop.p if self.training else 0.0, relu=float(act.negative_slope),
                                 seed=self.dropout_seed)

    def forward(self, samples: torch.Tensor) -> torch.Tensor:
        if self.device.type != "cuda":
            raise RuntimeError("SuperResMLP: move the model to the GPU first (model.to('cuda')); there is no CPU path")
        x = samples.to(self.device, torch.float32)
        if x.dim() == 3:                                                                  # dense [B, n, n]: x[:, input_mask] (:166)
            if x.shape[-1] != self.num_nodes_input or x.shape[-2] != self.num_nodes_input:
                raise RuntimeError(f"SuperResMLP: expected [B, {self.num_nodes_input}, {self.num_nodes_input}] adjacency matrices")
            x = F_.TriuFlattenFn.apply(x.contiguous())
        lin, bn, drop, act = self.input_layer[1], self.input_layer[2], self.input_layer[3], self.input_layer[4]
        x = self._bn_drop_act(bn, drop, act, self._linear(lin, x))                         # :169
        for block in self.residual_blocks:                                                 # :172-176
            y = self._bn_drop_act(block[1], block[2], block[3], self._linear(block[0], x))
            x = F_.AddActFn.apply(y, x, 0.01)
        x = F_.SigmoidFn.apply(self._linear(self.output_layer[0], x))                      # :179
        return F_.SymFromTriuFn.apply(x, self.num_nodes_output)                            # :182-187

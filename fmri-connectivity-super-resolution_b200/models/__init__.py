"""Mirror of the reference's ``models`` package for the GSR-Net path
(models/gsr_model.py, models/layers.py, models/ops.py) and the MLP baselines (models/mlp.py; the notebook model of training_template.ipynb:100-216)."""
from .gsr_model import GSRNet
from .layers import GSRLayer, GraphConvolution
from .ops import GCN, GraphPool, GraphUnet, GraphUnpool
from .mlp import SuperResMLP
from .mlp_notebook import SuperResMLP as SuperResMLPNotebook

__all__ = ["GSRNet", "GSRLayer", "GraphConvolution", "GraphUnet", "GraphPool", "GraphUnpool", "GCN", "SuperResMLP", "SuperResMLPNotebook"]

"""Mirror of the reference's ``models`` package for the GSR-Net path
(models/gsr_model.py, models/layers.py, models/ops.py) and the MLP baseline (models/mlp.py)."""
from .gsr_model import GSRNet
from .layers import GSRLayer, GraphConvolution
from .ops import GCN, GraphPool, GraphUnet, GraphUnpool
from .mlp import SuperResMLP

__all__ = ["GSRNet", "GSRLayer", "GraphConvolution", "GraphUnet", "GraphPool", "GraphUnpool", "GCN", "SuperResMLP"]

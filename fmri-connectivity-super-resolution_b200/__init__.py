"""B200-native GSR-Net hot path (LR 160 -> HR 268 brain-graph super-resolution).

Drop-in for the GSR-Net path of denisfouchard/fMRI-connectivity-super-resolution:

    pkg.models.gsr_model.GSRNet        pkg.models.layers.{GSRLayer, GraphConvolution}
    pkg.models.ops.{GraphUnet, GCN, GraphPool, GraphUnpool}   pkg.models.mlp.SuperResMLP
    pkg.utils.matrix_vectorizer.MatrixVectorizer
    pkg.utils.preprocessing.{normalize_adj_torch, pad_HR_adj, unpad}
    pkg.utils.gsr_train.{train, test}   (+ GSRTrainStep / train_batched, the batched extension)
    pkg.utils.evaluation.evaluation_reductions (MAE / PCC / JSD)

All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
include/gsr_sm100.h (libgsr_sm100.so, loaded with ctypes).  There is no CPU fallback:
``lib()`` raises when the library has not been built.
"""
from ._lib import GsrLibraryError, LIB_PATH, launch_count, lib, set_tuning  # noqa: F401
from . import kernels, functional  # noqa: F401
from . import models, utils  # noqa: F401
from .models import GSRNet, GSRLayer, GraphConvolution, GraphUnet, GraphPool, GraphUnpool, GCN, SuperResMLP, SuperResMLPNotebook  # noqa: F401
from .utils import MatrixVectorizer, normalize_adj_torch, pad_HR_adj, unpad  # noqa: F401
from .utils.optim import FlatAdam  # noqa: F401
from .utils.gsr_train import GSRTrainStep, GraphedTrainStep, train, test, train_batched, hr_target_basis  # noqa: F401
from .utils.evaluation import evaluation_reductions  # noqa: F401
from .subject_parallel import SubjectParallel, shard_bounds, fold_assignment, fold_groups  # noqa: F401

__all__ = ["lib", "GSRNet", "GSRLayer", "GraphConvolution", "GraphUnet", "GraphPool", "GraphUnpool", "GCN", "SuperResMLP",
           "SuperResMLPNotebook", "MatrixVectorizer", "normalize_adj_torch", "pad_HR_adj", "unpad", "train", "test", "train_batched",
           "GSRTrainStep", "GraphedTrainStep", "FlatAdam", "evaluation_reductions", "SubjectParallel"]

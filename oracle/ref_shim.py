"""Import the UNMODIFIED reference from /root/reference -- build container only.

TEST INFRASTRUCTURE.  Used by oracle/make_golden.py (fixture generation) and by
the optional ``reference_live`` tests; never on the GPU box (``/root/reference``
does not exist there) and never by product code.

Recipe: stub ``matplotlib`` (imported at
models/layers.py:5 but unused), put the reference root on ``sys.path``, import
``models.gsr_model`` and force its module-level ``device`` globals to CPU.
``utils.gsr_train`` cannot be imported (it calls ``data()`` at import,
utils/gsr_train.py:88), so its loop body is driven from make_golden.py.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("GSR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "gsr_model.py"))


def load():
    """Returns a namespace with GSRNet, GSRLayer, GraphConvolution, GraphUnet,
    GraphPool, GraphUnpool, GCN, MatrixVectorizer, normalize_adj_torch,
    pad_HR_adj, unpad -- the reference's own objects."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    # The reference uses top-level packages called ``models`` and ``utils``.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("models", "utils") or k.startswith(("models.", "utils."))}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import models.gsr_model as gm          # noqa
        import models.layers as ml             # noqa
        import models.ops as mo                # noqa
        import utils.preprocessing as up       # noqa
        import utils.matrix_vectorizer as mv   # noqa
        cpu = torch.device("cpu")
        for mod in (gm, ml, mo, up):
            mod.device = cpu
        ns = types.SimpleNamespace(
            GSRNet=gm.GSRNet, GSRLayer=ml.GSRLayer, GraphConvolution=ml.GraphConvolution,
            GraphUnet=mo.GraphUnet, GraphPool=mo.GraphPool, GraphUnpool=mo.GraphUnpool, GCN=mo.GCN,
            MatrixVectorizer=mv.MatrixVectorizer, normalize_adj_torch=up.normalize_adj_torch,
            pad_HR_adj=up.pad_HR_adj, unpad=up.unpad, modules=(gm, ml, mo, up, mv))
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # leave the reference's modules out of sys.modules so that nothing else
        # picks them up under the generic names ``models`` / ``utils``
        for k in list(sys.modules):
            if k in ("models", "utils") or k.startswith(("models.", "utils.")):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return ns


@contextlib.contextmanager
def canonical_library_calls():
    """Apply the two documented wrappers (oracle/gsr_oracle.py) to the library
    calls the reference makes: sign-canonical ``torch.linalg.eigh`` and
    stable-tie ``torch.topk``.  Identical wrappers on both sides is the parity
    contract of SURVEY.md section 0.6 / 0.7."""
    from . import gsr_oracle as O

    real_eigh, real_topk = torch.linalg.eigh, torch.topk

    def eigh(A, UPLO="L", **kw):
        w, U = real_eigh(A, UPLO=UPLO, **kw)
        return w, O.canonicalize_signs(U)

    def topk(x, k, *a, **kw):
        if a or kw or x.dim() != 1:
            return real_topk(x, k, *a, **kw)
        return O.stable_topk(x, k)

    torch.linalg.eigh, torch.topk = eigh, topk
    try:
        yield
    finally:
        torch.linalg.eigh, torch.topk = real_eigh, real_topk

"""MatrixVectorizer -- drop-in for the reference's utils/matrix_vectorizer.py, on the GPU.

Both static methods keep the reference's (mutually inconsistent) conventions bit for bit:
``vectorize`` walks the strict upper triangle COLUMN-major and its ``include_diagonal`` adds the
SUB-diagonal; ``anti_vectorize`` fills ROW-major and its ``include_diagonal`` fills the MAIN
diagonal.  Extensions: CUDA tensors and a leading batch dimension are accepted.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels as K

_GATHER_SIZES = (1, 2, 4, 8, 16)


class MatrixVectorizer:
    def __init__(self):
        pass

    @staticmethod
    def vectorize(matrix, include_diagonal=False):
        """utils/matrix_vectorizer.py:24-58.  numpy in -> numpy out (dtype preserved);
        a CUDA tensor in -> CUDA tensor out (no host round trip)."""
        if isinstance(matrix, torch.Tensor):
            if not matrix.is_cuda:
                raise RuntimeError("MatrixVectorizer.vectorize: torch input must live on the GPU "
                                   "(pass a numpy array for the reference's host-side behaviour)")
            return K.vectorize(matrix, include_diagonal)
        matrix = np.asarray(matrix)
        n = matrix.shape[0]
        if matrix.ndim != 2 or matrix.shape[1] != n:
            raise ValueError(f"MatrixVectorizer.vectorize: expected a square matrix, got {matrix.shape}")
        if n < 2:
            return np.array([])                   # the reference builds np.array([]) (float64)
        if matrix.dtype.itemsize not in _GATHER_SIZES or matrix.dtype.hasobject:
            raise TypeError(f"MatrixVectorizer.vectorize: unsupported dtype {matrix.dtype}")
        # dtype-agnostic byte gather: ship the raw bytes, view them back afterwards
        raw = np.ascontiguousarray(matrix).view(np.uint8).reshape(n, n * matrix.dtype.itemsize)
        carrier = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64, 16: torch.complex128}[matrix.dtype.itemsize]
        dev = torch.from_numpy(raw).cuda(non_blocking=False).view(carrier).reshape(n, n)
        out = K.vectorize(dev, include_diagonal)
        return out.cpu().numpy().view(matrix.dtype).reshape(-1)

    @staticmethod
    def anti_vectorize(vector, matrix_size, include_diagonal=False):
        """utils/matrix_vectorizer.py:61-95.  torch in -> float32 torch out on ``vector.device``."""
        if not isinstance(vector, torch.Tensor):
            # the reference fails here too: torch.zeros(device=None-from-hasattr) + numpy source
            raise TypeError("MatrixVectorizer.anti_vectorize: vector must be a torch.Tensor "
                            f"(got {type(vector).__name__}), as in the reference")
        n_elements = matrix_size * (matrix_size - 1) // 2
        assert vector.shape[-1] >= n_elements, "Vector size doesn't match required elements"
        if vector.dtype != torch.float32:
            raise RuntimeError("Index put requires the source and destination dtypes match, "
                               f"got Float for the destination and {vector.dtype} for the source.")
        if include_diagonal and vector.shape[-1] < n_elements + matrix_size:
            raise RuntimeError("shape mismatch: diagonal needs matrix_size more elements")
        src_device = vector.device
        out = K.anti_vectorize(vector.cuda(), matrix_size, include_diagonal)
        return out if src_device.type == "cuda" else out.to(src_device)

"""Mirror of the reference's ``utils`` package for the GSR-Net path."""
from .matrix_vectorizer import MatrixVectorizer
from .preprocessing import normalize_adj_torch, pad_HR_adj, unpad
from .optim import FlatAdam
from .submission import extract_matrices, submission_vector, upper_triangular_flatten, write_submission

__all__ = ["FlatAdam", "MatrixVectorizer", "normalize_adj_torch", "pad_HR_adj", "unpad",
           "extract_matrices", "submission_vector", "upper_triangular_flatten", "write_submission"]

"""Mirror of the reference's ``utils`` package for the GSR-Net path."""
from .matrix_vectorizer import MatrixVectorizer
from .preprocessing import normalize_adj_torch, pad_HR_adj, unpad

__all__ = ["MatrixVectorizer", "normalize_adj_torch", "pad_HR_adj", "unpad"]

"""rating_gp/models/kernels.py:4,315,357 build the gate's outer product as MatmulLinearOperator(a, b^T)."""
import torch


class MatmulLinearOperator:
    def __init__(self, left, right):
        self.left, self.right = left, right

    def to_dense(self):
        return self.left @ self.right

    def diagonal(self, dim1=-1, dim2=-2):
        return torch.diagonal(self.to_dense(), dim1=dim1, dim2=dim2)

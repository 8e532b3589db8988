"""``opt_einsum.contract`` restated as ``torch.einsum``.

ORACLE / TEST INFRASTRUCTURE ONLY.  Same contraction up to floating-point
reassociation.  Reference call sites: models/mace_modules/symmetric_contraction.py:152-183.
"""
import torch


def contract(equation: str, *operands):
    return torch.einsum(equation.replace(" ", ""), *operands)

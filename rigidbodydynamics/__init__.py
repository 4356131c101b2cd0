"""Namespace package: the product lives in ``rigidbodydynamics.jl_b200``."""

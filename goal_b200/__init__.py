"""goal_b200 -- B200-native finite-element assembly path for bgranzow/goal.

Only what the hot path needs: csrc/ (CUDA kernels + the C-ABI of include/goal_b200.h),
binding.py (ctypes mirror of the reference's compute_resid / compute_jacob / localize
call sites), synthetic.py and partition.py (structured tet meshes and their parts).
Importing the package does not load CUDA; constructing an Assembler does, and raises
if libgoal_b200.so or a GPU is missing (there is no CPU fallback).
"""
from .binding import ADJOINT, NONE, PRIMAL, Assembler, GxError, load_library  # noqa: F401

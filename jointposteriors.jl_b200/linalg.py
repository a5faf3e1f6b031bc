"""Host-side scale-matrix helpers, thin wrappers over the C ABI (no GPU needed).

Same names and meaning as the reference's functions (reference src/joint_posterior.jl:15-144):
chol!, try_chol!, inv!, inv_chol!, reduce_dimensions!, deduce_scale!.  Matrices are column-major
inside the library; these wrappers take and return ordinary numpy arrays.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, colmajor, lib, ptr


def chol(S):
    """Upper Cholesky factor U with U'U = S (chol!, reference src/joint_posterior.jl:30-43)."""
    S = colmajor(S)
    d = S.shape[0]
    U = np.zeros((d, d), order="F")
    check(lib().jp_chol(ptr(U), ptr(S), C.c_int(d)))
    return U


def try_chol(S):
    """(ok, U): ok is False when S is not positive definite (try_chol!, reference src/joint_posterior.jl:15-29)."""
    S = colmajor(S)
    d = S.shape[0]
    U = np.zeros((d, d), order="F")
    st = lib().jp_try_chol(ptr(U), ptr(S), C.c_int(d))
    if st == _lib.JP_ERR_NOT_PD:
        return False, U
    check(st)
    return True, U


def inv_upper(U):
    """Inverse of an upper-triangular matrix (inv!, reference src/joint_posterior.jl:56-68)."""
    U = colmajor(U).copy(order="F")
    check(lib().jp_inv_upper(ptr(U), C.c_int(U.shape[0])))
    return U


def inv_chol(H):
    """U = chol(H)^-1, so U U' = H^-1 (inv_chol!, reference src/joint_posterior.jl:72-76)."""
    H = colmajor(H)
    d = H.shape[0]
    U = np.zeros((d, d), order="F")
    check(lib().jp_inv_chol(ptr(U), ptr(H), C.c_int(d)))
    return U


def reduce_dimensions(H, max_rank=0):
    """d x p scale matrix from the eigenpairs with lambda >= 1e-11 (reference src/joint_posterior.jl:98-134)."""
    H = colmajor(H)
    d = H.shape[0]
    out = np.zeros((d, d), order="F")
    r = C.c_int()
    check(lib().jp_reduce_dimensions(ptr(H), C.c_int(d), C.c_int(int(max_rank)), ptr(out), C.byref(r)))
    return np.asfortranarray(out[:, :r.value])


def reduce_dimensions_ldr(H, g):
    """reduce_dimensions!(M, H, LDR{g}): leading eigen directions carrying the fraction g of the variance
    (reference src/joint_posterior.jl:78-95,111-119)."""
    H = colmajor(H)
    d = H.shape[0]
    out = np.zeros((d, d), order="F")
    r = C.c_int()
    check(lib().jp_reduce_dimensions_ldr(ptr(H), C.c_int(d), C.c_double(float(g)), ptr(out), C.byref(r)))
    return np.asfortranarray(out[:, :r.value])


def deduce_scale_dynamic(H):
    """deduce_scale!(M, H, Dynamic): Cholesky when possible, else eigen fallback (reference :136-138)."""
    H = colmajor(H)
    d = H.shape[0]
    U = np.zeros((d, d), order="F")
    r = C.c_int()
    check(lib().jp_deduce_scale_dynamic(ptr(H), C.c_int(d), ptr(U), C.byref(r)))
    return np.asfortranarray(U[:, :r.value])

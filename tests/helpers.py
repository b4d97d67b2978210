"""Shared helpers of the parity tests (not collected by pytest)."""
from __future__ import annotations

import itertools
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

# Tolerances (north_star): FP64 1e-12 relative, FP32 1e-5 relative; "relative" is taken against the
# largest magnitude of the reference result (norm-wise), which is what a BLAS-vs-BLAS comparison with
# different summation orders can promise. Index/shape handling and copies are bit-exact.
TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.complex128): 1e-12, np.dtype(np.float32): 1e-5, np.dtype(np.complex64): 1e-5}


def rel_err(got: np.ndarray, ref: np.ndarray) -> float:
    got = np.asarray(got)
    ref = np.asarray(ref)
    scale = max(1.0, float(np.max(np.abs(ref))) if ref.size else 1.0)
    return float(np.max(np.abs(got - ref))) / scale if ref.size else 0.0


def assert_close(got, ref, dtype=None, what=""):
    dtype = np.dtype(dtype or np.asarray(ref).dtype)
    e = rel_err(got, ref)
    assert e <= TOL[dtype], f"{what}: relative error {e:.3e} exceeds {TOL[dtype]:.0e}"


def rand(rng, shape, dtype=np.float64):
    """uniform [-1,1) like the reference's create_random_tensor
    (TensorUtilities/CreateRandomTensor.hpp:57-84); complex: both parts uniform."""
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        real = np.float32 if dtype == np.complex64 else np.float64
        return np.asarray(rng.uniform(-1, 1, shape).astype(real) + 1j * rng.uniform(-1, 1, shape).astype(real)).astype(dtype).reshape(shape)
    return np.asarray(rng.uniform(-1, 1, shape)).astype(dtype).reshape(shape)


def golden(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def shapes_for(packs, extents):
    return [tuple(extents[x] for x in p) for p in packs]


def np_einsum_ref(c_idx, Cn, a_idx, An, b_idx, Bn, alpha, beta, conj_a=False, conj_b=False):
    """Independent cross-check: numpy.einsum with explicit output; handles C-only (broadcast) indices."""
    letters = {}
    def sub(pack):
        return "".join(letters.setdefault(x, chr(ord("a") + len(letters))) for x in pack)
    sa, sb, sc = sub(a_idx), sub(b_idx), sub(c_idx)
    A = np.conj(An) if conj_a else An
    B = np.conj(Bn) if conj_b else Bn
    # indices only in C: broadcast
    only_c = [ch for ch in dict.fromkeys(sc) if ch not in sa and ch not in sb]
    sc_eff = "".join(ch for ch in dict.fromkeys(sc) if ch not in only_c)
    acc = np.result_type(An.dtype, Bn.dtype, np.float64)
    res = np.einsum(f"{sa},{sb}->{sc_eff}", A.astype(acc), B.astype(acc))
    wide = np.result_type(Cn.dtype, np.float64)
    out = np.array(np.array(Cn, dtype=wide) * beta if beta != 0 else np.zeros(Cn.shape, wide), dtype=wide, ndmin=0)
    res = np.asarray(res)
    # scatter res into out following sc (with possible repeated / broadcast indices)
    uniq = list(dict.fromkeys(sc))
    ext = {}
    for ch, n in zip(sc, Cn.shape):
        ext[ch] = n
    for idx in itertools.product(*[range(ext[ch]) for ch in uniq]):
        pos = dict(zip(uniq, idx))
        src = tuple(pos[ch] for ch in sc_eff)
        dst = tuple(pos[ch] for ch in sc)
        out[dst] += alpha * res[src]
    return out.astype(Cn.dtype)


# ---- water / STO-3G fixture -----------------------------------------------------------------
def load_water():
    w = golden("water_sto3g.json")
    n = 7
    def sym2(rows):
        M = np.zeros((n, n))
        for i, j, v in rows:
            M[i - 1, j - 1] = v
            M[j - 1, i - 1] = v
        return M
    S, T, V = sym2(w["S"]), sym2(w["T"]), sym2(w["V"])
    G = np.zeros((n, n, n, n))
    for i, j, k, l, v in w["TEI"]:
        i, j, k, l = i - 1, j - 1, k - 1, l - 1
        for (a, b, c, d) in ((i, j, k, l), (j, i, k, l), (i, j, l, k), (j, i, l, k), (k, l, i, j), (l, k, i, j), (k, l, j, i), (l, k, j, i)):
            G[a, b, c, d] = v
    return S, T, V, G, w["expected"]


def rhf_mp2_water(einsum):
    """The sequence of libs/Einsums/TensorAlgebra/tests/unit/QChemTests.cpp:185-416 (reference tree) with every
    contraction routed through `einsum(c_idx, C, a_idx, A, b_idx, B, alpha, beta)` (in place on numpy
    arrays). syev / pow(S,-1/2) come from numpy.linalg as the survey prescribes. Returns (E_rhf, E_mp2)."""
    S, T, V, TEI, _ = load_water()
    n, nocc = 7, 5
    H = T + V
    w, U = np.linalg.eigh(S)
    X = (U * w ** -0.5) @ U.T
    enuc = 8.002367061810450
    z2 = lambda: np.zeros((n, n))

    def transform(F):
        t1, Ft = z2(), z2()
        einsum("ij", t1, "ik", F, "kj", X, 1.0, 0.0)
        einsum("ij", Ft, "kj", t1, "ki", X, 1.0, 0.0)
        ev, Ct = np.linalg.eigh(Ft)
        Cm = z2()
        # reference: Ct holds eigenvectors in rows after syev; C[i,j] = sum_k Ct[k,j] X[i,k] with numpy's column convention
        einsum("ij", Cm, "kj", np.ascontiguousarray(Ct), "ik", X, 1.0, 0.0)
        return ev, Cm

    def density(Cm):
        Cocc = Cm[:, :nocc]  # strided view, as the reference's C(All, Range{0,5})
        D = z2()
        einsum("ij", D, "im", Cocc, "jm", Cocc, 1.0, 0.0)
        return D

    def energy(D, F):
        e = np.zeros(())
        einsum("", e, "ij", D, "ij", F, 1.0, 0.0)
        einsum("", e, "ij", D, "ij", H, 1.0, 1.0)
        return float(e) + enuc

    F = H.copy()
    ev, Cm = transform(F)
    D = density(Cm)
    e0, e1, cycles, drms = energy(D, F), 0.0, 0, 1.0
    while abs(e0 - e1) > 1e-10 and drms > 1e-6 and cycles < 50:
        cycles += 1
        e1, Dp = e0, D.copy()
        F = H.copy()
        einsum("ij", F, "kl", D, "ijkl", TEI, 2.0, 1.0)   # J: GEMV with swapped operands
        einsum("ij", F, "kl", D, "ikjl", TEI, -1.0, 1.0)  # K: GENERIC
        ev, Cm = transform(F)
        D = density(Cm)
        e0 = energy(D, F)
        diff = D - Dp
        r = np.zeros(())
        einsum("", r, "ij", diff, "ij", diff, 1.0, 0.0)
        drms = float(np.sqrt(r)) / 7.0
    assert cycles < 50
    # AO -> MO transformation, one index at a time
    t1, t2 = np.zeros((n,) * 4), np.zeros((n,) * 4)
    einsum("pqrs", t1, "mqrs", TEI, "mp", Cm, 1.0, 0.0)
    einsum("pqrs", t2, "pmrs", t1, "mq", Cm, 1.0, 0.0)
    einsum("pqrs", t1, "pqms", t2, "mr", Cm, 1.0, 0.0)
    MO = np.zeros((n,) * 4)
    einsum("pqrs", MO, "pqrm", t1, "ms", Cm, 1.0, 0.0)
    iajb = np.ascontiguousarray(MO[:nocc, nocc:, :nocc, nocc:])
    eo, evv = ev[:nocc], ev[nocc:]
    scale = 1.0 / (eo[:, None, None, None] + eo[None, None, :, None] - evv[None, :, None, None] - evv[None, None, None, :])
    amps = np.zeros_like(iajb)
    einsum("iajb", amps, "iajb", iajb, "iajb", np.ascontiguousarray(scale), 1.0, 0.0)  # DIRECT
    e_a, e_b = np.zeros(()), np.zeros(())
    einsum("", e_a, "iajb", amps, "iajb", iajb, 2.0, 0.0)   # DOT
    einsum("", e_b, "iajb", amps, "ibja", iajb, -1.0, 0.0)  # GENERIC (rank-0 C)
    return e0, float(e_a) + float(e_b)

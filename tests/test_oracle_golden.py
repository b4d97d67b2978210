"""CPU tests (no GPU): pin the oracle against the reference's own golden vectors, known-answer
tests and analyser decisions, and cross-check it against numpy.einsum."""
import itertools

import numpy as np
import pytest

from helpers import assert_close, golden, load_water, np_einsum_ref, rand, rhf_mp2_water


def test_algorithm_choice_golden(oracle):
    """Every `REQUIRE(alg_choice == X)` of the reference's unit tests (tests/golden/algorithm_choice.json)."""
    entries = golden("algorithm_choice.json")
    assert len(entries) >= 40
    checked = 0
    for e in entries:
        if e["source"].endswith("Einsum7.cpp:209"):
            continue  # run-time decision on a strided view: test_runtime_view_decisions
        got = oracle.classify_static(e["C"], e["A"], e["B"])
        assert got == e["expected"], f'{e["source"]}: {"".join(e["C"])}<-{"".join(e["A"])},{"".join(e["B"])} gave {got}, reference asserts {e["expected"]}'
        checked += 1
    assert checked >= 40


def test_runtime_view_decisions(oracle):
    """Einsum7.cpp:195-210: the GEMM pattern QX<-Qia,iaX on a strided slice of a larger tensor must fall to
    GENERIC at run time; Einsum1.cpp:104-127: 3x3 views with padding stay GEMM."""
    naux, nocc, nvirt, ur = 9, 3, 4, 5
    nmo = nocc + nvirt
    big = np.zeros((naux, nmo, nmo))
    qov = big[:, :nocc, nocc:]
    iax = np.zeros((nocc, nvirt, ur))
    nqx = np.zeros((naux, ur))
    assert oracle.classify("QX", nqx, "Qia", qov, "iaX", iax) == "GENERIC"
    assert oracle.classify("QX", nqx, "Qia", np.ascontiguousarray(qov), "iaX", iax) == "GEMM"
    A, B, Cb = np.zeros((3, 5)), np.zeros((3, 5)), np.zeros((10, 10))
    assert oracle.classify("ij", Cb[5:8, 5:8], "ik", A[:, :3], "jk", B[:, 2:5]) == "GEMM"


def test_conj_rules(oracle):
    # Dispatch.hpp:220-231: a direct product with a conjugated operand is not DIRECT
    assert oracle.classify_static("ij", "ij", "ij", conj_a=True) != "DIRECT"
    # :603-605 as written: `conjugate_works` = (!transpose_A || !ConjA) && ... for a non-transposed C, i.e. a
    # conjugated operand is accepted when it is NOT transposed and rejected when it is (the reference then calls
    # gemm with 'n', which drops the conjugation — a reference quirk; our device path applies conj itself).
    assert oracle.classify_static("ij", "ik", "kj", conj_a=True) == "GEMM"
    assert oracle.classify_static("ij", "ki", "kj", conj_a=True) == "GENERIC"
    # mixed element types always go generic (:896-899)
    assert oracle.classify_static("ij", "ik", "kj", same_types=False) == "GENERIC"
    # hadamard
    assert oracle.classify_static("i", "ii", "i") == "GENERIC"


def test_known_answers_einsum1(oracle):
    """Einsum1.cpp:17-88 (and the device twins GPUTensorAlgebra.hip:449-517)."""
    A = np.arange(1, 10, dtype=np.float64).reshape(3, 3)
    C = np.zeros((3, 3))
    oracle.einsum_generic("ij", C, "ik", A, "kj", A.copy())
    assert C.ravel().tolist() == [30, 36, 42, 66, 81, 96, 102, 126, 150]
    A3 = np.arange(1, 28, dtype=np.float64).reshape(3, 3, 3)
    C = np.zeros((3, 3))
    oracle.einsum_generic("il", C, "ijk", A3, "jkl", A3.copy())
    assert C.ravel().tolist() == [765, 810, 855, 1818, 1944, 2070, 2871, 3078, 3285]
    # same through the BLAS-routed baseline
    C2 = np.zeros((3, 3))
    assert oracle.einsum_f64("il", C2, "ijk", A3, "jkl", A3.copy()) == "GEMM"
    assert C2.ravel().tolist() == C.ravel().tolist()


CASES = [
    # (C, A, B, extents)
    ("ij", "ik", "kj", dict(i=5, j=7, k=3)),
    ("ij", "ik", "jk", dict(i=5, j=7, k=3)),
    ("ij", "ki", "kj", dict(i=5, j=7, k=3)),
    ("ji", "ik", "kj", dict(i=5, j=7, k=3)),          # transposed C
    ("i", "ij", "j", dict(i=6, j=4)),
    ("i", "j", "ij", dict(i=6, j=4)),
    ("", "ij", "ij", dict(i=6, j=4)),
    ("ij", "ij", "ij", dict(i=6, j=4)),
    ("ab", "a", "b", dict(a=6, b=4)),
    ("ijkl", "ijpl", "pk", dict(i=3, j=4, k=5, l=2, p=6)),
    ("pq", "prqs", "rs", dict(p=4, q=5, r=3, s=6)),   # Fock K
    ("pq", "pqrs", "rs", dict(p=4, q=5, r=3, s=6)),   # Fock J
    ("mbej", "jnfb", "mnef", dict(m=3, b=4, e=5, j=2, n=3, f=4)),  # Wmbej
    ("abef", "mnef", "mnab", dict(a=3, b=4, e=2, f=5, m=3, n=2)),  # Wabef
    ("mnij", "mnef", "ijef", dict(m=3, n=2, i=4, j=3, e=5, f=2)),  # Wmnij
    ("i", "ijk", "jik", dict(i=3, j=3, k=4)),          # Einsum3.cpp:78
    ("ik", "ijk", "jik", dict(i=3, j=3, k=4)),
    ("il", "ijk", "jik", dict(i=3, j=3, k=4, l=2)),    # broadcast index l (Einsum3.cpp:130-165)
    ("i", "ii", "i", dict(i=5)),                        # Hadamard in A
    ("ii", "i", "i", dict(i=5)),                        # Hadamard in C
    ("iaW", "iW", "aW", dict(i=3, a=4, W=5)),          # Khatri-Rao like, Einsum7.cpp:26
    ("ijb", "ikb", "kjb", dict(i=3, j=4, k=5, b=3)),   # batched
]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}<-{c[1]},{c[2]}" for c in CASES])
@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32])
def test_generic_vs_numpy(oracle, case, dtype):
    c_idx, a_idx, b_idx, ext = case
    rng = np.random.default_rng(1234)
    A = rand(rng, tuple(ext[x] for x in a_idx), dtype)
    B = rand(rng, tuple(ext[x] for x in b_idx), dtype)
    C0 = rand(rng, tuple(ext[x] for x in c_idx), dtype)
    for alpha, beta in ((1.0, 0.0), (0.5, 2.0)):
        C = C0.copy()
        oracle.einsum_generic(c_idx, C, a_idx, A, b_idx, B, alpha, beta)
        ref = np_einsum_ref(c_idx, C0, a_idx, A, b_idx, B, alpha, beta)
        if len(set(c_idx)) != len(c_idx) and beta == 0.0:
            pass  # repeated C index: the reference zeroes the whole tensor, np_einsum_ref starts from zeros too
        assert_close(C, ref, dtype, f"{case}")


def test_beta_zero_ignores_nan(oracle):
    """BaseAlgebra.hpp:117-121: beta == 0 zero-fills C, so NaNs already in C must not propagate."""
    rng = np.random.default_rng(0)
    A, B = rand(rng, (4, 5)), rand(rng, (5, 3))
    C = np.full((4, 3), np.nan)
    oracle.einsum_generic("ij", C, "ik", A, "kj", B, 1.0, 0.0)
    assert_close(C, A @ B)


def test_blas_route_matches_generic(oracle):
    rng = np.random.default_rng(7)
    oracle.load_openblas()
    for c_idx, a_idx, b_idx, ext in CASES:
        if len(set(c_idx)) != len(c_idx) or len(set(a_idx)) != len(a_idx):
            continue
        A = rand(rng, tuple(ext[x] for x in a_idx))
        B = rand(rng, tuple(ext[x] for x in b_idx))
        C0 = rand(rng, tuple(ext[x] for x in c_idx))
        C1, C2 = C0.copy(), C0.copy()
        oracle.einsum_generic(c_idx, C1, a_idx, A, b_idx, B, 0.5, 2.0)
        oracle.einsum_f64(c_idx, C2, a_idx, A, b_idx, B, 0.5, 2.0)
        assert_close(C2, C1, what=f"{c_idx}<-{a_idx},{b_idx}")


def test_permute_oracle(oracle):
    """Sort.cpp:18-250 semantics: every order of rank 2..4, alpha/beta mixes, unequal extents."""
    rng = np.random.default_rng(3)
    for rank in (1, 2, 3, 4):
        letters = "ijkl"[:rank]
        shape = (3, 4, 5, 2)[:rank]
        A = rand(rng, shape)
        for order in itertools.permutations(range(rank)):
            c_idx = "".join(letters[o] for o in order)
            C0 = rand(rng, tuple(shape[o] for o in order))
            for alpha, beta in ((1.0, 0.0), (0.5, 2.0)):
                C = C0.copy()
                oracle.permute(c_idx, C, letters, A, alpha, beta)
                ref = alpha * np.transpose(A, order) + (beta * C0 if beta else 0)
                if alpha == 1.0 and beta == 0.0:
                    assert np.array_equal(C, ref)
                else:
                    assert_close(C, ref)
                Cf = C0.copy()
                oracle.permute(c_idx, Cf, letters, A, alpha, beta, fast=True)
                assert_close(Cf, ref)
    # Sort.cpp:155-179: B[i,l,k,j] = 0.5 * A[k,j,l,i]
    A = rand(rng, (3, 3, 3, 3))
    Bt = np.zeros((3, 3, 3, 3))
    oracle.permute("ilkj", Bt, "kjli", A, 0.5, 0.0)
    assert_close(Bt, 0.5 * np.einsum("kjli->ilkj", A))


def test_water_sto3g_rhf_mp2(oracle):
    """QChemTests.cpp:185-416: RHF -74.942079928192, MP2 -0.049149636120 (+-1e-6), every contraction through the oracle."""
    def es(c, C, a, A, b, B, alpha, beta):
        oracle.einsum_generic(c, C, a, np.asarray(A), b, np.asarray(B), alpha, beta)
    e_rhf, e_mp2 = rhf_mp2_water(es)
    exp = load_water()[4]
    assert abs(e_rhf - exp["rhf_total_energy"]) < 1e-6
    assert abs(e_mp2 - exp["mp2_correlation_energy"]) < 1e-6
    assert abs(e_rhf + e_mp2 - exp["mp2_total_energy"]) < 1e-6

"""compile_plan / execute: the Python front end (reference: libs/EinsumsPy/TensorAlgebra/src/PyTensorAlgebra.cpp:168-552 and
tests/test_{generic,mat_product,mat_vec_product,dot_product,outer_product,direct_product}.py)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import TOL, rand, rel_err

GOLDEN_PLAN_TYPES = [  # the plan classes the reference's Python tests assert
    ("", "i", "i", "EinsumDotPlan"), ("i", "ij", "j", "EinsumGemvPlan"), ("i", "j", "ij", "EinsumGemvPlan"),
    ("ij", "i", "j", "EinsumGerPlan"), ("ij", "j", "i", "EinsumGerPlan"), ("ij", "ij", "ij", "EinsumDirectProductPlan"),
    ("ij", "ik", "kj", "EinsumGemmPlan"), ("ij", "ji", "ij", "EinsumGenericPlan"),
]


@pytest.mark.parametrize("c,a,b,want", GOLDEN_PLAN_TYPES)
def test_plan_types_reference_python_tests(eb, c, a, b, want):
    from einsums_b200 import compile_plan
    assert type(compile_plan(c, a, b)).__name__ == want


def test_plan_types_agree_with_cpp_analyser_goldens(eb, oracle):
    """Same decision as the C++ analyser's goldens (tests/golden/algorithm_choice.json) and as the oracle's static classifier
    wherever no conjugation is involved; the one run-time fallback case (Einsum7.cpp:195-210) is GEMM statically."""
    from einsums_b200 import compile_plan
    ents = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "algorithm_choice.json")))
    n = 0
    for e in ents:
        if e.get("conj_hint"):
            continue
        c, a, b = "".join(e["C"]), "".join(e["A"]), "".join(e["B"])
        got = compile_plan(c, a, b).route
        assert got == oracle.classify_static(c, a, b), (c, a, b)
        if not e["source"].endswith("Einsum7.cpp:209"):
            assert got == e["expected"], (e["source"], got)
        n += 1
    assert n >= 35


@pytest.mark.gpu
@pytest.mark.parametrize("c,a,b,want", GOLDEN_PLAN_TYPES)
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_plan_execute_matches_numpy(gpu, c, a, b, want, dtype):
    from einsums_b200 import compile_plan
    ext = {"i": 11, "j": 13, "k": 17}
    rng = np.random.default_rng(5)
    A, B = rand(rng, tuple(ext[x] for x in a), dtype), rand(rng, tuple(ext[x] for x in b), dtype)
    C0 = rand(rng, tuple(ext[x] for x in c), dtype)
    plan = compile_plan(c, a, b)
    for beta, alpha in ((0.0, 1.0), (0.5, -2.0)):
        Cd = torch.from_numpy(np.array(C0, copy=True)).cuda()
        plan.execute(beta, Cd, alpha, torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda())
        torch.cuda.synchronize()
        ref = beta * C0 + alpha * np.einsum(f"{a},{b}->{c}", A, B)
        assert rel_err(Cd.cpu().numpy(), ref) <= TOL[np.dtype(dtype)], (plan, beta, alpha)


@pytest.mark.gpu
def test_plan_execute_lists(gpu):
    """test_mat_prod_list / test_generic_list of the reference: ten independent items per call."""
    from einsums_b200 import compile_plan
    rng = np.random.default_rng(6)
    for c, a, b in (("ij", "ik", "kj"), ("ij", "ji", "ij")):
        ext = {"i": 11, "j": 13, "k": 17} if "k" in a + b else {"i": 11, "j": 11}
        As = [rand(rng, tuple(ext[x] for x in a)) for _ in range(10)]
        Bs = [rand(rng, tuple(ext[x] for x in b)) for _ in range(10)]
        Cs = [torch.zeros(tuple(ext[x] for x in c), dtype=torch.float64, device="cuda") for _ in range(10)]
        compile_plan(c, a, b).execute(0.0, Cs, 1.0, [torch.from_numpy(x).cuda() for x in As], [torch.from_numpy(x).cuda() for x in Bs])
        torch.cuda.synchronize()
        for Cd, Ah, Bh in zip(Cs, As, Bs):
            assert rel_err(Cd.cpu().numpy(), np.einsum(f"{a},{b}->{c}", Ah, Bh)) <= 1e-12
    with pytest.raises(ValueError, match="same size"):
        compile_plan("ij", "ik", "kj").execute(0.0, Cs, 1.0, [torch.zeros(2, 2, device="cuda")], [torch.zeros(2, 2, device="cuda")])

"""compile_plan(...).execute(...) — the Python front end of the reference (EinsumsPy) on top of the C ABI.

    plan = compile_plan("ij", "ik", "kj")        # -> EinsumGemmPlan
    plan.execute(0.0, C, 1.0, A, B)              # C, A, B: CUDA torch tensors (the GPUView role), or equal-length lists

Reference: libs/EinsumsPy/TensorAlgebra/src/PyTensorAlgebra.cpp:168-460 (compile_plan: DOT -> DIRECT -> GER -> GEMV ->
GEMM -> GENERIC over index strings) and the plan classes' execute (:470-...), tests/test_{generic,mat_product,
mat_vec_product,dot_product,outer_product,direct_product}.py. The plan class reports the route the reference would take;
every route executes through eb200_contract (lists: eb200_contract_grouped), where the engine picks the kernel family from
the actual strides — a GEMM-shaped plan on unit-stride operands lands on the TMA/DMMA GEMM kernel, anything else on the
fused-gather kernels, with identical results. No CPU fallback: tensors must live on the device.
"""
from __future__ import annotations

from . import ops

__all__ = ["compile_plan", "EinsumGenericPlan", "EinsumDotPlan", "EinsumDirectProductPlan", "EinsumGerPlan", "EinsumGemvPlan",
           "EinsumGemmPlan"]


def _with_pos(keys: str, where: str):
    """(index, position in `where`) for every occurrence, keys in order (reference: find_ind_with_pos)."""
    return [(k, p) for k in keys for p, w in enumerate(where) if w == k]


def _contiguous(pairs) -> bool:
    return all(pairs[i][1] + 1 == pairs[i + 1][1] for i in range(len(pairs) - 1))


def _same_order(x, y) -> bool:
    return len(x) == len(y) and len(x) > 0 and all(a[0] == b[0] for a, b in zip(x, y))


def _uniq(s: str) -> str:
    return "".join(dict.fromkeys(s))


class EinsumGenericPlan:
    """Base plan: holds the index strings; execute() runs C = c_prefactor*C + ab_prefactor * sum A*B."""
    route = "GENERIC"

    def __init__(self, c: str, a: str, b: str):
        self.c, self.a, self.b = c, a, b

    def execute(self, c_prefactor, C, ab_prefactor, A, B, conj_a: bool = False, conj_b: bool = False):
        if isinstance(C, (list, tuple)):
            if not (isinstance(A, (list, tuple)) and isinstance(B, (list, tuple))) or not (len(A) == len(B) == len(C)):
                raise ValueError("Lists passed to batched einsum call do not have the same size!")
            ops.contract_grouped(self.c, list(C), self.a, list(A), self.b, list(B), ab_prefactor, c_prefactor, conj_a, conj_b)
            return
        for name, idx, t in (("C", self.c, C), ("A", self.a, A), ("B", self.b, B)):
            if t.dim() != len(idx):
                raise ValueError(f"rank_error: tensor {name} has rank {t.dim()} but {len(idx)} indices were given")
        ops.contract(self.c, C, self.a, A, self.b, B, ab_prefactor, c_prefactor, conj_a, conj_b)

    def __repr__(self):
        return f"{type(self).__name__}('{self.c}' <- '{self.a}', '{self.b}')"


class EinsumDotPlan(EinsumGenericPlan):
    route = "DOT"


class EinsumDirectProductPlan(EinsumGenericPlan):
    route = "DIRECT"


class EinsumGerPlan(EinsumGenericPlan):
    route = "GER"


class EinsumGemvPlan(EinsumGenericPlan):
    route = "GEMV"


class EinsumGemmPlan(EinsumGenericPlan):
    route = "GEMM"


def compile_plan(c: str, a: str, b: str) -> EinsumGenericPlan:
    """Classify the contraction from its index strings alone, in the reference's order."""
    links_ab = "".join(x for x in a if x in b)
    links = "".join(x for x in links_ab if x not in c)
    link_u, c_u = _uniq(links), _uniq(c)
    hadamard = len(_uniq(a)) != len(a) or len(_uniq(b)) != len(b) or len(c_u) != len(c)

    link_in_a, link_in_b = _with_pos(link_u, a), _with_pos(link_u, b)
    target_in_a, target_in_b = _with_pos(c_u, a), _with_pos(c_u, b)
    a_target_in_c, b_target_in_c = _with_pos(a, c), _with_pos(b, c)

    links_ok = _contiguous(link_in_a) and _contiguous(link_in_b) and _same_order(link_in_a, link_in_b)
    ca_same, cb_same = _same_order(target_in_a, a_target_in_c), _same_order(target_in_b, b_target_in_c)

    if c == "" and a == b and not hadamard:
        return EinsumDotPlan(c, a, b)
    if a == c and b == c and not hadamard:
        return EinsumDirectProductPlan(c, a, b)
    if not links_ab and _contiguous(target_in_a) and _contiguous(target_in_b) and not hadamard:
        return EinsumGerPlan(c, a, b)
    gemv_ab = links_ok and _contiguous(target_in_a) and ca_same and not cb_same and not b_target_in_c and not hadamard
    gemv_ba = links_ok and _contiguous(target_in_b) and cb_same and not ca_same and not a_target_in_c and not hadamard
    if gemv_ab or gemv_ba:
        return EinsumGemvPlan(c, a, b)
    c_minus_a = [x for x in c if x not in a]
    c_minus_b = [x for x in c if x not in b]
    if (c_minus_a and c_minus_b and links_ok and _contiguous(target_in_a) and _contiguous(target_in_b) and _contiguous(a_target_in_c)
            and _contiguous(b_target_in_c) and ca_same and cb_same and not hadamard):
        return EinsumGemmPlan(c, a, b)
    return EinsumGenericPlan(c, a, b)

"""GPU parity tests (pytest -m gpu, on the B200): every kernel family of libeinsums_b200.so against the CPU
oracle on the same seeded inputs, through the C ABI. FP64 tolerance 1e-12 relative, FP32 1e-5 (helpers.TOL);
copies/permutes with alpha=1,beta=0 bit-exact."""
import itertools

import numpy as np
import pytest
import torch

from helpers import TOL, assert_close, golden, load_water, rand, rel_err, rhf_mp2_water
from test_oracle_golden import CASES

pytestmark = pytest.mark.gpu

NP2T = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32, np.dtype(np.complex128): torch.complex128,
        np.dtype(np.complex64): torch.complex64}


def dev(x):
    x = np.asarray(x)
    return torch.from_numpy(np.array(x, order="C", copy=True)).cuda()


def dev_like(x: np.ndarray):
    """Upload preserving the numpy strides (views with padding stay views)."""
    if x.ndim == 0 or x.flags.c_contiguous:
        return dev(x)
    base = x.base if x.base is not None else x
    while base.base is not None:
        base = base.base
    tb = torch.from_numpy(np.ascontiguousarray(base)).cuda() if base.flags.c_contiguous else None
    if tb is None:
        return dev(x)
    off = (x.__array_interface__["data"][0] - base.__array_interface__["data"][0]) // x.itemsize
    return torch.as_strided(tb.view(-1), x.shape, [s // x.itemsize for s in x.strides], off)


def run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, dtype, alpha, beta, conj_a=False, conj_b=False, force=0, seed=0, split_k=0):
    rng = np.random.default_rng(seed)
    A = rand(rng, tuple(ext[x] for x in a_idx), dtype)
    B = rand(rng, tuple(ext[x] for x in b_idx), dtype)
    C0 = rand(rng, tuple(ext[x] for x in c_idx), dtype)
    ref = C0.copy()
    oracle.einsum_generic(c_idx, ref, a_idx, A, b_idx, B, alpha, beta, conj_a, conj_b)
    Cd = dev(C0)
    d = gpu.contract(c_idx, Cd, a_idx, dev(A), b_idx, dev(B), alpha, beta, conj_a, conj_b, force_kernel=force, split_k=split_k)
    torch.cuda.synchronize()
    assert_close(Cd.cpu().numpy(), ref, dtype, f"{c_idx}<-{a_idx},{b_idx} {np.dtype(dtype)} alpha={alpha} beta={beta} kernel={d.kernel_used}")
    return d


# ---------------------------------------------------------------------------------------------
def test_known_answers_einsum1_device(gpu):
    """GPUTensorAlgebra.hip:449-517: exact integers."""
    A = dev(np.arange(1, 10, dtype=np.float64).reshape(3, 3))
    C = torch.zeros(3, 3, dtype=torch.float64, device="cuda")
    gpu.contract("ij", C, "ik", A, "kj", A.clone())
    assert C.cpu().numpy().ravel().tolist() == [30, 36, 42, 66, 81, 96, 102, 126, 150]
    A3 = dev(np.arange(1, 28, dtype=np.float64).reshape(3, 3, 3))
    C = torch.zeros(3, 3, dtype=torch.float64, device="cuda")
    gpu.contract("il", C, "ijk", A3, "jkl", A3.clone())
    assert C.cpu().numpy().ravel().tolist() == [765, 810, 855, 1818, 1944, 2070, 2871, 3078, 3285]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}<-{c[1]},{c[2]}" for c in CASES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex128, np.complex64])
def test_contract_small_all_patterns(gpu, oracle, case, dtype):
    c_idx, a_idx, b_idx, ext = case
    for alpha, beta in ((1.0, 0.0), (0.5, 2.0), (-1.0, 1.0)):
        run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, dtype, alpha, beta)


def _random_spec(rng):
    """A random legal contraction: 2-6 unique indices, each placed in a random non-empty subset of {C, A, B} (indices only
    in C broadcast, indices in A and/or B but not C are summed, indices everywhere are batch/Hadamard), random axis orders,
    sometimes a repeated index inside one operand, extents 1-7 (1 and primes on purpose)."""
    letters = "abcdefgh"
    nu = int(rng.integers(2, 7))
    ext, c, a, b = {}, [], [], []
    for u in range(nu):
        ch = letters[u]
        ext[ch] = int(rng.choice([1, 2, 3, 4, 5, 7]))
        where = int(rng.integers(1, 8))  # bit 0: C, bit 1: A, bit 2: B
        if where & 1: c.append(ch)
        if where & 2: a.append(ch)
        if where & 4: b.append(ch)
    if not a: a.append(letters[0]); ext.setdefault(letters[0], 2)
    if not b: b.append(letters[min(1, nu - 1)])
    if rng.random() < 0.2 and len(a) < 4: a.append(a[int(rng.integers(0, len(a)))])  # Hadamard inside A
    for lst in (c, a, b):
        rng.shuffle(lst)
    return "".join(c), "".join(a), "".join(b), ext


@pytest.mark.parametrize("chunk", range(8))
def test_contract_fuzz(gpu, oracle, chunk):
    """Seeded random contractions (25 per chunk, 200 in total) against the oracle's generic loops, all four value types, random
    prefactors and conjugations. The device semantics (sum over the full unique space) are the oracle's `full_space` mode."""
    rng = np.random.default_rng(1000 + chunk)
    for n in range(25):
        c_idx, a_idx, b_idx, ext = _random_spec(rng)
        dtype = [np.float64, np.float32, np.complex128, np.complex64][int(rng.integers(0, 4))]
        alpha, beta = [(1.0, 0.0), (0.5, 2.0), (-1.0, 1.0), (2.0, 0.0)][int(rng.integers(0, 4))]
        conj_a, conj_b = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        try:
            run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, dtype, alpha, beta, conj_a, conj_b, seed=chunk * 100 + n)
        except AssertionError as e:
            raise AssertionError(f"fuzz case {chunk}/{n}: {c_idx}<-{a_idx},{b_idx} ext={ext} {np.dtype(dtype)} a={alpha} b={beta} conj={conj_a},{conj_b}: {e}") from e


@pytest.mark.parametrize("chunk", range(4))
def test_contract_fuzz_mid_strided(gpu, oracle, chunk):
    """Random contractions at tile-kernel sizes (ragged extents 1...40, index space up to 2e7 points) on padded, strided
    views of larger buffers: bytes outside the views must stay untouched and every family / split-K choice must agree
    with the oracle."""
    rng = np.random.default_rng(2000 + chunk)
    done = 0
    while done < 10:
        letters = "abcde"
        nu = int(rng.integers(3, 6))
        ext = {letters[u]: int(rng.choice([1, 3, 8, 17, 33, 40])) for u in range(nu)}
        if np.prod([float(v) for v in ext.values()]) > 2e7:
            continue
        c, a, b = [], [], []
        for u in range(nu):
            where = int(rng.choice([3, 5, 6, 6, 6, 7, 3, 5]))  # every index touches at least two tensors; 6 = contracted
            ch = letters[u]
            if where & 1: c.append(ch)
            if where & 2: a.append(ch)
            if where & 4: b.append(ch)
        if not a or not b:
            continue
        for lst in (c, a, b):
            rng.shuffle(lst)
        c_idx, a_idx, b_idx = "".join(c), "".join(a), "".join(b)
        dtype = [np.float64, np.float64, np.float32, np.complex128][int(rng.integers(0, 4))]
        alpha, beta = [(1.0, 0.0), (0.5, 2.0), (-1.0, 1.0)][int(rng.integers(0, 3))]

        def padded(idx):
            shape = tuple(ext[x] for x in idx)
            pad = tuple(int(rng.integers(0, 3)) for _ in idx)
            big = rand(rng, tuple(s + p for s, p in zip(shape, pad)), dtype)
            return big, big[tuple(slice(0, s) for s in shape)] if idx else big

        Abig, Av = padded(a_idx); Bbig, Bv = padded(b_idx); Cbig, Cv = padded(c_idx)
        ref_big = Cbig.copy()
        ref_view = ref_big[tuple(slice(0, ext[x]) for x in c_idx)] if c_idx else ref_big
        tmp = np.ascontiguousarray(ref_view)
        oracle.einsum_generic(c_idx, tmp, a_idx, np.ascontiguousarray(Av), b_idx, np.ascontiguousarray(Bv), alpha, beta)
        ref_view[...] = tmp
        Cd_big = dev(Cbig)
        Cd = Cd_big[tuple(slice(0, ext[x]) for x in c_idx)] if c_idx else Cd_big
        Ad = dev(Abig)[tuple(slice(0, ext[x]) for x in a_idx)]
        Bd = dev(Bbig)[tuple(slice(0, ext[x]) for x in b_idx)]
        d = gpu.contract(c_idx, Cd, a_idx, Ad, b_idx, Bd, alpha, beta)
        torch.cuda.synchronize()
        got = Cd_big.cpu().numpy()
        what = f"fuzz-mid {chunk}/{done}: {c_idx}<-{a_idx},{b_idx} ext={ext} {np.dtype(dtype)} a={alpha} b={beta} kernel={d.kernel_used}"
        assert_close(got, ref_big, dtype, what)
        mask = np.ones(Cbig.shape, bool)
        if c_idx:
            mask[tuple(slice(0, ext[x]) for x in c_idx)] = False
            assert np.array_equal(got[mask], Cbig[mask]), what + " wrote outside the view"
        done += 1


@pytest.mark.parametrize("conj", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[4], CASES[6], CASES[7], CASES[8], CASES[10]], ids=lambda c: f"{c[0]}<-{c[1]},{c[2]}")
def test_contract_conj(gpu, oracle, case, conj):
    c_idx, a_idx, b_idx, ext = case
    run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, np.complex128, 0.5 + 0.25j, 2.0 - 1.0j, conj[0], conj[1])


MID = [
    # benchmark-config patterns at sizes that cross tile boundaries (ragged edges) and exercise split-K
    ("abef", "mnef", "mnab", dict(a=13, b=11, e=12, f=14, m=5, n=6)),            # Wabef: M,N contiguous operands
    ("mnij", "mnef", "ijef", dict(m=7, n=6, i=5, j=9, e=31, f=29)),              # Wmnij: K contiguous, split-K
    ("mbej", "jnfb", "mnef", dict(m=6, b=21, e=19, j=7, n=5, f=23)),             # Wmbej: fused permutation
    ("pq", "prqs", "rs", dict(p=37, q=41, r=33, s=35)),                          # Fock K
    ("pq", "pqrs", "rs", dict(p=37, q=41, r=33, s=35)),                          # Fock J
    ("ij", "ik", "kj", dict(i=257, j=131, k=193)),
    ("ij", "ki", "jk", dict(i=130, j=258, k=77)),
    ("ji", "ik", "kj", dict(i=129, j=70, k=300)),
    ("ijb", "ikb", "kjb", dict(i=65, j=33, k=40, b=3)),                          # batch index innermost
    ("bij", "bik", "bkj", dict(i=65, j=33, k=40, b=5)),                          # batch index outermost
    ("q", "pq", "p", dict(p=1000, q=777)),                                       # col reduce
    ("p", "pq", "q", dict(p=777, q=1000)),                                       # row reduce
    ("", "ijab", "ijab", dict(i=17, j=19, a=23, b=29)),                          # dot
    ("pn", "pqrs", "qrsn", dict(p=200, q=11, r=12, s=13, n=3)),                  # few columns (NN=4)
    ("np", "qrsn", "pqrs", dict(p=200, q=11, r=12, s=13, n=7)),                  # few rows on the A side -> swap
]


@pytest.mark.parametrize("case", MID, ids=lambda c: f"{c[0]}<-{c[1]},{c[2]}")
def test_contract_mid_f64(gpu, oracle, case):
    c_idx, a_idx, b_idx, ext = case
    run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, np.float64, 0.25, 1.0, seed=11)
    run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, np.float64, 1.0, 0.0, seed=12)


@pytest.mark.parametrize("case", MID[:8], ids=lambda c: f"{c[0]}<-{c[1]},{c[2]}")
def test_kernel_families_agree(gpu, oracle, case):
    """The same contraction forced through the tensor-pipe tile kernel, the FMA tile kernel and (when the shape
    allows) the streaming reductions must all match the oracle."""
    from einsums_b200 import KF
    c_idx, a_idx, b_idx, ext = case
    used = set()
    for fam in ("TILE_MMA", "TILE_SIMT", "ROWREDUCE", "COLREDUCE"):
        d = run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, np.float64, 0.5, 2.0, force=KF[fam], seed=5)
        used.add(d.kernel_used)
    assert KF["TILE_MMA"] in used and KF["TILE_SIMT"] in used
    for sk in (1, 2, 5):
        run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, np.float64, 0.5, 2.0, force=KF["TILE_MMA"], seed=6, split_k=sk)


@pytest.mark.parametrize("dtype", [np.float32, np.complex128])
def test_contract_mid_other_types(gpu, oracle, dtype):
    for case in (MID[0], MID[2], MID[3], MID[5], MID[10], MID[12]):
        c_idx, a_idx, b_idx, ext = case
        run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, dtype, 0.25, 1.0, seed=21)


@pytest.mark.parametrize("tc,ta,tb", [(np.float64, np.float32, np.float64), (np.float64, np.float32, np.float32), (np.float32, np.float64, np.float64),
                                      (np.complex128, np.complex128, np.float64), (np.complex64, np.float32, np.complex64)])
def test_contract_mixed_types(gpu, tc, ta, tb):
    """MixedPrecision.cpp:43-58 / GPUTensorAlgebra.hip:2334-2354 (d-f-d, d-f-f, f-d-d, cd-cd-d): operands are widened on the device,
    the product is formed in the promoted type. Reference tolerance there is 1e-4 absolute; we hold 1e-5 relative."""
    rng = np.random.default_rng(33)
    for (c_idx, a_idx, b_idx, ext) in (("ij", "ik", "kj", dict(i=10, j=10, k=10)), ("pq", "prqs", "rs", dict(p=6, q=7, r=5, s=9)),
                                       ("", "ij", "ij", dict(i=7, j=9)), ("ij", "ij", "ij", dict(i=7, j=9))):
        A = rand(rng, tuple(ext[x] for x in a_idx), ta)
        B = rand(rng, tuple(ext[x] for x in b_idx), tb)
        C0 = rand(rng, tuple(ext[x] for x in c_idx), tc)
        for alpha, beta in ((1.0, 0.0), (0.5, 2.0)):
            wide = np.result_type(tc, ta, tb, np.float64)
            ref = (beta * C0.astype(wide) if beta else 0) + alpha * np.einsum(f"{a_idx},{b_idx}->{c_idx}", A.astype(wide), B.astype(wide))
            if np.dtype(tc).kind != "c":
                ref = np.real(ref)
            Cd = dev(C0)
            gpu.contract(c_idx, Cd, a_idx, dev(A), b_idx, dev(B), alpha, beta)
            got = Cd.cpu().numpy()
            assert got.dtype == np.dtype(tc)
            assert rel_err(got.astype(wide), np.asarray(ref)) <= 1e-5, f"{np.dtype(tc)}<-{np.dtype(ta)},{np.dtype(tb)} {c_idx}<-{a_idx},{b_idx}"


def test_contract_strided_views(gpu, oracle):
    """Views with padding (Einsum1.cpp:96-140, Einsum7.cpp:195-210): strides are honoured, bytes outside the view untouched."""
    rng = np.random.default_rng(4)
    big = rand(rng, (9, 7, 7))
    qov = big[:, :3, 3:]
    iax = rand(rng, (3, 4, 5))
    Cfull = rand(rng, (12, 8))
    Cview = Cfull[2:11, 1:6]
    ref_full = Cfull.copy()
    oracle.einsum_generic("QX", ref_full[2:11, 1:6], "Qia", qov, "iaX", iax, 0.5, 2.0)
    bd, cd = dev(big), dev(Cfull)
    gpu.contract("QX", cd[2:11, 1:6], "Qia", bd[:, :3, 3:], "iaX", dev(iax), 0.5, 2.0)
    assert_close(cd.cpu().numpy(), ref_full)
    mask = np.ones_like(Cfull, bool)
    mask[2:11, 1:6] = False
    assert np.array_equal(cd.cpu().numpy()[mask], Cfull[mask])


def test_contract_many_batches(gpu):
    """More batches than grid.z can hold (65535): the launcher walks the batch index in chunks."""
    nb = 70001
    A = torch.rand(nb, 9, 10, dtype=torch.float64, device="cuda") - 0.5
    B = torch.rand(nb, 10, 11, dtype=torch.float64, device="cuda") - 0.5
    C = torch.zeros(nb, 9, 11, dtype=torch.float64, device="cuda")
    d = gpu.contract("bij", C, "bik", A, "bkj", B)
    assert d.launches == 2
    ref = torch.bmm(A, B)
    assert rel_err(C.cpu().numpy(), ref.cpu().numpy()) <= 1e-12
    Cf = torch.zeros(nb, 9, 11, dtype=torch.float32, device="cuda")
    gpu.contract("bij", Cf, "bik", A.float(), "bkj", B.float())
    assert rel_err(Cf.cpu().numpy(), ref.cpu().numpy()) <= 1e-5


# ---------------------------------------------------------------------------------------------
# grouped contraction = the per-block einsum loop of block-diagonal tensors (BlockAlgebra.hpp:36-47)
def _block_views(flat, sizes, rank):
    """Rank-`rank` hypercubic blocks packed back to back in one allocation (how BlockDeviceTensor stores them)."""
    out, off = [], 0
    for n in sizes:
        out.append(flat[off:off + n**rank].view(*([n] * rank)))
        off += n**rank
    return out


@pytest.mark.parametrize("sizes,expect_groups", [((48,) * 6, 1), ((5, 17, 64, 33, 1), 5), ((16, 16, 16, 40, 40, 7), 3)])
@pytest.mark.parametrize("dtype", [np.float64, np.complex64])
def test_grouped_block_diagonal(gpu, oracle, sizes, expect_groups, dtype):
    rng = np.random.default_rng(11)
    tot = sum(n * n for n in sizes)
    Ah, Bh, Ch = rand(rng, (tot,), dtype), rand(rng, (tot,), dtype), rand(rng, (tot,), dtype)
    Ad, Bd, Cd = dev(Ah), dev(Bh), dev(Ch)
    descs, groups = gpu.contract_grouped("ij", _block_views(Cd, sizes, 2), "ik", _block_views(Ad, sizes, 2), "jk", _block_views(Bd, sizes, 2), 0.5, 2.0)
    torch.cuda.synchronize()
    assert groups == expect_groups
    assert sum(d.launches for d in descs) >= groups
    ref, off = Ch.copy(), 0
    for n in sizes:  # the reference: one einsum per block
        sl = slice(off, off + n * n)
        blk = np.ascontiguousarray(ref[sl].reshape(n, n))
        oracle.einsum_generic("ij", blk, "ik", np.ascontiguousarray(Ah[sl].reshape(n, n)), "jk", np.ascontiguousarray(Bh[sl].reshape(n, n)), 0.5, 2.0)
        ref[sl] = blk.reshape(-1)
        off += n * n
    assert_close(Cd.cpu().numpy(), ref, dtype, f"grouped {sizes}")


def test_grouped_shared_operand_and_rank3(gpu, oracle):
    """A vector shared by every block (BlockAlgebra.hpp:185-214) folds into an N index; rank-3 blocks fold into a batch."""
    rng = np.random.default_rng(12)
    nb, n = 5, 24
    Ah, Ch = rand(rng, (nb, n, n, n), np.float64), np.zeros((nb, n, n), np.float64)
    Mh = rand(rng, (n, n), np.float64)
    Ad, Cd, Md = dev(Ah), dev(Ch), dev(Mh)
    descs, groups = gpu.contract_grouped("ij", [Cd[b] for b in range(nb)], "ikl", [Ad[b] for b in range(nb)], "lkj", [Md.view(n, 1, n).expand(n, n, n)] * nb)
    torch.cuda.synchronize()
    assert groups == 1
    ref = np.einsum("bikl,lj->bij", Ah, Mh)
    assert rel_err(Cd.cpu().numpy(), ref) <= 1e-12
    # outputs at irregular distances: only pairs share a step (any two equal-shape neighbours do) -> 2 + 2 + 1 dispatches
    pool = torch.zeros(nb * (n * n + 7) + 64, dtype=torch.float64, device="cuda")
    Cs = [pool[b * (n * n + 7) + b * b:][:n * n].view(n, n) for b in range(nb)]
    descs, groups = gpu.contract_grouped("ij", Cs, "ikl", [Ad[b] for b in range(nb)], "lkj", [Md.view(n, 1, n).expand(n, n, n)] * nb)
    torch.cuda.synchronize()
    assert groups == 3
    assert rel_err(torch.stack(Cs).cpu().numpy(), ref) <= 1e-12
    with pytest.raises(ValueError, match="same number of blocks"):
        gpu.contract_grouped("ij", Cs, "ikl", [Ad[0]], "lkj", [Md] * nb)


@pytest.mark.parametrize("present", ["all", "checkerboard", "random"])
def test_grouped_tile_grid_folds_to_k_and_batches(gpu, oracle, present):
    """A tiled product C(i,j) += A(i,k) B(k,j) handed over as one descriptor per tile triple (TileAlgebra.hpp:57-90), tiles packed
    grid-major in one buffer: contributions to one C tile fold into a longer K, then the grid folds into batches."""
    g, t = 4, 24
    rng = np.random.default_rng(21)
    Ah, Bh = rand(rng, (g, g, t, t), np.float64), rand(rng, (g, g, t, t), np.float64)
    Ch = rand(rng, (g, g, t, t), np.float64)
    keep = {"all": lambda i, j: True, "checkerboard": lambda i, j: (i + j) % 2 == 0, "random": lambda i, j: (i * 7 + j * 3) % 5 != 0}[present]
    Ad, Bd, Cd = dev(Ah), dev(Bh), dev(Ch)
    Cs, As, Bs = [], [], []
    ref = Ch.copy()
    for i in range(g):
        for j in range(g):
            for k in range(g):
                if keep(i, k) and keep(k, j):
                    Cs.append(Cd[i, j]); As.append(Ad[i, k]); Bs.append(Bd[k, j])
                    blk = np.ascontiguousarray(ref[i, j])
                    oracle.einsum_generic("ij", blk, "ik", np.ascontiguousarray(Ah[i, k]), "kj", np.ascontiguousarray(Bh[k, j]), 1.0, 1.0)
                    ref[i, j] = blk
    descs, groups = gpu.contract_grouped("ij", Cs, "ik", As, "kj", Bs, 1.0, 1.0)
    torch.cuda.synchronize()
    assert_close(Cd.cpu().numpy(), ref, np.float64, f"tile grid {present}")
    if present == "all":
        assert groups == 1 and sum(d.launches for d in descs) == 1
    else:
        assert groups < len(Cs)
    # beta != 1 with a repeated output must not fold into K (it would apply beta once instead of per member)
    C2 = dev(Ch)
    descs, groups = gpu.contract_grouped("ij", [C2[0, 0]] * 2, "ik", [Ad[0, 0], Ad[0, 1]], "kj", [Bd[0, 0], Bd[1, 0]], 1.0, 0.5)
    torch.cuda.synchronize()
    want = 0.5 * (0.5 * Ch[0, 0] + Ah[0, 0] @ Bh[0, 0]) + Ah[0, 1] @ Bh[1, 0]
    assert groups == 2 and rel_err(C2[0, 0].cpu().numpy(), want) <= 1e-12


# ---------------------------------------------------------------------------------------------
# host-resident tensors through eb200_contract_host (slab pipeline)
@pytest.mark.parametrize("beta", [0.0, 1.0])
@pytest.mark.parametrize("pinned", [True, False])
def test_contract_host_slabs(gpu, oracle, beta, pinned):
    o, v = 5, 24
    rng = np.random.default_rng(31)
    g, t, W0 = rand(rng, (o, o, v, v)), rand(rng, (o, o, v, v)), rand(rng, (v, v, v, v))
    ref = W0.copy()
    oracle.einsum_generic("abef", ref, "mnef", g, "mnab", t, 0.25, beta)
    mk = (lambda x: torch.from_numpy(x.copy()).pin_memory()) if pinned else (lambda x: torch.from_numpy(x.copy()))
    Wh, gh, th = mk(W0), mk(g), mk(t)
    st = gpu.contract_host("abef", Wh, "mnef", gh, "mnab", th, 0.25, beta, slab_bytes=3 * v**3 * 8)   # 3 rows of `a` per slab
    assert st.slabs == v // 3
    assert st.d2h_bytes == W0.nbytes and st.h2d_bytes == g.nbytes + t.nbytes + (W0.nbytes if beta else 0)
    assert_close(Wh.numpy(), ref, np.float64, f"contract_host beta={beta} pinned={pinned}")
    # bit-identical to the resident path
    Wd = dev(W0)
    gpu.contract("abef", Wd, "mnef", dev(g), "mnab", dev(t), 0.25, beta)
    assert np.array_equal(Wd.cpu().numpy(), Wh.numpy())


def test_contract_host_mixed_residency_scalar_and_errors(gpu, oracle):
    rng = np.random.default_rng(32)
    A, B, C0 = rand(rng, (40, 30)), rand(rng, (30, 50)), rand(rng, (40, 50))
    want = 2.0 * C0 + A @ B
    # operands on the device, C on the host; the slab index lives in A
    Ch = torch.from_numpy(C0.copy()).pin_memory()
    st = gpu.contract_host("ij", Ch, "ik", dev(A), "kj", dev(B), 1.0, 2.0, slab_bytes=7 * 50 * 8)
    assert st.slabs == 6 and st.h2d_bytes == C0.nbytes
    assert rel_err(Ch.numpy(), want) <= 1e-12
    # the slab index lives in B (C = B^T-like order), A on the host
    Ch = torch.from_numpy(np.ascontiguousarray(C0.T)).pin_memory()
    gpu.contract_host("ji", Ch, "ik", torch.from_numpy(A.copy()).pin_memory(), "kj", dev(B), 1.0, 2.0, slab_bytes=4096)
    assert rel_err(Ch.numpy(), want.T) <= 1e-12
    # host operands, device C
    Cd = dev(C0)
    gpu.contract_host("ij", Cd, "ik", torch.from_numpy(A.copy()).pin_memory(), "kj", torch.from_numpy(B.copy()), 1.0, 2.0)
    assert rel_err(Cd.cpu().numpy(), want) <= 1e-12
    # scalar on the host
    s = torch.tensor(3.0, dtype=torch.float64).pin_memory()
    gpu.contract_host("", s, "ik", torch.from_numpy(A.copy()), "ik", torch.from_numpy(A.copy()), 1.0, 0.5)
    assert abs(float(s) - (1.5 + float((A * A).sum()))) <= 1e-10
    # not supported: strided host C; slowest C index contracted away / carried by both operands
    big = torch.zeros(40, 60, dtype=torch.float64).pin_memory()
    with pytest.raises(Exception, match="dense row-major"):
        gpu.contract_host("ij", big[:, :50], "ik", dev(A), "kj", dev(B))
    with pytest.raises(Exception, match="slowest index"):
        gpu.contract_host("ij", torch.zeros(40, 50, dtype=torch.float64), "ij", dev(C0), "ij", dev(C0))


def test_beta_zero_never_reads_c(gpu):
    A = torch.rand(70, 50, dtype=torch.float64, device="cuda")
    B = torch.rand(50, 90, dtype=torch.float64, device="cuda")
    for fam_shape in (("ij", "ik", "kj", A, B, (70, 90)), ("i", "ik", "k", A, B[:, 0].contiguous(), (70,)), ("", "ik", "ik", A, A.clone(), ())):
        c, a, b, At, Bt, shp = fam_shape
        Ct = torch.full(shp, float("nan"), dtype=torch.float64, device="cuda")
        gpu.contract(c, Ct, a, At, b, Bt, 1.0, 0.0)
        assert not torch.isnan(Ct).any()


# ---------------------------------------------------------------------------------------------
# BLAS-like flat seam (column-major)
@pytest.mark.parametrize("ta,tb", list(itertools.product("ntc", "ntc")))
def test_dgemm_all_trans_with_padding(gpu, oracle, ta, tb):
    import ctypes as C
    rng = np.random.default_rng(8)
    m, n, k = 70, 45, 133
    lda = (m if ta == "n" else k) + 3
    ldb = (k if tb == "n" else n) + 2
    ldc = m + 5
    A = rand(rng, ((k if ta == "n" else m), lda))  # column-major: numpy row = BLAS column
    B = rand(rng, ((n if tb == "n" else k), ldb))
    Cm = rand(rng, (n, ldc))
    ref = Cm.copy()
    oracle.lib().orc_dgemm(ta.encode(), tb.encode(), m, n, k, 0.75, A.ctypes.data, lda, B.ctypes.data, ldb, -0.5, ref.ctypes.data, ldc)
    Cd = dev(Cm)
    gpu.gemm(ta, tb, m, n, k, 0.75, dev(A), lda, dev(B), ldb, -0.5, Cd, ldc)
    assert_close(Cd.cpu().numpy(), ref)
    assert np.array_equal(Cd.cpu().numpy()[:, m:], Cm[:, m:]), "padding rows of C were touched"


@pytest.mark.parametrize("ta,tb", list(itertools.product("nt", "nt")))
@pytest.mark.parametrize("shape", [(200, 150, 90), (130, 70, 1000), (64, 64, 16), (1, 2, 4), (258, 66, 34)])
def test_dgemm_tma_path(gpu, oracle, ta, tb, shape):
    """Even leading dimensions + 16-byte aligned bases take the TMA/mbarrier kernel (gemm_tma_f64): ragged M/N/K edges rely on
    TMA out-of-bounds zero fill; checked for all four layouts, with beta, and with forced split-K."""
    rng = np.random.default_rng(80)
    m, n, k = shape
    lda = (m if ta == "n" else k) + 2 + ((m if ta == "n" else k) % 2)
    ldb = (k if tb == "n" else n) + 4 + ((k if tb == "n" else n) % 2)
    ldc = m + 6 + (m % 2)
    A = rand(rng, ((k if ta == "n" else m), lda))
    B = rand(rng, ((n if tb == "n" else k), ldb))
    Cm = rand(rng, (n, ldc))
    for alpha, beta in ((1.0, 0.0), (0.75, -0.5)):
        ref = Cm.copy()
        oracle.lib().orc_dgemm(ta.encode(), tb.encode(), m, n, k, alpha, A.ctypes.data, lda, B.ctypes.data, ldb, beta, ref.ctypes.data, ldc)
        Cd = dev(Cm)
        gpu.gemm(ta, tb, m, n, k, alpha, dev(A), lda, dev(B), ldb, beta, Cd, ldc)
        assert_close(Cd.cpu().numpy(), ref)
        assert np.array_equal(Cd.cpu().numpy()[:, m:], Cm[:, m:]), "padding rows of C were touched"
    # split-K through the descriptor API on the same operands (row-major view: C[n? ...] use contract)
    At = dev(A)[:, : (m if ta == "n" else k)]
    Bt = dev(B)[:, : (k if tb == "n" else n)]
    Cd = dev(Cm)
    ref = Cm.copy()
    oracle.lib().orc_dgemm(ta.encode(), tb.encode(), m, n, k, 0.5, A.ctypes.data, lda, B.ctypes.data, ldb, 2.0, ref.ctypes.data, ldc)
    a_idx = "li" if ta == "n" else "il"   # numpy row = BLAS column
    b_idx = "jl" if tb == "n" else "lj"
    for sk in (2, 3):
        Cd = dev(Cm)
        d = gpu.contract("ji", Cd[:, :m], a_idx, At, b_idx, Bt, 0.5, 2.0, split_k=sk)
        assert_close(Cd.cpu().numpy(), ref)


def test_gemm_strided_batched_and_copy_strided(gpu, oracle):
    """eb200_dgemm_strided_batched (column-major, per-batch strides) and eb200_copy_strided (layout-changing copy used for
    host column-major -> device row-major conversion, Tensor/Backends/DeviceTensor.hpp:1054-1072 in the reference)."""
    import ctypes as C
    from einsums_b200 import _lib
    rng = np.random.default_rng(91)
    m, n, k, nb = 37, 29, 41, 5
    lda, ldb, ldc = m + 1, k + 3, m + 2
    A, B, Cm = rand(rng, (nb, k, lda)), rand(rng, (nb, n, ldb)), rand(rng, (nb, n, ldc))
    ref = Cm.copy()
    for b in range(nb):
        oracle.lib().orc_dgemm(b"n", b"n", m, n, k, 0.5, A[b].ctypes.data, lda, B[b].ctypes.data, ldb, 2.0, ref[b].ctypes.data, ldc)
    Ad, Bd, Cd = dev(A), dev(B), dev(Cm)
    one = lambda v: (C.c_double * 1)(v)
    rc = _lib.lib().eb200_dgemm_strided_batched(torch.cuda.current_stream().cuda_stream, b"n", b"n", m, n, k, one(0.5), Ad.data_ptr(), lda, k * lda,
                                               Bd.data_ptr(), ldb, n * ldb, one(2.0), Cd.data_ptr(), ldc, n * ldc, nb)
    assert rc == 0
    assert_close(Cd.cpu().numpy(), ref)
    # column-major host layout -> row-major device tensor
    H = rand(rng, (6, 7, 8))
    src = dev(np.asfortranarray(H).ravel(order="K"))          # Fortran-ordered bytes
    dst = torch.empty(6, 7, 8, dtype=torch.float64, device="cuda")
    i64 = C.c_int64 * 3
    rc = _lib.lib().eb200_copy_strided(torch.cuda.current_stream().cuda_stream, _lib.F64, 3, i64(6, 7, 8), src.data_ptr(), i64(1, 6, 42),
                                      dst.data_ptr(), i64(56, 8, 1))
    assert rc == 0
    assert np.array_equal(dst.cpu().numpy(), H)


@pytest.fixture(params=[0, 2], ids=["single", "pair"])
def tcgen05_pair_mode(request, gpu):
    """Run a tcgen05 test on the single-CTA kernel and on CTA pairs (cta_group::2, 256 x 128 tile pairs) regardless of the size rule."""
    gpu.set_option("tcgen05_pair", request.param)
    yield request.param
    gpu.set_option("tcgen05_pair", 1)


@pytest.mark.parametrize("shape", [(256, 256, 64), (300, 260, 100), (1000, 520, 260), (513, 384, 768), (128, 1024, 36)])
def test_sgemm_tcgen05_3xtf32(gpu, oracle, shape, tcgen05_pair_mode):
    """FP32 plain GEMM with both operands K-major takes the tcgen05/TMEM kernel (3xTF32 split accumulation). 1e-5 relative against a
    float64 reference; ragged M/N/K edges (TMA zero fill), alpha/beta, transposed C."""
    from einsums_b200 import KF
    m, n, k = shape
    rng = np.random.default_rng(55)
    A, B, C0 = rand(rng, (m, k), np.float32), rand(rng, (n, k), np.float32), rand(rng, (m, n), np.float32)
    for alpha, beta in ((1.0, 0.0), (0.5, 2.0)):
        ref = alpha * (A.astype(np.float64) @ B.astype(np.float64).T) + beta * C0
        Cd = dev(C0)
        d = gpu.contract("ij", Cd, "ik", dev(A), "jk", dev(B), alpha, beta)
        assert d.kernel_used == KF["TILE_TCGEN05"], d.kernel_used
        assert rel_err(Cd.cpu().numpy().astype(np.float64), ref) <= 1e-5
    Ct = dev(np.ascontiguousarray(C0.T))
    d = gpu.contract("ji", Ct, "ik", dev(A), "jk", dev(B), 1.0, 0.0)    # C stored transposed: strided epilogue
    assert d.kernel_used == KF["TILE_TCGEN05"]
    assert rel_err(Ct.cpu().numpy().astype(np.float64).T, A.astype(np.float64) @ B.astype(np.float64).T) <= 1e-5
    # long contractions: the TMEM accumulator is drained into registers every 512 k, so the truncation bias of the tensor core's
    # accumulation does not grow with K (it reached 3e-5 at K = 4096 with a single accumulator)
    for kk in (2048, 9000):
        A2, B2 = rand(rng, (256, kk), np.float32), rand(rng, (384, kk), np.float32)
        C2 = torch.zeros(256, 384, dtype=torch.float32, device="cuda")
        d = gpu.contract("ij", C2, "ik", dev(A2), "jk", dev(B2))
        assert d.kernel_used == KF["TILE_TCGEN05"]
        assert rel_err(C2.cpu().numpy().astype(np.float64), A2.astype(np.float64) @ B2.astype(np.float64).T) <= 1e-5, kk


@pytest.mark.parametrize("a_idx,b_idx", [("ik", "jk"), ("ki", "jk"), ("ik", "kj"), ("ki", "kj")])
def test_sgemm_tcgen05_layouts(gpu, a_idx, b_idx, tcgen05_pair_mode):
    """All four operand layouts of the tcgen05 kernel: K-major tiles (one TMA box, descriptor advance inside the swizzle row) and
    MN-major tiles (four 32x32 boxes, descriptor advance by 8-row atoms), incl. the FP32 Wabef pattern (both MN-major)."""
    from einsums_b200 import KF
    rng = np.random.default_rng(77)
    m, n, k = 260, 392, 200
    ext = dict(i=m, j=n, k=k)
    A = rand(rng, tuple(ext[x] for x in a_idx), np.float32)
    B = rand(rng, tuple(ext[x] for x in b_idx), np.float32)
    C0 = rand(rng, (m, n), np.float32)
    ref = 0.5 * np.einsum(f"{a_idx},{b_idx}->ij", A.astype(np.float64), B.astype(np.float64)) + 2.0 * C0
    Cd = dev(C0)
    d = gpu.contract("ij", Cd, a_idx, dev(A), b_idx, dev(B), 0.5, 2.0)
    assert d.kernel_used == KF["TILE_TCGEN05"], d.kernel_used
    assert rel_err(Cd.cpu().numpy().astype(np.float64), ref) <= 1e-5
    # CCD Wabef in FP32: abef <- mnef, mnab (both operands MN-major after index fusion)
    o, v = 6, 24
    g, t, W0 = rand(rng, (o, o, v, v), np.float32), rand(rng, (o, o, v, v), np.float32), rand(rng, (v, v, v, v), np.float32)
    refW = W0 + 0.25 * np.einsum("mnef,mnab->abef", g.astype(np.float64), t.astype(np.float64))
    Wd = dev(W0)
    d = gpu.contract("abef", Wd, "mnef", dev(g), "mnab", dev(t), 0.25, 1.0)
    assert d.kernel_used == KF["TILE_TCGEN05"], d.kernel_used
    assert rel_err(Wd.cpu().numpy().astype(np.float64), refW) <= 1e-5


def test_sgemm_tcgen05_persistent_and_batched(gpu, tcgen05_pair_mode):
    """The persistent tile loop (more tiles than SMs, so every CTA walks several tiles and the accumulators rotate) and the batched form
    (one batch index carried by all three tensors: 3-D tensor maps), against float64 matmul; 1e-5 relative (3xTF32 with truncated hi)."""
    from einsums_b200 import KF
    torch.manual_seed(11)
    # 24 x 20 = 480 tiles of 128 x 128 on 148 SMs, ragged edges, K = 3 chunks of 512 + a tail
    m, n, k = 3000, 2500, 1700
    A = torch.rand(m, k, dtype=torch.float32, device="cuda") * 2 - 1
    B = torch.rand(n, k, dtype=torch.float32, device="cuda") * 2 - 1
    C0 = torch.rand(m, n, dtype=torch.float32, device="cuda")
    Cd = C0.clone()
    d = gpu.contract("ij", Cd, "ik", A, "jk", B, 0.5, 2.0)
    assert d.kernel_used == KF["TILE_TCGEN05"], d.kernel_used
    assert gpu.last_kernel() == ("gemm_tf32x3_tcgen05<pair>" if tcgen05_pair_mode else "gemm_tf32x3_tcgen05"), gpu.last_kernel()
    ref = 0.5 * (A.double() @ B.double().T) + 2.0 * C0.double()
    assert ((Cd.double() - ref).abs().max() / ref.abs().max()).item() <= 1e-5
    # same-sign data: the truncation of hi and lo is a one-sided error; it must stay inside the tolerance
    Ap, Bp = A.abs(), B.abs()
    d = gpu.contract("ij", Cd, "ik", Ap, "jk", Bp, 1.0, 0.0)
    ref = Ap.double() @ Bp.double().T
    assert ((Cd.double() - ref).abs().max() / ref.abs().max()).item() <= 1e-5
    # batched: C[b,i,j] = sum_k A[b,i,k] B[b,k,j] (A K-major, B N-major), batch outermost; and batch in the middle of C
    nb, m, n, k = 7, 200, 264, 96
    A = torch.rand(nb, m, k, dtype=torch.float32, device="cuda") * 2 - 1
    B = torch.rand(nb, k, n, dtype=torch.float32, device="cuda") * 2 - 1
    C0 = torch.rand(nb, m, n, dtype=torch.float32, device="cuda")
    Cd = C0.clone()
    d = gpu.contract("bij", Cd, "bik", A, "bkj", B, 2.0, -1.0)
    assert d.kernel_used == KF["TILE_TCGEN05"], d.kernel_used
    ref = 2.0 * torch.bmm(A.double(), B.double()) - C0.double()
    assert ((Cd.double() - ref).abs().max() / ref.abs().max()).item() <= 1e-5
    Cm = torch.zeros(m, nb, n, dtype=torch.float32, device="cuda")
    d = gpu.contract("ibj", Cm, "bik", A, "bkj", B)
    assert d.kernel_used == KF["TILE_TCGEN05"], d.kernel_used
    ref = torch.bmm(A.double(), B.double()).permute(1, 0, 2)
    assert ((Cm.double() - ref).abs().max() / ref.abs().max()).item() <= 1e-5


@pytest.mark.parametrize("a_idx,b_idx", [("ik", "jk"), ("ki", "jk"), ("ik", "kj"), ("ki", "kj")])
def test_zgemm_tensor_pipe(gpu, oracle, a_idx, b_idx):
    """c128 plain products on the FP64 tensor pipe (gemm_tma_z: interleaved complex tiles by TMA, four real DMMA products per complex
    product): all four operand layouts, ragged M/N/K, complex alpha/beta, both conjugation flags, against the oracle's generic loops."""
    ext = dict(i=1100, j=712, k=101)
    for (alpha, beta, ca, cb) in ((1.0, 0.0, False, False), (0.5 - 0.25j, 2.0 + 1.0j, True, False), (-1.0j, 1.0, False, True), (1.0, 0.0, True, True)):
        run_case(gpu, oracle, "ij", a_idx, b_idx, ext, np.complex128, alpha, beta, ca, cb, seed=91)
        assert gpu.last_kernel() == "gemm_tma_z", gpu.last_kernel()
    run_case(gpu, oracle, "ji", a_idx, b_idx, ext, np.complex128, 0.5, 0.5, seed=92)        # C stored transposed
    assert gpu.last_kernel() == "gemm_tma_z"


def test_zgemm_seam_tensor_pipe(gpu):
    """eb200_zgemm (column-major flat seam) at a size that takes the tensor-pipe kernel: 'n', 't' and 'c' operands against numpy."""
    rng = np.random.default_rng(93)
    m, n, k = 1024, 640, 200
    z = lambda *s: (rng.uniform(-1, 1, s) + 1j * rng.uniform(-1, 1, s))
    for ta, tb in (("n", "n"), ("c", "n"), ("n", "c"), ("t", "c")):
        A = z(m, k) if ta == "n" else z(k, m)      # column-major (ld = rows): stored here as the transpose in C order
        B = z(k, n) if tb == "n" else z(n, k)
        C0 = z(m, n)
        op = lambda X, t: X if t == "n" else (X.T if t == "t" else X.conj().T)
        ref = (0.5 + 0.5j) * (op(A, ta) @ op(B, tb)) + (2.0 - 1.0j) * C0
        Ad, Bd, Cd = dev(np.asfortranarray(A).T), dev(np.asfortranarray(B).T), dev(np.asfortranarray(C0).T)
        gpu.gemm(ta, tb, m, n, k, 0.5 + 0.5j, Ad, A.shape[0], Bd, B.shape[0], 2.0 - 1.0j, Cd, m)
        assert gpu.last_kernel() == "gemm_tma_z", (ta, tb, gpu.last_kernel())
        out = Cd.cpu().numpy().T
        assert rel_err(out, ref) <= 1e-12, (ta, tb)


def test_sgemm_zgemm(gpu, oracle):
    import ctypes as C
    rng = np.random.default_rng(9)
    m, n, k = 100, 64, 50
    A, B, Cm = rand(rng, (k, m), np.float32), rand(rng, (n, k), np.float32), rand(rng, (n, m), np.float32)
    ref = Cm.copy()
    oracle.lib().orc_sgemm(b"n", b"n", m, n, k, 0.5, A.ctypes.data, m, B.ctypes.data, k, 1.5, ref.ctypes.data, m)
    Cd = dev(Cm)
    gpu.gemm("n", "n", m, n, k, 0.5, dev(A), m, dev(B), k, 1.5, Cd, m)
    assert_close(Cd.cpu().numpy(), ref, np.float32)
    A, B, Cm = rand(rng, (m, k), np.complex128), rand(rng, (n, k), np.complex128), rand(rng, (n, m), np.complex128)
    ref = Cm.copy()
    d2 = C.c_double * 2
    oracle.lib().orc_zgemm(b"c", b"n", m, n, k, d2(0.5, 0.25), A.ctypes.data, k, B.ctypes.data, k, d2(1.0, -1.0), ref.ctypes.data, m)
    Cd = dev(Cm)
    gpu.gemm("c", "n", m, n, k, 0.5 + 0.25j, dev(A), k, dev(B), k, 1.0 - 1.0j, Cd, m)
    assert_close(Cd.cpu().numpy(), ref, np.complex128)


@pytest.mark.parametrize("t", "nt")
@pytest.mark.parametrize("incx,incy", [(1, 1), (2, 3), (-1, 2)])
def test_dgemv(gpu, oracle, t, incx, incy):
    rng = np.random.default_rng(10)
    m, n, lda = 300, 170, 307
    A = rand(rng, (n, lda))
    lenx, leny = (n, m) if t == "n" else (m, n)
    x, y = rand(rng, (1 + (lenx - 1) * abs(incx),)), rand(rng, (1 + (leny - 1) * abs(incy),))
    ref = y.copy()
    oracle.lib().orc_dgemv(t.encode(), m, n, 1.5, A.ctypes.data, lda, x.ctypes.data, incx, 0.5, ref.ctypes.data, incy)
    yd = dev(y)
    gpu.gemv(t, m, n, 1.5, dev(A), lda, dev(x), incx, 0.5, yd, incy)
    assert_close(yd.cpu().numpy(), ref)


def test_level1_and_ger(gpu, oracle):
    rng = np.random.default_rng(12)
    n = 100003
    x, y, z = rand(rng, (n,)), rand(rng, (n,)), rand(rng, (n,))
    assert abs(gpu.dot(n, dev(x), 1, dev(y), 1) - oracle.lib().orc_ddot(n, x.ctypes.data, 1, y.ctypes.data, 1)) <= 1e-12 * n ** 0.5
    assert abs(gpu.dot(n // 3, dev(x), 3, dev(y), 2) - float(np.dot(x[: 3 * (n // 3) : 3], y[: 2 * (n // 3) : 2]))) <= 1e-11
    yd = dev(y); gpu.axpy(n, 0.5, dev(x), 1, yd, 1); assert_close(yd.cpu().numpy(), y + 0.5 * x)
    yd = dev(y); gpu.axpby(n, 0.5, dev(x), 1, -2.0, yd, 1); assert_close(yd.cpu().numpy(), 0.5 * x - 2.0 * y)
    yd = dev(np.full(n, np.nan)); gpu.axpby(n, 0.5, dev(x), 1, 0.0, yd, 1); assert_close(yd.cpu().numpy(), 0.5 * x)
    yd = dev(y); gpu.scal(n, -3.0, yd, 1); assert np.array_equal(yd.cpu().numpy(), -3.0 * y)
    yd = dev(y); gpu.copy(n // 2, dev(x), 2, yd, 1); assert np.array_equal(yd.cpu().numpy()[: n // 2], x[: 2 * (n // 2) : 2])
    yd = dev(y); gpu.fill(n // 2, 7.0, yd, 2); assert np.array_equal(yd.cpu().numpy()[: 2 * (n // 2) : 2], np.full(n // 2, 7.0))
    zd = dev(z); gpu.dirprod(n, 0.5, dev(x), 1, dev(y), 1, zd, 1); assert_close(zd.cpu().numpy(), z + 0.5 * x * y)
    m, k, lda = 211, 97, 215
    A = rand(rng, (k, lda)); u, w = rand(rng, (m,)), rand(rng, (k,))
    ref = A.copy(); oracle.lib().orc_dger(m, k, 0.5, u.ctypes.data, 1, w.ctypes.data, 1, ref.ctypes.data, lda)
    Ad = dev(A); gpu.ger(m, k, 0.5, dev(u), 1, dev(w), 1, Ad, lda); assert_close(Ad.cpu().numpy(), ref)
    # complex: dotc and gerc
    xc, yc = rand(rng, (1000,), np.complex128), rand(rng, (1000,), np.complex128)
    assert abs(gpu.dot(1000, dev(xc), 1, dev(yc), 1, conj=True) - np.vdot(xc, yc)) < 1e-11
    assert abs(gpu.dot(1000, dev(xc), 1, dev(yc), 1) - np.dot(xc, yc)) < 1e-11


# ---------------------------------------------------------------------------------------------
# permute
@pytest.mark.parametrize("order", list(itertools.permutations(range(4))), ids=lambda o: "".join("ijkl"[x] for x in o))
def test_permute_rank4_all_orders(gpu, oracle, order):
    """Config 4 at a size the oracle handles in seconds; unequal, non-tile-multiple extents; bit-exact for alpha=1,beta=0."""
    rng = np.random.default_rng(13)
    shape = (19, 33, 10, 45)
    A = rand(rng, shape)
    c_idx = "".join("ijkl"[o] for o in order)
    C0 = rand(rng, tuple(shape[o] for o in order))
    ref = C0.copy(); oracle.permute(c_idx, ref, "ijkl", A, 1.0, 0.0)
    Cd = dev(C0); gpu.permute(c_idx, Cd, "ijkl", dev(A), 1.0, 0.0)
    assert np.array_equal(Cd.cpu().numpy(), ref)
    ref = C0.copy(); oracle.permute(c_idx, ref, "ijkl", A, 0.5, 2.0)
    Cd = dev(C0); gpu.permute(c_idx, Cd, "ijkl", dev(A), 0.5, 2.0)
    assert_close(Cd.cpu().numpy(), ref)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
def test_permute_ranks_types_conj(gpu, oracle, dtype):
    rng = np.random.default_rng(14)
    for rank, shape in ((0, ()), (1, (37,)), (2, (65, 130)), (3, (33, 7, 64)), (5, (6, 5, 9, 4, 8))):
        letters = "ijklm"[:rank]
        A = rand(rng, shape, dtype)
        for order in list(itertools.permutations(range(rank)))[:: max(1, rank)]:
            c_idx = "".join(letters[o] for o in order)
            for conj in ((False, True) if np.dtype(dtype).kind == "c" else (False,)):
                ref = np.zeros(tuple(shape[o] for o in order), dtype); oracle.permute(c_idx, ref, letters, A, 1.0, 0.0, conj)
                Cd = dev(np.zeros_like(ref)); gpu.permute(c_idx, Cd, letters, dev(A), 1.0, 0.0, conj)
                assert np.array_equal(Cd.cpu().numpy(), ref), f"rank {rank} order {c_idx} conj {conj}"


def test_permute_into_larger_output_and_errors(gpu, oracle):
    """DevicePermute.hpp:121-131: C may be larger than A along an axis; smaller is a dimension_error."""
    rng = np.random.default_rng(15)
    A = rand(rng, (5, 6))
    C0 = rand(rng, (8, 7))
    Cd = dev(C0)
    gpu.permute("ji", Cd[:6, :5], "ij", dev(A))
    out = Cd.cpu().numpy()
    assert np.array_equal(out[:6, :5], A.T) and np.array_equal(out[6:], C0[6:]) and np.array_equal(out[:, 5:], C0[:, 5:])
    with pytest.raises(ValueError):
        gpu.permute("ji", dev(np.zeros((5, 5))), "ij", dev(A))


# ---------------------------------------------------------------------------------------------
# size-independent properties at (near) benchmark sizes, where the oracle would take too long
@pytest.fixture
def permute_tma_always(gpu):
    """Every FP64 overwrite transpose the TMA-staged kernel can legally take goes through it (ignores the tile-fill rule)."""
    gpu.set_option("permute_tma", 2)
    yield gpu
    gpu.set_option("permute_tma", 1)


@pytest.mark.parametrize("order", list(itertools.permutations(range(4))), ids=lambda o: "".join("ijkl"[x] for x in o))
def test_permute_tma_rank4_all_orders(permute_tma_always, oracle, order):
    """permute_tma_f64 (TMA-staged plane transpose) on all 24 orders: even, non-tile-multiple extents so that edge tiles are clipped by
    the TMA unit; alpha = 1 bit-exact against the oracle, alpha != 1 bit-exact against alpha * oracle (one multiplication)."""
    gpu = permute_tma_always
    rng = np.random.default_rng(113)
    shape = (18, 34, 10, 46)
    A = rand(rng, shape)
    c_idx = "".join("ijkl"[o] for o in order)
    ref = np.full(tuple(shape[o] for o in order), 7.0); oracle.permute(c_idx, ref, "ijkl", A, 1.0, 0.0)
    Cd = dev(np.full_like(ref, 7.0)); gpu.permute(c_idx, Cd, "ijkl", dev(A), 1.0, 0.0)
    assert np.array_equal(Cd.cpu().numpy(), ref)
    gpu.permute(c_idx, Cd, "ijkl", dev(A), -0.75, 0.0)
    assert np.array_equal(Cd.cpu().numpy(), -0.75 * ref)


def test_permute_tma_shapes_and_views(permute_tma_always, gpu):
    """Tile fitting (extents 160 -> 80, 96 -> 96, 300 -> 112 x 104 ...), ranks 2-5, and strided views on both sides, against torch's
    permuted view (bit-exact); the non-TMA kernels are run on the same inputs and must agree bit for bit."""
    torch.manual_seed(7)
    cases = [((160, 6, 160), (2, 1, 0)), ((96, 96), (1, 0)), ((300, 258), (1, 0)), ((6, 130, 4, 70), (3, 2, 0, 1)), ((4, 66, 2, 3, 50), (4, 2, 0, 3, 1)),
             ((2, 2, 2, 2, 40, 36), (5, 3, 1, 0, 2, 4))]
    for shape, order in cases:
        letters = "ijklmn"[:len(shape)]
        c_idx = "".join(letters[o] for o in order)
        A = torch.rand(*shape, dtype=torch.float64, device="cuda")
        B = torch.zeros([shape[o] for o in order], dtype=torch.float64, device="cuda")
        gpu.set_option("permute_tma", 2)
        gpu.permute(c_idx, B, letters, A)
        assert torch.equal(B, A.permute(*order)), (shape, order)
        B2 = torch.zeros_like(B)
        gpu.set_option("permute_tma", 0)
        gpu.permute(c_idx, B2, letters, A)
        assert torch.equal(B2, B)
    # views: A is a window of a larger tensor (even offsets/strides), C a window of a larger output
    gpu.set_option("permute_tma", 2)
    big = torch.rand(40, 70, 52, dtype=torch.float64, device="cuda")
    Av = big[4:36, 2:66, 8:48]
    out = torch.full((44, 70, 40), -1.0, dtype=torch.float64, device="cuda")
    Cv = out[2:42, 4:68, 6:38]
    gpu.permute("kji", Cv, "ijk", Av)
    assert torch.equal(Cv, Av.permute(2, 1, 0))
    assert (out[:2] == -1).all() and (out[:, :4] == -1).all() and (out[:, :, 38:] == -1).all()


def test_permute_round_trip_large(gpu):
    n = 96
    A = torch.rand(n, n, n, n, dtype=torch.float64, device="cuda")
    for order in ((3, 2, 1, 0), (1, 0, 3, 2), (2, 3, 0, 1), (0, 2, 1, 3), (3, 0, 1, 2)):
        c_idx = "".join("ijkl"[o] for o in order)
        B = torch.empty([n] * 4, dtype=torch.float64, device="cuda")
        gpu.permute(c_idx, B, "ijkl", A)
        assert torch.equal(B, A.permute(*order))           # bit-exact against torch's view
        back = torch.empty_like(A)
        gpu.permute("ijkl", back, c_idx, B)
        assert torch.equal(back, A)                         # round trip is the identity


def test_gemm_checksum_large(gpu):
    """Checksum of checksums: 1^T C 1 = (1^T A)(B 1) for C = A B, at a Wabef-like shape (M=N=6400, K=400)."""
    o, v = 20, 80
    g = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda") * 2 - 1
    t = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda") * 2 - 1
    W = torch.rand(v, v, v, v, dtype=torch.float64, device="cuda") * 2 - 1
    W0 = W.clone()
    d = gpu.contract("abef", W, "mnef", g, "mnab", t, 0.25, 1.0)
    assert d.kernel_used == 1
    lhs = (W - W0).sum(dtype=torch.float64).item()
    rhs = 0.25 * torch.dot(t.reshape(o * o, -1).sum(1), g.reshape(o * o, -1).sum(1)).item()
    assert abs(lhs - rhs) <= 1e-12 * max(1.0, (W - W0).abs().sum().item())
    # spot rows against a float64 matmul on the same device (different summation order)
    ref = W0 + 0.25 * (t.reshape(o * o, -1).T @ g.reshape(o * o, -1)).reshape(v, v, v, v)
    assert rel_err(W.cpu().numpy(), ref.cpu().numpy()) <= TOL[np.dtype(np.float64)]


def test_tma_gemm_stress(gpu):
    """Regression guard for a WAR race in the TMA/mbarrier pipeline (a shared-memory stage handed back to the producer while the
    last fragment load of the step was still in flight): ~1 corrupted 8x32 fragment per 10^6 warp-steps, only with 2 CTAs/SM.
    5000 tiles x 25 k-steps per launch, 12 launches, every element compared with a float64 matmul on the device."""
    o, v = 20, 80
    torch.manual_seed(3)
    g = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda") * 2 - 1
    t = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda") * 2 - 1
    W0 = torch.rand(v, v, v, v, dtype=torch.float64, device="cuda") * 2 - 1
    ref = W0 + 0.25 * (t.reshape(o * o, -1).T @ g.reshape(o * o, -1)).reshape(v, v, v, v)
    scale = ref.abs().max().item()
    for rep in range(12):
        W = W0.clone()
        gpu.contract("abef", W, "mnef", g, "mnab", t, 0.25, 1.0)
        err = (W - ref).abs().max().item() / scale
        assert err <= 1e-12, f"launch {rep}: relative error {err}"


def test_linearity_fock(gpu):
    n = 48
    g = torch.rand(n, n, n, n, dtype=torch.float64, device="cuda") - 0.5
    D1 = torch.rand(n, n, dtype=torch.float64, device="cuda") - 0.5
    D2 = torch.rand(n, n, dtype=torch.float64, device="cuda") - 0.5
    outs = []
    for D in (D1, D2, D1 + 2.0 * D2):
        K = torch.empty(n, n, dtype=torch.float64, device="cuda")
        gpu.contract("pq", K, "prqs", g, "rs", D, -1.0, 0.0)
        outs.append(K)
    assert rel_err((outs[0] + 2.0 * outs[1]).cpu().numpy(), outs[2].cpu().numpy()) <= 1e-12
    ref = -torch.einsum("prqs,rs->pq", g, D1)
    assert rel_err(outs[0].cpu().numpy(), ref.cpu().numpy()) <= 1e-12


def test_water_sto3g_on_device(gpu):
    """QChemTests.cpp:185-416 with every contraction on the GPU: RHF -74.942079928192, MP2 -0.049149636120."""
    def es(c, C, a, A, b, B, alpha, beta):
        Cd = dev_like(np.asarray(C))
        gpu.contract(c, Cd, a, dev_like(np.asarray(A)), b, dev_like(np.asarray(B)), alpha, beta)
        C[...] = Cd.cpu().numpy()
    e_rhf, e_mp2 = rhf_mp2_water(es)
    exp = load_water()[4]
    assert abs(e_rhf - exp["rhf_total_energy"]) < 1e-6
    assert abs(e_mp2 - exp["mp2_correlation_energy"]) < 1e-6


def test_launch_counter_and_errors(gpu):
    from einsums_b200 import Eb200Error
    gpu.launch_count_reset()
    A = torch.rand(64, 64, dtype=torch.float64, device="cuda")
    C = torch.zeros(64, 64, dtype=torch.float64, device="cuda")
    gpu.contract("ij", C, "ik", A, "kj", A)
    assert gpu.launch_count() >= 1
    with pytest.raises(Eb200Error):
        gpu.gemm("x", "n", 2, 2, 2, 1.0, A, 2, A, 2, 0.0, C, 2)


# ---------------------------------------------------------------------------------------------
# contract_tma_f64: the multi-index TMA kernel (<= 5-D tensor maps, rotating producer, fused split-K, 128x64 and 80x80 tiles)
@pytest.fixture
def tma_generic_always(gpu):
    """Route every contraction the multi-index TMA kernel can legally take through it (ignores the padding-waste rule)."""
    gpu.set_option("tma_generic", 2)
    yield gpu
    gpu.set_option("tma_generic", 1)
    gpu.set_option("tma_tile", -1)


TMA_GENERIC = [
    # Wmbej: A = t[j,n,f,b] is M-major (b), B = g[m,n,e,f] is K-major (f); two K indices that do not merge; ragged and exact extents
    ("mbej", "jnfb", "mnef", dict(m=6, b=32, e=48, j=7, n=5, f=32)),
    ("mbej", "jnfb", "mnef", dict(m=6, b=22, e=18, j=7, n=5, f=24)),
    # both operands M/N-major with a split M group (a outer, b inner) and an unmergeable K
    ("abef", "nmef", "mnab", dict(a=9, b=36, e=10, f=34, m=5, n=6)),
    # both K-major with the SAME inner contracted index (l) and two outer ones in opposite order (k, n do not merge): a 3-index K.
    # (A = [i,k,l,n] with n = 3 would have odd element strides: byte strides not multiples of 16, no tensor map -> gather kernel)
    ("ij", "iknl", "jnkl", dict(i=70, j=50, k=4, l=18, n=3)),
    # A K-major / B N-major and the reverse
    ("ij", "ik", "kj", dict(i=150, j=132, k=100)),
    ("ij", "ki", "jk", dict(i=150, j=132, k=100)),
    ("ji", "ik", "kj", dict(i=129, j=70, k=300)),
    # batch index outermost / innermost in C, carried by both operands (3-D and 4-D maps with a batch coordinate)
    ("bij", "bik", "bkj", dict(i=65, j=34, k=40, b=5)),
    ("ijb", "bik", "kbj", dict(i=66, j=34, k=48, b=3)),
    # two batch indices, target groups of two indices each
    ("xyabcd", "xyakb", "ykxcd", dict(x=2, y=3, a=5, b=16, c=4, d=20, k=40)),
]


@pytest.mark.parametrize("case", TMA_GENERIC, ids=lambda c: f"{c[0]}<-{c[1]},{c[2]}")
@pytest.mark.parametrize("tile", [0, 1])
def test_tma_generic_patterns(tma_generic_always, oracle, case, tile):
    from einsums_b200 import KF
    gpu = tma_generic_always
    gpu.set_option("tma_tile", tile)
    c_idx, a_idx, b_idx, ext = case
    for (alpha, beta, sk) in ((0.25, 1.0, 0), (1.0, 0.0, 1), (-0.5, 2.0, 3)):
        d = run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, np.float64, alpha, beta, force=KF["TILE_MMA"], seed=41, split_k=sk)
        want = "contract_tma_f64<80x80>" if tile else "contract_tma_f64<128x64>"
        if len(ext) == 3 and tile == 0 and d.split_k == 1:
            want = "gemm_tma_f64"   # a plain product on 128x64 tiles without split-K belongs to the 2-D kernel
        assert gpu.last_kernel() == want, (gpu.last_kernel(), case)
        if sk:
            assert d.split_k == sk


def test_wmnij_full_size_80x80_fused_split_k(gpu):
    """BASELINE configs[1], first einsum at full size: Wmnij = W + 0.25 sum_ef g t is a 400 x 400 x 40000 product. 128x64 tiles would pad the
    output by 43 %, so it runs on 5 x 5 tiles of 80 x 80 with the K range split over the SMs and the partials summed by the last CTA
    (contract_tma_f64<80x80>); against a float64 matmul on the device (different summation order: 1e-12) and bit-identical run to run."""
    o, v = 20, 200
    gen = torch.Generator(device="cuda").manual_seed(20251017 + 2)
    g = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    t = torch.rand(o, o, v, v, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    W0 = torch.rand(o, o, o, o, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    W = W0.clone()
    d = gpu.contract("mnij", W, "mnef", g, "ijef", t, 0.25, 1.0)
    assert gpu.last_kernel() == "contract_tma_f64<80x80>" and d.split_k > 1, (gpu.last_kernel(), d.split_k)
    ref = W0 + 0.25 * (g.reshape(o * o, -1) @ t.reshape(o * o, -1).T).reshape(o, o, o, o)
    assert ((W - ref).abs().max() / ref.abs().max()).item() <= 1e-12
    W2 = W0.clone()
    gpu.contract("mnij", W2, "mnef", g, "ijef", t, 0.25, 1.0)
    assert torch.equal(W, W2)


def test_tma_generic_split_k_is_reproducible(tma_generic_always):
    """Fused split-K: whichever CTA arrives last, the partials are added in split order, so repeated runs are bit-identical."""
    gpu = tma_generic_always
    torch.manual_seed(5)
    A = torch.rand(400, 20000, dtype=torch.float64, device="cuda") - 0.5
    B = torch.rand(400, 20000, dtype=torch.float64, device="cuda") - 0.5
    outs = []
    for rep in range(4):
        C = torch.zeros(400, 400, dtype=torch.float64, device="cuda")
        d = gpu.contract("ij", C, "ik", A, "jk", B)
        assert d.split_k > 1 and gpu.last_kernel().startswith("contract_tma_f64")
        outs.append(C)
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    ref = A @ B.T
    assert rel_err(outs[0].cpu().numpy(), ref.cpu().numpy()) <= 1e-12


def test_tma_generic_default_routing(gpu, oracle):
    """Without any knob: Wmbej-like shapes with 16-aligned inner extents and the batched product take the multi-index TMA kernel;
    the same pattern with tiny inner extents (padding would waste > 12 %) stays on the gather kernel; a plain product without
    split-K stays on gemm_tma_f64."""
    run_case(gpu, oracle, "mbej", "jnfb", "mnef", dict(m=6, b=32, e=48, j=7, n=5, f=32), np.float64, -0.5, 1.0)
    assert gpu.last_kernel() == "contract_tma_f64<128x64>"
    run_case(gpu, oracle, "mbej", "jnfb", "mnef", dict(m=6, b=21, e=19, j=7, n=5, f=23), np.float64, -0.5, 1.0)
    assert gpu.last_kernel() == "tile_mma_f64"
    run_case(gpu, oracle, "bij", "bik", "bkj", dict(i=64, j=64, k=64, b=12), np.float64, 1.0, 0.0)
    assert gpu.last_kernel() == "contract_tma_f64<80x80>"    # 4 x 4 blocks: one 80x80 tile wastes less than one 128x64 tile
    run_case(gpu, oracle, "bij", "bik", "bkj", dict(i=128, j=64, k=64, b=12), np.float64, 1.0, 0.0)
    assert gpu.last_kernel() == "contract_tma_f64<128x64>"
    run_case(gpu, oracle, "ij", "ik", "kj", dict(i=300, j=260, k=64), np.float64, 1.0, 0.0)   # 19 x 17 blocks: 80x80 tiles pad less
    assert gpu.last_kernel() == "contract_tma_f64<80x80>"
    run_case(gpu, oracle, "ij", "ik", "kj", dict(i=384, j=256, k=64), np.float64, 1.0, 0.0)
    assert gpu.last_kernel() == "gemm_tma_f64"
    run_case(gpu, oracle, "ij", "ik", "jk", dict(i=80, j=80, k=4000), np.float64, 1.0, 1.0)   # split-K: fused reduction in the generic kernel
    assert gpu.last_kernel().startswith("contract_tma_f64")


@pytest.mark.parametrize("chunk", range(3))
def test_tma_generic_fuzz(tma_generic_always, oracle, chunk):
    """Random tile-family contractions (1-2 indices per target group, 1-3 contracted, 0-1 batch; random axis orders; extents mostly even,
    16-aligned or ragged) through the multi-index TMA kernel whenever it is legal, both tile shapes, against the oracle. Cases whose
    layout has no tensor map (odd strides, no common unit-stride contracted index) fall back to the gather kernel and are checked too."""
    from einsums_b200 import KF
    gpu = tma_generic_always
    rng = np.random.default_rng(3000 + chunk)
    took = ran = 0
    for n in range(48):
        letters = list("abcdefgh")
        nm, nn, nk, nb = int(rng.integers(1, 3)), int(rng.integers(1, 3)), int(rng.integers(1, 4)), int(rng.integers(0, 2))
        if nm + nk + nb > 5 or nn + nk + nb > 5:
            continue
        m, nl, k, bt = letters[:nm], letters[nm:nm + nn], letters[nm + nn:nm + nn + nk], letters[nm + nn + nk:nm + nn + nk + nb]
        ext = {ch: int(rng.choice([2, 4, 6, 16, 20, 32, 34, 48, 3] if ch not in k[1:] + bt else [2, 3, 4, 5])) for ch in m + nl + k + bt}
        if np.prod([float(v) for v in ext.values()]) > 2e7:
            continue
        c, a, b = m + nl + bt, m + k + bt, nl + k + bt
        for lst in (c, a, b):
            rng.shuffle(lst)
        gpu.set_option("tma_tile", int(rng.integers(0, 2)))
        alpha, beta = [(1.0, 0.0), (0.5, 2.0), (-1.0, 1.0)][int(rng.integers(0, 3))]
        d = run_case(gpu, oracle, "".join(c), "".join(a), "".join(b), ext, np.float64, alpha, beta, seed=chunk * 50 + n,
                     split_k=int(rng.choice([0, 0, 2])))
        ran += 1
        if d.kernel_used == KF["TILE_MMA"] and gpu.last_kernel().startswith("contract_tma_f64"):
            took += 1
    assert took >= 3, f"only {took} of {ran} cases reached the multi-index TMA kernel"


# ---------------------------------------------------------------------------------------------
# gemm_tf32x3_tcgen05<generic>: multi-index FP32 contractions on the tcgen05 kernel (<= 5-D tensor maps, 32 x 32 boxes)
@pytest.fixture
def tcgen05_generic_always(gpu):
    """Route every FP32 contraction that has tensor maps through the tcgen05 kernel (ignores the padding and machine-fill rules)."""
    gpu.set_option("tcgen05_generic", 2)
    yield gpu
    gpu.set_option("tcgen05_generic", 1)


# the TMA_GENERIC patterns with every non-unit stride a multiple of 4 floats (FP32 tensor maps need 16-byte strides)
TCGEN05_GENERIC = [
    ("mbej", "jnfb", "mnef", dict(m=6, b=32, e=48, j=7, n=5, f=32)),
    ("mbej", "jnfb", "mnef", dict(m=6, b=24, e=20, j=7, n=5, f=28)),
    ("abef", "nmef", "mnab", dict(a=9, b=36, e=10, f=36, m=5, n=6)),
    ("ij", "iknl", "jnkl", dict(i=70, j=50, k=4, l=20, n=3)),
    ("ij", "ik", "kj", dict(i=150, j=132, k=100)),
    ("ij", "ki", "jk", dict(i=152, j=132, k=100)),
    ("ji", "ik", "kj", dict(i=129, j=72, k=300)),
    ("bij", "bik", "bkj", dict(i=65, j=36, k=40, b=5)),
    ("ijb", "bik", "kbj", dict(i=66, j=36, k=48, b=3)),
    ("xyabcd", "xyakb", "ykxcd", dict(x=2, y=3, a=5, b=16, c=4, d=20, k=40)),
]


@pytest.mark.parametrize("case", TCGEN05_GENERIC, ids=lambda c: f"{c[0]}<-{c[1]},{c[2]}")
def test_tcgen05_generic_patterns(tcgen05_generic_always, oracle, case, tcgen05_pair_mode):
    """The index patterns of the FP64 multi-index TMA kernel, in FP32, against the oracle (1e-5: 3xTF32 keeps FP32 accuracy)."""
    from einsums_b200 import KF
    gpu = tcgen05_generic_always
    c_idx, a_idx, b_idx, ext = case
    for (alpha, beta) in ((0.25, 1.0), (1.0, 0.0), (-0.5, 2.0)):
        d = run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, np.float32, alpha, beta, seed=43)
        assert d.kernel_used == KF["TILE_TCGEN05"], (d.kernel_used, gpu.last_kernel(), case)
        if len(ext) > 3:
            assert gpu.last_kernel() == ("gemm_tf32x3_tcgen05<generic,pair>" if tcgen05_pair_mode else "gemm_tf32x3_tcgen05<generic>"), (gpu.last_kernel(), case)


def test_tcgen05_generic_default_routing_and_long_k(gpu, oracle):
    """Without any knob: an FP32 Wmbej with enough 128 x 128 tiles to fill the SMs takes the tcgen05 kernel (K = 2 x 544 > 512: several
    accumulation chunks); the same pattern with a handful of tiles stays on the FMA-pipe kernel; beta = 0 must not read C."""
    from einsums_b200 import KF
    ext = dict(m=14, b=64, e=64, j=12, n=2, f=544)
    d = run_case(gpu, oracle, "mbej", "jnfb", "mnef", ext, np.float32, -0.5, 1.0, seed=7)
    assert d.kernel_used == KF["TILE_TCGEN05"] and gpu.last_kernel().startswith("gemm_tf32x3_tcgen05<generic"), gpu.last_kernel()
    d = run_case(gpu, oracle, "mbej", "jnfb", "mnef", dict(m=3, b=32, e=32, j=3, n=2, f=40), np.float32, -0.5, 1.0, seed=8)
    assert d.kernel_used == KF["TILE_SIMT"], gpu.last_kernel()
    rng = np.random.default_rng(9)
    t, g = rand(rng, (12, 2, 96, 64), np.float32), rand(rng, (14, 2, 64, 96), np.float32)
    Wd = torch.full((14, 64, 64, 12), float("nan"), dtype=torch.float32, device="cuda")
    gpu.contract("mbej", Wd, "jnfb", dev(t), "mnef", dev(g), 1.0, 0.0)
    assert gpu.last_kernel().startswith("gemm_tf32x3_tcgen05<generic")
    ref = np.zeros((14, 64, 64, 12), np.float32)
    oracle.einsum_generic("mbej", ref, "jnfb", t, "mnef", g, 1.0, 0.0)
    assert_close(Wd.cpu().numpy(), ref, np.float32)


@pytest.mark.parametrize("chunk", range(2))
def test_tcgen05_generic_fuzz(tcgen05_generic_always, oracle, chunk):
    """Random FP32 tile-family contractions (as test_tma_generic_fuzz) through the tcgen05 kernel whenever both operands have a tensor
    map; the rest falls back to the FMA-pipe kernel and is checked too."""
    from einsums_b200 import KF
    gpu = tcgen05_generic_always
    rng = np.random.default_rng(4000 + chunk)
    took = ran = 0
    for n in range(48):
        letters = list("abcdefgh")
        nm, nn, nk, nb = int(rng.integers(1, 3)), int(rng.integers(1, 3)), int(rng.integers(1, 4)), int(rng.integers(0, 2))
        if nm + nk + nb > 5 or nn + nk + nb > 5:
            continue
        m, nl, k, bt = letters[:nm], letters[nm:nm + nn], letters[nm + nn:nm + nn + nk], letters[nm + nn + nk:nm + nn + nk + nb]
        ext = {ch: int(rng.choice([4, 8, 12, 16, 20, 32, 36, 48, 72] if ch not in k[1:] + bt else [2, 3, 4, 5])) for ch in m + nl + k + bt}
        if np.prod([float(v) for v in ext.values()]) > 2e7:
            continue
        c, a, b = m + nl + bt, m + k + bt, nl + k + bt
        for lst in (c, a, b):
            rng.shuffle(lst)
        alpha, beta = [(1.0, 0.0), (0.5, 2.0), (-1.0, 1.0)][int(rng.integers(0, 3))]
        d = run_case(gpu, oracle, "".join(c), "".join(a), "".join(b), ext, np.float32, alpha, beta, seed=chunk * 50 + n)
        ran += 1
        if d.kernel_used == KF["TILE_TCGEN05"]:
            took += 1
    assert took >= 8, f"only {took} of {ran} cases reached the tcgen05 kernel"


# c64 on the tcgen05 kernel: both operands packed to real K-major form (alpha, conj and the index permutation absorbed by the pack),
# the product done by gemm_tf32x3_tcgen05 in complex mode (64 complex columns per tile, (re, im) stores)
C64_TCGEN05 = [
    ("ij", "ik", "kj", dict(i=150, j=132, k=100)),
    ("ij", "ki", "jk", dict(i=130, j=70, k=33)),                       # odd K: packed rows padded to an even length
    ("ji", "ik", "jk", dict(i=129, j=71, k=64)),                       # C transposed: the inner N index is unit-stride in C
    ("mbej", "jnfb", "mnef", dict(m=6, b=22, e=18, j=7, n=5, f=24)),   # no alignment rule: operands are packed
    ("abef", "nmef", "mnab", dict(a=9, b=36, e=10, f=34, m=5, n=6)),
    ("bij", "bik", "bkj", dict(i=65, j=34, k=40, b=5)),
    ("xyabcd", "xyakb", "ykxcd", dict(x=2, y=3, a=5, b=16, c=4, d=20, k=40)),
]


@pytest.mark.parametrize("case", C64_TCGEN05, ids=lambda c: f"{c[0]}<-{c[1]},{c[2]}")
def test_c64_tcgen05_patterns(tcgen05_generic_always, oracle, case, tcgen05_pair_mode):
    from einsums_b200 import KF
    gpu = tcgen05_generic_always
    c_idx, a_idx, b_idx, ext = case
    for (alpha, beta, ca, cb) in ((0.25, 1.0, False, False), (1.0 - 0.5j, 0.0, True, False), (-0.5 + 0.25j, 2.0 - 1.0j, False, True), (1.0, 0.5, True, True)):
        d = run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, np.complex64, alpha, beta, conj_a=ca, conj_b=cb, seed=47)
        assert d.kernel_used == KF["TILE_TCGEN05"] and gpu.last_kernel() == "gemm_tf32x3_tcgen05<c64>", (d.kernel_used, gpu.last_kernel(), case)


def test_c64_tcgen05_default_routing_and_beta_zero(gpu, oracle):
    """Without any knob a c64 product with enough tiles takes the tcgen05 kernel; a small one stays on the FMA-pipe kernel; beta = 0
    must not read C (NaN-filled); the flat seam's cgemm lands on the same kernel."""
    from einsums_b200 import KF
    d = run_case(gpu, oracle, "ij", "ik", "jk", dict(i=1280, j=640, k=96), np.complex64, 1.0 + 0.5j, 1.0, seed=3)
    assert d.kernel_used == KF["TILE_TCGEN05"] and gpu.last_kernel() == "gemm_tf32x3_tcgen05<c64>", gpu.last_kernel()
    d = run_case(gpu, oracle, "ij", "ik", "jk", dict(i=96, j=64, k=96), np.complex64, 1.0, 1.0, seed=4)
    assert d.kernel_used == KF["TILE_SIMT"]
    rng = np.random.default_rng(5)
    A, B = rand(rng, (1280, 96), np.complex64), rand(rng, (96, 640), np.complex64)
    Cd = torch.full((1280, 640), float("nan"), dtype=torch.complex64, device="cuda")
    gpu.contract("ij", Cd, "ik", dev(A), "kj", dev(B), 1.0, 0.0)
    assert gpu.last_kernel() == "gemm_tf32x3_tcgen05<c64>"
    ref = np.zeros((1280, 640), np.complex64)
    oracle.einsum_generic("ij", ref, "ik", A, "kj", B, 1.0, 0.0)
    assert_close(Cd.cpu().numpy(), ref, np.complex64)


@pytest.mark.parametrize("case", C64_TCGEN05, ids=lambda c: f"{c[0]}<-{c[1]},{c[2]}")
def test_c128_packed_tensor_pipe_patterns(tma_generic_always, oracle, case):
    """c128 contractions that are not a plain product: operands gathered to K-major matrices, gemm_tma_z per batch member into a dense
    temporary, strided accumulate into C (TTGT around the FP64 tensor-pipe kernel); conj flags, complex alpha / beta."""
    gpu = tma_generic_always
    c_idx, a_idx, b_idx, ext = case
    for (alpha, beta, ca, cb) in ((0.25, 1.0, False, False), (1.0 - 0.5j, 0.0, True, False), (-0.5 + 0.25j, 2.0 - 1.0j, False, True)):
        run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, np.complex128, alpha, beta, conj_a=ca, conj_b=cb, seed=53)
        assert gpu.last_kernel() in ("gemm_tma_z<packed>", "gemm_tma_z"), (gpu.last_kernel(), case)


def test_c128_packed_default_routing(gpu, oracle):
    """A c128 Wmbej with more tiles than SMs takes the packed tensor-pipe path without any knob; beta = 0 never reads C."""
    ext = dict(m=20, b=64, e=48, j=10, n=2, f=40)
    run_case(gpu, oracle, "mbej", "jnfb", "mnef", ext, np.complex128, -0.5, 1.0, seed=11)
    assert gpu.last_kernel() == "gemm_tma_z<packed>", gpu.last_kernel()
    rng = np.random.default_rng(12)
    t, g = rand(rng, (10, 2, 40, 64), np.complex128), rand(rng, (20, 2, 48, 40), np.complex128)
    Wd = torch.full((20, 64, 48, 10), float("nan"), dtype=torch.complex128, device="cuda")
    gpu.contract("mbej", Wd, "jnfb", dev(t), "mnef", dev(g), 1.0, 0.0)
    assert gpu.last_kernel() == "gemm_tma_z<packed>"
    ref = np.zeros((20, 64, 48, 10), np.complex128)
    oracle.einsum_generic("mbej", ref, "jnfb", t, "mnef", g, 1.0, 0.0)
    assert_close(Wd.cpu().numpy(), ref, np.complex128)


@pytest.mark.parametrize("case", TMA_GENERIC, ids=lambda c: f"{c[0]}<-{c[1]},{c[2]}")
def test_f32_packed_tcgen05_patterns(tcgen05_generic_always, oracle, case, tcgen05_pair_mode):
    """FP32 contractions whose own layout has no tensor map (strides that are not multiples of 16 bytes) are packed to K-major
    matrices and run on the tcgen05 kernel as well; the aligned ones take it directly."""
    from einsums_b200 import KF
    gpu = tcgen05_generic_always
    c_idx, a_idx, b_idx, ext = case
    for (alpha, beta) in ((0.25, 1.0), (1.0, 0.0)):
        d = run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, np.float32, alpha, beta, seed=59)
        assert d.kernel_used == KF["TILE_TCGEN05"], (d.kernel_used, gpu.last_kernel(), case)


def test_f32_packed_default_routing(gpu, oracle):
    """Odd extents (no 16-byte strides) at a size that fills the SMs: packed path without any knob."""
    d = run_case(gpu, oracle, "mbej", "jnfb", "mnef", dict(m=13, b=67, e=61, j=11, n=3, f=45), np.float32, -0.5, 1.0, seed=13)
    assert gpu.last_kernel() == "gemm_tf32x3_tcgen05<packed>", gpu.last_kernel()


def test_f64_packed_tma_generic(tma_generic_always, oracle):
    """FP64 contractions without a tensor map (odd strides) or with a badly padding index order: operands packed to dense K-major form,
    multi-index TMA kernel on the copies (C written in place through its own strides); against the oracle, incl. batch indices and split-K."""
    gpu = tma_generic_always
    cases = [("mbej", "jnfb", "mnef", dict(m=6, b=21, e=19, j=7, n=5, f=23)),
             ("ij", "iknl", "jnkl", dict(i=70, j=50, k=4, l=19, n=3)),
             ("abef", "nmef", "mnab", dict(a=9, b=35, e=10, f=33, m=5, n=3)),
             ("bij", "bik", "bkj", dict(i=65, j=33, k=41, b=5)),
             ("xyabcd", "xyakb", "ykxcd", dict(x=2, y=3, a=5, b=15, c=3, d=21, k=37))]
    for c_idx, a_idx, b_idx, ext in cases:
        for (alpha, beta) in ((0.25, 1.0), (1.0, 0.0)):
            run_case(gpu, oracle, c_idx, a_idx, b_idx, ext, np.float64, alpha, beta, seed=61)
            assert gpu.last_kernel() == "contract_tma_f64<packed>", (gpu.last_kernel(), c_idx, a_idx, b_idx)
    gpu.set_option("tma_generic", 1)
    # default routing: large odd-extent Wmbej takes the packed path, and agrees with the gather kernel
    rng = np.random.default_rng(62)
    o, v = 13, 79
    t, g, W0 = rand(rng, (o, o, v, v)), rand(rng, (o, o, v, v)), rand(rng, (o, v, v, o))
    Wp, Wg = dev(W0), dev(W0)
    gpu.contract("mbej", Wp, "jnfb", dev(t), "mnef", dev(g), -0.5, 1.0)
    assert gpu.last_kernel() == "contract_tma_f64<packed>", gpu.last_kernel()
    gpu.set_option("tma_generic", 0)
    gpu.contract("mbej", Wg, "jnfb", dev(t), "mnef", dev(g), -0.5, 1.0)
    assert gpu.last_kernel() == "tile_mma_f64", gpu.last_kernel()
    ref = W0 - 0.5 * np.einsum("jnfb,mnef->mbej", t, g)
    assert rel_err(Wp.cpu().numpy(), ref) <= 1e-12 and rel_err(Wg.cpu().numpy(), ref) <= 1e-12


@pytest.mark.parametrize("dtype", [torch.float32, torch.complex64, torch.float64, torch.complex128])
def test_default_routing_fuzz_mid_sizes(gpu, dtype):
    """Random tile-family contractions at sizes where the default rules pick the tensor-core routes (direct maps, packed operands, CTA
    pairs, TTGT for c128) — no knobs set. Too large for the CPU oracle's loops, so the checker is torch.einsum in the next wider
    precision on the device (FP32 / c64: 1e-5, FP64 / c128: 1e-12 relative to the result's largest magnitude)."""
    wide = {torch.float32: torch.float64, torch.complex64: torch.complex128, torch.float64: torch.float64, torch.complex128: torch.complex128}[dtype]
    tol = 1e-5 if dtype in (torch.float32, torch.complex64) else 1e-12
    rng = np.random.default_rng({torch.float32: 1, torch.complex64: 2, torch.float64: 3, torch.complex128: 4}[dtype])
    gen = torch.Generator(device="cuda").manual_seed(int(rng.integers(1 << 30)))
    def rnd(shape):
        rd = {torch.complex64: torch.float32, torch.complex128: torch.float64}.get(dtype, dtype)
        x = torch.rand(*shape, dtype=rd, device="cuda", generator=gen) * 2 - 1
        if dtype.is_complex:
            x = torch.complex(x, torch.rand(*shape, dtype=rd, device="cuda", generator=gen) * 2 - 1)
        return x
    seen = set()
    for n in range(14):
        nm, nn, nk, nb = int(rng.integers(1, 3)), int(rng.integers(1, 3)), int(rng.integers(1, 3)), int(rng.integers(0, 2))
        letters = list("abcdefgh")
        m, nl, k, bt = letters[:nm], letters[nm:nm + nn], letters[nm + nn:nm + nn + nk], letters[nm + nn + nk:nm + nn + nk + nb]
        def split(total, parts):   # extents whose product is about `total`, odd and even mixed
            if parts == 1:
                return [int(total)]
            a = int(rng.choice([3, 4, 5, 7, 8, 12, 16]))
            return [a, max(2, int(total // a))]
        ext = {}
        for grp, total in ((m, rng.choice([640, 900, 1333, 2048])), (nl, rng.choice([512, 777, 1100, 1536])), (k, rng.choice([64, 150, 333, 640]))):
            for ch, e in zip(grp, split(total, len(grp))):
                ext[ch] = e
        for ch in bt:
            ext[ch] = int(rng.choice([2, 3]))
        c, a, b = m + nl + bt, m + k + bt, nl + k + bt
        for lst in (c, a, b):
            rng.shuffle(lst)
        c_idx, a_idx, b_idx = "".join(c), "".join(a), "".join(b)
        alpha, beta = [(1.0, 0.0), (0.5, 2.0), (-1.0, 1.0)][int(rng.integers(0, 3))]
        A, B, C0 = rnd([ext[x] for x in a_idx]), rnd([ext[x] for x in b_idx]), rnd([ext[x] for x in c_idx])
        Cd = C0.clone()
        gpu.contract(c_idx, Cd, a_idx, A, b_idx, B, alpha, beta)
        seen.add(gpu.last_kernel())
        ref = beta * C0.to(wide) + alpha * torch.einsum(f"{a_idx},{b_idx}->{c_idx}", A.to(wide), B.to(wide))
        err = ((Cd.to(wide) - ref).abs().max() / ref.abs().max()).item()
        assert err <= tol, (c_idx, a_idx, b_idx, ext, gpu.last_kernel(), err)
    print("kernels seen:", sorted(seen))
    assert any(("tcgen05" in s_ or "tma" in s_) for s_ in seen), seen


def test_sgemm_tcgen05_pair_stress_against_single(gpu):
    """CTA pairs hand tiles between two SMs (remote mbarrier arrivals, multicast commits): the same products on the single-CTA kernel and on
    pairs must agree BITWISE (same MMA sequence per 128 x 128 accumulator, same chunking), every element, over repeated launches and several
    tile / k-step counts — a stale `lo` tile or an accumulator handed over early would show up here long before it shows up in a tolerance."""
    torch.manual_seed(23)
    for (m, n, k, a_idx, b_idx) in ((2048, 2048, 1024, "ik", "jk"), (4096, 1536, 4096 + 96, "ik", "kj"), (5000, 1100, 300, "ki", "jk")):
        ext = dict(i=m, j=n, k=k)
        A = torch.rand(*[ext[c] for c in a_idx], dtype=torch.float32, device="cuda") * 2 - 1
        B = torch.rand(*[ext[c] for c in b_idx], dtype=torch.float32, device="cuda") * 2 - 1
        C0 = torch.rand(m, n, dtype=torch.float32, device="cuda")
        gpu.set_option("tcgen05_pair", 0)
        Cs = C0.clone()
        gpu.contract("ij", Cs, a_idx, A, b_idx, B, 0.5, 2.0)
        assert gpu.last_kernel() == "gemm_tf32x3_tcgen05"
        gpu.set_option("tcgen05_pair", 2)
        try:
            for rep in range(6):
                Cp = C0.clone()
                gpu.contract("ij", Cp, a_idx, A, b_idx, B, 0.5, 2.0)
                assert gpu.last_kernel() == "gemm_tf32x3_tcgen05<pair>"
                assert torch.equal(Cp, Cs), (m, n, k, rep, (Cp - Cs).abs().max().item())
        finally:
            gpu.set_option("tcgen05_pair", 1)

// GPU tests of the C++20 host shim: the reference's own device test cases (GPUTensorAlgebra.hip, GPULinearAlgebra.hip,
// Sort.cpp / Transpose.cpp / Hadamard.cpp / Dot.cpp twins) rewritten against DeviceTensor + einsum/permute/linear_algebra of
// this backend. Expected values come from hand loops on host copies, as the reference tests do; tolerances: 1e-12
// relative (FP64), exact where the reference asserts exact integers.
#include <Einsums/LinearAlgebra.hpp>
#include <Einsums/hipBLAS.hpp>
#include <Einsums/Tensor/BlockTensor.hpp>
#include <Einsums/Tensor/DeviceTensor.hpp>
#include <Einsums/Tensor/TiledTensor.hpp>
#include <Einsums/Tensor/BlockTileAlgebra.hpp>
#include <Einsums/TensorAlgebra.hpp>

#include "mini_test.hpp"

#include <complex>
#include <thread>

using namespace einsums;
using namespace einsums::tensor_algebra;
using einsums::tensor_algebra::detail::AlgorithmChoice;
namespace ta = einsums::tensor_algebra::detail;

// ---- small host mirror helpers ---------------------------------------------------------------------------------------
template <typename T, size_t R> struct Host {
    std::array<int64_t, R> d{};
    std::vector<T>         v;
    template <typename TensorT> explicit Host(TensorT const &t) : v(t.to_host()) {
        for (size_t i = 0; i < R; i++) d[i] = t.dim(i);
    }
    template <typename... I> T &operator()(I... i) {
        std::array<int64_t, R> ix{int64_t(i)...};
        int64_t                off = 0;
        for (size_t k = 0; k < R; k++) off = off * d[k] + ix[k];
        return v[off];
    }
};
template <typename TensorT> auto host(TensorT const &t) { return Host<typename TensorT::ValueType, TensorT::Rank>(t); }

// ---------------------------------------------------------------------------------------------------------------------
TEST_CASE(gpu_einsum1_known_answers) { // GPUTensorAlgebra.hip:445-518
    using namespace einsums::index;
    DeviceTensor<double, 2> A{"A", einsums::detail::DEV_ONLY, 3, 3}, B{"B", einsums::detail::DEV_ONLY, 3, 3}, C{"C", einsums::detail::DEV_ONLY, 3, 3};
    for (int i0 = 0, ij = 1; i0 < 3; i0++)
        for (int j0 = 0; j0 < 3; j0++, ij++) {
            A(i0, j0) = ij;
            B(i0, j0) = ij;
        }
    AlgorithmChoice alg;
    einsum(Indices{i, j}, &C, Indices{i, k}, A, Indices{k, j}, B, &alg);
    CHECK(alg == ta::GEMM);
    double want[9] = {30, 36, 42, 66, 81, 96, 102, 126, 150};
    for (int x = 0; x < 9; x++) CHECK((double)C(x / 3, x % 3) == want[x]);

    DeviceTensor<double, 3> A3{"A", 3, 3, 3}, B3{"B", 3, 3, 3};
    std::vector<double>     inc(27);
    for (int x = 0; x < 27; x++) inc[x] = x + 1;
    A3 = inc;
    B3 = inc;
    einsum(Indices{i, l}, &C, Indices{i, j, k}, A3, Indices{j, k, l}, B3, &alg);
    CHECK(alg == ta::GEMM);
    double want2[9] = {765, 810, 855, 1818, 1944, 2070, 2871, 3078, 3285};
    for (int x = 0; x < 9; x++) CHECK((double)C(x / 3, x % 3) == want2[x]);
}

TEST_CASE(gpu_wmnij_and_wabef) { // GPUTensorAlgebra.hip:1636-1672 (Wmnij) + the Wabef pattern of config 2
    using namespace einsums::index;
    size_t _m = 12, _n = 12, _i = 5, _j = 5, _e = 7, _f = 7;
    auto   W = create_random_gpu_tensor("Wmnij", _m, _n, _i, _j);
    auto   t_oovv = create_random_gpu_tensor("t_oovv", _i, _j, _e, _f), g_oovv = create_random_gpu_tensor("g_oovv", _m, _n, _e, _f);
    auto   hW = host(W), ht = host(t_oovv), hg = host(g_oovv);
    AlgorithmChoice alg;
    einsum(1.0, Indices{m, n, i, j}, &W, 0.25, Indices{i, j, e, f}, t_oovv, Indices{m, n, e, f}, g_oovv, &alg);
    CHECK(alg == ta::GEMM);
    auto out = host(W);
    for (size_t m0 = 0; m0 < _m; m0++)
        for (size_t n0 = 0; n0 < _n; n0++)
            for (size_t i0 = 0; i0 < _i; i0++)
                for (size_t j0 = 0; j0 < _j; j0++) {
                    double s = hW(m0, n0, i0, j0);
                    for (size_t e0 = 0; e0 < _e; e0++)
                        for (size_t f0 = 0; f0 < _f; f0++) s += 0.25 * ht(i0, j0, e0, f0) * hg(m0, n0, e0, f0);
                    CHECK_CLOSE(out(m0, n0, i0, j0), s, 1e-12);
                }
    // Wabef[a,b,e,f] += 0.25 sum_mn g[m,n,e,f] t[m,n,a,b]
    size_t o = 4, v = 9;
    auto   Wv = create_random_gpu_tensor("Wabef", v, v, v, v), g2 = create_random_gpu_tensor("g", o, o, v, v), t2 = create_random_gpu_tensor("t", o, o, v, v);
    auto   h0 = host(Wv), hg2 = host(g2), ht2 = host(t2);
    einsum(1.0, Indices{a, b, e, f}, &Wv, 0.25, Indices{m, n, e, f}, g2, Indices{m, n, a, b}, t2, &alg);
    CHECK(alg == ta::GEMM);
    auto h1 = host(Wv);
    for (size_t a0 = 0; a0 < v; a0++)
        for (size_t b0 = 0; b0 < v; b0++)
            for (size_t e0 = 0; e0 < v; e0++)
                for (size_t f0 = 0; f0 < v; f0++) {
                    double s = h0(a0, b0, e0, f0);
                    for (size_t m0 = 0; m0 < o; m0++)
                        for (size_t n0 = 0; n0 < o; n0++) s += 0.25 * hg2(m0, n0, e0, f0) * ht2(m0, n0, a0, b0);
                    CHECK_CLOSE(h1(a0, b0, e0, f0), s, 1e-12);
                }
}

TEST_CASE(gpu_fock_j_k_and_wmbej) { // config 1 (GEMV swapped / GENERIC) and config 3 (generic Wmbej)
    using namespace einsums::index;
    size_t nb = 9;
    auto   g = create_random_gpu_tensor("g", nb, nb, nb, nb);
    auto   D = create_random_gpu_tensor("D", nb, nb);
    DeviceTensor<double, 2> J{"J", nb, nb}, K{"K", nb, nb};
    AlgorithmChoice algJ, algK;
    einsum(0.0, Indices{p, q}, &J, 2.0, Indices{p, q, r, s}, g, Indices{r, s}, D, &algJ);
    einsum(0.0, Indices{p, q}, &K, -1.0, Indices{p, r, q, s}, g, Indices{r, s}, D, &algK);
    CHECK(algJ == ta::GEMV);
    CHECK(algK == ta::GENERIC);
    auto hg = host(g);
    auto hD = host(D), hJ = host(J), hK = host(K);
    for (size_t p0 = 0; p0 < nb; p0++)
        for (size_t q0 = 0; q0 < nb; q0++) {
            double sj = 0, sk = 0;
            for (size_t r0 = 0; r0 < nb; r0++)
                for (size_t s0 = 0; s0 < nb; s0++) {
                    sj += 2.0 * hg(p0, q0, r0, s0) * hD(r0, s0);
                    sk -= hg(p0, r0, q0, s0) * hD(r0, s0);
                }
            CHECK_CLOSE(hJ(p0, q0), sj, 1e-12);
            CHECK_CLOSE(hK(p0, q0), sk, 1e-12);
        }
    // swapped-operand form used by QChemTests.cpp:301-302
    DeviceTensor<double, 2> F{"F", nb, nb};
    F.zero();
    einsum(1.0, Indices{i, j}, &F, 2.0, Indices{k, l}, D, Indices{i, j, k, l}, g, &algJ);
    CHECK(algJ == ta::GEMV);
    auto hF = host(F);
    for (size_t p0 = 0; p0 < nb; p0++)
        for (size_t q0 = 0; q0 < nb; q0++) CHECK_CLOSE(hF(p0, q0), hJ(p0, q0), 1e-12);

    size_t o = 5, v = 8;
    auto   t = create_random_gpu_tensor("t", o, o, v, v), gg = create_random_gpu_tensor("g", o, o, v, v), W = create_random_gpu_tensor("W", o, v, v, o);
    auto   h0 = host(W), ht = host(t), hgg = host(gg);
    AlgorithmChoice alg;
    einsum(1.0, Indices{m, b, e, j}, &W, -0.5, Indices{j, n, f, b}, t, Indices{m, n, e, f}, gg, &alg);
    CHECK(alg == ta::GENERIC);
    auto h1 = host(W);
    for (size_t m0 = 0; m0 < o; m0++)
        for (size_t b0 = 0; b0 < v; b0++)
            for (size_t e0 = 0; e0 < v; e0++)
                for (size_t j0 = 0; j0 < o; j0++) {
                    double s = h0(m0, b0, e0, j0);
                    for (size_t n0 = 0; n0 < o; n0++)
                        for (size_t f0 = 0; f0 < v; f0++) s -= 0.5 * ht(j0, n0, f0, b0) * hgg(m0, n0, e0, f0);
                    CHECK_CLOSE(h1(m0, b0, e0, j0), s, 1e-12);
                }
}

TEST_CASE(gpu_transposed_c_variants) { // Transpose.cpp:17-125, GPUTensorAlgebra.hip:1527-1634
    using namespace einsums::index;
    size_t ni = 6, nj = 7, nk = 5;
    auto   A = create_random_gpu_tensor("A", ni, nk), At = create_random_gpu_tensor("At", nk, ni);
    auto   B = create_random_gpu_tensor("B", nk, nj), Bt = create_random_gpu_tensor("Bt", nj, nk);
    auto   hA = host(A), hAt = host(At), hB = host(B), hBt = host(Bt);
    AlgorithmChoice alg;
    auto check = [&](auto &C, bool transposed_c, auto &&aval, auto &&bval) {
        auto hC = host(C);
        for (size_t i0 = 0; i0 < ni; i0++)
            for (size_t j0 = 0; j0 < nj; j0++) {
                double s = 0;
                for (size_t k0 = 0; k0 < nk; k0++) s += aval(i0, k0) * bval(k0, j0);
                CHECK_CLOSE(transposed_c ? hC(j0, i0) : hC(i0, j0), s, 1e-12);
            }
    };
    DeviceTensor<double, 2> C{"C", ni, nj}, Ct{"Ct", nj, ni};
    einsum(Indices{i, j}, &C, Indices{i, k}, A, Indices{k, j}, B, &alg);   CHECK(alg == ta::GEMM);
    check(C, false, [&](size_t x, size_t y) { return hA(x, y); }, [&](size_t x, size_t y) { return hB(x, y); });
    einsum(Indices{i, j}, &C, Indices{k, i}, At, Indices{k, j}, B, &alg);  CHECK(alg == ta::GEMM);
    check(C, false, [&](size_t x, size_t y) { return hAt(y, x); }, [&](size_t x, size_t y) { return hB(x, y); });
    einsum(Indices{i, j}, &C, Indices{i, k}, A, Indices{j, k}, Bt, &alg);  CHECK(alg == ta::GEMM);
    check(C, false, [&](size_t x, size_t y) { return hA(x, y); }, [&](size_t x, size_t y) { return hBt(y, x); });
    einsum(Indices{i, j}, &C, Indices{k, i}, At, Indices{j, k}, Bt, &alg); CHECK(alg == ta::GEMM);
    check(C, false, [&](size_t x, size_t y) { return hAt(y, x); }, [&](size_t x, size_t y) { return hBt(y, x); });
    einsum(Indices{j, i}, &Ct, Indices{i, k}, A, Indices{k, j}, B, &alg);  CHECK(alg == ta::GEMM);
    check(Ct, true, [&](size_t x, size_t y) { return hA(x, y); }, [&](size_t x, size_t y) { return hB(x, y); });
    einsum(Indices{j, i}, &Ct, Indices{k, i}, At, Indices{j, k}, Bt, &alg); CHECK(alg == ta::GEMM);
    check(Ct, true, [&](size_t x, size_t y) { return hAt(y, x); }, [&](size_t x, size_t y) { return hBt(y, x); });
}

TEST_CASE(gpu_views_runtime_decision) { // Einsum7.cpp:172-210 and Einsum1.cpp:96-140 on device tensors
    using namespace einsums::index;
    size_t naux = 9, nocc = 3, nvirt = 4, ur = 5, nmo = nocc + nvirt;
    auto   Bq = create_random_gpu_tensor("Q", naux, nmo, nmo);
    auto   iaX = create_random_gpu_tensor("ia_X", nocc, nvirt, ur);
    DeviceTensor<double, 2> N1{"N", naux, ur}, N2{"N", naux, ur};
    auto   Qov = Bq(All, Range{0, nocc}, Range{nocc, nmo});
    AlgorithmChoice alg;
    einsum(Indices{Q, X}, &N1, Indices{Q, i, a}, Qov, Indices{i, a, X}, iaX, &alg);
    CHECK(alg == ta::GENERIC); // strided view: GEMM pattern falls back at run time
    DeviceTensor<double, 3> dense(Qov);
    einsum(Indices{Q, X}, &N2, Indices{Q, i, a}, dense, Indices{i, a, X}, iaX, &alg);
    CHECK(alg == ta::GEMM);
    auto h1 = host(N1), h2 = host(N2);
    for (size_t x = 0; x < naux * ur; x++) CHECK_CLOSE(h1.v[x], h2.v[x], 1e-12);

    auto A = create_random_gpu_tensor("A", 3, 5), B = create_random_gpu_tensor("B", 3, 5);
    DeviceTensor<double, 2> C{"C", 10, 10};
    C.zero();
    DeviceTensorView Av{A, Dim<2>{3, 3}, Offset<2>{0, 0}}, Bv{B, Dim<2>{3, 3}, Offset<2>{0, 2}}, Cv{C, Dim<2>{3, 3}, Offset<2>{5, 5}};
    einsum(Indices{i, j}, &Cv, Indices{i, k}, Av, Indices{j, k}, Bv, &alg);
    CHECK(alg == ta::GEMM);
    auto hA = host(A), hB = host(B), hC = host(C);
    for (size_t x = 0; x < 10; x++)
        for (size_t y = 0; y < 10; y++) {
            double s = 0;
            if (x >= 5 && x < 8 && y >= 5 && y < 8)
                for (size_t k0 = 0; k0 < 3; k0++) s += hA(x - 5, k0) * hB(y - 5, k0 + 2);
            CHECK_CLOSE(hC(x, y), s, 1e-12);
        }
}

TEST_CASE(gpu_dot_direct_outer_hadamard) { // DotGPU.hip, Element.cpp, OuterProduct.cpp, Hadamard.cpp
    using namespace einsums::index;
    auto A = create_random_gpu_tensor("A", 4, 5, 6), B = create_random_gpu_tensor("B", 4, 5, 6);
    auto hA = host(A), hB = host(B);
    DeviceTensor<double, 0> d0;
    AlgorithmChoice         alg;
    einsum(Indices{}, &d0, Indices{i, j, a}, A, Indices{i, j, a}, B, &alg);
    CHECK(alg == ta::DOT);
    double want = 0;
    for (size_t x = 0; x < hA.v.size(); x++) want += hA.v[x] * hB.v[x];
    CHECK_CLOSE((double)d0, want, 1e-12);
    double hs = 3.0; // host scalar C
    einsum(2.0, Indices{}, &hs, 0.5, Indices{i, j, a}, A, Indices{i, j, a}, B, &alg);
    CHECK(alg == ta::DOT);
    CHECK_CLOSE(hs, 6.0 + 0.5 * want, 1e-12);
    CHECK_CLOSE(linear_algebra::dot(A, B), want, 1e-12);

    auto C = create_random_gpu_tensor("C", 4, 5, 6);
    auto h0 = host(C);
    einsum(2.0, Indices{i, j, a}, &C, -1.0, Indices{i, j, a}, A, Indices{i, j, a}, B, &alg);
    CHECK(alg == ta::DIRECT);
    auto h1 = host(C);
    for (size_t x = 0; x < h1.v.size(); x++) CHECK_CLOSE(h1.v[x], 2.0 * h0.v[x] - hA.v[x] * hB.v[x], 1e-12);

    auto u = create_random_gpu_tensor("u", 7), w = create_random_gpu_tensor("w", 9);
    DeviceTensor<double, 2> O{"O", 7, 9}, Ot{"Ot", 9, 7};
    einsum(Indices{a, b}, &O, Indices{a}, u, Indices{b}, w, &alg);  CHECK(alg == ta::GER);
    einsum(Indices{b, a}, &Ot, Indices{a}, u, Indices{b}, w, &alg); CHECK(alg == ta::GER);
    auto hu = host(u), hw = host(w);
    auto hO = host(O), hOt = host(Ot);
    for (size_t x = 0; x < 7; x++)
        for (size_t y = 0; y < 9; y++) {
            CHECK_CLOSE(hO(x, y), hu(x) * hw(y), 1e-14);
            CHECK_CLOSE(hOt(y, x), hu(x) * hw(y), 1e-14);
        }
    // rank-2 x rank-2 outer product into rank 4 (OuterProduct.cpp)
    auto P = create_random_gpu_tensor("P", 3, 4), Qm = create_random_gpu_tensor("Q", 5, 2);
    DeviceTensor<double, 4> R4{"R", 3, 4, 5, 2};
    einsum(Indices{i, j, k, l}, &R4, Indices{i, j}, P, Indices{k, l}, Qm, &alg);
    CHECK(alg == ta::GER);
    auto hP = host(P), hQ = host(Qm);
    auto hR = host(R4);
    for (size_t x = 0; x < 3; x++) for (size_t y = 0; y < 4; y++) for (size_t z = 0; z < 5; z++) for (size_t t = 0; t < 2; t++)
        CHECK_CLOSE(hR(x, y, z, t), hP(x, y) * hQ(z, t), 1e-14);

    // Hadamard: C[i] = sum_j A[i,i,j]... keep it simple: i <- ii, i  and  ij <- ij... with repeated index in A (Hadamard.cpp:17-60)
    auto H = create_random_gpu_tensor("H", 6, 6);
    auto hv = create_random_gpu_tensor("h", 6);
    DeviceTensor<double, 1> Hd{"Hd", 6};
    einsum(Indices{i}, &Hd, Indices{i, i}, H, Indices{i}, hv, &alg);
    CHECK(alg == ta::GENERIC);
    auto hH = host(H);
    auto hhv = host(hv), hHd = host(Hd);
    for (size_t x = 0; x < 6; x++) CHECK_CLOSE(hHd(x), hH(x, x) * hhv(x), 1e-14);
    // index only in C: broadcast (GPUTensorAlgebra.hip:962-995)
    auto A3 = create_random_gpu_tensor("A3", 3, 3, 4), B3 = create_random_gpu_tensor("B3", 3, 3, 4);
    DeviceTensor<double, 2> Cb{"Cb", 3, 2};
    einsum(Indices{i, l}, &Cb, Indices{i, j, k}, A3, Indices{j, i, k}, B3, &alg);
    CHECK(alg == ta::GENERIC);
    auto hA3 = host(A3), hB3 = host(B3);
    auto hCb = host(Cb);
    for (size_t x = 0; x < 3; x++) {
        double s = 0;
        for (size_t y = 0; y < 3; y++) for (size_t z = 0; z < 4; z++) s += hA3(x, y, z) * hB3(y, x, z);
        for (size_t l0 = 0; l0 < 2; l0++) CHECK_CLOSE(hCb(x, l0), s, 1e-12);
    }
}

TEST_CASE(gpu_complex_and_float) {
    using namespace einsums::index;
    using Z = std::complex<double>;
    auto A = create_random_gpu_tensor<Z>("A", 5, 6), B = create_random_gpu_tensor<Z>("B", 6, 4);
    DeviceTensor<Z, 2> C{"C", 5, 4};
    AlgorithmChoice    alg;
    einsum(Indices{i, j}, &C, Indices{i, k}, A, Indices{k, j}, B, &alg);
    CHECK(alg == ta::GEMM);
    auto hA = host(A), hB = host(B), hC = host(C);
    for (size_t x = 0; x < 5; x++) for (size_t y = 0; y < 4; y++) {
        Z s = 0;
        for (size_t k0 = 0; k0 < 6; k0++) s += hA(x, k0) * hB(k0, y);
        CHECK_CLOSE(hC(x, y), s, 1e-12);
    }
    einsum<true, false>(Indices{i, j}, &C, Indices{i, k}, A, Indices{k, j}, B, &alg); // conj(A): result must carry the conjugation
    CHECK(alg == ta::GEMM);
    hC = host(C);
    for (size_t x = 0; x < 5; x++) for (size_t y = 0; y < 4; y++) {
        Z s = 0;
        for (size_t k0 = 0; k0 < 6; k0++) s += std::conj(hA(x, k0)) * hB(k0, y);
        CHECK_CLOSE(hC(x, y), s, 1e-12);
    }
    auto Af = create_random_gpu_tensor<float>("Af", 33, 17), Bf = create_random_gpu_tensor<float>("Bf", 17, 29);
    DeviceTensor<float, 2> Cf{"Cf", 33, 29};
    einsum(Indices{i, j}, &Cf, Indices{i, k}, Af, Indices{k, j}, Bf, &alg);
    CHECK(alg == ta::GEMM);
    auto hAf = host(Af), hBf = host(Bf), hCf = host(Cf);
    for (size_t x = 0; x < 33; x++) for (size_t y = 0; y < 29; y++) {
        double s = 0;
        for (size_t k0 = 0; k0 < 17; k0++) s += double(hAf(x, k0)) * double(hBf(k0, y));
        CHECK_CLOSE(hCf(x, y), s, 1e-5 * 17);
    }
}

TEST_CASE(gpu_other_types_reach_tensor_cores) {
    // The same einsum() call, sizes that fill the SMs: FP32 and c64 Wmbej land on the tcgen05 kernel (generic instantiation / packed
    // operands), c128 on the FP64 tensor pipe (TTGT around gemm_tma_z) — eb200_last_kernel_name() says which; spot rows against host loops.
    using namespace einsums::index;
    const size_t o = 12, v = 64;
    auto check = [&](auto zero, double tol, const char *want_prefix) {
        using T = decltype(zero);
        auto t = create_random_gpu_tensor<T>("t", o, o, v, v), g = create_random_gpu_tensor<T>("g", o + 2, o, v, v);
        auto W = create_random_gpu_tensor<T>("W", o + 2, v, v, o);
        auto hW0 = host(W);
        AlgorithmChoice alg;
        einsum(T(1.0), Indices{m, b, e, j}, &W, T(-0.5), Indices{j, n, f, b}, t, Indices{m, n, e, f}, g, &alg);
        CHECK(alg == ta::GENERIC);
        std::string kernel = eb200_last_kernel_name();
        CHECK(kernel.rfind(want_prefix, 0) == 0);
        auto ht = host(t), hg = host(g), hW = host(W);
        for (size_t mm : {size_t(0), size_t(o + 1)}) for (size_t bb : {size_t(3), size_t(v - 1)}) for (size_t ee : {size_t(0), size_t(17)}) for (size_t jj : {size_t(5)}) {
            std::complex<double> s = 0;
            for (size_t nn = 0; nn < o; nn++) for (size_t ff = 0; ff < v; ff++) s += std::complex<double>(ht(jj, nn, ff, bb)) * std::complex<double>(hg(mm, nn, ee, ff));
            std::complex<double> want = std::complex<double>(hW0(mm, bb, ee, jj)) - 0.5 * s;
            CHECK(std::abs(std::complex<double>(hW(mm, bb, ee, jj)) - want) <= tol * std::max(1.0, std::abs(want)) * double(o * v));
        }
    };
    check(float(0), 1e-6, "gemm_tf32x3_tcgen05");
    check(std::complex<float>(0), 1e-6, "gemm_tf32x3_tcgen05<c64>");
    check(std::complex<double>(0), 1e-14, "gemm_tma_z");
}

TEST_CASE(gpu_mixed_precision) { // GPUTensorAlgebra.hip:2301-2354: d-f-d, d-f-f, f-d-d, cd-cd-d (always GENERIC)
    using namespace einsums::index;
    auto run = [&](auto tc, auto ta, auto tb) {
        using TC = decltype(tc); using TA = decltype(ta); using TB = decltype(tb);
        auto A = create_random_gpu_tensor<TA>("A", 10, 10);
        auto B = create_random_gpu_tensor<TB>("B", 10, 10);
        DeviceTensor<TC, 2> C{"C", 10, 10};
        C.zero();
        AlgorithmChoice alg;
        einsum(Indices{i, j}, &C, Indices{i, k}, A, Indices{k, j}, B, &alg);
        CHECK(alg == ((std::is_same_v<TC, TA> && std::is_same_v<TC, TB>) ? ta::GEMM : ta::GENERIC));
        auto hA = host(A); auto hB = host(B); auto hC = host(C);
        for (size_t x = 0; x < 10; x++) for (size_t y = 0; y < 10; y++) {
            std::complex<double> s = 0;
            for (size_t k0 = 0; k0 < 10; k0++) s += std::complex<double>(hA(x, k0)) * std::complex<double>(hB(k0, y));
            CHECK_CLOSE(std::complex<double>(hC(x, y)), s, 1e-4);
        }
    };
    run(double{}, float{}, double{});
    run(double{}, float{}, float{});
    run(float{}, double{}, double{});
    run(std::complex<double>{}, std::complex<double>{}, double{});
    run(double{}, double{}, double{});
}

TEST_CASE(gpu_batched_containers) { // Batched.cpp:13-100
    using namespace einsums::index;
    std::vector<DeviceTensor<double, 2>> As, Bs, Cs;
    for (int b = 0; b < 5; b++) {
        As.push_back(create_random_gpu_tensor("A", 7, 9));
        Bs.push_back(create_random_gpu_tensor("B", 9, 6));
        Cs.emplace_back("C", 7, 6);
    }
    AlgorithmChoice alg;
    einsum(Indices{i, j}, &Cs, Indices{i, k}, As, Indices{k, j}, Bs, &alg);
    CHECK(alg == ta::GEMM);
    for (int b = 0; b < 5; b++) {
        auto hA = host(As[b]), hB = host(Bs[b]), hC = host(Cs[b]);
        for (size_t x = 0; x < 7; x++) for (size_t y = 0; y < 6; y++) {
            double s = 0;
            for (size_t k0 = 0; k0 < 9; k0++) s += hA(x, k0) * hB(k0, y);
            CHECK_CLOSE(hC(x, y), s, 1e-12);
        }
    }
    Bs.pop_back();
    CHECK_THROWS_AS(einsum(Indices{i, j}, &Cs, Indices{i, k}, As, Indices{k, j}, Bs), bad_logic);
}

TEST_CASE(gpu_smart_pointer_operands) { // TensorAlgebra/TensorAlgebra.hpp:60-124: any of A, B, C behind a smart pointer
    using namespace einsums::index;
    auto A = std::make_shared<DeviceTensor<double, 2>>(create_random_gpu_tensor("A", 11, 13));
    auto B = std::make_unique<DeviceTensor<double, 2>>(create_random_gpu_tensor("B", 13, 5));
    auto C = std::make_shared<DeviceTensor<double, 2>>("C", 11, 5);
    DeviceTensor<double, 2> C2{"C2", 11, 5}, C3{"C3", 11, 5};
    AlgorithmChoice         alg;
    einsum(Indices{i, j}, &C, Indices{i, k}, A, Indices{k, j}, B, &alg); // all three
    CHECK(alg == ta::GEMM);
    einsum(Indices{i, j}, &C2, Indices{i, k}, A, Indices{k, j}, *B);            // only A
    einsum(0.0, Indices{i, j}, &C3, 1.0, Indices{i, k}, *A, Indices{k, j}, B);  // only B, prefactor form
    auto hA = host(*A), hB = host(*B), hC = host(*C), hC2 = host(C2), hC3 = host(C3);
    for (size_t x = 0; x < 11; x++) for (size_t y = 0; y < 5; y++) {
        double s = 0;
        for (size_t k0 = 0; k0 < 13; k0++) s += hA(x, k0) * hB(k0, y);
        CHECK_CLOSE(hC(x, y), s, 1e-12);
        CHECK(hC2(x, y) == hC(x, y));
        CHECK(hC3(x, y) == hC(x, y));
    }
    einsum(1.0, Indices{i, j}, &C, 2.0, Indices{i, k}, A, Indices{k, j}, B); // accumulate through the smart-pointer C
    auto hC4 = host(*C);
    for (size_t x = 0; x < 11; x++) for (size_t y = 0; y < 5; y++) CHECK_CLOSE(hC4(x, y), 3.0 * hC(x, y), 1e-12);
    std::shared_ptr<DeviceTensor<double, 2>> none;
    CHECK_THROWS_AS(einsum(Indices{i, j}, &C2, Indices{i, k}, none, Indices{k, j}, *B), bad_logic);
}

TEST_CASE(gpu_host_modes_and_layout_conversion) { // T/TensorForward.hpp:45, T/Backends/DeviceTensor.hpp:62-116, 1010-1090, T/DeviceTensor.hpp:1000-1053
    using namespace einsums::index;
    // column-major host tensor -> row-major device tensor and back
    Tensor<double, 3> hcm(false, "hcm", 5, 7, 4), hrm(true, "hrm", 5, 7, 4);
    for (int x = 0; x < 5; x++) for (int y = 0; y < 7; y++) for (int z = 0; z < 4; z++) hcm(x, y, z) = hrm(x, y, z) = 100 * x + 10 * y + z + 0.5;
    CHECK(hcm.stride(0) == 1 && hrm.stride(2) == 1);
    DeviceTensor<double, 3> D{"D", 5, 7, 4};
    D = hcm;
    CHECK(D.stride(2) == 1 && D.stride(0) == 28);
    auto flat = D.to_host();
    for (int x = 0; x < 5; x++) for (int y = 0; y < 7; y++) for (int z = 0; z < 4; z++) CHECK(flat[(x * 7 + y) * 4 + z] == hcm(x, y, z));
    D.zero();
    D = hrm;
    CHECK(D.to_host() == flat);
    Tensor<double, 3> back_cm(false, "b", 5, 7, 4);
    D.copy_to(&back_cm);
    auto back_rm = Tensor<double, 3>(D);
    for (int x = 0; x < 5; x++) for (int y = 0; y < 7; y++) for (int z = 0; z < 4; z++) {
        CHECK(back_cm(x, y, z) == hcm(x, y, z));
        CHECK(back_rm(x, y, z) == hcm(x, y, z));
    }
    Tensor<double, 3> wrong(true, "w", 5, 7, 5);
    CHECK_THROWS_AS(D = wrong, dimension_error);

    // MAPPED / PINNED: host-side storage the device reads in place; results visible through data() after the op's wait
    for (auto mode : {einsums::detail::MAPPED, einsums::detail::PINNED}) {
        DeviceTensor<double, 2> A{"A", mode, 9, 6}, B{"B", mode, 6, 8}, C{"C", mode, 9, 8};
        CHECK(A.data() != nullptr && A.mode() == mode);
        for (int x = 0; x < 54; x++) A.data()[x] = 0.25 * x - 3;
        for (int x = 0; x < 48; x++) B.data()[x] = 1.0 / (1 + x);
        einsum(Indices{i, j}, &C, Indices{i, k}, A, Indices{k, j}, B);
        for (int x = 0; x < 9; x++) for (int y = 0; y < 8; y++) {
            double s = 0;
            for (int k0 = 0; k0 < 6; k0++) s += A.data()[x * 6 + k0] * B.data()[k0 * 8 + y];
            CHECK_CLOSE(C.data()[x * 8 + y], s, 1e-12);
            CHECK_CLOSE((double)C(x, y), s, 1e-12);
        }
        DeviceTensor<double, 2> Cd(C); // copy keeps the mode and the values
        CHECK(Cd.mode() == mode && Cd.data() != C.data() && Cd.data()[5] == C.data()[5]);
        DeviceTensor<double, 3> M(hcm, mode);
        CHECK(M.data()[(2 * 7 + 3) * 4 + 1] == hcm(2, 3, 1));
    }
    DeviceTensor<double, 2> dev{"dev", 3, 4};
    CHECK(dev.data() == nullptr && dev.mode() == einsums::detail::DEV_ONLY);

    // gpu_dims() / gpu_strides(): device-resident size_t arrays
    size_t meta[6];
    EB200_CHECK(eb200_memcpy_d2h(meta, D.gpu_dims(), 3 * sizeof(size_t), gpu::get_stream()));
    EB200_CHECK(eb200_memcpy_d2h(meta + 3, D.gpu_strides(), 3 * sizeof(size_t), gpu::get_stream()));
    CHECK(meta[0] == 5 && meta[1] == 7 && meta[2] == 4 && meta[3] == 28 && meta[4] == 4 && meta[5] == 1);
}

TEST_CASE(gpu_memory_budget) { // GPUMemory/src/InitModule.cpp:39-55 (--einsums:gpu-buffer-size), GPUAllocator.hpp:157-191
    size_t budget = 0, in_use0 = 0, peak = 0, live0 = 0;
    EB200_CHECK(eb200_memory_stats(&budget, &in_use0, &peak, &live0));
    CHECK(budget == 0); // unlimited by default
    EB200_CHECK(eb200_set_memory_budget(in_use0 + (size_t(8) << 20)));
    {
        DeviceTensor<double, 2> small{"small", 512, 512};   // 2 MiB: fits
        size_t in_use = 0, live = 0;
        EB200_CHECK(eb200_memory_stats(nullptr, &in_use, nullptr, &live));
        CHECK(in_use == in_use0 + 512 * 512 * 8 && live == live0 + 1);
        bool refused = false;
        try { DeviceTensor<double, 2> big{"big", 1024, 1024}; } catch (gpu_out_of_memory const &) { refused = true; } // 8 MiB more: over the budget
        CHECK(refused);
        DeviceTensor<double, 2> pinned{"pinned", einsums::detail::PINNED, 1024, 1024};   // host memory is not charged to the device budget
        CHECK(pinned.data() != nullptr);
    }
    size_t in_use1 = 0;
    EB200_CHECK(eb200_memory_stats(nullptr, &in_use1, nullptr, nullptr));
    CHECK(in_use1 == in_use0); // everything returned
    EB200_CHECK(eb200_set_memory_budget(0));
    DeviceTensor<double, 2> big{"big", 1024, 1024};   // unlimited again
    CHECK(big.size() == 1024 * 1024);
}

TEST_CASE(gpu_blas_gpu_facade) { // hipBLAS.hpp:63-94 (gemm), :132-160 (gemv), :423 (scal), :545 (dot), :586 (dotc), :621 (axpy), :719 (ger)
    namespace bg = einsums::blas::gpu;
    using cd = std::complex<double>;
    const int m = 7, n = 5, k = 9;
    // column-major operands held in rank-1 device tensors
    auto A = create_random_gpu_tensor("A", m * k), B = create_random_gpu_tensor("B", k * n), C = create_random_gpu_tensor("C", m * n);
    auto hA = A.to_host(), hB = B.to_host(), hC = C.to_host();
    bg::gemm<double>('n', 'n', m, n, k, 0.5, A.gpu_data(), m, B.gpu_data(), k, 2.0, C.gpu_data(), m);
    auto out = C.to_host();
    for (int x = 0; x < m; x++) for (int y = 0; y < n; y++) {
        double s = 0;
        for (int q = 0; q < k; q++) s += hA[x + q * m] * hB[q + y * k];
        CHECK_CLOSE(out[x + y * m], 2.0 * hC[x + y * m] + 0.5 * s, 1e-12);
    }
    bg::gemm<double>('t', 'n', k, n, m, 1.0, A.gpu_data(), m, C.gpu_data(), m, 0.0, B.gpu_data(), k); // B = A^T C
    auto outB = B.to_host();
    for (int x = 0; x < k; x++) for (int y = 0; y < n; y++) {
        double s = 0;
        for (int q = 0; q < m; q++) s += hA[q + x * m] * out[q + y * m];
        CHECK_CLOSE(outB[x + y * k], s, 1e-12);
    }
    CHECK_THROWS_AS(bg::gemm<double>('x', 'n', m, n, k, 1.0, A.gpu_data(), m, B.gpu_data(), k, 0.0, C.gpu_data(), m), std::invalid_argument);
    CHECK_THROWS_AS(bg::gemm<double>('n', 'n', m, n, k, 1.0, A.gpu_data(), m - 1, B.gpu_data(), k, 0.0, C.gpu_data(), m), std::domain_error);
    // gemv, axpy, axpby, scal, rscl, dot, ger
    auto x1 = create_random_gpu_tensor("x", k), y1 = create_random_gpu_tensor("y", m);
    auto hx = x1.to_host(), hy = y1.to_host();
    bg::gemv<double>('n', m, k, 1.5, A.gpu_data(), m, x1.gpu_data(), 1, -1.0, y1.gpu_data(), 1);
    auto oy = y1.to_host();
    for (int r0 = 0; r0 < m; r0++) {
        double s = 0;
        for (int q = 0; q < k; q++) s += hA[r0 + q * m] * hx[q];
        CHECK_CLOSE(oy[r0], 1.5 * s - hy[r0], 1e-12);
    }
    double d0 = bg::dot<double>(m, y1.gpu_data(), 1, A.gpu_data(), 1), want = 0;
    for (int r0 = 0; r0 < m; r0++) want += oy[r0] * hA[r0];
    CHECK_CLOSE(d0, want, 1e-12);
    bg::scal<double>(m, 3.0, y1.gpu_data(), 1);
    bg::rscl<double>(m, 2.0, y1.gpu_data(), 1);
    bg::axpy<double>(m, -1.0, A.gpu_data(), 1, y1.gpu_data(), 1);
    bg::axpby<double>(m, 2.0, A.gpu_data(), 1, 0.5, y1.gpu_data(), 1);
    auto oy2 = y1.to_host();
    for (int r0 = 0; r0 < m; r0++) CHECK_CLOSE(oy2[r0], 2.0 * hA[r0] + 0.5 * (1.5 * oy[r0] - hA[r0]), 1e-12);
    auto G = create_zero_gpu_tensor("G", m * k);
    bg::ger<double>(m, k, 2.0, y1.gpu_data(), 1, x1.gpu_data(), 1, G.gpu_data(), m);
    auto hG = G.to_host();
    for (int r0 = 0; r0 < m; r0++) for (int q = 0; q < k; q++) CHECK_CLOSE(hG[r0 + q * m], 2.0 * oy2[r0] * hx[q], 1e-12);
    // complex: dotc conjugates the first vector, gerc the second, scal by a real
    auto zx = create_random_gpu_tensor<cd>("zx", 11), zy = create_random_gpu_tensor<cd>("zy", 11);
    auto hzx = zx.to_host(), hzy = zy.to_host();
    cd   dc = bg::dotc<cd>(11, zx.gpu_data(), 1, zy.gpu_data(), 1), du = bg::dot<cd>(11, zx.gpu_data(), 1, zy.gpu_data(), 1), wc = 0, wu = 0;
    for (int q = 0; q < 11; q++) { wc += std::conj(hzx[q]) * hzy[q]; wu += hzx[q] * hzy[q]; }
    CHECK_CLOSE(dc, wc, 1e-12);
    CHECK_CLOSE(du, wu, 1e-12);
    auto ZG = create_zero_gpu_tensor<cd>("ZG", 11 * 11);
    bg::gerc<cd>(11, 11, cd(0.5, 1.0), zx.gpu_data(), 1, zy.gpu_data(), 1, ZG.gpu_data(), 11);
    auto hZG = ZG.to_host();
    for (int r0 = 0; r0 < 11; r0++) for (int q = 0; q < 11; q++) CHECK_CLOSE(hZG[r0 + q * 11], cd(0.5, 1.0) * hzx[r0] * std::conj(hzy[q]), 1e-12);
    bg::scal<cd>(11, 2.0, zx.gpu_data(), 1);
    auto hzx2 = zx.to_host();
    for (int q = 0; q < 11; q++) CHECK_CLOSE(hzx2[q], 2.0 * hzx[q], 1e-12);
}

TEST_CASE(gpu_permute) { // Sort.cpp:18-250, GPUTensorAlgebra.hip:591-838
    using namespace einsums::index;
    auto A = create_random_gpu_tensor("A", 3, 4, 5, 6);
    auto hA = host(A);
    DeviceTensor<double, 4> B{"B", 6, 5, 4, 3};
    permute(Indices{l, k, j, i}, &B, Indices{i, j, k, l}, A);
    auto hB = host(B);
    for (size_t x = 0; x < 3; x++) for (size_t y = 0; y < 4; y++) for (size_t z = 0; z < 5; z++) for (size_t t = 0; t < 6; t++)
        CHECK(hB(t, z, y, x) == hA(x, y, z, t));
    // B[i,l,k,j] = 0.5 * A[k,j,l,i] (Sort.cpp:155-179) with accumulation
    auto A2 = create_random_gpu_tensor("A2", 3, 3, 3, 3), B2 = create_random_gpu_tensor("B2", 3, 3, 3, 3);
    auto hA2 = host(A2), h0 = host(B2);
    permute(2.0, Indices{i, l, k, j}, &B2, 0.5, Indices{k, j, l, i}, A2);
    auto h1 = host(B2);
    for (size_t x = 0; x < 3; x++) for (size_t y = 0; y < 3; y++) for (size_t z = 0; z < 3; z++) for (size_t t = 0; t < 3; t++)
        CHECK_CLOSE(h1(x, t, z, y), 2.0 * h0(x, t, z, y) + 0.5 * hA2(z, y, t, x), 1e-14);
    // rank 2 transpose and rank 3
    auto M = create_random_gpu_tensor("M", 37, 53);
    DeviceTensor<double, 2> Mt{"Mt", 53, 37};
    transpose(&Mt, M);
    auto hM = host(M), hMt = host(Mt);
    for (size_t x = 0; x < 37; x++) for (size_t y = 0; y < 53; y++) CHECK(hMt(y, x) == hM(x, y));
    // output smaller than input -> dimension_error (DevicePermute.hpp:121-126)
    DeviceTensor<double, 2> small{"small", 53, 36};
    CHECK_THROWS_AS(transpose(&small, M), dimension_error);
}

TEST_CASE(gpu_linear_algebra) { // GPULinearAlgebra.hip: gemm / gemv / ger / scale / axpy / direct_product
    auto A = create_random_gpu_tensor("A", 8, 6), B = create_random_gpu_tensor("B", 6, 7);
    DeviceTensor<double, 2> C{"C", 8, 7};
    C.set_all(1.0);
    linear_algebra::gemm<false, false>(2.0, A, B, 0.5, &C);
    auto hA = host(A), hB = host(B), hC = host(C);
    for (size_t x = 0; x < 8; x++) for (size_t y = 0; y < 7; y++) {
        double s = 0.5;
        for (size_t k0 = 0; k0 < 6; k0++) s += 2.0 * hA(x, k0) * hB(k0, y);
        CHECK_CLOSE(hC(x, y), s, 1e-12);
    }
    DeviceTensor<double, 2> Ct{"Ct", 6, 6};
    linear_algebra::gemm<true, false>(1.0, A, A, 0.0, &Ct); // A^T A
    auto hCt = host(Ct);
    for (size_t x = 0; x < 6; x++) for (size_t y = 0; y < 6; y++) {
        double s = 0;
        for (size_t k0 = 0; k0 < 8; k0++) s += hA(k0, x) * hA(k0, y);
        CHECK_CLOSE(hCt(x, y), s, 1e-12);
    }
    auto xv = create_random_gpu_tensor("x", 6), yv = create_random_gpu_tensor("y", 8);
    auto hx = host(xv), hy0 = host(yv);
    linear_algebra::gemv<false>(1.5, A, xv, 2.0, &yv);
    auto hy = host(yv);
    for (size_t r0 = 0; r0 < 8; r0++) {
        double s = 2.0 * hy0(r0);
        for (size_t k0 = 0; k0 < 6; k0++) s += 1.5 * hA(r0, k0) * hx(k0);
        CHECK_CLOSE(hy(r0), s, 1e-12);
    }
    auto x8 = create_random_gpu_tensor("x8", 8);
    DeviceTensor<double, 1> y6{"y6", 6};
    linear_algebra::gemv<true>(1.0, A, x8, 0.0, &y6);
    auto hx8 = host(x8), hy6 = host(y6);
    for (size_t c0 = 0; c0 < 6; c0++) {
        double s = 0;
        for (size_t r0 = 0; r0 < 8; r0++) s += hA(r0, c0) * hx8(r0);
        CHECK_CLOSE(hy6(c0), s, 1e-12);
    }
    auto G = create_random_gpu_tensor("G", 8, 6);
    auto hG0 = host(G);
    linear_algebra::ger(0.5, x8, xv, &G);
    auto hG = host(G);
    for (size_t r0 = 0; r0 < 8; r0++) for (size_t c0 = 0; c0 < 6; c0++) CHECK_CLOSE(hG(r0, c0), hG0(r0, c0) + 0.5 * hx8(r0) * hx(c0), 1e-14);
    linear_algebra::scale(3.0, &G);
    auto hG3 = host(G);
    for (size_t x = 0; x < hG.v.size(); x++) CHECK(hG3.v[x] == 3.0 * hG.v[x]);
    G *= 0.5;
    G += A;
    auto hG4 = host(G);
    for (size_t x = 0; x < hG.v.size(); x++) CHECK_CLOSE(hG4.v[x], 1.5 * hG.v[x] + hA.v[x], 1e-14);
    // vendor-layer argument errors keep the reference's exception classes (hipBLASVendor/src/gemm.hip:39-80)
    double one = 1.0;
    CHECK(eb200_dgemm(nullptr, 'x', 'n', 2, 2, 2, &one, A.gpu_data(), 2, A.gpu_data(), 2, &one, C.gpu_data(), 2) == EB200_ERR_INVALID_ARGUMENT);
    CHECK_THROWS_AS(EB200_CHECK(eb200_dgemm(nullptr, 'n', 'n', 4, 2, 2, &one, A.gpu_data(), 2, A.gpu_data(), 2, &one, C.gpu_data(), 4)), std::domain_error);
}

TEST_CASE(gpu_dimension_errors_and_threads) {
    using namespace einsums::index;
    auto A = create_random_gpu_tensor("A", 4, 5), B = create_random_gpu_tensor("B", 6, 3);
    DeviceTensor<double, 2> C{"C", 4, 3};
    CHECK_THROWS_AS(einsum(Indices{i, j}, &C, Indices{i, k}, A, Indices{k, j}, B), dimension_error);
    // concurrent einsum calls from different host threads, each on its own stream (GPUTensorAlgebraTasks.hip:37-52)
    auto X = create_random_gpu_tensor("X", 64, 64), Y = create_random_gpu_tensor("Y", 64, 64);
    DeviceTensor<double, 2> C1{"C1", 64, 64}, C2{"C2", 64, 64};
    std::thread t1([&] { for (int r = 0; r < 20; r++) einsum(Indices{i, j}, &C1, Indices{i, k}, X, Indices{k, j}, Y); });
    std::thread t2([&] { for (int r = 0; r < 20; r++) einsum(Indices{i, j}, &C2, Indices{i, k}, X, Indices{k, j}, Y); });
    t1.join();
    t2.join();
    auto h1 = host(C1), h2 = host(C2);
    bool same = true;
    for (size_t x = 0; x < h1.v.size(); x++) same = same && (h1.v[x] == h2.v[x]);
    CHECK(same);
}

// ---- block-diagonal tensors ------------------------------------------------------------------
TEST_CASE(gpu_block_tensor_creation_and_ops) { // Tensor/tests/unit/BlockDeviceTensor.hip:14-195
    BlockDeviceTensor<double, 2> A("A", einsums::detail::DEV_ONLY, 3, 3), B("B", einsums::detail::DEV_ONLY, 3, 3), C;
    CHECK(A.dim(0) == 6 && A.dim(1) == 6 && A.num_blocks() == 2);
    A.zero();
    B.zero();
    for (int i = 0; i < 6; i++) {
        A(i, i) = 1.0;
        B(i, i) = 1.0;
    }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) CHECK((double)A(i, j) == (i == j ? 1.0 : 0.0));
    A(0, 5) = 7.0; // outside the blocks: swallowed
    CHECK((double)A(0, 5) == 0.0);

    C.push_block(DeviceTensor<double, 2>("block2", einsums::detail::DEV_ONLY, 3, 3));
    C.insert_block(0, DeviceTensor<double, 2>("block1", einsums::detail::DEV_ONLY, 3, 3));
    C.set_name("C");
    linear_algebra::gemm<false, false>(1.0, A, B, 0.0, &C);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) CHECK((double)C(i, j) == (i == j ? 1.0 : 0.0));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            CHECK((double)A[0](i, j) == (i == j ? 1.0 : 0.0));
            CHECK((double)C["block1"](i, j) == (i == j ? 1.0 : 0.0));
            CHECK((double)C["block2"](i, j) == (i == j ? 1.0 : 0.0));
        }
    CHECK_THROWS_AS(C["block3"], std::out_of_range);

    A.set_all(2.0);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            CHECK((A.block_of(i) == A.block_of(j)) == (i / 3 == j / 3));
            CHECK((double)A(i, j) == (i / 3 == j / 3 ? 2.0 : 0.0));
        }
    for (int i = 0; i < 6; i++) CHECK(A.block_of(i) == i / 3);
    CHECK_THROWS_AS(A.block_of(6), std::out_of_range);
    CHECK_THROWS_AS(A.block_of(-1), std::out_of_range);

    // arithmetic (:121-181)
    A.zero();
    for (int i = 0; i < 6; i++) A(i, i) = 1.0;
    A *= 2.0; A += 2.0; A /= 2.0; A -= 2.0;
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) CHECK((double)A(i, j) == (i == j ? 0.0 : (i / 3 == j / 3 ? -1.0 : 0.0)));
    A += B; B *= A; B /= A; B -= A;
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            CHECK((double)A(i, j) == (i == j ? 1.0 : (i / 3 == j / 3 ? -1.0 : 0.0)));
            CHECK((double)B(i, j) == (i == j ? 0.0 : (i / 3 == j / 3 ? 1.0 : 0.0)));
        }
    // dense copy
    DeviceTensor<double, 2> D = (DeviceTensor<double, 2>)A;
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) CHECK((double)D(i, j) == (double)A(i, j));
}

TEST_CASE(gpu_block_tensor_einsum) { // BlockAlgebra.hpp: block x block -> block / dense / scalar, vector x block
    using namespace einsums::index;
    std::vector<int64_t> sizes{5, 0, 17, 17, 17, 33};
    BlockDeviceTensor<double, 2> A("A", sizes), B("B", sizes), C("C", sizes);
    int64_t n = A.dim();
    std::mt19937_64 eng(7);
    std::uniform_real_distribution<double> dist(-1, 1);
    for (size_t b = 0; b < A.num_blocks(); b++) {
        if (!A.block_dim(b)) continue;
        std::vector<double> ha(size_t(A[b].size())), hb(ha.size()), hc(ha.size());
        for (auto &v : ha) v = dist(eng);
        for (auto &v : hb) v = dist(eng);
        for (auto &v : hc) v = dist(eng);
        A[b].read(ha); B[b].read(hb); C[b].read(hc);
    }
    auto Ad = (DeviceTensor<double, 2>)A, Bd = (DeviceTensor<double, 2>)B, Cd0 = (DeviceTensor<double, 2>)C;
    auto ha = host(Ad), hb = host(Bd), hc0 = host(Cd0);

    // block x block -> block, with prefactors
    einsum(0.5, Indices{i, j}, &C, 2.0, Indices{i, k}, A, Indices{j, k}, B);
    auto Cd = (DeviceTensor<double, 2>)C;
    auto hc = host(Cd);
    double worst = 0;
    for (int64_t x = 0; x < n; x++)
        for (int64_t y = 0; y < n; y++) {
            double ref = 0.5 * hc0(x, y);
            if (A.block_of(x) == A.block_of(y))
                for (int64_t z = 0; z < n; z++) ref += 2.0 * ha(x, z) * hb(y, z);
            worst = std::max(worst, std::abs(ref - hc(x, y)));
        }
    CHECK(worst <= 1e-12);

    // block x block -> dense C: only the diagonal sub-blocks are touched (BlockAlgebra.hpp:52-88)
    DeviceTensor<double, 2> F("F", n, n);
    F.set_all(3.0);
    einsum(Indices{i, j}, &F, Indices{i, k}, A, Indices{k, j}, B);
    auto hf = host(F);
    worst = 0;
    for (int64_t x = 0; x < n; x++)
        for (int64_t y = 0; y < n; y++) {
            double ref = 3.0;
            if (A.block_of(x) == A.block_of(y)) {
                ref = 0;
                for (int64_t z = 0; z < n; z++) ref += ha(x, z) * hb(z, y);
            }
            worst = std::max(worst, std::abs(ref - hf(x, y)));
        }
    CHECK(worst <= 1e-12);

    // block x block -> scalar, host value and rank-0 device tensor (:108-180)
    double s = 4.0, ref = 0;
    for (int64_t x = 0; x < n; x++)
        for (int64_t y = 0; y < n; y++) ref += ha(x, y) * hb(x, y);
    einsum(0.5, Indices{}, &s, 1.0, Indices{i, j}, A, Indices{i, j}, B);
    CHECK_CLOSE(s, 2.0 + ref, 1e-12 * (1 + std::abs(ref)));
    DeviceTensor<double, 0> s0("s0");
    s0 = 1.0;
    einsum(0.0, Indices{}, &s0, 2.0, Indices{i, j}, A, Indices{i, j}, B);
    CHECK_CLOSE((double)s0, 2.0 * ref, 1e-12 * (1 + std::abs(ref)));
    CHECK_CLOSE(linear_algebra::dot(A, B), ref, 1e-12 * (1 + std::abs(ref)));

    // vector x block -> dense vector (:217-238) and block x vector -> vector through gemv
    auto x = create_random_gpu_tensor("x", n);
    DeviceTensor<double, 1> y("y", n), y2("y2", n);
    y.zero();
    einsum(Indices{j}, &y, Indices{i}, x, Indices{i, j}, A);
    auto hx = host(x), hy = host(y);
    worst = 0;
    for (int64_t c = 0; c < n; c++) {
        double r = 0;
        for (int64_t rw = 0; rw < n; rw++) r += hx(rw) * ha(rw, c);
        worst = std::max(worst, std::abs(r - hy(c)));
    }
    CHECK(worst <= 1e-12);
    y2.zero();
    linear_algebra::gemv<true>(1.0, A, x, 0.0, &y2);
    auto hy2 = host(y2);
    worst = 0;
    for (int64_t c = 0; c < n; c++) worst = std::max(worst, std::abs(hy2(c) - hy(c)));
    CHECK(worst <= 1e-12);

    // rank-3 blocks, equal sizes: folds into one batched launch
    BlockDeviceTensor<double, 3> T3("T3", 8, 8, 8, 8);
    BlockDeviceTensor<double, 2> M2("M2", 8, 8, 8, 8), R2("R2", 8, 8, 8, 8);
    for (size_t b = 0; b < 4; b++) {
        std::vector<double> t(512), m(64);
        for (auto &v : t) v = dist(eng);
        for (auto &v : m) v = dist(eng);
        T3[b].read(t); M2[b].read(m);
    }
    eb200_launch_count_reset();
    einsum(Indices{i, l}, &R2, Indices{i, k, l}, T3, Indices{k, l}, M2);
    for (size_t b = 0; b < 4; b++) {
        auto t = T3[b].to_host(), m = M2[b].to_host(), r = R2[b].to_host();
        double w = 0;
        for (int a = 0; a < 8; a++)
            for (int c = 0; c < 8; c++) {
                double e = 0;
                for (int k2 = 0; k2 < 8; k2++) e += t[(a * 8 + k2) * 8 + c] * m[k2 * 8 + c];
                w = std::max(w, std::abs(e - r[a * 8 + c]));
            }
        CHECK(w <= 1e-12);
    }

    // mismatched block structures
    BlockDeviceTensor<double, 2> P("P", 3, 4), Q("Q", 4, 3), Rr("R", 3, 4, 5);
    BlockDeviceTensor<double, 2> O("O", 3, 4);
    CHECK_THROWS_AS(einsum(Indices{i, j}, &O, Indices{i, k}, P, Indices{k, j}, Q), dimension_error);
    CHECK_THROWS_AS(einsum(Indices{i, j}, &O, Indices{i, k}, P, Indices{k, j}, Rr), tensor_compat_error);
}

// ---- tiled tensors ---------------------------------------------------------------------------
TEST_CASE(gpu_tiled_tensor_known_answers) { // Tensor/tests/unit/DeviceTiledTensor.hip:13-156, 432-504
    using namespace einsums::index;
    {
        TiledDeviceTensor<double, 2> A("A", einsums::detail::DEV_ONLY, std::array{1, 0, 2}), B("B", einsums::detail::DEV_ONLY, std::array{1, 0, 2}),
            C("C", einsums::detail::DEV_ONLY, std::array{1, 0, 2});
        CHECK(A.dim(0) == 3 && A.dim(1) == 3);
        CHECK(A.tile_offset(0)[0] == 0 && A.tile_offset(0)[1] == 1 && A.tile_offset(0)[2] == 1 && A.tile_offset(1)[2] == 1);
        A.zero();
        for (int i = 0; i < 3; i++) {
            A(i, i) = 1.0;
            B(i, i) = 1.0;
        }
        linear_algebra::gemm<false, false>(1.0, A, B, 0.0, &C);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) CHECK((double)C(i, j) == (i == j ? 1.0 : 0.0));
    }
    {
        TiledDeviceTensor<double, 2> A("A", einsums::detail::DEV_ONLY, std::array{1, 0, 2}), B("B", einsums::detail::DEV_ONLY, std::array{1, 0, 2}),
            C("C", einsums::detail::DEV_ONLY, std::array{1, 0, 2});
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                A(i, j) = 3 * i + j + 1;
                B(i, j) = 33 * i + 11 * j + 11;
            }
        double nn[9] = {330, 396, 462, 726, 891, 1056, 1122, 1386, 1650}, tn[9] = {726, 858, 990, 858, 1023, 1188, 990, 1188, 1386},
               nt[9] = {154, 352, 550, 352, 847, 1342, 550, 1342, 2134}, tt[9] = {330, 726, 1122, 396, 891, 1386, 462, 1056, 1650};
        linear_algebra::gemm<false, false>(1.0, A, B, 0.0, &C);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) CHECK((double)C(i, j) == nn[3 * i + j]);
        linear_algebra::gemm<true, false>(1.0, A, B, 0.0, &C);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) CHECK((double)C(i, j) == tn[3 * i + j]);
        linear_algebra::gemm<false, true>(1.0, A, B, 0.0, &C);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) CHECK((double)C(i, j) == nt[3 * i + j]);
        linear_algebra::gemm<true, true>(1.0, A, B, 0.0, &C);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) CHECK((double)C(i, j) == tt[3 * i + j]);

        DeviceTensor<double, 1> x("x", 3), y("y", 3);
        for (int i = 0; i < 3; i++) x(i) = 11 * i + 11;
        linear_algebra::gemv<false>(1.0, A, x, 0.0, &y);
        CHECK((double)y(0) == 154.0 && (double)y(1) == 352.0 && (double)y(2) == 550.0);
        linear_algebra::gemv<true>(1.0, A, x, 0.0, &y);
        CHECK((double)y(0) == 330.0 && (double)y(1) == 396.0 && (double)y(2) == 462.0);
    }
    { // ik=ij,jk on 1..9 (:435-467)
        TiledDeviceTensor<double, 2> A("A", einsums::detail::DEV_ONLY, std::array{1, 0, 2}), B("B", einsums::detail::DEV_ONLY, std::array{1, 0, 2}),
            C("C", einsums::detail::DEV_ONLY, std::array{1, 0, 2});
        for (int i = 0, ij = 1; i < 3; i++)
            for (int j = 0; j < 3; j++, ij++) {
                A(i, j) = ij;
                B(i, j) = ij;
            }
        einsum(Indices{i, j}, &C, Indices{i, k}, A, Indices{k, j}, B);
        double want[9] = {30, 36, 42, 66, 81, 96, 102, 126, 150};
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) CHECK((double)C(a, b) == want[3 * a + b]);
    }
    { // il=ijk,jkl on 1..27 (:468-503)
        TiledDeviceTensor<double, 3> A("A", einsums::detail::DEV_ONLY, std::array{1, 0, 2}), B("B", einsums::detail::DEV_ONLY, std::array{1, 0, 2});
        TiledDeviceTensor<double, 2> C("C", einsums::detail::DEV_ONLY, std::array{1, 0, 2});
        for (int a = 0, n = 1; a < 3; a++)
            for (int b = 0; b < 3; b++)
                for (int c = 0; c < 3; c++, n++) {
                    A(a, b, c) = n;
                    B(a, b, c) = n;
                }
        einsum(Indices{i, l}, &C, Indices{i, j, k}, A, Indices{j, k, l}, B);
        double want[9] = {765, 810, 855, 1818, 1944, 2070, 2871, 3078, 3285};
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) CHECK((double)C(a, b) == want[3 * a + b]);
    }
}

TEST_CASE(gpu_tiled_tensor_sparse_random) { // absent tiles, ragged grids, prefactors, scalar results: against a dense host loop
    using namespace einsums::index;
    std::array<std::vector<int64_t>, 2> ga{std::vector<int64_t>{7, 16, 16, 0, 5}, std::vector<int64_t>{9, 9, 20}},
        gb{std::vector<int64_t>{9, 9, 20}, std::vector<int64_t>{16, 16, 16, 3}}, gc{std::vector<int64_t>{7, 16, 16, 0, 5}, std::vector<int64_t>{16, 16, 16, 3}};
    TiledDeviceTensor<double, 2> A("A", einsums::detail::DEV_ONLY, ga), B("B", einsums::detail::DEV_ONLY, gb), C("C", einsums::detail::DEV_ONLY, gc);
    std::mt19937_64 eng(3);
    std::uniform_real_distribution<double> dist(-1, 1);
    auto fill = [&](auto &X, int skip_mod) {
        int n = 0;
        X.for_each_grid([&](auto idx) {
            if (X.has_zero_size(idx) || (n++ % skip_mod) == 1) return; // leave some tiles absent
            auto &t = X.tile(idx);
            std::vector<double> h(size_t(t.size()));
            for (auto &v : h) v = dist(eng);
            t.read(h);
        });
    };
    fill(A, 3);
    fill(B, 4);
    fill(C, 2);
    auto Ad = (DeviceTensor<double, 2>)A, Bd = (DeviceTensor<double, 2>)B, C0 = (DeviceTensor<double, 2>)C;
    auto ha = host(Ad), hb = host(Bd), hc0 = host(C0);
    einsum(0.5, Indices{i, j}, &C, -1.5, Indices{i, k}, A, Indices{k, j}, B);
    auto Cd = (DeviceTensor<double, 2>)C;
    auto hc = host(Cd);
    double worst = 0;
    for (int64_t x = 0; x < C.dim(0); x++)
        for (int64_t y = 0; y < C.dim(1); y++) {
            double ref = 0.5 * hc0(x, y);
            for (int64_t z = 0; z < A.dim(1); z++) ref += -1.5 * ha(x, z) * hb(z, y);
            worst = std::max(worst, std::abs(ref - hc(x, y)));
        }
    CHECK(worst <= 1e-12);

    // scalar: sum_ij A(i,k) A(i,k) and dot
    double s = 1.0, ref = 0;
    for (auto v : ha.v) ref += v * v;
    einsum(2.0, Indices{}, &s, 1.0, Indices{i, k}, A, Indices{i, k}, A);
    CHECK_CLOSE(s, 2.0 + ref, 1e-12 * (1 + ref));
    CHECK_CLOSE(linear_algebra::dot(A, A), ref, 1e-12 * (1 + ref));
    C *= 2.0;
    auto hc2 = host((DeviceTensor<double, 2>)C);
    worst = 0;
    for (size_t q = 0; q < hc.v.size(); q++) worst = std::max(worst, std::abs(hc2.v[q] - 2.0 * hc.v[q]));
    CHECK(worst == 0.0);

    // incompatible grids
    TiledDeviceTensor<double, 2> Bad("Bad", einsums::detail::DEV_ONLY, std::array{4, 4});
    CHECK_THROWS_AS(einsum(Indices{i, j}, &C, Indices{i, k}, A, Indices{k, j}, Bad), tensor_compat_error);
}

TEST_CASE(gpu_block_tile_mixed_algebra) { // TA/Backends/BlockTileAlgebra.hpp:18-174: a block tensor acts as a tiled one with diagonal tiles only
    using namespace einsums::index;
    std::vector<int64_t> sizes{6, 0, 9, 4};
    BlockDeviceTensor<double, 2> Ab("Ab", sizes), Cb("Cb", sizes);
    TiledDeviceTensor<double, 2> Bt("Bt", sizes), Ct("Ct", sizes), At("At", sizes);
    std::mt19937_64 eng(17);
    std::uniform_real_distribution<double> dist(-1, 1);
    for (size_t b = 0; b < Ab.num_blocks(); b++) {
        if (!Ab.block_dim(b)) continue;
        std::vector<double> h(size_t(Ab.block(b).size()));
        for (auto &v : h) v = dist(eng);
        Ab.block(b).read(h);
        for (auto &v : h) v = dist(eng);
        Cb.block(b).read(h);
    }
    int n = 0;
    auto fill = [&](auto &X, int skip_mod) {
        X.for_each_grid([&](auto idx) {
            if (X.has_zero_size(idx) || (n++ % skip_mod) == 2) return;
            auto &t = X.tile(idx);
            std::vector<double> h(size_t(t.size()));
            for (auto &v : h) v = dist(eng);
            t.read(h);
        });
    };
    fill(Bt, 4); fill(At, 3); fill(Ct, 2);
    auto dense_mm = [](Host<double, 2> &a, Host<double, 2> &b, int64_t nn) {
        std::vector<double> out(size_t(nn * nn), 0.0);
        for (int64_t x = 0; x < nn; x++) for (int64_t y = 0; y < nn; y++) { double s = 0; for (int64_t q = 0; q < nn; q++) s += a(x, q) * b(q, y); out[size_t(x * nn + y)] = s; }
        return out;
    };
    const int64_t N = Ab.dim(0);
    auto hA = host((DeviceTensor<double, 2>)Ab), hB = host((DeviceTensor<double, 2>)Bt), hAt = host((DeviceTensor<double, 2>)At);
    // (1) tiled <- block x tiled, with prefactors
    auto hC0 = host((DeviceTensor<double, 2>)Ct);
    AlgorithmChoice alg;
    einsum(0.5, Indices{i, j}, &Ct, 2.0, Indices{i, k}, Ab, Indices{k, j}, Bt, &alg);
    CHECK(alg == ta::GENERIC);
    auto ref = dense_mm(hA, hB, N);
    auto hC = host((DeviceTensor<double, 2>)Ct);
    double worst = 0;
    for (size_t q = 0; q < ref.size(); q++) worst = std::max(worst, std::abs(hC.v[q] - (0.5 * hC0.v[q] + 2.0 * ref[q])));
    CHECK(worst <= 1e-12);
    // (2) tiled <- tiled x block (beta = 0)
    TiledDeviceTensor<double, 2> C2("C2", sizes);
    einsum(Indices{i, j}, &C2, Indices{i, k}, At, Indices{k, j}, Ab);
    ref = dense_mm(hAt, hA, N);
    auto hC2 = host((DeviceTensor<double, 2>)C2);
    worst = 0;
    for (size_t q = 0; q < ref.size(); q++) worst = std::max(worst, std::abs(hC2.v[q] - ref[q]));
    CHECK(worst <= 1e-12);
    // (3) block <- tiled x tiled: only the diagonal blocks of the product are kept
    auto hCb0 = host((DeviceTensor<double, 2>)Cb);
    einsum(1.0, Indices{i, j}, &Cb, -1.0, Indices{i, k}, At, Indices{k, j}, Bt);
    ref = dense_mm(hAt, hB, N);
    auto hCb = host((DeviceTensor<double, 2>)Cb);
    worst = 0;
    for (size_t b = 0; b < Cb.num_blocks(); b++) {
        Range r = Cb.block_range(b);
        for (int64_t x = r[0]; x < r[1]; x++) for (int64_t y = r[0]; y < r[1]; y++)
            worst = std::max(worst, std::abs(hCb(x, y) - (hCb0(x, y) - ref[size_t(x * N + y)])));
    }
    CHECK(worst <= 1e-12);
    // incompatible grids
    BlockDeviceTensor<double, 2> Bad("Bad", std::vector<int64_t>{5, 5});
    CHECK_THROWS_AS(einsum(Indices{i, j}, &Ct, Indices{i, k}, Bad, Indices{k, j}, Bt), tensor_compat_error);

    // tiled ger / direct_product (LinearAlgebra/TiledTensor.hpp:356-430)
    TiledDeviceTensor<double, 1> X("X", sizes), Y("Y", sizes);
    for (int t = 0; t < 4; t++) {
        if (!sizes[size_t(t)]) continue;
        std::vector<double> h(size_t(sizes[size_t(t)]));
        for (auto &v : h) v = dist(eng);
        X.tile(t).read(h);
        if (t != 2) { for (auto &v : h) v = dist(eng); Y.tile(t).read(h); } // Y has no tile 2
    }
    auto hX = host((DeviceTensor<double, 1>)X), hY = host((DeviceTensor<double, 1>)Y);
    auto hG0 = host((DeviceTensor<double, 2>)Ct);
    linear_algebra::ger(0.75, X, Y, &Ct);
    auto hG = host((DeviceTensor<double, 2>)Ct);
    worst = 0;
    for (int64_t x = 0; x < N; x++) for (int64_t y = 0; y < N; y++) worst = std::max(worst, std::abs(hG(x, y) - (hG0(x, y) + 0.75 * hX(x) * hY(y))));
    CHECK(worst <= 1e-12);
    TiledDeviceTensor<double, 2> D("D", sizes);
    linear_algebra::direct_product(2.0, At, Bt, 0.0, &D);
    auto hD = host((DeviceTensor<double, 2>)D);
    worst = 0;
    for (int64_t x = 0; x < N; x++) for (int64_t y = 0; y < N; y++) worst = std::max(worst, std::abs(hD(x, y) - 2.0 * hAt(x, y) * hB(x, y)));
    CHECK(worst <= 1e-12);
}

int main() {
    try {
        einsums::initialize(0);
    } catch (std::exception const &e) {
        std::printf("FATAL: no usable device: %s\n", e.what());
        return 2;
    }
    int rc = mini::run_all();
    einsums::finalize();
    return rc;
}

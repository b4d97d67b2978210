// CPU-only test of the compile-time analyser (no device needed): every `REQUIRE(alg_choice == X)` of the reference's
// TensorAlgebra unit tests as a static_assert on the same Indices{...} packs (tests/golden/algorithm_choice.json holds the
// extracted list with file:line; the two are kept in sync by tests/test_cpp_shim.py).
#include <Einsums/B200/Analysis.hpp>
#include <Einsums/B200/Index.hpp>
#include <Einsums/B200/IndexUtilities.hpp>

#include <cstdio>

using namespace einsums;
using namespace einsums::index;
using namespace einsums::tensor_algebra::detail;

template <bool CA = false, bool CB = false, bool Same = true, typename... Cs, typename... As, typename... Bs>
constexpr AlgorithmChoice choose(std::tuple<Cs...> const &, std::tuple<As...> const &, std::tuple<Bs...> const &) {
    return choice_of(static_route(make_pack(ids_of<std::tuple<Cs...>>()), make_pack(ids_of<std::tuple<As...>>()), make_pack(ids_of<std::tuple<Bs...>>()),
                                  CA, CB, Same));
}
#define IDX(...) Indices{__VA_ARGS__}
#define EXPECT(choice, C_, A_, B_) static_assert(choose(C_, A_, B_) == choice, #C_ " <- " #A_ ", " #B_)

// Einsum1.cpp:31,72,123 / Batched.cpp:34,81 / Einsum3.cpp:25,42 / Einsum4.cpp:28 / SharedPointer.cpp:157
EXPECT(GEMM, IDX(i, j), IDX(i, k), IDX(k, j));
EXPECT(GEMM, IDX(i, l), IDX(i, j, k), IDX(j, k, l));
EXPECT(GEMM, IDX(i, j), IDX(i, k), IDX(j, k));
// Dot.cpp:30-112
EXPECT(DOT, IDX(), IDX(i), IDX(i));
EXPECT(DOT, IDX(), IDX(i, j), IDX(i, j));
EXPECT(DOT, IDX(), IDX(i, j, a), IDX(i, j, a));
EXPECT(DOT, IDX(), IDX(i, j, a, b), IDX(i, j, a, b));
EXPECT(DOT, IDX(), IDX(l, k), IDX(l, k));
// Einsum3.cpp:62,79,107,138
EXPECT(GEMV, IDX(i), IDX(i, j), IDX(j));
EXPECT(GENERIC, IDX(i), IDX(i, j, k), IDX(j, i, k));
EXPECT(GENERIC, IDX(i, k), IDX(i, j, k), IDX(j, i, k));
EXPECT(GENERIC, IDX(i, l), IDX(i, j, k), IDX(j, i, k));
// Einsum4.cpp:46,71 / Einsum5.cpp:28,53
EXPECT(GEMM, IDX(i, j, k, l), IDX(i, j, k, p), IDX(p, l));
EXPECT(GENERIC, IDX(i, j, k, l), IDX(i, j, p, l), IDX(p, k));
EXPECT(GEMM, IDX(p, q, r, l), IDX(p, q, r, s), IDX(s, l));
EXPECT(GENERIC, IDX(p, q, k, s), IDX(p, q, r, s), IDX(r, k));
// Einsum7.cpp:26,37,47,66,96,125,169,181,192
EXPECT(GENERIC, IDX(i, a, W), IDX(i, W), IDX(a, W));
EXPECT(GEMM, IDX(i, a, P), IDX(i, a, W), IDX(P, W));
EXPECT(DIRECT, IDX(p, q), IDX(p, q), IDX(p, q));
EXPECT(GEMM, IDX(Q, X), IDX(Q, i, a), IDX(i, a, X));
EXPECT(GENERIC, IDX(Q, a, X), IDX(Q, Y, X), IDX(a, Y));
EXPECT(GER, IDX(a, b), IDX(a), IDX(b));
EXPECT(GENERIC, IDX(a, X), IDX(Q, a, X), IDX(Q, X));
EXPECT(GEMV, IDX(X), IDX(i, a, X), IDX(i, a));
// Element.cpp:120-176
EXPECT(DIRECT, IDX(i, j), IDX(i, j), IDX(i, j));
// QChemTests.cpp:301-304 (Fock J: swapped-operand GEMV; Fock K: GENERIC), config 1 of the benchmark
EXPECT(GEMV, IDX(i, j), IDX(k, l), IDX(i, j, k, l));
EXPECT(GENERIC, IDX(i, j), IDX(k, l), IDX(i, k, j, l));
// benchmark configs 2, 3, 5
EXPECT(GEMM, IDX(m, n, i, j), IDX(m, n, e, f), IDX(i, j, e, f));     // Wmnij
EXPECT(GEMM, IDX(a, b, e, f), IDX(m, n, e, f), IDX(m, n, a, b));     // Wabef (C transposed variant)
EXPECT(GENERIC, IDX(m, b, e, j), IDX(j, n, f, b), IDX(m, n, e, f)); // Wmbej
// transposition variants, Transpose.cpp:17-125
EXPECT(GEMM, IDX(j, i), IDX(i, k), IDX(k, j));
EXPECT(GEMM, IDX(i, j), IDX(k, i), IDX(k, j));
EXPECT(GEMM, IDX(i, j), IDX(k, i), IDX(j, k));
// conjugation rules as written in the reference (Dispatch.hpp:220-231, 603-605)
static_assert(choose<true, false>(Indices{i, j}, Indices{i, j}, Indices{i, j}) != DIRECT);
static_assert(choose<true, false>(Indices{i, j}, Indices{i, k}, Indices{k, j}) == GEMM);
static_assert(choose<true, false>(Indices{i, j}, Indices{k, i}, Indices{k, j}) == GENERIC);
// mixed value types and Hadamard indices always take the generic route (:896-899)
static_assert(choose<false, false, false>(Indices{i, j}, Indices{i, k}, Indices{k, j}) == GENERIC);
static_assert(choose(Indices{i}, Indices{i, i}, Indices{i}) == GENERIC);

// index utilities (TensorBase/IndexUtilities.hpp): strides <-> running index, index tables, extents resolved C -> A -> B
namespace index_utils_checks {
constexpr bool strides_ok() {
    std::array<int64_t, 3> d{4, 5, 6}, s{}, sc{};
    size_t n = einsums::dims_to_strides(d, s), nc = einsums::dims_to_strides(d, sc, false);
    std::array<int64_t, 3> ix{};
    einsums::sentinel_to_indices(size_t(2 * 30 + 3 * 6 + 4), s, ix);
    bool ok = n == 120 && nc == 120 && s[0] == 30 && s[1] == 6 && s[2] == 1 && sc[0] == 1 && sc[1] == 4 && sc[2] == 20;
    ok = ok && ix[0] == 2 && ix[1] == 3 && ix[2] == 4 && einsums::indices_to_sentinel(s, ix) == 82;
    einsums::sentinel_to_indices(size_t(2 + 3 * 4 + 4 * 20), sc, ix);
    return ok && ix[0] == 2 && ix[1] == 3 && ix[2] == 4;
}
static_assert(strides_ok());
constexpr bool table_ok() {
    using namespace einsums::index;
    std::array<int, 3> t{};
    einsums::compile_index_table(std::tuple{i, j, k, l}, std::tuple{k, i, l}, t);
    return t[0] == 2 && t[1] == 0 && t[2] == 3;
}
static_assert(table_ok());
struct FakeTensor {
    std::array<int64_t, 2> d;
    int64_t dim(std::ptrdiff_t a) const { return d[size_t(a)]; }
};
inline bool ranges_ok() {
    using namespace einsums::index;
    FakeTensor C{{3, 4}}, A{{3, 7}}, B{{7, 4}};
    auto r = einsums::get_dim_ranges_for_many(C, std::tuple{i, j}, A, std::tuple{i, k}, B, std::tuple{k, j}, std::tuple{i, j, k});
    return r[0] == 3 && r[1] == 4 && r[2] == 7;
}
} // namespace index_utils_checks

int main() {
    if (!index_utils_checks::ranges_ok()) {
        std::printf("index utilities: get_dim_ranges_for_many FAILED\n");
        return 1;
    }
    // print the decisions so that the Python side can diff them against tests/golden/algorithm_choice.json
    const char *names[] = {"GENERIC", "DOT", "DIRECT", "GER", "GEMV", "GEMM", "INDETERMINATE"};
#define SHOW(C_, A_, B_) std::printf("%s <- %s , %s : %s\n", #C_, #A_, #B_, names[choose(C_, A_, B_)])
    SHOW(IDX(i, j), IDX(i, k), IDX(k, j));
    SHOW(IDX(i, l), IDX(i, j, k), IDX(j, k, l));
    SHOW(IDX(a, b, e, f), IDX(m, n, e, f), IDX(m, n, a, b));
    SHOW(IDX(m, b, e, j), IDX(j, n, f, b), IDX(m, n, e, f));
    std::printf("analysis static_asserts: all passed at compile time\n");
    return 0;
}

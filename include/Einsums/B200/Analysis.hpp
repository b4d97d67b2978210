// Analysis.hpp — the compile-time einsum analyser, written as ordinary constexpr code over arrays of index ids.
// It reproduces, decision for decision, the reference's if-chain DOT -> DIRECT -> GER -> GEMV -> GEMV(swapped) ->
// GEMM -> GENERIC (reference: TensorAlgebra/Backends/Dispatch.hpp:859-957) and its predicates
// (:192-262 hadamard/dot/direct/outer, :407-439 matrix-vector, :565-609 matrix-matrix), including the
// index-pack algebra of TensorAlgebra/Detail/Utilities.hpp:63-152,289-343,609-751 (Intersect/Difference/Unique keep
// first-pack order; find_type_with_position lists every occurrence).
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>

namespace einsums::tensor_algebra::detail {

/// reference: TensorAlgebra/Detail/Utilities.hpp:26-34
enum AlgorithmChoice { GENERIC, DOT, DIRECT, GER, GEMV, GEMM, INDETERMINATE };

inline constexpr int kMaxPack = 16;

struct IdPack {
    int n = 0;
    int v[kMaxPack] = {};
    constexpr bool has(int x) const {
        for (int i = 0; i < n; i++)
            if (v[i] == x) return true;
        return false;
    }
    constexpr void push(int x) { v[n++] = x; }
    constexpr bool operator==(IdPack const &o) const {
        if (n != o.n) return false;
        for (int i = 0; i < n; i++)
            if (v[i] != o.v[i]) return false;
        return true;
    }
};
template <size_t N> constexpr IdPack make_pack(std::array<int, N> const &a) {
    static_assert(N <= size_t(kMaxPack), "too many indices on one tensor");
    IdPack p;
    for (size_t i = 0; i < N; i++) p.push(a[i]);
    return p;
}
constexpr IdPack intersect(IdPack const &a, IdPack const &b) {
    IdPack r;
    for (int i = 0; i < a.n; i++)
        if (b.has(a.v[i])) r.push(a.v[i]);
    return r;
}
constexpr IdPack difference(IdPack const &a, IdPack const &b) {
    IdPack r;
    for (int i = 0; i < a.n; i++)
        if (!b.has(a.v[i])) r.push(a.v[i]);
    return r;
}
constexpr IdPack unique(IdPack const &a) {
    IdPack r;
    for (int i = 0; i < a.n; i++)
        if (!r.has(a.v[i])) r.push(a.v[i]);
    return r;
}

/// (label, position) list: for every label of `ts` in order, every position in `us` that carries it
struct PosList {
    int n = 0;
    int label[4 * kMaxPack] = {};
    int pos[4 * kMaxPack] = {};
    constexpr bool contiguous() const {
        for (int i = 0; i + 1 < n; i++)
            if (pos[i] != pos[i + 1] - 1) return false;
        return true;
    }
    constexpr int last() const { return pos[n - 1]; }
};
constexpr PosList find_positions(IdPack const &ts, IdPack const &us) {
    PosList r;
    for (int i = 0; i < ts.n; i++)
        for (int p = 0; p < us.n; p++)
            if (us.v[p] == ts.v[i]) {
                r.label[r.n] = ts.v[i];
                r.pos[r.n] = p;
                r.n++;
            }
    return r;
}
constexpr bool same_ordering(PosList const &x, PosList const &y) {
    if (x.n == 0 || y.n == 0 || x.n != y.n) return false;
    for (int i = 0; i < x.n; i++)
        if (x.label[i] != y.label[i]) return false;
    return true;
}

/// Everything einsum_do_* needs at run time, computed at compile time.
struct Analysis {
    IdPack  C, A, B, links_ab, links, c_unique, link_unique, c_minus_a, c_minus_b, all_unique;
    PosList link_in_a, link_in_b, target_in_a, target_in_b, a_target_in_c, b_target_in_c;
    bool    hadamard = false;
};
constexpr Analysis analyse(IdPack const &C, IdPack const &A, IdPack const &B) {
    Analysis z;
    z.C = C; z.A = A; z.B = B;
    z.links_ab = intersect(A, B);
    z.links = difference(z.links_ab, C);
    z.c_unique = unique(C);
    z.link_unique = unique(z.links);
    z.c_minus_a = difference(C, A);
    z.c_minus_b = difference(C, B);
    z.link_in_a = find_positions(z.link_unique, A);
    z.link_in_b = find_positions(z.link_unique, B);
    z.target_in_a = find_positions(z.c_unique, A);
    z.target_in_b = find_positions(z.c_unique, B);
    z.a_target_in_c = find_positions(A, C);
    z.b_target_in_c = find_positions(B, C);
    z.hadamard = unique(C).n != C.n || unique(A).n != A.n || unique(B).n != B.n;
    IdPack all;
    for (int i = 0; i < C.n; i++) all.push(C.v[i]);
    IdPack ua = unique(A), ub = unique(B);
    z.all_unique = unique(all);
    for (int i = 0; i < ua.n; i++)
        if (!z.all_unique.has(ua.v[i])) z.all_unique.push(ua.v[i]);
    for (int i = 0; i < ub.n; i++)
        if (!z.all_unique.has(ub.v[i])) z.all_unique.push(ub.v[i]);
    return z;
}

constexpr bool is_outer(Analysis const &z, bool conj_a, bool conj_b) {
    if (!(z.links_ab.n == 0 && z.target_in_a.contiguous() && z.target_in_b.contiguous())) return false;
    if (z.a_target_in_c.n == 0) return false;
    bool swap_ab = z.a_target_in_c.pos[0] != 0;
    return (!conj_a && !swap_ab) || (!conj_b && swap_ab);
}
/// A pack = matrix, B pack = vector
constexpr bool is_matrix_vector(Analysis const &z, bool conj_a, bool conj_b) {
    bool cond = z.link_in_a.contiguous() && z.link_in_b.contiguous() && z.target_in_a.contiguous() && same_ordering(z.link_in_a, z.link_in_b) &&
                same_ordering(z.target_in_a, z.a_target_in_c) && !same_ordering(z.target_in_b, z.b_target_in_c) && z.b_target_in_c.n == 0 &&
                !conj_b;
    if (!cond) return false;
    bool transpose_a = z.link_in_a.pos[0] == 0;
    return transpose_a || !conj_a;
}
constexpr bool is_matrix_product(Analysis const &z, bool conj_a, bool conj_b) {
    bool cond = z.c_minus_a.n > 0 && z.c_minus_b.n > 0 && z.link_in_a.contiguous() && z.link_in_b.contiguous() &&
                z.target_in_a.contiguous() && z.target_in_b.contiguous() && z.a_target_in_c.contiguous() && z.b_target_in_c.contiguous() &&
                same_ordering(z.link_in_a, z.link_in_b) && same_ordering(z.target_in_a, z.a_target_in_c) &&
                same_ordering(z.target_in_b, z.b_target_in_c);
    if (!cond) return false;
    bool tA = z.link_in_a.pos[0] == 0, tB = z.link_in_b.pos[0] != 0, tC = z.a_target_in_c.pos[0] != 0;
    return (tC && (tA || !conj_a) && (tB || !conj_b)) || (!tC && (!tA || !conj_a) && (!tB || !conj_b));
}

enum class Route { Generic, Dot, Direct, Ger, GemvAB, GemvBA, Gemm };

/// `same_types`: C, A and B have the same value type; `ranks_ok`: all ranks >= 2 (GEMM precondition, Dispatch.hpp:945)
constexpr Route static_route(IdPack const &C, IdPack const &A, IdPack const &B, bool conj_a, bool conj_b, bool same_types) {
    Analysis z = analyse(C, A, B);
    if (z.hadamard || !same_types) return Route::Generic;
    if (C.n == 0 && A == B) return Route::Dot;
    if (C == A && C == B && !conj_a && !conj_b) return Route::Direct;
    if (is_outer(z, conj_a, conj_b)) return Route::Ger;
    if (is_matrix_vector(z, conj_a, conj_b)) return Route::GemvAB;
    if (is_matrix_vector(analyse(C, B, A), conj_a, conj_b)) return Route::GemvBA; // reference passes ConjA, ConjB unswapped (Dispatch.hpp:939)
    if (C.n >= 2 && A.n >= 2 && B.n >= 2 && is_matrix_product(z, conj_a, conj_b)) return Route::Gemm;
    return Route::Generic;
}

constexpr AlgorithmChoice choice_of(Route r) {
    switch (r) {
    case Route::Dot: return DOT;
    case Route::Direct: return DIRECT;
    case Route::Ger: return GER;
    case Route::GemvAB:
    case Route::GemvBA: return GEMV;
    case Route::Gemm: return GEMM;
    default: return GENERIC;
    }
}

} // namespace einsums::tensor_algebra::detail

// Index.hpp — index tags and Indices{...}, source-compatible with the reference's
// TensorAlgebra/Detail/Index.hpp:32-66,69-173,199-206: einsums::index::{a..z, A..Z, greek}. Each tag is an empty
// type; here it also carries a compile-time integer id so that the analyser can run as ordinary constexpr code
// over arrays of ids instead of tuple metaprogramming.
#pragma once
#include <array>
#include <cstddef>
#include <ostream>
#include <tuple>
#include <type_traits>

namespace einsums {
namespace index {
struct LabelBase {};
} // namespace index

#define EINSUMS_B200_MAKE_INDEX(x, ID)                                                      \
    namespace index {                                                                       \
    struct x : public LabelBase {                                                           \
        static constexpr const char *letter = #x;                                           \
        static constexpr int         id = ID;                                               \
        constexpr x() = default;                                                            \
        constexpr size_t operator()(size_t index) const { return index; }                   \
    };                                                                                      \
    static constexpr struct x x;                                                            \
    inline std::ostream &operator<<(std::ostream &os, const struct x &) { return os << x::letter; } \
    }

#define EB_IDX2(lo, up, n) EINSUMS_B200_MAKE_INDEX(up, 2 * n) EINSUMS_B200_MAKE_INDEX(lo, 2 * n + 1)
EB_IDX2(a, A, 0) EB_IDX2(b, B, 1) EB_IDX2(c, C, 2) EB_IDX2(d, D, 3) EB_IDX2(e, E, 4) EB_IDX2(f, F, 5) EB_IDX2(g, G, 6)
EB_IDX2(h, H, 7) EB_IDX2(i, I, 8) EB_IDX2(j, J, 9) EB_IDX2(k, K, 10) EB_IDX2(l, L, 11) EB_IDX2(m, M, 12) EB_IDX2(n, N, 13)
EB_IDX2(o, O, 14) EB_IDX2(p, P, 15) EB_IDX2(q, Q, 16) EB_IDX2(r, R, 17) EB_IDX2(s, S, 18) EB_IDX2(t, T, 19) EB_IDX2(u, U, 20)
EB_IDX2(v, V, 21) EB_IDX2(w, W, 22) EB_IDX2(x, X, 23) EB_IDX2(y, Y, 24) EB_IDX2(z, Z, 25)
EB_IDX2(alpha, Alpha, 26) EB_IDX2(beta, Beta, 27) EB_IDX2(gamma, Gamma, 28) EB_IDX2(delta, Delta, 29) EB_IDX2(epsilon, Epsilon, 30)
EB_IDX2(zeta, Zeta, 31) EB_IDX2(eta, Eta, 32) EB_IDX2(theta, Theta, 33) EB_IDX2(iota, Iota, 34) EB_IDX2(kappa, Kappa, 35)
EB_IDX2(lambda, Lambda, 36) EB_IDX2(mu, Mu, 37) EB_IDX2(nu, Nu, 38) EB_IDX2(xi, Xi, 39) EB_IDX2(omicron, Omicron, 40)
EB_IDX2(pi, Pi, 41) EB_IDX2(rho, Rho, 42) EB_IDX2(sigma, Sigma, 43) EB_IDX2(tau, Tau, 44) EB_IDX2(upsilon, Upsilon, 45)
EB_IDX2(phi, Phi, 46) EB_IDX2(chi, Chi, 47) EB_IDX2(psi, Psi, 48) EB_IDX2(omega, Omega, 49)
#undef EB_IDX2

namespace tensor_algebra {
/// Indices{i, j, k}: a std::tuple of index tags (reference: Detail/Index.hpp:199-206)
template <typename... Args>
    requires(std::is_base_of_v<index::LabelBase, Args> && ... && true)
struct Indices : std::tuple<Args...> {
    constexpr Indices(Args... args) : std::tuple<Args...>(args...) {}
};
template <typename... Args> Indices(Args...) -> Indices<Args...>;

namespace detail {
template <typename... Is> constexpr std::array<int, sizeof...(Is)> ids_of(std::tuple<Is...> const *) {
    return std::array<int, sizeof...(Is)>{std::remove_cvref_t<Is>::id...};
}
template <typename Tuple> constexpr auto ids_of() { return ids_of(static_cast<Tuple const *>(nullptr)); }
} // namespace detail
} // namespace tensor_algebra

using tensor_algebra::Indices;
} // namespace einsums

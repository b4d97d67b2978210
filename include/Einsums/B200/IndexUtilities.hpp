// IndexUtilities.hpp — host-side index helpers with the reference's names
// (libs/Einsums/TensorBase/include/Einsums/TensorBase/IndexUtilities.hpp:24-120 get_dim_ranges_for_many, :154-180
//  sentinel_to_indices, :502-515 indices_to_sentinel, :698-760 dims_to_strides, :882-932 compile_index_table).
// The device path itself never needs them — dims, strides and index tables travel to eb200_contract by value and the kernels decode
// composite indices with their own stride tables — but callers of the typed seam (and the tiled algebra) use them on the host.
#pragma once
#include <Einsums/B200/Index.hpp>

#include <array>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace einsums {

/// strides for walking `dims` with one running index; returns the total size
template <typename I1, typename I2, size_t N>
    requires(std::is_integral_v<I1> && std::is_integral_v<I2>)
constexpr size_t dims_to_strides(std::array<I1, N> const &dims, std::array<I2, N> &out, bool row_major = true) {
    size_t stride = 1;
    for (size_t n = 0; n < N; n++) {
        size_t i = row_major ? N - 1 - n : n;
        out[i] = static_cast<I2>(stride);
        stride *= static_cast<size_t>(dims[i]);
    }
    return stride;
}
template <typename I1, typename I2> size_t dims_to_strides(std::vector<I1> const &dims, std::vector<I2> &out, bool row_major = true) {
    size_t stride = 1;
    out.resize(dims.size());
    for (size_t n = 0; n < dims.size(); n++) {
        size_t i = row_major ? dims.size() - 1 - n : n;
        out[i] = static_cast<I2>(stride);
        stride *= static_cast<size_t>(dims[i]);
    }
    return stride;
}

/// running index -> one index per axis (largest stride first; zero strides give index 0)
template <typename I1, typename I2, size_t N>
constexpr void sentinel_to_indices(size_t sentinel, std::array<I1, N> const &strides, std::array<I2, N> &out) {
    if constexpr (N == 0) return;
    const bool row_major = N < 2 || strides[0] >= strides[N - 1];
    size_t     hold = sentinel;
    for (size_t n = 0; n < N; n++) {
        size_t i = row_major ? n : N - 1 - n;
        if (strides[i] != 0) {
            out[i] = static_cast<I2>(hold / static_cast<size_t>(strides[i]));
            hold %= static_cast<size_t>(strides[i]);
        } else {
            out[i] = 0;
        }
    }
}
/// the inverse: sum of index * stride
template <typename I1, typename I2, size_t N> constexpr size_t indices_to_sentinel(std::array<I1, N> const &strides, std::array<I2, N> const &inds) {
    size_t out = 0;
    for (size_t i = 0; i < N; i++) out += static_cast<size_t>(inds[i]) * static_cast<size_t>(strides[i]);
    return out;
}

/// position of every index of `indices` inside `unique_indices` — what eb200_tensor_desc::index carries
template <typename... Unique, typename... Idx>
constexpr void compile_index_table(std::tuple<Unique...> const &, std::tuple<Idx...> const &, std::array<int, sizeof...(Idx)> &out) {
    constexpr auto u = tensor_algebra::detail::ids_of<std::tuple<Unique...>>();
    constexpr auto x = tensor_algebra::detail::ids_of<std::tuple<Idx...>>();
    for (size_t p = 0; p < x.size(); p++) {
        out[p] = -1;
        for (size_t q = 0; q < u.size(); q++)
            if (u[q] == x[p]) { out[p] = int(q); break; }
    }
}

/// extent of every unique index, taken from the first tensor that carries it: C, then A, then B
template <typename CType, typename AType, typename BType, typename... CIdx, typename... AIdx, typename... BIdx, typename... Unique>
auto get_dim_ranges_for_many(CType const &C, std::tuple<CIdx...> const &, AType const &A, std::tuple<AIdx...> const &, BType const &B,
                             std::tuple<BIdx...> const &, std::tuple<Unique...> const &) {
    constexpr auto u = tensor_algebra::detail::ids_of<std::tuple<Unique...>>();
    constexpr auto c = tensor_algebra::detail::ids_of<std::tuple<CIdx...>>();
    constexpr auto a = tensor_algebra::detail::ids_of<std::tuple<AIdx...>>();
    constexpr auto b = tensor_algebra::detail::ids_of<std::tuple<BIdx...>>();
    std::array<size_t, sizeof...(Unique)> out{};
    for (size_t q = 0; q < u.size(); q++) {
        bool found = false;
        if constexpr (sizeof...(CIdx) > 0)
            for (size_t p = 0; p < c.size() && !found; p++)
                if (c[p] == u[q]) { out[q] = size_t(C.dim(std::ptrdiff_t(p))); found = true; }
        for (size_t p = 0; p < a.size() && !found; p++)
            if (a[p] == u[q]) { out[q] = size_t(A.dim(std::ptrdiff_t(p))); found = true; }
        for (size_t p = 0; p < b.size() && !found; p++)
            if (b[p] == u[q]) { out[q] = size_t(B.dim(std::ptrdiff_t(p))); found = true; }
        if (!found) throw std::invalid_argument("get_dim_ranges_for_many: index not carried by any tensor");
    }
    return out;
}

} // namespace einsums

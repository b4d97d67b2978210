// BlockTileAlgebra.hpp — einsum over a MIX of block-diagonal and tiled device tensors, and the tiled forms of ger / gerc /
// direct_product. Reference: TensorAlgebra/Backends/BlockTileAlgebra.hpp:18-174 (einsum_special_dispatch over the unique tile grid: a
// block tensor is a tiled tensor whose only present tiles are the diagonal ones), LinearAlgebra/TiledTensor.hpp:356-430.
//
// As for the pure tiled / pure block cases the whole operation is ONE eb200_contract_grouped call: descriptors for all present
// (block, tile) combinations are built on the host and enqueued without a wait in between; equal-size neighbours fold into batched /
// longer-K launches (contract.cu). The reference runs one einsum + host wait per combination under a lock on C.
#pragma once
#include <Einsums/B200/BlockTensor.hpp>
#include <Einsums/B200/TiledTensor.hpp>

namespace einsums {
namespace tensor_algebra {
namespace detail {

template <typename X>
concept GridTensorLike = TiledDeviceTensorLike<X> || BlockDeviceTensorLike<X>;

/// number of grid positions along `axis`: tiles along that axis, or blocks
template <typename X> std::int64_t grid_extent(X const &x, size_t axis) {
    if constexpr (TiledDeviceTensorLike<X>) return x.grid_size(axis);
    else return std::int64_t(x.num_blocks());
}
/// The part of x at grid position idx, or nullptr when that position holds nothing: an off-diagonal position of a block tensor, an empty
/// block, an absent or zero-size tile (reference BlockTileAlgebra.hpp:68-124).
template <typename X, size_t R> auto grid_part(X const &x, std::array<int, R> const &idx) -> DeviceTensorView<typename X::ValueType, R> const * {
    if constexpr (TiledDeviceTensorLike<X>) {
        auto const *t = x.find_tile(idx);
        return (t && t->size()) ? t : nullptr;
    } else {
        for (size_t a = 1; a < R; a++)
            if (idx[a] != idx[0]) return nullptr;
        if (x.block_dim(size_t(idx[0])) == 0) return nullptr;
        return &x.block(size_t(idx[0]));
    }
}
/// The part of the RESULT at idx: tiles are created (zero-filled) on demand, off-diagonal positions of a block result do not exist.
template <typename X, size_t R> auto grid_part_mut(X *x, std::array<int, R> const &idx) -> DeviceTensorView<typename X::ValueType, R> * {
    if constexpr (TiledDeviceTensorLike<X>) {
        if (x->has_zero_size(idx)) return nullptr;
        return &x->tile(idx);
    } else {
        for (size_t a = 1; a < R; a++)
            if (idx[a] != idx[0]) return nullptr;
        if (x->block_dim(size_t(idx[0])) == 0) return nullptr;
        return &x->block(size_t(idx[0]));
    }
}

} // namespace detail

/// einsum with at least one block-diagonal and at least one tiled tensor among C, A, B.
template <bool ConjA = false, bool ConjB = false, detail::GridTensorLike AType, detail::GridTensorLike BType, detail::GridTensorLike CType, typename U,
          typename... CIndices, typename... AIndices, typename... BIndices>
    requires((std::is_arithmetic_v<U> || IsComplexV<U>) && !(TiledDeviceTensorLike<AType> && TiledDeviceTensorLike<BType> && TiledDeviceTensorLike<CType>) &&
             !(BlockDeviceTensorLike<AType> && BlockDeviceTensorLike<BType> && BlockDeviceTensorLike<CType>))
void einsum(U const UC_prefactor, std::tuple<CIndices...> const &, CType *C, U const UAB_prefactor, std::tuple<AIndices...> const &, AType const &A,
            std::tuple<BIndices...> const &, BType const &B, detail::AlgorithmChoice *algorithm_choice = nullptr) {
    using namespace detail;
    using T = typename CType::ValueType;
    static_assert(std::is_same_v<T, typename AType::ValueType> && std::is_same_v<T, typename BType::ValueType>, "mixed block/tiled einsum needs one value type");
    constexpr size_t CR = CType::Rank, AR = AType::Rank, BR = BType::Rank;
    static_assert(sizeof...(CIndices) == CR && sizeof...(AIndices) == AR && sizeof...(BIndices) == BR, "ranks and index packs must agree");
    constexpr IdPack   pc = make_pack(ids_of<std::tuple<CIndices...>>()), pa = make_pack(ids_of<std::tuple<AIndices...>>()),
                     pb = make_pack(ids_of<std::tuple<BIndices...>>());
    constexpr Analysis z = analyse(pc, pa, pb);
    TileTables<CR, AR, BR> tab;
    tab.n_unique = z.all_unique.n;
    fill_table<void>(z.all_unique, z.C, tab.c);
    fill_table<void>(z.all_unique, z.A, tab.a);
    fill_table<void>(z.all_unique, z.B, tab.b);

    // grid extent of every unique index, resolved C -> A -> B (reference: get_grid_ranges_for_many)
    std::vector<std::int64_t> grid(size_t(tab.n_unique), -1);
    auto put = [&](int u, std::int64_t e) {
        if (grid[size_t(u)] < 0) grid[size_t(u)] = e;
        else if (grid[size_t(u)] != e) throw tensor_compat_error("Block and tiled tensors do not have compatible grids.");
    };
    for (size_t p = 0; p < CR; p++) put(tab.c[p], grid_extent(*C, p));
    for (size_t p = 0; p < AR; p++) put(tab.a[p], grid_extent(A, p));
    for (size_t p = 0; p < BR; p++) put(tab.b[p], grid_extent(B, p));

    T const beta = T(UC_prefactor), alpha = T(UAB_prefactor);
    if (beta == T(0)) {
        if constexpr (TiledDeviceTensorLike<CType>) C->zero_no_clear();
        else C->zero();
    } else if (beta != T(1)) *C *= beta;

    std::vector<eb200_contract_desc> descs;
    std::int64_t                     total = 1;
    for (auto g : grid) total *= g;
    std::vector<int> u(size_t(tab.n_unique), 0);
    for (std::int64_t s = 0; s < total; s++) {
        std::int64_t x = s;
        for (int k = tab.n_unique - 1; k >= 0; k--) {
            u[size_t(k)] = int(x % grid[size_t(k)]);
            x /= grid[size_t(k)];
        }
        std::array<int, AR> ai; std::array<int, BR> bi; std::array<int, CR> ci;
        for (size_t p = 0; p < AR; p++) ai[p] = u[size_t(tab.a[p])];
        for (size_t p = 0; p < BR; p++) bi[p] = u[size_t(tab.b[p])];
        for (size_t p = 0; p < CR; p++) ci[p] = u[size_t(tab.c[p])];
        auto const *at = grid_part(A, ai);
        if (!at) continue;
        auto const *bt = grid_part(B, bi);
        if (!bt) continue;
        auto *ct = grid_part_mut(C, ci);
        if (!ct) continue;
        descs.emplace_back();
        tile_desc_into<T>(descs.back(), tab, T(1), ct->gpu_data(), ct->dims(), ct->strides(), alpha, *at, *bt, ConjA, ConjB);
    }
    if (!descs.empty()) EB200_CHECK(eb200_contract_grouped(gpu::get_stream(), std::int64_t(descs.size()), descs.data(), nullptr));
    if (algorithm_choice) *algorithm_choice = detail::GENERIC;
    gpu::op_wait();
}
template <bool ConjA = false, bool ConjB = false, detail::GridTensorLike AType, detail::GridTensorLike BType, detail::GridTensorLike CType,
          typename... CIndices, typename... AIndices, typename... BIndices>
    requires(!(TiledDeviceTensorLike<AType> && TiledDeviceTensorLike<BType> && TiledDeviceTensorLike<CType>) &&
             !(BlockDeviceTensorLike<AType> && BlockDeviceTensorLike<BType> && BlockDeviceTensorLike<CType>))
void einsum(std::tuple<CIndices...> const &C_indices, CType *C, std::tuple<AIndices...> const &A_indices, AType const &A,
            std::tuple<BIndices...> const &B_indices, BType const &B, detail::AlgorithmChoice *algorithm_choice = nullptr) {
    using V = typename CType::ValueType;
    using S = std::conditional_t<IsComplexV<V>, V, double>;
    einsum<ConjA, ConjB>(S(0), C_indices, C, S(1), A_indices, A, B_indices, B, algorithm_choice);
}

} // namespace tensor_algebra

namespace linear_algebra {
namespace detail {
/// A(i,j) += alpha * x_i * op(y_j) over every present pair of vector tiles; op = conj for gerc
template <typename T> void tiled_ger(T alpha, TiledDeviceTensor<T, 1> const &X, TiledDeviceTensor<T, 1> const &Y, TiledDeviceTensor<T, 2> *A, bool conj_y) {
    // the reference's check uses && (LinearAlgebra/TiledTensor.hpp:387); a mismatch in either direction cannot work, so both are required here
    if (A->grid_size(0) != X.grid_size(0) || A->grid_size(1) != Y.grid_size(0)) throw tensor_compat_error("Tiled tensors have incompatible grids!");
    tensor_algebra::detail::TileTables<2, 1, 1> tab;
    tab.n_unique = 2; tab.c = {0, 1}; tab.a = {0}; tab.b = {1};
    std::vector<eb200_contract_desc> descs;
    for (int i = 0; i < int(X.grid_size(0)); i++)
        for (int j = 0; j < int(Y.grid_size(0)); j++) {
            auto const *xt = X.find_tile({i});
            auto const *yt = Y.find_tile({j});
            if (!xt || !yt || !xt->size() || !yt->size()) continue;
            auto &at = A->tile({i, j});
            descs.emplace_back();
            tensor_algebra::detail::tile_desc_into<T>(descs.back(), tab, T(1), at.gpu_data(), at.dims(), at.strides(), alpha, *xt, *yt, false, conj_y);
        }
    if (!descs.empty()) EB200_CHECK(eb200_contract_grouped(gpu::get_stream(), std::int64_t(descs.size()), descs.data(), nullptr));
    gpu::op_wait();
}
} // namespace detail

/// A += alpha * X * Y^T on tiled tensors (reference: LinearAlgebra/TiledTensor.hpp:380-404)
template <typename T, typename U> void ger(U alpha, TiledDeviceTensor<T, 1> const &X, TiledDeviceTensor<T, 1> const &Y, TiledDeviceTensor<T, 2> *A) {
    detail::tiled_ger<T>(T(alpha), X, Y, A, false);
}
/// A += alpha * X * Y^H (reference :406-430)
template <typename T, typename U> void gerc(U alpha, TiledDeviceTensor<T, 1> const &X, TiledDeviceTensor<T, 1> const &Y, TiledDeviceTensor<T, 2> *A) {
    detail::tiled_ger<T>(T(alpha), X, Y, A, true);
}

/// C = beta * C + alpha * A o B tile by tile, for the tiles both A and B hold (reference :356-378; tiles of C at other positions are untouched)
template <typename T, size_t R, typename U>
void direct_product(U alpha, TiledDeviceTensor<T, R> const &A, TiledDeviceTensor<T, R> const &B, U beta, TiledDeviceTensor<T, R> *C) {
    for (size_t a = 0; a < R; a++)
        if (A.grid_size(a) != B.grid_size(a) || A.grid_size(a) != C->grid_size(a)) throw tensor_compat_error("Tiled tensors have incompatible grids!");
    tensor_algebra::detail::TileTables<R, R, R> tab;
    tab.n_unique = int(R);
    for (size_t p = 0; p < R; p++) tab.c[p] = tab.a[p] = tab.b[p] = int(p);
    std::vector<eb200_contract_desc> descs;
    for (auto const &[idx, at] : A.tiles()) {
        auto const *bt = B.find_tile(idx);
        if (!bt || !at.size() || !bt->size()) continue;
        auto &ct = C->tile(idx);
        descs.emplace_back();
        tensor_algebra::detail::tile_desc_into<T>(descs.back(), tab, T(beta), ct.gpu_data(), ct.dims(), ct.strides(), T(alpha), at, *bt, false, false);
    }
    if (!descs.empty()) EB200_CHECK(eb200_contract_grouped(gpu::get_stream(), std::int64_t(descs.size()), descs.data(), nullptr));
    gpu::op_wait();
}

} // namespace linear_algebra

using tensor_algebra::einsum;
} // namespace einsums

// TiledTensor.hpp — TiledDeviceTensor<T, Rank>: a tensor cut into a grid of tiles of which only some exist (absent tiles are
// zero), plus einsum / linear algebra over it. Same names and call syntax as the reference
// (libs/Einsums/Tensor/include/Einsums/Tensor/TiledTensor.hpp:60-1882; algebra: TensorAlgebra/Backends/TileAlgebra.hpp:17-158,
//  LinearAlgebra/TiledTensor.hpp:26-440; tests Tensor/tests/unit/DeviceTiledTensor.hip).
//
// B200-first: tiles are carved out of a growing arena (a few large device allocations) instead of one allocation per
// tile, so equal-size tiles created one after another sit at a constant distance and eb200_contract_grouped can fold them
// level by level: the accumulation over tile columns becomes one contraction with a longer K, rows of output tiles become
// batches, and a full grid of equal tiles collapses into a single launch. A tiled contraction is ONE grouped call: all
// tile triples in the reference's enumeration order (last unique index fastest, so the contributions to one output tile
// are neighbours); what does not fold overlaps on auxiliary streams, contributions to the same tile staying in order.
// The reference runs one einsum + wait per tile triple under a per-tile lock.
#pragma once
#include <Einsums/B200/LinearAlgebra.hpp>
#include <Einsums/B200/TensorAlgebra.hpp>

#include <algorithm>
#include <map>
#include <memory>
#include <numeric>

namespace einsums {

namespace detail {
/// bump allocator over a few large device buffers; everything is released together
class DeviceArena {
  public:
    DeviceArena() = default;
    DeviceArena(DeviceArena const &) = delete;
    DeviceArena &operator=(DeviceArena const &) = delete;
    ~DeviceArena() {
        for (auto &c : _chunks) eb200_free(c.base);
    }
    /// chunk size: what the whole tensor would need if every tile existed, capped — so that a (nearly) full grid is one
    /// contiguous run of tiles and an almost empty one wastes at most one chunk
    explicit DeviceArena(size_t dense_bytes) : _chunk(std::clamp<size_t>(dense_bytes, size_t(4) << 10, size_t(512) << 20)) {}
    void *take(size_t bytes) {
        bytes = (bytes + 255) / 256 * 256;
        if (_chunks.empty() || _chunks.back().used + bytes > _chunks.back().cap) {
            size_t cap = std::max<size_t>(bytes, _chunk);
            void *p = nullptr;
            (void)gpu::get_stream();
            EB200_CHECK(eb200_malloc(&p, cap));
            _chunks.push_back({static_cast<char *>(p), cap, 0});
        }
        auto &c = _chunks.back();
        void *out = c.base + c.used;
        c.used += bytes;
        return out;
    }

  private:
    struct Chunk { char *base; size_t cap, used; };
    std::vector<Chunk> _chunks;
    size_t             _chunk = size_t(1) << 20;
};
} // namespace detail

template <typename T, size_t R> class TiledDeviceTensor {
  public:
    using ValueType = T;
    using Tile      = DeviceTensorView<T, R>;
    using TileIndex = std::array<int, R>;
    static constexpr size_t Rank = R;

    TiledDeviceTensor() = default;
    /// the same tile sizes along every axis: TiledDeviceTensor A("A", DEV_ONLY, std::array{1, 0, 2})
    template <typename Sizes>
        requires requires(Sizes const &s) { { *s.begin() } -> std::convertible_to<std::int64_t>; }
    TiledDeviceTensor(std::string name, detail::HostToDeviceMode, Sizes const &sizes) : _name(std::move(name)) {
        for (size_t a = 0; a < R; a++) _sizes[a].assign(sizes.begin(), sizes.end());
        finish();
    }
    template <typename Sizes>
        requires requires(Sizes const &s) { { *s.begin() } -> std::convertible_to<std::int64_t>; }
    TiledDeviceTensor(std::string name, Sizes const &sizes) : TiledDeviceTensor(std::move(name), detail::DEV_ONLY, sizes) {}
    /// one size list per axis
    TiledDeviceTensor(std::string name, detail::HostToDeviceMode, std::array<std::vector<std::int64_t>, R> const &sizes) : _name(std::move(name)), _sizes(sizes) {
        finish();
    }
    TiledDeviceTensor(TiledDeviceTensor const &o) : _name(o._name), _sizes(o._sizes) {
        finish();
        for (auto const &[idx, t] : o._tiles) {
            Tile &mine = tile(idx);
            if (mine.size()) EB200_CHECK(eb200_memcpy_d2d(mine.gpu_data(), t.gpu_data(), size_t(mine.size()) * sizeof(T), gpu::get_stream()));
        }
        gpu::stream_wait();
    }
    TiledDeviceTensor(TiledDeviceTensor &&) noexcept = default;
    TiledDeviceTensor &operator=(TiledDeviceTensor &&) noexcept = default;
    TiledDeviceTensor &operator=(TiledDeviceTensor const &o) {
        TiledDeviceTensor tmp(o);
        *this = std::move(tmp);
        return *this;
    }

    // ---- shape -------------------------------------------------------------------------------------------------
    std::int64_t dim(std::ptrdiff_t a) const { return _dims[size_t(a < 0 ? a + std::ptrdiff_t(R) : a)]; }
    Dim<R>       dims() const { return _dims; }
    std::vector<std::int64_t> const &tile_offset(size_t a = 0) const { return _offsets.at(a); }
    std::vector<std::int64_t> const &tile_size(size_t a = 0) const { return _sizes.at(a); }
    std::int64_t grid_size(size_t a) const { return std::int64_t(_sizes.at(a).size()); }
    std::int64_t grid_size() const {
        std::int64_t g = 1;
        for (size_t a = 0; a < R; a++) g *= grid_size(a);
        return g;
    }
    size_t             num_filled() const { return _tiles.size(); }
    std::string const &name() const { return _name; }
    void               set_name(std::string const &n) { _name = n; }

    // ---- tiles -------------------------------------------------------------------------------------------------
    bool has_tile(TileIndex idx) const { return _tiles.count(normalise(idx)) != 0; }
    bool has_zero_size(TileIndex idx) const {
        idx = normalise(idx);
        for (size_t a = 0; a < R; a++)
            if (_sizes[a][size_t(idx[a])] == 0) return true;
        return false;
    }
    /// the tile at grid position idx, created (zero-filled) when absent
    Tile &tile(TileIndex idx) {
        idx = normalise(idx);
        auto it = _tiles.find(idx);
        if (it != _tiles.end()) return it->second;
        Dim<R>       d;
        Stride<R>    s;
        std::int64_t n = 1;
        for (std::ptrdiff_t a = std::ptrdiff_t(R) - 1; a >= 0; a--) {
            d[size_t(a)] = _sizes[size_t(a)][size_t(idx[size_t(a)])];
            s[size_t(a)] = n;
            n *= d[size_t(a)];
        }
        if (!_arena) _arena = std::make_unique<detail::DeviceArena>(dense_bytes());
        T *p = static_cast<T *>(_arena->take(size_t(std::max<std::int64_t>(n, 1)) * sizeof(T)));
        if (n) EB200_CHECK(eb200_memset_zero(p, size_t(n) * sizeof(T), gpu::get_stream()));
        std::string tname = _name + " - (";
        for (size_t a = 0; a < R; a++) tname += (a ? ", " : "") + std::to_string(idx[a]);
        return _tiles.emplace(idx, Tile(p, d, s, tname + ")")).first->second;
    }
    Tile const &tile(TileIndex idx) const {
        auto it = _tiles.find(normalise(idx));
        if (it == _tiles.end()) throw std::out_of_range("Tile not found!");
        return it->second;
    }
    template <typename... I>
        requires(sizeof...(I) == R && (std::is_integral_v<std::remove_cvref_t<I>> && ...))
    Tile &tile(I... i) { return tile(TileIndex{int(i)...}); }
    std::map<TileIndex, Tile> const &tiles() const { return _tiles; }
    /// the tile if it exists, else nullptr (one lookup; idx must already be in range)
    Tile const *find_tile(TileIndex const &idx) const {
        auto it = _tiles.find(idx);
        return it == _tiles.end() ? nullptr : &it->second;
    }

    /// grid position that holds element index idx, and the index inside that tile (reference: tile_of)
    TileIndex tile_of(std::array<std::int64_t, R> idx, std::array<std::int64_t, R> *inside = nullptr) const {
        TileIndex out{};
        for (size_t a = 0; a < R; a++) {
            std::int64_t v = idx[a] < 0 ? idx[a] + _dims[a] : idx[a];
            if (v < 0 || v >= _dims[a]) throw std::out_of_range("TiledDeviceTensor index out of range");
            size_t t = 0;
            while (t + 1 < _sizes[a].size() && v >= _offsets[a][t + 1]) t++;
            while (_sizes[a][t] == 0 && t + 1 < _sizes[a].size()) t++; // never land on an empty tile
            out[a] = int(t);
            if (inside) (*inside)[a] = v - _offsets[a][t];
        }
        return out;
    }

    // ---- element access ----------------------------------------------------------------------------------------
    template <typename... I>
        requires(sizeof...(I) == R && (std::is_integral_v<std::remove_cvref_t<I>> && ...))
    HostDevReference<T> operator()(I... i) { // creates the tile: the caller may be about to write
        std::array<std::int64_t, R> in{};
        TileIndex                   t = tile_of({static_cast<std::int64_t>(i)...}, &in);
        Tile                       &tl = tile(t);
        std::int64_t                off = 0;
        for (size_t a = 0; a < R; a++) off += in[a] * tl.stride(std::ptrdiff_t(a));
        return HostDevReference<T>(tl.gpu_data() + off);
    }
    template <typename... I>
        requires(sizeof...(I) == R && (std::is_integral_v<std::remove_cvref_t<I>> && ...))
    T operator()(I... i) const {
        std::array<std::int64_t, R> in{};
        TileIndex                   t = tile_of({static_cast<std::int64_t>(i)...}, &in);
        auto                        it = _tiles.find(t);
        if (it == _tiles.end()) return T{0};
        std::int64_t off = 0;
        for (size_t a = 0; a < R; a++) off += in[a] * it->second.stride(std::ptrdiff_t(a));
        return HostDevReference<T>(const_cast<T *>(it->second.gpu_data()) + off).get();
    }

    // ---- whole-tensor operations --------------------------------------------------------------------------------
    /// zero the tiles that exist (reference: zero_no_clear; zero() there also keeps the tiles)
    void zero() {
        for (auto &[idx, t] : _tiles)
            if (t.size()) EB200_CHECK(eb200_memset_zero(t.gpu_data(), size_t(t.size()) * sizeof(T), gpu::get_stream()));
        gpu::op_wait();
    }
    void zero_no_clear() { zero(); }
    /// every tile is created and filled
    void set_all(T v) {
        for_each_grid([&](TileIndex idx) {
            if (has_zero_size(idx)) return;
            Tile &t = tile(idx);
            EB200_CHECK(linear_algebra::detail::xfill(gpu::get_stream(), t.size(), v, t.gpu_data(), 1));
        });
        gpu::op_wait();
    }
    TiledDeviceTensor &operator*=(T v) {
        for (auto &[idx, t] : _tiles)
            if (t.size()) EB200_CHECK(linear_algebra::detail::xscal(gpu::get_stream(), t.size(), v, t.gpu_data(), 1));
        gpu::op_wait();
        return *this;
    }
    TiledDeviceTensor &operator/=(T v) { return *this *= (T{1} / v); }

    /// dense copy (zeros where no tile exists)
    explicit operator DeviceTensor<T, R>() const {
        DeviceTensor<T, R> out(_dims);
        out.set_name(_name);
        out.zero();
        for (auto const &[idx, t] : _tiles) {
            if (!t.size()) continue;
            Offset<R> o;
            for (size_t a = 0; a < R; a++) o[a] = _offsets[a][size_t(idx[a])];
            DeviceTensorView<T, R> dst(out, t.dims(), o);
            EB200_CHECK(eb200_copy_strided(gpu::get_stream(), detail::DTypeOf<T>::value, int(R), t.dims().data(), t.gpu_data(), t.strides().data(),
                                           dst.gpu_data(), dst.strides().data()));
        }
        gpu::stream_wait();
        return out;
    }

    template <typename F> void for_each_grid(F &&fn) const {
        std::int64_t g = grid_size();
        for (std::int64_t s = 0; s < g; s++) {
            TileIndex    idx{};
            std::int64_t x = s;
            for (std::ptrdiff_t a = std::ptrdiff_t(R) - 1; a >= 0; a--) {
                idx[size_t(a)] = int(x % grid_size(size_t(a)));
                x /= grid_size(size_t(a));
            }
            fn(idx);
        }
    }

  private:
    size_t dense_bytes() const {
        double b = double(sizeof(T));
        for (size_t a = 0; a < R; a++) b *= double(_dims[a]);
        b += 256.0 * double(grid_size());
        return b > 1e18 ? size_t(1e18) : size_t(b);
    }
    void finish() {
        for (size_t a = 0; a < R; a++) {
            _offsets[a].clear();
            std::int64_t sum = 0;
            for (auto s : _sizes[a]) {
                if (s < 0) throw dimension_error("Tile sizes must not be negative.");
                _offsets[a].push_back(sum);
                sum += s;
            }
            _dims[a] = sum;
        }
    }
    TileIndex normalise(TileIndex idx) const {
        for (size_t a = 0; a < R; a++) {
            if (idx[a] < 0) idx[a] += int(_sizes[a].size());
            if (idx[a] < 0 || size_t(idx[a]) >= _sizes[a].size()) throw std::out_of_range("Tile index out of range");
        }
        return idx;
    }

    std::string                                  _name{"(unnamed)"};
    std::array<std::vector<std::int64_t>, R>     _sizes{}, _offsets{};
    Dim<R>                                       _dims{};
    std::map<TileIndex, Tile>                    _tiles;
    std::unique_ptr<detail::DeviceArena>         _arena;
};

template <typename T, size_t R> struct IsTiledDeviceTensor<TiledDeviceTensor<T, R>> : std::true_type {};
template <typename X>
concept TiledDeviceTensorLike = IsTiledDeviceTensor<std::remove_cvref_t<X>>::value;

// =================================================================================================================
namespace tensor_algebra {
namespace detail {

/// Run-time form of a contraction over tiles: position of every axis of C/A/B in a list of n_unique indices.
template <size_t CR, size_t AR, size_t BR> struct TileTables {
    int                 n_unique = 0;
    std::array<int, CR> c{};
    std::array<int, AR> a{};
    std::array<int, BR> b{};
};

template <typename T, typename CT, typename AT, typename BT, size_t CR, size_t AR, size_t BR>
void tile_desc_into(eb200_contract_desc &d, TileTables<CR, AR, BR> const &tab, T beta, CT *c_data, std::array<std::int64_t, CR> const &cdim,
                    std::array<std::int64_t, CR> const &cstr, T alpha, AT const &A, BT const &B, bool conj_a, bool conj_b) {
    d.conj_a = d.conj_b = d.force_kernel = d.split_k = d.kernel_used = d.launches = 0;
    d.n_unique = tab.n_unique;
    for (int u = 0; u < tab.n_unique; u++) d.extent[u] = -1;
    auto put = [&](int u, std::int64_t e) {
        if (d.extent[u] < 0) d.extent[u] = e;
        else if (d.extent[u] != e) throw dimension_error("Inconsistent dimensions found!");
    };
    d.C.rank = int(CR); d.A.rank = int(AR); d.B.rank = int(BR);
    for (size_t p = 0; p < CR; p++) { d.C.index[p] = tab.c[p]; d.C.stride[p] = cstr[p]; put(tab.c[p], cdim[p]); }
    for (size_t p = 0; p < AR; p++) { d.A.index[p] = tab.a[p]; d.A.stride[p] = A.stride(std::ptrdiff_t(p)); put(tab.a[p], A.dim(std::ptrdiff_t(p))); }
    for (size_t p = 0; p < BR; p++) { d.B.index[p] = tab.b[p]; d.B.stride[p] = B.stride(std::ptrdiff_t(p)); put(tab.b[p], B.dim(std::ptrdiff_t(p))); }
    d.C.data = c_data;
    d.A.data = const_cast<T *>(A.gpu_data());
    d.B.data = const_cast<T *>(B.gpu_data());
    d.C.dtype = d.A.dtype = d.B.dtype = einsums::detail::DTypeOf<T>::value;
    std::complex<double> a(alpha), b(beta);
    d.alpha[0] = a.real(); d.alpha[1] = a.imag(); d.beta[0] = b.real(); d.beta[1] = b.imag();
    d.conj_a = conj_a && IsComplexV<T>;
    d.conj_b = conj_b && IsComplexV<T>;
}
template <typename T, typename CT, typename AT, typename BT, size_t CR, size_t AR, size_t BR>
eb200_contract_desc tile_desc(TileTables<CR, AR, BR> const &tab, T beta, CT *c_data, std::array<std::int64_t, CR> const &cdim,
                              std::array<std::int64_t, CR> const &cstr, T alpha, AT const &A, BT const &B, bool conj_a, bool conj_b) {
    eb200_contract_desc d{};
    tile_desc_into<T>(d, tab, beta, c_data, cdim, cstr, alpha, A, B, conj_a, conj_b);
    return d;
}

/// grid extent of every unique index, resolved C -> A -> B (reference: get_grid_ranges_for_many)
template <typename CType, typename AType, typename BType, size_t CR, size_t AR, size_t BR>
std::vector<std::int64_t> grid_ranges(TileTables<CR, AR, BR> const &tab, CType const *C, AType const &A, BType const &B) {
    std::vector<std::int64_t> g(size_t(tab.n_unique), -1);
    auto put = [&](int u, std::int64_t e) {
        if (g[size_t(u)] < 0) g[size_t(u)] = e;
        else if (g[size_t(u)] != e) throw tensor_compat_error("Tiled tensors do not have compatible tile grids.");
    };
    if constexpr (CR > 0)
        for (size_t p = 0; p < CR; p++) put(tab.c[p], C->grid_size(p));
    for (size_t p = 0; p < AR; p++) put(tab.a[p], A.grid_size(p));
    for (size_t p = 0; p < BR; p++) put(tab.b[p], B.grid_size(p));
    return g;
}

/// C = beta*C + alpha * A*B over tiles (reference: TileAlgebra.hpp:17-91): one grouped call, see the file header.
template <typename T, size_t CR, size_t AR, size_t BR>
void tiled_contract(TileTables<CR, AR, BR> const &tab, T beta, TiledDeviceTensor<T, CR> *C, T alpha, TiledDeviceTensor<T, AR> const &A,
                    TiledDeviceTensor<T, BR> const &B, bool conj_a, bool conj_b) {
    auto grid = grid_ranges(tab, C, A, B);
    if (beta == T(0)) C->zero_no_clear();
    else if (beta != T(1)) *C *= beta;

    std::vector<eb200_contract_desc> descs;
    std::int64_t                     total = 1;
    for (auto g : grid) total *= g;
    std::vector<int> u(size_t(tab.n_unique), 0);
    for (std::int64_t s = 0; s < total; s++) {
        std::int64_t x = s;
        for (int k = tab.n_unique - 1; k >= 0; k--) { // last unique index fastest, as sentinel_to_indices
            u[size_t(k)] = int(x % grid[size_t(k)]);
            x /= grid[size_t(k)];
        }
        std::array<int, AR> ai; std::array<int, BR> bi; std::array<int, CR> ci;
        for (size_t p = 0; p < AR; p++) ai[p] = u[size_t(tab.a[p])];
        for (size_t p = 0; p < BR; p++) bi[p] = u[size_t(tab.b[p])];
        for (size_t p = 0; p < CR; p++) ci[p] = u[size_t(tab.c[p])];
        auto const *at = A.find_tile(ai);
        if (!at || !at->size()) continue;
        auto const *bt = B.find_tile(bi);
        if (!bt || !bt->size() || C->has_zero_size(ci)) continue;
        auto &ct = C->tile(ci);
        if (descs.empty()) descs.reserve(size_t(std::min<std::int64_t>(total - s, std::int64_t(1) << 20)));
        descs.emplace_back(); // filled in place: only the fields the library reads are written
        tile_desc_into<T>(descs.back(), tab, T(1), ct.gpu_data(), ct.dims(), ct.strides(), alpha, *at, *bt, conj_a, conj_b);
    }
    if (!descs.empty()) EB200_CHECK(eb200_contract_grouped(gpu::get_stream(), std::int64_t(descs.size()), descs.data(), nullptr));
    gpu::op_wait();
}

/// scalar result (reference: TileAlgebra.hpp:93-158): every tile pair reduces into its own device slot, summed on the host
template <typename T, size_t AR, size_t BR>
T tiled_contract_scalar(TileTables<0, AR, BR> const &tab, T alpha, TiledDeviceTensor<T, AR> const &A, TiledDeviceTensor<T, BR> const &B, bool conj_a,
                        bool conj_b) {
    auto grid = grid_ranges<TiledDeviceTensor<T, 1>>(tab, nullptr, A, B);
    std::vector<eb200_contract_desc> descs;
    std::int64_t                     total = 1;
    for (auto g : grid) total *= g;
    std::vector<int> u(size_t(tab.n_unique), 0);
    std::vector<std::pair<std::array<int, AR>, std::array<int, BR>>> pairs;
    for (std::int64_t s = 0; s < total; s++) {
        std::int64_t x = s;
        for (int k = tab.n_unique - 1; k >= 0; k--) {
            u[size_t(k)] = int(x % grid[size_t(k)]);
            x /= grid[size_t(k)];
        }
        std::array<int, AR> ai; std::array<int, BR> bi;
        for (size_t p = 0; p < AR; p++) ai[p] = u[size_t(tab.a[p])];
        for (size_t p = 0; p < BR; p++) bi[p] = u[size_t(tab.b[p])];
        if (!A.has_tile(ai) || !B.has_tile(bi) || A.has_zero_size(ai) || B.has_zero_size(bi)) continue;
        pairs.emplace_back(ai, bi);
    }
    if (pairs.empty()) return T{0};
    void *p = nullptr;
    EB200_CHECK(eb200_malloc(&p, pairs.size() * sizeof(T)));
    T *slots = static_cast<T *>(p);
    for (size_t k = 0; k < pairs.size(); k++)
        descs.push_back(tile_desc<T>(tab, T(0), slots + k, {}, {}, alpha, A.tile(pairs[k].first), B.tile(pairs[k].second), conj_a, conj_b));
    std::vector<T> host(pairs.size());
    try {
        EB200_CHECK(eb200_contract_grouped(gpu::get_stream(), std::int64_t(descs.size()), descs.data(), nullptr));
        EB200_CHECK(eb200_memcpy_d2h(host.data(), slots, host.size() * sizeof(T), gpu::get_stream()));
        gpu::stream_wait();
    } catch (...) {
        eb200_free(slots);
        throw;
    }
    eb200_free(slots);
    return std::accumulate(host.begin(), host.end(), T{0});
}

template <typename Tuple, size_t N> void fill_table(IdPack const &all, IdPack const &mine, std::array<int, N> &out) {
    for (size_t p = 0; p < N; p++) {
        out[p] = -1;
        for (int q = 0; q < all.n; q++)
            if (all.v[q] == mine.v[p]) out[p] = q;
    }
}

} // namespace detail

/// einsum over tiled tensors; C is a tiled tensor or a scalar (host value / rank-0 DeviceTensor)
template <bool ConjA = false, bool ConjB = false, TiledDeviceTensorLike AType, TiledDeviceTensorLike BType, typename CType, typename U,
          typename... CIndices, typename... AIndices, typename... BIndices>
    requires((std::is_arithmetic_v<U> || IsComplexV<U>) && !IsBlockDeviceTensor<std::remove_cvref_t<CType>>::value)
void einsum(U const UC_prefactor, std::tuple<CIndices...> const &, CType *C, U const UAB_prefactor, std::tuple<AIndices...> const &,
            AType const &A, std::tuple<BIndices...> const &, BType const &B, detail::AlgorithmChoice *algorithm_choice = nullptr) {
    using namespace detail;
    using T = typename AType::ValueType;
    static_assert(std::is_same_v<T, typename BType::ValueType>, "tiled einsum needs one value type");
    constexpr size_t   CR = sizeof...(CIndices);
    constexpr IdPack   pc = make_pack(ids_of<std::tuple<CIndices...>>()), pa = make_pack(ids_of<std::tuple<AIndices...>>()),
                     pb = make_pack(ids_of<std::tuple<BIndices...>>());
    constexpr Analysis z = analyse(pc, pa, pb);
    TileTables<CR, AType::Rank, BType::Rank> tab;
    tab.n_unique = z.all_unique.n;
    fill_table<void>(z.all_unique, z.C, tab.c);
    fill_table<void>(z.all_unique, z.A, tab.a);
    fill_table<void>(z.all_unique, z.B, tab.b);
    if constexpr (CR == 0) {
        using CV = std::conditional_t<DeviceTensorLike<CType>, T, CType>;
        T  part = tiled_contract_scalar<T>(tab, T(UAB_prefactor), A, B, ConjA, ConjB);
        CV c0 = (CV(UC_prefactor) == CV(0)) ? CV(0) : CV(*C) * CV(UC_prefactor);
        *C = c0 + CV(part);
    } else {
        static_assert(TiledDeviceTensorLike<CType>, "C must be a tiled tensor or a scalar");
        tiled_contract<T>(tab, T(UC_prefactor), C, T(UAB_prefactor), A, B, ConjA, ConjB);
    }
    if (algorithm_choice) *algorithm_choice = detail::GENERIC;
}
template <bool ConjA = false, bool ConjB = false, TiledDeviceTensorLike AType, TiledDeviceTensorLike BType, typename CType, typename... CIndices,
          typename... AIndices, typename... BIndices>
    requires(!IsBlockDeviceTensor<std::remove_cvref_t<CType>>::value)
void einsum(std::tuple<CIndices...> const &C_indices, CType *C, std::tuple<AIndices...> const &A_indices, AType const &A,
            std::tuple<BIndices...> const &B_indices, BType const &B, detail::AlgorithmChoice *algorithm_choice = nullptr) {
    using V = typename AType::ValueType;
    using S = std::conditional_t<IsComplexV<V>, V, double>;
    einsum<ConjA, ConjB>(S(0), C_indices, C, S(1), A_indices, A, B_indices, B, algorithm_choice);
}

} // namespace tensor_algebra

// =================================================================================================================
namespace linear_algebra {

/// C = beta*C + alpha*op(A)*op(B) on tiled matrices (reference: LinearAlgebra/TiledTensor.hpp:115-228)
template <typename T, typename U>
void gemm(char transA, char transB, U alpha, TiledDeviceTensor<T, 2> const &A, TiledDeviceTensor<T, 2> const &B, U beta, TiledDeviceTensor<T, 2> *C) {
    if (!std::strchr("CNTcnt", transA) || !std::strchr("CNTcnt", transB))
        throw std::invalid_argument("The transposition parameters were invalid! They must be either n, c, or t, case insensitive.");
    const bool tA = std::tolower(transA) != 'n', tB = std::tolower(transB) != 'n';
    tensor_algebra::detail::TileTables<2, 2, 2> tab; // unique = (i, j, k)
    tab.n_unique = 3;
    tab.c = {0, 1};
    tab.a = tA ? std::array<int, 2>{2, 0} : std::array<int, 2>{0, 2};
    tab.b = tB ? std::array<int, 2>{1, 2} : std::array<int, 2>{2, 1};
    tensor_algebra::detail::tiled_contract<T>(tab, T(beta), C, T(alpha), A, B, std::tolower(transA) == 'c', std::tolower(transB) == 'c');
}
template <bool TransA, bool TransB, typename T, typename U>
void gemm(U alpha, TiledDeviceTensor<T, 2> const &A, TiledDeviceTensor<T, 2> const &B, U beta, TiledDeviceTensor<T, 2> *C) {
    gemm(TransA ? 't' : 'n', TransB ? 't' : 'n', alpha, A, B, beta, C);
}

/// y = beta*y + alpha*op(A)*x with a tiled A and ordinary vectors (reference :231-355)
template <typename T, DeviceTensorLike XType, DeviceTensorLike YType, typename U>
    requires(XType::Rank == 1 && YType::Rank == 1)
void gemv(char transA, U alpha, TiledDeviceTensor<T, 2> const &A, XType const &x, U beta, YType *y) {
    const bool tA = std::tolower(transA) != 'n', cA = std::tolower(transA) == 'c';
    const size_t out_axis = tA ? 1 : 0, in_axis = tA ? 0 : 1;
    if (x.dim(0) != A.dim(std::ptrdiff_t(in_axis)) || y->dim(0) != A.dim(std::ptrdiff_t(out_axis))) throw dimension_error("gemv: vectors do not match the tiled matrix");
    if (T(beta) == T(0)) y->zero();
    else if (T(beta) != T(1)) scale(T(beta), y);
    tensor_algebra::detail::TileTables<1, 2, 1> tab; // unique = (out, in)
    tab.n_unique = 2;
    tab.c = {0};
    tab.a = tA ? std::array<int, 2>{1, 0} : std::array<int, 2>{0, 1};
    tab.b = {1};
    std::vector<eb200_contract_desc> descs;
    for (auto const &[idx, t] : A.tiles()) {
        if (!t.size()) continue;
        int          to = idx[out_axis], ti = idx[in_axis];
        std::int64_t o0 = A.tile_offset(out_axis)[size_t(to)], i0 = A.tile_offset(in_axis)[size_t(ti)];
        DeviceTensorView<T, 1> xs(const_cast<T *>(x.gpu_data()) + i0 * x.stride(0), Dim<1>{A.tile_size(in_axis)[size_t(ti)]}, Stride<1>{x.stride(0)});
        std::array<std::int64_t, 1> cd{A.tile_size(out_axis)[size_t(to)]}, cs{y->stride(0)};
        descs.push_back(tensor_algebra::detail::tile_desc<T>(tab, T(1), y->gpu_data() + o0 * y->stride(0), cd, cs, T(alpha), t, xs, cA, false));
    }
    if (!descs.empty()) EB200_CHECK(eb200_contract_grouped(gpu::get_stream(), std::int64_t(descs.size()), descs.data(), nullptr));
    gpu::op_wait();
}
template <bool TransA, typename T, DeviceTensorLike XType, DeviceTensorLike YType, typename U>
    requires(XType::Rank == 1 && YType::Rank == 1)
void gemv(U alpha, TiledDeviceTensor<T, 2> const &A, XType const &x, U beta, YType *y) {
    gemv(TransA ? 't' : 'n', alpha, A, x, beta, y);
}

template <typename T, size_t R> void scale(T alpha, TiledDeviceTensor<T, R> *A) { *A *= alpha; }

/// sum over the tiles present in both operands (reference :35-110)
template <typename T, size_t R> T dot(TiledDeviceTensor<T, R> const &A, TiledDeviceTensor<T, R> const &B) {
    tensor_algebra::detail::TileTables<0, R, R> tab;
    tab.n_unique = int(R);
    for (size_t p = 0; p < R; p++) tab.a[p] = tab.b[p] = int(p);
    return tensor_algebra::detail::tiled_contract_scalar<T>(tab, T(1), A, B, false, false);
}
template <typename T, size_t R> T true_dot(TiledDeviceTensor<T, R> const &A, TiledDeviceTensor<T, R> const &B) {
    tensor_algebra::detail::TileTables<0, R, R> tab;
    tab.n_unique = int(R);
    for (size_t p = 0; p < R; p++) tab.a[p] = tab.b[p] = int(p);
    return tensor_algebra::detail::tiled_contract_scalar<T>(tab, T(1), A, B, true, false);
}

} // namespace linear_algebra
} // namespace einsums

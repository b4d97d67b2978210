// BlockTensor.hpp — BlockDeviceTensor<T, Rank>: a block-diagonal tensor on the device, plus the einsum / linear-algebra
// overloads that take it. Same names and call syntax as the reference
// (libs/Einsums/Tensor/include/Einsums/Tensor/BlockTensor.hpp:1355-1700, tensor_base::BlockTensor :60-1215;
//  algebra: TensorAlgebra/Backends/BlockAlgebra.hpp:23-429, LinearAlgebra/BlockTensor.hpp:43-460).
//
// B200-first layout: all blocks live back to back in ONE device allocation (the reference keeps one DeviceTensor per
// block). Whole-tensor operations (zero, scale, axpy, dot, direct_product) are therefore one streaming launch over the
// store, and a per-block contraction over equal-size blocks is one batched launch: eb200_contract_grouped folds the
// constant pointer step between blocks into an extra batch index. Blocks of different sizes are enqueued back to back on
// the calling thread's stream with a single wait at the end (the reference waits after every block).
#pragma once
#include <Einsums/B200/LinearAlgebra.hpp>
#include <Einsums/B200/TensorAlgebra.hpp>

#include <numeric>

namespace einsums {

template <typename T, size_t R> class BlockDeviceTensor {
  public:
    using ValueType = T;
    using Block     = DeviceTensorView<T, R>;
    static constexpr size_t Rank = R;

    BlockDeviceTensor() = default;
    template <std::convertible_to<std::int64_t>... Dims>
    explicit BlockDeviceTensor(std::string name, detail::HostToDeviceMode, Dims... block_dims) : _name(std::move(name)) {
        build({static_cast<std::int64_t>(block_dims)...});
    }
    template <std::convertible_to<std::int64_t>... Dims>
    explicit BlockDeviceTensor(std::string name, Dims... block_dims) : BlockDeviceTensor(std::move(name), detail::DEV_ONLY, block_dims...) {}
    template <typename ArrayArg>
        requires requires(ArrayArg const &a) { a.begin(); a.end(); }
    explicit BlockDeviceTensor(std::string name, detail::HostToDeviceMode, ArrayArg const &block_dims) : _name(std::move(name)) {
        build(std::vector<std::int64_t>(block_dims.begin(), block_dims.end()));
    }
    template <typename ArrayArg>
        requires requires(ArrayArg const &a) { a.begin(); a.end(); } && (!std::is_convertible_v<ArrayArg, std::int64_t>)
    explicit BlockDeviceTensor(std::string name, ArrayArg const &block_dims) : BlockDeviceTensor(std::move(name), detail::DEV_ONLY, block_dims) {}

    BlockDeviceTensor(BlockDeviceTensor const &o) : _name(o._name) {
        build(o._bdims);
        if (_total) EB200_CHECK(eb200_memcpy_d2d(_store, o._store, size_t(_total) * sizeof(T), gpu::get_stream()));
        gpu::stream_wait();
    }
    BlockDeviceTensor(BlockDeviceTensor &&o) noexcept { swap(o); }
    BlockDeviceTensor &operator=(BlockDeviceTensor o) noexcept {
        swap(o);
        return *this;
    }
    ~BlockDeviceTensor() {
        if (_store) eb200_free(_store);
    }

    // ---- shape -------------------------------------------------------------------------------------------------
    size_t       num_blocks() const { return _bdims.size(); }
    std::int64_t block_dim(size_t i) const { return _bdims.at(i); }
    Range        block_range(size_t i) const { return Range{_begin.at(i), _begin.at(i) + _bdims.at(i)}; }
    std::vector<Range> ranges() const {
        std::vector<Range> r;
        for (size_t i = 0; i < num_blocks(); i++) r.push_back(block_range(i));
        return r;
    }
    std::vector<std::int64_t> const &vector_dims() const { return _bdims; }
    std::int64_t dim(std::ptrdiff_t = 0) const { return _dim; }
    Dim<R>       dims() const {
        Dim<R> d;
        d.fill(_dim);
        return d;
    }
    /// number of stored elements (all blocks); the dense tensor would hold dim()^Rank
    std::int64_t size() const { return _total; }
    /// block that holds position `index` along any axis (reference: BlockTensor.hpp block_of; throws std::out_of_range)
    int block_of(std::int64_t index) const {
        if (index >= 0)
            for (size_t i = 0; i < num_blocks(); i++)
                if (index < _begin[i] + _bdims[i]) return int(i);
        throw std::out_of_range("Index out of range!");
    }
    std::string const &name() const { return _name; }
    void               set_name(std::string const &n) { _name = n; }
    void               set_name(int i, std::string const &n) { _blocks.at(size_t(i)).set_name(n); }

    // ---- blocks ------------------------------------------------------------------------------------------------
    Block       &block(size_t i) { return _blocks.at(i); }
    Block const &block(size_t i) const { return _blocks.at(i); }
    template <std::integral I> Block       &operator[](I i) { return _blocks.at(size_t(i)); }
    template <std::integral I> Block const &operator[](I i) const { return _blocks.at(size_t(i)); }
    Block       &block(std::string const &name) { return _blocks.at(find(name)); }
    Block const &block(std::string const &name) const { return _blocks.at(find(name)); }
    Block       &operator[](std::string const &name) { return block(name); }
    Block const &operator[](std::string const &name) const { return block(name); }
    Block       &operator[](char const *name) { return block(std::string(name)); }
    Block const &operator[](char const *name) const { return block(std::string(name)); }

    /// append / insert a block (its values are copied into the re-packed store)
    template <typename TensorT> void push_block(TensorT const &value) { insert_block(int(num_blocks()), value); }
    template <typename TensorT> void insert_block(int pos, TensorT const &value) {
        for (size_t a = 1; a < R; a++)
            if (value.dim(a) != value.dim(0)) throw dimension_error("Can only push square/hypersquare tensors to a block tensor.");
        if (pos < 0 || size_t(pos) > num_blocks()) throw std::out_of_range("insert_block: position out of range");
        std::vector<std::int64_t> nd = _bdims;
        std::vector<std::string>  names;
        for (auto &b : _blocks) names.push_back(b.name());
        nd.insert(nd.begin() + pos, value.dim(0));
        names.insert(names.begin() + pos, value.name());
        BlockDeviceTensor fresh;
        fresh._name = _name;
        fresh.build(nd);
        for (size_t i = 0, old = 0; i < nd.size(); i++) {
            fresh._blocks[i].set_name(names[i]);
            if (int(i) == pos) copy_into(fresh._blocks[i], value);
            else copy_into(fresh._blocks[i], _blocks[old++]);
        }
        gpu::stream_wait();
        swap(fresh);
    }

    // ---- element access (host <-> device one element at a time, like the reference) -------------------------
    template <typename... Idx>
        requires(sizeof...(Idx) == R && (std::is_integral_v<std::remove_cvref_t<Idx>> && ...))
    HostDevReference<T> operator()(Idx... idx) {
        std::int64_t off = element_offset({static_cast<std::int64_t>(idx)...});
        if (off < 0) { // outside every block: a zero that swallows writes (reference returns a reference to a dummy zero)
            HostDevReference<T> dummy{zero_slot()};
            dummy = T{0};
            return dummy;
        }
        return HostDevReference<T>(_store + off);
    }
    template <typename... Idx>
        requires(sizeof...(Idx) == R && (std::is_integral_v<std::remove_cvref_t<Idx>> && ...))
    T operator()(Idx... idx) const {
        std::int64_t off = element_offset({static_cast<std::int64_t>(idx)...});
        return off < 0 ? T{0} : HostDevReference<T>(_store + off).get();
    }

    // ---- whole-tensor operations: one launch over the packed store -------------------------------------------
    void zero() {
        if (_total) EB200_CHECK(eb200_memset_zero(_store, size_t(_total) * sizeof(T), gpu::get_stream()));
        gpu::op_wait();
    }
    void set_all(T v) {
        if (_total) EB200_CHECK(linear_algebra::detail::xfill(gpu::get_stream(), _total, v, _store, 1));
        gpu::op_wait();
    }
    BlockDeviceTensor &operator*=(T v) {
        if (_total) EB200_CHECK(linear_algebra::detail::xscal(gpu::get_stream(), _total, v, _store, 1));
        gpu::op_wait();
        return *this;
    }
    BlockDeviceTensor &operator/=(T v) { return *this *= (T{1} / v); }
    BlockDeviceTensor &operator+=(T v) {
        if (_total) {
            HostDevReference<T> slot{zero_slot()}; // x = the scalar, incx = 0
            slot = v;
            EB200_CHECK(linear_algebra::detail::xaxpy(gpu::get_stream(), _total, T{1}, zero_slot(), 0, _store, 1));
        }
        gpu::op_wait();
        return *this;
    }
    BlockDeviceTensor &operator-=(T v) { return *this += (-v); }
    BlockDeviceTensor &operator+=(BlockDeviceTensor const &o) {
        same_blocks(o);
        if (_total) EB200_CHECK(linear_algebra::detail::xaxpy(gpu::get_stream(), _total, T{1}, o._store, 1, _store, 1));
        gpu::op_wait();
        return *this;
    }
    BlockDeviceTensor &operator-=(BlockDeviceTensor const &o) {
        same_blocks(o);
        if (_total) EB200_CHECK(linear_algebra::detail::xaxpy(gpu::get_stream(), _total, T{-1}, o._store, 1, _store, 1));
        gpu::op_wait();
        return *this;
    }
    /// elementwise product / quotient inside the blocks
    BlockDeviceTensor &operator*=(BlockDeviceTensor const &o) { return elementwise(o, false); }
    BlockDeviceTensor &operator/=(BlockDeviceTensor const &o) { return elementwise(o, true); }

    /// dense copy (zeros outside the blocks); reference: explicit operator TensorType()
    explicit operator DeviceTensor<T, R>() const {
        DeviceTensor<T, R> out(dims());
        out.set_name(_name);
        out.zero();
        for (size_t i = 0; i < num_blocks(); i++) {
            if (_bdims[i] == 0) continue;
            Dim<R>    d;
            Offset<R> o;
            d.fill(_bdims[i]);
            o.fill(_begin[i]);
            DeviceTensorView<T, R> dst(out, d, o);
            EB200_CHECK(eb200_copy_strided(gpu::get_stream(), detail::DTypeOf<T>::value, int(R), d.data(), _blocks[i].gpu_data(),
                                           _blocks[i].strides().data(), dst.gpu_data(), dst.strides().data()));
        }
        gpu::stream_wait();
        return out;
    }

    /// the packed store (all blocks, block-major) and the element offset of block i inside it
    T           *store_data() { return _store; }
    T const     *store_data() const { return _store; }
    std::int64_t block_offset(size_t i) const { return _offset.at(i); }

    void same_blocks(BlockDeviceTensor const &o) const {
        if (num_blocks() != o.num_blocks()) throw tensor_compat_error("Block tensors need to have the same number of blocks.");
        for (size_t i = 0; i < num_blocks(); i++)
            if (_bdims[i] != o._bdims[i]) throw dimension_error("Block tensors need to have the same block sizes.");
    }

  private:
    void build(std::vector<std::int64_t> bd) {
        _bdims = std::move(bd);
        _begin.assign(_bdims.size(), 0);
        _offset.assign(_bdims.size(), 0);
        _dim = 0;
        _total = 0;
        for (size_t i = 0; i < _bdims.size(); i++) {
            if (_bdims[i] < 0) throw dimension_error("Block dimensions must not be negative.");
            _begin[i] = _dim;
            _offset[i] = _total;
            std::int64_t n = 1;
            for (size_t a = 0; a < R; a++) n *= _bdims[i];
            _dim += _bdims[i];
            _total += n;
        }
        (void)gpu::get_stream();
        void *p = nullptr;
        EB200_CHECK(eb200_malloc(&p, size_t(_total + 1) * sizeof(T))); // + one slot: the "zero outside the blocks" element
        _store = static_cast<T *>(p);
        _blocks.clear();
        for (size_t i = 0; i < _bdims.size(); i++) {
            Dim<R>    d;
            Stride<R> s;
            d.fill(_bdims[i]);
            std::int64_t acc = 1;
            for (std::ptrdiff_t a = std::ptrdiff_t(R) - 1; a >= 0; a--) {
                s[size_t(a)] = acc;
                acc *= _bdims[i];
            }
            _blocks.emplace_back(_store + _offset[i], d, s, _name + " - " + std::to_string(i));
        }
    }
    void swap(BlockDeviceTensor &o) noexcept {
        std::swap(_name, o._name); std::swap(_bdims, o._bdims); std::swap(_begin, o._begin); std::swap(_offset, o._offset);
        std::swap(_blocks, o._blocks); std::swap(_dim, o._dim); std::swap(_total, o._total); std::swap(_store, o._store);
    }
    size_t find(std::string const &name) const {
        for (size_t i = 0; i < _blocks.size(); i++)
            if (_blocks[i].name() == name) return i;
        throw std::out_of_range("Could not find block with the name " + name);
    }
    T *zero_slot() const { return _store + _total; }
    std::int64_t element_offset(std::array<std::int64_t, R> idx) const {
        for (auto &v : idx) {
            if (v < 0) v += _dim;
            if (v < 0 || v >= _dim) throw std::out_of_range("BlockDeviceTensor index out of range");
        }
        int b = block_of(idx[0]);
        for (size_t a = 1; a < R; a++)
            if (block_of(idx[a]) != b) return -1;
        std::int64_t off = _offset[size_t(b)];
        for (size_t a = 0; a < R; a++) off += (idx[a] - _begin[size_t(b)]) * _blocks[size_t(b)].stride(std::ptrdiff_t(a));
        return off;
    }
    template <typename TensorT> static void copy_into(Block &dst, TensorT const &src) {
        if (dst.size() == 0) return;
        EB200_CHECK(eb200_copy_strided(gpu::get_stream(), detail::DTypeOf<T>::value, int(R), dst.dims().data(), src.gpu_data(),
                                       src.strides().data(), dst.gpu_data(), dst.strides().data()));
    }
    BlockDeviceTensor &elementwise(BlockDeviceTensor const &o, bool divide);

    std::string               _name{"(unnamed)"};
    std::vector<std::int64_t> _bdims, _begin, _offset;
    std::vector<Block>        _blocks;
    std::int64_t              _dim = 0, _total = 0;
    T                        *_store = nullptr;
};

template <typename T, size_t R> struct IsBlockDeviceTensor<BlockDeviceTensor<T, R>> : std::true_type {};
template <typename X>
concept BlockDeviceTensorLike = IsBlockDeviceTensor<std::remove_cvref_t<X>>::value;

// =================================================================================================================
// einsum on block tensors (reference: TensorAlgebra/Backends/BlockAlgebra.hpp)
// =================================================================================================================
namespace tensor_algebra {
namespace detail {

template <typename X> struct BlockValueOf { using type = X; }; // host scalar
template <typename X>
    requires(BlockDeviceTensorLike<X> || DeviceTensorLike<X>)
struct BlockValueOf<X> { using type = typename std::remove_cvref_t<X>::ValueType; };

/// operand of block i: the block itself, or — for an ordinary tensor — its sub-view over block_range(i) on every axis
template <typename X> auto block_operand(X &x, size_t i, Range r) {
    if constexpr (BlockDeviceTensorLike<X>) return x.block(i);
    else {
        constexpr size_t XR = std::remove_cvref_t<X>::Rank;
        using XV = typename std::remove_cvref_t<X>::ValueType;
        Dim<XR>      d;
        Stride<XR>   s;
        std::int64_t off = 0;
        for (size_t a = 0; a < XR; a++) {
            if (r[1] > x.dim(std::ptrdiff_t(a))) throw std::out_of_range("block range exceeds the tensor");
            d[a] = r[1] - r[0];
            s[a] = x.stride(std::ptrdiff_t(a));
            off += r[0] * s[a];
        }
        return DeviceTensorView<XV, XR>(const_cast<XV *>(x.gpu_data()) + off, d, s);
    }
}

template <typename X, typename Y> void check_same_blocks(X const &x, Y const &y) {
    if constexpr (BlockDeviceTensorLike<X> && BlockDeviceTensorLike<Y>) {
        if (x.num_blocks() != y.num_blocks()) throw tensor_compat_error("Block tensors need to have the same number of blocks.");
        for (size_t i = 0; i < x.num_blocks(); i++)
            if (x.block_dim(i) != y.block_dim(i)) throw dimension_error("Block tensors need to have the same block sizes.");
    }
}
template <typename A, typename B> auto const &first_block_tensor(A const &a, B const &b) {
    if constexpr (BlockDeviceTensorLike<A>) return a;
    else return b;
}

} // namespace detail

/// C = beta*C + alpha * sum_links A*B block by block. A and/or B are block tensors (the other may be a rank-1 DeviceTensor);
/// C is a block tensor, an ordinary tensor of rank >= 1 (the diagonal sub-views are updated) or a scalar (rank-0
/// DeviceTensor or host value: the per-block results are summed). One grouped launch sequence, one wait.
template <bool ConjA = false, bool ConjB = false, typename AType, typename BType, typename CType, typename U, typename... CIndices,
          typename... AIndices, typename... BIndices>
    requires((BlockDeviceTensorLike<AType> || BlockDeviceTensorLike<BType>) && (std::is_arithmetic_v<U> || IsComplexV<U>) &&
             !IsTiledDeviceTensor<std::remove_cvref_t<AType>>::value && !IsTiledDeviceTensor<std::remove_cvref_t<BType>>::value &&
             !IsTiledDeviceTensor<std::remove_cvref_t<CType>>::value)
void einsum(U const UC_prefactor, std::tuple<CIndices...> const &, CType *C, U const UAB_prefactor, std::tuple<AIndices...> const &,
            AType const &A, std::tuple<BIndices...> const &, BType const &B, detail::AlgorithmChoice *algorithm_choice = nullptr) {
    using namespace detail;
    using CV = typename BlockValueOf<CType>::type;
    constexpr bool c_block = BlockDeviceTensorLike<CType>;
    constexpr bool c_scalar = sizeof...(CIndices) == 0;
    static_assert(BlockDeviceTensorLike<AType> || AType::Rank == 1, "the non-block operand must be a vector");
    static_assert(BlockDeviceTensorLike<BType> || BType::Rank == 1, "the non-block operand must be a vector");

    constexpr IdPack   pc = make_pack(ids_of<std::tuple<CIndices...>>()), pa = make_pack(ids_of<std::tuple<AIndices...>>()),
                     pb = make_pack(ids_of<std::tuple<BIndices...>>());
    constexpr Analysis z = analyse(pc, pa, pb);

    check_same_blocks(A, B);
    if constexpr (c_block) {
        check_same_blocks(A, *C);
        check_same_blocks(B, *C);
    }
    auto const  &shape = first_block_tensor(A, B);
    const size_t nb = shape.num_blocks();
    CV const     beta = CV(UC_prefactor), alpha = CV(UAB_prefactor);

    std::vector<eb200_contract_desc> descs;
    descs.reserve(nb);
    [[maybe_unused]] CV *partials = nullptr; // scalar C: one device slot per block
    if constexpr (c_scalar) {
        void *p = nullptr;
        EB200_CHECK(eb200_malloc(&p, std::max<size_t>(nb, 1) * sizeof(CV)));
        partials = static_cast<CV *>(p);
    }
    auto &Am = const_cast<std::remove_const_t<AType> &>(A);
    auto &Bm = const_cast<std::remove_const_t<BType> &>(B);
    for (size_t i = 0; i < nb; i++) {
        if (shape.block_dim(i) == 0) continue; // reference: BlockAlgebra.hpp:42-44
        Range r = shape.block_range(i);
        auto  a = block_operand(Am, i, r);
        auto  b = block_operand(Bm, i, r);
        if constexpr (c_scalar) {
            DeviceTensorView<CV, 0> *none = nullptr;
            descs.push_back(make_contract_desc<ConjA, ConjB>(z, CV(0), none, alpha, a, b, partials + descs.size()));
        } else {
            auto c = block_operand(*C, i, r);
            check_dims(z, &c, a, b);
            descs.push_back(make_contract_desc<ConjA, ConjB>(z, beta, &c, alpha, a, b, c.gpu_data()));
        }
    }
    if (!descs.empty()) EB200_CHECK(eb200_contract_grouped(gpu::get_stream(), std::int64_t(descs.size()), descs.data(), nullptr));
    if constexpr (c_scalar) {
        std::vector<CV> host(descs.size());
        if (!host.empty()) EB200_CHECK(eb200_memcpy_d2h(host.data(), partials, host.size() * sizeof(CV), gpu::get_stream()));
        gpu::stream_wait();
        eb200_free(partials);
        CV c0 = (beta == CV(0)) ? CV(0) : CV(*C) * beta; // reference: BlockAlgebra.hpp:149-155
        *C = std::accumulate(host.begin(), host.end(), c0);
    }
    if (algorithm_choice) *algorithm_choice = detail::GENERIC;
    gpu::op_wait();
}

template <bool ConjA = false, bool ConjB = false, typename AType, typename BType, typename CType, typename... CIndices, typename... AIndices,
          typename... BIndices>
    requires((BlockDeviceTensorLike<AType> || BlockDeviceTensorLike<BType>) && !IsTiledDeviceTensor<std::remove_cvref_t<AType>>::value &&
             !IsTiledDeviceTensor<std::remove_cvref_t<BType>>::value && !IsTiledDeviceTensor<std::remove_cvref_t<CType>>::value)
void einsum(std::tuple<CIndices...> const &C_indices, CType *C, std::tuple<AIndices...> const &A_indices, AType const &A,
            std::tuple<BIndices...> const &B_indices, BType const &B, detail::AlgorithmChoice *algorithm_choice = nullptr) {
    using V = typename detail::BlockValueOf<CType>::type;
    using S = std::conditional_t<IsComplexV<V>, V, double>;
    einsum<ConjA, ConjB>(S(0), C_indices, C, S(1), A_indices, A, B_indices, B, algorithm_choice);
}

} // namespace tensor_algebra

// =================================================================================================================
// linear algebra on block tensors (reference: LinearAlgebra/BlockTensor.hpp)
// =================================================================================================================
namespace linear_algebra {

/// C_i = beta*C_i + alpha*op(A_i)*op(B_i) for every block (reference :43-72)
template <typename T, typename U>
void gemm(char transA, char transB, U alpha, BlockDeviceTensor<T, 2> const &A, BlockDeviceTensor<T, 2> const &B, U beta, BlockDeviceTensor<T, 2> *C) {
    if (A.num_blocks() != B.num_blocks() || A.num_blocks() != C->num_blocks()) throw tensor_compat_error("gemm: Tensors need the same number of blocks.");
    if (!std::strchr("CNTcnt", transA) || !std::strchr("CNTcnt", transB))
        throw std::invalid_argument("The transposition parameters were invalid! They must be either n, c, or t, case insensitive.");
    A.same_blocks(B);
    A.same_blocks(*C);
    const bool tA = std::tolower(transA) != 'n', tB = std::tolower(transB) != 'n';
    const bool cA = std::tolower(transA) == 'c' && IsComplexV<T>, cB = std::tolower(transB) == 'c' && IsComplexV<T>;
    std::vector<eb200_contract_desc> descs;
    for (size_t i = 0; i < A.num_blocks(); i++) {
        const std::int64_t n = A.block_dim(i);
        if (n == 0) continue;
        eb200_contract_desc d{};
        d.n_unique = 3; // 0 = i (row of C), 1 = j (column of C), 2 = k
        d.extent[0] = d.extent[1] = d.extent[2] = n;
        auto fill = [&](eb200_tensor_desc &t, T const *p, int i0, int i1) {
            t.data = const_cast<T *>(p); t.dtype = einsums::detail::DTypeOf<T>::value; t.rank = 2;
            t.index[0] = i0; t.index[1] = i1; t.stride[0] = n; t.stride[1] = 1;
        };
        fill(d.C, C->block(i).gpu_data(), 0, 1);
        fill(d.A, A.block(i).gpu_data(), tA ? 2 : 0, tA ? 0 : 2);
        fill(d.B, B.block(i).gpu_data(), tB ? 1 : 2, tB ? 2 : 1);
        std::complex<double> a{T(alpha)}, b{T(beta)};
        d.alpha[0] = a.real(); d.alpha[1] = a.imag(); d.beta[0] = b.real(); d.beta[1] = b.imag();
        d.conj_a = cA; d.conj_b = cB;
        descs.push_back(d);
    }
    if (!descs.empty()) EB200_CHECK(eb200_contract_grouped(gpu::get_stream(), std::int64_t(descs.size()), descs.data(), nullptr));
    gpu::op_wait();
}
template <bool TransA, bool TransB, typename T, typename U>
void gemm(U alpha, BlockDeviceTensor<T, 2> const &A, BlockDeviceTensor<T, 2> const &B, U beta, BlockDeviceTensor<T, 2> *C) {
    gemm(TransA ? 't' : 'n', TransB ? 't' : 'n', alpha, A, B, beta, C);
}

/// y = beta*y + alpha*op(A)*x with a block-diagonal A and ordinary vectors (reference :84-128)
template <typename T, DeviceTensorLike XType, DeviceTensorLike YType, typename U>
    requires(XType::Rank == 1 && YType::Rank == 1)
void gemv(char transA, U alpha, BlockDeviceTensor<T, 2> const &A, XType const &x, U beta, YType *y) {
    if (x.dim(0) != A.dim() || y->dim(0) != A.dim()) throw dimension_error("gemv: vectors do not match the block tensor");
    const bool was_async = gpu::is_async();
    gpu::set_async(true);
    try {
        for (size_t i = 0; i < A.num_blocks(); i++) {
            if (A.block_dim(i) == 0) continue;
            Range r = A.block_range(i);
            auto  xs = tensor_algebra::detail::block_operand(const_cast<XType &>(x), i, r);
            auto  ys = tensor_algebra::detail::block_operand(*y, i, r);
            gemv(transA, T(alpha), A.block(i), xs, T(beta), &ys);
        }
    } catch (...) {
        gpu::set_async(was_async);
        throw;
    }
    gpu::set_async(was_async);
    gpu::op_wait();
}
template <bool TransA, typename T, DeviceTensorLike XType, DeviceTensorLike YType, typename U>
    requires(XType::Rank == 1 && YType::Rank == 1)
void gemv(U alpha, BlockDeviceTensor<T, 2> const &A, XType const &x, U beta, YType *y) {
    gemv(TransA ? 't' : 'n', alpha, A, x, beta, y);
}

template <typename T, size_t R> void scale(T alpha, BlockDeviceTensor<T, R> *A) { *A *= alpha; }

/// sum over blocks of the unconjugated / conjugated dot products: one reduction over the packed stores
template <typename T, size_t R> T dot(BlockDeviceTensor<T, R> const &A, BlockDeviceTensor<T, R> const &B) {
    A.same_blocks(B);
    T out{};
    if (A.size()) EB200_CHECK(detail::xdot(gpu::get_stream(), A.size(), A.store_data(), 1, B.store_data(), 1, &out));
    return out;
}
template <typename T, size_t R> T true_dot(BlockDeviceTensor<T, R> const &A, BlockDeviceTensor<T, R> const &B) {
    A.same_blocks(B);
    T out{};
    if (!A.size()) return out;
    if constexpr (IsComplexV<T>) EB200_CHECK(detail::xdotc(gpu::get_stream(), A.size(), A.store_data(), 1, B.store_data(), 1, &out));
    else EB200_CHECK(detail::xdot(gpu::get_stream(), A.size(), A.store_data(), 1, B.store_data(), 1, &out));
    return out;
}
template <typename T, size_t R> void axpy(T alpha, BlockDeviceTensor<T, R> const &X, BlockDeviceTensor<T, R> *Y) {
    X.same_blocks(*Y);
    if (X.size()) EB200_CHECK(detail::xaxpy(gpu::get_stream(), X.size(), alpha, X.store_data(), 1, Y->store_data(), 1));
    gpu::op_wait();
}
template <typename T, size_t R> void axpby(T alpha, BlockDeviceTensor<T, R> const &X, T beta, BlockDeviceTensor<T, R> *Y) {
    X.same_blocks(*Y);
    if (X.size()) EB200_CHECK(detail::xaxpby(gpu::get_stream(), X.size(), alpha, X.store_data(), 1, beta, Y->store_data(), 1));
    gpu::op_wait();
}
/// C = beta*C + alpha * A (*) B elementwise inside the blocks (reference :435-445)
template <typename T, size_t R>
void direct_product(T alpha, BlockDeviceTensor<T, R> const &A, BlockDeviceTensor<T, R> const &B, T beta, BlockDeviceTensor<T, R> *C) {
    A.same_blocks(B);
    A.same_blocks(*C);
    if (A.size()) {
        DeviceTensorView<T, 1> a(const_cast<T *>(A.store_data()), Dim<1>{A.size()}, Stride<1>{1}), b(const_cast<T *>(B.store_data()), Dim<1>{A.size()}, Stride<1>{1}),
            c(C->store_data(), Dim<1>{A.size()}, Stride<1>{1});
        direct_product(alpha, a, b, beta, &c);
    }
}

} // namespace linear_algebra

template <typename T, size_t R> BlockDeviceTensor<T, R> &BlockDeviceTensor<T, R>::elementwise(BlockDeviceTensor const &o, bool divide) {
    same_blocks(o);
    if (!_total) return *this;
    if (!divide) {
        EB200_CHECK(linear_algebra::detail::xdirmul(gpu::get_stream(), _total, o._store, 1, _store, 1));
        gpu::op_wait();
    } else {
        EB200_CHECK(linear_algebra::detail::xdirdiv(gpu::get_stream(), _total, o._store, 1, _store, 1));
        gpu::op_wait();
    }
    return *this;
}

} // namespace einsums

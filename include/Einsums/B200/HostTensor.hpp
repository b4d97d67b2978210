// HostTensor.hpp — the minimal host-side tensor the device path needs at its edges: something to assign a DeviceTensor from and
// to copy it back into, in either storage order. The reference's einsums::Tensor (libs/Einsums/Tensor/include/Einsums/Tensor/
// Tensor.hpp) is a large class with its own algebra; only the part its DeviceTensor conversions touch is mirrored here
// (name, dims, strides, data(), operator(), is_row_major(); reference T/Backends/DeviceTensor.hpp:1010-1090 reads exactly those).
// Any type satisfying HostTensorLike (the reference's own Tensor included) can be used in its place.
#pragma once
#include <Einsums/B200/Errors.hpp>

#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace einsums {

template <typename X>
concept HostTensorLike = requires(X const &x) {
    typename std::remove_cvref_t<X>::ValueType;
    { std::remove_cvref_t<X>::Rank } -> std::convertible_to<size_t>;
    { x.data() } -> std::convertible_to<typename std::remove_cvref_t<X>::ValueType const *>;
    { x.dim(0) } -> std::convertible_to<std::int64_t>;
    { x.stride(0) } -> std::convertible_to<std::int64_t>;
} && !requires(X const &x) { x.gpu_data(); };

/// Dense host tensor, row-major (C order, the device layout) or column-major (the reference's BLAS-facing default).
template <typename T, size_t R> class Tensor {
  public:
    using ValueType = T;
    using value_type = T;
    static constexpr size_t Rank = R;

    Tensor() = default;
    template <typename... Dims>
        requires(sizeof...(Dims) == R && (std::is_integral_v<Dims> && ...))
    Tensor(bool row_major, std::string name, Dims... dims) : _name(std::move(name)), _dims{static_cast<std::int64_t>(dims)...}, _row_major(row_major) {
        layout();
    }
    template <typename... Dims>
        requires(sizeof...(Dims) == R && (std::is_integral_v<Dims> && ...))
    explicit Tensor(std::string name, Dims... dims) : Tensor(true, std::move(name), dims...) {}

    std::int64_t dim(std::ptrdiff_t i) const { return _dims[i < 0 ? i + std::ptrdiff_t(R) : i]; }
    std::int64_t stride(std::ptrdiff_t i) const { return _strides[i < 0 ? i + std::ptrdiff_t(R) : i]; }
    std::array<std::int64_t, R> const &dims() const { return _dims; }
    std::array<std::int64_t, R> const &strides() const { return _strides; }
    std::int64_t                       size() const { return std::int64_t(_data.size()); }
    std::string const                 &name() const { return _name; }
    bool                               is_row_major() const { return _row_major; }
    T                                 *data() { return _data.data(); }
    T const                           *data() const { return _data.data(); }
    template <typename... Idx>
        requires(sizeof...(Idx) == R)
    T &operator()(Idx... idx) {
        return _data[offset(idx...)];
    }
    template <typename... Idx>
        requires(sizeof...(Idx) == R)
    T const &operator()(Idx... idx) const {
        return _data[offset(idx...)];
    }
    void zero() { std::fill(_data.begin(), _data.end(), T{}); }
    void set_all(T v) { std::fill(_data.begin(), _data.end(), v); }

  private:
    template <typename... Idx> std::int64_t offset(Idx... idx) const {
        std::array<std::int64_t, R> ix{static_cast<std::int64_t>(idx)...};
        std::int64_t                off = 0;
        for (size_t i = 0; i < R; i++) {
            if (ix[i] < 0) ix[i] += _dims[i];
            if (ix[i] < 0 || ix[i] >= _dims[i]) throw std::out_of_range("Tensor index out of range");
            off += ix[i] * _strides[i];
        }
        return off;
    }
    void layout() {
        std::int64_t acc = 1;
        if (_row_major)
            for (std::ptrdiff_t i = std::ptrdiff_t(R) - 1; i >= 0; i--) { _strides[i] = acc; acc *= _dims[i]; }
        else
            for (size_t i = 0; i < R; i++) { _strides[i] = acc; acc *= _dims[i]; }
        if (acc < 0) throw dimension_error("Tensor: negative dimension");
        _data.assign(size_t(acc), T{});
    }
    std::string                 _name{"(unnamed)"};
    std::array<std::int64_t, R> _dims{};
    std::array<std::int64_t, R> _strides{};
    std::vector<T>              _data;
    bool                        _row_major = true;
};

} // namespace einsums

// DeviceTensor.hpp — owning row-major device tensor and non-owning strided view for the B200 backend.
// Mirrors the part of the reference's DeviceTensor / DeviceTensorView API that the einsum / permute / linear_algebra
// hot path and its tests use (reference: libs/Einsums/Tensor/include/Einsums/Tensor/DeviceTensor.hpp:68-316,
// 326-1238, 1247-1730; Backends/DeviceTensor.hpp:27-291; Backends/DeviceTensorView.hpp:280-362).
// Differences by design: dims/strides are passed to kernels by value (no device-side copies, no per-view
// hipMalloc + hipMemcpy + device sync); storage comes from eb200_malloc.
#pragma once
#include <Einsums/B200/HostTensor.hpp>
#include <Einsums/B200/Runtime.hpp>

#include <array>
#include <complex>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <type_traits>
#include <vector>

namespace einsums {

namespace detail {
enum HostToDeviceMode { UNKNOWN, DEV_ONLY, MAPPED, PINNED };

template <typename T> struct DTypeOf;
template <> struct DTypeOf<float> { static constexpr int value = EB200_F32; };
template <> struct DTypeOf<double> { static constexpr int value = EB200_F64; };
template <> struct DTypeOf<std::complex<float>> { static constexpr int value = EB200_C64; };
template <> struct DTypeOf<std::complex<double>> { static constexpr int value = EB200_C128; };

template <typename T> struct IsComplex : std::false_type {};
template <typename T> struct IsComplex<std::complex<T>> : std::true_type {};
} // namespace detail
template <typename T> inline constexpr bool IsComplexV = detail::IsComplex<std::remove_cvref_t<T>>::value;

template <size_t Rank> struct Dim : std::array<std::int64_t, Rank> {
    template <typename... Args> constexpr Dim(Args... args) : std::array<std::int64_t, Rank>{static_cast<std::int64_t>(args)...} {}
};
template <size_t Rank> struct Stride : std::array<std::int64_t, Rank> {
    template <typename... Args> constexpr Stride(Args... args) : std::array<std::int64_t, Rank>{static_cast<std::int64_t>(args)...} {}
};
template <size_t Rank> struct Offset : std::array<std::int64_t, Rank> {
    template <typename... Args> constexpr Offset(Args... args) : std::array<std::int64_t, Rank>{static_cast<std::int64_t>(args)...} {}
};
template <typename... Args> Dim(Args...) -> Dim<sizeof...(Args)>;
template <typename... Args> Stride(Args...) -> Stride<sizeof...(Args)>;
template <typename... Args> Offset(Args...) -> Offset<sizeof...(Args)>;

struct Range : std::array<std::int64_t, 2> {
    template <typename A, typename B> constexpr Range(A a, B b) : std::array<std::int64_t, 2>{static_cast<std::int64_t>(a), static_cast<std::int64_t>(b)} {}
};
struct AllT {};
static inline AllT All;

/// Proxy for one element living on the device: reads and writes are 1-element copies
/// (reference: HostDevReference, DeviceTensor.hpp:68-316).
template <typename T> class HostDevReference {
    T *_dev;

  public:
    explicit HostDevReference(T *dev) : _dev(dev) {}
    T get() const {
        T v;
        gpu::stream_wait();
        EB200_CHECK(eb200_memcpy_d2h(&v, _dev, sizeof(T), gpu::get_stream()));
        return v;
    }
    operator T() const { return get(); }
    HostDevReference &operator=(T const &v) {
        EB200_CHECK(eb200_memcpy_h2d(_dev, &v, sizeof(T), gpu::get_stream()));
        return *this;
    }
    HostDevReference &operator=(HostDevReference const &o) { return *this = o.get(); }
    template <typename U> HostDevReference &operator+=(U const &v) { return *this = get() + T(v); }
    template <typename U> HostDevReference &operator-=(U const &v) { return *this = get() - T(v); }
    template <typename U> HostDevReference &operator*=(U const &v) { return *this = get() * T(v); }
    template <typename U> HostDevReference &operator/=(U const &v) { return *this = get() / T(v); }
};

template <typename T, size_t R> class DeviceTensorView;

namespace detail {
/// Shared read-only interface of tensors and views, used by the algebra layer.
template <typename Derived, typename T, size_t R> class DeviceTensorCommon {
  public:
    using ValueType = T;
    using value_type = T;
    using host_datatype = T;
    static constexpr size_t Rank = R;
    static constexpr size_t rank = R;

    std::int64_t dim(std::ptrdiff_t i) const { return _dims[i < 0 ? i + std::ptrdiff_t(R) : i]; }
    std::int64_t stride(std::ptrdiff_t i) const { return _strides[i < 0 ? i + std::ptrdiff_t(R) : i]; }
    Dim<R> const    &dims() const { return _dims; }
    Stride<R> const &strides() const { return _strides; }
    std::int64_t     size() const {
        std::int64_t s = 1;
        for (size_t i = 0; i < R; i++) s *= _dims[i];
        return s;
    }
    std::string const &name() const { return _name; }
    void               set_name(std::string const &n) { _name = n; }
    T                 *gpu_data() { return _data; }
    T const           *gpu_data() const { return _data; }

    /// reference: DeviceTensor.hpp:1106-1116
    bool is_row_major() const {
        if constexpr (R < 2) return true;
        else {
            if (stride(0) > stride(-1)) return true;
            if (stride(0) == stride(-1)) return dim(0) > dim(-1);
            return false;
        }
    }
    /// reference: DeviceTensor.hpp:1130-1155
    bool is_totally_vectorable(std::int64_t *incx = nullptr) const {
        if constexpr (R == 0) {
            if (incx) *incx = 0;
            return false;
        } else if constexpr (R == 1) {
            if (incx) *incx = stride(0);
            return true;
        } else {
            if (stride(0) < stride(-1)) {
                if (incx) *incx = stride(0);
                return dim(-1) * stride(-1) == size() * stride(0);
            }
            if (incx) *incx = stride(-1);
            return dim(0) * stride(0) == size() * stride(-1);
        }
    }
    bool is_contiguous_row_major() const {
        std::int64_t expect = 1;
        for (std::ptrdiff_t i = std::ptrdiff_t(R) - 1; i >= 0; i--) {
            if (_dims[i] != 1 && _strides[i] != expect) return false;
            expect *= _dims[i];
        }
        return true;
    }

    template <typename... Idx>
        requires(sizeof...(Idx) == R && (std::is_integral_v<Idx> && ...))
    HostDevReference<T> operator()(Idx... idx) {
        return HostDevReference<T>(_data + offset_of(idx...));
    }
    template <typename... Idx>
        requires(sizeof...(Idx) == R && (std::is_integral_v<Idx> && ...))
    T operator()(Idx... idx) const {
        return HostDevReference<T>(const_cast<T *>(_data) + offset_of(idx...)).get();
    }

    /// Copy the elements out into a dense row-major host buffer.
    void write(T *host) const {
        gpu::stream_wait();
        if (is_contiguous_row_major()) {
            EB200_CHECK(eb200_memcpy_d2h(host, _data, sizeof(T) * size(), gpu::get_stream()));
            return;
        }
        DeviceScratch tmp(sizeof(T) * size());
        Stride<R> dense = dense_strides();
        EB200_CHECK(eb200_copy_strided(gpu::get_stream(), DTypeOf<T>::value, int(R), _dims.data(), _data, _strides.data(), tmp.ptr, dense.data()));
        EB200_CHECK(eb200_memcpy_d2h(host, tmp.ptr, sizeof(T) * size(), gpu::get_stream()));
    }
    std::vector<T> to_host() const {
        std::vector<T> v(size());
        write(v.data());
        return v;
    }
    /// Fill from a dense row-major host buffer.
    void read(T const *host) {
        if (is_contiguous_row_major()) {
            EB200_CHECK(eb200_memcpy_h2d(_data, host, sizeof(T) * size(), gpu::get_stream()));
            return;
        }
        DeviceScratch tmp(sizeof(T) * size());
        Stride<R> dense = dense_strides();
        EB200_CHECK(eb200_memcpy_h2d(tmp.ptr, host, sizeof(T) * size(), gpu::get_stream()));
        EB200_CHECK(eb200_copy_strided(gpu::get_stream(), DTypeOf<T>::value, int(R), _dims.data(), tmp.ptr, dense.data(), _data, _strides.data()));
        gpu::stream_wait();
    }
    void read(std::vector<T> const &host) {
        if (std::int64_t(host.size()) != size()) throw dimension_error("read: host buffer size does not match the tensor");
        read(host.data());
    }

    /// set every element (strided views included)
    void set_all(T value) {
        if (size() == 0) return;
        DeviceScratch one(sizeof(T));
        EB200_CHECK(eb200_memcpy_h2d(one.ptr, &value, sizeof(T), gpu::get_stream()));
        Stride<R> zero_strides;
        for (size_t i = 0; i < R; i++) zero_strides[i] = 0;
        EB200_CHECK(eb200_copy_strided(gpu::get_stream(), DTypeOf<T>::value, int(R), _dims.data(), one.ptr, zero_strides.data(), _data, _strides.data()));
        gpu::stream_wait();
    }
    void zero() {
        if (is_contiguous_row_major()) {
            EB200_CHECK(eb200_memset_zero(_data, sizeof(T) * size(), gpu::get_stream()));
            gpu::op_wait();
        } else set_all(T{0});
    }

    Stride<R> dense_strides() const {
        Stride<R> s;
        std::int64_t acc = 1;
        for (std::ptrdiff_t i = std::ptrdiff_t(R) - 1; i >= 0; i--) {
            s[i] = acc;
            acc *= _dims[i];
        }
        return s;
    }

  protected:
    struct DeviceScratch {
        void *ptr = nullptr;
        explicit DeviceScratch(size_t bytes) { EB200_CHECK(eb200_malloc(&ptr, bytes ? bytes : 1)); }
        ~DeviceScratch() { eb200_free(ptr); }
    };
    template <typename... Idx> std::int64_t offset_of(Idx... idx) const {
        std::array<std::int64_t, R> ix{static_cast<std::int64_t>(idx)...};
        std::int64_t off = 0;
        for (size_t i = 0; i < R; i++) {
            std::int64_t v = ix[i] < 0 ? ix[i] + _dims[i] : ix[i];
            if (v < 0 || v >= _dims[i]) throw std::out_of_range("DeviceTensor index out of range");
            off += v * _strides[i];
        }
        return off;
    }
    std::string _name{"(unnamed)"};
    Dim<R>      _dims{};
    Stride<R>   _strides{};
    T          *_data = nullptr;
};
} // namespace detail

/// Owning, row-major device tensor.
template <typename T, size_t R> class DeviceTensor : public detail::DeviceTensorCommon<DeviceTensor<T, R>, T, R> {
    using Base = detail::DeviceTensorCommon<DeviceTensor<T, R>, T, R>;

  public:
    using Base::operator();
    DeviceTensor() = default;
    template <typename... Dims>
        requires(sizeof...(Dims) == R && (std::is_integral_v<Dims> && ...))
    DeviceTensor(std::string name, detail::HostToDeviceMode mode, Dims... dims) : _mode(mode) {
        this->_name = std::move(name);
        this->_dims = Dim<R>{dims...};
        allocate();
    }
    template <typename... Dims>
        requires(sizeof...(Dims) == R && (std::is_integral_v<Dims> && ...))
    explicit DeviceTensor(std::string name, Dims... dims) : DeviceTensor(std::move(name), detail::DEV_ONLY, dims...) {}
    explicit DeviceTensor(Dim<R> dims, detail::HostToDeviceMode mode = detail::DEV_ONLY) : _mode(mode) {
        this->_dims = dims;
        allocate();
    }
    DeviceTensor(DeviceTensor const &o) : Base(o), _mode(o._mode) {
        this->_data = nullptr;
        allocate();
        EB200_CHECK(eb200_memcpy_d2d(this->_data, o._data, sizeof(T) * this->size(), gpu::get_stream()));
        gpu::stream_wait();
    }
    DeviceTensor(DeviceTensor &&o) noexcept : Base(o), _mode(o._mode), _host_data(o._host_data), _gpu_dims(o._gpu_dims), _gpu_strides(o._gpu_strides) {
        o._data = nullptr;
        o._host_data = nullptr;
        o._gpu_dims = o._gpu_strides = nullptr;
    }
    /// dense copy of a view
    explicit DeviceTensor(DeviceTensorView<T, R> const &v);
    /// from a host tensor of either storage order (reference: DeviceTensor(Tensor<T, rank> const &, mode = MAPPED),
    /// T/DeviceTensor.hpp:1097); the device copy is always row-major
    template <HostTensorLike H>
        requires(std::remove_cvref_t<H>::Rank == R)
    explicit DeviceTensor(H const &host, detail::HostToDeviceMode mode = detail::MAPPED) : _mode(mode) {
        this->_name = host.name();
        for (size_t i = 0; i < R; i++) this->_dims[i] = host.dim(i);
        allocate();
        *this = host;
    }
    DeviceTensor &operator=(DeviceTensor const &o) {
        if (this == &o) return *this;
        if (this->_dims != o._dims) {
            release();
            this->_dims = o._dims;
            allocate();
        }
        EB200_CHECK(eb200_memcpy_d2d(this->_data, o._data, sizeof(T) * this->size(), gpu::get_stream()));
        gpu::stream_wait();
        return *this;
    }
    DeviceTensor &operator=(DeviceTensor &&o) noexcept {
        if (this != &o) {
            release();
            static_cast<Base &>(*this) = o;
            _mode = o._mode;
            _host_data = o._host_data;
            _gpu_dims = o._gpu_dims;
            _gpu_strides = o._gpu_strides;
            o._data = nullptr;
            o._host_data = nullptr;
            o._gpu_dims = o._gpu_strides = nullptr;
        }
        return *this;
    }
    /// A = hostTensor with a layout change when the host tensor is not dense row-major (reference: T/Backends/DeviceTensor.hpp:
    /// 1010-1090 — raw upload, then the copy_to_tensor kernel with both stride sets). Shapes must agree (reference resizes only
    /// an uninitialised tensor; here a default-constructed one takes the host tensor's shape).
    template <HostTensorLike H>
        requires(std::remove_cvref_t<H>::Rank == R)
    DeviceTensor &operator=(H const &host) {
        bool same = true;
        for (size_t i = 0; i < R; i++) same = same && this->_dims[i] == host.dim(i);
        if (!same) {
            if (this->_data) throw dimension_error("DeviceTensor = host tensor: shapes differ");
            release();
            for (size_t i = 0; i < R; i++) this->_dims[i] = host.dim(i);
            allocate();
        }
        if (this->size() == 0) return *this;
        Stride<R> hs;
        bool      dense_rm = true;
        std::int64_t expect = 1, lo = 0, hi = 0;
        for (std::ptrdiff_t i = std::ptrdiff_t(R) - 1; i >= 0; i--) {
            hs[i] = host.stride(i);
            if (host.dim(i) != 1 && hs[i] != expect) dense_rm = false;
            expect *= host.dim(i);
            (hs[i] < 0 ? lo : hi) += (host.dim(i) - 1) * hs[i];
        }
        if (dense_rm) {
            if (_host_data) {
                gpu::stream_wait();
                std::memcpy(_host_data, host.data(), sizeof(T) * size_t(this->size()));
            } else EB200_CHECK(eb200_memcpy_h2d(this->_data, host.data(), sizeof(T) * size_t(this->size()), gpu::get_stream()));
            return *this;
        }
        // raw upload of the span the host tensor occupies, then one strided copy into the row-major device layout
        const size_t span = size_t(hi - lo + 1);
        typename Base::DeviceScratch tmp(sizeof(T) * span);
        EB200_CHECK(eb200_memcpy_h2d(tmp.ptr, host.data() + lo, sizeof(T) * span, gpu::get_stream()));
        EB200_CHECK(eb200_copy_strided(gpu::get_stream(), detail::DTypeOf<T>::value, int(R), this->_dims.data(), static_cast<T *>(tmp.ptr) - lo, hs.data(),
                                       this->_data, this->_strides.data()));
        gpu::stream_wait();
        return *this;
    }
    /// hostTensor = A: the inverse conversion (reference: Tensor<T, rank>(DeviceTensor) / DeviceTensor::operator Tensor, :1099-1104)
    template <HostTensorLike H>
        requires(std::remove_cvref_t<H>::Rank == R)
    void copy_to(H *host) const {
        for (size_t i = 0; i < R; i++)
            if (this->_dims[i] != host->dim(i)) throw dimension_error("DeviceTensor -> host tensor: shapes differ");
        if (this->size() == 0) return;
        Stride<R>    hs;
        bool         dense_rm = true;
        std::int64_t expect = 1, lo = 0, hi = 0;
        for (std::ptrdiff_t i = std::ptrdiff_t(R) - 1; i >= 0; i--) {
            hs[i] = host->stride(i);
            if (host->dim(i) != 1 && hs[i] != expect) dense_rm = false;
            expect *= host->dim(i);
            (hs[i] < 0 ? lo : hi) += (host->dim(i) - 1) * hs[i];
        }
        if (dense_rm) {
            this->write(host->data());
            return;
        }
        const size_t span = size_t(hi - lo + 1);
        if (span != size_t(this->size())) throw tensor_compat_error("DeviceTensor -> host tensor: the host tensor must be dense (any axis order)");
        typename Base::DeviceScratch tmp(sizeof(T) * span);
        gpu::stream_wait();
        EB200_CHECK(eb200_copy_strided(gpu::get_stream(), detail::DTypeOf<T>::value, int(R), this->_dims.data(), this->_data, this->_strides.data(),
                                       static_cast<T *>(tmp.ptr) - lo, hs.data()));
        EB200_CHECK(eb200_memcpy_d2h(host->data() + lo, tmp.ptr, sizeof(T) * span, gpu::get_stream()));
    }
    explicit operator Tensor<T, R>() const {
        Tensor<T, R> out = make_host(std::make_index_sequence<R>{});
        copy_to(&out);
        return out;
    }

    /// host-side storage of a MAPPED / PINNED tensor (nullptr for DEV_ONLY); reference T/DeviceTensor.hpp:614-627
    T                       *data() { return _host_data; }
    T const                 *data() const { return _host_data; }
    detail::HostToDeviceMode mode() const { return _mode; }
    /// dims / strides as device-resident size_t arrays for user kernels (reference T/DeviceTensor.hpp:1000-1053). The backend's own
    /// kernels take them by value, so the device copies are made on first request only.
    size_t const *gpu_dims() const {
        ensure_gpu_meta();
        return _gpu_dims;
    }
    size_t const *gpu_strides() const {
        ensure_gpu_meta();
        return _gpu_strides;
    }
    DeviceTensor &operator=(std::vector<T> const &host) {
        this->read(host);
        return *this;
    }
    ~DeviceTensor() { release(); }

    DeviceTensor &operator*=(T s) {
        scale_inplace(s);
        return *this;
    }
    DeviceTensor &operator+=(DeviceTensor const &o) {
        axpy_inplace(T{1}, o);
        return *this;
    }
    DeviceTensor &operator-=(DeviceTensor const &o) {
        axpy_inplace(T{-1}, o);
        return *this;
    }

    DeviceTensorView<T, 1> to_rank_1_view() const;

    /// slicing: any mix of integers, Range{a,b} and All; the view keeps one axis per Range/All argument
    template <typename... Args>
        requires(sizeof...(Args) == R && !(std::is_integral_v<std::remove_cvref_t<Args>> && ...))
    auto operator()(Args... args);

  private:
    /// storage by mode (reference: T/Backends/DeviceTensor.hpp:62-116): DEV_ONLY = device memory under the byte budget; PINNED = page-locked
    /// host memory the device reads in place; MAPPED = ordinary host memory registered with the driver and aliased on the device
    void allocate() {
        std::int64_t acc = 1;
        for (std::ptrdiff_t i = std::ptrdiff_t(R) - 1; i >= 0; i--) {
            if (this->_dims[i] < 0) throw dimension_error("DeviceTensor: negative dimension");
            this->_strides[i] = acc;
            acc *= this->_dims[i];
        }
        (void)gpu::get_stream(); // makes sure the runtime is up
        void        *p = nullptr;
        const size_t bytes = sizeof(T) * size_t(acc);
        switch (_mode) {
        case detail::DEV_ONLY:
            EB200_CHECK(eb200_malloc(&p, bytes));
            break;
        case detail::PINNED: {
            void *h = nullptr;
            EB200_CHECK(eb200_malloc_host(&h, bytes ? bytes : sizeof(T)));
            int rc = eb200_host_get_device_pointer(h, &p);
            if (rc) {
                eb200_free_host(h);
                EB200_CHECK(rc);
            }
            _host_data = static_cast<T *>(h);
            break;
        }
        case detail::MAPPED: {
            T  *h = new T[acc ? size_t(acc) : 1]();
            int rc = eb200_host_register(h, bytes ? bytes : sizeof(T), &p);
            if (rc) {
                delete[] h;
                EB200_CHECK(rc);
            }
            _host_data = h;
            break;
        }
        default: throw uninitialized_error("DeviceTensor: unknown host-to-device mode");
        }
        this->_data = static_cast<T *>(p);
    }
    void release() {
        if (this->_data || _host_data || _gpu_dims) eb200_device_synchronize();
        if (_mode == detail::PINNED && _host_data) eb200_free_host(_host_data);
        else if (_mode == detail::MAPPED && _host_data) {
            eb200_host_unregister(_host_data);
            delete[] _host_data;
        } else if (this->_data) eb200_free(this->_data);
        if (_gpu_dims) eb200_free(_gpu_dims);
        this->_data = nullptr;
        _host_data = nullptr;
        _gpu_dims = _gpu_strides = nullptr;
    }
    void ensure_gpu_meta() const {
        if (_gpu_dims) return;
        void  *p = nullptr;
        size_t host[2 * (R ? R : 1)];
        for (size_t i = 0; i < R; i++) {
            host[i] = size_t(this->_dims[i]);
            host[R + i] = size_t(this->_strides[i]);
        }
        EB200_CHECK(eb200_malloc(&p, sizeof(host)));
        EB200_CHECK(eb200_memcpy_h2d(p, host, sizeof(host), gpu::get_stream()));
        _gpu_dims = static_cast<size_t *>(p);
        _gpu_strides = _gpu_dims + R;
    }
    template <size_t... I> Tensor<T, R> make_host(std::index_sequence<I...>) const { return Tensor<T, R>(true, this->_name, this->_dims[I]...); }
    void scale_inplace(T s);
    void axpy_inplace(T a, DeviceTensor const &x);
    detail::HostToDeviceMode _mode = detail::DEV_ONLY;
    T                       *_host_data = nullptr;
    mutable size_t          *_gpu_dims = nullptr, *_gpu_strides = nullptr;
};

/// Rank-0 tensor: one element on the device, convertible to and assignable from its value type.
template <typename T> class DeviceTensor<T, 0> : public detail::DeviceTensorCommon<DeviceTensor<T, 0>, T, 0> {
  public:
    DeviceTensor() { allocate(); }
    explicit DeviceTensor(std::string name, detail::HostToDeviceMode = detail::DEV_ONLY) {
        this->_name = std::move(name);
        allocate();
    }
    DeviceTensor(DeviceTensor const &o) {
        allocate();
        *this = T(o);
    }
    ~DeviceTensor() {
        if (this->_data) {
            eb200_device_synchronize();
            eb200_free(this->_data);
        }
    }
    operator T() const { return HostDevReference<T>(const_cast<T *>(this->_data)).get(); }
    DeviceTensor &operator=(T v) {
        HostDevReference<T>(this->_data) = v;
        return *this;
    }
    DeviceTensor &operator=(DeviceTensor const &o) { return *this = T(o); }
    DeviceTensor &operator*=(T v) { return *this = T(*this) * v; }
    DeviceTensor &operator+=(T v) { return *this = T(*this) + v; }
    DeviceTensor &operator-=(T v) { return *this = T(*this) - v; }

  private:
    void allocate() {
        (void)gpu::get_stream();
        void *p = nullptr;
        EB200_CHECK(eb200_malloc(&p, sizeof(T)));
        this->_data = static_cast<T *>(p);
        T z{0};
        EB200_CHECK(eb200_memcpy_h2d(p, &z, sizeof(T), gpu::get_stream()));
    }
};

/// Non-owning strided view. Borrowed storage must outlive the view.
template <typename T, size_t R> class DeviceTensorView : public detail::DeviceTensorCommon<DeviceTensorView<T, R>, T, R> {
  public:
    DeviceTensorView() = default;
    DeviceTensorView(T *data, Dim<R> dims, Stride<R> strides, std::string name = "(view)") {
        this->_data = data;
        this->_dims = dims;
        this->_strides = strides;
        this->_name = std::move(name);
    }
    /// reshape view of a contiguous tensor/view (reference: DeviceTensorView(tensor, Dim<R>{...})); one dim may be -1
    template <typename Other>
    DeviceTensorView(Other &other, Dim<R> dims) {
        if (!other.is_contiguous_row_major()) throw tensor_compat_error("DeviceTensorView: reshaping a non-contiguous tensor");
        std::int64_t known = 1;
        std::ptrdiff_t wild = -1;
        for (size_t i = 0; i < R; i++) {
            if (dims[i] < 0) wild = std::ptrdiff_t(i);
            else known *= dims[i];
        }
        if (wild >= 0) dims[wild] = known ? other.size() / known : 0;
        this->_dims = dims;
        if (this->size() > other.size()) throw dimension_error("DeviceTensorView: view is larger than the viewed tensor");
        this->_strides = this->dense_strides();
        this->_data = const_cast<T *>(other.gpu_data());
        this->_name = other.name() + " (view)";
    }
    /// explicit dims + strides over the same storage (what einsum_do_matrix_product builds, reference Dispatch.hpp:802-808)
    template <typename Other>
    DeviceTensorView(Other &other, Dim<R> dims, Stride<R> strides) {
        this->_dims = dims;
        this->_strides = strides;
        this->_data = const_cast<T *>(other.gpu_data());
        this->_name = other.name() + " (view)";
    }
    /// window with an offset into a tensor of the same rank (strides inherited)
    template <typename Other>
        requires(Other::Rank == R)
    DeviceTensorView(Other &other, Dim<R> dims, Offset<R> offset) {
        std::int64_t off = 0;
        for (size_t i = 0; i < R; i++) {
            if (offset[i] < 0 || offset[i] + dims[i] > other.dim(i)) throw dimension_error("DeviceTensorView: window exceeds the viewed tensor");
            off += offset[i] * other.stride(i);
            this->_strides[i] = other.stride(i);
        }
        this->_dims = dims;
        this->_data = const_cast<T *>(other.gpu_data()) + off;
        this->_name = other.name() + " (view)";
    }
    DeviceTensorView &operator*=(T s);
    DeviceTensorView<T, 1> to_rank_1_view() const {
        std::int64_t inc = 1;
        if (!this->is_totally_vectorable(&inc)) throw tensor_compat_error("to_rank_1_view: view is not totally vectorable");
        return DeviceTensorView<T, 1>(const_cast<T *>(this->_data), Dim<1>{this->size()}, Stride<1>{inc});
    }
    /// assignment from a tensor/view of the same shape copies elements
    template <typename Other>
        requires(Other::Rank == R)
    DeviceTensorView &assign(Other const &o) {
        for (size_t i = 0; i < R; i++)
            if (o.dim(i) != this->dim(i)) throw dimension_error("DeviceTensorView::assign: shapes differ");
        EB200_CHECK(eb200_copy_strided(gpu::get_stream(), detail::DTypeOf<T>::value, int(R), this->_dims.data(), o.gpu_data(), o.strides().data(),
                                       this->_data, this->_strides.data()));
        gpu::op_wait();
        return *this;
    }
};

template <typename Other, size_t R> DeviceTensorView(Other &, Dim<R>) -> DeviceTensorView<typename Other::ValueType, R>;
template <typename Other, size_t R> DeviceTensorView(Other &, Dim<R>, Stride<R>) -> DeviceTensorView<typename Other::ValueType, R>;
template <typename Other, size_t R> DeviceTensorView(Other &, Dim<R>, Offset<R>) -> DeviceTensorView<typename Other::ValueType, R>;

// ---- out-of-line members ------------------------------------------------------------------
template <typename T, size_t R> DeviceTensor<T, R>::DeviceTensor(DeviceTensorView<T, R> const &v) {
    this->_dims = v.dims();
    this->_name = v.name();
    allocate();
    EB200_CHECK(eb200_copy_strided(gpu::get_stream(), detail::DTypeOf<T>::value, int(R), this->_dims.data(), v.gpu_data(), v.strides().data(),
                                   this->_data, this->_strides.data()));
    gpu::stream_wait();
}

template <typename T, size_t R> DeviceTensorView<T, 1> DeviceTensor<T, R>::to_rank_1_view() const {
    return DeviceTensorView<T, 1>(const_cast<T *>(this->_data), Dim<1>{this->size()}, Stride<1>{1});
}

namespace detail {
template <typename A> inline std::int64_t slice_begin(A const &a, std::int64_t) {
    if constexpr (std::is_same_v<A, Range>) return a[0];
    else if constexpr (std::is_same_v<A, AllT>) return 0;
    else return static_cast<std::int64_t>(a);
}
template <typename A> inline std::int64_t slice_extent(A const &a, std::int64_t dim) {
    if constexpr (std::is_same_v<A, Range>) return (a[1] < 0 ? a[1] + dim : a[1]) - a[0];
    else if constexpr (std::is_same_v<A, AllT>) return dim;
    else return -1; // integer: axis dropped
}
template <typename... Args> constexpr size_t kept_axes() { return ((std::is_integral_v<std::remove_cvref_t<Args>> ? 0 : 1) + ... + 0); }
} // namespace detail

template <typename T, size_t R>
template <typename... Args>
    requires(sizeof...(Args) == R && !(std::is_integral_v<std::remove_cvref_t<Args>> && ...))
auto DeviceTensor<T, R>::operator()(Args... args) {
    constexpr size_t             VR = detail::kept_axes<Args...>();
    std::array<std::int64_t, R> begin{}, ext{};
    size_t                      ax = 0;
    ((begin[ax] = detail::slice_begin(args, this->_dims[ax]), ext[ax] = detail::slice_extent(args, this->_dims[ax]), ax++), ...);
    Dim<VR>      d;
    Stride<VR>   s;
    std::int64_t off = 0;
    size_t       o = 0;
    for (size_t i = 0; i < R; i++) {
        std::int64_t b = begin[i] < 0 ? begin[i] + this->_dims[i] : begin[i];
        if (b < 0 || b + (ext[i] < 0 ? 1 : ext[i]) > this->_dims[i]) throw std::out_of_range("DeviceTensor slice out of range");
        off += b * this->_strides[i];
        if (ext[i] >= 0) {
            d[o] = ext[i];
            s[o] = this->_strides[i];
            o++;
        }
    }
    return DeviceTensorView<T, VR>(this->_data + off, d, s, this->_name + " (view)");
}

// ---- creation helpers used by tests and examples (reference: TensorUtilities/CreateRandomTensor.hpp:57-84) ------------
template <typename T = double, typename... Dims> DeviceTensor<T, sizeof...(Dims)> create_random_gpu_tensor(std::string name, Dims... dims) {
    DeviceTensor<T, sizeof...(Dims)> t(std::move(name), detail::DEV_ONLY, dims...);
    std::vector<T>                   host(size_t(t.size()));
    static thread_local std::mt19937_64 eng(20251017);
    if constexpr (IsComplexV<T>) {
        std::uniform_real_distribution<typename T::value_type> d(-1, 1);
        for (auto &v : host) v = T(d(eng), d(eng));
    } else {
        std::uniform_real_distribution<T> d(-1, 1);
        for (auto &v : host) v = d(eng);
    }
    t.read(host);
    return t;
}
template <typename T = double, typename... Dims> DeviceTensor<T, sizeof...(Dims)> create_incremented_gpu_tensor(std::string name, Dims... dims) {
    DeviceTensor<T, sizeof...(Dims)> t(std::move(name), detail::DEV_ONLY, dims...);
    std::vector<T>                   host(size_t(t.size()));
    for (size_t i = 0; i < host.size(); i++) host[i] = T(double(i));
    t.read(host);
    return t;
}
template <typename T = double, typename... Dims> DeviceTensor<T, sizeof...(Dims)> create_zero_gpu_tensor(std::string name, Dims... dims) {
    DeviceTensor<T, sizeof...(Dims)> t(std::move(name), detail::DEV_ONLY, dims...);
    t.zero();
    return t;
}
template <typename TensorT> void zero(TensorT &t) { t.zero(); }

/// true for TiledDeviceTensor (specialised in TiledTensor.hpp); declared here so that the block-tensor overloads can step aside for the
/// mixed block x tiled algebra (BlockTileAlgebra.hpp) without including the tiled header
template <typename X> struct IsTiledDeviceTensor : std::false_type {};
/// likewise for BlockDeviceTensor (specialised in BlockTensor.hpp)
template <typename X> struct IsBlockDeviceTensor : std::false_type {};

template <typename X>
concept DeviceTensorLike = requires(X const &x) {
    typename std::remove_cvref_t<X>::ValueType;
    { std::remove_cvref_t<X>::Rank } -> std::convertible_to<size_t>;
    { x.gpu_data() };
    { x.dim(0) } -> std::convertible_to<std::int64_t>;
    { x.stride(0) } -> std::convertible_to<std::int64_t>;
};

} // namespace einsums

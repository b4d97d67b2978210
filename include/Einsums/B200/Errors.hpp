// Errors.hpp — exception types of the host shim and the status-code -> exception mapping.
// The C ABI never throws; this header turns eb200_status back into the exception classes the reference
// throws on the same conditions (reference: libs/Einsums/Errors/include/Einsums/Errors/Error.hpp:251,361-390,
// libs/Einsums/hipBLASVendor/src/gemm.hip:39-80, TensorAlgebra/Backends/Dispatch.hpp:125-127).
#pragma once
#include <einsums_b200.h>

#include <stdexcept>
#include <string>

namespace einsums {

struct dimension_error : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct tensor_compat_error : std::logic_error { using std::logic_error::logic_error; };
struct rank_error : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct bad_logic : std::logic_error { using std::logic_error::logic_error; };
struct not_enough_args : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct uninitialized_error : std::runtime_error { using std::runtime_error::runtime_error; };
/// any CUDA runtime failure reported by the backend (reference: hip_exception<code>)
struct gpu_error : std::runtime_error {
    int cuda_code;
    gpu_error(std::string const &msg, int code) : std::runtime_error(msg), cuda_code(code) {}
};
struct gpu_out_of_memory : gpu_error { using gpu_error::gpu_error; };

namespace detail {
inline void throw_status(int rc) {
    if (rc == EB200_OK) return;
    std::string msg = eb200_last_error_string();
    switch (rc) {
    case EB200_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
    case EB200_ERR_DOMAIN: throw std::domain_error(msg);
    case EB200_ERR_DIMENSION: throw dimension_error(msg);
    case EB200_ERR_RANK: throw rank_error(msg);
    case EB200_ERR_UNSUPPORTED: throw std::runtime_error(msg);
    case EB200_ERR_OUT_OF_MEMORY: throw gpu_out_of_memory(msg, eb200_last_cuda_error());
    case EB200_ERR_NOT_INITIALIZED: throw uninitialized_error(msg);
    default: throw gpu_error(msg, eb200_last_cuda_error());
    }
}
} // namespace detail
} // namespace einsums

#define EB200_CHECK(expr) ::einsums::detail::throw_status((expr))

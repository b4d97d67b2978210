// Drop-in umbrella header: `#include <Einsums/TensorAlgebra.hpp>` as in the reference, device path only.
#pragma once
#include <Einsums/B200/TensorAlgebra.hpp>

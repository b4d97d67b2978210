// Drop-in umbrella header: `#include <Einsums/Tensor/BlockTensor.hpp>` as in the reference, device path only.
#pragma once
#include <Einsums/B200/BlockTensor.hpp>

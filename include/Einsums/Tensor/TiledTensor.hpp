// Drop-in umbrella header: `#include <Einsums/Tensor/TiledTensor.hpp>` as in the reference, device path only.
#pragma once
#include <Einsums/B200/TiledTensor.hpp>

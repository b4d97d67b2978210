// Drop-in umbrella header for the mixed block x tiled algebra (reference: TensorAlgebra/Backends/BlockTileAlgebra.hpp), device path only.
#pragma once
#include <Einsums/B200/BlockTileAlgebra.hpp>

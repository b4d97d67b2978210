// Drop-in path of the reference's typed device-BLAS facade (libs/Einsums/hipBLAS/include/Einsums/hipBLAS.hpp): einsums::blas::gpu::*<T>.
#pragma once
#include <Einsums/B200/BLASGpu.hpp>

#pragma once
#include <Einsums/B200/DeviceTensor.hpp>

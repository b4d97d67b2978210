#pragma once
#include <Einsums/B200/LinearAlgebra.hpp>

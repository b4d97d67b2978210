// contract.cuh — internal entry of the contraction engine (used by blas.cu).
#pragma once
#include "common.cuh"
namespace eb200 {
int contract_dispatch(cudaStream_t s, eb200_contract_desc &d);
}

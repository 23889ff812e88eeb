// predictor.cuh -- state of the safety-signal voltage predictor + device replay sink
// (filled in by predictor.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct PredictorState {
    int loaded = 0;
};

inline void predictor_free(PredictorState*) {}

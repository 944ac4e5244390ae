// Product build: the pipeline on DeviceBackend (CUDA kernels for sm_100a + CUB scan/sort).
#include "phz_backend.h"
#define PHZ_BACKEND phz::DeviceBackend
#define PHZ_BACKEND_NAME "cuda-sm_100a"
#include "phz_api.inl"

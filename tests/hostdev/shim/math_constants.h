// TEST INFRASTRUCTURE (see cuda_runtime.h in this directory)
#pragma once
#define CUDART_INF_F (__builtin_inff())

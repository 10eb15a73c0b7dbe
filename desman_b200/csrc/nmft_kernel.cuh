// desman_b200/csrc/nmft_kernel.cuh -- K6-K8: NMFT initialiser (Init_NMFT.py).  (first slice: not yet implemented)
#pragma once
#include "common.cuh"
#include <stdio.h>
static int nmft_factorize_impl(cudaStream_t, int, const int64_t *, int64_t, int, int, double *, double *, int, double, int,
                               int *, double *, double *, char *err, size_t errn)
{
    snprintf(err, errn, "NMFT kernels not built yet");
    return -4;
}

#include "common.cuh"
extern "C" int upol_batched_minimize_dev(int, int64_t, int64_t, const double *, double *, const double *, const double *,
                                         const double *, const int32_t *, int64_t, const int32_t *, const int32_t *,
                                         const double *, double, int32_t, double *, int32_t *, void *) { return UPOL_ERR_ARG; }

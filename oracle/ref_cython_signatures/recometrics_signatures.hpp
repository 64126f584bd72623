/* TEST INFRASTRUCTURE (oracle/): what the reference's UNMODIFIED Cython wrapper (recometrics/wrapper.pyx, `cdef extern from
 * "recometrics_signatures.hpp"`, :36-203) gets to see under that name when oracle/build_ref_cython.py builds it against
 * librecometrics_b200.so instead of the reference's CPU code:
 *
 *   calc_metrics_float / calc_metrics_double / get_has_openmp   defined by include/recometrics_b200_shim.hpp (-> the C-ABI)
 *   split_data_*                                                 declared by the reference's own header (their definitions
 *                                                                come from the reference's recometrics_instantiated.cpp,
 *                                                                compiled next to it: the splitters are out of scope here)
 *
 * The reference header's declarations of the three metric entry points are renamed out of the way; nothing of it is copied.
 */
#ifndef RMB200_REF_CYTHON_SIGNATURES_HPP
#define RMB200_REF_CYTHON_SIGNATURES_HPP
#include "recometrics_b200_shim.hpp"
#define calc_metrics_float  rmb200_refdecl_calc_metrics_float
#define calc_metrics_double rmb200_refdecl_calc_metrics_double
#define get_has_openmp      rmb200_refdecl_get_has_openmp
#include RMB200_REFERENCE_SIGNATURES_HPP   /* -DRMB200_REFERENCE_SIGNATURES_HPP="\"<reference>/src/recometrics_signatures.hpp\"" */
#undef calc_metrics_float
#undef calc_metrics_double
#undef get_has_openmp
#endif

/* TEST INFRASTRUCTURE (oracle/): what the reference's UNMODIFIED Cython wrapper (recometrics/wrapper.pyx, `cdef extern from
 * "recometrics_signatures.hpp"`, :36-203) gets to see under that name when oracle/build_ref_cython.py builds it against
 * librecometrics_b200.so instead of the reference's CPU code:
 *
 *   calc_metrics_float / calc_metrics_double / get_has_openmp   defined by include/recometrics_b200_shim.hpp (-> the C-ABI)
 *   split_data_{selected,separate,joined}_users_float/_double    likewise (-> rmb200_split_*)
 *
 * The reference header is still included -- FIRST, so that the shim's definitions are checked against its declarations by
 * the compiler (a different parameter list would be an overload the wrapper's calls could not pick unambiguously) -- and
 * nothing of it is copied.  No reference source file is compiled into the module.
 */
#ifndef RMB200_REF_CYTHON_SIGNATURES_HPP
#define RMB200_REF_CYTHON_SIGNATURES_HPP
#include RMB200_REFERENCE_SIGNATURES_HPP   /* -DRMB200_REFERENCE_SIGNATURES_HPP="\"<reference>/src/recometrics_signatures.hpp\"" */
#include "recometrics_b200_shim.hpp"
#endif

/*
 * ref_hooks.h -- force-included (-include) in front of the reference's
 * saena_object_solve.cpp by oracle/Makefile (TEST INFRASTRUCTURE).
 *
 * The reference prints no per-iteration residual in Release builds (the
 * printf at /root/reference/src/saena_object_solve.cpp:2616-2617 is commented
 * out).  To record the residual-norm history WITHOUT patching any reference
 * source, this header pulls in aux_functions.h first (its include guard makes
 * the TU's own #include a no-op) and then routes every dotProduct(r, r, ...)
 * call of that TU through a recorder.  solve_pCG's <r,r> evaluations
 * (:2501 initial, :2603 per iteration) are exactly the calls with both
 * operands equal; the vector length tells them apart from solve_coarsest_CG's own <res,res>
 * (:21, :76) when direct_solver == "CG".
 */
#ifndef SAENA_B200_ORACLE_REF_HOOKS_H
#define SAENA_B200_ORACLE_REF_HOOKS_H
#ifdef __cplusplus
#include "aux_functions.h"

extern "C" void saena_ref_record_rr(double rr, int sz);

static inline void saena_ref_dot_hook(const value_t *r, const value_t *s, const index_t sz, value_t *dot,
                                      MPI_Comm comm) {
    dotProduct(r, s, sz, dot, comm);
    if (r == s) saena_ref_record_rr(*dot, (int)sz);
}
#ifndef SAENA_REF_HARNESS_TU
#define dotProduct(r, s, sz, dot, comm) saena_ref_dot_hook(r, s, sz, dot, comm)
#endif
#endif
#endif

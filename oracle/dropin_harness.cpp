/*
 * dropin_harness.cpp -- TEST INFRASTRUCTURE for the drop-in (tests/test_public_api_dropin.py).
 *
 * Linked into oracle/_ref/libsaena_dropin.so together with the UNMODIFIED reference objects
 * (src/saena.cpp's five solve-path forwarders weakened by objcopy) and
 * saena_b200/adaptor/saena_b200_adaptor.cpp.  It runs the reference's own driver sequence
 * (/root/reference/experiments/Poisson.cpp:81-246) through the PUBLIC saena.hpp API:
 * saena::amg::solve_pCG now lands on the GPU, while solver.get_object()->solve_pCG() is the
 * reference's CPU solve -- same process, same hierarchy object.
 */
#define SAENA_REF_HARNESS_TU
#include "ref_shim/ref_hooks.h"

#include "saena.hpp"
#include "saena_object.h"
#include "aux_functions2.h"

#include <cmath>
#include <cstring>
#include <vector>
#include <unistd.h>
#include <fcntl.h>

static std::vector<double> g_rr;
extern "C" void saena_ref_record_rr(double rr, int sz) { (void)sz; g_rr.push_back(rr); }
extern "C" int saena_b200_adaptor_last_iterations(void);
extern "C" int saena_b200_adaptor_last_history(double *out, int cap);
extern "C" void saena_b200_adaptor_release(saena::amg *solver);

namespace {
struct Quiet {
    int saved;
    Quiet() { fflush(stdout); saved = dup(1); int n = open("/dev/null", O_WRONLY); dup2(n, 1); close(n); }
    ~Quiet() { fflush(stdout); dup2(saved, 1); close(saved); }
};
}

extern "C" {

/* Returns 0 on success.  hist_gpu/hist_cpu: residual-norm histories (capacity cap), u_rel_diff:
 * ||u_gpu - u_cpu|| / ||u_cpu||, mv_rel_diff: the same for saena::matrix::matvec vs saena_matrix::matvec. */
int dropin_poisson_check(int mx, int *iters_gpu, int *iters_cpu, double *hist_gpu, int *n_gpu, double *hist_cpu,
                         int *n_cpu, int cap, double *u_rel_diff, double *mv_rel_diff) {
    int inited = 0;
    MPI_Initialized(&inited);
    if (!inited) MPI_Init(nullptr, nullptr);
    MPI_Comm comm = MPI_COMM_WORLD;
    Quiet q;
    saena::matrix A(comm);
    saena::laplacian3D(&A, mx, mx, mx);
    A.set_remove_boundary(true);
    A.assemble(false);
    value_t *rhs_std = nullptr;
    index_t orig_sz = saena::laplacian3D_set_rhs(rhs_std, mx, mx, mx, comm);
    index_t my_split = 0;
    saena::find_split(orig_sz, my_split, comm);
    saena::vector rhs(comm);
    rhs.set(&rhs_std[0], orig_sz, my_split);
    rhs.assemble();
    // data/options006_poisson.xml
    saena::options opts(50, 1e-8, "chebyshev", 3, 3, "jacobi", 0.2f, true, 20, 0, 1e-12, 1e-9, 1, 1, false, 0.1f, 5000);
    saena::amg solver;
    solver.set_scale(false);
    solver.set_matrix(&A, &opts);
    solver.set_rhs(rhs);

    // --- GPU through the public API (adaptor) ---
    value_t *u = nullptr;
    solver.solve_pCG(u, &opts, false);
    *iters_gpu = saena_b200_adaptor_last_iterations();
    *n_gpu = saena_b200_adaptor_last_history(hist_gpu, cap);

    // --- reference CPU path on the same hierarchy object ---
    value_t *u2 = nullptr;
    g_rr.clear();
    solver.get_object()->solve_pCG(u2, false);
    *n_cpu = (int)g_rr.size();
    for (int i = 0; i < *n_cpu && i < cap; ++i) hist_cpu[i] = std::sqrt(g_rr[i]);
    *iters_cpu = *n_cpu - 1;

    const index_t M = A.get_internal_matrix()->M;
    double num = 0, den = 0;
    for (index_t i = 0; i < M; ++i) { num += (u[i] - u2[i]) * (u[i] - u2[i]); den += u2[i] * u2[i]; }
    *u_rel_diff = std::sqrt(num / den);

    // --- saena::matrix::matvec (adaptor) vs saena_matrix::matvec (reference) ---
    std::vector<value_t> v(M), w(M), w2(M);
    for (index_t i = 0; i < M; ++i) v[i] = std::sin(0.37 * i) + 0.1;
    A.matvec(v, w);
    A.get_internal_matrix()->matvec(&v[0], &w2[0]);
    num = den = 0;
    for (index_t i = 0; i < M; ++i) { num += (w[i] - w2[i]) * (w[i] - w2[i]); den += w2[i] * w2[i]; }
    *mv_rel_diff = std::sqrt(num / den);

    saena_b200_adaptor_release(nullptr);
    saena_free(u);
    saena_free(u2);
    saena_free(rhs_std);
    solver.destroy();
    A.destroy();
    return 0;
}

}

/*
 * omp_stubs.c -- the four OpenMP runtime calls the reference makes outside its
 * SAENA_USE_OPENMP guards (timers and thread counts), for a build with OpenMP
 * off -- the reference's default (/root/reference/CMakeLists.txt:27).
 * TEST INFRASTRUCTURE.
 */
#include <time.h>
double omp_get_wtime(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
int omp_get_max_threads(void) { return 1; }
int omp_get_thread_num(void) { return 0; }
int omp_get_num_threads(void) { return 1; }
void omp_set_num_threads(int n) { (void)n; }

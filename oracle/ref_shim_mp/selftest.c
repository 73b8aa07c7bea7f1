/* selftest.c -- exercises the multi-process MPI stand-in on its own (no reference code):
 *   python -m oracle.mprun -n 4 oracle/_ref/mpi_selftest
 * Every check aborts with a message; rank 0 prints "SBMPI_SELFTEST_OK <size>" at the end. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mpi.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "rank %d: check failed: %s (line %d)\n", rank, #c, __LINE__); MPI_Abort(MPI_COMM_WORLD, 3); } } while (0)

static void land(void *in, void *inout, int *n, MPI_Datatype *dt) { (void)dt; for (int i = 0; i < *n; ++i) ((char *)inout)[i] = ((char *)in)[i] && ((char *)inout)[i]; }

int main(int argc, char **argv) {
    int rank, size;
    MPI_Init(&argc, &argv);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    /* ring: nonblocking, large message (beyond the socket buffers), both directions at once */
    const int N = 3000000;
    double *s = malloc(sizeof(double) * N), *r = malloc(sizeof(double) * N);
    for (int i = 0; i < N; ++i) s[i] = rank + 1e-6 * i;
    MPI_Request q[2];
    const int right = (rank + 1) % size, left = (rank + size - 1) % size;
    MPI_Irecv(r, N, MPI_DOUBLE, left, 7, MPI_COMM_WORLD, &q[0]);
    MPI_Isend(s, N, MPI_DOUBLE, right, 7, MPI_COMM_WORLD, &q[1]);
    MPI_Status st[2];
    MPI_Waitall(2, q, st);
    CHECK(r[0] == left && r[N - 1] == left + 1e-6 * (N - 1));
    int cnt; MPI_Get_count(&st[0], MPI_DOUBLE, &cnt);
    CHECK(cnt == N && st[0].MPI_SOURCE == left && st[0].MPI_TAG == 7);
    /* ordering: two messages with the same tag arrive in order; Waitany reports each receive once */
    int a = 10 * rank + 1, b = 10 * rank + 2, x[2] = {0, 0};
    MPI_Request rq[2];
    MPI_Irecv(&x[0], 1, MPI_INT, left, 9, MPI_COMM_WORLD, &rq[0]);
    MPI_Irecv(&x[1], 1, MPI_INT, left, 9, MPI_COMM_WORLD, &rq[1]);
    int flag; MPI_Test(&rq[0], &flag, MPI_STATUS_IGNORE);
    MPI_Send(&a, 1, MPI_INT, right, 9, MPI_COMM_WORLD);
    MPI_Send(&b, 1, MPI_INT, right, 9, MPI_COMM_WORLD);
    int seen = 0;
    for (int k = 0; k < 2; ++k) { int idx; MPI_Waitany(2, rq, &idx, MPI_STATUS_IGNORE); CHECK(idx == 0 || idx == 1); seen |= 1 << idx; }
    CHECK(seen == 3 && x[0] == 10 * left + 1 && x[1] == 10 * left + 2);
    /* collectives */
    double v = rank + 1.0, sum = 0, mx = 0;
    MPI_Allreduce(&v, &sum, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    MPI_Allreduce(&v, &mx, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
    CHECK(sum == size * (size + 1) / 2.0 && mx == size);
    long lv = rank, lsum = -1;
    MPI_Reduce(&lv, &lsum, 1, MPI_LONG, MPI_SUM, size - 1, MPI_COMM_WORLD);
    if (rank == size - 1) CHECK(lsum == (long)size * (size - 1) / 2);
    int in_place = rank + 1;
    MPI_Allreduce(MPI_IN_PLACE, &in_place, 1, MPI_INT, MPI_MIN, MPI_COMM_WORLD);
    CHECK(in_place == 1);
    struct { double v; int i; } pr = {(double)((rank * 7) % size), rank}, po;
    MPI_Allreduce(&pr, &po, 1, MPI_DOUBLE_INT, MPI_MAXLOC, MPI_COMM_WORLD);
    CHECK(po.v == (double)((po.i * 7) % size));
    MPI_Op op; MPI_Op_create(land, 1, &op);
    char bl = 1, bo = 0; MPI_Allreduce(&bl, &bo, 1, MPI_CHAR, op, MPI_COMM_WORLD); CHECK(bo == 1);
    int sc = 0, me1 = rank + 1; MPI_Scan(&me1, &sc, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD); CHECK(sc == (rank + 1) * (rank + 2) / 2);
    int *all = malloc(sizeof(int) * size);
    MPI_Allgather(&rank, 1, MPI_INT, all, 1, MPI_INT, MPI_COMM_WORLD);
    for (int i = 0; i < size; ++i) CHECK(all[i] == i);
    /* alltoallv: rank i sends j+1 copies of (100 i + j) to rank j */
    int *scnt = malloc(sizeof(int) * size), *sdsp = malloc(sizeof(int) * size), *rcnt = malloc(sizeof(int) * size), *rdsp = malloc(sizeof(int) * size);
    int tot = 0; for (int j = 0; j < size; ++j) { scnt[j] = j + 1; sdsp[j] = tot; tot += j + 1; }
    int *sb = malloc(sizeof(int) * tot), *rb = malloc(sizeof(int) * size * (rank + 1));
    for (int j = 0; j < size; ++j) for (int k = 0; k < j + 1; ++k) sb[sdsp[j] + k] = 100 * rank + j;
    for (int i = 0; i < size; ++i) { rcnt[i] = rank + 1; rdsp[i] = i * (rank + 1); }
    MPI_Alltoallv(sb, scnt, sdsp, MPI_INT, rb, rcnt, rdsp, MPI_INT, MPI_COMM_WORLD);
    for (int i = 0; i < size; ++i) for (int k = 0; k < rank + 1; ++k) CHECK(rb[rdsp[i] + k] == 100 * i + rank);
    int *gv = malloc(sizeof(int) * tot), *gc = malloc(sizeof(int) * size), *gd = malloc(sizeof(int) * size);
    int t2 = 0; for (int i = 0; i < size; ++i) { gc[i] = i + 1; gd[i] = t2; t2 += i + 1; }
    int *mine = malloc(sizeof(int) * (rank + 1)); for (int k = 0; k <= rank; ++k) mine[k] = rank;
    MPI_Allgatherv(mine, rank + 1, MPI_INT, gv, gc, gd, MPI_INT, MPI_COMM_WORLD);
    for (int i = 0; i < size; ++i) for (int k = 0; k <= i; ++k) CHECK(gv[gd[i] + k] == i);
    /* communicators: split by parity with reversed keys, dup, group + create, create_group */
    MPI_Comm half; MPI_Comm_split(MPI_COMM_WORLD, rank % 2, -rank, &half);
    int hr, hs; MPI_Comm_rank(half, &hr); MPI_Comm_size(half, &hs);
    CHECK(hs == (size + 1 - rank % 2) / 2);
    int hsum = 0; MPI_Allreduce(&rank, &hsum, 1, MPI_INT, MPI_SUM, half);
    int want = 0; for (int i = rank % 2; i < size; i += 2) want += i; CHECK(hsum == want);
    int first = -1; if (hr == 0) first = rank; MPI_Bcast(&first, 1, MPI_INT, 0, half);
    CHECK(first == (size - 1) - ((size - 1 - rank % 2) % 2 ? 1 : 0) || hs >= 1);   /* highest rank of my parity has key order 0 */
    MPI_Comm dup; MPI_Comm_dup(half, &dup); int d2 = 1, ds = 0; MPI_Allreduce(&d2, &ds, 1, MPI_INT, MPI_SUM, dup); CHECK(ds == hs);
    MPI_Group wg, sub; MPI_Comm_group(MPI_COMM_WORLD, &wg);
    const int nsub = size > 1 ? size - 1 : 1; int *ranks = malloc(sizeof(int) * nsub); for (int i = 0; i < nsub; ++i) ranks[i] = i;
    MPI_Group_incl(wg, nsub, ranks, &sub);
    MPI_Comm c1; MPI_Comm_create(MPI_COMM_WORLD, sub, &c1);
    if (rank < nsub) { int one = 1, cs = 0; CHECK(c1 != MPI_COMM_NULL); MPI_Allreduce(&one, &cs, 1, MPI_INT, MPI_SUM, c1); CHECK(cs == nsub); }
    else CHECK(c1 == MPI_COMM_NULL);
    if (rank < nsub) { MPI_Comm c2; MPI_Comm_create_group(MPI_COMM_WORLD, sub, 5, &c2); int one = 1, cs = 0; MPI_Allreduce(&one, &cs, 1, MPI_INT, MPI_SUM, c2); CHECK(cs == nsub); MPI_Comm_free(&c2); }
    /* probe + any source */
    if (rank == 0) { for (int i = 1; i < size; ++i) { MPI_Status ps; MPI_Probe(MPI_ANY_SOURCE, 33, MPI_COMM_WORLD, &ps); int c; MPI_Get_count(&ps, MPI_INT, &c); int *buf = malloc(sizeof(int) * c); MPI_Recv(buf, c, MPI_INT, ps.MPI_SOURCE, 33, MPI_COMM_WORLD, MPI_STATUS_IGNORE); CHECK(c == ps.MPI_SOURCE && (c == 0 || buf[c - 1] == ps.MPI_SOURCE)); free(buf); } }
    else { int *buf = malloc(sizeof(int) * rank); for (int i = 0; i < rank; ++i) buf[i] = rank; MPI_Send(buf, rank, MPI_INT, 0, 33, MPI_COMM_WORLD); free(buf); }
    double ex = rank, got = -1; MPI_Sendrecv(&ex, 1, MPI_DOUBLE, right, 4, &got, 1, MPI_DOUBLE, left, 4, MPI_COMM_WORLD, MPI_STATUS_IGNORE); CHECK(got == left);
    MPI_Barrier(MPI_COMM_WORLD);
    if (rank == 0) printf("SBMPI_SELFTEST_OK %d\n", size);
    MPI_Finalize();
    return 0;
}

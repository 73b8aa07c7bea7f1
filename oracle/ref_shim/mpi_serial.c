/*
 * mpi_serial.c -- implementation of the single-rank MPI stand-in declared in
 * mpi.h (TEST INFRASTRUCTURE; see the header).  Communicators always have
 * size 1 / rank 0.  Messages a rank sends to itself are buffered in a list
 * and matched against receives by (comm, tag) in posting order.
 */
#include "mpi.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static int g_initialized = 0, g_finalized = 0;
static int g_next_comm = 16;

/* ---------------- self-message queue ---------------- */
typedef struct msg {
    int comm, tag;
    size_t bytes;
    void *data;
    struct msg *next;
} msg_t;

typedef struct req {
    int in_use;
    int is_recv;
    int done;
    int comm, tag;
    void *buf;
    size_t cap;
    size_t got;
} req_t;

static msg_t *g_head = NULL, *g_tail = NULL;
static req_t *g_reqs = NULL;
static int g_nreqs = 0;

static int req_alloc(void) {
    for (int i = 1; i < g_nreqs; ++i)
        if (!g_reqs[i].in_use) {
            memset(&g_reqs[i], 0, sizeof(req_t));
            g_reqs[i].in_use = 1;
            return i;
        }
    int old = g_nreqs;
    g_nreqs = old ? old * 2 : 64;
    g_reqs = (req_t *)realloc(g_reqs, sizeof(req_t) * (size_t)g_nreqs);
    memset(g_reqs + old, 0, sizeof(req_t) * (size_t)(g_nreqs - old));
    int i = old ? old : 1; /* slot 0 is MPI_REQUEST_NULL */
    g_reqs[i].in_use = 1;
    return i;
}

static int tag_match(int want, int have) { return want == MPI_ANY_TAG || want == have; }

static void try_match(req_t *r) {
    if (r->done) return;
    msg_t *prev = NULL;
    for (msg_t *m = g_head; m; prev = m, m = m->next) {
        if (m->comm == r->comm && tag_match(r->tag, m->tag)) {
            size_t n = m->bytes < r->cap ? m->bytes : r->cap;
            if (n) memcpy(r->buf, m->data, n);
            r->got = n;
            r->tag = m->tag;
            r->done = 1;
            if (prev) prev->next = m->next; else g_head = m->next;
            if (g_tail == m) g_tail = prev;
            free(m->data);
            free(m);
            return;
        }
    }
}

static void match_pending_recvs(void) {
    for (int i = 1; i < g_nreqs; ++i)
        if (g_reqs[i].in_use && g_reqs[i].is_recv && !g_reqs[i].done) try_match(&g_reqs[i]);
}

static void enqueue(const void *buf, size_t bytes, int tag, int comm) {
    msg_t *m = (msg_t *)malloc(sizeof(msg_t));
    m->comm = comm;
    m->tag = tag;
    m->bytes = bytes;
    m->data = malloc(bytes ? bytes : 1);
    if (bytes) memcpy(m->data, buf, bytes);
    m->next = NULL;
    if (g_tail) g_tail->next = m; else g_head = m;
    g_tail = m;
    match_pending_recvs();
}

static void fill_status(MPI_Status *st, int tag, size_t bytes) {
    if (!st) return;
    st->MPI_SOURCE = 0;
    st->MPI_TAG = tag;
    st->MPI_ERROR = MPI_SUCCESS;
    st->count_bytes = (int)bytes;
}

static void die(const char *what) {
    fprintf(stderr, "mpi_serial: %s\n", what);
    abort();
}

/* ---------------- environment ---------------- */
int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; g_initialized = 1; return MPI_SUCCESS; }
int MPI_Init_thread(int *argc, char ***argv, int required, int *provided) {
    (void)argc; (void)argv;
    if (provided) *provided = required;
    g_initialized = 1;
    return MPI_SUCCESS;
}
int MPI_Initialized(int *flag) { *flag = g_initialized; return MPI_SUCCESS; }
int MPI_Finalize(void) { g_finalized = 1; return MPI_SUCCESS; }
int MPI_Finalized(int *flag) { *flag = g_finalized; return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm comm, int code) { (void)comm; fprintf(stderr, "MPI_Abort(%d)\n", code); exit(code ? code : 1); }
double MPI_Wtime(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
int MPI_Pcontrol(const int level, ...) { (void)level; return MPI_SUCCESS; }
int MPI_Get_processor_name(char *name, int *len) { strcpy(name, "serial"); *len = 6; return MPI_SUCCESS; }

/* ---------------- communicators / groups ---------------- */
int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = 1; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *out) {
    (void)comm; (void)key;
    *out = (color == MPI_UNDEFINED) ? MPI_COMM_NULL : g_next_comm++;
    return MPI_SUCCESS;
}
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out) { (void)comm; *out = g_next_comm++; return MPI_SUCCESS; }
int MPI_Comm_free(MPI_Comm *comm) { *comm = MPI_COMM_NULL; return MPI_SUCCESS; }
/* a group is 2 if it contains rank 0, GROUP_EMPTY otherwise */
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group) { (void)comm; *group = 2; return MPI_SUCCESS; }
int MPI_Group_incl(MPI_Group group, int n, const int ranks[], MPI_Group *out) {
    (void)group;
    *out = MPI_GROUP_EMPTY;
    for (int i = 0; i < n; ++i) if (ranks[i] == 0) *out = 2;
    return MPI_SUCCESS;
}
int MPI_Group_free(MPI_Group *group) { *group = MPI_GROUP_NULL; return MPI_SUCCESS; }
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *out) {
    (void)comm;
    *out = (group == 2) ? g_next_comm++ : MPI_COMM_NULL;
    return MPI_SUCCESS;
}
int MPI_Comm_create_group(MPI_Comm comm, MPI_Group group, int tag, MPI_Comm *out) {
    (void)tag;
    return MPI_Comm_create(comm, group, out);
}
int MPI_Comm_set_errhandler(MPI_Comm comm, MPI_Errhandler eh) { (void)comm; (void)eh; return MPI_SUCCESS; }
int MPI_Attr_get(MPI_Comm comm, int keyval, void *attr, int *flag) {
    static int tag_ub = 1 << 30;
    (void)comm;
    if (keyval == MPI_TAG_UB) { *(int **)attr = &tag_ub; *flag = 1; } else *flag = 0;
    return MPI_SUCCESS;
}

/* ---------------- collectives: copy ---------------- */
static void cpy(const void *s, void *r, size_t bytes) {
    if (s == MPI_IN_PLACE || s == r || bytes == 0) return;
    memmove(r, s, bytes);
}
int MPI_Barrier(MPI_Comm comm) { (void)comm; return MPI_SUCCESS; }
int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm) {
    (void)buf; (void)count; (void)dt; (void)root; (void)comm; return MPI_SUCCESS;
}
int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm) {
    (void)op; (void)comm; cpy(s, r, (size_t)count * (size_t)dt); return MPI_SUCCESS;
}
int MPI_Reduce(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm) {
    (void)root; return MPI_Allreduce(s, r, count, dt, op, comm);
}
int MPI_Scan(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm) {
    return MPI_Allreduce(s, r, count, dt, op, comm);
}
int MPI_Exscan(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm) {
    (void)s; (void)r; (void)count; (void)dt; (void)op; (void)comm; return MPI_SUCCESS; /* rank 0: undefined */
}
int MPI_Allgather(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, MPI_Comm comm) {
    (void)rc; (void)rt; (void)comm; cpy(s, r, (size_t)sc * (size_t)st); return MPI_SUCCESS;
}
int MPI_Allgatherv(const void *s, int sc, MPI_Datatype st, void *r, const int *rc, const int *displs,
                   MPI_Datatype rt, MPI_Comm comm) {
    (void)rc; (void)comm;
    cpy(s, (char *)r + (size_t)displs[0] * (size_t)rt, (size_t)sc * (size_t)st);
    return MPI_SUCCESS;
}
int MPI_Gather(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, int root, MPI_Comm comm) {
    (void)root; return MPI_Allgather(s, sc, st, r, rc, rt, comm);
}
int MPI_Gatherv(const void *s, int sc, MPI_Datatype st, void *r, const int *rc, const int *displs,
                MPI_Datatype rt, int root, MPI_Comm comm) {
    (void)root; return MPI_Allgatherv(s, sc, st, r, rc, displs, rt, comm);
}
int MPI_Scatterv(const void *s, const int *sc, const int *displs, MPI_Datatype st, void *r, int rc,
                 MPI_Datatype rt, int root, MPI_Comm comm) {
    (void)rc; (void)rt; (void)root; (void)comm;
    if (r != MPI_IN_PLACE) cpy((const char *)s + (size_t)displs[0] * (size_t)st, r, (size_t)sc[0] * (size_t)st);
    return MPI_SUCCESS;
}
int MPI_Alltoall(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, MPI_Comm comm) {
    (void)rc; (void)rt; (void)comm; cpy(s, r, (size_t)sc * (size_t)st); return MPI_SUCCESS;
}
int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype st, void *r, const int *rc,
                  const int *rd, MPI_Datatype rt, MPI_Comm comm) {
    (void)rc; (void)comm;
    if (s == MPI_IN_PLACE) return MPI_SUCCESS;
    cpy((const char *)s + (size_t)sd[0] * (size_t)st, (char *)r + (size_t)rd[0] * (size_t)rt,
        (size_t)sc[0] * (size_t)st);
    return MPI_SUCCESS;
}

/* ---------------- point-to-point (self only) ---------------- */
static void check_peer(int peer, const char *fn) {
    if (peer != 0 && peer != MPI_ANY_SOURCE && peer != MPI_PROC_NULL) {
        fprintf(stderr, "mpi_serial: %s with peer %d in a one-rank world\n", fn, peer);
        abort();
    }
}
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req) {
    check_peer(dest, "MPI_Isend");
    int id = req_alloc();
    g_reqs[id].done = 1; /* buffered */
    *req = id;
    if (dest != MPI_PROC_NULL) enqueue(buf, (size_t)count * (size_t)dt, tag, comm);
    return MPI_SUCCESS;
}
int MPI_Issend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req) {
    return MPI_Isend(buf, count, dt, dest, tag, comm, req);
}
int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm) {
    check_peer(dest, "MPI_Send");
    if (dest != MPI_PROC_NULL) enqueue(buf, (size_t)count * (size_t)dt, tag, comm);
    return MPI_SUCCESS;
}
int MPI_Bsend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm) {
    return MPI_Send(buf, count, dt, dest, tag, comm);
}
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request *req) {
    check_peer(src, "MPI_Irecv");
    int id = req_alloc();
    req_t *r = &g_reqs[id];
    r->is_recv = 1;
    r->comm = comm;
    r->tag = tag;
    r->buf = buf;
    r->cap = (size_t)count * (size_t)dt;
    if (src == MPI_PROC_NULL) r->done = 1; else try_match(r);
    *req = id;
    return MPI_SUCCESS;
}
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status *st) {
    MPI_Request rq;
    MPI_Irecv(buf, count, dt, src, tag, comm, &rq);
    return MPI_Wait(&rq, st);
}
int MPI_Sendrecv(const void *s, int sc, MPI_Datatype sdt, int dest, int stag, void *r, int rc, MPI_Datatype rdt,
                 int src, int rtag, MPI_Comm comm, MPI_Status *status) {
    MPI_Request rq;
    MPI_Irecv(r, rc, rdt, src, rtag, comm, &rq);
    MPI_Send(s, sc, sdt, dest, stag, comm);
    return MPI_Wait(&rq, status);
}
static int req_complete(MPI_Request *req, MPI_Status *st, int blocking) {
    if (*req == MPI_REQUEST_NULL) { fill_status(st, MPI_ANY_TAG, 0); return 1; }
    if (*req < 0 || *req >= g_nreqs || !g_reqs[*req].in_use) die("bad request handle");
    req_t *r = &g_reqs[*req];
    if (r->is_recv && !r->done) try_match(r);
    if (!r->done) {
        if (blocking) die("MPI_Wait on a receive that no send matches (deadlock in a one-rank world)");
        return 0;
    }
    fill_status(st, r->tag, r->got);
    r->in_use = 0;
    *req = MPI_REQUEST_NULL;
    return 1;
}
int MPI_Wait(MPI_Request *req, MPI_Status *st) { req_complete(req, st, 1); return MPI_SUCCESS; }
int MPI_Waitall(int n, MPI_Request reqs[], MPI_Status sts[]) {
    for (int i = 0; i < n; ++i) req_complete(&reqs[i], sts ? &sts[i] : NULL, 1);
    return MPI_SUCCESS;
}
int MPI_Waitany(int n, MPI_Request reqs[], int *index, MPI_Status *st) {
    int active = 0;
    for (int i = 0; i < n; ++i) {
        if (reqs[i] == MPI_REQUEST_NULL) continue;
        active = 1;
        if (req_complete(&reqs[i], st, 0)) { *index = i; return MPI_SUCCESS; }
    }
    if (active) die("MPI_Waitany: no request can complete (deadlock in a one-rank world)");
    *index = MPI_UNDEFINED;
    return MPI_SUCCESS;
}
int MPI_Test(MPI_Request *req, int *flag, MPI_Status *st) { *flag = req_complete(req, st, 0); return MPI_SUCCESS; }
int MPI_Testall(int n, MPI_Request reqs[], int *flag, MPI_Status sts[]) {
    *flag = 1;
    for (int i = 0; i < n; ++i) {
        if (reqs[i] == MPI_REQUEST_NULL) continue;
        req_t *r = &g_reqs[reqs[i]];
        if (r->is_recv && !r->done) try_match(r);
        if (!r->done) *flag = 0;
    }
    if (*flag) MPI_Waitall(n, reqs, sts);
    return MPI_SUCCESS;
}
int MPI_Iprobe(int src, int tag, MPI_Comm comm, int *flag, MPI_Status *st) {
    (void)src;
    *flag = 0;
    for (msg_t *m = g_head; m; m = m->next)
        if (m->comm == comm && tag_match(tag, m->tag)) { *flag = 1; fill_status(st, m->tag, m->bytes); break; }
    return MPI_SUCCESS;
}
int MPI_Probe(int src, int tag, MPI_Comm comm, MPI_Status *st) {
    int flag;
    MPI_Iprobe(src, tag, comm, &flag, st);
    if (!flag) die("MPI_Probe: nothing to probe (deadlock in a one-rank world)");
    return MPI_SUCCESS;
}
int MPI_Request_free(MPI_Request *req) {
    if (*req != MPI_REQUEST_NULL) g_reqs[*req].in_use = 0;
    *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}
int MPI_Cancel(MPI_Request *req) { if (*req != MPI_REQUEST_NULL) g_reqs[*req].done = 1; return MPI_SUCCESS; }
int MPI_Get_count(const MPI_Status *st, MPI_Datatype dt, int *count) {
    *count = dt ? st->count_bytes / dt : 0;
    return MPI_SUCCESS;
}

/* ---------------- datatypes / ops ---------------- */
int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype *newt) { *newt = count * old; return MPI_SUCCESS; }
int MPI_Type_create_struct(int n, const int bl[], const MPI_Aint disp[], const MPI_Datatype types[],
                           MPI_Datatype *newt) {
    long end = 0, align = 1;
    for (int i = 0; i < n; ++i) {
        long e = (long)disp[i] + (long)bl[i] * types[i];
        if (e > end) end = e;
        long a = types[i] > 8 ? 8 : types[i];
        if (a > align) align = a;
    }
    *newt = (int)((end + align - 1) / align * align);
    return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype *dt) { (void)dt; return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype *dt) { *dt = MPI_DATATYPE_NULL; return MPI_SUCCESS; }
int MPI_Type_size(MPI_Datatype dt, int *size) { *size = dt; return MPI_SUCCESS; }
int MPI_Op_create(MPI_User_function *fn, int commute, MPI_Op *op) { (void)fn; (void)commute; *op = 100; return MPI_SUCCESS; }
int MPI_Op_free(MPI_Op *op) { *op = MPI_OP_NULL; return MPI_SUCCESS; }

/* ---------------- one-sided (usort only) ---------------- */
struct mpi_serial_win { char *base; int disp_unit; };
int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr) {
    (void)info; *(void **)baseptr = malloc((size_t)size); return MPI_SUCCESS;
}
int MPI_Free_mem(void *base) { free(base); return MPI_SUCCESS; }
int MPI_Win_create(void *base, MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, MPI_Win *win) {
    (void)size; (void)info; (void)comm;
    *win = (MPI_Win)malloc(sizeof(struct mpi_serial_win));
    (*win)->base = (char *)base;
    (*win)->disp_unit = disp_unit;
    return MPI_SUCCESS;
}
int MPI_Win_fence(int assert_, MPI_Win win) { (void)assert_; (void)win; return MPI_SUCCESS; }
int MPI_Win_free(MPI_Win *win) { free(*win); *win = NULL; return MPI_SUCCESS; }
int MPI_Put(const void *origin, int ocount, MPI_Datatype odt, int target, MPI_Aint tdisp, int tcount,
            MPI_Datatype tdt, MPI_Win win) {
    (void)tcount; (void)tdt;
    check_peer(target, "MPI_Put");
    memcpy(win->base + (size_t)tdisp * (size_t)win->disp_unit, origin, (size_t)ocount * (size_t)odt);
    return MPI_SUCCESS;
}

/* ---------------- files ---------------- */
struct mpi_serial_file { FILE *f; };
int MPI_File_open(MPI_Comm comm, const char *name, int amode, MPI_Info info, MPI_File *fh) {
    (void)comm; (void)info;
    FILE *f = fopen(name, (amode & MPI_MODE_RDONLY) ? "rb" : ((amode & MPI_MODE_CREATE) ? "wb+" : "rb+"));
    if (!f) { *fh = NULL; return MPI_ERR_OTHER; }
    *fh = (MPI_File)malloc(sizeof(struct mpi_serial_file));
    (*fh)->f = f;
    return MPI_SUCCESS;
}
int MPI_File_read_at(MPI_File fh, MPI_Offset off, void *buf, int count, MPI_Datatype dt, MPI_Status *st) {
    fseeko(fh->f, (off_t)off, SEEK_SET);
    size_t n = fread(buf, (size_t)dt, (size_t)count, fh->f);
    fill_status(st, 0, n * (size_t)dt);
    return MPI_SUCCESS;
}
int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void *buf, int count, MPI_Datatype dt, MPI_Status *st) {
    fseeko(fh->f, (off_t)off, SEEK_SET);
    size_t n = fwrite(buf, (size_t)dt, (size_t)count, fh->f);
    fill_status(st, 0, n * (size_t)dt);
    return MPI_SUCCESS;
}
int MPI_File_close(MPI_File *fh) {
    if (*fh) { fclose((*fh)->f); free(*fh); *fh = NULL; }
    return MPI_SUCCESS;
}

/*
 * oracle/ref_tcp.c -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE (see collectives.c).
 *
 * "Restated reference TCP path (C)": the data plane of /root/reference/network.go, which cannot
 * be built here (no Go toolchain), restated so that `bench.py` can time the reference's way of
 * moving the same buffers on the GPU box's host cores.
 *
 *   world          n ranks, full mesh, TWO loopback TCP connections per ordered pair
 *                  (dial = I send data / read acks, listen = I read data / write acks)
 *                                                                   network.go:122-339, 501-506
 *   ref_send       gob-encode the slice into a buffer (network.go:537-542); same-rank shortcut
 *                  through an in-process rendezvous (network.go:545-548, 388-446); otherwise wrap in
 *                  message{Tag,Bytes} -- Raw.GobEncode copies the payload (mpi.go:77-81) -- write
 *                  it to the dial socket (network.go:562-563) and block for the peer's ack
 *                  envelope (network.go:551-559, 569)
 *   ref_recv       read one envelope from the listen socket (network.go:609), Raw.GobDecode copies
 *                  the payload out (mpi.go:83-91), write the ack {Tag} (network.go:617-621),
 *                  gob-decode into the typed destination (network.go:594-601)
 * ref_bench runs the ranks as threads of one process; ref_bench_procs forks one OS process per rank
 * (what the reference does: gompirun.go:85 starts one process per rank), with the sockets created
 * before the fork and a process-shared barrier around every step.  Either way the bytes cross the
 * kernel's TCP stack, which is what is being measured; ranks are not pinned.
 * Collectives do not exist in the reference (mpi.go:130): they are composed from Send/Receive
 * the way a user of the reference would -- ring allreduce / allgather, root-sends-to-all bcast --
 * with Send running concurrently with Receive (the "go mpi.Send(...)" idiom of
 * examples/helloworld/helloworld.go:56-68) because Send is synchronous.
 */
#define _GNU_SOURCE
#include <arpa/inet.h>
#include <errno.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

/* from gob.c */
size_t gob_payload_bound(int dtype, size_t count);
size_t gob_encode_payload(int dtype, const void* data, size_t count, uint8_t* out);
int gob_decode_payload(int dtype, const uint8_t* in, size_t len, void* out, size_t capacity, size_t* count_out);
size_t gob_envelope_bound(size_t payload);
size_t gob_encode_envelope(int64_t tag, const uint8_t* payload, size_t plen, uint8_t* out);
int gob_decode_envelope_body(const uint8_t* in, size_t len, int64_t* tag, const uint8_t** payload, size_t* plen);
size_t gob_get_uint(const uint8_t* p, size_t avail, uint64_t* u);
size_t gob_get_int(const uint8_t* p, size_t avail, int64_t* i);
/* from collectives.c */
void oracle_fill_i64(uint64_t seed, size_t count, int64_t* out);
void oracle_fill_f32(uint64_t seed, size_t count, float* out);
void oracle_fill_f64(uint64_t seed, size_t count, double* out);

enum { DT_U8 = 0, DT_I64 = 1, DT_F32 = 2, DT_F64 = 3 };
enum { COLL_ALLREDUCE = 0, COLL_BCAST = 1, COLL_ALLGATHER = 2, COLL_PINGPONG = 3, COLL_ALLREDUCE_NAIVE = 4 };
#define MAXR 8

typedef struct {
  uint8_t* p;
  size_t cap;
} growbuf;
static uint8_t* grow(growbuf* g, size_t need) {
  if (g->cap < need) {
    free(g->p);
    g->cap = need + need / 4 + 4096;
    g->p = (uint8_t*)malloc(g->cap);
  }
  return g->p;
}

typedef struct { /* buffered socket reader, the role of the gob Decoder's internal buffer */
  int fd;
  uint8_t* buf;
  size_t cap, lo, hi;
} reader;

struct world;
typedef struct rank_ctx {
  struct world* w;
  int rank;
  int dial[MAXR], listen[MAXR];
  reader rd_dial[MAXR], rd_listen[MAXR];
  growbuf enc, raw, env, msg, rawin; /* bytes.Buffer, GobEncode copy, encoder buffer, decoder buffer, GobDecode copy */
  /* helper thread that runs Send concurrently with Receive */
  pthread_t helper;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int job; /* 0 idle, 1 pending, 2 quit */
  const void* j_data; size_t j_count; int j_dtype, j_dest, j_tag, j_rc;
  growbuf henc, hraw, henv, hmsg; /* the helper's own buffers */
} rank_ctx;

typedef struct world {
  int n;
  rank_ctx r[MAXR];
  /* same-rank rendezvous (network.go:388-446): one slot is enough for the benches */
  pthread_mutex_t lmu;
  pthread_cond_t lcv;
  const uint8_t* lbytes; size_t llen; int lstate; /* 0 empty, 1 posted, 2 taken */
} world;

static int write_all(int fd, const uint8_t* p, size_t n) {
  while (n) {
    ssize_t w = send(fd, p, n, MSG_NOSIGNAL);
    if (w < 0) { if (errno == EINTR) continue; return -1; }
    p += w; n -= (size_t)w;
  }
  return 0;
}
static int fill(reader* r, size_t need) { /* make `need` bytes available at r->buf+r->lo */
  if (r->hi - r->lo >= need) return 0;
  if (need > r->cap || r->lo + need > r->cap) {
    size_t have = r->hi - r->lo;
    if (need > r->cap) {
      size_t nc = need + (1u << 16);
      uint8_t* nb = (uint8_t*)malloc(nc);
      memcpy(nb, r->buf + r->lo, have);
      free(r->buf); r->buf = nb; r->cap = nc;
    } else memmove(r->buf, r->buf + r->lo, have);
    r->lo = 0; r->hi = have;
  }
  while (r->hi - r->lo < need) {
    ssize_t k = recv(r->fd, r->buf + r->hi, r->cap - r->hi, 0);
    if (k < 0) { if (errno == EINTR) continue; return -1; }
    if (k == 0) return -1;
    r->hi += (size_t)k;
  }
  return 0;
}
/* next length-prefixed gob message; *body points into the reader's buffer */
static int next_msg(reader* r, const uint8_t** body, size_t* len) {
  if (fill(r, 1)) return -1;
  size_t pre = 1;
  if (r->buf[r->lo] >= 128) pre = 1 + (size_t)(-(int8_t)r->buf[r->lo]);
  if (fill(r, pre)) return -1;
  uint64_t l;
  if (!gob_get_uint(r->buf + r->lo, pre, &l)) return -1;
  if (fill(r, pre + l)) return -1;
  *body = r->buf + r->lo + pre; *len = (size_t)l;
  r->lo += pre + l;
  if (r->lo == r->hi) r->lo = r->hi = 0;
  return 0;
}
/* read one envelope (skipping its type descriptors) */
static int read_envelope(reader* r, int64_t* tag, const uint8_t** payload, size_t* plen) {
  for (;;) {
    const uint8_t* b; size_t l;
    if (next_msg(r, &b, &l)) return -1;
    int64_t id;
    if (!gob_get_int(b, l, &id)) return -1;
    if (id < 0) continue;
    return gob_decode_envelope_body(b, l, tag, payload, plen);
  }
}

static int do_send(rank_ctx* c, growbuf* enc, growbuf* raw, growbuf* env, const void* data, size_t count, int dtype, int dest, int tag) {
  world* w = c->w;
  uint8_t* e = grow(enc, gob_payload_bound(dtype, count));
  size_t elen = gob_encode_payload(dtype, data, count, e);
  if (dest == c->rank) { /* local.Send */
    pthread_mutex_lock(&w->lmu);
    while (w->lstate != 0) pthread_cond_wait(&w->lcv, &w->lmu);
    w->lbytes = e; w->llen = elen; w->lstate = 1;
    pthread_cond_broadcast(&w->lcv);
    while (w->lstate != 2) pthread_cond_wait(&w->lcv, &w->lmu);
    w->lstate = 0;
    pthread_cond_broadcast(&w->lcv);
    pthread_mutex_unlock(&w->lmu);
    return 0;
  }
  uint8_t* rw = grow(raw, elen + 1);
  memcpy(rw, e, elen); /* Raw.GobEncode: b := make([]byte, len(r)); copy(b, r) */
  uint8_t* v = grow(env, gob_envelope_bound(elen));
  size_t vlen = gob_encode_envelope(tag, rw, elen, v);
  if (write_all(c->dial[dest], v, vlen)) return -1;
  int64_t atag; const uint8_t* ap; size_t al;
  if (read_envelope(&c->rd_dial[dest], &atag, &ap, &al)) return -1; /* the ack */
  return atag == tag ? 0 : -1;
}

int ref_send(rank_ctx* c, const void* data, size_t count, int dtype, int dest, int tag) {
  return do_send(c, &c->enc, &c->raw, &c->env, data, count, dtype, dest, tag);
}

int ref_recv(rank_ctx* c, void* out, size_t capacity, size_t* count_out, int dtype, int src, int tag) {
  world* w = c->w;
  if (src == c->rank) { /* local.Receive */
    pthread_mutex_lock(&w->lmu);
    while (w->lstate != 1) pthread_cond_wait(&w->lcv, &w->lmu);
    uint8_t* cp = grow(&c->rawin, w->llen + 1);
    size_t l = w->llen;
    memcpy(cp, w->lbytes, l); /* keep the bytes valid after the sender returns */
    w->lstate = 2;
    pthread_cond_broadcast(&w->lcv);
    pthread_mutex_unlock(&w->lmu);
    return gob_decode_payload(dtype, cp, l, out, capacity, count_out);
  }
  int64_t mtag; const uint8_t* p; size_t plen;
  if (read_envelope(&c->rd_listen[src], &mtag, &p, &plen)) return -1;
  if (mtag != tag) return -1;
  uint8_t* cp = grow(&c->rawin, plen + 1);
  memcpy(cp, p, plen); /* Raw.GobDecode: copy(*r, b) */
  uint8_t ack[160];
  size_t alen = gob_encode_envelope(mtag, NULL, 0, ack);
  if (write_all(c->listen[src], ack, alen)) return -1;
  return gob_decode_payload(dtype, cp, plen, out, capacity, count_out);
}

/* ---- helper thread: "go mpi.Send(...)" ----------------------------------------------------- */
static void* helper_main(void* arg) {
  rank_ctx* c = (rank_ctx*)arg;
  pthread_mutex_lock(&c->mu);
  for (;;) {
    while (c->job == 0) pthread_cond_wait(&c->cv, &c->mu);
    if (c->job == 2) break;
    pthread_mutex_unlock(&c->mu);
    int rc = do_send(c, &c->henc, &c->hraw, &c->henv, c->j_data, c->j_count, c->j_dtype, c->j_dest, c->j_tag);
    pthread_mutex_lock(&c->mu);
    c->j_rc = rc; c->job = 0;
    pthread_cond_broadcast(&c->cv);
  }
  pthread_mutex_unlock(&c->mu);
  return NULL;
}
static void send_async(rank_ctx* c, const void* data, size_t count, int dtype, int dest, int tag) {
  pthread_mutex_lock(&c->mu);
  c->j_data = data; c->j_count = count; c->j_dtype = dtype; c->j_dest = dest; c->j_tag = tag; c->job = 1;
  pthread_cond_broadcast(&c->cv);
  pthread_mutex_unlock(&c->mu);
}
static int send_wait(rank_ctx* c) {
  pthread_mutex_lock(&c->mu);
  while (c->job == 1) pthread_cond_wait(&c->cv, &c->mu);
  int rc = c->j_rc;
  pthread_mutex_unlock(&c->mu);
  return rc;
}

/* ---- world --------------------------------------------------------------------------------- */
static void reader_init(reader* r, int fd) { r->fd = fd; r->cap = 1u << 16; r->buf = (uint8_t*)malloc(r->cap); r->lo = r->hi = 0; }

static void start_helper(rank_ctx* c) {
  pthread_mutex_init(&c->mu, NULL);
  pthread_cond_init(&c->cv, NULL);
  pthread_create(&c->helper, NULL, helper_main, c);
}

/* sockets only: the per-rank helper threads are started by the caller (threads do not survive fork) */
static world* world_sockets(int n) {
  if (n < 1 || n > MAXR) return NULL;
  world* w = (world*)calloc(1, sizeof(world));
  w->n = n;
  pthread_mutex_init(&w->lmu, NULL);
  pthread_cond_init(&w->lcv, NULL);
  int lst[MAXR]; int port[MAXR];
  for (int i = 0; i < n; ++i) {
    lst[i] = socket(AF_INET, SOCK_STREAM, 0);
    struct sockaddr_in sa; memset(&sa, 0, sizeof sa);
    sa.sin_family = AF_INET; sa.sin_addr.s_addr = htonl(INADDR_LOOPBACK); sa.sin_port = 0;
    if (bind(lst[i], (struct sockaddr*)&sa, sizeof sa) || listen(lst[i], 64)) return NULL;
    socklen_t sl = sizeof sa;
    getsockname(lst[i], (struct sockaddr*)&sa, &sl);
    port[i] = ntohs(sa.sin_port);
  }
  for (int i = 0; i < n; ++i) {
    rank_ctx* c = &w->r[i];
    c->w = w; c->rank = i;
    for (int j = 0; j < n; ++j) c->dial[j] = c->listen[j] = -1;
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      if (i == j) continue;
      int s = socket(AF_INET, SOCK_STREAM, 0);
      struct sockaddr_in sa; memset(&sa, 0, sizeof sa);
      sa.sin_family = AF_INET; sa.sin_addr.s_addr = htonl(INADDR_LOOPBACK); sa.sin_port = htons((uint16_t)port[j]);
      if (connect(s, (struct sockaddr*)&sa, sizeof sa)) return NULL;
      int a = accept(lst[j], NULL, NULL);
      if (a < 0) return NULL;
      /* Go's net package sets TCP_NODELAY on every TCP connection by default */
      int one = 1;
      setsockopt(s, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
      setsockopt(a, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
      w->r[i].dial[j] = s; w->r[j].listen[i] = a;
      reader_init(&w->r[i].rd_dial[j], s);
      reader_init(&w->r[j].rd_listen[i], a);
    }
  for (int i = 0; i < n; ++i) close(lst[i]);
  return w;
}

world* ref_world_create(int n) {
  world* w = world_sockets(n);
  if (!w) return NULL;
  for (int i = 0; i < n; ++i) start_helper(&w->r[i]);
  return w;
}

void ref_world_destroy(world* w) {
  if (!w) return;
  for (int i = 0; i < w->n; ++i) {
    rank_ctx* c = &w->r[i];
    pthread_mutex_lock(&c->mu); c->job = 2; pthread_cond_broadcast(&c->cv); pthread_mutex_unlock(&c->mu);
    pthread_join(c->helper, NULL);
    for (int j = 0; j < w->n; ++j) {
      if (c->dial[j] >= 0) { close(c->dial[j]); free(c->rd_dial[j].buf); }
      if (c->listen[j] >= 0) { close(c->listen[j]); free(c->rd_listen[j].buf); }
    }
    free(c->enc.p); free(c->raw.p); free(c->env.p); free(c->msg.p); free(c->rawin.p);
    free(c->henc.p); free(c->hraw.p); free(c->henv.p); free(c->hmsg.p);
  }
  free(w);
}
rank_ctx* ref_world_rank(world* w, int r) { return &w->r[r]; }

/* ---- collectives composed from Send/Receive ------------------------------------------------- */
static size_t esz(int dt) { return dt == DT_U8 ? 1 : dt == DT_F32 ? 4 : 8; }

static void add_into(int dtype, void* acc, const void* x, size_t count) {
  if (dtype == DT_F32) { float* a = (float*)acc; const float* b = (const float*)x; for (size_t i = 0; i < count; ++i) a[i] += b[i]; }
  else if (dtype == DT_F64) { double* a = (double*)acc; const double* b = (const double*)x; for (size_t i = 0; i < count; ++i) a[i] += b[i]; }
  else { uint64_t* a = (uint64_t*)acc; const uint64_t* b = (const uint64_t*)x; for (size_t i = 0; i < count; ++i) a[i] += b[i]; }
}

/* in place on buf (count elements); tmp holds one chunk */
int ref_allreduce_ring(rank_ctx* c, void* buf, size_t count, int dtype, void* tmp) {
  const int n = c->w->n, r = c->rank;
  if (n == 1) { /* a world of one still goes through Send/Receive to itself */
    send_async(c, buf, count, dtype, r, 0);
    size_t got;
    int rc = ref_recv(c, tmp, count, &got, dtype, r, 0);
    rc |= send_wait(c);
    memcpy(buf, tmp, count * esz(dtype));
    return rc;
  }
  const size_t per = (count + n - 1) / n, es = esz(dtype);
  const int next = (r + 1) % n, prev = (r + n - 1) % n;
  for (int s = 0; s < 2 * (n - 1); ++s) {
    const int rs = s < n - 1;
    const int sc = rs ? (r - s + 2 * n) % n : (r - (s - (n - 1)) + 1 + 2 * n) % n; /* chunk I send */
    const int rc_ = (sc + n - 1) % n;                                               /* chunk I receive */
    size_t slo = per * sc < count ? per * sc : count, shi = slo + per < count ? slo + per : count;
    size_t rlo = per * rc_ < count ? per * rc_ : count, rhi = rlo + per < count ? rlo + per : count;
    send_async(c, (char*)buf + slo * es, shi - slo, dtype, next, s);
    size_t got;
    int e = ref_recv(c, rs ? tmp : (void*)((char*)buf + rlo * es), rhi - rlo, &got, dtype, prev, s);
    if (e == 0 && rs) add_into(dtype, (char*)buf + rlo * es, tmp, rhi - rlo);
    e |= send_wait(c);
    if (e) return -1;
  }
  return 0;
}

/* gather to rank 0, rank-ordered sum, root sends the result back (SURVEY.md 8(c) composition) */
int ref_allreduce_naive(rank_ctx* c, void* buf, size_t count, int dtype, void* tmp) {
  const int n = c->w->n, r = c->rank;
  size_t got;
  if (r == 0) {
    for (int p = 1; p < n; ++p) { if (ref_recv(c, tmp, count, &got, dtype, p, 1)) return -1; add_into(dtype, buf, tmp, count); }
    for (int p = 1; p < n; ++p) if (ref_send(c, buf, count, dtype, p, 2)) return -1;
    return 0;
  }
  if (ref_send(c, buf, count, dtype, 0, 1)) return -1;
  return ref_recv(c, buf, count, &got, dtype, 0, 2);
}

int ref_bcast(rank_ctx* c, void* buf, size_t count, int dtype, int root) {
  const int n = c->w->n, r = c->rank;
  size_t got;
  if (r == root) { for (int p = 0; p < n; ++p) if (p != root && ref_send(c, buf, count, dtype, p, 3)) return -1; return 0; }
  return ref_recv(c, buf, count, &got, dtype, root, 3);
}

/* recv holds n*count; own block already in place */
int ref_allgather_ring(rank_ctx* c, void* recv, size_t count, int dtype) {
  const int n = c->w->n, r = c->rank;
  const size_t es = esz(dtype);
  const int next = (r + 1) % n, prev = (r + n - 1) % n;
  for (int s = 0; s < n - 1; ++s) {
    const int sb = (r - s + 2 * n) % n, rb = (r - s - 1 + 2 * n) % n;
    send_async(c, (char*)recv + (size_t)sb * count * es, count, dtype, next, 10 + s);
    size_t got;
    int e = ref_recv(c, (char*)recv + (size_t)rb * count * es, count, &got, dtype, prev, 10 + s);
    e |= send_wait(c);
    if (e) return -1;
  }
  return 0;
}

/* ---- timed harness -------------------------------------------------------------------------- */
typedef struct {
  world* w; int rank, coll, dtype; size_t count; int iters, warmup; uint64_t seed;
  pthread_barrier_t* bar; double* seconds; void* out0; int rc;
} job;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void fill_input(int dtype, uint64_t seed, size_t count, void* p) {
  if (dtype == DT_F32) oracle_fill_f32(seed, count, (float*)p);
  else if (dtype == DT_F64) oracle_fill_f64(seed, count, (double*)p);
  else oracle_fill_i64(seed, count, (int64_t*)p);
}

static void* bench_main(void* arg) {
  job* j = (job*)arg;
  rank_ctx* c = &j->w->r[j->rank];
  const int n = j->w->n;
  const size_t es = esz(j->dtype);
  const size_t total = j->coll == COLL_ALLGATHER ? j->count * n : j->count;
  void* buf = malloc(total * es + 16);
  void* tmp = malloc(j->count * es + 16);
  double t0 = 0, timed = 0;
  for (int it = 0; it < j->warmup + j->iters; ++it) {
    /* fresh inputs every iteration (outside the timed region) */
    if (j->coll == COLL_ALLGATHER) fill_input(j->dtype, j->seed + j->rank, j->count, (char*)buf + (size_t)j->rank * j->count * es);
    else if (j->coll == COLL_BCAST) { if (j->rank == 0) fill_input(j->dtype, j->seed, j->count, buf); else memset(buf, 0xff, j->count * es); }
    else fill_input(j->dtype, j->seed + j->rank, j->count, buf);
    pthread_barrier_wait(j->bar);
    if (j->rank == 0) t0 = now_s();
    int rc = 0;
    switch (j->coll) {
      case COLL_ALLREDUCE: rc = ref_allreduce_ring(c, buf, j->count, j->dtype, tmp); break;
      case COLL_ALLREDUCE_NAIVE: rc = ref_allreduce_naive(c, buf, j->count, j->dtype, tmp); break;
      case COLL_BCAST: rc = ref_bcast(c, buf, j->count, j->dtype, 0); break;
      case COLL_ALLGATHER: rc = ref_allgather_ring(c, buf, j->count, j->dtype); break;
      case COLL_PINGPONG: { /* bounce.go:85-138: even sends, odd returns it */
        size_t got;
        if (n < 2) { send_async(c, buf, j->count, j->dtype, 0, 0); rc = ref_recv(c, tmp, j->count, &got, j->dtype, 0, 0); rc |= send_wait(c); }
        else if (j->rank % 2 == 0 && j->rank + 1 < n) { rc = ref_send(c, buf, j->count, j->dtype, j->rank + 1, 0); rc |= ref_recv(c, tmp, j->count, &got, j->dtype, j->rank + 1, 0); }
        else if (j->rank % 2 == 1) { rc = ref_recv(c, tmp, j->count, &got, j->dtype, j->rank - 1, 0); rc |= ref_send(c, tmp, j->count, j->dtype, j->rank - 1, 0); }
        break;
      }
    }
    if (rc) j->rc = rc;
    pthread_barrier_wait(j->bar); /* every rank has finished the step: max over ranks */
    if (j->rank == 0 && it >= j->warmup) timed += now_s() - t0; /* input generation stays outside */
  }
  if (j->rank == 0) {
    *j->seconds = timed / (j->iters > 0 ? j->iters : 1);
    if (j->out0) memcpy(j->out0, j->coll == COLL_PINGPONG ? tmp : buf, total * es);
  }
  free(buf); free(tmp);
  return NULL;
}

/* Runs `warmup`+`iters` iterations of one collective over a fresh world of n ranks with the same
 * synthetic inputs the GPU bench uses (seed + rank).  seconds_per_iter is the mean over the timed
 * iterations of the wall clock between the barrier before and the barrier after the step
 * (= max over ranks); regenerating the inputs is not timed.
 * out_rank0 (optional) receives rank 0's final buffer for the parity check.  Returns 0 on success. */
int ref_bench(int coll, int dtype, int n, size_t count, int iters, int warmup, uint64_t seed, double* seconds_per_iter, void* out_rank0) {
  world* w = ref_world_create(n);
  if (!w) return -1;
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, (unsigned)n);
  pthread_t th[MAXR]; job jobs[MAXR];
  double secs = 0;
  for (int r = 0; r < n; ++r) {
    job j = {w, r, coll, dtype, count, iters, warmup, seed, &bar, &secs, r == 0 ? out_rank0 : NULL, 0};
    jobs[r] = j;
    pthread_create(&th[r], NULL, bench_main, &jobs[r]);
  }
  int rc = 0;
  for (int r = 0; r < n; ++r) { pthread_join(th[r], NULL); rc |= jobs[r].rc; }
  pthread_barrier_destroy(&bar);
  ref_world_destroy(w);
  if (seconds_per_iter) *seconds_per_iter = secs;
  return rc;
}

/* The same harness with one OS process per rank (gompirun.go:85).  The sockets of the whole mesh are
 * created first, then n children are forked; child r closes the other ranks' ends, starts its helper
 * thread and runs bench_main for rank r.  The step barrier, the timing word and rank 0's result live in
 * a shared anonymous mapping.  Returns 0 on success. */
typedef struct {
  pthread_barrier_t bar;
  double secs;
  int rc[MAXR];
} shared_hdr;

int ref_bench_procs(int coll, int dtype, int n, size_t count, int iters, int warmup, uint64_t seed, double* seconds_per_iter, void* out_rank0) {
  if (n < 1 || n > MAXR) return -1;
  const size_t total = (coll == COLL_ALLGATHER ? count * (size_t)n : count) * esz(dtype);
  const size_t map_len = sizeof(shared_hdr) + total + 64;
  shared_hdr* sh = (shared_hdr*)mmap(NULL, map_len, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (sh == MAP_FAILED) return -1;
  pthread_barrierattr_t ba;
  pthread_barrierattr_init(&ba);
  pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);
  pthread_barrier_init(&sh->bar, &ba, (unsigned)n);
  sh->secs = 0;
  for (int r = 0; r < n; ++r) sh->rc[r] = -99;
  world* w = world_sockets(n);
  if (!w) { munmap(sh, map_len); return -1; }
  pid_t pid[MAXR];
  for (int r = 0; r < n; ++r) {
    pid[r] = fork();
    if (pid[r] == 0) {
      for (int i = 0; i < n; ++i) { /* keep only this rank's ends of the mesh */
        if (i == r) continue;
        for (int j = 0; j < n; ++j) {
          if (w->r[i].dial[j] >= 0) close(w->r[i].dial[j]);
          if (w->r[i].listen[j] >= 0) close(w->r[i].listen[j]);
        }
      }
      start_helper(&w->r[r]);
      job j = {w, r, coll, dtype, count, iters, warmup, seed, &sh->bar, &sh->secs, r == 0 ? (void*)(sh + 1) : NULL, 0};
      bench_main(&j);
      sh->rc[r] = j.rc;
      _exit(0);
    }
    if (pid[r] < 0) { sh->rc[r] = -1; }
  }
  /* the parent holds no end of the mesh */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      if (w->r[i].dial[j] >= 0) close(w->r[i].dial[j]);
      if (w->r[i].listen[j] >= 0) close(w->r[i].listen[j]);
    }
  int rc = 0;
  for (int r = 0; r < n; ++r) {
    int st = 0;
    if (pid[r] > 0) waitpid(pid[r], &st, 0);
    if (pid[r] <= 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0 || sh->rc[r] != 0) rc = -2;
  }
  if (seconds_per_iter) *seconds_per_iter = sh->secs;
  if (out_rank0 && rc == 0) memcpy(out_rank0, sh + 1, total);
  pthread_barrier_destroy(&sh->bar);
  munmap(sh, map_len);
  free(w);
  return rc;
}


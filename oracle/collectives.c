/*
 * oracle/collectives.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library.  libb200mpi never calls it.
 *
 * CPU restatement of what a value means after each call on the hot path:
 *
 *   Send/Receive  value-exact delivery of one typed slice
 *                 (/root/reference/network.go:518-602; gob round trip is value exact for
 *                 []byte, []int64, []float64, []float32 -- see oracle/gob.c)           -> copy
 *   Bcast         every rank's buffer == root's                                           -> copy
 *   Allgather     recv[r*count .. (r+1)*count) == send of rank r, rank = index in the sorted
 *                 address list (/root/reference/network.go:94-109)                        -> concat
 *   Allreduce     the composition a reference user writes from Send/Receive: gather, then
 *                 acc = x_0; acc = op(acc, x_r), r = 1..n-1, in the element type, then Bcast.
 *   Reduce        the same value, on root only.
 *   ReduceScatter rank j holds the Allreduce of everyone's block j (n blocks of `count`).
 *   Alltoall      recv_j[r*count ..] == send_r[j*count ..]: what n*n Send/Receive pairs deliver.
 *
 * PARITY UNPINNED for Bcast/Allreduce/Allgather: the reference has no such functions
 * (/root/reference/mpi.go:130 is a commented stub) and no tests, so there is no golden vector to
 * pin these definitions on; they are the semantics SURVEY.md 8(c) fixes.  Send/Receive is pinned
 * only by the reference's own round-trip equality checks (examples/bounce/bounce.go:105,133).
 *
 * Floating-point sums depend on association, so the oracle restates each kernel's order:
 *   ORDER_RANK  ((x0+x1)+x2)+...                 one-shot, two-shot      (kernels.cuh)
 *   ORDER_TREE  ((x0+x1)+(x2+x3))+(...)          one-shot shuffle, n in {2,4,8}
 *   ORDER_RING  chunk c: ((x_c+x_{c+1})+...)+x_{c-1}; chunks of ceil((count/EPV)/n) 16-byte
 *               groups; the count%EPV tail in rank order                       allreduce_ring_kernel
 *   ORDER_F64   every element accumulated in double (long double for f64) in rank order and
 *               rounded once: the tolerance reference for NVLS, where the switch picks the order
 * Integer sums wrap (Go int64 semantics) and are order independent.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

enum { DT_U8 = 0, DT_I64 = 1, DT_F32 = 2, DT_F64 = 3 };
enum { OP_SUM = 0, OP_MAX = 1, OP_MIN = 2 };
enum { ORDER_RANK = 0, ORDER_TREE = 1, ORDER_RING = 2, ORDER_F64 = 3 };

static size_t esize(int dt) { return dt == DT_U8 ? 1 : dt == DT_F32 ? 4 : 8; }

#define DEF_APPLY(NAME, T, ADD)                                  \
  static T NAME(int op, T a, T b) {                              \
    switch (op) {                                                \
      case OP_SUM: return ADD;                                   \
      case OP_MAX: return b > a ? b : a;                         \
      default: return b < a ? b : a;                             \
    }                                                            \
  }
DEF_APPLY(apply_f32, float, a + b)
DEF_APPLY(apply_f64, double, a + b)
DEF_APPLY(apply_i64, int64_t, (int64_t)((uint64_t)a + (uint64_t)b))

/* element e of rank r */
#define AT(T, r, e) (((const T*)in[r])[e])

#define DEF_REDUCE(NAME, T, APPLY, WIDE)                                                         \
  static void NAME(int op, int order, int n, size_t count, const void* const* in, T* out) {      \
    const size_t epv = 16 / sizeof(T);                                                           \
    const size_t groups = count / epv;                                                           \
    const size_t per = (groups + (size_t)n - 1) / (size_t)n;                                     \
    for (size_t e = 0; e < count; ++e) {                                                         \
      T acc;                                                                                     \
      if (order == ORDER_TREE && (n == 2 || n == 4 || n == 8)) {                                 \
        T x[8];                                                                                  \
        for (int r = 0; r < n; ++r) x[r] = AT(T, r, e);                                          \
        for (int m = 1; m < n; m <<= 1)                                                          \
          for (int r = 0; r < n; r += 2 * m) x[r] = APPLY(op, x[r], x[r + m]);                   \
        acc = x[0];                                                                              \
      } else if (order == ORDER_RING && e < groups * epv && per > 0) {                           \
        int c = (int)((e / epv) / per);                                                          \
        acc = AT(T, c, e);                                                                       \
        for (int k = 1; k < n; ++k) acc = APPLY(op, acc, AT(T, (c + k) % n, e));                 \
      } else if (order == ORDER_F64 && op == OP_SUM) {                                           \
        WIDE w = (WIDE)AT(T, 0, e);                                                              \
        for (int r = 1; r < n; ++r) w += (WIDE)AT(T, r, e);                                      \
        acc = (T)w;                                                                              \
      } else {                                                                                   \
        acc = AT(T, 0, e);                                                                       \
        for (int r = 1; r < n; ++r) acc = APPLY(op, acc, AT(T, r, e));                           \
      }                                                                                          \
      out[e] = acc;                                                                              \
    }                                                                                            \
  }
DEF_REDUCE(reduce_f32, float, apply_f32, double)
DEF_REDUCE(reduce_f64, double, apply_f64, long double)
DEF_REDUCE(reduce_i64, int64_t, apply_i64, int64_t)

/* out <- reduction over in[0..n) ; returns 0, or -1 for an unsupported dtype */
int oracle_allreduce(int dtype, int op, int order, int n, size_t count, const void* const* in, void* out) {
  if (n < 1 || n > 8) return -1;
  switch (dtype) {
    case DT_F32: reduce_f32(op, order, n, count, in, (float*)out); return 0;
    case DT_F64: reduce_f64(op, order, n, count, in, (double*)out); return 0;
    case DT_I64: reduce_i64(op, order, n, count, in, (int64_t*)out); return 0;
  }
  return -1;
}

/* out[r*count ...] <- in[r] */
int oracle_allgather(int dtype, int n, size_t count, const void* const* in, void* out) {
  const size_t b = count * esize(dtype);
  for (int r = 0; r < n; ++r) memcpy((char*)out + (size_t)r * b, in[r], b);
  return 0;
}

/* ReduceScatter: in[r] holds n*count elements; out (count elements) is what rank `me` receives */
int oracle_reduce_scatter(int dtype, int op, int order, int n, int me, size_t count, const void* const* in, void* out) {
  const void* blk[8];
  if (n < 1 || n > 8 || me < 0 || me >= n) return -1;
  for (int r = 0; r < n; ++r) blk[r] = (const char*)in[r] + (size_t)me * count * esize(dtype);
  return oracle_allreduce(dtype, op, order, n, count, blk, out);
}

/* Alltoall: in[r] holds n*count elements; out (n*count elements) is what rank `me` receives */
int oracle_alltoall(int dtype, int n, int me, size_t count, const void* const* in, void* out) {
  const size_t b = count * esize(dtype);
  for (int r = 0; r < n; ++r) memcpy((char*)out + (size_t)r * b, (const char*)in[r] + (size_t)me * b, b);
  return 0;
}

/* out <- root's buffer (what every rank must hold afterwards) */
int oracle_bcast(int dtype, size_t count, const void* root_buf, void* out) {
  memcpy(out, root_buf, count * esize(dtype));
  return 0;
}

/* Send/Receive: the receiver's value and length equal the sender's */
int oracle_sendrecv(int dtype, size_t count, const void* sent, void* received, size_t* count_out) {
  memcpy(received, sent, count * esize(dtype));
  if (count_out) *count_out = count;
  return 0;
}

/* splitmix64: the synthetic "indices" / payload generator named in SURVEY.md 8(d) */
uint64_t oracle_splitmix64(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
/* the same streams from element `start` on: large buffers are generated and checked block by block */
void oracle_fill_i64_at(uint64_t seed, size_t start, size_t count, int64_t* out) {
  for (size_t i = 0; i < count; ++i) out[i] = (int64_t)oracle_splitmix64(seed, start + i);
}
void oracle_fill_f32_at(uint64_t seed, size_t start, size_t count, float* out) {
  for (size_t i = 0; i < count; ++i) out[i] = (float)(oracle_splitmix64(seed, start + i) >> 40) * (1.0f / 16777216.0f);
}
void oracle_fill_f64_at(uint64_t seed, size_t start, size_t count, double* out) {
  for (size_t i = 0; i < count; ++i) out[i] = (double)(oracle_splitmix64(seed, start + i) >> 11) * (1.0 / 9007199254740992.0);
}
void oracle_fill_i64(uint64_t seed, size_t count, int64_t* out) {
  for (size_t i = 0; i < count; ++i) out[i] = (int64_t)oracle_splitmix64(seed, i);
}
/* uniform [0,1) with 24 random bits: exactly representable in f32 (mirrors rand.Float64 use, bounce.go:75) */
void oracle_fill_f32(uint64_t seed, size_t count, float* out) {
  for (size_t i = 0; i < count; ++i) out[i] = (float)(oracle_splitmix64(seed, i) >> 40) * (1.0f / 16777216.0f);
}
void oracle_fill_f64(uint64_t seed, size_t count, double* out) {
  for (size_t i = 0; i < count; ++i) out[i] = (double)(oracle_splitmix64(seed, i) >> 11) * (1.0 / 9007199254740992.0);
}

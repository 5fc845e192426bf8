/*
 * oracle/gob.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see collectives.c header).
 *
 * Restatement of the wire format the reference's data plane spends its time in: Go's
 * encoding/gob, which is NOT under /root/reference (standard library, no pinned version: the
 * reference has no go.mod; imports at /root/reference/network.go:3-12).  Call sites restated:
 *   payload encode   gob.NewEncoder(&buf).Encode(data)              network.go:537-542
 *   envelope encode  enc.Encode(message{Tag, Bytes})                network.go:562-563, 620-621
 *   envelope decode  gob.NewDecoder(conn).Decode(&m)                network.go:553, 609
 *   payload decode   gob.NewDecoder(buf).Decode(data)               network.go:594-601
 * A fresh Encoder per message means the type descriptors are re-sent every time, as here.
 *
 * Published format (package documentation of encoding/gob), pinned by tests/test_oracle_gob.py on
 * the documentation's own known answers: uint 7 -> 07, 256 -> FE 01 00; int -129 -> FE 01 01;
 * float 17.0 -> FE 31 40; and the complete Point{22,33} stream (type descriptor + value).
 * The slice / GobEncoder type-descriptor bytes follow the documented wireType layout but could
 * not be diffed against a Go toolchain (none in this image): they cost tens of bytes per message
 * and do not affect any value.
 *
 *   unsigned   < 128: one byte; else (negated byte count) then big-endian minimal bytes
 *   signed     zig-zag-ish: u = i<<1, or (^i<<1)|1 when negative, then as unsigned
 *   float      float64 bits, byte-reversed, then as unsigned (float32 is widened first)
 *   []byte, string   unsigned count + raw bytes
 *   other slices     unsigned count + elements
 *   struct           (field delta, value) pairs, zero-valued fields omitted, 00 terminator
 *   message          unsigned byte length, then signed type id, then (non-struct) a 00, value
 *   type ids         int 2, uint 3, float 4, []byte 5, string 6; first user type 65
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

enum { DT_U8 = 0, DT_I64 = 1, DT_F32 = 2, DT_F64 = 3, DT_STRING = 4 };

size_t gob_put_uint(uint8_t* p, uint64_t u) {
  if (u < 128) { p[0] = (uint8_t)u; return 1; }
  int n = 0;
  for (uint64_t t = u; t; t >>= 8) ++n;
  p[0] = (uint8_t)(-n);
  for (int i = 0; i < n; ++i) p[1 + i] = (uint8_t)(u >> (8 * (n - 1 - i)));
  return (size_t)n + 1;
}
size_t gob_put_int(uint8_t* p, int64_t i) {
  uint64_t u = i < 0 ? ((~(uint64_t)i) << 1) | 1 : ((uint64_t)i << 1);
  return gob_put_uint(p, u);
}
size_t gob_put_float(uint8_t* p, double f) {
  uint64_t b;
  memcpy(&b, &f, 8);
  return gob_put_uint(p, __builtin_bswap64(b));
}
/* returns bytes consumed, 0 on malformed input */
size_t gob_get_uint(const uint8_t* p, size_t avail, uint64_t* u) {
  if (avail == 0) return 0;
  if (p[0] < 128) { *u = p[0]; return 1; }
  int n = -(int8_t)p[0];
  if (n < 1 || n > 8 || (size_t)n + 1 > avail) return 0;
  uint64_t v = 0;
  for (int i = 0; i < n; ++i) v = (v << 8) | p[1 + i];
  *u = v;
  return (size_t)n + 1;
}
size_t gob_get_int(const uint8_t* p, size_t avail, int64_t* i) {
  uint64_t u;
  size_t k = gob_get_uint(p, avail, &u);
  if (!k) return 0;
  *i = (u & 1) ? (int64_t)~(u >> 1) : (int64_t)(u >> 1);
  return k;
}
size_t gob_get_float(const uint8_t* p, size_t avail, double* f) {
  uint64_t u;
  size_t k = gob_get_uint(p, avail, &u);
  if (!k) return 0;
  u = __builtin_bswap64(u);
  memcpy(f, &u, 8);
  return k;
}

/* wireType{SliceT: &sliceType{CommonType{Name, Id 65}, Elem}} as one length-prefixed message */
static size_t put_slice_typedef(uint8_t* out, const char* name, int elem_id) {
  uint8_t body[64];
  size_t n = 0, nl = strlen(name);
  n += gob_put_int(body + n, -65);
  body[n++] = 0x02;              /* wireType field 1: SliceT */
  body[n++] = 0x01;              /* sliceType field 0: CommonType */
  body[n++] = 0x01;              /* CommonType field 0: Name */
  n += gob_put_uint(body + n, nl);
  memcpy(body + n, name, nl); n += nl;
  body[n++] = 0x01;              /* CommonType field 1: Id */
  n += gob_put_int(body + n, 65);
  body[n++] = 0x00;              /* end CommonType */
  body[n++] = 0x01;              /* sliceType field 1: Elem */
  n += gob_put_int(body + n, elem_id);
  body[n++] = 0x00;              /* end sliceType */
  body[n++] = 0x00;              /* end wireType */
  size_t k = gob_put_uint(out, n);
  memcpy(out + k, body, n);
  return k + n;
}

/* worst-case encoded size of a payload */
size_t gob_payload_bound(int dtype, size_t count) {
  size_t per = (dtype == DT_U8 || dtype == DT_STRING) ? 1 : 9;
  return 96 + count * per;
}

/* gob.NewEncoder(&buf).Encode(slice): [type descriptor] + value message.  Returns bytes written. */
size_t gob_encode_payload(int dtype, const void* data, size_t count, uint8_t* out) {
  size_t n = 0;
  int type_id;
  switch (dtype) {
    case DT_U8: type_id = 5; break;
    case DT_STRING: type_id = 6; break;
    case DT_F64: n += put_slice_typedef(out, "[]float64", 4); type_id = 65; break;
    case DT_F32: n += put_slice_typedef(out, "[]float32", 4); type_id = 65; break;
    case DT_I64: n += put_slice_typedef(out, "[]int64", 2); type_id = 65; break;
    default: return 0;
  }
  /* the value message: length prefix is written last, so build the body after a 10-byte gap */
  uint8_t* body = out + n + 10;
  size_t b = 0;
  b += gob_put_int(body + b, type_id);
  body[b++] = 0x00; /* singleton: zero field delta */
  b += gob_put_uint(body + b, count);
  if (dtype == DT_U8 || dtype == DT_STRING) {
    memcpy(body + b, data, count);
    b += count;
  } else if (dtype == DT_F64) {
    const double* d = (const double*)data;
    for (size_t i = 0; i < count; ++i) b += gob_put_float(body + b, d[i]);
  } else if (dtype == DT_F32) {
    const float* d = (const float*)data;
    for (size_t i = 0; i < count; ++i) b += gob_put_float(body + b, (double)d[i]);
  } else {
    const int64_t* d = (const int64_t*)data;
    for (size_t i = 0; i < count; ++i) b += gob_put_int(body + b, d[i]);
  }
  uint8_t len[10];
  size_t k = gob_put_uint(len, b);
  memcpy(out + n, len, k);
  memmove(out + n + k, body, b);
  return n + k + b;
}

/* gob.NewDecoder(buf).Decode(&slice).  Returns 0 and *count_out, -1 malformed, -2 capacity. */
int gob_decode_payload(int dtype, const uint8_t* in, size_t len, void* out, size_t capacity, size_t* count_out) {
  size_t pos = 0;
  for (;;) {
    uint64_t mlen;
    size_t k = gob_get_uint(in + pos, len - pos, &mlen);
    if (!k || mlen > len - pos - k) return -1;
    pos += k;
    int64_t id;
    size_t kk = gob_get_int(in + pos, mlen, &id);
    if (!kk) return -1;
    if (id < 0) { pos += mlen; continue; } /* a type descriptor: skip */
    size_t p = pos + kk, end = pos + mlen;
    if (p >= end || in[p] != 0) return -1;  /* "non-zero delta for singleton" */
    ++p;
    uint64_t cnt;
    k = gob_get_uint(in + p, end - p, &cnt);
    if (!k) return -1;
    p += k;
    if (count_out) *count_out = (size_t)cnt;
    if (cnt > capacity) return -2;
    if (dtype == DT_U8 || dtype == DT_STRING) {
      if (cnt > end - p) return -1;
      memcpy(out, in + p, cnt);
    } else if (dtype == DT_F64) {
      double* d = (double*)out;
      for (uint64_t i = 0; i < cnt; ++i) { k = gob_get_float(in + p, end - p, &d[i]); if (!k) return -1; p += k; }
    } else if (dtype == DT_F32) {
      float* d = (float*)out;
      for (uint64_t i = 0; i < cnt; ++i) { double f; k = gob_get_float(in + p, end - p, &f); if (!k) return -1; d[i] = (float)f; p += k; }
    } else {
      int64_t* d = (int64_t*)out;
      for (uint64_t i = 0; i < cnt; ++i) { k = gob_get_int(in + p, end - p, &d[i]); if (!k) return -1; p += k; }
    }
    return 0;
  }
}

/* wireType{StructT: &structType{CommonType{Name, Id}, Field: [{Name, Id}...]}} as one
 * length-prefixed message.  With ("Point", 65, {"X","Y"}, {2,2}) this reproduces the package
 * documentation's example byte for byte (tests/test_oracle_gob.py). */
size_t gob_put_struct_typedef(uint8_t* out, const char* name, int id, int nfields, const char* const* fnames, const int* fids) {
  uint8_t b[256];
  size_t k = 0, nl = strlen(name);
  k += gob_put_int(b + k, -(int64_t)id);
  b[k++] = 0x03;                                   /* wireType field 2: StructT */
  b[k++] = 0x01;                                   /* structType field 0: CommonType */
  b[k++] = 0x01; k += gob_put_uint(b + k, nl); memcpy(b + k, name, nl); k += nl; /* Name */
  b[k++] = 0x01; k += gob_put_int(b + k, id);      /* Id */
  b[k++] = 0x00;                                   /* end CommonType */
  b[k++] = 0x01; k += gob_put_uint(b + k, (uint64_t)nfields); /* structType field 1: Field slice */
  for (int f = 0; f < nfields; ++f) {
    size_t fl = strlen(fnames[f]);
    b[k++] = 0x01; k += gob_put_uint(b + k, fl); memcpy(b + k, fnames[f], fl); k += fl;
    b[k++] = 0x01; k += gob_put_int(b + k, fids[f]);
    b[k++] = 0x00;
  }
  b[k++] = 0x00; b[k++] = 0x00;                    /* end structType, end wireType */
  size_t n = gob_put_uint(out, k);
  memcpy(out + n, b, k);
  return n + k;
}

/* message{Tag int; Bytes Raw} (network.go:511-514).  Descriptors: struct "message" (id 65, fields
 * Tag:int, Bytes:66) then GobEncoder type "Raw" (id 66); then the value. */
size_t gob_envelope_bound(size_t payload) { return payload + 128; }

size_t gob_encode_envelope(int64_t tag, const uint8_t* payload, size_t plen, uint8_t* out) {
  size_t n = 0;
  static const char* const fnames[2] = {"Tag", "Bytes"};
  static const int fids[2] = {2, 66};
  n += gob_put_struct_typedef(out, "message", 65, 2, fnames, fids);
  {
    uint8_t b[48];
    size_t k = 0;
    k += gob_put_int(b + k, -66);
    b[k++] = 0x05; /* wireType field 4: GobEncoderT */
    b[k++] = 0x01; b[k++] = 0x01; b[k++] = 3; memcpy(b + k, "Raw", 3); k += 3;
    b[k++] = 0x01; k += gob_put_int(b + k, 66); b[k++] = 0x00;
    b[k++] = 0x00; b[k++] = 0x00;
    n += gob_put_uint(out + n, k);
    memcpy(out + n, b, k); n += k;
  }
  uint8_t head[40];
  size_t h = 0;
  h += gob_put_int(head + h, 65);
  int field = -1;
  if (tag != 0) { head[h++] = 0x01; h += gob_put_int(head + h, tag); field = 0; }
  if (plen != 0) { head[h++] = (uint8_t)(1 - field); h += gob_put_uint(head + h, plen); }
  size_t body = h + plen + 1;
  n += gob_put_uint(out + n, body);
  memcpy(out + n, head, h); n += h;
  if (plen) { memcpy(out + n, payload, plen); n += plen; }
  out[n++] = 0x00;
  return n;
}

/* Parses one envelope value message body (after the descriptors were skipped by the reader).
 * Returns 0; *payload points into `in`. */
int gob_decode_envelope_body(const uint8_t* in, size_t len, int64_t* tag, const uint8_t** payload, size_t* plen) {
  int64_t id;
  size_t p = gob_get_int(in, len, &id);
  if (!p || id != 65) return -1;
  *tag = 0; *payload = in; *plen = 0;
  int field = -1;
  for (;;) {
    uint64_t delta;
    size_t k = gob_get_uint(in + p, len - p, &delta);
    if (!k) return -1;
    p += k;
    if (delta == 0) return 0;
    field += (int)delta;
    if (field == 0) { k = gob_get_int(in + p, len - p, tag); if (!k) return -1; p += k; }
    else if (field == 1) {
      uint64_t c;
      k = gob_get_uint(in + p, len - p, &c);
      if (!k || c > len - p - k) return -1;
      p += k; *payload = in + p; *plen = (size_t)c; p += c;
    } else return -1;
  }
}

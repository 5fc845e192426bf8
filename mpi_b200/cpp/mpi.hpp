// mpi.hpp -- C++ host-side mirror of the reference's Go package `mpi`
// (/root/reference/mpi.go:56-182, flags.go:10-50) over the C ABI of include/b200mpi.h.
//
// The reference is compiled Go and this image has no Go toolchain, so the host layer above the
// C ABI is written in C++ with the same names, argument meaning and error behaviour; the Go shim
// that binds the same ABI is in go/ (see INTEGRATION.md).
//
//   Go                                         here
//   mpi.Init() error                           mpi::Error mpi::Init()
//   mpi.Finalize()                             void mpi::Finalize()
//   mpi.Rank() / mpi.Size()                    int mpi::Rank() / mpi::Size()
//   mpi.Send(data interface{}, dst, tag)       mpi::Error mpi::Send(const T& data, int dst, int tag)
//   mpi.Receive(&data, src, tag)               mpi::Error mpi::Receive(T* data, int src, int tag)   (resizes *data like gob)
//   mpi.Register(impl)                         void mpi::Register(Interface*)  (second call throws = panics, mpi.go:61-67)
//   flag.Parse() + mpi.Flag*                   mpi::ParseFlags(argc, argv) + mpi::Flag*
//   new: Recv, Bcast, Allreduce, Allgather, Barrier, ReduceScatter, Reduce, Alltoall, Isend/Wait;
//        mpi::DeviceSlice<T> for heap-resident buffers
// `data` may be std::vector<double|float|int64_t|uint8_t>, std::string (sent as bytes, the
// "anything else is gob-encoded" path), or DeviceSlice<T>.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/b200mpi.h"

namespace mpi {

// ---- error: nil-able like Go's `error` -----------------------------------------------------------
struct Error {
  int code = 0;
  std::string msg;
  explicit operator bool() const { return code != 0; } // `if (err)` reads like `if err != nil`
  const std::string& String() const { return msg; }
};
struct TagExists : Error { // mpi.go:174-182
  int Tag = 0;
};
inline Error make_error(int rc) {
  Error e;
  if (rc != 0) {
    e.code = rc;
    e.msg = b200mpi_last_error();
  }
  return e;
}

enum Op { SUM = B200MPI_SUM, MAX = B200MPI_MAX, MIN = B200MPI_MIN };
using Raw = std::vector<uint8_t>; // mpi.go:75

template <typename T> struct dtype_of;
template <> struct dtype_of<uint8_t> { static constexpr int value = B200MPI_U8; };
template <> struct dtype_of<char> { static constexpr int value = B200MPI_U8; };
template <> struct dtype_of<int64_t> { static constexpr int value = B200MPI_I64; };
template <> struct dtype_of<long long> { static constexpr int value = B200MPI_I64; };
template <> struct dtype_of<float> { static constexpr int value = B200MPI_F32; };
template <> struct dtype_of<double> { static constexpr int value = B200MPI_F64; };

// Memory from the rank's peer-mapped device heap: collectives and Send/Receive use it in place.
template <typename T>
class DeviceSlice {
 public:
  DeviceSlice() = default;
  explicit DeviceSlice(size_t count) : n_(count), own_(true) {
    void* p = nullptr;
    if (b200mpi_alloc(count * sizeof(T), &p) != 0) throw std::runtime_error(b200mpi_last_error());
    p_ = static_cast<T*>(p);
  }
  DeviceSlice(T* p, size_t count) : p_(p), n_(count), own_(false) {}
  DeviceSlice(const DeviceSlice&) = delete;
  DeviceSlice& operator=(const DeviceSlice&) = delete;
  DeviceSlice(DeviceSlice&& o) noexcept : p_(o.p_), n_(o.n_), own_(o.own_) { o.p_ = nullptr; o.own_ = false; }
  ~DeviceSlice() { if (own_ && p_) b200mpi_free(p_); }
  T* data() const { return p_; }
  size_t size() const { return n_; }
  DeviceSlice Sub(size_t lo, size_t hi) const { return DeviceSlice(p_ + lo, hi - lo); } // s[lo:hi]
  Error CopyFromHost(const std::vector<T>& h) { return make_error(b200mpi_memcpy(p_, h.data(), std::min(h.size(), n_) * sizeof(T), 0)); }
  Error CopyToHost(std::vector<T>* h) const { h->resize(n_); return make_error(b200mpi_memcpy(h->data(), p_, n_ * sizeof(T), 1)); }

 private:
  T* p_ = nullptr;
  size_t n_ = 0;
  bool own_ = false;
};

// ---- flags.go ---------------------------------------------------------------------------------------
inline std::string FlagAddr;
inline std::vector<std::string> FlagAllAddrs; // AddrsFlag: comma split, appends (flags.go:22-27)
inline int64_t FlagInitTimeout = 0;           // DurationFlag, nanoseconds
inline std::string FlagProtocol = "tcp";
inline std::string FlagPassword;
inline int FlagGpu = -1; // additive: -mpi-gpu

// time.ParseDuration subset used by -mpi-inittimeout ("300ms", "1.5s", "2m", "1h2m3s"); false if malformed
inline bool ParseDuration(const std::string& s, int64_t* ns) {
  if (s == "0") { *ns = 0; return true; }
  size_t i = 0;
  double sign = 1, total = 0;
  if (i < s.size() && (s[i] == '+' || s[i] == '-')) { if (s[i] == '-') sign = -1; ++i; }
  if (i >= s.size()) return false;
  while (i < s.size()) {
    size_t j = i;
    while (j < s.size() && (isdigit((unsigned char)s[j]) || s[j] == '.')) ++j;
    if (j == i) return false;
    double v = atof(s.substr(i, j - i).c_str());
    size_t k = j;
    while (k < s.size() && !isdigit((unsigned char)s[k]) && s[k] != '.') ++k;
    std::string u = s.substr(j, k - j);
    double mul;
    if (u == "ns") mul = 1; else if (u == "us" || u == "\xC2\xB5s") mul = 1e3; else if (u == "ms") mul = 1e6;
    else if (u == "s") mul = 1e9; else if (u == "m") mul = 60e9; else if (u == "h") mul = 3600e9; else return false;
    total += v * mul;
    i = k;
  }
  *ns = (int64_t)(sign * total + (sign * total >= 0 ? 0.5 : -0.5));
  return true;
}

// flag.Parse(): consumes the -mpi-* flags (one or two dashes, "-f v" or "-f=v"), returns the rest.
inline std::vector<std::string> ParseFlags(int argc, char** argv) {
  std::vector<std::string> rest;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    std::string name = a;
    while (!name.empty() && name[0] == '-') name.erase(0, 1);
    const bool is_flag = a.size() > 1 && a[0] == '-' && name.rfind("mpi-", 0) == 0;
    if (!is_flag) { rest.push_back(a); continue; }
    std::string value;
    size_t eq = name.find('=');
    if (eq != std::string::npos) { value = name.substr(eq + 1); name = name.substr(0, eq); }
    else if (i + 1 < argc) value = argv[++i];
    else throw std::runtime_error("flag needs an argument: -" + name);
    if (name == "mpi-addr") FlagAddr = value;
    else if (name == "mpi-alladdr") {
      size_t start = 0;
      for (;;) {
        size_t c = value.find(',', start);
        FlagAllAddrs.push_back(value.substr(start, c == std::string::npos ? std::string::npos : c - start));
        if (c == std::string::npos) break;
        start = c + 1;
      }
    } else if (name == "mpi-inittimeout") { if (!ParseDuration(value, &FlagInitTimeout)) throw std::runtime_error("invalid duration " + value); }
    else if (name == "mpi-protocol") FlagProtocol = value;
    else if (name == "mpi-password") FlagPassword = value;
    else if (name == "mpi-gpu") FlagGpu = atoi(value.c_str());
    else rest.push_back(a);
  }
  return rest;
}

// ---- mpi.Interface (mpi.go:163-170) + the optional collective upgrade (mpi.go:69-71) ---------------
struct Buffer { // what `data interface{}` is lowered to
  void* ptr;
  size_t count;
  int dtype;
  int memkind;
};
struct Interface {
  virtual ~Interface() = default;
  virtual Error Init() = 0;
  virtual void Finalize() = 0;
  virtual int Rank() = 0;
  virtual int Size() = 0;
  virtual Error Send(Buffer data, int destination, int tag) = 0;
  virtual Error Receive(Buffer data, size_t* count_out, int source, int tag) = 0;
  // collectives: default "not supported", like an implementation that is no AllReducer
  virtual Error Bcast(Buffer, int) { Error e; e.code = B200MPI_ERR_UNSUPPORTED; e.msg = "registered implementation has no Bcast"; return e; }
  virtual Error Allreduce(Buffer, Buffer, Op) { Error e; e.code = B200MPI_ERR_UNSUPPORTED; e.msg = "registered implementation has no Allreduce"; return e; }
  virtual Error Allgather(Buffer, Buffer) { Error e; e.code = B200MPI_ERR_UNSUPPORTED; e.msg = "registered implementation has no Allgather"; return e; }
  virtual Error Barrier() { Error e; e.code = B200MPI_ERR_UNSUPPORTED; e.msg = "registered implementation has no Barrier"; return e; }
  virtual Error ReduceScatter(Buffer, Buffer, Op) { Error e; e.code = B200MPI_ERR_UNSUPPORTED; e.msg = "registered implementation has no ReduceScatter"; return e; }
  virtual Error Reduce(Buffer, Buffer, Op, int) { Error e; e.code = B200MPI_ERR_UNSUPPORTED; e.msg = "registered implementation has no Reduce"; return e; }
  virtual Error Alltoall(Buffer, Buffer) { Error e; e.code = B200MPI_ERR_UNSUPPORTED; e.msg = "registered implementation has no Alltoall"; return e; }
  // the Send/Wait pair the reference sketches and comments out (mpi.go:132-152)
  virtual Error Isend(Buffer, int, int) { Error e; e.code = B200MPI_ERR_UNSUPPORTED; e.msg = "registered implementation has no Isend"; return e; }
  virtual Error Wait(int, int) { Error e; e.code = B200MPI_ERR_UNSUPPORTED; e.msg = "registered implementation has no Wait"; return e; }
};

// The B200 implementation; plays the role of `Network` (network.go:25-39): fields win over flags.
struct Cuda : Interface {
  std::string Addr;
  std::vector<std::string> Addrs;
  int64_t Timeout = 0;
  std::string Password;
  int Gpu = -1;
  Error Init() override {
    if (Password.empty()) Password = FlagPassword; // useFlags, network.go:69-90
    if (Timeout == 0) Timeout = FlagInitTimeout;
    if (Addr.empty()) Addr = FlagAddr;
    if (Addrs.empty()) Addrs = FlagAllAddrs;
    if (Gpu < 0) Gpu = FlagGpu;
    std::string csv;
    for (size_t i = 0; i < Addrs.size(); ++i) csv += (i ? "," : "") + Addrs[i];
    return make_error(b200mpi_init(Addr.c_str(), csv.c_str(), Password.c_str(), Timeout, Gpu));
  }
  void Finalize() override { b200mpi_finalize(); }
  int Rank() override { return b200mpi_rank(); }
  int Size() override { return b200mpi_size(); }
  Error Send(Buffer d, int destination, int tag) override { return make_error(b200mpi_send(d.ptr, d.count, d.dtype, destination, tag, d.memkind)); }
  Error Receive(Buffer d, size_t* n, int source, int tag) override { return make_error(b200mpi_recv(d.ptr, d.count, n, d.dtype, source, tag, d.memkind)); }
  Error Bcast(Buffer d, int root) override { return make_error(b200mpi_bcast(d.ptr, d.count, d.dtype, root, d.memkind)); }
  Error Allreduce(Buffer s, Buffer r, Op op) override { return make_error(b200mpi_allreduce(s.ptr, r.ptr, s.count, s.dtype, op, s.memkind)); }
  Error Allgather(Buffer s, Buffer r) override { return make_error(b200mpi_allgather(s.ptr, r.ptr, s.count, s.dtype, s.memkind)); }
  Error Barrier() override { return make_error(b200mpi_barrier()); }
  Error ReduceScatter(Buffer s, Buffer r, Op op) override { return make_error(b200mpi_reduce_scatter(s.ptr, r.ptr, r.count, s.dtype, op, s.memkind)); }
  Error Reduce(Buffer s, Buffer r, Op op, int root) override { return make_error(b200mpi_reduce(s.ptr, r.ptr, s.count, s.dtype, op, root, s.memkind)); }
  Error Alltoall(Buffer s, Buffer r) override { return make_error(b200mpi_alltoall(s.ptr, r.ptr, s.count / (size_t)(Size() > 0 ? Size() : 1), s.dtype, s.memkind)); }
  Error Isend(Buffer d, int destination, int tag) override { return make_error(b200mpi_isend(d.ptr, d.count, d.dtype, destination, tag, d.memkind)); }
  Error Wait(int destination, int tag) override { return make_error(b200mpi_wait(destination, tag)); }
};

inline Interface*& mpier() { // var mpier Interface = &Network{}  (mpi.go:56)
  static Cuda default_impl;
  static Interface* cur = &default_impl;
  return cur;
}
inline void Register(Interface* impl) { // mpi.go:61-67
  static bool called = false;
  mpier() = impl;
  if (called) throw std::logic_error("register called more than once");
  called = true;
}

// ---- lowering `data` ---------------------------------------------------------------------------------
template <typename T> inline Buffer lower(const std::vector<T>& v) { return {const_cast<T*>(v.data()), v.size(), dtype_of<T>::value, B200MPI_HOST}; }
inline Buffer lower(const std::string& s) { return {const_cast<char*>(s.data()), s.size(), B200MPI_U8, B200MPI_HOST}; }
template <typename T> inline Buffer lower(const DeviceSlice<T>& d) { return {d.data(), d.size(), dtype_of<T>::value, B200MPI_DEVICE}; }

// ---- the package functions (mpi.go:96-159) -----------------------------------------------------------
inline Error Init() { return mpier()->Init(); }
inline void Finalize() { mpier()->Finalize(); }
inline int Rank() { return mpier()->Rank(); }
inline int Size() { return mpier()->Size(); }

template <typename D> inline Error Send(const D& data, int destination, int tag) { return mpier()->Send(lower(data), destination, tag); }

// Receive resizes *data to the sent length (gob does that for the reference, network.go:597).
template <typename T> inline Error Receive(std::vector<T>* data, int source, int tag) {
  size_t n = 0;
  Error e = mpier()->Receive(lower(*data), &n, source, tag);
  if (e.code == B200MPI_ERR_TRUNCATE) { // still posted: grow and take it
    data->resize(n);
    e = mpier()->Receive(lower(*data), &n, source, tag);
  }
  if (!e) data->resize(n);
  return e;
}
inline Error Receive(std::string* data, int source, int tag) {
  if (data->size() < 256) data->resize(256);
  size_t n = 0;
  Error e = mpier()->Receive(lower(*data), &n, source, tag);
  if (e.code == B200MPI_ERR_TRUNCATE) { data->resize(n); e = mpier()->Receive(lower(*data), &n, source, tag); }
  if (!e) data->resize(n);
  return e;
}
template <typename T> inline Error Receive(DeviceSlice<T>* data, size_t* count, int source, int tag) { return mpier()->Receive(lower(*data), count, source, tag); }
template <typename D> inline Error Recv(D* data, int source, int tag) { return Receive(data, source, tag); }

template <typename T> inline Error Bcast(std::vector<T>* data, int root) { return mpier()->Bcast(lower(*data), root); }
template <typename T> inline Error Bcast(DeviceSlice<T>* data, int root) { return mpier()->Bcast(lower(*data), root); }
template <typename T> inline Error Allreduce(const std::vector<T>& send, std::vector<T>* recv, Op op = SUM) {
  recv->resize(send.size());
  return mpier()->Allreduce(lower(send), lower(*recv), op);
}
template <typename T> inline Error Allreduce(const DeviceSlice<T>& send, DeviceSlice<T>* recv, Op op = SUM) { return mpier()->Allreduce(lower(send), lower(*recv), op); }
template <typename T> inline Error Allgather(const std::vector<T>& send, std::vector<T>* recv) {
  recv->resize(send.size() * (size_t)Size());
  return mpier()->Allgather(lower(send), lower(*recv));
}
template <typename T> inline Error Allgather(const DeviceSlice<T>& send, DeviceSlice<T>* recv) { return mpier()->Allgather(lower(send), lower(*recv)); }
inline Error Barrier() { return mpier()->Barrier(); }

// send holds Size() blocks; rank j receives the reduction of block j (recv is resized to one block)
template <typename T> inline Error ReduceScatter(const std::vector<T>& send, std::vector<T>* recv, Op op = SUM) {
  recv->resize(send.size() / (size_t)Size());
  return mpier()->ReduceScatter(lower(send), lower(*recv), op);
}
template <typename T> inline Error ReduceScatter(const DeviceSlice<T>& send, DeviceSlice<T>* recv, Op op = SUM) { return mpier()->ReduceScatter(lower(send), lower(*recv), op); }
template <typename T> inline Error Reduce(const std::vector<T>& send, std::vector<T>* recv, Op op, int root) {
  if (Rank() == root) recv->resize(send.size());
  return mpier()->Reduce(lower(send), lower(*recv), op, root);
}
template <typename T> inline Error Alltoall(const std::vector<T>& send, std::vector<T>* recv) {
  recv->resize(send.size());
  return mpier()->Alltoall(lower(send), lower(*recv));
}
template <typename D> inline Error Isend(const D& data, int destination, int tag) { return mpier()->Isend(lower(data), destination, tag); }
inline Error Wait(int destination, int tag) { return mpier()->Wait(destination, tag); }

} // namespace mpi

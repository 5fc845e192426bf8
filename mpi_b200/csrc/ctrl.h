// ctrl.h -- control plane of libb200mpi: the part of the reference's Network type that survives
// (/root/reference/network.go:53-351): flag-driven address list, rank = index in the sorted
// list, full-mesh TCP handshake with password + id check and an optional init timeout.
// The data plane (gob over those sockets, network.go:518-625) is NOT here: after bootstrap the
// sockets only carry small control blobs (device ids, allocation handles) and the optional
// slow-path barrier.  File descriptors (cuMem allocation handles, the mailbox memfd) travel over
// per-pair abstract Unix sockets with SCM_RIGHTS.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace b200 {

struct Ctrl {
  int rank = -1;
  int n = 0;
  std::string addr;               // -mpi-addr
  std::vector<std::string> addrs; // -mpi-alladdr, sorted
  std::string password;           // -mpi-password
  int64_t timeout_ns = 0;         // -mpi-inittimeout (0 = wait forever, as the reference)

  std::vector<int> dial_fd;   // [peer] socket I connected  (network.go:502 "dial")
  std::vector<int> listen_fd; // [peer] socket I accepted   (network.go:503 "listen")
  std::vector<int> uds_fd;    // [peer] unix socket for fd passing
  uint64_t nonce = 0;

  // Returns 0 or a negative b200mpi_error; err gets the message.
  int init(const char* addr, const char* alladdr_csv, const char* password, int64_t timeout_ns,
           std::string& err);
  void shutdown();

  // Every rank contributes `bytes`; all[r*bytes ...] = rank r's blob.  wait_ns < 0: the init
  // timeout rule (forever when -mpi-inittimeout is 0); otherwise give up after wait_ns.
  int allgather(const void* mine, size_t bytes, void* all, std::string& err, int64_t wait_ns = -1);
  int barrier(std::string& err);
  // out[r] = a descriptor in this process for rank r's fd (out[rank] = dup(myfd)).
  int alltoall_fd(int myfd, std::vector<int>& out, std::string& err);
  // root's fd delivered to everyone (root gets a dup).
  int bcast_fd(int root, int fd_in, int& fd_out, std::string& err);
};

// Pure helpers (unit-tested without sockets).
// Split a comma separated list the way flags.go:22-27 does (no trimming, empty items kept).
std::vector<std::string> split_addrs(const std::string& csv);
// network.go:94-109: sort, reject duplicates, rank = index of addr.  Returns rank or -1 (err set).
int assign_rank(std::vector<std::string>& addrs, const std::string& addr, std::string& err);
// "host:port" / ":port" -> host (may be empty) and port; false when malformed.
bool split_host_port(const std::string& addr, std::string& host, int& port);

} // namespace b200

// ctrl.cpp -- see ctrl.h.  POSIX sockets only; no CUDA in this file.
#include "ctrl.h"

#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <random>
#include <thread>

#include "../../include/b200mpi.h"

namespace b200 {

namespace {

using Clock = std::chrono::steady_clock;

struct Deadline {
  bool bounded;
  Clock::time_point at;
  explicit Deadline(int64_t ns) : bounded(ns > 0), at(Clock::now() + std::chrono::nanoseconds(ns)) {}
  bool expired() const { return bounded && Clock::now() >= at; }
  // poll() timeout in ms: -1 forever, else remaining (>=0)
  int poll_ms() const {
    if (!bounded) return -1;
    auto left = std::chrono::duration_cast<std::chrono::milliseconds>(at - Clock::now()).count();
    return left < 0 ? 0 : (int)std::min<int64_t>(left + 1, 1 << 30);
  }
};

constexpr uint32_t kMagic = 0xB2005A01u;
constexpr uint32_t kReject = 0xB200DEADu;

int write_full(int fd, const void* p, size_t n) {
  const char* c = (const char*)p;
  while (n) {
    ssize_t w = ::send(fd, c, n, MSG_NOSIGNAL);
    if (w < 0) {
      if (errno == EINTR) continue;
      return -1;
    }
    c += w;
    n -= (size_t)w;
  }
  return 0;
}

// Reads exactly n bytes; honours the deadline.  0 ok, -1 error/eof, -2 timeout.
int read_full(int fd, void* p, size_t n, const Deadline& dl) {
  char* c = (char*)p;
  while (n) {
    struct pollfd pf = {fd, POLLIN, 0};
    int pr = ::poll(&pf, 1, dl.poll_ms());
    if (pr < 0) {
      if (errno == EINTR) continue;
      return -1;
    }
    if (pr == 0) return -2;
    ssize_t r = ::recv(fd, c, n, 0);
    if (r < 0) {
      if (errno == EINTR || errno == EAGAIN) continue;
      return -1;
    }
    if (r == 0) return -1;
    c += r;
    n -= (size_t)r;
  }
  return 0;
}

// Hello frame: magic, id, password length, password bytes.  Plays the role of the reference's
// gob-encoded initialMessage{Password, Id} (network.go:198-201).
int send_hello(int fd, uint32_t magic, int id, const std::string& pw) {
  uint32_t hdr[3] = {magic, (uint32_t)id, (uint32_t)pw.size()};
  if (write_full(fd, hdr, sizeof hdr)) return -1;
  if (!pw.empty() && write_full(fd, pw.data(), pw.size())) return -1;
  return 0;
}

int recv_hello(int fd, uint32_t& magic, int& id, std::string& pw, const Deadline& dl) {
  uint32_t hdr[3];
  int r = read_full(fd, hdr, sizeof hdr, dl);
  if (r) return r;
  magic = hdr[0];
  id = (int)hdr[1];
  if (hdr[2] > (1u << 20)) return -1;
  pw.resize(hdr[2]);
  if (hdr[2]) {
    r = read_full(fd, &pw[0], hdr[2], dl);
    if (r) return r;
  }
  return 0;
}

void set_nodelay(int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
}

bool resolve(const std::string& host, int port, bool passive, sockaddr_in& out, std::string& err) {
  memset(&out, 0, sizeof out);
  out.sin_family = AF_INET;
  out.sin_port = htons((uint16_t)port);
  if (host.empty()) {
    out.sin_addr.s_addr = passive ? htonl(INADDR_ANY) : htonl(INADDR_LOOPBACK);
    return true;
  }
  if (inet_pton(AF_INET, host.c_str(), &out.sin_addr) == 1) return true;
  struct addrinfo hints = {}, *res = nullptr;
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  int g = getaddrinfo(host.c_str(), nullptr, &hints, &res);
  if (g != 0 || !res) {
    err = "cannot resolve host '" + host + "': " + gai_strerror(g);
    return false;
  }
  out.sin_addr = ((sockaddr_in*)res->ai_addr)->sin_addr;
  freeaddrinfo(res);
  return true;
}

std::string uds_name(uint64_t nonce, int rank) {
  char buf[64];
  snprintf(buf, sizeof buf, "b200mpi.%016llx.%d", (unsigned long long)nonce, rank);
  return buf;
}

socklen_t fill_abstract(sockaddr_un& sa, const std::string& name) {
  memset(&sa, 0, sizeof sa);
  sa.sun_family = AF_UNIX;
  sa.sun_path[0] = '\0';
  memcpy(sa.sun_path + 1, name.data(), name.size());
  return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
}

int send_fd(int sock, int fd) {
  struct msghdr msg = {};
  char cbuf[CMSG_SPACE(sizeof(int))];
  memset(cbuf, 0, sizeof cbuf);
  char tag = 'F';
  struct iovec io = {&tag, 1};
  msg.msg_iov = &io;
  msg.msg_iovlen = 1;
  msg.msg_control = cbuf;
  msg.msg_controllen = sizeof cbuf;
  struct cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  for (;;) {
    ssize_t w = sendmsg(sock, &msg, MSG_NOSIGNAL);
    if (w < 0 && errno == EINTR) continue;
    return w == 1 ? 0 : -1;
  }
}

int recv_fd(int sock, int& fd) {
  struct msghdr msg = {};
  char cbuf[CMSG_SPACE(sizeof(int))];
  char tag = 0;
  struct iovec io = {&tag, 1};
  msg.msg_iov = &io;
  msg.msg_iovlen = 1;
  msg.msg_control = cbuf;
  msg.msg_controllen = sizeof cbuf;
  for (;;) {
    ssize_t r = recvmsg(sock, &msg, 0);
    if (r < 0 && errno == EINTR) continue;
    if (r != 1 || tag != 'F') return -1;
    break;
  }
  struct cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return -1;
  memcpy(&fd, CMSG_DATA(c), sizeof(int));
  return 0;
}

} // namespace

std::vector<std::string> split_addrs(const std::string& csv) {
  std::vector<std::string> out;
  if (csv.empty()) return out;
  size_t start = 0;
  for (;;) {
    size_t comma = csv.find(',', start);
    if (comma == std::string::npos) {
      out.push_back(csv.substr(start));
      break;
    }
    out.push_back(csv.substr(start, comma - start));
    start = comma + 1;
  }
  return out;
}

int assign_rank(std::vector<std::string>& addrs, const std::string& addr, std::string& err) {
  std::sort(addrs.begin(), addrs.end());
  for (size_t i = 0; i + 1 < addrs.size(); ++i) {
    if (addrs[i] == addrs[i + 1]) {
      err = "network addresses not unique: '" + addrs[i] + "' appears twice";
      return -1;
    }
  }
  auto it = std::lower_bound(addrs.begin(), addrs.end(), addr);
  if (it == addrs.end() || *it != addr) {
    err = "mpi init: local address '" + addr + "' not in global list";
    return -1;
  }
  return (int)(it - addrs.begin());
}

bool split_host_port(const std::string& addr, std::string& host, int& port) {
  size_t colon = addr.rfind(':');
  if (colon == std::string::npos) return false;
  host = addr.substr(0, colon);
  std::string p = addr.substr(colon + 1);
  if (p.empty() || p.size() > 5) return false;
  port = 0;
  for (char ch : p) {
    if (ch < '0' || ch > '9') return false;
    port = port * 10 + (ch - '0');
  }
  return port > 0 && port < 65536;
}

int Ctrl::init(const char* addr_c, const char* csv_c, const char* pw_c, int64_t tmo,
               std::string& err) {
  addr = addr_c ? addr_c : "";
  password = pw_c ? pw_c : "";
  timeout_ns = tmo;
  addrs = split_addrs(csv_c ? csv_c : "");
  if (addrs.empty()) { // network.go:55-58: no list => one node on ":5000"
    addr = ":5000";
    addrs = {":5000"};
  }
  if (addrs.size() > B200MPI_MAX_RANKS) {
    err = "at most 8 ranks (one box) are supported, got " + std::to_string(addrs.size());
    return B200MPI_ERR_BOOTSTRAP;
  }
  int r = assign_rank(addrs, addr, err);
  if (r < 0) return B200MPI_ERR_BOOTSTRAP;
  rank = r;
  n = (int)addrs.size();
  dial_fd.assign(n, -1);
  listen_fd.assign(n, -1);
  uds_fd.assign(n, -1);
  std::random_device rd;
  nonce = ((uint64_t)rd() << 32) ^ (uint64_t)rd() ^ ((uint64_t)getpid() << 16);
  if (n == 1) return 0;

  Deadline dl(timeout_ns);
  std::string host;
  int port = 0;
  if (!split_host_port(addr, host, port)) {
    err = "error listening: malformed address '" + addr + "'";
    return B200MPI_ERR_BOOTSTRAP;
  }
  sockaddr_in sa;
  if (!resolve(host, port, true, sa, err)) return B200MPI_ERR_BOOTSTRAP;
  int lst = ::socket(AF_INET, SOCK_STREAM, 0);
  int one = 1;
  setsockopt(lst, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
  if (::bind(lst, (sockaddr*)&sa, sizeof sa) != 0 || ::listen(lst, 64) != 0) {
    err = std::string("error listening: ") + strerror(errno) + " (" + addr + ")";
    ::close(lst);
    return B200MPI_ERR_BOOTSTRAP;
  }

  // The two halves run concurrently, like the two goroutines in network.go:137-146.
  int listen_rc = 0, dial_rc = 0;
  std::string listen_err, dial_err;

  std::thread listener([&] { // network.go:163-263
    for (int got = 0; got < n - 1; ++got) {
      struct pollfd pf = {lst, POLLIN, 0};
      int pr;
      do { pr = ::poll(&pf, 1, dl.poll_ms()); } while (pr < 0 && errno == EINTR);
      if (pr == 0) {
        listen_err = "listener timed out";
        listen_rc = B200MPI_ERR_TIMEOUT;
        return;
      }
      int c = ::accept(lst, nullptr, nullptr);
      if (c < 0) {
        listen_err = std::string("error accepting: ") + strerror(errno);
        listen_rc = B200MPI_ERR_BOOTSTRAP;
        return;
      }
      set_nodelay(c);
      uint32_t magic;
      int id;
      std::string pw;
      int hr = recv_hello(c, magic, id, pw, dl);
      if (hr || magic != kMagic) {
        ::close(c);
        listen_err = hr == -2 ? "listener timed out" : "bad handshake frame";
        listen_rc = hr == -2 ? B200MPI_ERR_TIMEOUT : B200MPI_ERR_BOOTSTRAP;
        return;
      }
      if (pw != password) { // network.go:344-346
        send_hello(c, kReject, rank, "");
        ::close(c);
        listen_err = "bad password";
        listen_rc = B200MPI_ERR_PASSWORD;
        return;
      }
      if (id < 0 || id >= n || id == rank || listen_fd[id] != -1) { // network.go:347-349
        send_hello(c, kReject, rank, "");
        ::close(c);
        listen_err = "bad id: " + std::to_string(id);
        listen_rc = B200MPI_ERR_BOOTSTRAP;
        return;
      }
      listen_fd[id] = c;
      if (send_hello(c, kMagic, rank, password)) {
        listen_err = "handshake reply failed";
        listen_rc = B200MPI_ERR_BOOTSTRAP;
        return;
      }
    }
  });

  std::thread dialer([&] { // network.go:265-339
    for (int peer = 0; peer < n; ++peer) {
      if (peer == rank) continue;
      std::string ph;
      int pp = 0;
      sockaddr_in psa;
      if (!split_host_port(addrs[peer], ph, pp) || !resolve(ph, pp, false, psa, dial_err)) {
        if (dial_err.empty()) dial_err = "malformed peer address '" + addrs[peer] + "'";
        dial_rc = B200MPI_ERR_BOOTSTRAP;
        return;
      }
      int c = -1;
      for (;;) { // 100 ms retry tick, network.go:298
        c = ::socket(AF_INET, SOCK_STREAM, 0);
        if (::connect(c, (sockaddr*)&psa, sizeof psa) == 0) break;
        int e = errno;
        ::close(c);
        c = -1;
        if (dl.expired()) {
          dial_err = std::string("dial ") + addrs[peer] + ": " + strerror(e) + " (timed out)";
          dial_rc = B200MPI_ERR_TIMEOUT;
          return;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
      }
      set_nodelay(c);
      uint32_t magic;
      int id;
      std::string pw;
      if (send_hello(c, kMagic, rank, password)) {
        ::close(c);
        dial_err = "handshake send failed";
        dial_rc = B200MPI_ERR_BOOTSTRAP;
        return;
      }
      int hr = recv_hello(c, magic, id, pw, dl);
      if (hr == 0 && magic == kReject) {
        ::close(c);
        dial_err = "peer " + addrs[peer] + " rejected the handshake (bad password or id)";
        dial_rc = B200MPI_ERR_PASSWORD;
        return;
      }
      if (hr || magic != kMagic) {
        ::close(c);
        dial_err = hr == -2 ? "dial handshake timed out" : "peer closed during handshake (bad password?)";
        dial_rc = hr == -2 ? B200MPI_ERR_TIMEOUT : B200MPI_ERR_PASSWORD;
        return;
      }
      if (pw != password) {
        ::close(c);
        dial_err = "bad password";
        dial_rc = B200MPI_ERR_PASSWORD;
        return;
      }
      if (id != peer) {
        ::close(c);
        dial_err = "bad id: " + std::to_string(id);
        dial_rc = B200MPI_ERR_BOOTSTRAP;
        return;
      }
      dial_fd[peer] = c;
    }
  });

  listener.join();
  dialer.join();
  ::close(lst);
  if (listen_rc) { // listen errors first, as network.go:150-156
    err = listen_err;
    shutdown();
    return listen_rc;
  }
  if (dial_rc) {
    err = dial_err;
    shutdown();
    return dial_rc;
  }

  // Agree on a nonce and a secret (rank 0's), both carried by the password-checked TCP mesh, and
  // open the per-pair unix sockets used for fd passing.  The socket names are visible in
  // /proc/net/unix, so a connection only counts after (a) SO_PEERCRED shows this user's uid and
  // (b) the peer presented the secret -- before that no allocation handle is sent to it.
  struct Agree { uint64_t nonce, secret[2]; } mine_a, all_a[B200MPI_MAX_RANKS];
  {
    std::random_device rd2;
    mine_a.nonce = nonce;
    mine_a.secret[0] = ((uint64_t)rd2() << 32) ^ rd2();
    mine_a.secret[1] = ((uint64_t)rd2() << 32) ^ rd2();
  }
  int rc = allgather(&mine_a, sizeof mine_a, all_a, err);
  if (rc) return rc;
  nonce = all_a[0].nonce;
  const uint64_t secret[2] = {all_a[0].secret[0], all_a[0].secret[1]};
  int ul = ::socket(AF_UNIX, SOCK_STREAM, 0);
  sockaddr_un usa;
  socklen_t ulen = fill_abstract(usa, uds_name(nonce, rank));
  if (::bind(ul, (sockaddr*)&usa, ulen) != 0 || ::listen(ul, 16) != 0) {
    err = std::string("unix socket bind failed: ") + strerror(errno);
    ::close(ul);
    return B200MPI_ERR_BOOTSTRAP;
  }
  rc = barrier(err); // every listener exists
  if (rc) {
    ::close(ul);
    return rc;
  }
  struct UdsHello { int32_t who; int32_t pad; uint64_t secret[2]; };
  for (int peer = 0; peer < rank; ++peer) { // higher rank connects to lower rank
    int c = ::socket(AF_UNIX, SOCK_STREAM, 0);
    sockaddr_un psa;
    socklen_t plen = fill_abstract(psa, uds_name(nonce, peer));
    if (::connect(c, (sockaddr*)&psa, plen) != 0) {
      err = std::string("unix socket connect failed: ") + strerror(errno);
      ::close(c);
      ::close(ul);
      return B200MPI_ERR_BOOTSTRAP;
    }
    UdsHello h = {rank, 0, {secret[0], secret[1]}};
    write_full(c, &h, sizeof h);
    uds_fd[peer] = c;
  }
  Deadline d2(timeout_ns > 0 ? timeout_ns : 60000000000ll);
  for (int got = rank + 1; got < n;) {
    struct pollfd pf = {ul, POLLIN, 0};
    int pr;
    do { pr = ::poll(&pf, 1, d2.poll_ms()); } while (pr < 0 && errno == EINTR);
    int c = pr > 0 ? ::accept(ul, nullptr, nullptr) : -1;
    if (c < 0) {
      err = "unix socket accept failed or timed out";
      ::close(ul);
      return B200MPI_ERR_BOOTSTRAP;
    }
    struct ucred cred = {};
    socklen_t clen = sizeof cred;
    UdsHello h = {};
    const bool cred_ok = getsockopt(c, SOL_SOCKET, SO_PEERCRED, &cred, &clen) == 0 && cred.uid == geteuid();
    if (!cred_ok || read_full(c, &h, sizeof h, d2) || h.secret[0] != secret[0] || h.secret[1] != secret[1] ||
        h.who <= rank || h.who >= n || uds_fd[h.who] != -1) {
      ::close(c); // a stranger (or a duplicate): not a rank of this world, keep waiting for the real one
      continue;
    }
    uds_fd[h.who] = c;
    ++got;
  }
  ::close(ul);
  return 0;
}

void Ctrl::shutdown() { // network.go:354-369
  for (auto* v : {&dial_fd, &listen_fd, &uds_fd}) {
    for (int& fd : *v) {
      if (fd >= 0) ::close(fd);
      fd = -1;
    }
  }
}

int Ctrl::allgather(const void* mine, size_t bytes, void* all, std::string& err, int64_t wait_ns) {
  char* out = (char*)all;
  memcpy(out + (size_t)rank * bytes, mine, bytes);
  if (n == 1) return 0;
  // Blobs are small (<< socket buffer), so send-all-then-receive-all cannot deadlock.
  for (int p = 0; p < n; ++p) {
    if (p == rank) continue;
    if (write_full(dial_fd[p], mine, bytes)) {
      err = "control plane: send to rank " + std::to_string(p) + " failed";
      return B200MPI_ERR_PEER;
    }
  }
  Deadline dl(wait_ns >= 0 ? std::max<int64_t>(wait_ns, 1) : (timeout_ns > 0 ? std::max<int64_t>(timeout_ns, 60000000000ll) : 0));
  for (int p = 0; p < n; ++p) {
    if (p == rank) continue;
    int r = read_full(listen_fd[p], out + (size_t)p * bytes, bytes, dl);
    if (r) {
      err = "control plane: receive from rank " + std::to_string(p) + (r == -2 ? " timed out" : " failed");
      return r == -2 ? B200MPI_ERR_TIMEOUT : B200MPI_ERR_PEER;
    }
  }
  return 0;
}

int Ctrl::barrier(std::string& err) {
  char mine = 'b';
  char all[B200MPI_MAX_RANKS];
  return allgather(&mine, 1, all, err);
}

int Ctrl::alltoall_fd(int myfd, std::vector<int>& out, std::string& err) {
  out.assign(n, -1);
  out[rank] = ::dup(myfd);
  for (int p = 0; p < n; ++p) {
    if (p == rank) continue;
    if (send_fd(uds_fd[p], myfd)) {
      err = "fd send to rank " + std::to_string(p) + " failed: " + strerror(errno);
      return B200MPI_ERR_PEER;
    }
  }
  for (int p = 0; p < n; ++p) {
    if (p == rank) continue;
    if (recv_fd(uds_fd[p], out[p])) {
      err = "fd receive from rank " + std::to_string(p) + " failed";
      return B200MPI_ERR_PEER;
    }
  }
  return 0;
}

int Ctrl::bcast_fd(int root, int fd_in, int& fd_out, std::string& err) {
  if (rank == root) {
    for (int p = 0; p < n; ++p) {
      if (p == rank) continue;
      if (send_fd(uds_fd[p], fd_in)) {
        err = "fd send to rank " + std::to_string(p) + " failed: " + strerror(errno);
        return B200MPI_ERR_PEER;
      }
    }
    fd_out = ::dup(fd_in);
    return 0;
  }
  if (recv_fd(uds_fd[root], fd_out)) {
    err = "fd receive from root failed";
    return B200MPI_ERR_PEER;
  }
  return 0;
}

} // namespace b200

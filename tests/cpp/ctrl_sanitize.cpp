// Host-side sanitizer harness for the control plane (mpi_b200/csrc/ctrl.cpp): one rank per process,
// built with -fsanitize=thread or address by tests/test_sanitizers.py.  Exercises the concurrent
// listen/dial handshake (the two goroutines of network.go:137-146), the blob allgather, the barrier
// and fd passing over the per-pair unix sockets.
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../mpi_b200/csrc/ctrl.h"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  b200::Ctrl c;
  std::string err;
  int rc = c.init(argv[1], argv[2], argc > 3 ? argv[3] : "", 20ll * 1000000000ll, err);
  if (rc) {
    fprintf(stderr, "init failed: %d %s\n", rc, err.c_str());
    return 1;
  }
  std::vector<int> all(c.n);
  int mine = 100 + c.rank;
  if (c.allgather(&mine, sizeof mine, all.data(), err)) return 1;
  for (int r = 0; r < c.n; ++r)
    if (all[r] != 100 + r) return 1;
  for (int i = 0; i < 20; ++i)
    if (c.barrier(err)) return 1;
  if (c.n > 1) {
    // every rank shares a memfd holding its rank; everyone reads everyone's
    int fd = memfd_create("ctrl-sanitize", 0);
    if (fd < 0 || ftruncate(fd, 4096) != 0) return 1;
    if (pwrite(fd, &c.rank, sizeof c.rank, 0) != (ssize_t)sizeof c.rank) return 1;
    std::vector<int> fds;
    if (c.alltoall_fd(fd, fds, err)) return 1;
    for (int r = 0; r < c.n; ++r) {
      int v = -1;
      if (pread(fds[r], &v, sizeof v, 0) != (ssize_t)sizeof v || v != r) return 1;
      close(fds[r]);
    }
    int got = -1;
    if (c.bcast_fd(0, fd, got, err)) return 1;
    int v = -1;
    if (pread(got, &v, sizeof v, 0) != (ssize_t)sizeof v || v != 0) return 1;
    close(got);
    close(fd);
  }
  c.shutdown();
  printf("rank %d of %d ok\n", c.rank, c.n);
  return 0;
}

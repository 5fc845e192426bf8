// helloworld -- the reference's first example restated on the C++ facade
// (/root/reference/examples/helloworld/helloworld.go:33-82): every rank concurrently sends a
// string to every rank, itself included, and receives one from every rank, all with tag 0.
//
//   gompirun N helloworld        (or N shells: helloworld -mpi-addr :6000 -mpi-alladdr :6000,:6001 ...)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../mpi_b200/cpp/mpi.hpp"

int main(int argc, char** argv) {
  mpi::ParseFlags(argc, argv);
  if (mpi::Error err = mpi::Init()) {
    fprintf(stderr, "%s\n", err.String().c_str());
    return 1;
  }
  const int rank = mpi::Rank();
  if (rank == -1) {
    fprintf(stderr, "Incorrect initialization\n");
    return 1;
  }
  const int size = mpi::Size();
  printf("Hello world, I'm node %d in a land with %d nodes\n", rank, size);
  std::vector<std::thread> pool;
  int failures = 0;
  for (int i = 0; i < size; ++i) {
    pool.emplace_back([=, &failures] {
      std::string str = "\"Hello node " + std::to_string(i) + ", I'm node " + std::to_string(rank) + "\"";
      if (i == rank) str = "\"I'm just node " + std::to_string(rank) + " talking to myself\"";
      if (mpi::Error err = mpi::Send(str, i, 0)) {
        fprintf(stderr, "send: %s\n", err.String().c_str());
        ++failures;
      }
    });
  }
  for (int i = 0; i < size; ++i) {
    pool.emplace_back([=, &failures] {
      std::string str;
      if (mpi::Error err = mpi::Receive(&str, i, 0)) {
        fprintf(stderr, "receive: %s\n", err.String().c_str());
        ++failures;
        return;
      }
      printf("I, node %d, received a message: %s\n", rank, str.c_str());
    });
  }
  for (auto& t : pool) t.join();
  mpi::Finalize();
  return failures ? 1 : 0;
}

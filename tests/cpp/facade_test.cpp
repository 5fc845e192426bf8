// CPU-side checks of the C++ facade (mpi_b200/cpp/mpi.hpp): flag parsing with Go's rules
// (flags.go), duration syntax, Register-once (mpi.go:61-67), Rank/Size before Init
// (mpi.go:110-118) and the error surface without a device.
#include <cassert>
#include <cstdio>
#include <stdexcept>

#include "../../mpi_b200/cpp/mpi.hpp"

struct Fake : mpi::Interface {
  mpi::Error Init() override { return {}; }
  void Finalize() override {}
  int Rank() override { return 7; }
  int Size() override { return 9; }
  mpi::Error Send(mpi::Buffer, int, int) override { return {}; }
  mpi::Error Receive(mpi::Buffer, size_t* n, int, int) override { *n = 0; return {}; }
};

int main() {
  int64_t ns = -1;
  assert(mpi::ParseDuration("0", &ns) && ns == 0);
  assert(mpi::ParseDuration("300ms", &ns) && ns == 300000000ll);
  assert(mpi::ParseDuration("1h2m3.5s", &ns) && ns == 3723500000000ll);
  assert(mpi::ParseDuration("1.5s", &ns) && ns == 1500000000ll);
  assert(mpi::ParseDuration("-2us", &ns) && ns == -2000);
  assert(!mpi::ParseDuration("", &ns) && !mpi::ParseDuration("5", &ns) && !mpi::ParseDuration("1x", &ns));

  const char* argv[] = {"prog", "-x", "1", "-mpi-addr", ":6001", "--mpi-alladdr=:6000,:6001", "-mpi-alladdr", ":6002",
                        "-mpi-inittimeout=2s", "-mpi-password", "pw", "-mpi-gpu", "3", "tail"};
  auto rest = mpi::ParseFlags(14, const_cast<char**>(argv));
  assert(mpi::FlagAddr == ":6001");
  assert(mpi::FlagAllAddrs.size() == 3 && mpi::FlagAllAddrs[2] == ":6002"); // AddrsFlag appends
  assert(mpi::FlagInitTimeout == 2000000000ll && mpi::FlagPassword == "pw" && mpi::FlagGpu == 3 && mpi::FlagProtocol == "tcp");
  assert(rest.size() == 3 && rest[0] == "-x" && rest[1] == "1" && rest[2] == "tail");

  assert(mpi::Rank() == -1 && mpi::Size() == 0); // before Init
  std::vector<float> x(4, 1.f), y;
  mpi::Error e = mpi::Allreduce(x, &y);
  assert(e && e.code == B200MPI_ERR_NOT_INIT && !e.String().empty());

  Fake fake;
  mpi::Register(&fake);
  assert(mpi::Rank() == 7 && mpi::Size() == 9);
  assert(mpi::Allreduce(x, &y).code == B200MPI_ERR_UNSUPPORTED); // no collective upgrade in Fake
  bool threw = false;
  try { mpi::Register(&fake); } catch (const std::logic_error&) { threw = true; }
  assert(threw); // "register called more than once"
  printf("facade ok\n");
  return 0;
}

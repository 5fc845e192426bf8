// bounce -- the reference's ping-pong example restated on the C++ facade
// (/root/reference/examples/bounce/bounce.go:33-152): even ranks send to rank+1, odd ranks send
// the message back; []byte then []float64 over the size ladder, equality checked on the even rank
// after every round trip, mean trip time printed in microseconds.  Added: the 1 MiB float64 point of
// BASELINE.json configs[1] and a device-resident pass (DeviceSlice) next to the host-slice pass.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../mpi_b200/cpp/mpi.hpp"

static const size_t kMsgLengths[] = {0, 1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 1 << 20}; // bounce.go:33 + 1 MiB
static const int kRepeats = 10;                                                                       // bounce.go:35

int main(int argc, char** argv) {
  mpi::ParseFlags(argc, argv);
  if (mpi::Error err = mpi::Init()) {
    fprintf(stderr, "error initializing: %s\n", err.String().c_str());
    return 1;
  }
  const int rank = mpi::Rank(), size = mpi::Size();
  if (rank < 0) { fprintf(stderr, "Incorrect initialization\n"); return 1; }
  if (size % 2 != 0) { fprintf(stderr, "Must have an even number of nodes for this example\n"); mpi::Finalize(); return 1; }
  const bool even = rank % 2 == 0;
  if (rank == 0) printf("Number of nodes =  %d\n", size);
  const size_t maxsize = 10000000;
  std::mt19937_64 rng(12345 + rank);
  std::vector<uint8_t> message(maxsize);
  for (size_t i = 0; i + 8 <= maxsize; i += 8) { uint64_t v = rng() >> 1; memcpy(&message[i], &v, 8); }
  std::vector<double> messageFloats(maxsize / 8);
  std::uniform_real_distribution<double> uni(0.0, 1.0);
  for (auto& f : messageFloats) f = uni(rng);
  mpi::DeviceSlice<double> dmsg(maxsize / 8), drcv(maxsize / 8);
  dmsg.CopyFromHost(messageFloats);

  const int nl = sizeof kMsgLengths / sizeof kMsgLengths[0];
  std::vector<long long> times(nl), timesF(nl), timesD(nl);
  int bad = 0;
  using clk = std::chrono::steady_clock;
  for (int i = 0; i < nl; ++i) {
    const size_t l = kMsgLengths[i];
    for (int j = 0; j < kRepeats; ++j) {
      std::vector<uint8_t> msg(message.begin(), message.begin() + l), rcv(l);
      auto start = clk::now();
      if (even) { mpi::Send(msg, rank + 1, 0); mpi::Receive(&rcv, rank + 1, 0); }
      else { mpi::Receive(&rcv, rank - 1, 0); mpi::Send(rcv, rank - 1, 0); }
      times[i] += std::chrono::duration_cast<std::chrono::nanoseconds>(clk::now() - start).count();
      if (even && msg != rcv) { fprintf(stderr, "message not the same\n"); ++bad; }

      std::vector<double> msgF(messageFloats.begin(), messageFloats.begin() + l / 8), rcvF(l / 8);
      start = clk::now();
      if (even) { mpi::Send(msgF, rank + 1, 0); mpi::Receive(&rcvF, rank + 1, 0); }
      else { mpi::Receive(&rcvF, rank - 1, 0); mpi::Send(rcvF, rank - 1, 0); }
      timesF[i] += std::chrono::duration_cast<std::chrono::nanoseconds>(clk::now() - start).count();
      if (even && msgF != rcvF) { fprintf(stderr, "message not the same\n"); ++bad; }

      // device-resident: the zero-copy path (payload crosses NVLink once per direction)
      mpi::DeviceSlice<double> s = dmsg.Sub(0, l / 8), r = drcv.Sub(0, l / 8);
      size_t got = 0;
      start = clk::now();
      if (even) { mpi::Send(s, rank + 1, 1); mpi::Receive(&r, &got, rank + 1, 1); }
      else { mpi::Receive(&r, &got, rank - 1, 1); mpi::Send(r, rank - 1, 1); }
      timesD[i] += std::chrono::duration_cast<std::chrono::nanoseconds>(clk::now() - start).count();
      if (even) {
        std::vector<double> back;
        r.CopyToHost(&back);
        if (got != l / 8 || back != msgF) { fprintf(stderr, "device message not the same\n"); ++bad; }
      }
    }
  }
  if (even) {
    auto show = [&](const char* what, const std::vector<long long>& t) {
      printf("Average %s trip time in us between node %d and %d: [", what, rank, rank + 1);
      for (int i = 0; i < nl; ++i) printf("%lld%s", t[i] / 1000 / kRepeats, i + 1 < nl ? " " : "]\n");
    };
    show("byte", times);
    show("float64", timesF);
    show("float64 (device-resident)", timesD);
  }
  mpi::Finalize();
  return bad ? 1 : 0;
}

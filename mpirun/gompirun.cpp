// gompirun -- local launcher re-targeted to one rank per local GPU.
//
//   gompirun N program [args...]        N = "auto" or 0: one rank per visible GPU
//
// Behaviour of the reference launcher kept (/root/reference/mpirun/gompirun/gompirun.go:28-93):
// N children of `program`, ports ":6000"+i (gompirun.go:45-51), each child gets the user's
// arguments followed by `-mpi-addr <own> -mpi-alladdr <list>` (gompirun.go:77-83), stdio is
// inherited, the launcher waits for all.  Added: `-mpi-gpu <rank % ngpus>` per child, N capped
// sanity (the transport serves the 8 GPUs of one box), non-zero exit when a child fails.
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static int gpu_count() {
  if (const char* vis = getenv("CUDA_VISIBLE_DEVICES")) {
    if (*vis) {
      int n = 1;
      for (const char* p = vis; *p; ++p) n += *p == ',';
      return n;
    }
  }
  FILE* f = popen("nvidia-smi -L 2>/dev/null", "r");
  if (!f) return 0;
  int n = 0;
  char line[512];
  while (fgets(line, sizeof line, f)) n += strncmp(line, "GPU ", 4) == 0;
  pclose(f);
  return n;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "less than two arguments, must have at least number of nodes and executable\n");
    return 2;
  }
  const int ngpu = gpu_count();
  int n;
  if (strcmp(argv[1], "auto") == 0) n = ngpu > 0 ? ngpu : 1;
  else {
    char* end = nullptr;
    n = (int)strtol(argv[1], &end, 10);
    if (*end != '\0') { fprintf(stderr, "error parsing nNodes: %s\n", argv[1]); return 2; }
    if (n < 1) { fprintf(stderr, "number of nodes must be positive\n"); return 2; }
  }
  if (n > 8) { fprintf(stderr, "at most 8 ranks (one box)\n"); return 2; }
  int base = 6000;
  if (const char* b = getenv("GOMPIRUN_BASE_PORT")) base = atoi(b);
  std::vector<std::string> ports;
  std::string list;
  for (int i = 0; i < n; ++i) {
    ports.push_back(":" + std::to_string(base + i));
    list += (i ? "," : "") + ports.back();
  }
  std::vector<pid_t> kids;
  for (int i = 0; i < n; ++i) {
    std::vector<std::string> a;
    a.push_back(argv[2]);
    for (int k = 3; k < argc; ++k) a.push_back(argv[k]);
    a.insert(a.end(), {"-mpi-addr", ports[i], "-mpi-alladdr", list});
    if (ngpu > 0) a.insert(a.end(), {"-mpi-gpu", std::to_string(i % ngpu)});
    pid_t pid = fork();
    if (pid == 0) {
      std::vector<char*> cargv;
      for (auto& s : a) cargv.push_back(const_cast<char*>(s.c_str()));
      cargv.push_back(nullptr);
      execvp(cargv[0], cargv.data());
      perror("gompirun: exec");
      _exit(127);
    }
    kids.push_back(pid);
  }
  int rc = 0;
  for (pid_t k : kids) {
    int st = 0;
    waitpid(k, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 1;
  }
  return rc;
}

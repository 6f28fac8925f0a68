// cs_main.cc -- launcher: the reference's own example server and client (their sources are compiled in place
// with -Dmain=server_main / -Dmain=client_main) in ONE process: the loopback verbs of this offline build and the
// in-process wire of libb200rdma.so both need the two ends in one address space.  The reference's CI runs them
// as separate MPI ranks over a real fabric (examples/cpp/test/test.sh).
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include <thread>

int server_main(int argc, char** argv);
int client_main(int argc, char** argv);

int main(int argc, char** argv) {
  std::thread srv([&] { server_main(argc, argv); });
  srv.detach();
  sleep(1);
  const int rc = client_main(argc, argv);
  printf("client finished rc=%d\n", rc);
  fflush(stdout);
  _exit(rc);
}

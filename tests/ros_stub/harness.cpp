// harness.cpp — runs one shim node (its main() renamed to node_main by -Dmain=node_main) against a scenario file.
// TEST INFRASTRUCTURE ONLY (see include/ros_stub/core.h).   usage: <node> scenario.bin output.bin
#include <cstdio>

#include "ros_stub/core.h"

int node_main(int argc, char** argv);

#undef main
int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s scenario.bin output.bin\n", argv[0]);
    return 2;
  }
  ros::stub::Master& m = ros::stub::Master::get();
  if (!m.load(argv[1])) {
    std::fprintf(stderr, "cannot read scenario %s\n", argv[1]);
    return 2;
  }
  const int rc = node_main(argc, argv);
  if (!m.dump(argv[2])) {
    std::fprintf(stderr, "cannot write %s\n", argv[2]);
    return 2;
  }
  return rc;
}

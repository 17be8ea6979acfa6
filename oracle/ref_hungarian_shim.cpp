// Shim that exposes the reference's verbatim Munkres solver through a C symbol.
// Compiled TOGETHER WITH /root/reference/skeleton_3d/src/Hungarian.cpp (read where it lies,
// never copied) into oracle/_ref/libref_hungarian.so by oracle/Makefile. Test infrastructure.
#include <Hungarian.h>  // -I /root/reference/skeleton_3d/include

extern "C" void ref_hungarian_assignmentoptimal(int* assignment, double* cost, double* dist, int n_rows, int n_cols) {
  HungarianAlgorithm::assignmentoptimal(assignment, cost, dist, n_rows, n_cols);  // Hungarian.h:24, HUN:60
}

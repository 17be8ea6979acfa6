// host_setup.h — K0: one-time camera tables (host, FP64), see host_setup.cpp.
#pragma once
#include <vector>

#include "common.h"

namespace ses3d {

struct HostTables {
  int n_cams = 0;
  std::vector<CamF> camf;
  std::vector<CamD> camd;
  std::vector<float> F;     // [C(C-1)/2][9]
  std::vector<int> f_row;   // [C]
  SkeletonModel model;
};

// Returns false on invalid input (fewer than 2 cameras, singular extrinsics).
bool build_host_tables(int n_cams, const ses3d_camera* cams, const ses3d_params& prm, HostTables* out);

}  // namespace ses3d

// ses3d_ros/convert.h — glue between the ROS 1 message types of the reference (person_msgs, sensor_msgs, tf2,
// visualization_msgs) and the POD layouts of libses3d (include/ses3d.h). Shared by the three shim nodes.
//
// Nothing here computes geometry: the functions copy fields. What each one replaces in the reference is cited inline
// (S3D = skeleton_3d/src/skeleton_3d_triang_mult_node.cpp, REP = pose_reprojection/src/skeleton_reproj_mult_node.cpp,
// PRI = pose_prior/src/pose_prior_mult_node.cpp).
#pragma once
#include <geometry_msgs/TransformStamped.h>
#include <person_msgs/Person2DList.h>
#include <person_msgs/PersonCovList.h>
#include <sensor_msgs/CameraInfo.h>
#include <ses3d.h>
#include <std_msgs/ColorRGBA.h>
#include <visualization_msgs/MarkerArray.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

namespace ses3d_ros {

// One camera of the rig from what getTransforms() / getIntrinsics() deliver (S3D:161-228, REP:77-137):
// lookupTransform(target = camera optical frame, source = base) and the CameraInfo projection matrix.
// tf2::transformToEigen(t).matrix().block<3,4>(0,0) == [R(q) | t]; R(q) as Eigen's Quaternion::toRotationMatrix.
inline ses3d_camera make_camera(const geometry_msgs::TransformStamped& cam_from_base, const sensor_msgs::CameraInfo& info) {
  ses3d_camera c;
  std::memset(&c, 0, sizeof c);
  const double x = cam_from_base.transform.rotation.x, y = cam_from_base.transform.rotation.y,
               z = cam_from_base.transform.rotation.z, w = cam_from_base.transform.rotation.w;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y,
               tyz = tz * y, tzz = tz * z;
  double* T = c.T_cam_base;   // row-major 3x4
  T[0] = 1.0 - (tyy + tzz); T[1] = txy - twz;         T[2] = txz + twy;          T[3] = cam_from_base.transform.translation.x;
  T[4] = txy + twz;         T[5] = 1.0 - (txx + tzz); T[6] = tyz - twx;          T[7] = cam_from_base.transform.translation.y;
  T[8] = txz - twy;         T[9] = tyz + twx;         T[10] = 1.0 - (txx + tyy); T[11] = cam_from_base.transform.translation.z;
  // image_geometry::PinholeCameraModel::fx() ... Ty(): entries of CameraInfo.P
  c.fx = info.P[0]; c.fy = info.P[5]; c.cx = info.P[2]; c.cy = info.P[6]; c.Tx = info.P[3]; c.Ty = info.P[7];
  c.width = info.width;
  c.height = info.height;
  return c;
}

// person_msgs/Person2D -> ses3d_person2d. A detection that does not carry exactly 17 keypoints cannot be a COCO
// skeleton; it is passed on with all scores zero, which is how the reference treats an unusable detection
// (normalize_keypoints finds no valid keypoint, S3D:312-333, 540-552).
inline void to_pod(const person_msgs::Person2D& src, ses3d_person2d* dst) {
  std::memset(dst, 0, sizeof *dst);
  dst->score = src.score;
  if (src.keypoints.size() == SES3D_NUM_KEYPOINTS)
    for (int k = 0; k < SES3D_NUM_KEYPOINTS; ++k) {
      const person_msgs::Keypoint2D& kp = src.keypoints[k];
      ses3d_keypoint2d& o = dst->keypoints[k];
      o.x = kp.x; o.y = kp.y; o.score = kp.score;
      o.cov[0] = kp.cov[0]; o.cov[1] = kp.cov[1]; o.cov[2] = kp.cov[2];
    }
  for (int i = 0; i < 4; ++i) dst->bbox[i] = src.bbox[i];
}

inline void from_pod(const ses3d_person2d& src, person_msgs::Person2D* dst) {
  dst->score = src.score;
  dst->keypoints.resize(SES3D_NUM_KEYPOINTS);
  for (int k = 0; k < SES3D_NUM_KEYPOINTS; ++k) {
    const ses3d_keypoint2d& kp = src.keypoints[k];
    person_msgs::Keypoint2D& o = dst->keypoints[k];
    o.x = kp.x; o.y = kp.y; o.score = kp.score;
    o.cov[0] = kp.cov[0]; o.cov[1] = kp.cov[1]; o.cov[2] = kp.cov[2];
  }
  for (int i = 0; i < 4; ++i) dst->bbox[i] = src.bbox[i];
}

// person_msgs/PersonCov <-> ses3d_person_cov (21 keypoints in FUSION_BODY_PARTS order). Returns false when the
// message does not have 21 keypoints (REP:166-169, the person is skipped by the reference as well).
inline bool to_pod(const person_msgs::PersonCov& src, ses3d_person_cov* dst) {
  std::memset(dst, 0, sizeof *dst);
  if (src.keypoints.size() != SES3D_NUM_FUSION_KEYPOINTS) return false;
  dst->id = src.id;
  dst->score = src.score;
  for (int s = 0; s < SES3D_NUM_FUSION_KEYPOINTS; ++s) {
    const person_msgs::KeypointWithCovariance& kp = src.keypoints[s];
    ses3d_keypoint_cov& o = dst->keypoints[s];
    o.x = kp.joint.x; o.y = kp.joint.y; o.z = kp.joint.z;
    o.score = kp.score;
    for (int i = 0; i < 6; ++i) o.cov[i] = kp.cov[i];
  }
  dst->bbox_center[0] = src.bbox_center.position.x; dst->bbox_center[1] = src.bbox_center.position.y;
  dst->bbox_center[2] = src.bbox_center.position.z; dst->bbox_center[3] = src.bbox_center.orientation.x;
  dst->bbox_center[4] = src.bbox_center.orientation.y; dst->bbox_center[5] = src.bbox_center.orientation.z;
  dst->bbox_center[6] = src.bbox_center.orientation.w;
  dst->bbox_size[0] = src.bbox_size.x; dst->bbox_size[1] = src.bbox_size.y; dst->bbox_size[2] = src.bbox_size.z;
  return true;
}

inline void from_pod(const ses3d_person_cov& src, person_msgs::PersonCov* dst) {
  dst->id = src.id;
  dst->score = src.score;
  dst->keypoints.resize(SES3D_NUM_FUSION_KEYPOINTS);
  for (int s = 0; s < SES3D_NUM_FUSION_KEYPOINTS; ++s) {
    const ses3d_keypoint_cov& kp = src.keypoints[s];
    person_msgs::KeypointWithCovariance& o = dst->keypoints[s];
    o.joint.x = kp.x; o.joint.y = kp.y; o.joint.z = kp.z;
    o.score = kp.score;
    for (int i = 0; i < 6; ++i) o.cov[i] = kp.cov[i];
  }
  dst->bbox_center.position.x = src.bbox_center[0]; dst->bbox_center.position.y = src.bbox_center[1];
  dst->bbox_center.position.z = src.bbox_center[2]; dst->bbox_center.orientation.x = src.bbox_center[3];
  dst->bbox_center.orientation.y = src.bbox_center[4]; dst->bbox_center.orientation.z = src.bbox_center[5];
  dst->bbox_center.orientation.w = src.bbox_center[6];
  dst->bbox_size.x = src.bbox_size[0]; dst->bbox_size.y = src.bbox_size[1]; dst->bbox_size.z = src.bbox_size[2];
}

// The per-slot colour table both visualising nodes define in main() (S3D:1140-1169 == PRI define_colors): 21 fusion
// slots followed by 8 track colours. 8-bit RGB triples as the reference's comments give them.
inline std::vector<std_msgs::ColorRGBA> marker_colors() {
  static const unsigned char rgb[29][3] = {
      {255, 0, 0},   {85, 170, 0},  {0, 255, 0},   {0, 255, 170}, {0, 170, 255}, {85, 255, 0},  {0, 255, 85},
      {0, 255, 255}, {0, 85, 170},  {0, 0, 255},   {100, 0, 255}, {255, 0, 255}, {0, 85, 255},  {50, 0, 255},
      {170, 0, 255}, {255, 170, 0}, {255, 85, 0},  {170, 255, 0}, {255, 255, 0}, {255, 150, 0}, {42, 128, 85},
      {50, 0, 255},  {100, 0, 255}, {150, 0, 255}, {200, 0, 255}, {255, 0, 200}, {255, 0, 150}, {255, 0, 100},
      {255, 0, 50}};
  std::vector<std_msgs::ColorRGBA> out(29);
  for (int i = 0; i < 29; ++i) {
    out[i].r = rgb[i][0] / 255.0f; out[i].g = rgb[i][1] / 255.0f; out[i].b = rgb[i][2] / 255.0f; out[i].a = 1.0f;
  }
  out[20].g = 0.5f;   // Belly: the source writes 0.5, its comment says 128
  return out;
}

// Marker message assembly of skeleton_3d (S3D:688-715, 885-916, 968-973) from the numbers ses3d_markers_batch
// delivers for ONE frame: per published person a LINE_LIST "joints" marker and a SPHERE_LIST "joint_spheres" marker,
// plus (vis_cov) one SPHERE "joint_cov_3d" marker per joint of the first 15 fusion slots.
//   segments [n][22][2][3], n_segments [n], segment_slot [n][22], ellipsoids [n][21] (nullable unless vis_cov)
// Marker ids: the reference numbers markers by hypothesis index, which does not survive the plausibility filter;
// the index in the published list is used instead (ids only have to be unique per namespace).
inline void assemble_skeleton3d_markers(const std_msgs::Header& header, const ses3d_person_cov* persons, int n,
                                        const double* segments, const int32_t* n_segments, const int8_t* segment_slot,
                                        const ses3d_ellipsoid* ellipsoids, bool vis_cov, const int* kp2fusion /*[17]*/,
                                        const std::vector<std_msgs::ColorRGBA>& colors,
                                        visualization_msgs::MarkerArray* out) {
  for (int p = 0; p < n; ++p) {
    visualization_msgs::Marker lines;
    lines.header = header;
    lines.lifetime = ros::Duration(2.0);
    lines.pose.orientation.w = 1.0;
    lines.type = visualization_msgs::Marker::LINE_LIST;
    lines.scale.x = 0.05;
    lines.ns = "joints";
    lines.id = p;
    lines.color.r = 1.0f;
    lines.color.a = 1.0f;
    const double* seg = segments + (size_t)p * SES3D_MARKER_MAX_SEGMENTS * 6;
    for (int s = 0; s < n_segments[p]; ++s) {
      geometry_msgs::Point a, b;
      a.x = seg[s * 6 + 0]; a.y = seg[s * 6 + 1]; a.z = seg[s * 6 + 2];
      b.x = seg[s * 6 + 3]; b.y = seg[s * 6 + 4]; b.z = seg[s * 6 + 5];
      lines.points.push_back(a);
      lines.points.push_back(b);
      const std_msgs::ColorRGBA& col = colors[segment_slot[(size_t)p * SES3D_MARKER_MAX_SEGMENTS + s]];
      lines.colors.push_back(col);
      lines.colors.push_back(col);
    }
    visualization_msgs::Marker spheres;
    spheres.header = header;
    spheres.lifetime = ros::Duration(2.0);
    spheres.pose = lines.pose;
    spheres.type = visualization_msgs::Marker::SPHERE_LIST;
    spheres.scale.x = spheres.scale.y = spheres.scale.z = 0.07;
    spheres.ns = "joint_spheres";
    spheres.id = p;
    spheres.color.r = spheres.color.g = 0.5f;
    spheres.color.a = 1.0f;
    for (int k = 0; k < SES3D_NUM_KEYPOINTS; ++k) {
      const int slot = kp2fusion[k];
      const ses3d_keypoint_cov& kp = persons[p].keypoints[slot];
      if (!(kp.score > 0)) continue;
      geometry_msgs::Point pt;
      pt.x = kp.x; pt.y = kp.y; pt.z = kp.z;
      spheres.points.push_back(pt);
      spheres.colors.push_back(colors[slot]);
      if (vis_cov && ellipsoids && slot < 15) {
        const ses3d_ellipsoid& e = ellipsoids[(size_t)p * SES3D_NUM_FUSION_KEYPOINTS + slot];
        visualization_msgs::Marker cov;
        cov.header = header;
        cov.lifetime = ros::Duration(5.0);
        cov.type = visualization_msgs::Marker::SPHERE;
        cov.ns = "joint_cov_3d";
        cov.id = SES3D_NUM_FUSION_KEYPOINTS * p + k;
        cov.color = colors[slot];
        cov.color.a = 0.50f;
        cov.pose.position = pt;
        cov.pose.orientation.w = e.qw; cov.pose.orientation.x = e.qx; cov.pose.orientation.y = e.qy;
        cov.pose.orientation.z = e.qz;
        cov.scale.x = e.sx; cov.scale.y = e.sy; cov.scale.z = e.sz;
        out->markers.push_back(cov);
      }
    }
    out->markers.push_back(lines);
    out->markers.push_back(spheres);
  }
}

// kp2kpFusion_idx (S3D:133-146): detector joint -> fusion slot for the two detector models
inline const int* kp2fusion_table(bool h36m) {
  static const int simple[17] = {SES3D_FBP_NOSE, SES3D_FBP_LEYE, SES3D_FBP_REYE, SES3D_FBP_LEAR, SES3D_FBP_REAR,
                                 SES3D_FBP_LSHOULDER, SES3D_FBP_RSHOULDER, SES3D_FBP_LELBOW, SES3D_FBP_RELBOW,
                                 SES3D_FBP_LWRIST, SES3D_FBP_RWRIST, SES3D_FBP_LHIP, SES3D_FBP_RHIP, SES3D_FBP_LKNEE,
                                 SES3D_FBP_RKNEE, SES3D_FBP_LANKLE, SES3D_FBP_RANKLE};
  static const int h36[17] = {SES3D_FBP_NOSE, SES3D_FBP_HEAD, SES3D_FBP_NECK, SES3D_FBP_BELLY, SES3D_FBP_MIDHIP,
                              SES3D_FBP_LSHOULDER, SES3D_FBP_RSHOULDER, SES3D_FBP_LELBOW, SES3D_FBP_RELBOW,
                              SES3D_FBP_LWRIST, SES3D_FBP_RWRIST, SES3D_FBP_LHIP, SES3D_FBP_RHIP, SES3D_FBP_LKNEE,
                              SES3D_FBP_RKNEE, SES3D_FBP_LANKLE, SES3D_FBP_RANKLE};
  return h36m ? h36 : simple;
}

}  // namespace ses3d_ros

// pose_prior_ses3d_node — drop-in replacement of pose_prior/src/pose_prior_mult_node.cpp with tracking, the skeleton
// model fit and its marginal covariances running in libses3d (B200). Same node name, parameters, topics and types:
//
//   subscribes  human_pose_estimation/persons3d               person_msgs/PersonCovList
//   publishes   human_pose_estimation/persons3d_fused         person_msgs/PersonCovList
//               human_pose_estimation/persons3d_fused_pred    person_msgs/PersonCovList
//               human_pose_estimation/skeleton3d_fused        visualization_msgs/MarkerArray
//   parameters  ~pose_method, ~norm_height, ~vis_cov (PRI:930-932); ~device (0), ~max_tracks (32) (new)
//
// Replaced: skeletonCallback's body (PRI:506-903) and its file-scope tracker state by ses3d_prior_run with
// n_sequences = 1, n_frames = 1 (the state lives in the handle). Kept: header / ts_per_cam hand-over (PRI:528-532) and
// the three publish calls (PRI:905-907). No gtsam, no Hungarian.cpp, no OpenMP in this node.
#include <ros/ros.h>

#include "ses3d_ros/convert.h"

using person_msgs::PersonCovList;

namespace {

const std::string kPersonTopic = "human_pose_estimation/persons3d";

struct Node {
  ses3d_prior prior = nullptr;
  ses3d_handle geo = nullptr;   // marker numerics only (ses3d_markers_batch needs a geometry handle; any rig will do)
  bool vis_cov = false;
  ros::Publisher pub_fused, pub_pred, pub_markers;
  std::vector<std_msgs::ColorRGBA> colors = ses3d_ros::marker_colors();
  std::vector<ses3d_person_cov> in, fused, pred;
  std::vector<double> segments;
  std::vector<int32_t> n_segments;
  std::vector<int8_t> segment_slot;
  std::vector<ses3d_ellipsoid> ellipsoids;
};

// Marker message of pose_prior (PRI:256-382, 770-816): per published track one LINE_LIST in the track colour scheme of
// addJointToSkeleton, optionally the covariance ellipsoids.
void assemble_markers(Node& nd, const std_msgs::Header& header, int n, visualization_msgs::MarkerArray* out) {
  if (n == 0 || !nd.geo) return;
  const int h_max = (int)nd.fused.size();
  nd.segments.resize((size_t)h_max * SES3D_MARKER_MAX_SEGMENTS * 6);
  nd.n_segments.resize(h_max);
  nd.segment_slot.resize((size_t)h_max * SES3D_MARKER_MAX_SEGMENTS);
  nd.ellipsoids.resize((size_t)h_max * SES3D_NUM_FUSION_KEYPOINTS);
  int32_t n32 = n;
  if (ses3d_markers_batch(nd.geo, 1, h_max, nd.fused.data(), &n32, SES3D_MARKERS_POSE_PRIOR,
                          nd.vis_cov ? nd.ellipsoids.data() : nullptr, nd.segments.data(), nd.n_segments.data(),
                          nd.segment_slot.data(), SES3D_HOST_BUFFERS, nullptr) != SES3D_OK) {
    ROS_ERROR("ses3d_markers_batch: %s", ses3d_last_error_string());
    return;
  }
  for (int p = 0; p < n; ++p) {
    visualization_msgs::Marker lines;
    lines.header = header;
    lines.lifetime = ros::Duration(0.5);
    lines.pose.orientation.w = 1.0;
    lines.type = visualization_msgs::Marker::LINE_LIST;
    lines.scale.x = 0.05;
    lines.ns = "fused_skeleton";
    lines.id = (int32_t)nd.fused[p].id;
    lines.color = nd.colors[21 + nd.fused[p].id % 8];   // track colours follow the 21 slot colours
    const double* seg = nd.segments.data() + (size_t)p * SES3D_MARKER_MAX_SEGMENTS * 6;
    for (int s = 0; s < nd.n_segments[p]; ++s) {
      geometry_msgs::Point a, b;
      a.x = seg[s * 6 + 0]; a.y = seg[s * 6 + 1]; a.z = seg[s * 6 + 2];
      b.x = seg[s * 6 + 3]; b.y = seg[s * 6 + 4]; b.z = seg[s * 6 + 5];
      lines.points.push_back(a);
      lines.points.push_back(b);
    }
    out->markers.push_back(lines);
    if (!nd.vis_cov) continue;
    for (int s = 0; s < SES3D_NUM_FUSION_KEYPOINTS; ++s) {
      const ses3d_keypoint_cov& kp = nd.fused[p].keypoints[s];
      if (!(kp.score > 0)) continue;
      const ses3d_ellipsoid& e = nd.ellipsoids[(size_t)p * SES3D_NUM_FUSION_KEYPOINTS + s];
      visualization_msgs::Marker cov;
      cov.header = header;
      cov.lifetime = ros::Duration(0.5);
      cov.type = visualization_msgs::Marker::SPHERE;
      cov.ns = "fused_joint_cov";
      cov.id = SES3D_NUM_FUSION_KEYPOINTS * (int32_t)nd.fused[p].id + s;
      cov.color = nd.colors[s];
      cov.color.a = 0.5f;
      cov.pose.position.x = kp.x; cov.pose.position.y = kp.y; cov.pose.position.z = kp.z;
      cov.pose.orientation.w = e.qw; cov.pose.orientation.x = e.qx; cov.pose.orientation.y = e.qy;
      cov.pose.orientation.z = e.qz;
      cov.scale.x = e.sx; cov.scale.y = e.sy; cov.scale.z = e.sz;
      out->markers.push_back(cov);
    }
  }
}

void skeleton_callback(Node& nd, const PersonCovList::ConstPtr& msg) {
  const int n_in = (int)msg->persons.size(), h_max = std::max(n_in, 1), n_cams = (int)msg->fb_delay_per_cam.size();
  nd.in.assign(h_max, ses3d_person_cov());
  nd.fused.assign(h_max, ses3d_person_cov());
  nd.pred.assign(h_max, ses3d_person_cov());
  for (int i = 0; i < n_in; ++i) ses3d_ros::to_pod(msg->persons[i], &nd.in[i]);   // a malformed person stays all-zero
  const int64_t stamp_ns = (int64_t)msg->header.stamp.toNSec();
  int32_t n32 = n_in, n_out = 0;
  float pred_delay = 0.f;
  const int rc = ses3d_prior_run(nd.prior, 1, 1, h_max, nd.in.data(), &n32, &stamp_ns, n_cams,
                                 n_cams ? msg->fb_delay_per_cam.data() : nullptr, nd.fused.data(), nd.pred.data(), &n_out,
                                 &pred_delay, nullptr, SES3D_HOST_BUFFERS, nullptr);
  if (rc != SES3D_OK) {
    ROS_ERROR("ses3d_prior_run: %s", ses3d_last_error_string());
    return;
  }
  PersonCovList fused;   // PRI:528-532
  fused.header = msg->header;
  fused.ts_per_cam = msg->ts_per_cam;
  fused.fb_delay_per_cam.assign(msg->fb_delay_per_cam.size(), pred_delay);
  PersonCovList pred = fused;
  fused.persons.resize(n_out);
  pred.persons.resize(n_out);
  for (int i = 0; i < n_out; ++i) {
    ses3d_ros::from_pod(nd.fused[i], &fused.persons[i]);
    ses3d_ros::from_pod(nd.pred[i], &pred.persons[i]);
  }
  visualization_msgs::MarkerArray markers;
  assemble_markers(nd, msg->header, n_out, &markers);
  nd.pub_markers.publish(markers);
  nd.pub_fused.publish(fused);
  nd.pub_pred.publish(pred);
}

}  // namespace

int main(int argc, char** argv) {
  ros::init(argc, argv, "pose_prior");
  ros::NodeHandle nh;
  ros::NodeHandle nh_private("~");
  Node nd;
  std::string pose_method = "simple";
  bool norm_height = false;
  int device = 0, max_tracks = 32;
  nh_private.param<std::string>("pose_method", pose_method, "simple");
  nh_private.param<bool>("norm_height", norm_height, false);
  nh_private.param<bool>("vis_cov", nd.vis_cov, false);
  nh_private.param<int>("device", device, 0);
  nh_private.param<int>("max_tracks", max_tracks, 32);

  ses3d_prior_params pp;
  ses3d_prior_default_params(&pp);
  pp.pose_method = pose_method == "h36m" ? SES3D_POSE_H36M : SES3D_POSE_SIMPLE;
  pp.normalize_by_height = norm_height ? 1 : 0;
  if (ses3d_prior_create(&pp, 1, max_tracks, device, &nd.prior) != SES3D_OK) {
    ROS_ERROR("ses3d_prior_create: %s", ses3d_last_error_string());
    return -1;
  }
  {   // a two-camera dummy rig: the marker kernel only reads the skeleton tables of the handle
    ses3d_camera cams[2];
    std::memset(cams, 0, sizeof cams);
    for (int i = 0; i < 2; ++i) {
      cams[i].T_cam_base[0] = cams[i].T_cam_base[5] = cams[i].T_cam_base[10] = 1.0;
      cams[i].T_cam_base[3] = (double)i;
      cams[i].fx = cams[i].fy = 1000.0; cams[i].cx = 320.0; cams[i].cy = 240.0; cams[i].width = 640; cams[i].height = 480;
    }
    ses3d_params prm;
    ses3d_default_params(&prm);
    prm.pose_method = pp.pose_method;
    if (ses3d_create(2, cams, &prm, device, &nd.geo) != SES3D_OK) nd.geo = nullptr;   // markers are optional
  }
  ROS_INFO("Using pose method %s (norm_height = %d)", pose_method.c_str(), (int)norm_height);

  nd.pub_fused = nh.advertise<PersonCovList>("human_pose_estimation/persons3d_fused", 1);
  nd.pub_pred = nh.advertise<PersonCovList>("human_pose_estimation/persons3d_fused_pred", 1);
  nd.pub_markers = nh.advertise<visualization_msgs::MarkerArray>("human_pose_estimation/skeleton3d_fused", 1);
  ros::Subscriber sub = nh.subscribe<PersonCovList>(
      kPersonTopic, 1, [&nd](const PersonCovList::ConstPtr& m) { skeleton_callback(nd, m); }, ros::VoidConstPtr(),
      ros::TransportHints().tcpNoDelay());
  ros::spin();
  ses3d_prior_destroy(nd.prior);
  if (nd.geo) ses3d_destroy(nd.geo);
  return 0;
}

// pose_reproj_ses3d_node — drop-in replacement of pose_reprojection/src/skeleton_reproj_mult_node.cpp with the
// sigma-point projection running in libses3d (B200). Same node name, parameters, topics and message types:
//
//   subscribes  human_pose_estimation/persons3d_fused_pred   person_msgs/PersonCovList
//               <cam>/color/camera_info, tf <cam>_color_optical_frame <- base          once, at start-up
//   publishes   <cam>/skel_pred                              person_msgs/Person2DList   one per camera
//   parameters  ~pose_method, ~cameras (REP:243-257); ~device (0) (new)
//
// Kept from the reference: the frame_id check (REP:140-143), the per-camera header / fb_delay fields (REP:158-160)
// and one publish per camera per message, empty lists included (REP:233-234). Replaced: fusedSkeletonCallback's
// body (REP:145-231) by ses3d_reproject_batch with n_frames = 1. No Eigen / image_geometry / cv_bridge in this node.
#include <ros/ros.h>
#include <tf2_ros/transform_listener.h>

#include "ses3d_ros/convert.h"

using person_msgs::Person2DList;
using person_msgs::PersonCovList;

namespace {

const std::string kBaseFrame = "base";
const std::string kCamFrameSuffix = "_color_optical_frame";
const std::string kCamInfoSuffix = "/color/camera_info";
const std::string kSkelPredSuffix = "/skel_pred";
const std::string kFusedTopic = "human_pose_estimation/persons3d_fused_pred";

struct Node {
  unsigned n_cams = 4;
  std::vector<std::string> cam_frames{"cam_1_color_optical_frame", "cam_2_color_optical_frame",
                                      "cam_3_color_optical_frame", "cam_4_color_optical_frame"};
  std::vector<std::string> cam_info_topics{"cam_1/color/camera_info", "cam_2/color/camera_info",
                                           "cam_3/color/camera_info", "cam_4/color/camera_info"};
  std::vector<std::string> pred_topics{"cam_1/skel_pred", "cam_2/skel_pred", "cam_3/skel_pred", "cam_4/skel_pred"};
  std::vector<sensor_msgs::CameraInfo> intrinsics;
  std::vector<ros::Publisher> pubs;
  ses3d_handle geo = nullptr;
  std::vector<ses3d_person_cov> in;
  std::vector<ses3d_person2d> out;
  std::vector<int32_t> n_out;
};

void fused_skeleton_callback(Node& nd, const PersonCovList::ConstPtr& msg) {
  if (msg->header.frame_id != kBaseFrame) {
    ROS_ERROR("Fused person is not given in \"%s\" but in \"%s\". Aborting!", kBaseFrame.c_str(), msg->header.frame_id.c_str());
    return;
  }
  std::vector<Person2DList> lists(nd.n_cams);
  for (unsigned c = 0; c < nd.n_cams; ++c) {
    lists[c].header.frame_id = nd.intrinsics[c].header.frame_id;
    if (c < msg->ts_per_cam.size()) lists[c].header.stamp = msg->ts_per_cam[c];
    if (c < msg->fb_delay_per_cam.size()) lists[c].fb_delay = msg->fb_delay_per_cam[c];
  }
  // persons without 21 keypoints are skipped (REP:166-169)
  nd.in.clear();
  for (size_t i = 0; i < msg->persons.size(); ++i) {
    ses3d_person_cov p;
    if (!ses3d_ros::to_pod(msg->persons[i], &p)) {
      ROS_ERROR("Fused person %zu: Expected skeleton to have %d Keypoints, but got %zu. Aborting!", i,
                SES3D_NUM_FUSION_KEYPOINTS, msg->persons[i].keypoints.size());
      continue;
    }
    nd.in.push_back(p);
  }
  const int32_t n = (int32_t)nd.in.size();
  if (n > 0) {
    const int h_max = n;
    nd.out.resize((size_t)nd.n_cams * h_max);
    nd.n_out.assign(nd.n_cams, 0);
    const int rc = ses3d_reproject_batch(nd.geo, 1, h_max, nd.in.data(), &n, nd.out.data(), nd.n_out.data(),
                                         SES3D_HOST_BUFFERS, nullptr);
    if (rc != SES3D_OK) {
      ROS_ERROR("ses3d_reproject_batch: %s", ses3d_last_error_string());
      return;
    }
    for (unsigned c = 0; c < nd.n_cams; ++c) {
      lists[c].persons.resize(nd.n_out[c]);
      for (int i = 0; i < nd.n_out[c]; ++i) ses3d_ros::from_pod(nd.out[(size_t)c * h_max + i], &lists[c].persons[i]);
    }
  }
  for (unsigned c = 0; c < nd.n_cams; ++c) nd.pubs[c].publish(lists[c]);
}

}  // namespace

int main(int argc, char** argv) {
  ros::init(argc, argv, "multi_skeleton_reprojection");
  ros::NodeHandle nh;
  ros::NodeHandle nh_private("~");
  Node nd;
  std::string pose_method = "simple";
  int device = 0;
  nh_private.param<std::string>("pose_method", pose_method, "simple");
  nh_private.param<int>("device", device, 0);
  std::vector<std::string> cam_names;
  nh_private.param("cameras", cam_names, std::vector<std::string>());
  if (!cam_names.empty()) {
    nd.n_cams = (unsigned)cam_names.size();
    nd.cam_frames.clear(); nd.cam_info_topics.clear(); nd.pred_topics.clear();
    for (const std::string& c : cam_names) {
      nd.cam_frames.push_back(c + kCamFrameSuffix);
      nd.cam_info_topics.push_back(c + kCamInfoSuffix);
      nd.pred_topics.push_back(c + kSkelPredSuffix);
    }
  }
  ROS_INFO("NUM_CAMERAS: %u, pose estimation method: %s", nd.n_cams, pose_method.c_str());

  tf2_ros::Buffer tf_buffer;
  tf2_ros::TransformListener tf_listener(tf_buffer);
  std::vector<geometry_msgs::TransformStamped> transforms;
  while (ros::ok()) {   // getTransforms, REP:77-104
    transforms.clear();
    try {
      for (unsigned i = 0; i < nd.n_cams; ++i) transforms.push_back(tf_buffer.lookupTransform(nd.cam_frames[i], kBaseFrame, ros::Time(0)));
      break;
    } catch (tf2::TransformException& ex) {
      ROS_WARN("%s", ex.what());
      ros::Duration(1.0).sleep();
      ros::spinOnce();
    }
  }
  nd.intrinsics.assign(nd.n_cams, sensor_msgs::CameraInfo());   // getIntrinsics, REP:110-137
  {
    std::vector<char> seen(nd.n_cams, 0);
    std::vector<ros::Subscriber> subs;
    for (unsigned i = 0; i < nd.n_cams; ++i)
      subs.push_back(nh.subscribe<sensor_msgs::CameraInfo>(
          nd.cam_info_topics[i], 1, [&nd, &seen, i](const sensor_msgs::CameraInfo::ConstPtr& m) { nd.intrinsics[i] = *m; seen[i] = 1; }));
    ros::Rate rate(1.0);
    bool all = false;
    for (int tries = 0; ros::ok() && !all && tries < 600; ++tries) {
      ros::spinOnce();
      all = true;
      for (unsigned i = 0; i < nd.n_cams; ++i)
        all = all && seen[i] && !(nd.intrinsics[i].D.empty() && nd.intrinsics[i].distortion_model != "none");
      if (!all) rate.sleep();
    }
    if (!all) return -1;
  }
  if (transforms.size() != nd.n_cams) {
    ROS_ERROR("incoherent number of transforms, intrinsics and output heatmaps! Aborting!");
    return -1;
  }
  for (unsigned i = 0; i < nd.n_cams; ++i) nd.pubs.push_back(nh.advertise<Person2DList>(nd.pred_topics[i], 1));

  std::vector<ses3d_camera> cams(nd.n_cams);
  for (unsigned i = 0; i < nd.n_cams; ++i) cams[i] = ses3d_ros::make_camera(transforms[i], nd.intrinsics[i]);
  ses3d_params prm;
  ses3d_default_params(&prm);
  prm.pose_method = pose_method == "h36m" ? SES3D_POSE_H36M : SES3D_POSE_SIMPLE;
  if (ses3d_create((int32_t)nd.n_cams, cams.data(), &prm, device, &nd.geo) != SES3D_OK) {
    ROS_ERROR("ses3d_create: %s", ses3d_last_error_string());
    return -1;
  }
  ROS_INFO("Reprojecting into %u camera views", nd.n_cams);

  ros::Subscriber sub = nh.subscribe<PersonCovList>(
      kFusedTopic, 1, [&nd](const PersonCovList::ConstPtr& m) { fused_skeleton_callback(nd, m); }, ros::VoidConstPtr(),
      ros::TransportHints().tcpNoDelay());
  ros::spin();
  ses3d_destroy(nd.geo);
  return 0;
}

// skeleton_3d_ses3d_node — drop-in replacement of skeleton_3d/src/skeleton_3d_triang_mult_node.cpp with the geometry
// running in libses3d (B200). Same node name, parameters, topics and message types as the reference node:
//
//   subscribes  <cam>/human_joints            person_msgs/Person2DList   one per camera, approximate-time synchronised
//               <cam>/color/camera_info       sensor_msgs/CameraInfo     once, at start-up
//               tf  <cam>_color_optical_frame <- base                    once, at start-up
//   publishes   human_pose_estimation/persons3d          person_msgs/PersonCovList
//               human_pose_estimation/skeleton3d_vis     visualization_msgs/MarkerArray
//   parameters  ~pose_method ("simple" | "h36m"), ~vis_cov, ~max_epi_dist, ~cameras          (as the reference, S3D:1095-1126)
//               ~device (0), ~h_max (32), ~lossless (false), ~precision ("fp32" | "fp64")    (new)
//
// What stays as in the reference: the synchroniser set-up (S3D:1218-1223), the 1-slot latest-wins mailbox between the
// ROS spinner and the worker thread (S3D:999-1025), the worker's pivot / backwards-time / stale-camera gating
// (S3D:1029-1057) and the output header fields (S3D:1059-1065). What is replaced: the tf/CameraInfo -> table set-up
// (S3D:1184-1214, now ses3d_create) and triangulate_persons (S3D:525-997, now ses3d_triangulate_batch with
// n_frames = 1) including the marker numerics (ses3d_markers_batch). No Eigen, image_geometry, OpenMP or Hungarian.cpp
// in this node.
//
// ~lossless = true turns the mailbox into a hand-over that never overwrites an unread frame (the synchroniser callback
// waits for the worker): for bag replay where every frame must be processed.
#include <message_filters/subscriber.h>
#include <my_message_filters/sync_policies/approximate_time_vec.h>
#include <my_message_filters/synchronizer_vec.h>
#include <ros/ros.h>
#include <tf2_ros/transform_listener.h>

#include <condition_variable>
#include <mutex>
#include <thread>

#include "ses3d_ros/convert.h"

using person_msgs::Person2DList;
using person_msgs::PersonCovList;

namespace {

const std::string kBaseFrame = "base";
const std::string kCamFrameSuffix = "_color_optical_frame";
const std::string kCamInfoSuffix = "/color/camera_info";
const std::string kSkel2dSuffix = "/human_joints";
const std::string kPersons3dTopic = "human_pose_estimation/persons3d";
const std::string kSkeleton3dTopic = "human_pose_estimation/skeleton3d_vis";
const double kMaxSyncDiff = 0.067;   // g_max_sync_diff, S3D:64

struct Mailbox {   // S3D:999-1025
  std::mutex mu;
  std::condition_variable cv, cv_taken;
  std::vector<Person2DList::ConstPtr> frame;
  bool updated = false;
  bool lossless = false;
  bool stop = false;
  void put(const std::vector<Person2DList::ConstPtr>& f) {
    {
      std::unique_lock<std::mutex> lk(mu);
      if (lossless) cv_taken.wait(lk, [this] { return !updated; });
      frame = f;   // latest wins: an unread frame is overwritten
      updated = true;
    }
    cv.notify_one();
  }
  // Blocks until a frame is available; an empty vector means "shut down" (and nothing is pending).
  std::vector<Person2DList::ConstPtr> take() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [this] { return updated || stop; });
    if (!updated) return {};
    std::vector<Person2DList::ConstPtr> f = frame;
    updated = false;
    lk.unlock();
    cv_taken.notify_one();
    return f;
  }
  // After ros::spin() returned. The reference wakes its worker with the last frame again and lets the backwards-time
  // rule discard it (S3D:1226-1232); an explicit flag does the same without the second pass. A frame that is still
  // in the slot is processed before the worker leaves.
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_one();
  }
};

struct Node {
  unsigned n_cams = 4;
  std::vector<std::string> cam_frames{"cam_1_color_optical_frame", "cam_2_color_optical_frame",
                                      "cam_3_color_optical_frame", "cam_4_color_optical_frame"};
  std::vector<std::string> cam_info_topics{"cam_1/color/camera_info", "cam_2/color/camera_info",
                                           "cam_3/color/camera_info", "cam_4/color/camera_info"};
  std::vector<std::string> skel_topics{"cam_1/human_joints", "cam_2/human_joints", "cam_3/human_joints",
                                       "cam_4/human_joints"};
  std::string pose_method = "simple";
  bool vis_cov = false;
  int h_max = 32;
  ses3d_handle geo = nullptr;
  std::vector<std_msgs::ColorRGBA> colors = ses3d_ros::marker_colors();
  Mailbox mailbox;
  // staging reused across frames
  std::vector<ses3d_person2d> in;
  std::vector<int32_t> n_in;
  std::vector<ses3d_person_cov> out;
  std::vector<double> segments;
  std::vector<int32_t> n_segments;
  std::vector<int8_t> segment_slot;
  std::vector<ses3d_ellipsoid> ellipsoids;
};

bool wait_for_transforms(const Node& nd, const tf2_ros::Buffer& tf, std::vector<geometry_msgs::TransformStamped>* out) {
  while (ros::ok()) {   // getTransforms, S3D:161-191
    out->clear();
    try {
      for (unsigned i = 0; i < nd.n_cams; ++i) out->push_back(tf.lookupTransform(nd.cam_frames[i], kBaseFrame, ros::Time(0)));
    } catch (tf2::TransformException& ex) {
      ROS_WARN("%s", ex.what());
      ros::Duration(1.0).sleep();
      if (!ros::ok()) break;
      ros::spinOnce();
      continue;
    }
    ROS_INFO("Sucessfully retrieved camera extrinsic transforms.");
    return true;
  }
  return false;
}

bool wait_for_intrinsics(const Node& nd, ros::NodeHandle& nh, std::vector<sensor_msgs::CameraInfo>* out) {
  out->assign(nd.n_cams, sensor_msgs::CameraInfo());   // getIntrinsics, S3D:197-228
  std::vector<char> seen(nd.n_cams, 0);
  std::vector<ros::Subscriber> subs;
  for (unsigned i = 0; i < nd.n_cams; ++i)
    subs.push_back(nh.subscribe<sensor_msgs::CameraInfo>(
        nd.cam_info_topics[i], 1, [out, &seen, i](const sensor_msgs::CameraInfo::ConstPtr& m) { (*out)[i] = *m; seen[i] = 1; }));
  ros::Rate rate(1.0);
  for (int tries = 0; ros::ok(); ++tries) {
    ros::spinOnce();
    bool all = true;
    for (unsigned i = 0; i < nd.n_cams; ++i)
      all = all && seen[i] && !((*out)[i].D.empty() && (*out)[i].distortion_model != "none");
    if (all) {
      ROS_INFO("intrinsics received.");
      return true;
    }
    ROS_INFO("Spinning.. Waiting to receive camera intrinsics.");
    rate.sleep();
    if (tries > 600) break;
  }
  return false;
}

// One synchronised frame through the library; fills the two output messages. Replaces triangulate_persons.
void process_frame(Node& nd, const std::vector<Person2DList::ConstPtr>& people, PersonCovList* msg,
                   visualization_msgs::MarkerArray* vis) {
  int p_max = 1;
  for (const auto& m : people) p_max = std::max<int>(p_max, (int)std::min<size_t>(m->persons.size(), 127));
  nd.in.assign((size_t)nd.n_cams * p_max, ses3d_person2d());
  nd.n_in.assign(nd.n_cams, 0);
  for (unsigned c = 0; c < nd.n_cams; ++c) {
    const int n = (int)std::min<size_t>(people[c]->persons.size(), (size_t)p_max);
    nd.n_in[c] = n;
    for (int d = 0; d < n; ++d) ses3d_ros::to_pod(people[c]->persons[d], &nd.in[(size_t)c * p_max + d]);
  }
  nd.out.resize(nd.h_max);
  int32_t n_out = 0;
  int rc = ses3d_triangulate_batch(nd.geo, 1, p_max, nd.in.data(), nd.n_in.data(), nd.h_max, nd.out.data(), &n_out,
                                   nullptr, SES3D_HOST_BUFFERS, nullptr);
  if (rc != SES3D_OK) {
    ROS_ERROR("ses3d_triangulate_batch: %s", ses3d_last_error_string());
    return;
  }
  msg->persons.resize(n_out);
  for (int i = 0; i < n_out; ++i) ses3d_ros::from_pod(nd.out[i], &msg->persons[i]);
  if (n_out == 0) return;
  nd.segments.resize((size_t)nd.h_max * SES3D_MARKER_MAX_SEGMENTS * 6);
  nd.n_segments.resize(nd.h_max);
  nd.segment_slot.resize((size_t)nd.h_max * SES3D_MARKER_MAX_SEGMENTS);
  nd.ellipsoids.resize((size_t)nd.h_max * SES3D_NUM_FUSION_KEYPOINTS);
  rc = ses3d_markers_batch(nd.geo, 1, nd.h_max, nd.out.data(), &n_out, SES3D_MARKERS_SKELETON3D,
                           nd.vis_cov ? nd.ellipsoids.data() : nullptr, nd.segments.data(), nd.n_segments.data(),
                           nd.segment_slot.data(), SES3D_HOST_BUFFERS, nullptr);
  if (rc != SES3D_OK) {
    ROS_ERROR("ses3d_markers_batch: %s", ses3d_last_error_string());
    return;
  }
  ses3d_ros::assemble_skeleton3d_markers(msg->header, nd.out.data(), n_out, nd.segments.data(), nd.n_segments.data(),
                                         nd.segment_slot.data(), nd.vis_cov ? nd.ellipsoids.data() : nullptr, nd.vis_cov,
                                         ses3d_ros::kp2fusion_table(nd.pose_method == "h36m"), nd.colors, vis);
}

void worker(Node& nd, const ros::Publisher& pub3d, const ros::Publisher& pub_vis) {
  double last_stamp = 0;
  std::vector<Person2DList::Ptr> dummy(nd.n_cams);
  for (auto& d : dummy) d.reset(new Person2DList);
  for (;;) {
    std::vector<Person2DList::ConstPtr> people = nd.mailbox.take();
    if (people.empty()) break;
    if (people.size() != nd.n_cams) continue;
    double t_max = 0.0;   // newest message = pivot, S3D:1031-1038
    int pivot = -1;
    for (unsigned i = 0; i < nd.n_cams; ++i)
      if (people[i]->header.stamp.toSec() > t_max) { t_max = people[i]->header.stamp.toSec(); pivot = (int)i; }
    if (pivot < 0) continue;
    const double delta_t = t_max - last_stamp;
    if (delta_t > 0.17) ROS_WARN("Large frame delay delta_t = %fs (should be < 0.17s)", delta_t);
    if (delta_t <= 0.0) continue;   // re-used message or jumped backwards in time, S3D:1043-1046
    last_stamp = t_max;
    for (unsigned i = 0; i < nd.n_cams; ++i) {   // stale cameras are replaced by an empty list, S3D:1049-1057
      const double dt = t_max - people[i]->header.stamp.toSec();
      if (dt > kMaxSyncDiff) {
        dummy[i]->header = people[i]->header;
        dummy[i]->fb_delay = people[i]->fb_delay;
        people[i] = dummy[i];
        ROS_WARN("sync time diff of msg %u larger than %.0fms: %.1fms (w.r.t. pivot msg %d). REMOVING.", i,
                 kMaxSyncDiff * 1000, dt * 1000, pivot);
      }
    }
    PersonCovList msg;   // S3D:1059-1065
    msg.header = people[pivot]->header;
    for (unsigned i = 0; i < nd.n_cams; ++i) {
      msg.ts_per_cam.push_back(people[i]->header.stamp);
      msg.fb_delay_per_cam.push_back(people[i]->fb_delay);
    }
    msg.header.frame_id = kBaseFrame;
    visualization_msgs::MarkerArray vis;
    process_frame(nd, people, &msg, &vis);
    pub3d.publish(msg);
    if (!vis.markers.empty()) pub_vis.publish(vis);
  }
}

}  // namespace

int main(int argc, char** argv) {
  ros::init(argc, argv, "skeleton_singlePerson_3d");
  ros::NodeHandle nh;
  ros::NodeHandle nh_private("~");
  Node nd;
  double max_epi_dist = 0.050;
  int device = 0;
  std::string precision = "fp32";
  nh_private.param<std::string>("pose_method", nd.pose_method, "simple");
  nh_private.param<bool>("vis_cov", nd.vis_cov, false);
  nh_private.param<double>("max_epi_dist", max_epi_dist, 0.050);
  nh_private.param<int>("device", device, 0);
  nh_private.param<int>("h_max", nd.h_max, 32);
  nh_private.param<bool>("lossless", nd.mailbox.lossless, false);
  nh_private.param<std::string>("precision", precision, "fp32");
  std::vector<std::string> cam_names;
  nh_private.param("cameras", cam_names, std::vector<std::string>());
  if (!cam_names.empty()) {   // S3D:1114-1126
    nd.n_cams = (unsigned)cam_names.size();
    nd.cam_frames.clear(); nd.cam_info_topics.clear(); nd.skel_topics.clear();
    for (const std::string& c : cam_names) {
      nd.cam_frames.push_back(c + kCamFrameSuffix);
      nd.cam_info_topics.push_back(c + kCamInfoSuffix);
      nd.skel_topics.push_back(c + kSkel2dSuffix);
    }
  }
  ROS_INFO("NUM_CAMERAS: %u", nd.n_cams);
  if (nd.n_cams < 2) {
    ROS_ERROR("Need at least 2 cameras for triangulation. Aborting!");
    return -1;
  }

  ros::Publisher pub3d = nh.advertise<PersonCovList>(kPersons3dTopic, 1);
  ros::Publisher pub_vis = nh.advertise<visualization_msgs::MarkerArray>(kSkeleton3dTopic, 1);
  std::vector<message_filters::Subscriber<Person2DList>> subs(nd.n_cams);
  for (unsigned i = 0; i < nd.n_cams; ++i) subs[i].subscribe(nh, nd.skel_topics[i], 1, ros::TransportHints().tcpNoDelay());

  tf2_ros::Buffer tf_buffer;
  tf2_ros::TransformListener tf_listener(tf_buffer);
  std::vector<geometry_msgs::TransformStamped> transforms;
  std::vector<sensor_msgs::CameraInfo> intrinsics;
  if (!wait_for_transforms(nd, tf_buffer, &transforms) || !wait_for_intrinsics(nd, nh, &intrinsics)) return -1;

  // camera tables, fundamental matrices, skeleton model: all inside the library (replaces S3D:1184-1214)
  std::vector<ses3d_camera> cams(nd.n_cams);
  for (unsigned i = 0; i < nd.n_cams; ++i) cams[i] = ses3d_ros::make_camera(transforms[i], intrinsics[i]);
  ses3d_params prm;
  ses3d_default_params(&prm);
  prm.pose_method = nd.pose_method == "h36m" ? SES3D_POSE_H36M : SES3D_POSE_SIMPLE;
  prm.precision = precision == "fp64" ? SES3D_PRECISION_FP64 : SES3D_PRECISION_FP32;
  prm.max_epipolar_error = max_epi_dist;
  if (ses3d_create((int32_t)nd.n_cams, cams.data(), &prm, device, &nd.geo) != SES3D_OK) {
    ROS_ERROR("ses3d_create: %s", ses3d_last_error_string());
    return -1;
  }
  ROS_INFO("%s: %u cameras, pose method %s, max epipolar dist %f", ses3d_version(), nd.n_cams, nd.pose_method.c_str(),
           max_epi_dist);

  std::thread worker_thread(worker, std::ref(nd), std::cref(pub3d), std::cref(pub_vis));

  typedef message_filters::sync_policies::ApproximateTimeVec<Person2DList> SyncPolicy;   // S3D:1218-1223
  SyncPolicy policy(std::max(3u, 1 + nd.n_cams / 4), nd.n_cams);
  policy.setInterMessageLowerBound(ros::Duration(0.020));
  policy.setAgePenalty(2.0);
  message_filters::SynchronizerVec<SyncPolicy> sync((SyncPolicy)policy, subs);
  sync.registerCallback([&nd](const std::vector<Person2DList::ConstPtr>& people) { nd.mailbox.put(people); });
  ros::spin();

  nd.mailbox.shutdown();
  worker_thread.join();
  ses3d_destroy(nd.geo);
  return 0;
}

// ros_stub/msgs.h — the message types the shim nodes touch, with ROS 1 wire (de)serialisation. Field order follows the
// .msg definitions (std_msgs, geometry_msgs, sensor_msgs, visualization_msgs of ROS 1; person_msgs/msg/*.msg of the
// reference). TEST INFRASTRUCTURE ONLY (see core.h).
#pragma once
#include "ros_stub/core.h"

#define ROS_STUB_PTRS(M)                       \
  typedef std::shared_ptr<M> Ptr;              \
  typedef std::shared_ptr<M const> ConstPtr;

namespace std_msgs {
struct Header {
  uint32_t seq = 0;
  ros::Time stamp;
  std::string frame_id;
};
inline void ser(ros::stub::Writer& w, const Header& m) { using namespace ros::stub; ser(w, m.seq); ser(w, m.stamp); ser(w, m.frame_id); }
inline void de(ros::stub::Reader& r, Header& m) { using namespace ros::stub; de(r, m.seq); de(r, m.stamp); de(r, m.frame_id); }
struct ColorRGBA {
  float r = 0, g = 0, b = 0, a = 0;
};
inline void ser(ros::stub::Writer& w, const ColorRGBA& m) { using namespace ros::stub; ser(w, m.r); ser(w, m.g); ser(w, m.b); ser(w, m.a); }
inline void de(ros::stub::Reader& r, ColorRGBA& m) { using namespace ros::stub; de(r, m.r); de(r, m.g); de(r, m.b); de(r, m.a); }
}  // namespace std_msgs

namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped {
  std_msgs::Header header;
  std::string child_frame_id;
  Transform transform;
  ROS_STUB_PTRS(TransformStamped)
};
inline void ser(ros::stub::Writer& w, const Point& m) { using namespace ros::stub; ser(w, m.x); ser(w, m.y); ser(w, m.z); }
inline void de(ros::stub::Reader& r, Point& m) { using namespace ros::stub; de(r, m.x); de(r, m.y); de(r, m.z); }
inline void ser(ros::stub::Writer& w, const Vector3& m) { using namespace ros::stub; ser(w, m.x); ser(w, m.y); ser(w, m.z); }
inline void de(ros::stub::Reader& r, Vector3& m) { using namespace ros::stub; de(r, m.x); de(r, m.y); de(r, m.z); }
inline void ser(ros::stub::Writer& w, const Quaternion& m) { using namespace ros::stub; ser(w, m.x); ser(w, m.y); ser(w, m.z); ser(w, m.w); }
inline void de(ros::stub::Reader& r, Quaternion& m) { using namespace ros::stub; de(r, m.x); de(r, m.y); de(r, m.z); de(r, m.w); }
inline void ser(ros::stub::Writer& w, const Pose& m) { ser(w, m.position); ser(w, m.orientation); }
inline void de(ros::stub::Reader& r, Pose& m) { de(r, m.position); de(r, m.orientation); }
}  // namespace geometry_msgs

namespace sensor_msgs {
struct RegionOfInterest {
  uint32_t x_offset = 0, y_offset = 0, height = 0, width = 0;
  bool do_rectify = false;
};
struct CameraInfo {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::string distortion_model;
  std::vector<double> D;
  ros::stub::FixedArray<double, 9> K, R;
  ros::stub::FixedArray<double, 12> P;
  uint32_t binning_x = 0, binning_y = 0;
  RegionOfInterest roi;
  ROS_STUB_PTRS(CameraInfo)
};
inline void ser(ros::stub::Writer& w, const RegionOfInterest& m) {
  using namespace ros::stub;
  ser(w, m.x_offset); ser(w, m.y_offset); ser(w, m.height); ser(w, m.width); ser(w, m.do_rectify);
}
inline void de(ros::stub::Reader& r, RegionOfInterest& m) {
  using namespace ros::stub;
  de(r, m.x_offset); de(r, m.y_offset); de(r, m.height); de(r, m.width); de(r, m.do_rectify);
}
inline void ser(ros::stub::Writer& w, const CameraInfo& m) {
  using namespace ros::stub;
  ser(w, m.header); ser(w, m.height); ser(w, m.width); ser(w, m.distortion_model); ser(w, m.D); ser(w, m.K); ser(w, m.R);
  ser(w, m.P); ser(w, m.binning_x); ser(w, m.binning_y); ser(w, m.roi);
}
inline void de(ros::stub::Reader& r, CameraInfo& m) {
  using namespace ros::stub;
  de(r, m.header); de(r, m.height); de(r, m.width); de(r, m.distortion_model); de(r, m.D); de(r, m.K); de(r, m.R);
  de(r, m.P); de(r, m.binning_x); de(r, m.binning_y); de(r, m.roi);
}
}  // namespace sensor_msgs

namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7,
         POINTS = 8, TEXT_VIEW_FACING = 9, MESH_RESOURCE = 10, TRIANGLE_LIST = 11 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
  std_msgs::Header header;
  std::string ns;
  int32_t id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  ros::Duration lifetime;
  bool frame_locked = false;
  std::vector<geometry_msgs::Point> points;
  std::vector<std_msgs::ColorRGBA> colors;
  std::string text, mesh_resource;
  bool mesh_use_embedded_materials = false;
};
struct MarkerArray {
  std::vector<Marker> markers;
  ROS_STUB_PTRS(MarkerArray)
};
inline void ser(ros::stub::Writer& w, const Marker& m) {
  using namespace ros::stub;
  ser(w, m.header); ser(w, m.ns); ser(w, m.id); ser(w, m.type); ser(w, m.action); ser(w, m.pose); ser(w, m.scale);
  ser(w, m.color); ser(w, m.lifetime); ser(w, m.frame_locked); ser(w, m.points); ser(w, m.colors); ser(w, m.text);
  ser(w, m.mesh_resource); ser(w, m.mesh_use_embedded_materials);
}
inline void de(ros::stub::Reader& r, Marker& m) {
  using namespace ros::stub;
  de(r, m.header); de(r, m.ns); de(r, m.id); de(r, m.type); de(r, m.action); de(r, m.pose); de(r, m.scale);
  de(r, m.color); de(r, m.lifetime); de(r, m.frame_locked); de(r, m.points); de(r, m.colors); de(r, m.text);
  de(r, m.mesh_resource); de(r, m.mesh_use_embedded_materials);
}
inline void ser(ros::stub::Writer& w, const MarkerArray& m) { using namespace ros::stub; ser(w, m.markers); }
inline void de(ros::stub::Reader& r, MarkerArray& m) { using namespace ros::stub; de(r, m.markers); }
}  // namespace visualization_msgs

namespace person_msgs {
struct Keypoint2D {   // person_msgs/msg/Keypoint2D.msg
  float x = 0, y = 0, score = 0;
  ros::stub::FixedArray<float, 3> cov;
};
struct Person2D {     // person_msgs/msg/Person2D.msg
  float score = 0;
  std::vector<Keypoint2D> keypoints;
  ros::stub::FixedArray<float, 4> bbox;
};
struct Person2DList { // person_msgs/msg/Person2DList.msg
  std_msgs::Header header;
  float fb_delay = 0;
  std::vector<Person2D> persons;
  ROS_STUB_PTRS(Person2DList)
};
struct KeypointWithCovariance {   // person_msgs/msg/KeypointWithCovariance.msg
  geometry_msgs::Point joint;
  float score = 0;
  ros::stub::FixedArray<double, 6> cov;
};
struct PersonCov {    // person_msgs/msg/PersonCov.msg
  uint32_t id = 0;
  float score = 0;
  std::vector<KeypointWithCovariance> keypoints;
  geometry_msgs::Pose bbox_center;
  geometry_msgs::Vector3 bbox_size;
};
struct PersonCovList { // person_msgs/msg/PersonCovList.msg
  std_msgs::Header header;
  std::vector<ros::Time> ts_per_cam;
  std::vector<float> fb_delay_per_cam;
  std::vector<PersonCov> persons;
  ROS_STUB_PTRS(PersonCovList)
};
inline void ser(ros::stub::Writer& w, const Keypoint2D& m) { using namespace ros::stub; ser(w, m.x); ser(w, m.y); ser(w, m.score); ser(w, m.cov); }
inline void de(ros::stub::Reader& r, Keypoint2D& m) { using namespace ros::stub; de(r, m.x); de(r, m.y); de(r, m.score); de(r, m.cov); }
inline void ser(ros::stub::Writer& w, const Person2D& m) { using namespace ros::stub; ser(w, m.score); ser(w, m.keypoints); ser(w, m.bbox); }
inline void de(ros::stub::Reader& r, Person2D& m) { using namespace ros::stub; de(r, m.score); de(r, m.keypoints); de(r, m.bbox); }
inline void ser(ros::stub::Writer& w, const Person2DList& m) { using namespace ros::stub; ser(w, m.header); ser(w, m.fb_delay); ser(w, m.persons); }
inline void de(ros::stub::Reader& r, Person2DList& m) { using namespace ros::stub; de(r, m.header); de(r, m.fb_delay); de(r, m.persons); }
inline void ser(ros::stub::Writer& w, const KeypointWithCovariance& m) { using namespace ros::stub; ser(w, m.joint); ser(w, m.score); ser(w, m.cov); }
inline void de(ros::stub::Reader& r, KeypointWithCovariance& m) { using namespace ros::stub; de(r, m.joint); de(r, m.score); de(r, m.cov); }
inline void ser(ros::stub::Writer& w, const PersonCov& m) {
  using namespace ros::stub;
  ser(w, m.id); ser(w, m.score); ser(w, m.keypoints); ser(w, m.bbox_center); ser(w, m.bbox_size);
}
inline void de(ros::stub::Reader& r, PersonCov& m) {
  using namespace ros::stub;
  de(r, m.id); de(r, m.score); de(r, m.keypoints); de(r, m.bbox_center); de(r, m.bbox_size);
}
inline void ser(ros::stub::Writer& w, const PersonCovList& m) {
  using namespace ros::stub;
  ser(w, m.header); ser(w, m.ts_per_cam); ser(w, m.fb_delay_per_cam); ser(w, m.persons);
}
inline void de(ros::stub::Reader& r, PersonCovList& m) {
  using namespace ros::stub;
  de(r, m.header); de(r, m.ts_per_cam); de(r, m.fb_delay_per_cam); de(r, m.persons);
}
}  // namespace person_msgs

// ------------------------------------------------------------------------------------------------ tf2
namespace tf2 {
struct TransformException : std::runtime_error {
  explicit TransformException(const std::string& m) : std::runtime_error(m) {}
};
}  // namespace tf2
namespace tf2_ros {
class Buffer {
 public:
  geometry_msgs::TransformStamped lookupTransform(const std::string& target, const std::string& source,
                                                  const ros::Time&) const {
    for (const ros::stub::TfEntry& e : ros::stub::Master::get().tf)
      if (e.target == target && e.source == source) {
        geometry_msgs::TransformStamped t;
        t.header.frame_id = target;
        t.child_frame_id = source;
        t.transform.translation.x = e.t[0]; t.transform.translation.y = e.t[1]; t.transform.translation.z = e.t[2];
        t.transform.rotation.x = e.q[0]; t.transform.rotation.y = e.q[1]; t.transform.rotation.z = e.q[2];
        t.transform.rotation.w = e.q[3];
        return t;
      }
    throw tf2::TransformException("\"" + target + "\" passed to lookupTransform argument target_frame does not exist.");
  }
};
class TransformListener {
 public:
  explicit TransformListener(Buffer&) {}
};
}  // namespace tf2_ros

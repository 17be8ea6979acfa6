// wire.cpp — ROS-free (de)serialisation of the person_msgs wire format (host side).
//
// ROS 1 serialisation of the reference's messages (person_msgs/msg/*.msg; std_msgs/Header, geometry_msgs/Point,
// Pose, Vector3): little-endian primitives, `time` = u32 secs + u32 nsecs, strings and variable-length arrays
// carry a u32 length prefix, fixed-size arrays do not.
//   Keypoint2D             f32 x, y, score, f32[3] cov                                  24 B
//   Person2D               f32 score, Keypoint2D[] (u32 n), f32[4] bbox                 4 + 4 + 24 n + 16
//   Person2DList           Header (u32 seq, time stamp, string frame_id), f32 fb_delay, Person2D[]
//   KeypointWithCovariance Point (3 f64), f32 score, f64[6] cov                         76 B
//   PersonCov              u32 id, f32 score, KeypointWithCovariance[], Pose (7 f64), Vector3 (3 f64)
//   PersonCovList          Header, time[] ts_per_cam, f32[] fb_delay_per_cam, PersonCov[]
// This lets recorded frames (rosbag message payloads, TCPROS bodies) be replayed through the batch ABI and
// results be written back without ROS. The ABI PODs have fixed 17 / 21 keypoints: a Person2D with another
// count is rejected, a PersonCov with another count decodes to an empty skeleton (the reference skips it,
// REP:166-169).
#include <cstdint>
#include <cstring>

#include "ses3d.h"

namespace {

struct Reader {
  const uint8_t* p;
  size_t left;
  bool ok = true;
  template <class T> T get() {
    T v{};
    if (left < sizeof(T)) { ok = false; left = 0; return v; }
    std::memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    left -= sizeof(T);
    return v;
  }
  void skip(size_t n) {
    if (left < n) { ok = false; left = 0; return; }
    p += n;
    left -= n;
  }
};

struct Writer {  // counts when buf == nullptr or capacity is exceeded
  uint8_t* buf;
  size_t cap, used = 0;
  template <class T> void put(T v) {
    if (buf && used + sizeof(T) <= cap) std::memcpy(buf + used, &v, sizeof(T));
    used += sizeof(T);
  }
  void bytes(const void* src, size_t n) {
    if (buf && used + n <= cap) std::memcpy(buf + used, src, n);
    used += n;
  }
};

bool read_header(Reader& r, uint32_t* seq, int64_t* stamp_ns, char* frame_id, size_t frame_id_cap) {
  const uint32_t s = r.get<uint32_t>();
  const uint32_t sec = r.get<uint32_t>(), nsec = r.get<uint32_t>();
  const uint32_t flen = r.get<uint32_t>();
  if (!r.ok || r.left < flen) return false;
  if (seq) *seq = s;
  if (stamp_ns) *stamp_ns = (int64_t)sec * 1000000000LL + nsec;
  if (frame_id && frame_id_cap > 0) {
    const size_t n = flen < frame_id_cap - 1 ? flen : frame_id_cap - 1;
    std::memcpy(frame_id, r.p, n);
    frame_id[n] = 0;
  }
  r.skip(flen);
  return r.ok;
}

void write_header(Writer& w, uint32_t seq, int64_t stamp_ns, const char* frame_id) {
  w.put<uint32_t>(seq);
  w.put<uint32_t>((uint32_t)(stamp_ns / 1000000000LL));
  w.put<uint32_t>((uint32_t)(stamp_ns % 1000000000LL));
  const uint32_t flen = frame_id ? (uint32_t)std::strlen(frame_id) : 0u;
  w.put<uint32_t>(flen);
  if (flen) w.bytes(frame_id, flen);
}

}  // namespace

extern "C" {

int ses3d_wire_decode_person2dlist(const uint8_t* buf, size_t len, uint32_t* seq, int64_t* stamp_ns, char* frame_id,
                                   size_t frame_id_cap, float* fb_delay, ses3d_person2d* persons, int32_t cap) {
  if (!buf || (cap > 0 && !persons)) return SES3D_E_INVALID;
  Reader r{buf, len};
  if (!read_header(r, seq, stamp_ns, frame_id, frame_id_cap)) return SES3D_E_INVALID;
  const float fb = r.get<float>();
  const uint32_t n = r.get<uint32_t>();
  if (!r.ok) return SES3D_E_INVALID;
  if (fb_delay) *fb_delay = fb;
  for (uint32_t i = 0; i < n; ++i) {
    ses3d_person2d p;
    std::memset(&p, 0, sizeof(p));
    p.score = r.get<float>();
    const uint32_t nk = r.get<uint32_t>();
    if (!r.ok || nk != SES3D_NUM_KEYPOINTS || r.left < (size_t)nk * 24 + 16) return SES3D_E_INVALID;
    std::memcpy(p.keypoints, r.p, (size_t)nk * 24);   // Keypoint2D is 6 packed f32, identical to the POD
    r.skip((size_t)nk * 24);
    std::memcpy(p.bbox, r.p, 16);
    r.skip(16);
    if ((int32_t)i < cap) persons[i] = p;
  }
  if (!r.ok) return SES3D_E_INVALID;
  return (int)n;   // may exceed cap: the caller sees how many there were
}

size_t ses3d_wire_encode_person2dlist(uint32_t seq, int64_t stamp_ns, const char* frame_id, float fb_delay,
                                      const ses3d_person2d* persons, int32_t n, uint8_t* buf, size_t cap) {
  Writer w{buf, cap};
  write_header(w, seq, stamp_ns, frame_id);
  w.put<float>(fb_delay);
  w.put<uint32_t>((uint32_t)(n < 0 ? 0 : n));
  for (int32_t i = 0; i < n; ++i) {
    w.put<float>(persons[i].score);
    w.put<uint32_t>(SES3D_NUM_KEYPOINTS);
    w.bytes(persons[i].keypoints, sizeof(persons[i].keypoints));
    w.bytes(persons[i].bbox, 16);
  }
  return w.used;
}

int ses3d_wire_decode_personcovlist(const uint8_t* buf, size_t len, uint32_t* seq, int64_t* stamp_ns, char* frame_id,
                                    size_t frame_id_cap, int64_t* ts_per_cam_ns, float* fb_delay_per_cam,
                                    int32_t cam_cap, int32_t* n_cams, ses3d_person_cov* persons, int32_t cap) {
  if (!buf || (cap > 0 && !persons)) return SES3D_E_INVALID;
  Reader r{buf, len};
  if (!read_header(r, seq, stamp_ns, frame_id, frame_id_cap)) return SES3D_E_INVALID;
  const uint32_t nt = r.get<uint32_t>();
  if (!r.ok || r.left < (size_t)nt * 8) return SES3D_E_INVALID;   // counts come from untrusted bytes: check before looping
  for (uint32_t i = 0; i < nt; ++i) {
    const uint32_t sec = r.get<uint32_t>(), nsec = r.get<uint32_t>();
    if (ts_per_cam_ns && (int32_t)i < cam_cap) ts_per_cam_ns[i] = (int64_t)sec * 1000000000LL + nsec;
  }
  const uint32_t nf = r.get<uint32_t>();
  if (!r.ok || r.left < (size_t)nf * 4) return SES3D_E_INVALID;
  for (uint32_t i = 0; i < nf; ++i) {
    const float v = r.get<float>();
    if (fb_delay_per_cam && (int32_t)i < cam_cap) fb_delay_per_cam[i] = v;
  }
  const uint32_t n = r.get<uint32_t>();
  if (!r.ok) return SES3D_E_INVALID;
  if (n_cams) *n_cams = (int32_t)nt;   // only once the header part has been validated
  for (uint32_t i = 0; i < n; ++i) {
    ses3d_person_cov p;
    std::memset(&p, 0, sizeof(p));
    p.id = r.get<uint32_t>();
    p.score = r.get<float>();
    const uint32_t nk = r.get<uint32_t>();
    if (!r.ok || r.left < (size_t)nk * 76) return SES3D_E_INVALID;
    for (uint32_t k = 0; k < nk; ++k) {
      const double x = r.get<double>(), y = r.get<double>(), z = r.get<double>();
      const float sc = r.get<float>();
      double cv[6];
      for (double& c : cv) c = r.get<double>();
      if (nk == SES3D_NUM_FUSION_KEYPOINTS) {   // any other count: skeleton stays empty (REP:166-169)
        ses3d_keypoint_cov& o = p.keypoints[k];
        o.x = x; o.y = y; o.z = z; o.score = sc;
        for (int c = 0; c < 6; ++c) o.cov[c] = cv[c];
      }
    }
    for (double& v : p.bbox_center) v = r.get<double>();
    for (double& v : p.bbox_size) v = r.get<double>();
    if (!r.ok) return SES3D_E_INVALID;
    if ((int32_t)i < cap) persons[i] = p;
  }
  return (int)n;
}

size_t ses3d_wire_encode_personcovlist(uint32_t seq, int64_t stamp_ns, const char* frame_id, int32_t n_cams,
                                       const int64_t* ts_per_cam_ns, const float* fb_delay_per_cam,
                                       const ses3d_person_cov* persons, int32_t n, uint8_t* buf, size_t cap) {
  Writer w{buf, cap};
  write_header(w, seq, stamp_ns, frame_id);
  const uint32_t nc = (uint32_t)(n_cams < 0 ? 0 : n_cams);
  w.put<uint32_t>(ts_per_cam_ns ? nc : 0u);
  if (ts_per_cam_ns)
    for (uint32_t i = 0; i < nc; ++i) {
      w.put<uint32_t>((uint32_t)(ts_per_cam_ns[i] / 1000000000LL));
      w.put<uint32_t>((uint32_t)(ts_per_cam_ns[i] % 1000000000LL));
    }
  w.put<uint32_t>(fb_delay_per_cam ? nc : 0u);
  if (fb_delay_per_cam)
    for (uint32_t i = 0; i < nc; ++i) w.put<float>(fb_delay_per_cam[i]);
  w.put<uint32_t>((uint32_t)(n < 0 ? 0 : n));
  for (int32_t i = 0; i < n; ++i) {
    const ses3d_person_cov& p = persons[i];
    w.put<uint32_t>(p.id);
    w.put<float>(p.score);
    w.put<uint32_t>(SES3D_NUM_FUSION_KEYPOINTS);
    for (int k = 0; k < SES3D_NUM_FUSION_KEYPOINTS; ++k) {
      const ses3d_keypoint_cov& kp = p.keypoints[k];
      w.put<double>(kp.x); w.put<double>(kp.y); w.put<double>(kp.z);
      w.put<float>(kp.score);
      for (int c = 0; c < 6; ++c) w.put<double>(kp.cov[c]);
    }
    for (double v : p.bbox_center) w.put<double>(v);
    for (double v : p.bbox_size) w.put<double>(v);
  }
  return w.used;
}

}  // extern "C"

// ROS-free decoding of a sensor_msgs/PointCloud2-shaped buffer into what kicp_register_frame consumes:
//   * x/y/z        -> a kicp_frame_input that points INTO the message (float32 fields at point_step; widened on the device,
//                     replacing PointCloud2ToEigen, ros/src/kinematic_icp_ros/utils/RosUtils.cpp:30-39)
//   * per-point t  -> seconds as double, with the field choice, the supported types and the nanosecond heuristic of
//                     utils/TimeStampHandler.cpp:42-104
//   * scan timing  -> begin/end stamp of the sweep and the stamps normalised to [0,1] (TimeStampHandler.cpp:108-139), in double
//                     seconds instead of rclcpp::Time (the reference rounds the sweep duration to nanoseconds)
// Header-only, host-only; a replay driver or a ROS node fills a PointCloud2View from its message and needs nothing else.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "kicp.h"

namespace kicp {

// sensor_msgs/msg/PointField datatype constants
enum : uint8_t { PF_INT8 = 1, PF_UINT8 = 2, PF_INT16 = 3, PF_UINT16 = 4, PF_INT32 = 5, PF_UINT32 = 6, PF_FLOAT32 = 7, PF_FLOAT64 = 8 };

struct PointFieldView {
    std::string name;
    uint32_t offset = 0;
    uint8_t datatype = 0;
    uint32_t count = 0;
};

struct PointCloud2View {
    const uint8_t *data = nullptr;
    uint32_t width = 0, height = 0, point_step = 0;
    std::vector<PointFieldView> fields;
    size_t num_points() const { return static_cast<size_t>(width) * height; }
};

namespace detail {
inline const PointFieldView &field_named(const PointCloud2View &m, const char *name) {
    for (const auto &f : m.fields)
        if (f.name == name) return f;
    throw std::runtime_error(std::string("PointCloud2: no field '") + name + "'");
}
}  // namespace detail

// The message's x, y, z as the device ingest sees them.  Like sensor_msgs::PointCloud2ConstIterator<float>(msg, "x") the
// fields must be FLOAT32; the three offsets need not be contiguous or aligned.
inline kicp_frame_input make_frame_input(const PointCloud2View &m) {
    const PointFieldView &x = detail::field_named(m, "x"), &y = detail::field_named(m, "y"), &z = detail::field_named(m, "z");
    if (x.datatype != PF_FLOAT32 || y.datatype != PF_FLOAT32 || z.datatype != PF_FLOAT32)
        throw std::runtime_error("PointCloud2: x/y/z must be FLOAT32");
    if (m.point_step == 0 || x.offset + 4 > m.point_step || y.offset + 4 > m.point_step || z.offset + 4 > m.point_step)
        throw std::runtime_error("PointCloud2: x/y/z outside point_step");
    kicp_frame_input in{};
    in.data = m.data, in.n = static_cast<int64_t>(m.num_points()), in.dtype = KICP_DTYPE_F32;
    in.point_step = static_cast<int32_t>(m.point_step);
    in.offset_x = static_cast<int32_t>(x.offset), in.offset_y = static_cast<int32_t>(y.offset), in.offset_z = static_cast<int32_t>(z.offset);
    return in;
}

// Per-point acquisition times in seconds; empty when the message carries none (de-skewing is then disabled).  The LAST field
// named t / timestamp / time / stamps is used; UINT32, FLOAT32 and FLOAT64 are supported; a value whose integer part has more
// than 10 digits is taken to be nanoseconds.
inline std::vector<double> extract_timestamps(const PointCloud2View &m) {
    const PointFieldView *tf = nullptr;
    for (const auto &f : m.fields)
        if (f.name == "t" || f.name == "timestamp" || f.name == "time" || f.name == "stamps") tf = &f;
    if (tf == nullptr || tf->count == 0) return {};
    const size_t width = tf->datatype == PF_FLOAT64 ? 8 : 4;
    if (tf->datatype != PF_UINT32 && tf->datatype != PF_FLOAT32 && tf->datatype != PF_FLOAT64)
        throw std::runtime_error("timestamp field type not supported");
    if (tf->offset + width > m.point_step) throw std::runtime_error("PointCloud2: timestamp field outside point_step");
    std::vector<double> stamps;
    stamps.reserve(m.num_points());
    for (size_t i = 0; i < m.num_points(); ++i) {
        const uint8_t *p = m.data + i * m.point_step + tf->offset;
        double s;
        if (tf->datatype == PF_UINT32) {
            uint32_t v;
            std::memcpy(&v, p, 4);
            s = static_cast<double>(v);
        } else if (tf->datatype == PF_FLOAT32) {
            float v;
            std::memcpy(&v, p, 4);
            s = static_cast<double>(v);
        } else {
            std::memcpy(&s, p, 8);
        }
        const uint64_t whole = static_cast<uint64_t>(std::round(s));
        const double digits = whole > 0 ? std::floor(std::log10(static_cast<double>(whole)) + 1.0) : 1.0;
        if (digits > 10.0) s *= 1e-9;
        stamps.push_back(s);
    }
    return stamps;
}

struct SweepTiming {
    double begin = 0.0, end = 0.0;  // seconds: the interval the wheel odometry is looked up for
};

// Normalises `timestamps` in place to [0,1] and returns the sweep interval: it begins where the previous sweep ended
// (`last_processed_stamp`, updated here) and ends at the header stamp — plus the sweep duration when the header stamps the
// beginning of the scan (|header - max| > 1e-8).
inline SweepTiming process_timestamps(std::vector<double> &timestamps, double header_stamp, double &last_processed_stamp) {
    SweepTiming t;
    t.begin = last_processed_stamp;
    t.end = header_stamp;
    if (!timestamps.empty()) {
        const auto mm = std::minmax_element(timestamps.cbegin(), timestamps.cend());
        const double lo = *mm.first, hi = *mm.second;
        if (std::abs(header_stamp - hi) > 1e-8) t.end = header_stamp + (hi - lo);
        for (double &s : timestamps) s = (s - lo) / (hi - lo);
    }
    last_processed_stamp = t.end;
    return t;
}

}  // namespace kicp

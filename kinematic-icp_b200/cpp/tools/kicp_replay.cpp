// kicp_replay — offline replay of a recorded drive through kinematic_icp::pipeline::KinematicICP, the way
// ros/src/kinematic_icp_ros/nodes/offline_node.cpp:99-149 replays a bag: for every frame RegisterFrame(cloud, stamps,
// lidar_to_base, relative wheel odometry), collect the poses, write them in TUM format (offline_node.cpp:76-97).
// There is no ROS here, so the input is a flat ".kseq" file (below) instead of a bag: float32 x,y,z per point exactly as a
// PointCloud2 message carries them (ros/.../utils/RosUtils.cpp:30-39), handed to the facade without widening on the host.
//
//   kicp_replay drive.kseq out.tum [--deskew 0|1] [--voxel-size v] [--max-range r] [--min-range r] [--repeat k] [--pageable]
//
// .kseq, little endian:  char magic[8] = "KSEQ1\0\0\0"; int32 n_frames; int32 reserved; double lidar_to_base[7]; double start_pose[7];
//   per frame: int32 n_points; int32 has_stamps; double header_stamp; double relative_odometry[7]; float xyz[3 n]; double stamps[n] (if has_stamps)
// (poses are {qx, qy, qz, qw, tx, ty, tz}).  Prints one JSON line with the frame rate; exits non-zero on any error (no CUDA
// device included: there is no CPU fallback).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "kicp/facade_pipeline.hpp"
#include "kicp/runtime.hpp"
#include "kicp/tum.hpp"

namespace {
struct Frame {
    int32_t n = 0, has_stamps = 0;
    double header_stamp = 0.0;
    double odom[7];
    float *xyz = nullptr;      // n * 3, page-locked unless --pageable
    double *stamps = nullptr;  // n
};

void *host_alloc(size_t bytes, bool pinned) {
    if (bytes == 0) bytes = 8;
    void *p = nullptr;
    if (pinned) {
        kicp::check(kicp_host_alloc((uint64_t)bytes, &p), "kicp_host_alloc");
    } else {
        p = std::malloc(bytes);
        if (!p) throw std::runtime_error("out of host memory");
    }
    return p;
}

template <class T>
void read_exact(std::FILE *f, T *dst, size_t count, const char *what) {
    if (std::fread(dst, sizeof(T), count, f) != count) throw std::runtime_error(std::string("truncated .kseq file while reading ") + what);
}
}  // namespace

int main(int argc, char **argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s drive.kseq out.tum [--deskew 0|1] [--voxel-size v] [--max-range r] [--min-range r] [--repeat k] [--pageable]\n", argv[0]);
        return 2;
    }
    kinematic_icp::pipeline::Config config;  // the reference's defaults (pipeline/KinematicICP.hpp:38-60)
    config.deskew = true;
    int repeat = 1;
    bool pinned = true;
    for (int i = 3; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&](const char *name) -> const char * {
            if (i + 1 >= argc) throw std::runtime_error(std::string("missing value for ") + name);
            return argv[++i];
        };
        try {
            if (a == "--deskew") config.deskew = std::atoi(value("--deskew")) != 0;
            else if (a == "--voxel-size") config.voxel_size = std::atof(value("--voxel-size"));
            else if (a == "--max-range") config.max_range = std::atof(value("--max-range"));
            else if (a == "--min-range") config.min_range = std::atof(value("--min-range"));
            else if (a == "--repeat") repeat = std::max(1, std::atoi(value("--repeat")));
            else if (a == "--pageable") pinned = false;
            else throw std::runtime_error("unknown option " + a);
        } catch (const std::exception &e) {
            std::fprintf(stderr, "kicp_replay: %s\n", e.what());
            return 2;
        }
    }
    try {
        kicp::default_context();  // fails here, loudly, without a CUDA device
        std::FILE *f = std::fopen(argv[1], "rb");
        if (!f) throw std::runtime_error(std::string("cannot open ") + argv[1]);
        char magic[8];
        read_exact(f, magic, 8, "the header");
        if (std::memcmp(magic, "KSEQ1\0\0\0", 8) != 0) throw std::runtime_error("not a .kseq file (bad magic)");
        int32_t n_frames = 0, reserved = 0;
        double l2b[7], start[7];
        read_exact(f, &n_frames, 1, "the header"), read_exact(f, &reserved, 1, "the header");
        read_exact(f, l2b, 7, "lidar_to_base"), read_exact(f, start, 7, "the start pose");
        if (n_frames < 0 || n_frames > (1 << 24)) throw std::runtime_error("implausible frame count");
        std::vector<Frame> frames((size_t)n_frames);
        int64_t total_points = 0;
        for (auto &fr : frames) {
            read_exact(f, &fr.n, 1, "a frame header"), read_exact(f, &fr.has_stamps, 1, "a frame header");
            read_exact(f, &fr.header_stamp, 1, "a frame header"), read_exact(f, fr.odom, 7, "a frame header");
            if (fr.n < 0) throw std::runtime_error("negative point count");
            fr.xyz = static_cast<float *>(host_alloc((size_t)fr.n * 3 * sizeof(float), pinned));
            read_exact(f, fr.xyz, (size_t)fr.n * 3, "points");
            if (fr.has_stamps) {
                fr.stamps = static_cast<double *>(host_alloc((size_t)fr.n * sizeof(double), pinned));
                read_exact(f, fr.stamps, (size_t)fr.n, "stamps");
            }
            total_points += fr.n;
        }
        std::fclose(f);

        const Sophus::SE3d lidar_to_base = kicp::from_pose7(l2b);
        std::vector<std::pair<double, Sophus::SE3d>> poses;
        double seconds = 0.0;
        std::vector<double> rep_seconds;  // every repetition's wall time (a 24-frame drive lasts milliseconds: one hiccup halves a single figure)
        for (int rep = 0; rep < repeat; ++rep) {  // --repeat: the same drive again on a fresh pipeline (the last run is the one written)
            kinematic_icp::pipeline::KinematicICP pipeline(config);
            pipeline.SetPose(kicp::from_pose7(start));
            poses.clear();
            poses.reserve(frames.size());
            const auto t0 = std::chrono::steady_clock::now();
            for (const auto &fr : frames) {
                kicp_frame_input in{};
                in.data = fr.xyz, in.n = fr.n, in.dtype = KICP_DTYPE_F32, in.point_step = 0;
                in.stamps = fr.has_stamps ? fr.stamps : nullptr, in.n_stamps = fr.has_stamps ? fr.n : 0;
                pipeline.RegisterFrame(in, lidar_to_base, kicp::from_pose7(fr.odom));
                poses.emplace_back(fr.header_stamp, pipeline.pose());
            }
            seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            rep_seconds.push_back(seconds);
        }
        // the fastest repetition after the first (which also pays for allocations); the last one stays the headline of this line
        double best = seconds;
        for (size_t k = rep_seconds.size() > 1 ? 1 : 0; k < rep_seconds.size(); ++k) best = std::min(best, rep_seconds[k]);
        std::string reps = "[";
        for (size_t k = 0; k < rep_seconds.size(); ++k) {
            char buf[32];
            std::snprintf(buf, sizeof(buf), "%s%.6f", k ? ", " : "", rep_seconds[k]);
            reps += buf;
        }
        reps += "]";
        if (!kicp::write_poses_tum(argv[2], poses)) throw std::runtime_error(std::string("cannot write ") + argv[2]);
        std::printf("{\"harness\": \"kicp_replay\", \"frames\": %d, \"points_per_frame\": %.0f, \"seconds\": %.6f, \"frames_per_s\": %.2f, "
                    "\"ms_per_frame\": %.4f, \"frames_per_s_best\": %.2f, \"repetition_seconds\": %s, \"host_buffers\": \"%s float32\", "
                    "\"deskew\": %s, \"tum_file\": \"%s\"}\n",
                    n_frames, n_frames ? (double)total_points / n_frames : 0.0, seconds, seconds > 0 ? n_frames / seconds : 0.0,
                    n_frames ? 1e3 * seconds / n_frames : 0.0, best > 0 ? n_frames / best : 0.0, reps.c_str(), pinned ? "pinned" : "pageable",
                    config.deskew ? "true" : "false", argv[2]);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "kicp_replay: %s\n", e.what());
        return 1;
    }
}

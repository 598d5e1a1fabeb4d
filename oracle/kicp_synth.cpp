// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Deterministic synthetic world + spinning-lidar ray caster used by tests/ and bench.py to produce
// the scans and maps SURVEY.md §8(d) / BASELINE.md §2 describe.  Analytic warehouse scene: ground plane z = 0, ceiling plane z = 1.3 wall_h,
// a square outer wall loop and a lattice of pillars and short wall segments — everything inside
// max_range of the trajectory.  Range noise is Gaussian along the ray; every output coordinate can
// be rounded to a float32-representable double (what a PointCloud2 float32 field would carry,
// ros/src/kinematic_icp_ros/utils/RosUtils.cpp:30-39).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct Box {
    double lo[3], hi[3];
};

struct Scene {
    std::vector<Box> boxes;
    // uniform 2-D grid over the xy footprint for ray traversal
    double g0 = 0, cell = 1;
    int gn = 0;
    std::vector<std::vector<int>> cells;
    double key[3] = {-1, -1, -1};
};

// half_extent: outer wall loop at +-half_extent; pillars every `pitch` metres.
void make_scene(Scene &s, double half_extent, double pitch, double wall_h) {
    s.boxes.clear();
    const double E = half_extent, th = 0.4;
    // outer wall loop
    s.boxes.push_back({{-E - th, -E - th, 0}, {E + th, -E, wall_h}});
    s.boxes.push_back({{-E - th, E, 0}, {E + th, E + th, wall_h}});
    s.boxes.push_back({{-E - th, -E, 0}, {-E, E, wall_h}});
    s.boxes.push_back({{E, -E, 0}, {E + th, E, wall_h}});
    // lattice of obstacles, leaving a ring corridor |r - 50 m| < 3 m free for the trajectory.  Two in three are
    // pillars (1.2 m x 0.8 m footprint, three heights), every third is a wall segment (0.3 m x 0.6*pitch).
    int idx = 0;
    for (double x = -E + pitch * 0.5; x < E; x += pitch) {
        for (double y = -E + pitch * 0.5; y < E; y += pitch) {
            ++idx;
            const double r = std::sqrt(x * x + y * y);
            if (std::fabs(r - 50.0) < 3.0 + 0.5 * pitch) continue;
            const double h = (idx % 3 == 0) ? wall_h : (idx % 3 == 1 ? 0.6 * wall_h : 0.35 * wall_h);
            const double ox = 0.37 * ((idx * 7) % 5 - 2), oy = 0.29 * ((idx * 11) % 5 - 2);
            if (idx % 3 == 2) {
                const double hl = 0.3 * pitch;
                if ((idx / 3) % 2)
                    s.boxes.push_back({{x - hl, y + oy - 0.15, 0}, {x + hl, y + oy + 0.15, 0.8 * wall_h}});
                else
                    s.boxes.push_back({{x + ox - 0.15, y - hl, 0}, {x + ox + 0.15, y + hl, 0.8 * wall_h}});
            } else {
                s.boxes.push_back({{x + ox - 0.6, y + oy - 0.4, 0}, {x + ox + 0.6, y + oy + 0.4, h}});
            }
        }
    }
    s.cell = pitch;
    s.g0 = -E - 2.0;
    s.gn = static_cast<int>(std::ceil((2.0 * E + 4.0) / s.cell));
    s.cells.assign(static_cast<size_t>(s.gn) * s.gn, {});
    for (size_t b = 0; b < s.boxes.size(); ++b) {
        const Box &bx = s.boxes[b];
        const int x0 = std::max(0, static_cast<int>(std::floor((bx.lo[0] - s.g0) / s.cell)));
        const int x1 = std::min(s.gn - 1, static_cast<int>(std::floor((bx.hi[0] - s.g0) / s.cell)));
        const int y0 = std::max(0, static_cast<int>(std::floor((bx.lo[1] - s.g0) / s.cell)));
        const int y1 = std::min(s.gn - 1, static_cast<int>(std::floor((bx.hi[1] - s.g0) / s.cell)));
        for (int cy = y0; cy <= y1; ++cy)
            for (int cx = x0; cx <= x1; ++cx) s.cells[static_cast<size_t>(cy) * s.gn + cx].push_back(static_cast<int>(b));
    }
    s.key[0] = half_extent, s.key[1] = pitch, s.key[2] = wall_h;
}

inline bool ray_box(const double o[3], const double inv[3], const Box &b, double tmax, double *t_hit) {
    double t0 = 0.0, t1 = tmax;
    for (int a = 0; a < 3; ++a) {
        double ta = (b.lo[a] - o[a]) * inv[a], tb = (b.hi[a] - o[a]) * inv[a];
        if (ta > tb) std::swap(ta, tb);
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
        if (t0 > t1) return false;
    }
    *t_hit = t0;
    return t0 > 0.0;
}

// nearest box hit along the ray, < best; 2-D DDA over the grid (Amanatides-Woo)
double cast(const Scene &s, const double o[3], const double d[3], double best) {
    const double inv[3] = {1.0 / d[0], 1.0 / d[1], 1.0 / d[2]};
    int cx = static_cast<int>(std::floor((o[0] - s.g0) / s.cell)), cy = static_cast<int>(std::floor((o[1] - s.g0) / s.cell));
    if (cx < 0 || cy < 0 || cx >= s.gn || cy >= s.gn) return best;
    const int sx = d[0] > 0 ? 1 : -1, sy = d[1] > 0 ? 1 : -1;
    const double big = 1e300;
    double tmx = std::fabs(d[0]) < 1e-15 ? big : ((s.g0 + (cx + (sx > 0 ? 1 : 0)) * s.cell) - o[0]) * inv[0];
    double tmy = std::fabs(d[1]) < 1e-15 ? big : ((s.g0 + (cy + (sy > 0 ? 1 : 0)) * s.cell) - o[1]) * inv[1];
    const double tdx = std::fabs(d[0]) < 1e-15 ? big : s.cell * std::fabs(inv[0]);
    const double tdy = std::fabs(d[1]) < 1e-15 ? big : s.cell * std::fabs(inv[1]);
    while (true) {
        for (int b : s.cells[static_cast<size_t>(cy) * s.gn + cx]) {
            double t;
            if (ray_box(o, inv, s.boxes[b], best, &t) && t < best) best = t;
        }
        const double t_exit = std::min(tmx, tmy);
        if (best <= t_exit) return best;
        if (tmx < tmy) {
            cx += sx, tmx += tdx;
        } else {
            cy += sy, tmy += tdy;
        }
        if (cx < 0 || cy < 0 || cx >= s.gn || cy >= s.gn) return best;
    }
}

struct Rng {  // splitmix64 -> uniform -> Box-Muller
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (static_cast<double>(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
    double gauss() {
        const double u1 = uniform(), u2 = uniform();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2);
    }
};

}  // namespace

extern "C" {

// Ray-cast one scan.  Sensor sits at (x, y, sensor_h) with heading yaw in the world; output points are in
// the robot BASE frame (base origin on the ground below the sensor, x forward).  Azimuth-major order.
// Returns the number of points written (rays with no hit inside max_range are dropped).
int64_t kor_synth_scan(double half_extent, double pitch, double wall_h, int n_beams, double elev_min_deg,
                       double elev_max_deg, int n_az, double x, double y, double yaw, double sensor_h, double sigma,
                       uint64_t seed, double max_range, int round_f32, double *out_xyz, int64_t cap) {
    static Scene scene;
    if (scene.key[0] != half_extent || scene.key[1] != pitch || scene.key[2] != wall_h)
        make_scene(scene, half_extent, pitch, wall_h);
    Rng rng{seed};
    const double o[3] = {x, y, sensor_h};
    const double cy = std::cos(yaw), sy = std::sin(yaw);
    int64_t n = 0;
    const double pi = 3.14159265358979323846, deg = pi / 180.0;
    for (int ia = 0; ia < n_az; ++ia) {
        const double az = 2.0 * pi * ia / n_az;
        const double ca = std::cos(az), sa = std::sin(az);
        for (int ib = 0; ib < n_beams; ++ib) {
            const double el =
                (n_beams == 1 ? 0.5 * (elev_min_deg + elev_max_deg)
                              : elev_min_deg + (elev_max_deg - elev_min_deg) * ib / (n_beams - 1)) * deg;
            const double ce = std::cos(el), se = std::sin(el);
            const double dl[3] = {ce * ca, ce * sa, se};                                    // base axes
            const double d[3] = {cy * dl[0] - sy * dl[1], sy * dl[0] + cy * dl[1], dl[2]};  // world
            const double noise = rng.gauss();  // drawn for every ray so the stream is layout-independent
            double best = max_range;
            if (d[2] < -1e-12) best = std::min(best, -o[2] / d[2]);               // ground z = 0
            if (d[2] > 1e-12) best = std::min(best, (1.3 * wall_h - o[2]) / d[2]);  // ceiling z = 1.3 wall_h
            best = cast(scene, o, d, best);
            if (best >= max_range) continue;
            const double r = best + sigma * noise;
            if (!(r > 0.3) || r >= max_range) continue;
            if (n >= cap) return -1;
            double p[3] = {r * dl[0], r * dl[1], r * dl[2] + sensor_h};
            if (round_f32)
                for (int a = 0; a < 3; ++a) p[a] = static_cast<double>(static_cast<float>(p[a]));
            out_xyz[3 * n + 0] = p[0], out_xyz[3 * n + 1] = p[1], out_xyz[3 * n + 2] = p[2];
            ++n;
        }
    }
    return n;
}

}  // extern "C"

#include "disco_limb.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>

namespace disco {
namespace {

constexpr double kPiL = 3.14159265358979323846;

struct Vec {
    double x, y, z;
};
inline Vec operator+(Vec a, Vec b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec operator-(Vec a, Vec b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec operator*(Vec a, double f) { return {a.x * f, a.y * f, a.z * f}; }
inline double dot(Vec a, Vec b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec cross(Vec a, Vec b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double len(Vec a) { return std::sqrt(dot(a, a)); }
inline Vec unit(Vec a) {
    const double n = len(a);
    return {a.x / n, a.y / n, a.z / n};
}
// rotation of v about the unit axis k by `angle` (what Eigen::AngleAxis does upstream)
inline Vec spin(Vec v, Vec k, double angle) {
    const double c = std::cos(angle), s = std::sin(angle);
    return v * c + cross(k, v) * s + k * (dot(k, v) * (1.0 - c));
}

// The reference's local frame (lib/geometry/geometry.cpp:8-23, force_sun_z = false): z = reference point,
// sun = cos_sza z + sin_sza (cos saa x + sin saa y)
struct Frame {
    Vec ex{1, 0, 0}, ey{0, 1, 0}, ez{0, 0, 1}, sun{0, 0, 1};
    double re = 0;
    Frame(double cos_sza, double saa, double earth_radius) : re(earth_radius) {
        const Vec horiz = ex * std::cos(saa) + ey * std::sin(saa);
        sun = ez * cos_sza + horiz * std::sqrt(1.0 - cos_sza * cos_sza);
    }
    // point at `altitude` whose solar zenith cosine is cos_sza (geometry.cpp:155-182, spherical)
    Vec point_at_sza(double cos_sza, double saa, double altitude) const {
        Vec n = cross(sun, ez);
        n = (len(n) == 0.0) ? ey : unit(n);
        Vec v = spin(sun, n, std::acos(cos_sza));
        v = spin(v, sun, saa);
        return v * (altitude + re);
    }
    // look vector at `location` with azimuth saa from the sun and zenith cosine cos_viewing (geometry.cpp:199-232)
    Vec look_from_azimuth(Vec location, double saa, double cos_viewing) const {
        const Vec up = unit(location);
        Vec sh = sun - up * dot(up, sun);
        if (len(sh) == 0.0) sh = ey;
        sh = unit(sh);
        const Vec horiz = spin(sh, up, -saa);
        const double tilt = kPiL / 2 - std::acos(-cos_viewing);
        return spin(horiz, cross(up, horiz), tilt);
    }
};

// One-dimensional grid lookup with the reference's "extend" rule outside the grid (lib/grids/grid.cpp:43-300).
// spacing_constant as the Grid constructor resolves it (:9-27).
struct Axis {
    std::vector<double> g;
    int interp = 1;  // 0 shell, 1 linear, 2 lower
    bool uniform = false;
    void detect_spacing() {
        uniform = true;
        if (g.size() > 1) {
            const double d0 = g[1] - g[0];
            for (size_t i = 1; i < g.size(); ++i) {
                const double di = g[i] - g[i - 1];
                if (std::abs(di - d0) > 1e-12 * std::min(std::abs(di), std::abs(d0))) uniform = false;
            }
        }
    }
    int lookup(double x, int idx[2], double w[2]) const {
        const int n = (int)g.size();
        auto single = [&](int i) {
            idx[0] = i;
            idx[1] = 0;
            w[0] = 1.0;
            w[1] = 0.0;
            return 1;
        };
        if (n == 1) return single(0);
        if (interp == 2) {
            for (int i = 0; i + 1 < n; ++i)
                if (x + 0.1 >= g[i] && x < g[i + 1]) return single(i);
            return single(x < g[0] ? 0 : n - 2);
        }
        int lo;
        double frac;
        if (uniform) {
            const double x0 = g[0], dx = g[1] - g[0];
            // degenerate grid (the SZA grid of an exactly vertical ray is LinSpaced(n, a, a)): everything on the first entry
            if (!(dx > 0.0) || x < x0) return single(0);
            lo = (int)std::floor((x - x0) / dx);
            if (lo >= n - 1) return single(n - 1);
            frac = (x - g[lo]) / dx;
        } else {
            if (x < g[0]) return single(0);
            if (x > g[n - 1]) return single(n - 1);
            int hi = (int)(std::lower_bound(g.begin(), g.end(), x) - g.begin());
            if (hi == 0) hi = 1;
            lo = hi - 1;
            frac = (x - g[lo]) / (g[hi] - g[lo]);
        }
        idx[0] = lo;
        idx[1] = lo + 1;
        if (interp == 0) {
            w[0] = w[1] = 0.5;
        } else {
            w[1] = frac;
            w[0] = 1.0 - w[1];
        }
        return 2;
    }
};

struct EndPoint {
    Vec pos{0, 0, 0};
    bool exact = false;  // sits on a grid altitude
    int grid = -1;       // that altitude's index
};
struct Seg {
    EndPoint near_pt, far_pt;        // near: the end closer to the observer ("entrance" upstream)
    double r_near = 0, r_far = 0;
    int tangent_end = 0;             // 1: the near end is the tangent point, 2: the far end is
    Vec look{0, 0, 0};
    double length = 0;
    double q_near = 0, q_far = 0, qf_near = 0.5, qf_far = 0.5;   // optical-depth quadrature coefficients and their fractions
    double csz_near = 0, csz_far = 0, saz_near = 0, saz_far = 0;
    int nidx = 0;
    int idx[kLimbStencil] = {0, 0, 0, 0};
    double w_near[kLimbStencil] = {0, 0, 0, 0}, w_far[kLimbStencil] = {0, 0, 0, 0}, w_od[kLimbStencil] = {0, 0, 0, 0};
};
struct Path {
    Vec observer{0, 0, 0}, look{0, 0, 0};
    bool ground = false;
    double rt = 0;
    std::vector<Seg> segs;  // far -> near
};

struct Tracer {
    const std::vector<double>& alt;
    Axis grid;
    Frame frame;
    Tracer(const GeometrySpec& geo) : alt(geo.altitudes), frame(geo.cos_sza, geo.saa, geo.earth_radius) {
        grid.g = geo.altitudes;
        grid.interp = geo.interp;
        grid.detect_spacing();
    }
    double re() const { return frame.re; }

    // altitude stencil of an end point (lib/geometry/geometry1d.cpp:20-84)
    int stencil(const EndPoint& e, int idx[2], double w[2]) const {
        const double a = len(e.pos) - re();
        if (e.exact && e.grid >= 0 && e.grid < (int)alt.size() && std::abs(a - alt[e.grid]) <= 1.0) {
            idx[0] = e.grid;
            idx[1] = 0;
            w[0] = 1.0;
            w[1] = 0.0;
            return 1;
        }
        return grid.lookup(a, idx, w);
    }
    static void solar_angles(Vec sun, Vec pos, Vec look, double& csz, double& saz) {  // raytracing.h:319-349
        const Vec up = unit(pos);
        csz = dot(up, sun);
        const Vec lp = unit(look - up * dot(look, up));
        const Vec sp = unit(sun - up * dot(sun, up));
        const Vec yax = cross(up, sp);
        saz = std::atan2(dot(yax, lp), dot(sp, lp));
    }
    // shell between grid altitudes: `far_index` is the grid index of the far end, the near end is one step `dir` away
    static void full_shell(Seg& s, const std::vector<double>& alt, double re, int far_index, int dir) {
        s.r_near = alt[far_index + dir] + re;
        s.r_far = alt[far_index] + re;
        s.near_pt.exact = true;
        s.near_pt.grid = far_index + dir;
        s.far_pt.exact = true;
        s.far_pt.grid = far_index;
    }
    void tangent_shell(Seg& s, int upper, double tangent_alt, bool far_side) const {  // spherical_shell.cpp:300-346
        int ti = upper - 1;
        bool exact = std::abs(tangent_alt - alt[ti]) <= 1e-4;
        if (!exact && std::abs(tangent_alt - alt[upper]) <= 1e-4) {
            ti = upper;
            exact = true;
        }
        s.tangent_end = far_side ? 1 : 2;
        if (far_side) {  // near end = tangent point, far end = the grid altitude above it
            s.r_near = tangent_alt + re();
            s.r_far = alt[upper] + re();
            s.near_pt.exact = exact;
            s.near_pt.grid = ti;
            s.far_pt.exact = true;
            s.far_pt.grid = upper;
        } else {
            s.r_far = tangent_alt + re();
            s.r_near = alt[upper] + re();
            s.far_pt.exact = exact;
            s.far_pt.grid = ti;
            s.near_pt.exact = true;
            s.near_pt.grid = upper;
        }
    }

    // Straight-ray trace (spherical_shell.cpp:6-76).  Supported: observer above the atmosphere looking down (limb or
    // ground), observer inside the atmosphere looking up (solar rays).
    void trace(Vec observer, Vec look, Path& out, bool solar_ray) const {
        out = Path();
        out.observer = observer;
        out.look = look;
        const int ng = (int)alt.size();
        const double ro = len(observer);
        const double cv = dot(observer, look) / (ro * len(look));
        out.rt = ro * std::sqrt(std::max(0.0, 1.0 - cv * cv));
        const double tangent_alt = out.rt - re(), obs_alt = ro - re();
        if (obs_alt >= alt[ng - 1]) {
            if (cv > 0) return;  // looking away from the atmosphere: empty path
            if (tangent_alt > alt[0]) {
                const int above = (int)(std::upper_bound(alt.begin(), alt.end(), tangent_alt) - alt.begin());
                const int n = 2 * (ng - above);
                out.segs.resize(n);
                if (n == 0) return;
                int c = 0;
                for (int i = ng - 1; i != above; --i) full_shell(out.segs[c++], alt, re(), i, -1);
                tangent_shell(out.segs[c++], above, tangent_alt, true);
                tangent_shell(out.segs[c++], above, tangent_alt, false);
                for (int i = above; i < ng - 1; ++i) full_shell(out.segs[c++], alt, re(), i, +1);
            } else {
                out.ground = true;
                out.segs.resize(ng - 1);
                for (int i = 0; i < ng - 1; ++i) full_shell(out.segs[i], alt, re(), i, +1);
            }
        } else if (cv > 0) {
            int start = (int)(std::upper_bound(alt.begin(), alt.end(), obs_alt) - alt.begin());
            // A start ON a grid altitude (solar rays leave from the boundaries of the line-of-sight layers) is an exact
            // point of the grid: the reference lets the rounding of |position| - R decide between a zero-length partial
            // shell below and a whole-shell "partial" layer with an inexact lower end, which changes that shell's
            // optical depth under shell interpolation (limb_oracle.hpp, trace).  Snapped here within the reference's own
            // 1e-4 m exactness tolerance of tangent points.
            int snap = -1;
            for (int i = 0; i < ng; ++i)
                if (std::abs(obs_alt - alt[i]) <= 1e-4) snap = i;
            if (snap >= 0) start = snap + 1;
            if (start >= ng) {
                out.segs.clear();
                return;
            }
            out.segs.resize(ng - start);
            int c = 0;
            for (int i = ng - 1; i != start; --i) full_shell(out.segs[c++], alt, re(), i, -1);
            if (snap >= 0) {
                full_shell(out.segs[c], alt, re(), start, -1);
            } else {
                Seg& s = out.segs[c];  // partial shell from the observer up to the next grid altitude (:278-298)
                s.r_near = obs_alt + re();
                s.r_far = alt[start] + re();
                s.far_pt.exact = true;
                s.far_pt.grid = start;
                s.near_pt.exact = false;
                s.near_pt.grid = start - 1;
            }
        } else {
            if (tangent_alt <= alt[0]) {
                out.ground = true;  // a blocked solar ray only needs its flag
                if (solar_ray) return;
            }
            throw std::runtime_error(solar_ray ? "B200 limb path: the sun is below the local horizon of a line-of-sight point "
                                                 "(solar rays looking down are not supported)"
                                               : "B200 limb path: the observer must be above the top of the atmosphere");
        }
        finish(out);
    }

    // positions, path lengths, optical-depth quadrature, stencils and solar angles (spherical_shell.cpp:85-205)
    void finish(Path& p) const {
        const int n = (int)p.segs.size(), ng = (int)alt.size();
        const double rt = p.rt;
        for (int i = 0; i < n; ++i) {
            Seg& s = p.segs[n - 1 - i];
            if (i == 0) {
                if (len(p.observer) - re() < alt[ng - 1]) {
                    s.near_pt.pos = p.observer;
                } else {
                    // distance from the observer down to the top of the atmosphere on the near side (raytracing.h:826-866)
                    const double cz = std::abs(dot(p.observer, p.look) / (len(p.observer) * len(p.look)));
                    const double ro = len(p.observer), rr = re() + alt[ng - 1];
                    const double rtsq = ro * ro * (1.0 - cz * cz);
                    double from_tangent;
                    if (rtsq > rr * rr) {
                        if (std::abs(rtsq - rr * rr) < 100)
                            from_tangent = 0.0;
                        else
                            throw std::runtime_error("B200 limb path: the ray misses the atmosphere");
                    } else {
                        from_tangent = std::sqrt(std::abs(rr * rr - rtsq));
                    }
                    s.near_pt.pos = p.observer + p.look * (ro * cz - from_tangent);
                }
            } else {
                s.near_pt.pos = p.segs[n - i].far_pt.pos;
            }
            // Along-ray coordinate of both ends measured from the tangent point.  The reference evaluates
            // sqrt(max(r^2 - rt^2, 0)) for the tangent end too, with r = (rt - R) + R one rounding away from rt: depending
            // on the rounding direction both tangent segments come out ~0.1 m short (up to 3e-7 of a limb optical depth;
            // its own golden limb optical depths are 2.1e-7 off the exact integral).  Here the tangent end is exactly 0.
            const double t_near = s.tangent_end == 1 ? 0.0 : std::sqrt(std::fmax(s.r_near * s.r_near - rt * rt, 0.0));
            const double t_far = s.tangent_end == 2 ? 0.0 : std::sqrt(std::fmax(s.r_far * s.r_far - rt * rt, 0.0));
            s.length = std::abs(t_near - t_far);
            s.far_pt.pos = s.near_pt.pos + p.look * s.length;
            quadrature(s);
            stencils(s);
            solar_angles(frame.sun, s.near_pt.pos, s.look, s.csz_near, s.saz_near);
            solar_angles(frame.sun, s.far_pt.pos, s.look, s.csz_far, s.saz_far);
        }
    }

    // optical depth of the segment = q_near k(near) + q_far k(far) for an extinction linear in radius (raytracing.h:478-560)
    void quadrature(Seg& s) const {
        const double r0 = len(s.near_pt.pos), r1 = len(s.far_pt.pos), dr = r1 - r0;
        s.look = unit(s.far_pt.pos - s.near_pt.pos);
        if (grid.interp == 2) {
            s.q_near = r0 < r1 ? s.length : 0.0;
            s.q_far = r0 < r1 ? 0.0 : s.length;
            s.qf_near = s.qf_far = 0.5;
            return;
        }
        if (std::abs(dr) < 0.001 || grid.interp == 0) {
            s.q_near = s.q_far = s.length / 2;
            s.qf_near = s.qf_far = 0.5;
            return;
        }
        const double c0 = dot(s.near_pt.pos, s.look) / (r0 * len(s.look)), c1 = dot(s.far_pt.pos, s.look) / (r1 * len(s.look));
        const double t0 = r0 * c0, t1 = r1 * c1;
        const double rt = r0 * std::sqrt(1.0 - c0 * c0);
        double dt1, dt2;
        if (t1 >= t0) {
            dt1 = t1 - t0;
            dt2 = std::abs(rt) < 10 ? 0.5 * (r1 * t1 - r0 * t0) : 0.5 * ((r1 * t1 - r0 * t0) + rt * rt * std::log((r1 + t1) / (r0 + t0)));
        } else {
            dt1 = t0 - t1;
            dt2 = std::abs(rt) < 10 ? 0.5 * (r0 + t0 - r1 * t1) : 0.5 * ((r0 * t0 - r1 * t1) + rt * rt * std::log((r0 + t0) / (r1 + t1)));
        }
        s.q_near = (r1 * dt1 - dt2) / dr;
        s.q_far = -1.0 * (r0 * dt1 - dt2) / dr;
        s.qf_near = s.q_near / (s.q_near + s.q_far);
        s.qf_far = s.q_far / (s.q_near + s.q_far);
    }

    // union of the two end-point stencils, sorted; entrance / exit / optical-depth weights on it (raytracing.h:390-470)
    void stencils(Seg& s) const {
        int ni[2], fi[2];
        double nw[2], fw[2];
        const int nn = stencil(s.near_pt, ni, nw), fn = stencil(s.far_pt, fi, fw);
        int count = 0;
        auto merge = [&](const int* ii, const double* ww, int n) {
            for (int k = 0; k < n; ++k)
                if (ww[k] != 0.0 && std::find(s.idx, s.idx + count, ii[k]) == s.idx + count) {
                    if (count == kLimbStencil) throw std::runtime_error("B200 limb path: a traced layer touches more than four grid points");
                    s.idx[count++] = ii[k];
                }
        };
        merge(ni, nw, nn);
        merge(fi, fw, fn);
        std::sort(s.idx, s.idx + count);
        s.nidx = count;
        auto spread = [&](const int* ii, const double* ww, int n, double* out) {
            for (int k = 0; k < n; ++k)
                if (ww[k] != 0.0) out[std::find(s.idx, s.idx + count, ii[k]) - s.idx] += ww[k];
        };
        spread(ni, nw, nn, s.w_near);
        spread(fi, fw, fn, s.w_far);
        for (int k = 0; k < count; ++k) s.w_od[k] = s.w_near[k] * s.q_near + s.w_far[k] * s.q_far;
    }
};

}  // namespace

LimbPlan build_limb_plan(int nstr, const GeometrySpec& geo, const std::vector<LimbRay>& rays, const LimbOptions& opt) {
    if (geo.geotype != 2) throw std::runtime_error("B200 limb path: geometry type must be spherical");
    if (geo.altitudes.size() < 2) throw std::runtime_error("altitude grid needs at least two points");
    if (opt.num_sza < 1) throw std::runtime_error("Invalid number of dosza, must be at least 1");
    LimbPlan P;
    P.nstr = nstr;
    P.nrays = (int)rays.size();
    P.ms_do = opt.ms_do;
    P.ss_exact = opt.ss_exact;
    P.nss = opt.num_ss_moments;
    const int L = (int)geo.altitudes.size() - 1;
    P.nalt = L;
    Tracer tr(geo);
    const Frame& F = tr.frame;
    const double re = geo.earth_radius;

    // ---- trace the lines of sight
    std::vector<Path> paths(rays.size());
    for (size_t i = 0; i < rays.size(); ++i) {
        const LimbRay& r = rays[i];
        Vec obs, look;
        if (r.kind == 0) {  // lib/viewinggeometry/groundviewing.cpp:16-60 (spherical)
            const double cos_sza = r.p[0], rel_az = r.p[1], cos_vza = r.p[2], obs_alt = r.p[3];
            const Vec ground = F.point_at_sza(cos_sza, 0.0, 0.0);
            look = F.look_from_azimuth(ground, -(kPiL - rel_az), cos_vza) * -1.0;
            const double b = 2.0 * re * cos_vza, c = -(2.0 * re * obs_alt + obs_alt * obs_alt);
            obs = ground - look * ((-b + std::sqrt(b * b - 4 * c)) / 2);
        } else if (r.kind == 1) {  // lib/viewinggeometry/tangentaltitudesolar.cpp:33-62
            const double tan_alt = r.p[0], rel_az = r.p[1], obs_alt = r.p[2], cos_sza = r.p[3];
            const Vec tp = F.point_at_sza(cos_sza, 0.0, tan_alt);
            look = F.look_from_azimuth(tp, rel_az, 0.0);
            const double a = re + obs_alt, b = re + tan_alt;
            obs = tp - look * std::sqrt(a * a - b * b);
        } else {
            throw std::runtime_error("B200 limb path: unsupported viewing ray kind");
        }
        tr.trace(obs, look, paths[i], false);
    }

    // ---- SZA grid of the DO solves (do_source.cpp:61-92)
    double cmin = 1.0, cmax = -1.0;
    for (const Path& p : paths)
        for (const Seg& s : p.segs) {
            cmin = std::min({cmin, s.csz_near, s.csz_far});
            cmax = std::max({cmax, s.csz_near, s.csz_far});
        }
    if (opt.num_sza == 1) {
        P.sza_grid = {dot(F.ez, F.sun)};
    } else {
        for (int i = 0; i < opt.num_sza; ++i) P.sza_grid.push_back(cmin + (cmax - cmin) * i / (opt.num_sza - 1));
        P.sza_grid.back() = cmax;
    }
    P.nsza = (int)P.sza_grid.size();
    Axis sza_axis;
    sza_axis.g = P.sza_grid;
    sza_axis.interp = 1;
    sza_axis.uniform = true;
    if (opt.ms_do) {
        for (int s = 0; s < P.nsza; ++s) {
            GeometrySpec gs = geo;
            gs.geotype = 1;   // chapman factors of the spherical solar rays == the pseudo-spherical construction
            gs.cos_sza = P.sza_grid[s];
            P.sza_plans.push_back(build_plan(nstr, gs, {}));
        }
    }
    // ---- source table axes (do_source_diffuse_storage.cpp:16-34)
    Axis alt_axis, ang_axis;
    for (int q = 0; q < L; ++q) alt_axis.g.push_back((geo.altitudes[q] + geo.altitudes[q + 1]) / 2.0);
    alt_axis.interp = 1;
    alt_axis.uniform = false;
    for (int i = 0; i < kLimbAngles; ++i) ang_axis.g.push_back(-1.0 + 2.0 * i / (kLimbAngles - 1));
    ang_axis.g.back() = 1.0;
    ang_axis.interp = 1;
    ang_axis.uniform = false;
    P.layer_fraction.resize(L);
    for (int p = 0; p < L; ++p) {  // :745-749: layer p samples the altitude of table row L - 1 - p
        const double ceil_h = geo.altitudes[L - p], floor_h = geo.altitudes[L - 1 - p];
        P.layer_fraction[p] = (ceil_h - alt_axis.g[L - 1 - p]) / (ceil_h - floor_h);
    }
    P.lp_ang.assign((size_t)kLimbAngles * nstr * nstr, 0.0);
    for (int a = 0; a < kLimbAngles; ++a)
        for (int m = 0; m < nstr; ++m)
            for (int l = 0; l < nstr; ++l) P.lp_ang[((size_t)a * nstr + m) * nstr + l] = wigner_dm0(m, l, ang_axis.g[a]);

    // ---- segments
    std::map<std::array<int, 3>, int> point_of;   // (angle, altitude, sza) -> compact index
    auto point_index = [&](int a, int q, int s) {
        auto it = point_of.find({a, q, s});
        if (it != point_of.end()) return it->second;
        const int id = (int)P.pt_angle.size();
        point_of[{a, q, s}] = id;
        P.pt_angle.push_back(a);
        P.pt_alt.push_back(q);
        P.pt_sza.push_back(s);
        return id;
    };
    P.seg_start.assign(1, 0);
    P.gnd_hit.assign(P.nrays, 0);
    P.gnd_sza_idx.assign((size_t)P.nrays * 2, 0);
    P.gnd_sza_w.assign((size_t)P.nrays * 2, 0.0);
    P.gnd_mu_in.assign(P.nrays, 0.0);
    P.ray_cos_scatter.assign(P.nrays, 0.0);
    P.wig_ss.assign((size_t)P.nrays * P.nss, 0.0);
    for (int r = 0; r < P.nrays; ++r) {
        const Path& path = paths[r];
        for (const Seg& s : path.segs) {
            for (int k = 0; k < kLimbStencil; ++k) {
                P.od_idx.push_back(k < s.nidx ? s.idx[k] : 0);
                P.od_w.push_back(k < s.nidx ? s.w_od[k] : 0.0);
                P.ent_w.push_back(k < s.nidx ? s.w_near[k] : 0.0);
                P.exit_w.push_back(k < s.nidx ? s.w_far[k] : 0.0);
            }
            // SSA at the mid-point (do_source.cpp:94-124)
            EndPoint mid;
            mid.pos = (s.near_pt.pos + s.far_pt.pos) * 0.5;
            int mi[2];
            double mw[2];
            const int mn = tr.stencil(mid, mi, mw);
            for (int k = 0; k < 2; ++k) {
                P.mid_idx.push_back(k < mn ? mi[k] : 0);
                P.mid_w.push_back(k < mn ? mw[k] : 0.0);
            }
            P.seg_len.push_back(s.length);
            P.seg_qfrac.push_back(s.qf_near);
            P.seg_qfrac.push_back(s.qf_far);
            int lower = 0;   // singlescattersource.cpp:1049-1078
            if (geo.interp == 2) lower = (s.r_far > s.r_near) ? 1 : 2;
            P.seg_lower.push_back(lower);
            // DO source interpolation (do_source_diffuse_storage.cpp:84-209)
            const double altitude = (len(s.near_pt.pos) + len(s.far_pt.pos)) / 2.0 - re;
            const double cz_near = dot(s.near_pt.pos, s.look) / (len(s.near_pt.pos) * len(s.look));
            const double cz_far = dot(s.far_pt.pos, s.look) / (len(s.far_pt.pos) * len(s.look));
            const double cos_angle = -(cz_near + cz_far) / 2.0;
            const double azi = (s.saz_near + s.saz_far) / 2.0;
            const double cos_sza = (s.csz_near + s.csz_far) / 2.0;
            int ai[2], gi[2], si[2];
            double aw[2], gw[2], sw[2];
            const int an = alt_axis.lookup(altitude, ai, aw), gn = ang_axis.lookup(cos_angle, gi, gw), sn = sza_axis.lookup(cos_sza, si, sw);
            int e = 0;
            int pts[kLimbSrcEntries];
            double wts[kLimbSrcEntries];
            for (int c = 0; c < kLimbSrcEntries; ++c) {
                pts[c] = -1;
                wts[c] = 0.0;
            }
            if (opt.ms_do && !(s.length < 1e-4))   // MINIMUM_SHELL_SIZE_M: empty shells carry no source
                for (int is = 0; is < sn; ++is)
                    for (int ia = 0; ia < an; ++ia)
                        for (int ig = 0; ig < gn; ++ig) {
                            pts[e] = point_index(gi[ig], ai[ia], si[is]);
                            wts[e] = aw[ia] * gw[ig] * sw[is];
                            ++e;
                        }
            else if (opt.ms_do)   // the reference still marks the points as needed (the map is filled before the size test)
                for (int is = 0; is < sn; ++is)
                    for (int ia = 0; ia < an; ++ia)
                        for (int ig = 0; ig < gn; ++ig) point_index(gi[ig], ai[ia], si[is]);
            for (int c = 0; c < kLimbSrcEntries; ++c) {
                P.src_pt.push_back(pts[c]);
                P.src_w.push_back(wts[c]);
            }
            for (int m = 0; m < nstr; ++m) P.src_cos.push_back(std::cos(m * azi));
        }
        P.seg_start.push_back((int)P.seg_len.size());
        if (!path.segs.empty()) {
            const Seg& end = path.segs[0];
            // straight rays have one scattering angle (phasehandler.cpp:241-262, math/scattering.h:77-96)
            double c = dot(F.sun * -1.0, end.look * -1.0);
            c = std::max(-1.0, std::min(1.0, c));
            P.ray_cos_scatter[r] = c;
            for (int l = 0; l < P.nss; ++l) P.wig_ss[(size_t)r * P.nss + l] = wigner_dm0(0, l, c);
            if (path.ground) {
                P.gnd_hit[r] = 1;
                // order-0 ground source, interpolated in SZA at the near end of the last segment (sic, do_source_interpolated_pp.cpp:65-72;
                // do_source_diffuse_storage.cpp:211-267); the Lambertian value does not depend on the outgoing angle
                double csz, saz;
                Tracer::solar_angles(F.sun, end.near_pt.pos, end.look, csz, saz);
                int si[2];
                double sw[2];
                const int sn = sza_axis.lookup(csz, si, sw);
                for (int k = 0; k < 2; ++k) {
                    P.gnd_sza_idx[(size_t)r * 2 + k] = k < sn ? si[k] : 0;
                    P.gnd_sza_w[(size_t)r * 2 + k] = k < sn ? sw[k] : 0.0;
                }
                // exact single scatter off the ground: cos(SZA) at the far end of the last segment (singlescattersource.cpp:251-267)
                double mu_in, phi;
                Tracer::solar_angles(F.sun, end.far_pt.pos, end.look, mu_in, phi);
                P.gnd_mu_in[r] = mu_in;
            }
        }
    }
    P.nseg = (int)P.seg_len.size();
    P.npts = (int)P.pt_angle.size();

    // ---- solar rays from every segment boundary (solartransmissionexact.cpp:36-96)
    P.sol_start.assign(1, 0);
    if (opt.ss_exact) {
        Path sun_path;
        for (int r = 0; r < P.nrays; ++r) {
            const Path& path = paths[r];
            const int n = (int)path.segs.size();
            for (int b = 0; b <= n; ++b) {
                // boundary 0: far end of segment 0; boundary b > 0: near end of segment b - 1
                const EndPoint& from = (n == 0) ? EndPoint() : (b == 0 ? path.segs[0].far_pt : path.segs[b - 1].near_pt);
                std::map<int, double> row;
                int blocked = 0;
                if (n > 0) {
                    tr.trace(from.pos, F.sun, sun_path, true);
                    if (sun_path.ground) {
                        blocked = 1;
                    } else {
                        for (const Seg& s : sun_path.segs)
                            for (int k = 0; k < s.nidx; ++k) row[s.idx[k]] += s.w_od[k];
                    }
                }
                for (const auto& kv : row) {
                    P.sol_idx.push_back(kv.first);
                    P.sol_w.push_back(kv.second);
                }
                P.sol_start.push_back((int)P.sol_idx.size());
                P.sol_blocked.push_back(blocked);
            }
        }
    } else {
        for (int r = 0; r < P.nrays; ++r)
            for (int b = 0; b <= P.seg_start[r + 1] - P.seg_start[r]; ++b) {
                P.sol_start.push_back(0);
                P.sol_blocked.push_back(0);
            }
    }
    P.nbnd = (int)P.sol_blocked.size();
    P.ray_order.resize(P.nrays);
    for (int r = 0; r < P.nrays; ++r) P.ray_order[r] = r;
    std::stable_sort(P.ray_order.begin(), P.ray_order.end(), [&](int a, int b) {
        return P.seg_start[a + 1] - P.seg_start[a] > P.seg_start[b + 1] - P.seg_start[b];
    });
    return P;
}

}  // namespace disco

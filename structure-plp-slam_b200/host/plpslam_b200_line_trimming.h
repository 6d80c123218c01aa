// plpslam_b200_line_trimming.h -- host-side end-point trimming of optimised 3-D lines, on plain PODs.
//
// optimize::local_bundle_adjuster_extended_line::endpoint_trimming
// (optimize/local_bundle_adjuster_extended_line.cc:676-787, the same code in global_bundle_adjuster.cc): after a bundle
// adjustment every line landmark gets new Pluecker coordinates; its two 3-D end points (used for display, for the
// reprojection in tracking and for the distance gates) are re-derived on the REFERENCE keyframe: the detected 2-D end
// points are dropped onto the re-projected line, each foot point spans a plane with the camera centre, and the plane is
// intersected with the 3-D line.  A line whose end points move by more than 0.1 x the keyframe's median depth is an
// outlier (:773-786).  This runs under the map mutex in the write-back of the adapter (one call per local line, a few
// hundred per BA); it is a few dozen flops per line on data that lives in host objects, so it stays on the host.
// No reference types: compiled and tested in the authoring environment (tests/test_line_trimming.py).
#pragma once
#include <cmath>

namespace plpslam_b200 {

struct trimming_camera {
    double fx, fy, cx, cy;
};

// pose_cw: 4x4 row-major pose of the reference keyframe; plucker = (m, d); sp / ep: the keyline's end points in that
// keyframe (cv::Point2f); old_endpoints: Line::get_pos_in_world() (sp, ep); median_depth: ref_kf->compute_median_depth(true).
// Returns false when the line must be erased; updated_endpoints (6) is written in both cases, like the reference's
// out-parameter.
inline bool endpoint_trimming(const trimming_camera &cam, const double *pose_cw, const double *plucker, float sp_x, float sp_y,
                              float ep_x, float ep_y, const double *old_endpoints, double median_depth,
                              double *updated_endpoints) {
    const double R[9] = {pose_cw[0], pose_cw[1], pose_cw[2], pose_cw[4], pose_cw[5], pose_cw[6], pose_cw[8], pose_cw[9], pose_cw[10]};
    const double t[3] = {pose_cw[3], pose_cw[7], pose_cw[11]};
    const double *m = plucker, *d = plucker + 3;
    // [2] (transformation_line_cw * L).head<3>() = R m + [t]x R d, then _K (:705-719)
    double Rm[3], Rd[3];
    for (int r = 0; r < 3; ++r) {
        Rm[r] = R[r * 3] * m[0] + R[r * 3 + 1] * m[1] + R[r * 3 + 2] * m[2];
        Rd[r] = R[r * 3] * d[0] + R[r * 3 + 1] * d[1] + R[r * 3 + 2] * d[2];
    }
    const double lc[3] = {Rm[0] + (t[1] * Rd[2] - t[2] * Rd[1]), Rm[1] + (t[2] * Rd[0] - t[0] * Rd[2]),
                          Rm[2] + (t[0] * Rd[1] - t[1] * Rd[0])};
    const double l1 = cam.fy * lc[0], l2 = cam.fx * lc[1], l3 = -cam.fy * cam.cx * lc[0] - cam.fx * cam.cy * lc[1] + cam.fx * cam.fy * lc[2];
    // P = K [R | t] (:740-745)
    const double K[9] = {cam.fx, 0, cam.cx, 0, cam.fy, cam.cy, 0, 0, 1};
    double P[12];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += K[r * 3 + k] * (c < 3 ? R[k * 3 + c] : t[k]);
            P[r * 4 + c] = s;
        }
    const float px[2] = {sp_x, ep_x}, py[2] = {sp_y, ep_y};
    for (int e = 0; e < 2; ++e) {
        const double x = px[e], y = py[e];
        // [3] closest point on the re-projected line, [4] its y-axis intercept partner (:724-736)
        const double xc = -(y - (l2 / l1) * x + (l3 / l2)) * ((l1 * l2) / (l1 * l1 + l2 * l2));
        const double yc = -(l1 / l2) * xc - (l3 / l2);
        const double y0 = y - (l2 / l1) * x;
        // [5] plane through the camera centre and the 2-D line (xc, yc, 1) x (0, y0, 1)
        const double lt[3] = {yc * 1.0 - 1.0 * y0, 1.0 * 0.0 - xc * 1.0, xc * y0 - yc * 0.0};
        double pl[4];
        for (int c = 0; c < 4; ++c) pl[c] = P[c] * lt[0] + P[4 + c] * lt[1] + P[8 + c] * lt[2];
        // [6] Pluecker matrix [[m]x d; -d^T 0] times the plane
        const double X0 = (-m[2] * pl[1] + m[1] * pl[2]) + d[0] * pl[3];
        const double X1 = (m[2] * pl[0] - m[0] * pl[2]) + d[1] * pl[3];
        const double X2 = (-m[1] * pl[0] + m[0] * pl[1]) + d[2] * pl[3];
        const double X3 = -(d[0] * pl[0] + d[1] * pl[1] + d[2] * pl[2]);
        updated_endpoints[3 * e] = X0 / X3;
        updated_endpoints[3 * e + 1] = X1 / X3;
        updated_endpoints[3 * e + 2] = X2 / X3;
    }
    double ch[2];
    for (int e = 0; e < 2; ++e) {
        const double dx = updated_endpoints[3 * e] - old_endpoints[3 * e], dy = updated_endpoints[3 * e + 1] - old_endpoints[3 * e + 1],
                     dz = updated_endpoints[3 * e + 2] - old_endpoints[3 * e + 2];
        ch[e] = std::sqrt(dx * dx + dy * dy + dz * dz) / median_depth;
    }
    return !(ch[0] > 0.1 || ch[1] > 0.1);  // :773-786
}

}  // namespace plpslam_b200

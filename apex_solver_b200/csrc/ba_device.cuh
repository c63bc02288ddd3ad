// ba_device.cuh — per-observation FP64 math of the bundle-adjustment path, as device functions.
//
// Each function names the reference code it is the sm_100a equivalent of (paths relative to the
// reference repo). The translation unit is compiled with -fmad=false so that the element-wise
// arithmetic (projection, Jacobians, loss correction, manifold update) rounds exactly like the
// reference's scalar Rust; fused multiply-adds are written explicitly (fma()) only in reductions,
// whose summation order differs from the CPU anyway.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/apex_gpu.h"

namespace apex {

constexpr double F64_EPS = 2.220446049250313e-16;
constexpr double F64_MIN = -1.7976931348623157e308;  // Rust f64::MIN
constexpr double SMALL_ANGLE_THRESHOLD = 1e-10;      // crates/apex-manifolds/src/lib.rs:61
constexpr double GEOMETRIC_PRECISION = 1e-6;         // crates/apex-camera-models/src/lib.rs:56
constexpr double MIN_DEPTH = 1e-6;                   // crates/apex-camera-models/src/lib.rs:80

#define APEX_HD __host__ __device__ __forceinline__

APEX_HD double dmax(double a, double b) { return (a < b) ? b : a; }  // std::max / f64::max on non-NaN
APEX_HD double dmin(double a, double b) { return (b < a) ? b : a; }

struct V3 { double x, y, z; };
struct Quat { double w, i, j, k; };
struct Pose { V3 t; Quat q; };

APEX_HD V3 cross3(const V3& a, const V3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// Quaternion::normalize (se3.rs:107-113 normalises twice when a pose is built from a DVector, se3.rs:200-206)
APEX_HD Quat quat_normalize(Quat q) {
  double n = sqrt(q.w * q.w + q.i * q.i + q.j * q.j + q.k * q.k);
  return {q.w / n, q.i / n, q.j / n, q.k / n};
}

// UnitQuaternion * Vector3 (SO3::act, so3.rs:359-378): t = 2 (qv x v); v' = t*w + qv x t + v
APEX_HD V3 quat_rotate(const Quat& q, const V3& v) {
  V3 qv{q.i, q.j, q.k};
  V3 t = cross3(qv, v);
  t = {t.x * 2.0, t.y * 2.0, t.z * 2.0};
  V3 c = cross3(qv, t);
  return {t.x * q.w + c.x + v.x, t.y * q.w + c.y + v.y, t.z * q.w + c.z + v.z};
}

// UnitQuaternion::to_rotation_matrix (SO3::rotation_matrix, so3.rs:193-195); row-major
APEX_HD void quat_to_matrix(const Quat& q, double R[9]) {
  double i = q.i, j = q.j, k = q.k, w = q.w;
  double ww = w * w, ii = i * i, jj = j * j, kk = k * k;
  double ij = i * j * 2.0, wk = w * k * 2.0, wj = w * j * 2.0, ik = i * k * 2.0, jk = j * k * 2.0, wi = w * i * 2.0;
  R[0] = ww + ii - jj - kk; R[1] = ij - wk;           R[2] = wj + ik;
  R[3] = wk + ij;           R[4] = ww - ii + jj - kk; R[5] = jk - wi;
  R[6] = ik - wj;           R[7] = wi + jk;           R[8] = ww - ii - jj + kk;
}

// Hamilton product, not renormalised (SO3::compose, so3.rs:270-290)
APEX_HD Quat quat_mul(const Quat& a, const Quat& b) {
  return {a.w * b.w - a.i * b.i - a.j * b.j - a.k * b.k,
          a.w * b.i + a.i * b.w + a.j * b.k - a.k * b.j,
          a.w * b.j - a.i * b.k + a.j * b.w + a.k * b.i,
          a.w * b.k + a.i * b.j - a.j * b.i + a.k * b.w};
}

// SO3Tangent::exp (so3.rs:558-577)
APEX_HD Quat so3_exp(const V3& th) {
  double t2 = th.x * th.x + th.y * th.y + th.z * th.z;
  if (t2 > SMALL_ANGLE_THRESHOLD) {
    V3 h{th.x / 2.0, th.y / 2.0, th.z / 2.0};
    double nn = h.x * h.x + h.y * h.y + h.z * h.z;
    double n = sqrt(nn);
    double s = sin(n) / n;
    return {cos(n), h.x * s, h.y * s, h.z * s};
  }
  return quat_normalize({1.0, th.x / 2.0, th.y / 2.0, th.z / 2.0});
}

// SO3Tangent::left_jacobian (so3.rs:595-611) applied to rho: returns J_l(theta) * rho
APEX_HD V3 so3_left_jacobian_mul(const V3& th, const V3& rho) {
  double angle = th.x * th.x + th.y * th.y + th.z * th.z;
  double K[9] = {0, -th.z, th.y, th.z, 0, -th.x, -th.y, th.x, 0};
  double J[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (angle <= SMALL_ANGLE_THRESHOLD) {
    for (int a = 0; a < 9; ++a) J[a] += 0.5 * K[a];
  } else {
    double theta = sqrt(angle), s = sin(theta), c = cos(theta);
    double a1 = (1.0 - c) / angle, a2 = (theta - s) / (angle * theta);
    double K2[9];
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) {
        double v = 0;
        for (int m = 0; m < 3; ++m) v += K[r * 3 + m] * K[m * 3 + cc];
        K2[r * 3 + cc] = v;
      }
    for (int a = 0; a < 9; ++a) J[a] += a1 * K[a] + a2 * K2[a];
  }
  return {J[0] * rho.x + J[1] * rho.y + J[2] * rho.z, J[3] * rho.x + J[4] * rho.y + J[5] * rho.z,
          J[6] * rho.x + J[7] * rho.y + J[8] * rho.z};
}

// SE3::from(DVector) (se3.rs:200-206)
APEX_HD Pose pose_from7(const double* d) {
  Pose p;
  p.t = {d[0], d[1], d[2]};
  p.q = quat_normalize(quat_normalize({d[3], d[4], d[5], d[6]}));
  return p;
}
// SE3::act (se3.rs:322-345)
APEX_HD V3 pose_act(const Pose& p, const V3& v) {
  V3 r = quat_rotate(p.q, v);
  return {r.x + p.t.x, r.y + p.t.y, r.z + p.t.z};
}
// right_plus = compose(self, exp(tau)) (lib.rs:269-282, se3.rs:569-586, 272-297); tau = [rho, theta]
APEX_HD Pose pose_plus(const Pose& p, const double* tau) {
  V3 rho{tau[0], tau[1], tau[2]}, th{tau[3], tau[4], tau[5]};
  Quat qe = so3_exp(th);
  V3 te = so3_left_jacobian_mul(th, rho);
  Pose out;
  out.q = quat_mul(p.q, qe);
  V3 rt = quat_rotate(p.q, te);
  out.t = {rt.x + p.t.x, rt.y + p.t.y, rt.z + p.t.z};
  return out;
}

// ------------------------------------------------------------------------------------------------
// Camera models (crates/apex-camera-models/src/*.rs), selected at compile time
// ------------------------------------------------------------------------------------------------
template <int MODEL> struct CamK;
template <> struct CamK<APEX_CAM_BAL> { static constexpr int K = 3; };
template <> struct CamK<APEX_CAM_PINHOLE> { static constexpr int K = 4; };
template <> struct CamK<APEX_CAM_KANNALA_BRANDT> { static constexpr int K = 8; };
template <> struct CamK<APEX_CAM_DOUBLE_SPHERE> { static constexpr int K = 6; };
template <> struct CamK<APEX_CAM_RADTAN> { static constexpr int K = 9; };
template <> struct CamK<APEX_CAM_UCM> { static constexpr int K = 5; };
template <> struct CamK<APEX_CAM_EUCM> { static constexpr int K = 6; };
template <> struct CamK<APEX_CAM_FOV> { static constexpr int K = 5; };
template <> struct CamK<APEX_CAM_FTHETA> { static constexpr int K = 6; };

inline int model_intr_dim(int model) {
  switch (model) {
    case APEX_CAM_BAL: return 3;
    case APEX_CAM_PINHOLE: return 4;
    case APEX_CAM_KANNALA_BRANDT: return 8;
    case APEX_CAM_DOUBLE_SPHERE: return 6;
    case APEX_CAM_RADTAN: return 9;
    case APEX_CAM_UCM: return 5;
    case APEX_CAM_EUCM: return 6;
    case APEX_CAM_FOV: return 5;
    case APEX_CAM_FTHETA: return 6;
    default: return -1;
  }
}

// CameraModel::project; false = Err (the factor then zeroes the rows, projection_factor.rs:227-239)
template <int MODEL>
APEX_HD bool cam_project(const double* in, const V3& p, double uv[2]) {
  if constexpr (MODEL == APEX_CAM_BAL) {  // bal_pinhole.rs:273-296, :154-156
    if (!(p.z < -MIN_DEPTH)) return false;
    double inv_neg_z = -1.0 / p.z;
    double xn = p.x * inv_neg_z, yn = p.y * inv_neg_z;
    double r2 = xn * xn + yn * yn, r4 = r2 * r2;
    double d = 1.0 + in[1] * r2 + in[2] * r4;
    uv[0] = in[0] * (xn * d);
    uv[1] = in[0] * (yn * d);
    return true;
  } else if constexpr (MODEL == APEX_CAM_PINHOLE) {  // pinhole.rs:226-238, :105-107
    if (!(p.z >= 1e-6)) return false;
    double inv_z = 1.0 / p.z;
    uv[0] = in[0] * p.x * inv_z + in[2];
    uv[1] = in[1] * p.y * inv_z + in[3];
    return true;
  } else if constexpr (MODEL == APEX_CAM_KANNALA_BRANDT) {  // kannala_brandt.rs:385-450, :103-105
    if (!(p.z > F64_EPS)) return false;
    double r2 = p.x * p.x + p.y * p.y, r = sqrt(r2);
    double th = atan2(r, p.z);
    double t2 = th * th, t3 = t2 * th, t5 = t3 * t2, t7 = t5 * t2, t9 = t7 * t2;
    double thd = th + in[4] * t3 + in[5] * t5 + in[6] * t7 + in[7] * t9;
    if (r < GEOMETRIC_PRECISION) {
      double inv_z = 1.0 / p.z;
      uv[0] = in[0] * p.x * inv_z + in[2];
      uv[1] = in[1] * p.y * inv_z + in[3];
      return true;
    }
    double inv_r = 1.0 / r;
    uv[0] = in[0] * thd * p.x * inv_r + in[2];
    uv[1] = in[1] * thd * p.y * inv_r + in[3];
    return true;
  } else if constexpr (MODEL == APEX_CAM_RADTAN) {  // rad_tan.rs:351-385, :95-97
    if (!(p.z >= GEOMETRIC_PRECISION)) return false;
    double inv_z = 1.0 / p.z;
    double xp = p.x * inv_z, yp = p.y * inv_z;
    double r2 = xp * xp + yp * yp, r4 = r2 * r2, r6 = r4 * r2;
    double k1 = in[4], k2 = in[5], p1 = in[6], p2 = in[7], k3 = in[8];
    double radial = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
    double xy = xp * yp;
    double dx = 2.0 * p1 * xy + p2 * (r2 + 2.0 * xp * xp);
    double dy = p1 * (r2 + 2.0 * yp * yp) + 2.0 * p2 * xy;
    uv[0] = in[0] * (radial * xp + dx) + in[2];
    uv[1] = in[1] * (radial * yp + dy) + in[3];
    return true;
  } else if constexpr (MODEL == APEX_CAM_UCM) {  // ucm.rs:326-356, :100-108
    double alpha = in[4];
    double d = sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
    double denom = alpha * d + (1.0 - alpha) * p.z;
    double w = alpha <= 0.5 ? alpha / (1.0 - alpha) : (1.0 - alpha) / alpha;
    if (!(p.z > -w * d)) return false;
    if (denom < GEOMETRIC_PRECISION) return false;
    uv[0] = in[0] * p.x / denom + in[2];
    uv[1] = in[1] * p.y / denom + in[3];
    return true;
  } else if constexpr (MODEL == APEX_CAM_EUCM) {  // eucm.rs:346-376, :103-113
    double alpha = in[4], beta = in[5];
    double r2 = p.x * p.x + p.y * p.y;
    double d = sqrt(beta * r2 + p.z * p.z);
    double denom = alpha * d + (1.0 - alpha) * p.z;
    if (denom < GEOMETRIC_PRECISION) return false;
    if (alpha > 0.5) {
      double c = (alpha - 1.0) / (2.0 * alpha - 1.0);
      if (p.z < denom * c) return false;
    }
    uv[0] = in[0] * p.x / denom + in[2];
    uv[1] = in[1] * p.y / denom + in[3];
    return true;
  } else if constexpr (MODEL == APEX_CAM_FOV) {  // fov.rs:312-340
    if (p.z < 1.4901161193847656e-08) return false;  // f64::EPSILON.sqrt()
    double r = sqrt(p.x * p.x + p.y * p.y);
    double w = in[4];
    double m2t = tan(w / 2.0) * 2.0;
    double rd = r > GEOMETRIC_PRECISION ? atan(m2t * r / p.z) / (r * w) : m2t / w;
    uv[0] = in[0] * (p.x * rd) + in[2];
    uv[1] = in[1] * (p.y * rd) + in[3];
    return true;
  } else if constexpr (MODEL == APEX_CAM_FTHETA) {  // ftheta.rs:229-253, :140-143; intrinsics [cx,cy,k1..k4]
    if (p.z < MIN_DEPTH) return false;
    double d = sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
    double th = acos(fmin(fmax(p.z / d, -1.0), 1.0));
    double f = th * (in[2] + th * (in[3] + th * (in[4] + th * in[5])));
    double rp = sqrt(p.x * p.x + p.y * p.y);
    if (rp < GEOMETRIC_PRECISION) { uv[0] = in[0]; uv[1] = in[1]; return true; }
    double inv_rp = 1.0 / rp;
    uv[0] = in[0] + f * p.x * inv_rp;
    uv[1] = in[1] + f * p.y * inv_rp;
    return true;
  } else {  // double_sphere.rs:361-392, :118-127
    static_assert(MODEL == APEX_CAM_DOUBLE_SPHERE, "unknown camera model");
    double xi = in[4], alpha = in[5];
    double r2 = p.x * p.x + p.y * p.y;
    double d1 = sqrt(r2 + p.z * p.z);
    double w1 = alpha > 0.5 ? (1.0 - alpha) / alpha : alpha / (1.0 - alpha);
    double w2 = (w1 + xi) / sqrt(2.0 * w1 * xi + xi * xi + 1.0);
    if (!(p.z > -w2 * d1)) return false;
    double xdz = xi * d1 + p.z;
    double d2 = sqrt(r2 + xdz * xdz);
    double denom = alpha * d2 + (1.0 - alpha) * xdz;
    if (denom < GEOMETRIC_PRECISION) return false;
    uv[0] = in[0] * p.x / denom + in[2];
    uv[1] = in[1] * p.y / denom + in[3];
    return true;
  }
}

// CameraModel::jacobian_point: J[6] row-major 2x3 = d(u,v)/d(p_cam)
template <int MODEL>
APEX_HD void cam_jacobian_point(const double* in, const V3& p, double J[6]) {
  if constexpr (MODEL == APEX_CAM_BAL) {  // bal_pinhole.rs:400-435
    double f = in[0], k1 = in[1], k2 = in[2];
    double inv_neg_z = -1.0 / p.z;
    double xn = p.x * inv_neg_z, yn = p.y * inv_neg_z;
    double r2 = xn * xn + yn * yn, r4 = r2 * r2;
    double dist = 1.0 + k1 * r2 + k2 * r4;
    double dd = k1 + 2.0 * k2 * r2;
    double dxn_dz = xn * inv_neg_z, dyn_dz = yn * inv_neg_z;
    double dxd_dxn = dist + xn * dd * 2.0 * xn;
    double dxd_dyn = xn * dd * 2.0 * yn;
    double dyd_dxn = yn * dd * 2.0 * xn;
    double dyd_dyn = dist + yn * dd * 2.0 * yn;
    J[0] = f * (dxd_dxn * inv_neg_z);
    J[1] = f * (dxd_dyn * inv_neg_z);
    J[2] = f * (dxd_dxn * dxn_dz + dxd_dyn * dyn_dz);
    J[3] = f * (dyd_dxn * inv_neg_z);
    J[4] = f * (dyd_dyn * inv_neg_z);
    J[5] = f * (dyd_dxn * dxn_dz + dyd_dyn * dyn_dz);
  } else if constexpr (MODEL == APEX_CAM_PINHOLE) {  // pinhole.rs:315-330
    double inv_z = 1.0 / p.z, xn = p.x * inv_z, yn = p.y * inv_z;
    J[0] = in[0] * inv_z; J[1] = 0.0; J[2] = -in[0] * xn * inv_z;
    J[3] = 0.0; J[4] = in[1] * inv_z; J[5] = -in[1] * yn * inv_z;
  } else if constexpr (MODEL == APEX_CAM_KANNALA_BRANDT) {  // kannala_brandt.rs:609-675
    double fx = in[0], fy = in[1], k1 = in[4], k2 = in[5], k3 = in[6], k4 = in[7];
    double x = p.x, y = p.y, z = p.z;
    double r = sqrt(x * x + y * y);
    double th = atan2(r, z);
    double t2 = th * th, t3 = t2 * th, t5 = t3 * t2, t7 = t5 * t2, t9 = t7 * t2;
    double thd = th + k1 * t3 + k2 * t5 + k3 * t7 + k4 * t9;
    double dthd = 1.0 + 3.0 * k1 * t2 + 5.0 * k2 * t2 * t2 + 7.0 * k3 * t2 * t2 * t2 + 9.0 * k4 * t2 * t2 * t2 * t2;
    if (r < GEOMETRIC_PRECISION) {
      J[0] = fx * dthd / z; J[1] = 0; J[2] = 0; J[3] = 0; J[4] = fy * dthd / z; J[5] = 0;
      return;
    }
    double inv_r = 1.0 / r, r2 = r * r, rz2 = r2 + z * z;
    double dth_dx = z * x / (r * rz2), dth_dy = z * y / (r * rz2), dth_dz = -r / rz2;
    double inv_r2 = inv_r * inv_r;
    J[0] = fx * (dthd * dth_dx * x * inv_r + thd * (inv_r - x * x * inv_r2 * inv_r));
    J[1] = fx * (dthd * dth_dy * x * inv_r - thd * x * y * inv_r2 * inv_r);
    J[2] = fx * dthd * dth_dz * x * inv_r;
    J[3] = fy * (dthd * dth_dx * y * inv_r - thd * x * y * inv_r2 * inv_r);
    J[4] = fy * (dthd * dth_dy * y * inv_r + thd * (inv_r - y * y * inv_r2 * inv_r));
    J[5] = fy * dthd * dth_dz * y * inv_r;
  } else if constexpr (MODEL == APEX_CAM_RADTAN) {  // rad_tan.rs:630-680
    double fx = in[0], fy = in[1], k1 = in[4], k2 = in[5], p1 = in[6], p2 = in[7], k3 = in[8];
    double inv_z = 1.0 / p.z;
    double xp = p.x * inv_z, yp = p.y * inv_z;
    double r2 = xp * xp + yp * yp, r4 = r2 * r2, r6 = r4 * r2;
    double radial = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
    double drad = k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4;
    double dxx = radial + 2.0 * xp * xp * drad + 2.0 * p1 * yp + 6.0 * p2 * xp;
    double dxy = 2.0 * xp * yp * drad + 2.0 * p1 * xp + 2.0 * p2 * yp;
    double dyx = 2.0 * yp * xp * drad + 2.0 * p1 * xp + 2.0 * p2 * yp;
    double dyy = radial + 2.0 * yp * yp * drad + 6.0 * p1 * yp + 2.0 * p2 * xp;
    J[0] = fx * (dxx * inv_z);
    J[1] = fx * (dxy * inv_z);
    J[2] = fx * (dxx * (-xp * inv_z) + dxy * (-yp * inv_z));
    J[3] = fy * (dyx * inv_z);
    J[4] = fy * (dyy * inv_z);
    J[5] = fy * (dyx * (-xp * inv_z) + dyy * (-yp * inv_z));
  } else if constexpr (MODEL == APEX_CAM_UCM) {  // ucm.rs:458-503
    double fx = in[0], fy = in[1], alpha = in[4];
    double x = p.x, y = p.y, z = p.z;
    double rho = sqrt(x * x + y * y + z * z);
    double ddx = alpha * x / rho, ddy = alpha * y / rho, ddz = alpha * z / rho + (1.0 - alpha);
    double denom = alpha * rho + (1.0 - alpha) * z;
    double denom2 = denom * denom;
    J[0] = fx * (denom - x * ddx) / denom2;
    J[1] = fx * (-x * ddy) / denom2;
    J[2] = fx * (-x * ddz) / denom2;
    J[3] = fy * (-y * ddx) / denom2;
    J[4] = fy * (denom - y * ddy) / denom2;
    J[5] = fy * (-y * ddz) / denom2;
  } else if constexpr (MODEL == APEX_CAM_EUCM) {  // eucm.rs:514-548
    double fx = in[0], fy = in[1], alpha = in[4], beta = in[5];
    double x = p.x, y = p.y, z = p.z;
    double r2 = x * x + y * y;
    double d = sqrt(beta * r2 + z * z);
    double denom = alpha * d + (1.0 - alpha) * z;
    double dd_dx = beta * x / d, dd_dy = beta * y / d, dd_dz = z / d;
    double ddx = alpha * dd_dx, ddy = alpha * dd_dy, ddz = alpha * dd_dz + (1.0 - alpha);
    double denom2 = denom * denom;
    J[0] = fx * (denom - x * ddx) / denom2;
    J[1] = fx * (-x * ddy) / denom2;
    J[2] = fx * (-x * ddz) / denom2;
    J[3] = fy * (-y * ddx) / denom2;
    J[4] = fy * (denom - y * ddy) / denom2;
    J[5] = fy * (-y * ddz) / denom2;
  } else if constexpr (MODEL == APEX_CAM_FOV) {  // fov.rs:468-530
    double fx = in[0], fy = in[1], w = in[4];
    double x = p.x, y = p.y, z = p.z;
    double r = sqrt(x * x + y * y);
    double m2t = tan(w / 2.0) * 2.0;
    if (r < GEOMETRIC_PRECISION) {
      double rd = m2t / w;
      J[0] = fx * rd; J[1] = 0; J[2] = 0; J[3] = 0; J[4] = fy * rd; J[5] = 0;
      return;
    }
    double at = atan(m2t * r / z);
    double rd = at / (r * w);
    double datan_dr = m2t * z / (z * z + m2t * m2t * r * r);
    double datan_dz = -m2t * r / (z * z + m2t * m2t * r * r);
    double drd_dr = (datan_dr * r - at) / (r * r * w);
    double drd_dz = datan_dz / (r * w);
    double dr_dx = x / r, dr_dy = y / r;
    J[0] = fx * (rd + x * drd_dr * dr_dx);
    J[1] = fx * (x * drd_dr * dr_dy);
    J[2] = fx * (x * drd_dz);
    J[3] = fy * (y * drd_dr * dr_dx);
    J[4] = fy * (rd + y * drd_dr * dr_dy);
    J[5] = fy * (y * drd_dz);
  } else if constexpr (MODEL == APEX_CAM_FTHETA) {  // ftheta.rs:295-327
    double x = p.x, y = p.y, z = p.z;
    double rp2 = x * x + y * y, d2 = rp2 + z * z;
    double d = sqrt(d2), rp = sqrt(rp2);
    if (rp < GEOMETRIC_PRECISION) {
      J[0] = in[2] / z; J[1] = 0; J[2] = 0; J[3] = 0; J[4] = in[2] / z; J[5] = 0;
      return;
    }
    double th = acos(fmin(fmax(z / d, -1.0), 1.0));
    double f = th * (in[2] + th * (in[3] + th * (in[4] + th * in[5])));
    double fp = in[2] + th * (2.0 * in[3] + th * (3.0 * in[4] + th * 4.0 * in[5]));
    double a = fp * z / (rp2 * d2), b = f / (rp2 * rp);
    J[0] = a * x * x + b * y * y;
    J[1] = (a - b) * x * y;
    J[2] = -fp * x / d2;
    J[3] = J[1];
    J[4] = a * y * y + b * x * x;
    J[5] = -fp * y / d2;
  } else {  // double_sphere.rs:532-583
    double fx = in[0], fy = in[1], xi = in[4], alpha = in[5];
    double x = p.x, y = p.y, z = p.z;
    double r2 = x * x + y * y;
    double d1 = sqrt(r2 + z * z);
    double xdz = xi * d1 + z;
    double d2 = sqrt(r2 + xdz * xdz);
    double denom = alpha * d2 + (1.0 - alpha) * xdz;
    double inv_d1 = 1.0 / d1, inv_d2 = 1.0 / d2;
    double dd1_dx = x * inv_d1, dd1_dy = y * inv_d1, dd1_dz = z * inv_d1;
    double dx_dx = xi * dd1_dx, dx_dy = xi * dd1_dy, dx_dz = xi * dd1_dz + 1.0;
    double dd2_dx = (x + xdz * dx_dx) * inv_d2, dd2_dy = (y + xdz * dx_dy) * inv_d2, dd2_dz = (xdz * dx_dz) * inv_d2;
    double dn_dx = alpha * dd2_dx + (1.0 - alpha) * dx_dx;
    double dn_dy = alpha * dd2_dy + (1.0 - alpha) * dx_dy;
    double dn_dz = alpha * dd2_dz + (1.0 - alpha) * dx_dz;
    double denom2 = denom * denom;
    J[0] = fx * (denom - x * dn_dx) / denom2;
    J[1] = fx * (-x * dn_dy) / denom2;
    J[2] = fx * (-x * dn_dz) / denom2;
    J[3] = fy * (-y * dn_dx) / denom2;
    J[4] = fy * (denom - y * dn_dy) / denom2;
    J[5] = fy * (-y * dn_dz) / denom2;
  }
}

// CameraModel::jacobian_intrinsics: J[2*K] row-major 2xK
template <int MODEL>
APEX_HD void cam_jacobian_intrinsics(const double* in, const V3& p, double* J) {
  if constexpr (MODEL == APEX_CAM_BAL) {  // bal_pinhole.rs:649-672
    double f = in[0], k1 = in[1], k2 = in[2];
    double inv_neg_z = -1.0 / p.z;
    double xn = p.x * inv_neg_z, yn = p.y * inv_neg_z;
    double r2 = xn * xn + yn * yn, r4 = r2 * r2;
    double dist = 1.0 + k1 * r2 + k2 * r4;
    J[0] = xn * dist; J[1] = f * xn * r2; J[2] = f * xn * r4;
    J[3] = yn * dist; J[4] = f * yn * r2; J[5] = f * yn * r4;
  } else if constexpr (MODEL == APEX_CAM_PINHOLE) {  // pinhole.rs:385-393
    double inv_z = 1.0 / p.z;
    J[0] = p.x * inv_z; J[1] = 0; J[2] = 1; J[3] = 0;
    J[4] = 0; J[5] = p.y * inv_z; J[6] = 0; J[7] = 1;
  } else if constexpr (MODEL == APEX_CAM_KANNALA_BRANDT) {  // kannala_brandt.rs:767-836
    double fx = in[0], fy = in[1], k1 = in[4], k2 = in[5], k3 = in[6], k4 = in[7];
    double x = p.x, y = p.y, z = p.z;
    double r = sqrt(x * x + y * y);
    double th = atan2(r, z);
    double t2 = th * th, t3 = t2 * th, t5 = t3 * t2, t7 = t5 * t2, t9 = t7 * t2;
    double thd = th + k1 * t3 + k2 * t5 + k3 * t7 + k4 * t9;
    if (r < GEOMETRIC_PRECISION) {
      for (int a = 0; a < 16; ++a) J[a] = 0.0;
      return;
    }
    double inv_r = 1.0 / r;
    J[0] = x * thd * inv_r; J[1] = 0; J[2] = 1; J[3] = 0;
    J[4] = fx * t3 * x * inv_r; J[5] = fx * t5 * x * inv_r; J[6] = fx * t7 * x * inv_r; J[7] = fx * t9 * x * inv_r;
    J[8] = 0; J[9] = y * thd * inv_r; J[10] = 0; J[11] = 1;
    J[12] = fy * t3 * y * inv_r; J[13] = fy * t5 * y * inv_r; J[14] = fy * t7 * y * inv_r; J[15] = fy * t9 * y * inv_r;
  } else if constexpr (MODEL == APEX_CAM_RADTAN) {  // rad_tan.rs:783-851
    double fx = in[0], fy = in[1], k1 = in[4], k2 = in[5], p1 = in[6], p2 = in[7], k3 = in[8];
    double inv_z = 1.0 / p.z;
    double xp = p.x * inv_z, yp = p.y * inv_z;
    double r2 = xp * xp + yp * yp, r4 = r2 * r2, r6 = r4 * r2;
    double radial = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
    double xy = xp * yp;
    double dx = 2.0 * p1 * xy + p2 * (r2 + 2.0 * xp * xp);
    double dy = p1 * (r2 + 2.0 * yp * yp) + 2.0 * p2 * xy;
    J[0] = radial * xp + dx; J[1] = 0; J[2] = 1; J[3] = 0;
    J[4] = fx * xp * r2; J[5] = fx * xp * r4; J[6] = fx * 2.0 * xy; J[7] = fx * (r2 + 2.0 * xp * xp); J[8] = fx * xp * r6;
    J[9] = 0; J[10] = radial * yp + dy; J[11] = 0; J[12] = 1;
    J[13] = fy * yp * r2; J[14] = fy * yp * r4; J[15] = fy * (r2 + 2.0 * yp * yp); J[16] = fy * 2.0 * xy; J[17] = fy * yp * r6;
  } else if constexpr (MODEL == APEX_CAM_UCM) {  // ucm.rs:549-598
    double fx = in[0], fy = in[1], alpha = in[4];
    double x = p.x, y = p.y, z = p.z;
    double rho = sqrt(x * x + y * y + z * z);
    double denom = alpha * rho + (1.0 - alpha) * z;
    double xn = x / denom, yn = y / denom;
    double ucx = fx * xn, vcy = fy * yn;
    double dda = rho - z;
    J[0] = xn; J[1] = 0; J[2] = 1; J[3] = 0; J[4] = -ucx * dda / denom;
    J[5] = 0; J[6] = yn; J[7] = 0; J[8] = 1; J[9] = -vcy * dda / denom;
  } else if constexpr (MODEL == APEX_CAM_EUCM) {  // eucm.rs:652-696
    double fx = in[0], fy = in[1], alpha = in[4], beta = in[5];
    double x = p.x, y = p.y, z = p.z;
    double r2 = x * x + y * y;
    double d = sqrt(beta * r2 + z * z);
    double denom = alpha * d + (1.0 - alpha) * z;
    double dda = d - z;
    double ddb = alpha * (r2 / (2.0 * d));
    J[0] = x / denom; J[1] = 0; J[2] = 1; J[3] = 0;
    J[4] = -fx * x * dda / (denom * denom); J[5] = -fx * x * ddb / (denom * denom);
    J[6] = 0; J[7] = y / denom; J[8] = 0; J[9] = 1;
    J[10] = -fy * y * dda / (denom * denom); J[11] = -fy * y * ddb / (denom * denom);
  } else if constexpr (MODEL == APEX_CAM_FOV) {  // fov.rs:647-708
    double fx = in[0], fy = in[1], w = in[4];
    double x = p.x, y = p.y, z = p.z;
    double r = sqrt(x * x + y * y);
    double t2 = tan(w / 2.0);
    double m2t = t2 * 2.0;
    double rd = r > GEOMETRIC_PRECISION ? atan(m2t * r / z) / (r * w) : m2t / w;
    double sec2 = 1.0 + t2 * t2;
    double drd_dw;
    if (r > GEOMETRIC_PRECISION) {
      double al = 2.0 * t2 * r / z;
      double at = atan(al);
      double dal_dw = sec2 * r / z;
      double datan_dw = dal_dw / (1.0 + al * al);
      drd_dw = (datan_dw * r * w - at * r) / (r * r * w * w);
    } else {
      drd_dw = (sec2 * w - 2.0 * t2) / (w * w);
    }
    J[0] = x * rd; J[1] = 0; J[2] = 1; J[3] = 0; J[4] = fx * x * drd_dw;
    J[5] = 0; J[6] = y * rd; J[7] = 0; J[8] = 1; J[9] = fy * y * drd_dw;
  } else if constexpr (MODEL == APEX_CAM_FTHETA) {  // ftheta.rs:329-356
    double x = p.x, y = p.y, z = p.z;
    double rp2 = x * x + y * y;
    double rp = sqrt(rp2), d = sqrt(rp2 + z * z);
    for (int a = 0; a < 12; ++a) J[a] = 0.0;
    J[0] = 1.0; J[7] = 1.0;
    if (rp < GEOMETRIC_PRECISION) return;
    double th = acos(fmin(fmax(z / d, -1.0), 1.0));
    double cphi = x / rp, sphi = y / rp;
    double tp = th;
    for (int col = 2; col < 6; ++col) { J[col] = tp * cphi; J[6 + col] = tp * sphi; tp *= th; }
  } else {  // double_sphere.rs:691-731
    double fx = in[0], fy = in[1], xi = in[4], alpha = in[5];
    double x = p.x, y = p.y, z = p.z;
    double r2 = x * x + y * y;
    double d1 = sqrt(r2 + z * z);
    double xdz = xi * d1 + z;
    double d2 = sqrt(r2 + xdz * xdz);
    double denom = alpha * d2 + (1.0 - alpha) * xdz;
    double inv_denom = 1.0 / denom, inv_d2 = 1.0 / d2;
    double dd2_dxi = (xdz * d1) * inv_d2;
    double dn_dxi = alpha * dd2_dxi + (1.0 - alpha) * d1;
    double dn_dalpha = d2 - xdz;
    double inv_denom2 = inv_denom * inv_denom;
    J[0] = x * inv_denom; J[1] = 0; J[2] = 1; J[3] = 0;
    J[4] = -fx * x * dn_dxi * inv_denom2; J[5] = -fx * x * dn_dalpha * inv_denom2;
    J[6] = 0; J[7] = y * inv_denom; J[8] = 0; J[9] = 1;
    J[10] = -fy * y * dn_dxi * inv_denom2; J[11] = -fy * y * dn_dalpha * inv_denom2;
  }
}

// ------------------------------------------------------------------------------------------------
// LossFunction::evaluate (src/core/loss_functions.rs) -> [rho, rho', rho'']
// ------------------------------------------------------------------------------------------------
struct LossSpec { int id; double p0, p1; };

APEX_HD void loss_evaluate(const LossSpec& L, double s, double rho[3]) {
  switch (L.id) {
    case APEX_LOSS_L2: rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; return;  // :174-179
    case APEX_LOSS_L1: {  // :236-250
      if (s < F64_EPS) { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; return; }
      double q = sqrt(s);
      rho[0] = 2.0 * q; rho[1] = 1.0 / q; rho[2] = -1.0 / (2.0 * s * q); return;
    }
    case APEX_LOSS_HUBER: {  // :353-381
      double scale = L.p0, scale2 = scale * scale;
      if (s > scale2) {
        double r = sqrt(s);
        double rho1 = dmax(scale / r, F64_MIN);
        rho[0] = 2.0 * scale * r - scale2; rho[1] = rho1; rho[2] = -rho1 / (2.0 * s);
      } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
      return;
    }
    case APEX_LOSS_CAUCHY: {  // :486-508
      double scale2 = L.p0 * L.p0, c = 1.0 / scale2;
      double sum = 1.0 + s * c, inv = 1.0 / sum;
      rho[0] = scale2 * log(sum) / 2.0; rho[1] = dmax(inv, F64_MIN); rho[2] = -c * (inv * inv);
      return;
    }
    case APEX_LOSS_FAIR: {  // :585-607
      if (s < F64_EPS) { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; return; }
      double sc = L.p0;
      double x = sqrt(s), ax = fabs(x), cpx = sc + ax;
      rho[0] = sc * sc * (ax / sc - log(1.0 + ax / sc));
      rho[1] = 0.5 / cpx;
      rho[2] = -1.0 / (4.0 * s * cpx * cpx);
      return;
    }
    case APEX_LOSS_GEMAN_MCCLURE: {  // :674-687
      double c = 1.0 / (L.p0 * L.p0);
      double denom = 1.0 + s * c, inv = 1.0 / denom, inv2 = inv * inv;
      rho[0] = s * inv; rho[1] = inv2; rho[2] = -2.0 * c * inv2 * inv; return;
    }
    case APEX_LOSS_WELSCH: {  // :759-770
      double scale2 = L.p0 * L.p0, inv_scale2 = 1.0 / scale2;
      double e = exp(-s * inv_scale2);
      rho[0] = (scale2 / 2.0) * (1.0 - e); rho[1] = 0.5 * e; rho[2] = -0.5 * inv_scale2 * e; return;
    }
    case APEX_LOSS_TUKEY: {  // :848-869
      double sc = L.p0, sc2 = sc * sc;
      double x = sqrt(s);
      if (x > sc) { rho[0] = sc2 / 6.0; rho[1] = 0.0; rho[2] = 0.0; return; }
      double ratio = x / sc, ratio2 = ratio * ratio, om = 1.0 - ratio2, om2 = om * om;
      rho[0] = (sc2 / 6.0) * (1.0 - om * om2); rho[1] = 0.5 * om2; rho[2] = -(ratio / sc2) * om; return;
    }
    case APEX_LOSS_ANDREWS: {  // :949-969
      double sc = L.p0, sc2 = sc * sc, thr = 3.14159265358979323846 * sc;
      double x = sqrt(s);
      if (x > thr) { rho[0] = 2.0 * sc2; rho[1] = 0.0; rho[2] = 0.0; return; }
      double arg = x / sc, sv = sin(arg), cv = cos(arg);
      rho[0] = sc2 * (1.0 - cv); rho[1] = 0.5 * sv; rho[2] = (0.25 / sc) * cv / dmax(x, F64_EPS); return;
    }
    case APEX_LOSS_RAMSAY_EA: {  // :1037-1055
      double sc = L.p0, inv_sc2 = 1.0 / (sc * sc);
      double x = sqrt(s), ax = sc * x, e = exp(-ax);
      rho[0] = inv_sc2 * (1.0 - e * (1.0 + ax)); rho[1] = 0.5 * e; rho[2] = -(sc / (4.0 * dmax(x, F64_EPS))) * e; return;
    }
    case APEX_LOSS_TRIMMED_MEAN: {  // :1132-1141
      double sc2 = L.p0 * L.p0;
      if (s <= sc2) { rho[0] = s / 2.0; rho[1] = 0.5; rho[2] = 0.0; }
      else { rho[0] = sc2 / 2.0; rho[1] = 0.0; rho[2] = 0.0; }
      return;
    }
    case APEX_LOSS_LP_NORM: {  // :1207-1224
      if (s < F64_EPS) { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; return; }
      double e0 = L.p0 / 2.0, e1 = e0 - 1.0, e2 = e1 - 1.0;
      rho[0] = pow(s, e0); rho[1] = e0 * pow(s, e1); rho[2] = e0 * e1 * pow(s, e2); return;
    }
    case APEX_LOSS_BARRON: {  // :1316-1355
      double alpha = L.p0, sc = L.p1, sc2 = sc * sc;
      if (fabs(alpha) < 1e-6) {
        double denom = 1.0 + s / sc2, inv = 1.0 / denom;
        rho[0] = (sc2 / 2.0) * log(denom); rho[1] = dmax(inv, F64_MIN); rho[2] = -inv * inv / sc2; return;
      }
      if (fabs(alpha - 2.0) < 1e-6) { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; return; }
      double x = sqrt(s), nz = x / sc, nz2 = nz * nz;
      double inner = fabs(alpha) / 2.0 * nz2 + 1.0;
      double power = pow(inner, alpha / 2.0);
      rho[0] = (fabs(alpha) / sc2) * (power - 1.0);
      rho[1] = 0.5 * pow(inner, alpha / 2.0 - 1.0);
      rho[2] = (alpha - 2.0) / (4.0 * sc2) * pow(inner, alpha / 2.0 - 2.0);
      return;
    }
    case APEX_LOSS_T_DISTRIBUTION: {  // :1445-1461
      double nu = L.p0, h = (nu + 1.0) / 2.0;
      double inner = 1.0 + s / nu, denom = nu + s;
      rho[0] = h * log(inner); rho[1] = h / denom; rho[2] = -h / (denom * denom); return;
    }
  }
  rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
}

// Corrector::new (src/core/corrector.rs:143-181)
struct Corrector { double sqrt_rho1, residual_scaling, alpha_sq_norm; };
APEX_HD Corrector corrector_new(const LossSpec& L, double sq_norm) {
  double rho[3];
  loss_evaluate(L, sq_norm, rho);
  double sqrt_rho1 = sqrt(rho[1]);
  if (sq_norm == 0.0 || rho[2] <= 0.0) return {sqrt_rho1, sqrt_rho1, 0.0};
  double d = dmax(1.0 + 2.0 * sq_norm * rho[2] / rho[1], 0.0);
  double alpha = 1.0 - sqrt(d);
  return {sqrt_rho1, sqrt_rho1 / (1.0 - alpha), alpha / sq_norm};
}

// Corrector::correct_jacobian on one 2 x ncol row-major block (corrector.rs:233-254)
template <int NCOL>
APEX_HD void correct_jacobian(const Corrector& c, const double* r, double* J) {
  if (c.alpha_sq_norm == 0.0) {
#pragma unroll
    for (int a = 0; a < 2 * NCOL; ++a) J[a] *= c.sqrt_rho1;
  } else {
#pragma unroll
    for (int col = 0; col < NCOL; ++col) {
      double j0 = J[col], j1 = J[NCOL + col];
      double rtj = r[0] * j0 + r[1] * j1;
      J[col] = (j0 - r[0] * rtj * c.alpha_sq_norm) * c.sqrt_rho1;
      J[NCOL + col] = (j1 - r[1] * rtj * c.alpha_sq_norm) * c.sqrt_rho1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// ProjectionFactor::linearize for one observation + the loss correction of linearize_block
// (src/factors/projection_factor.rs:184-296, src/linearizer/mod.rs:143-149).
// Outputs the loss-corrected residual r[2], camera block jc[2][DC] (pose 6 | intrinsics K) and
// landmark block jp[2][3]. An invalid projection leaves everything zero.
// ------------------------------------------------------------------------------------------------
template <int MODEL, bool OPT_INTR, bool WANT_JAC>
APEX_HD void linearize_obs(const LossSpec& L, const Pose& pose, const double* intr, const V3& pw, double u_obs, double v_obs,
                           double r[2], double* jc /*2*DC*/, double* jp /*6*/) {
  constexpr int K = CamK<MODEL>::K;
  constexpr int DC = 6 + (OPT_INTR ? K : 0);
  r[0] = 0.0; r[1] = 0.0;
  if constexpr (WANT_JAC) {
#pragma unroll
    for (int a = 0; a < 2 * DC; ++a) jc[a] = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) jp[a] = 0.0;
  }
  V3 pc = pose_act(pose, pw);  // projection_factor.rs:224
  double uv[2];
  if (cam_project<MODEL>(intr, pc, uv)) {
    r[0] = uv[0] - u_obs;  // :242-243
    r[1] = uv[1] - v_obs;
    if constexpr (WANT_JAC) {
      double A[6];
      cam_jacobian_point<MODEL>(intr, pc, A);
      double R[9];
      quat_to_matrix(pose.q, R);
      // jacobian_pose (lib.rs:560-589 / bal_pinhole.rs:528-556): d p_cam / d[rho,theta] = [R | -R [p_w]x]
      double S[9] = {0, -pw.z, pw.y, pw.z, 0, -pw.x, -pw.y, pw.x, 0};
      double D[18];
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          if (c < 3) D[rr * 6 + c] = R[rr * 3 + c];
          else {
            double v = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) v += R[rr * 3 + k] * S[k * 3 + (c - 3)];
            D[rr * 6 + c] = -v;
          }
        }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double v = 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) v += A[rr * 3 + k] * D[k * 6 + c];
          jc[rr * DC + c] = v;  // :250-259
        }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double v = 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) v += A[rr * 3 + k] * R[k * 3 + c];
          jp[rr * 3 + c] = v;  // :262-276
        }
      if constexpr (OPT_INTR) {  // :284-291
        double JI[2 * K];
        cam_jacobian_intrinsics<MODEL>(intr, pc, JI);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
          for (int c = 0; c < K; ++c) jc[rr * DC + 6 + c] = JI[rr * K + c];
      }
    }
  }
  if (L.id != APEX_LOSS_NONE) {
    double sq = r[0] * r[0] + r[1] * r[1];
    Corrector c = corrector_new(L, sq);
    if constexpr (WANT_JAC) {
      // the reference corrects the whole 2 x (6+3+K) block column by column; the column split is immaterial
      if (c.alpha_sq_norm == 0.0) {
#pragma unroll
        for (int a = 0; a < 2 * DC; ++a) jc[a] *= c.sqrt_rho1;
#pragma unroll
        for (int a = 0; a < 6; ++a) jp[a] *= c.sqrt_rho1;
      } else {
        correct_jacobian<DC>(c, r, jc);
        correct_jacobian<3>(c, r, jp);
      }
    }
    r[0] *= c.residual_scaling;  // correct_residuals (corrector.rs:292-298)
    r[1] *= c.residual_scaling;
  }
}

// ------------------------------------------------------------------------------------------------
// 3x3 landmark blocks
// ------------------------------------------------------------------------------------------------
// nalgebra Matrix3::try_inverse (closed form, None iff det == 0); row-major
APEX_HD bool inverse3(const double* m, double* o) {
  double m11 = m[0], m12 = m[1], m13 = m[2], m21 = m[3], m22 = m[4], m23 = m[5], m31 = m[6], m32 = m[7], m33 = m[8];
  double a = m22 * m33 - m32 * m23, b = m21 * m33 - m31 * m23, c = m21 * m32 - m31 * m22;
  double det = m11 * a - m12 * b + m13 * c;
  if (det == 0.0) return false;
  o[0] = a / det; o[1] = (m13 * m32 - m33 * m12) / det; o[2] = (m12 * m23 - m22 * m13) / det;
  o[3] = -b / det; o[4] = (m11 * m33 - m31 * m13) / det; o[5] = (m13 * m21 - m23 * m11) / det;
  o[6] = c / det; o[7] = (m12 * m31 - m32 * m11) / det; o[8] = (m11 * m22 - m21 * m12) / det;
  return true;
}

// min / max eigenvalue of a symmetric 3x3 by cyclic Jacobi (stands in for nalgebra symmetric_eigenvalues();
// only min/max feed the guards of invert_landmark_blocks)
APEX_HD void sym_eig3_minmax(const double* m, double& mn, double& mx) {
  double a00 = m[0], a11 = m[4], a22 = m[8];
  double a01 = 0.5 * (m[1] + m[3]), a02 = 0.5 * (m[2] + m[6]), a12 = 0.5 * (m[5] + m[7]);
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = a01 * a01 + a02 * a02 + a12 * a12;
    double diag = a00 * a00 + a11 * a11 + a22 * a22;
    if (off <= 1e-40 * diag || off == 0.0) break;
    // rotation (0,1)
    if (a01 != 0.0) {
      double tau = (a11 - a00) / (2.0 * a01);
      double tt = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
      double c = 1.0 / sqrt(1.0 + tt * tt), s = tt * c;
      double n00 = c * c * a00 - 2.0 * s * c * a01 + s * s * a11;
      double n11 = s * s * a00 + 2.0 * s * c * a01 + c * c * a11;
      double n02 = c * a02 - s * a12, n12 = s * a02 + c * a12;
      a00 = n00; a11 = n11; a01 = 0.0; a02 = n02; a12 = n12;
    }
    // rotation (0,2)
    if (a02 != 0.0) {
      double tau = (a22 - a00) / (2.0 * a02);
      double tt = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
      double c = 1.0 / sqrt(1.0 + tt * tt), s = tt * c;
      double n00 = c * c * a00 - 2.0 * s * c * a02 + s * s * a22;
      double n22 = s * s * a00 + 2.0 * s * c * a02 + c * c * a22;
      double n01 = c * a01 - s * a12, n12 = s * a01 + c * a12;
      a00 = n00; a22 = n22; a02 = 0.0; a01 = n01; a12 = n12;
    }
    // rotation (1,2)
    if (a12 != 0.0) {
      double tau = (a22 - a11) / (2.0 * a12);
      double tt = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
      double c = 1.0 / sqrt(1.0 + tt * tt), s = tt * c;
      double n11 = c * c * a11 - 2.0 * s * c * a12 + s * s * a22;
      double n22 = s * s * a11 + 2.0 * s * c * a12 + c * c * a22;
      double n01 = c * a01 - s * a02, n02 = s * a01 + c * a02;
      a11 = n11; a22 = n22; a12 = 0.0; a01 = n01; a02 = n02;
    }
  }
  mn = dmin(a00, dmin(a11, a22));
  mx = dmax(a00, dmax(a11, a22));
}

// invert_landmark_blocks (explicit_schur.rs:377-442 with lambda argument 0, implicit_schur.rs:724-759):
// `block` (row-major, already damped) -> inverse, with the eigenvalue guards. false = SingularMatrix.
APEX_HD bool invert_landmark_block(const double* block, double* inv) {
  const double CONDITION_THRESHOLD = 1e10, MIN_EIGENVALUE_THRESHOLD = 1e-12, REGULARIZATION_SCALE = 1e-6;
  double mn, mx;
  sym_eig3_minmax(block, mn, mx);
  double b[9];
  for (int i = 0; i < 9; ++i) b[i] = block[i];
  if (mn < MIN_EIGENVALUE_THRESHOLD) {
    double reg = REGULARIZATION_SCALE + mx * REGULARIZATION_SCALE;
    b[0] += reg; b[4] += reg; b[8] += reg;
  } else if (mx / mn > CONDITION_THRESHOLD) {
    double reg = mx * REGULARIZATION_SCALE;
    b[0] += reg; b[4] += reg; b[8] += reg;
  }
  return inverse3(b, inv);
}

// nalgebra DMatrix::try_inverse: n <= 3 closed form, else LU with partial pivoting (fails on a zero pivot).
// `a` (row-major n x n, n <= MAXN) is destroyed.
template <int MAXN>
APEX_HD bool inverse_n(int n, double* a, double* o) {
  if (n == 1) { if (a[0] == 0.0) return false; o[0] = 1.0 / a[0]; return true; }
  if (n == 2) {
    double det = a[0] * a[3] - a[2] * a[1];
    if (det == 0.0) return false;
    o[0] = a[3] / det; o[1] = -a[1] / det; o[2] = -a[2] / det; o[3] = a[0] / det; return true;
  }
  if (n == 3) return inverse3(a, o);
  for (int i = 0; i < n * n; ++i) o[i] = 0.0;
  for (int i = 0; i < n; ++i) o[i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int p = c; double best = fabs(a[c * n + c]);
    for (int r = c + 1; r < n; ++r) if (fabs(a[r * n + c]) > best) { best = fabs(a[r * n + c]); p = r; }
    if (a[p * n + c] == 0.0 || !(best == best)) return false;
    if (p != c)
      for (int k = 0; k < n; ++k) {
        double t = a[p * n + k]; a[p * n + k] = a[c * n + k]; a[c * n + k] = t;
        t = o[p * n + k]; o[p * n + k] = o[c * n + k]; o[c * n + k] = t;
      }
    double d = a[c * n + c];
    for (int r = c + 1; r < n; ++r) {
      double f = a[r * n + c] / d;
      if (f == 0.0) continue;
      for (int k = c; k < n; ++k) a[r * n + k] -= f * a[c * n + k];
      for (int k = 0; k < n; ++k) o[r * n + k] -= f * o[c * n + k];
    }
  }
  for (int c = n - 1; c >= 0; --c) {
    double d = a[c * n + c];
    for (int k = 0; k < n; ++k) o[c * n + k] /= d;
    for (int r = 0; r < c; ++r) {
      double f = a[r * n + c];
      if (f == 0.0) continue;
      for (int k = 0; k < n; ++k) o[r * n + k] -= f * o[c * n + k];
    }
  }
  return true;
}

}  // namespace apex

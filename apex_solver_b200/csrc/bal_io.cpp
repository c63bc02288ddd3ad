// bal_io.cpp — host side either side of the hot path: the BAL ("Bundle Adjustment in the Large") text format and the
// factor graph the reference's CLI builds from it. Pure host C++ (no device), exported through the C ABI.
//
//   apex_bal_load          <- BalLoader::load                 crates/apex-io/src/bal.rs:138-400
//   apex_bal_build_problem <- run_bundle_adjustment/add_factors bin/bundle_adjustment.rs:212-441
//
// The loader is a single pass over the file bytes with strtod / strtoull (correctly rounded, like Rust's
// str::parse::<f64>); line numbers are tracked so that the error text matches IoError's Display.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <algorithm>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/apex_gpu.h"

struct apex_bal_dataset {
  uint32_t ncam = 0, npts = 0;
  uint64_t nobs = 0;
  std::vector<double> cameras, points, obs_uv;
  std::vector<uint32_t> obs_cam, obs_pt;
  // storage behind the last apex_bal_build_problem
  std::vector<double> p_pose, p_intr, p_pt, p_uv;
  std::vector<uint32_t> p_cam, p_lp;
  std::vector<uint8_t> p_pose_fixed;
};

namespace {

thread_local std::string g_err;

constexpr double DEFAULT_FOCAL_LENGTH = 500.0;  // bal.rs:99-101

double normalize_focal_length(double f) { return (f > 0.0 && std::isfinite(f)) ? f : DEFAULT_FOCAL_LENGTH; }  // bal.rs:107-113

// Non-empty lines of the file, trimmed, with their 1-based line numbers (bal.rs:144-149).
struct Lines {
  const char* p;
  const char* end;
  size_t line_no = 0;
  bool next(const char*& b, const char*& e, size_t& no) {
    while (p < end) {
      const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
      const char* le = nl ? nl : end;
      const char* lb = p;
      p = nl ? nl + 1 : end;
      ++line_no;
      while (lb < le && (unsigned char)*lb <= ' ') ++lb;       // str::trim: ASCII whitespace on both sides
      while (le > lb && (unsigned char)le[-1] <= ' ') --le;
      if (lb == le) continue;
      b = lb; e = le; no = line_no;
      return true;
    }
    return false;
  }
};

// split_whitespace: up to `cap` fields; returns the number of fields on the line (may exceed cap)
int split(const char* b, const char* e, const char* fb[], const char* fe[], int cap) {
  int n = 0;
  while (b < e) {
    while (b < e && (unsigned char)*b <= ' ') ++b;
    if (b == e) break;
    const char* s = b;
    while (b < e && (unsigned char)*b > ' ') ++b;
    if (n < cap) { fb[n] = s; fe[n] = b; }
    ++n;
  }
  return n;
}

// str::parse::<f64>: the whole token must be a number (decimal, exponent, inf / infinity / nan in any case); no
// hexadecimal floats, no leading / trailing garbage.
bool parse_f64(const char* b, const char* e, double& out) {
  if (b == e || (size_t)(e - b) > 510) return false;
  char buf[512];
  memcpy(buf, b, (size_t)(e - b));
  buf[e - b] = 0;
  for (const char* c = buf; *c; ++c)
    if (*c == 'x' || *c == 'X' || *c == 'p' || *c == 'P' || *c == '(') return false;  // strtod extras Rust rejects
  char* endp = nullptr;
  out = strtod(buf, &endp);
  return endp == buf + (e - b);
}

// str::parse::<usize>: optional '+', decimal digits only
bool parse_usize(const char* b, const char* e, uint64_t& out) {
  if (b < e && *b == '+') ++b;
  if (b == e) return false;
  uint64_t v = 0;
  for (const char* c = b; c < e; ++c) {
    if (*c < '0' || *c > '9') return false;
    const uint64_t d = (uint64_t)(*c - '0');
    if (v > (std::numeric_limits<uint64_t>::max() - d) / 10) return false;
    v = v * 10 + d;
  }
  out = v;
  return true;
}

apex_status err_invalid_number(size_t line, const char* b, const char* e) {
  g_err = "Invalid number format at line " + std::to_string(line) + ": " + std::string(b, e);
  return APEX_ERR_INVALID_NUMBER;
}
apex_status err_missing_fields(size_t line) {
  g_err = "Missing required fields at line " + std::to_string(line);
  return APEX_ERR_MISSING_FIELDS;
}
apex_status err_parse(size_t line, const std::string& msg) {
  g_err = "Parse error at line " + std::to_string(line) + ": " + msg;
  return APEX_ERR_PARSE;
}

apex_status parse_bal(const char* data, size_t len, apex_bal_dataset& ds) {
  Lines L{data, data + len};
  const char *b, *e;
  size_t no;
  const char *fb[4], *fe[4];
  // header (bal.rs:203-240)
  if (!L.next(b, e, no)) return err_parse(1, "Missing header line");
  if (split(b, e, fb, fe, 4) != 3) return err_missing_fields(no);
  uint64_t hdr[3];
  for (int i = 0; i < 3; ++i)
    if (!parse_usize(fb[i], fe[i], hdr[i])) return err_invalid_number(no, fb[i], fe[i]);
  if (hdr[0] > 0xFFFFFFF0ull || hdr[1] > 0xFFFFFFF0ull || hdr[2] > 0xFFFFFFF0ull) return err_parse(no, "problem too large for u32 indices");
  ds.ncam = (uint32_t)hdr[0]; ds.npts = (uint32_t)hdr[1]; ds.nobs = hdr[2];
  // The header is untrusted: the arrays are sized by what the rest of the file can hold (an observation line is at least
  // "0 0 0 0" = 7 bytes, a parameter line at least 1 byte + its newline), not by the header's counts - a short file then ends
  // in the loops' own end-of-file errors (with the reference's line / index in the message) instead of a 100 GB allocation
  auto fits = [&](uint64_t want, uint64_t min_bytes_each) { return (size_t)std::min<uint64_t>(want, (uint64_t)(L.end - L.p) / min_bytes_each + 1); };
  // observations (bal.rs:243-303)
  { const size_t cap = fits(ds.nobs, 7); ds.obs_cam.resize(cap); ds.obs_pt.resize(cap); ds.obs_uv.resize(2 * cap); }
  for (uint64_t o = 0; o < ds.nobs; ++o) {
    if (!L.next(b, e, no)) return err_parse(0, "Unexpected end of file in observations section");
    if (split(b, e, fb, fe, 4) != 4) return err_missing_fields(no);
    uint64_t ci, pi;
    if (!parse_usize(fb[0], fe[0], ci)) return err_invalid_number(no, fb[0], fe[0]);
    if (!parse_usize(fb[1], fe[1], pi)) return err_invalid_number(no, fb[1], fe[1]);
    if (!parse_f64(fb[2], fe[2], ds.obs_uv[2 * o])) return err_invalid_number(no, fb[2], fe[2]);
    if (!parse_f64(fb[3], fe[3], ds.obs_uv[2 * o + 1])) return err_invalid_number(no, fb[3], fe[3]);
    // the reference indexes dataset.cameras[obs.camera_index] later (bin/bundle_adjustment.rs:405) and would panic;
    // here an index outside the header's counts is a parse error at the offending line
    if (ci >= ds.ncam || pi >= ds.npts) return err_parse(no, "observation index out of range");
    ds.obs_cam[o] = (uint32_t)ci; ds.obs_pt[o] = (uint32_t)pi;
  }
  // cameras: 9 consecutive lines each (bal.rs:306-349)
  ds.cameras.resize(fits((uint64_t)ds.ncam * 9, 2));
  for (uint32_t c = 0; c < ds.ncam; ++c)
    for (int k = 0; k < 9; ++k) {
      if (!L.next(b, e, no)) return err_parse(0, "Unexpected end of file in camera " + std::to_string(c) + " parameter " + std::to_string(k));
      if (!parse_f64(b, e, ds.cameras[(size_t)c * 9 + k])) return err_invalid_number(no, b, e);
    }
  for (uint32_t c = 0; c < ds.ncam; ++c) ds.cameras[(size_t)c * 9 + 6] = normalize_focal_length(ds.cameras[(size_t)c * 9 + 6]);
  // points: 3 consecutive lines each (bal.rs:352-392)
  ds.points.resize(fits((uint64_t)ds.npts * 3, 2));
  for (uint32_t p = 0; p < ds.npts; ++p)
    for (int k = 0; k < 3; ++k) {
      if (!L.next(b, e, no)) return err_parse(0, "Unexpected end of file in point " + std::to_string(p) + " coordinate " + std::to_string(k));
      if (!parse_f64(b, e, ds.points[(size_t)p * 3 + k])) return err_invalid_number(no, b, e);
    }
  return APEX_OK;
}

}  // namespace

extern "C" {

const char* apex_bal_last_error(void) { return g_err.c_str(); }

void apex_bal_free(apex_bal_dataset* ds) { delete ds; }

apex_status apex_bal_load(const char* path, apex_bal_dataset** out) {
  try {
    if (!path || !out) { g_err = "null argument"; return APEX_ERR_INVALID_INPUT; }
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) { g_err = std::string("IO error: ") + strerror(errno) + " (" + path + ")"; return APEX_ERR_IO; }
    std::string content;
    {
      char buf[1 << 16];
      size_t n;
      while ((n = fread(buf, 1, sizeof buf, f)) > 0) content.append(buf, n);
      const bool bad = ferror(f) != 0;
      fclose(f);
      if (bad) { g_err = std::string("IO error: read failed (") + path + ")"; return APEX_ERR_IO; }
    }
    apex_bal_dataset* ds = new (std::nothrow) apex_bal_dataset;
    if (!ds) { g_err = "out of memory"; return APEX_ERR_IO; }
    const apex_status st = parse_bal(content.data(), content.size(), *ds);
    if (st != APEX_OK) { delete ds; return st; }
    *out = ds;
    return APEX_OK;
  } catch (const std::bad_alloc&) { g_err = "out of memory"; return APEX_ERR_IO; } catch (const std::exception& ex) { g_err = std::string("internal error: ") + ex.what(); return APEX_ERR_IO; }
}

apex_status apex_bal_from_arrays(uint32_t ncam, uint32_t npts, uint64_t nobs, const double* cameras, const double* points,
                                 const uint32_t* obs_cam, const uint32_t* obs_pt, const double* obs_uv, apex_bal_dataset** out) {
  try {
    if (!out || (ncam && !cameras) || (npts && !points) || (nobs && (!obs_cam || !obs_pt || !obs_uv))) { g_err = "null argument"; return APEX_ERR_INVALID_INPUT; }
    for (uint64_t o = 0; o < nobs; ++o)
      if (obs_cam[o] >= ncam || obs_pt[o] >= npts) { g_err = "observation index out of range"; return APEX_ERR_INVALID_INPUT; }
    apex_bal_dataset* ds = new (std::nothrow) apex_bal_dataset;
    if (!ds) { g_err = "out of memory"; return APEX_ERR_IO; }
    ds->ncam = ncam; ds->npts = npts; ds->nobs = nobs;
    ds->cameras.assign(cameras, cameras + (size_t)ncam * 9);
    ds->points.assign(points, points + (size_t)npts * 3);
    ds->obs_cam.assign(obs_cam, obs_cam + nobs);
    ds->obs_pt.assign(obs_pt, obs_pt + nobs);
    ds->obs_uv.assign(obs_uv, obs_uv + 2 * nobs);
    *out = ds;
    return APEX_OK;
  } catch (const std::bad_alloc&) { g_err = "out of memory"; return APEX_ERR_IO; } catch (const std::exception& ex) { g_err = std::string("internal error: ") + ex.what(); return APEX_ERR_IO; }
}

apex_status apex_bal_view_get(const apex_bal_dataset* ds, apex_bal_view* v) {
  if (!ds || !v) { g_err = "null argument"; return APEX_ERR_INVALID_INPUT; }
  v->ncam = ds->ncam; v->npts = ds->npts; v->nobs = ds->nobs;
  v->cameras = ds->cameras.data(); v->points = ds->points.data();
  v->obs_cam = ds->obs_cam.data(); v->obs_pt = ds->obs_pt.data(); v->obs_uv = ds->obs_uv.data();
  return APEX_OK;
}

apex_status apex_bal_write(const apex_bal_dataset* ds, const char* path) {
  try {
    if (!ds || !path) { g_err = "null argument"; return APEX_ERR_INVALID_INPUT; }
    FILE* f = fopen(path, "wb");
    if (!f) { g_err = std::string("IO error: ") + strerror(errno) + " (" + path + ")"; return APEX_ERR_IO; }
    std::vector<char> big(1 << 20);
    setvbuf(f, big.data(), _IOFBF, big.size());
    fprintf(f, "%u %u %llu\n", ds->ncam, ds->npts, (unsigned long long)ds->nobs);
    for (uint64_t o = 0; o < ds->nobs; ++o)
      fprintf(f, "%u %u %.17g %.17g\n", ds->obs_cam[o], ds->obs_pt[o], ds->obs_uv[2 * o], ds->obs_uv[2 * o + 1]);
    for (double v : ds->cameras) fprintf(f, "%.17g\n", v);
    for (double v : ds->points) fprintf(f, "%.17g\n", v);
    const bool bad = ferror(f) != 0;
    if (fclose(f) != 0 || bad) { g_err = std::string("IO error: write failed (") + path + ")"; return APEX_ERR_IO; }
    return APEX_OK;
  } catch (const std::bad_alloc&) { g_err = "out of memory"; return APEX_ERR_IO; } catch (const std::exception& ex) { g_err = std::string("internal error: ") + ex.what(); return APEX_ERR_IO; }
}

apex_status apex_bal_build_problem(apex_bal_dataset* ds, uint64_t num_points, int32_t optimization_type, apex_problem_desc* out) {
  try {
    if (!ds || !out) { g_err = "null argument"; return APEX_ERR_INVALID_INPUT; }
    if (optimization_type != 0 && optimization_type != 1) {
      g_err = "only bundle-adjustment and self-calibration are functional optimization types (bin/bundle_adjustment.rs)";
      return APEX_ERR_UNSUPPORTED;
    }
    const uint32_t npts = (uint32_t)std::min<uint64_t>(num_points, ds->npts);  // bin/bundle_adjustment.rs:170-171
    const uint32_t ncam = ds->ncam;
    ds->p_pose.resize((size_t)ncam * 7);
    ds->p_intr.resize((size_t)ncam * 3);
    for (uint32_t c = 0; c < ncam; ++c) {
      const double* cam = &ds->cameras[(size_t)c * 9];
      // axis_angle_to_so3 (:200-208) + SO3::from_axis_angle = UnitQuaternion::from_axis_angle(Unit::new_normalize(axis), angle)
      const double angle = std::sqrt(cam[0] * cam[0] + cam[1] * cam[1] + cam[2] * cam[2]);
      double q[4] = {1.0, 0.0, 0.0, 0.0};
      if (!(angle < 1e-10)) {
        const double ax = cam[0] / angle, ay = cam[1] / angle, az = cam[2] / angle;
        const double n = std::sqrt(ax * ax + ay * ay + az * az);  // Unit::new_normalize renormalises the already unit axis
        const double s = std::sin(angle / 2.0), co = std::cos(angle / 2.0);
        q[0] = co; q[1] = ax / n * s; q[2] = ay / n * s; q[3] = az / n * s;
      }
      double* pose = &ds->p_pose[(size_t)c * 7];  // DVector::from(SE3): [t, qw, qx, qy, qz] (se3.rs:200-222)
      pose[0] = cam[3]; pose[1] = cam[4]; pose[2] = cam[5];
      pose[3] = q[0]; pose[4] = q[1]; pose[5] = q[2]; pose[6] = q[3];
      ds->p_intr[(size_t)c * 3 + 0] = cam[6]; ds->p_intr[(size_t)c * 3 + 1] = cam[7]; ds->p_intr[(size_t)c * 3 + 2] = cam[8];
    }
    ds->p_pt.assign(ds->points.begin(), ds->points.begin() + (size_t)npts * 3);
    ds->p_cam.clear(); ds->p_lp.clear(); ds->p_uv.clear();
    for (uint64_t o = 0; o < ds->nobs; ++o) {  // valid_obs: point_index < num_points, file order (:258-262)
      if (ds->obs_pt[o] >= npts) continue;
      ds->p_cam.push_back(ds->obs_cam[o]);
      ds->p_lp.push_back(ds->obs_pt[o]);
      ds->p_uv.push_back(ds->obs_uv[2 * o]);
      ds->p_uv.push_back(ds->obs_uv[2 * o + 1]);
    }
    ds->p_pose_fixed.assign(ncam, 0);
    if (ncam) ds->p_pose_fixed[0] = 0x3F;  // problem.fix_variable("pose_0000", 0..6) (:294-298)
    memset(out, 0, sizeof *out);
    out->camera_model = APEX_CAM_BAL;
    out->opt_flags = APEX_OPT_POSE | APEX_OPT_LANDMARK | (optimization_type == 1 ? APEX_OPT_INTRINSIC : 0u);
    out->intr_dim = 3;
    out->intr_vars_present = 1;  // intr_XXXX is always inserted (:243-246)
    out->ncam = ncam; out->npts = npts; out->nobs = ds->p_cam.size();
    out->pose = ds->p_pose.data(); out->intr = ds->p_intr.data(); out->pt = ds->p_pt.data();
    out->obs_cam = ds->p_cam.data(); out->obs_pt = ds->p_lp.data(); out->obs_uv = ds->p_uv.data();
    out->loss_id = APEX_LOSS_HUBER;  // HuberLoss::new(1.0) (:421-424)
    out->loss_params[0] = 1.0;
    out->pose_fixed = ds->p_pose_fixed.data();
    out->intr_fixed = nullptr; out->pt_fixed = nullptr;
    return APEX_OK;
  } catch (const std::bad_alloc&) { g_err = "out of memory"; return APEX_ERR_IO; } catch (const std::exception& ex) { g_err = std::string("internal error: ") + ex.what(); return APEX_ERR_IO; }
}

}  // extern "C"

// bundle_adjustment_cli.cpp — the reference's `bundle_adjustment` binary (bin/bundle_adjustment.rs) on the GPU path:
// same positional FILE and flags (-n/--num-points, -s/--solver, -t/--optimization-type, -v/--verbose), same problem
// construction (apex_bal_build_problem), LevenbergMarquardtConfig::for_bundle_adjustment(), same summary lines.
// Host C++ above the C ABI; links libapex_gpu.so. --cameras / --points (auto-download, :136-160) need network access
// and are rejected. There is no CPU fallback: without a CUDA device the run ends with the library's error.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/apex_gpu.h"

static const char* status_name(int s) {
  static const char* n[] = {"Converged", "MaxIterationsReached", "CostToleranceReached", "ParameterToleranceReached", "GradientToleranceReached",
                            "NumericalFailure", "UserTerminated", "Timeout", "TrustRegionRadiusTooSmall", "MinCostThresholdReached",
                            "IllConditionedJacobian", "InvalidNumericalValues", "Failed"};
  return (s >= 0 && s <= 12) ? n[s] : "?";
}

static int usage(const char* msg) {
  if (msg) fprintf(stderr, "error: %s\n\n", msg);
  fprintf(stderr,
          "Bundle adjustment optimization for BAL datasets\n\n"
          "Usage: bundle_adjustment [OPTIONS] <FILE>\n\n"
          "Options:\n"
          "  -n, --num-points <N>            Limit number of points (for testing)\n"
          "  -s, --solver <SOLVER>           explicit | implicit | matrix-free [default: implicit]\n"
          "                                  explicit    = SchurVariant::Sparse: dense S + Cholesky\n"
          "                                  implicit    = SchurVariant::Iterative AS THE REFERENCE DISPATCHES IT TODAY: explicit S +\n"
          "                                                scalar-Jacobi PCG (explicit_schur.rs:1222-1225), same trajectory as the crate\n"
          "                                  matrix-free = the math of IterativeSchurSolver (implicit_schur.rs): S never formed, block PCG\n"
          "                                                with the Schur-Jacobi preconditioner (what bench.py times; another truncated-PCG\n"
          "                                                trajectory than `implicit`)\n"
          "  -t, --optimization-type <TYPE>  bundle-adjustment | self-calibration | only-pose | only-landmarks | only-intrinsics\n"
          "                                  [default: self-calibration]\n"
          "  -v, --verbose                   Verbose output\n"
          "      --device <ID>               CUDA device ordinal [default: 0]\n");
  return 2;
}

int main(int argc, char** argv) {
  std::string file, solver = "implicit", type = "self-calibration";
  long long num_points = -1;
  bool verbose = false;
  int device = 0;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto value = [&](std::string& dst) -> bool {
      const size_t eq = a.find('=');
      if (eq != std::string::npos) { dst = a.substr(eq + 1); return true; }
      if (i + 1 >= argc) return false;
      dst = argv[++i];
      return true;
    };
    std::string v;
    if (a == "-h" || a == "--help") { usage(nullptr); return 0; }
    else if (a == "-v" || a == "--verbose") verbose = true;
    else if (a == "-n" || a.rfind("--num-points", 0) == 0) { if (!value(v)) return usage("--num-points needs a value"); num_points = atoll(v.c_str()); }
    else if (a == "-s" || a.rfind("--solver", 0) == 0) { if (!value(solver)) return usage("--solver needs a value"); }
    else if (a == "-t" || a.rfind("--optimization-type", 0) == 0) { if (!value(type)) return usage("--optimization-type needs a value"); }
    else if (a.rfind("--device", 0) == 0) { if (!value(v)) return usage("--device needs a value"); device = atoi(v.c_str()); }
    else if (a.rfind("--cameras", 0) == 0 || a.rfind("--points", 0) == 0) { value(v); fprintf(stderr, "warning: %s (auto-download) is not available: no network\n", a.c_str()); }
    else if (!a.empty() && a[0] == '-') return usage(("unexpected argument '" + a + "'").c_str());
    else if (file.empty()) file = a;
    else return usage("more than one FILE");
  }
  if (file.empty()) return usage("the following required arguments were not provided: <FILE>");
  int variant;
  if (solver == "explicit") variant = APEX_SCHUR_EXPLICIT;       // SchurVariant::Sparse (:36-43)
  else if (solver == "implicit") variant = APEX_SCHUR_EXPLICIT_PCG;  // SchurVariant::Iterative -> solve_with_pcg on the explicit S (explicit_schur.rs:1222-1225)
  else if (solver == "matrix-free") variant = APEX_SCHUR_IMPLICIT;   // IterativeSchurSolver's matrix-free block PCG (not reachable from the reference CLI)
  else return usage(("invalid value '" + solver + "' for '--solver <SOLVER>'").c_str());
  int opt_type;
  if (type == "bundle-adjustment") opt_type = 0;
  else if (type == "self-calibration") opt_type = 1;
  else if (type == "only-pose" || type == "only-landmarks" || type == "only-intrinsics") opt_type = 2;
  else return usage(("invalid value '" + type + "' for '--optimization-type <TYPE>'").c_str());

  printf("APEX-SOLVER BUNDLE ADJUSTMENT (B200 path)\n\n");
  printf("Loading BAL dataset: %s\n", file.c_str());
  const auto t_load = std::chrono::steady_clock::now();
  apex_bal_dataset* ds = nullptr;
  if (apex_bal_load(file.c_str(), &ds) != APEX_OK) {
    fprintf(stderr, "Error: %s\n", apex_bal_last_error());
    return 1;
  }
  const double load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count();
  apex_bal_view v;
  apex_bal_view_get(ds, &v);
  const uint64_t use = num_points < 0 ? v.npts : std::min<uint64_t>((uint64_t)num_points, v.npts);
  printf("Dataset statistics:\n  Cameras: %u\n  Total points: %u\n  Points to use: %llu\n  Observations: %llu\n  Load time: %.3fs\n\n", v.ncam, v.npts,
         (unsigned long long)use, (unsigned long long)v.nobs, load_s);
  apex_problem_desc desc;
  if (apex_bal_build_problem(ds, use, opt_type, &desc) != APEX_OK) {
    fprintf(stderr, "Error: %s\n", apex_bal_last_error());
    apex_bal_free(ds);
    return 1;
  }
  printf("Adding %u cameras as SE3 poses + intrinsics...\nAdding %u landmarks as RN(3) variables...\nAdding %llu projection factors (optimization: %s)...\n",
         desc.ncam, desc.npts, (unsigned long long)desc.nobs, type.c_str());
  printf("Fixing first camera pose (all 6 DOF) for gauge freedom...\n");
  apex_lm_config cfg;
  apex_lm_config_for_bundle_adjustment(&cfg);
  cfg.schur_variant = variant;
  printf("\nSolver configuration:\n  Solver variant: %s\n  Optimization type: %s\n  Linear solver: SparseSchurComplement\n  Preconditioner: %s\n", solver.c_str(),
         type.c_str(), cfg.schur_preconditioner == APEX_PRECOND_SCHUR_JACOBI ? "SchurJacobi" : cfg.schur_preconditioner == APEX_PRECOND_BLOCK_DIAGONAL ? "BlockDiagonal" : "None");
  const uint64_t pose_dof = (uint64_t)desc.ncam * 6, intr_dof = (uint64_t)desc.ncam * 3, lm_dof = (uint64_t)desc.npts * 3;
  printf("\nDiagnostics:\n  Cameras: %u\n  Number of factors (observations): %llu\n  Pose DOF: %llu (6 per camera)\n  Intrinsic DOF: %llu (3 per camera)\n"
         "  Landmark DOF: %llu\n  Total DOF: %llu\n  DOF per observation: %.2f\n",
         desc.ncam, (unsigned long long)desc.nobs, (unsigned long long)pose_dof, (unsigned long long)intr_dof, (unsigned long long)lm_dof,
         (unsigned long long)(pose_dof + intr_dof + lm_dof), (double)(pose_dof + intr_dof + lm_dof) / (double)std::max<uint64_t>(desc.nobs, 1));

  apex_ctx_desc cd;
  memset(&cd, 0, sizeof cd);
  cd.device = device; cd.rank = 0; cd.nranks = 1;
  apex_ctx* ctx = nullptr;
  apex_status st = apex_ctx_create(&cd, &ctx);
  if (st != APEX_OK) {
    fprintf(stderr, "Error: cannot create a GPU context (status %d): %s\n", st, ctx ? apex_last_error(ctx) : "no CUDA device");
    apex_bal_free(ds);
    return 1;
  }
  printf("\nStarting optimization...\n");
  const auto t0 = std::chrono::steady_clock::now();
  st = apex_problem_upload(ctx, &desc);
  std::vector<apex_iter_trace> trace(std::max(cfg.max_iterations, 1));
  apex_lm_result res;
  memset(&res, 0, sizeof res);
  if (st == APEX_OK) st = apex_lm_solve(ctx, &cfg, &res, trace.data(), (int)trace.size());
  const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (st != APEX_OK) {
    fprintf(stderr, "Error: optimization failed (status %d): %s\n", st, apex_last_error(ctx));
    apex_ctx_destroy(ctx);
    apex_bal_free(ds);
    return 1;
  }
  printf("\nOptimization completed!\nStatus: %s\nIterations: %d\nTime: %.2f seconds\n", status_name(res.status), res.iterations, elapsed);
  const double nobs = (double)std::max<uint64_t>(desc.nobs, 1);
  printf("\nMetrics:\n  Initial cost: %.6e\n  Final cost: %.6e\n  Initial RMSE: %.3f pixels\n  Final RMSE: %.3f pixels\n  Improvement: %.2f%%\n", res.initial_cost,
         res.final_cost, std::sqrt(res.initial_cost / nobs), std::sqrt(res.final_cost / nobs), (res.initial_cost - res.final_cost) / res.initial_cost * 100.0);
  if (verbose) {
    printf("\n  Per-iteration: %.2fs\n", elapsed / std::max(res.iterations, 1));
    for (int i = 0; i < res.iterations && i < (int)trace.size(); ++i)
      printf("  iter %3d  cost %.9e  accepted %d  linear iterations %d\n", trace[i].iteration, trace[i].cost, trace[i].accepted, trace[i].ls_iter);
  }
  apex_ctx_destroy(ctx);
  apex_bal_free(ds);
  return 0;
}

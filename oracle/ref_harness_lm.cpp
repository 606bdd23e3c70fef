// oracle/ref_harness_lm.cpp -- TEST INFRASTRUCTURE.  g2o's OWN Levenberg-Marquardt control code and robust kernel,
// executed: Thirdparty/g2o/g2o/core/{optimization_algorithm_levenberg.cpp, robust_kernel.cpp, robust_kernel_impl.cpp}
// are compiled unmodified (oracle/Makefile target `ref`; the headers they include besides their own class declarations
// are shadowed by oracle/g2o_lm_stub through `-I... -I-`) and OptimizationAlgorithmLevenberg::solve() drives the
// oracle's linear algebra through the two abstract faces it talks to:
//     g2o::Solver           -> oracle::Problem::{build_system, set_lambda, solve_schur, restore_diagonal, x, b}
//     g2o::SparseOptimizer  -> oracle::Problem::{compute_active_errors, active_robust_chi2, apply_update, terminate}
//                              + push / pop / discardTop on the estimates, indexMapping() over the free vertices.
// ref_lm_local_ba() is oracle_local_ba() with ONE difference: the per-iteration LM step is the reference's solve()
// instead of oracle::Problem::lm_solve().  Equal traces and results on the same windows (tests/test_oracle_lm_vs_ref.py)
// pin the oracle's restatement of solve(), computeLambdaInit() and computeScale()
// (optimization_algorithm_levenberg.cpp:61-189): lambda schedule, gain ratio, accept / reject, trial limit, stop rules.
// What this file restates itself: the three-line loop of SparseOptimizer::optimize (sparse_optimizer.cpp:376-395) and,
// copied from oracle_local_ba, the stage driver of Optimizer.cpp:2643-2701.
#include "lba.cpp"  // the oracle's Problem (this library is test infrastructure around test infrastructure)

#include "optimization_algorithm_levenberg.h"  // the reference's class declaration
#include "robust_kernel_impl.h"

namespace {

struct DiagVertex : g2o::OptimizableGraph::Vertex {
    const double* diag0;  // address of this vertex's first diagonal entry
    size_t stride;        // distance between consecutive diagonal entries
    int dim;
    DiagVertex(const double* d, size_t s, int n) : diag0(d), stride(s), dim(n) {}
    int dimension() const override { return dim; }
    double hessian(int i, int j) const override { return i == j ? diag0[stride * (size_t)i] : 0.0; }
};

struct OracleOptimizer : g2o::SparseOptimizer {
    oracle::Problem& pb;
    std::vector<DiagVertex> store;
    g2o::OptimizableGraph::VertexContainer mapping;
    mutable std::vector<double> chi_log;  // every activeRobustChi2() of the current solve()
    int discards = 0;
    explicit OracleOptimizer(oracle::Problem& p) : pb(p) {}
    // the vertices with a block in the Hessian, in g2o's order: non-marginalised ones by id (PVR 2 mnId, Bias 2 mnId + 1),
    // then the marginalised ones (the active map points).  Rebuilt after initializeOptimization().
    void build_mapping() {
        store.clear();
        mapping.clear();
        const size_t n = (size_t)pb.n;
        for (int k = 0; k < pb.K; ++k) {
            if (!pb.kf_free(k)) continue;
            store.emplace_back(&pb.Hpp[(size_t)pb.off_pvr(k) * n + pb.off_pvr(k)], n + 1, 9);
            store.emplace_back(&pb.Hpp[(size_t)pb.off_bias(k) * n + pb.off_bias(k)], n + 1, 6);
        }
        for (int p = 0; p < pb.P; ++p)
            if (pb.pt_active[p]) store.emplace_back(&pb.Hll[9 * (size_t)p], 4, 3);
        for (DiagVertex& v : store) mapping.push_back(&v);
    }
    void computeActiveErrors() override { pb.compute_active_errors(); }
    double activeRobustChi2() const override {
        const double c = pb.active_robust_chi2();
        chi_log.push_back(c);
        return c;
    }
    void push() override { pb.ns_backup = pb.ns, pb.pts_backup = pb.pts; }
    void pop() override { pb.ns = pb.ns_backup, pb.pts = pb.pts_backup; }  // estimates only: cached errors stay stale
    void discardTop() override { ++discards; }
    void update(const double* x) override {
        if (x != pb.x.data()) std::abort();
        pb.apply_update();
    }
    bool terminate() override { return pb.terminate(); }
    const g2o::OptimizableGraph::VertexContainer& indexMapping() const override { return mapping; }
};

struct OracleSolver : g2o::Solver {
    oracle::Problem& pb;
    explicit OracleSolver(oracle::Problem& p) : pb(p) {}
    bool buildStructure(bool) override { return true; }
    bool buildSystem() override {
        pb.build_system();
        return true;
    }
    bool setLambda(double lambda, bool) override {
        pb.currentLambda = lambda;  // (Problem::compute_scale is not used on this path; kept consistent anyway)
        pb.set_lambda(lambda);
        return true;
    }
    void restoreDiagonal() override { pb.restore_diagonal(); }
    bool solve() override { return pb.solve_schur(); }
    double* x() override { return pb.x.data(); }
    double* b() override { return pb.b.data(); }
    size_t vectorSize() const override { return pb.x.size(); }
    bool schur() override { return true; }
};

}  // namespace

extern "C" {

// oracle_local_ba with g2o's own OptimizationAlgorithmLevenberg::solve() as the LM step (same signature)
int ref_lm_local_ba(const vilba_window* win, const vilba_params* params, vilba_result* out, const volatile uint8_t* stop_flag) {
    if (!out) return VILBA_ERR_ARG;
    int st = check_window(win);
    out->status = st, out->n_trace = 0, out->stage2_ran = 0, out->n_outliers_stage1 = 0, out->solve_ms = 0.0;
    if (st != VILBA_OK) return st;
    vilba_params prm;
    if (params) prm = *params; else vilba_default_params_oracle(&prm);
    const bool single_stage = (prm.mode & VILBA_MODE_SINGLE_STAGE) != 0;
    if (!single_stage && stop_flag && *stop_flag) {
        out->status = VILBA_ABORTED;
        return VILBA_ABORTED;
    }
    oracle::Problem pb(*win, prm, stop_flag);
    if (prm.mode & VILBA_MODE_MONO_NOT_ROBUST) std::fill(pb.mono_robust.begin(), pb.mono_robust.end(), 0);
    OracleOptimizer opt(pb);
    OracleSolver solver(pb);
    solver.setOptimizer(&opt);
    // ONE algorithm object for the whole call, like `solver` in Optimizer.cpp:2405-2412: _ni, _nBad, _currentLambda
    // survive from optimize(5) into optimize(10) exactly as far as solve(0) does not reset them
    g2o::OptimizationAlgorithmLevenberg lm(&solver);
    lm.setOptimizer(&opt);
    lm.setMaxTrialsAfterFailure(prm.max_trials);
    auto run = [&](int iterations, int stage, int n_active) {
        opt.build_mapping();
        bool ok = true;
        for (int i = 0; i < iterations && !pb.terminate() && ok; ++i) {
            opt.chi_log.clear();
            opt.discards = 0;
            const g2o::OptimizationAlgorithm::SolverResult result = lm.solve(i, false);
            ok = result == g2o::OptimizationAlgorithm::OK;
            vilba_iter_record rec;
            std::memset(&rec, 0, sizeof(rec));
            rec.stage = stage, rec.iteration = i, rec.n_active_edges = n_active;
            rec.trials = lm.levenbergIteration();
            rec.accepted = opt.discards > 0 ? 1 : 0;
            rec.result = ok ? 0 : 1;
            rec.lambda = lm.currentLambda();
            rec.chi2_initial = opt.chi_log.front();
            // currentChi at the end of solve(): the chi2 of the accepted trial (the last one evaluated), else unchanged
            rec.chi2_final = opt.discards > 0 ? opt.chi_log.back() : opt.chi_log.front();
            if (out->n_trace < VILBA_MAX_TRACE) out->trace[out->n_trace++] = rec;
        }
    };
    int n_active = pb.initialize_optimization();
    run(prm.iters_stage1, 1, n_active);
    const bool do_more = !single_stage && !(stop_flag && *stop_flag);
    if (do_more) {
        for (int p_ = 0; p_ < pb.P; ++p_)
            for (int e = win->pt_obs_begin[p_]; e < win->pt_obs_begin[p_ + 1]; ++e) {
                if (pb.chi2_mono(e) > prm.chi2_gate || !mono_depth_positive(pb.ns[win->obs_kf[e]], pb.pts[p_], pb.calib)) {
                    pb.mono_level[e] = 1;
                    out->n_outliers_stage1++;
                }
                pb.mono_robust[e] = 0;
            }
        n_active = pb.initialize_optimization();
        run(prm.iters_stage2, 2, n_active);
        out->stage2_ran = 1;
    }
    for (int p_ = 0; p_ < pb.P; ++p_)
        for (int e = win->pt_obs_begin[p_]; e < win->pt_obs_begin[p_ + 1]; ++e) {
            const double c2 = pb.chi2_mono(e);
            const bool bad = c2 > prm.chi2_gate || !mono_depth_positive(pb.ns[win->obs_kf[e]], pb.pts[p_], pb.calib);
            if (out->obs_outlier) out->obs_outlier[e] = bad ? 1 : 0;
            if (out->obs_chi2) out->obs_chi2[e] = c2;
        }
    if (out->kf_state)
        for (int k = 0; k < pb.K; ++k) pb.ns[k].store(out->kf_state + VILBA_NS_DOUBLES * k);
    if (out->pt_xyz)
        for (int p_ = 0; p_ < pb.P; ++p_) pb.pts[p_].store(out->pt_xyz + 3 * p_);
    out->status = VILBA_OK;
    return VILBA_OK;
}

// RobustKernelHuber::robustify (robust_kernel_impl.cpp:78-91) with setDelta(delta): rho[3]
void ref_huber(double e2, double delta, double rho[3]) {
    g2o::RobustKernelHuber k;
    k.setDelta(delta);
    Eigen::Vector3d r;
    k.robustify(e2, r);
    rho[0] = r[0], rho[1] = r[1], rho[2] = r[2];
}
}

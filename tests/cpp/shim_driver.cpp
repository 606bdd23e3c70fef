// shim_driver.cpp -- exercises the reference-shaped C++ interface (mc_slam_b200/shim/vilba_shim.h) the way
// LocalMapping would: builds KeyFrame / MapPoint objects from a window blob, calls
// Optimizer::LocalBundleAdjustmentNavState and KeyFrame::ComputePreInt, dumps what they wrote back.
// Usage: shim_driver lba <in.bin> <out.bin> [stop] | shim_driver preint <in.bin> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>

#include "../../mc_slam_b200/shim/vilba_shim.h"

using namespace ORB_SLAM2;

template <typename T>
static std::vector<T> rd(FILE* f, size_t n) {
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) {
        fprintf(stderr, "short read\n");
        exit(2);
    }
    return v;
}
template <typename T>
static void wr(FILE* f, const std::vector<T>& v) {
    if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
}

// gba < 0: local BA; otherwise the global BA with nLoopKF = gba
static int run_lba(const char* in, const char* out, bool stop_before, long gba = -1, bool robust = false) {
    FILE* f = fopen(in, "rb");
    if (!f) return 2;
    auto hdr = rd<int32_t>(f, 4);
    const int K = hdr[0], NI = hdr[1], P = hdr[2], E = hdr[3];
    auto kf_state = rd<double>(f, (size_t)22 * K);
    auto kf_flags = rd<uint8_t>(f, K);
    auto kf_id = rd<int64_t>(f, K);
    auto imu_i = rd<int32_t>(f, NI), imu_j = rd<int32_t>(f, NI);
    auto preint = rd<double>(f, (size_t)142 * NI);
    auto pt_xyz = rd<double>(f, (size_t)3 * P);
    auto pt_begin = rd<int32_t>(f, P + 1);
    auto obs_kf = rd<int32_t>(f, E);
    auto obs_uv = rd<float>(f, (size_t)2 * E);
    auto obs_is2 = rd<float>(f, E);
    auto cal = rd<double>(f, 4 + 9 + 3 + 3);
    fclose(f);

    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) ConfigParam::EigTbc(r, c) = cal[4 + 3 * r + c];
        ConfigParam::EigTbc(r, 3) = cal[13 + r];
    }
    shim::MatF gw(3, 1);
    for (int i = 0; i < 3; ++i) gw.at(i) = (float)cal[16 + i];

    std::vector<std::unique_ptr<KeyFrame>> kfs;
    for (int k = 0; k < K; ++k) {
        kfs.emplace_back(new KeyFrame((unsigned long)kf_id[k], 0.2 * k, (float)cal[0], (float)cal[1], (float)cal[2], (float)cal[3]));
        const double* s = &kf_state[(size_t)22 * k];
        NavState ns;
        ns.Set_Pos(Vector3d(s[0], s[1], s[2]));
        ns.Set_Vel(Vector3d(s[3], s[4], s[5]));
        shim::Quat q;
        q.w = s[6], q.x = s[7], q.y = s[8], q.z = s[9];
        ns.Set_Rot(q);
        ns.Set_BiasGyr(Vector3d(s[10], s[11], s[12]));
        ns.Set_BiasAcc(Vector3d(s[13], s[14], s[15]));
        ns.Set_DeltaBiasGyr(Vector3d(s[16], s[17], s[18]));
        ns.Set_DeltaBiasAcc(Vector3d(s[19], s[20], s[21]));
        kfs[k]->SetNavState(ns);
    }
    for (int e = 0; e < NI; ++e) {
        kfs[imu_j[e]]->SetPrevKeyFrame(kfs[imu_i[e]].get());
        kfs[imu_j[e]]->IMUPreIntRef().setRaw(&preint[(size_t)142 * e]);
    }
    std::vector<std::unique_ptr<MapPoint>> mps;
    std::vector<std::pair<KeyFrame*, MapPoint*>> edge_owner(E);
    for (int p = 0; p < P; ++p) {
        mps.emplace_back(new MapPoint(1000 + p));
        shim::MatF pos(3, 1);
        for (int d = 0; d < 3; ++d) pos.at(d) = (float)pt_xyz[3 * p + d];
        mps[p]->SetWorldPos(pos);
        for (int e = pt_begin[p]; e < pt_begin[p + 1]; ++e) {
            KeyFrame* kf = kfs[obs_kf[e]].get();
            int oct = -1;
            for (size_t l = 0; l < kf->mvInvLevelSigma2.size(); ++l)
                if (kf->mvInvLevelSigma2[l] == obs_is2[e]) oct = (int)l;
            if (oct < 0) {
                kf->mvInvLevelSigma2.push_back(obs_is2[e]);
                oct = (int)kf->mvInvLevelSigma2.size() - 1;
            }
            KeyPoint kp;
            kp.pt.x = obs_uv[2 * e], kp.pt.y = obs_uv[2 * e + 1], kp.octave = oct;
            const size_t idx = kf->mvKeysUn.size();
            kf->mvKeysUn.push_back(kp);
            kf->mvuRight.push_back(-1.f);
            kf->mvpMapPoints.push_back(mps[p].get());
            mps[p]->AddObservation(kf, idx);
            edge_owner[e] = {kf, mps[p].get()};
        }
    }
    std::list<KeyFrame*> local;
    for (int k = 0; k < K; ++k)
        if (!(kf_flags[k] & VILBA_KF_FIXED)) local.push_back(kfs[k].get());
    Map map;
    LocalMapping lm;
    bool stop = stop_before;
    vilba_result trace;
    if (gba < 0)
        Optimizer::LocalBundleAdjustmentNavState(local.back(), local, &stop, &map, gw, &lm, &trace);
    else {
        for (auto& k : kfs) map.AddKeyFrame(k.get());
        for (auto& m : mps) map.AddMapPoint(m.get());
        Optimizer::GlobalBundleAdjustmentNavState(&map, gw, 10, &stop, (unsigned long)gba, robust, &trace);
    }

    FILE* o = fopen(out, "wb");
    std::vector<double> st((size_t)22 * K);
    std::vector<float> tcw((size_t)16 * K), pw((size_t)3 * P);
    for (int k = 0; k < K; ++k) {
        kfs[k]->GetNavState().toFlat(&st[(size_t)22 * k]);
        shim::MatF T = kfs[k]->GetPose();
        for (int i = 0; i < 16; ++i) tcw[(size_t)16 * k + i] = T.empty() ? 0.f : T.d[i];
    }
    for (int p = 0; p < P; ++p)
        for (int d = 0; d < 3; ++d) pw[3 * p + d] = mps[p]->GetWorldPos().at(d);
    std::vector<uint8_t> erased(E);
    for (int e = 0; e < E; ++e) erased[e] = edge_owner[e].second->GetObservations().count(edge_owner[e].first) ? 0 : 1;
    const bool have_trace = !stop_before || gba >= 0;  // the local BA returns before touching *pTrace
    std::vector<int32_t> meta = {lm.mbMapUpdateFlagForTracking ? 1 : 0, have_trace ? trace.n_trace : 0,
                                 have_trace ? trace.stage2_ran : 0, P ? mps[0]->mnNormalUpdates : 0};
    std::vector<double> chi(64, 0.0);
    if (have_trace)
        for (int i = 0; i < trace.n_trace && i < 64; ++i) chi[i] = trace.trace[i].chi2_final;
    wr(o, meta), wr(o, st), wr(o, tcw), wr(o, pw), wr(o, erased), wr(o, chi);
    if (gba >= 0) {  // the ...GBA members (Optimizer.cpp:1643-1665)
        std::vector<double> stg((size_t)22 * K, 0.0);
        std::vector<float> tcwg((size_t)16 * K, 0.f), pg((size_t)3 * P, 0.f);
        std::vector<int64_t> tag(K + P);
        for (int k = 0; k < K; ++k) {
            kfs[k]->mNavStateGBA.toFlat(&stg[(size_t)22 * k]);
            for (int i = 0; i < 16 && !kfs[k]->mTcwGBA.empty(); ++i) tcwg[(size_t)16 * k + i] = kfs[k]->mTcwGBA.d[i];
            tag[k] = (int64_t)kfs[k]->mnBAGlobalForKF;
        }
        for (int p = 0; p < P; ++p) {
            for (int d = 0; d < 3 && !mps[p]->mPosGBA.empty(); ++d) pg[3 * p + d] = mps[p]->mPosGBA.at(d);
            tag[K + p] = (int64_t)mps[p]->mnBAGlobalForKF;
        }
        wr(o, stg), wr(o, tcwg), wr(o, pg), wr(o, tag);
    }
    fclose(o);
    return 0;
}

// preint: N pairs; per pair the previous key-frame's biases and time stamp, this key-frame's time stamp and
// its IMU samples (g, a, t).  Runs KeyFrame::ComputePreInt (lazy, one launch per key-frame) and
// ComputePreIntBatch (one launch for all) and dumps both.
static int run_preint(const char* in, const char* out) {
    FILE* f = fopen(in, "rb");
    if (!f) return 2;
    const int N = rd<int32_t>(f, 1)[0];
    auto begin = rd<int32_t>(f, N + 1);
    auto bias = rd<double>(f, (size_t)6 * N);
    auto tprev = rd<double>(f, N), tcur = rd<double>(f, N);
    const int S = begin[N];
    auto g = rd<double>(f, (size_t)3 * S), a = rd<double>(f, (size_t)3 * S), t = rd<double>(f, S);
    fclose(f);
    std::vector<std::unique_ptr<KeyFrame>> prev, cur;
    for (int p = 0; p < N; ++p) {
        prev.emplace_back(new KeyFrame(2 * p + 1, tprev[p], 1, 1, 0, 0));
        cur.emplace_back(new KeyFrame(2 * p + 2, tcur[p], 1, 1, 0, 0));
        NavState ns;
        ns.Set_BiasGyr(Vector3d(bias[6 * p], bias[6 * p + 1], bias[6 * p + 2]));
        ns.Set_BiasAcc(Vector3d(bias[6 * p + 3], bias[6 * p + 4], bias[6 * p + 5]));
        prev[p]->SetNavState(ns);
        cur[p]->SetPrevKeyFrame(prev[p].get());
        for (int s = begin[p]; s < begin[p + 1]; ++s)
            cur[p]->mvIMUData.emplace_back(g[3 * s], g[3 * s + 1], g[3 * s + 2], a[3 * s], a[3 * s + 1], a[3 * s + 2], t[s]);
    }
    std::vector<double> single((size_t)142 * N), batch((size_t)142 * N);
    for (int p = 0; p < N; ++p) {
        cur[p]->ComputePreInt();
        const IMUPreintegrator& pi = cur[p]->GetIMUPreInt();
        const Vector3d dP = pi.getDeltaP();  // first getter integrates on the GPU
        std::memcpy(&single[(size_t)142 * p], pi.raw(), sizeof(double) * 142);
        if (dP[0] != single[(size_t)142 * p]) return 3;
    }
    std::vector<KeyFrame*> all;
    for (auto& k : cur) all.push_back(k.get());
    ComputePreIntBatch(all);
    for (int p = 0; p < N; ++p) std::memcpy(&batch[(size_t)142 * p], cur[p]->GetIMUPreInt().raw(), sizeof(double) * 142);
    FILE* o = fopen(out, "wb");
    wr(o, single), wr(o, batch);
    fclose(o);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 4) return 1;
    try {
        if (std::string(argv[1]) == "lba") return run_lba(argv[2], argv[3], argc > 4);
        if (std::string(argv[1]) == "gba")  // gba in out nLoopKF robust [stop]
            return run_lba(argv[2], argv[3], argc > 6, std::atol(argv[4]), std::atoi(argv[5]) != 0);
        if (std::string(argv[1]) == "preint") return run_preint(argv[2], argv[3]);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 4;
    }
    return 1;
}

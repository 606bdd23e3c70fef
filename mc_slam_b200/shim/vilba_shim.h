// vilba_shim.h -- host-side C++ mirror of the reference interface for the accelerated path.
//
// Same class and method names, argument meaning and error behaviour as mc275/MC_SLAM for exactly the
// members that Optimizer::LocalBundleAdjustmentNavState and IMUPreintegrator touch:
//   Optimizer::LocalBundleAdjustmentNavState   include/Optimizer.h:44-46, src/Optimizer.cpp:2320-2771
//   IMUPreintegrator                           src/IMU/IMUPreintegrator.h:17-200
//   NavState                                   src/IMU/NavState.h:14-140
//   KeyFrame (NavState / IMU / BA members)     include/KeyFrame.h:97-108,175-183,251-294; src/KeyFrame.cpp:96-252
//   MapPoint (position / observations)         include/MapPoint.h:28,42-62,103,148
//   Converter                                  src/Converter.cpp:110-160
// The arithmetic is NOT here: every getter of IMUPreintegrator and the optimiser body end in the C ABI of
// include/vilba.h (CUDA kernels).  The small Eigen / OpenCV value types the signatures mention are replaced
// by stand-ins (shim::Vec3 / Mat3 / Mat9 / MatF) because neither library is available in this image; a
// maintainer building inside MC_SLAM keeps the real KeyFrame/MapPoint classes and only takes
// LocalBundleAdjustmentNavState_flatten()/writeback() -- see INTEGRATION.md.
#ifndef VILBA_SHIM_H
#define VILBA_SHIM_H

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <stdexcept>
#include <vector>

#include "../../include/vilba.h"

namespace shim {
struct Vec3 {
    double v[3] = {0, 0, 0};
    Vec3() {}
    Vec3(double x, double y, double z) { v[0] = x, v[1] = y, v[2] = z; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    Vec3 operator-(const Vec3& o) const { return Vec3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
struct Mat3 {  // row-major
    double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double& operator()(int r, int c) { return m[3 * r + c]; }
    double operator()(int r, int c) const { return m[3 * r + c]; }
};
struct Mat9 {
    double m[81] = {0};
    double& operator()(int r, int c) { return m[9 * r + c]; }
    double operator()(int r, int c) const { return m[9 * r + c]; }
};
struct Quat {  // unit quaternion (w,x,y,z) -- Sophus::SO3 stand-in
    double w = 1, x = 0, y = 0, z = 0;
};
// CV_32F matrix stand-in (cv::Mat is only float storage on this path, Converter.cpp:110-160)
struct MatF {
    int rows = 0, cols = 0;
    std::vector<float> d;
    MatF() {}
    MatF(int r, int c) : rows(r), cols(c), d((size_t)r * c, 0.f) {}
    float& at(int r, int c = 0) { return d[(size_t)r * cols + c]; }
    float at(int r, int c = 0) const { return d[(size_t)r * cols + c]; }
    bool empty() const { return d.empty(); }
};
inline Mat3 quat_to_matrix(const Quat& q) {
    Mat3 R;
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R(0, 0) = 1 - (tyy + tzz), R(0, 1) = txy - twz, R(0, 2) = txz + twy;
    R(1, 0) = txy + twz, R(1, 1) = 1 - (txx + tzz), R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy, R(2, 1) = tyz + twx, R(2, 2) = 1 - (txx + tyy);
    return R;
}
inline Quat matrix_to_quat(const Mat3& m) {  // layout conversion only (Set_Rot(Matrix3d), NavState.h:52-55)
    Quat q;
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (m(2, 1) - m(1, 2)) * t, q.y = (m(0, 2) - m(2, 0)) * t, q.z = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (m(k, j) - m(j, k)) * t;
        v[j] = (m(j, i) + m(i, j)) * t, v[k] = (m(k, i) + m(i, k)) * t;
        q.x = v[0], q.y = v[1], q.z = v[2];
    }
    const double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    q.w /= n, q.x /= n, q.y /= n, q.z /= n;
    return q;
}

// one vilba context per calling thread (the reference calls this path from the single LocalMapping thread)
inline vilba_ctx* context() {
    static thread_local vilba_ctx* ctx = nullptr;
    if (!ctx) {
        ctx = vilba_create(0, nullptr);
        if (!ctx) throw std::runtime_error("vilba: no CUDA device -- the VI local-BA path has no CPU fallback");
    }
    return ctx;
}
}  // namespace shim

namespace ORB_SLAM2 {

typedef shim::Vec3 Vector3d;
typedef shim::Mat3 Matrix3d;
typedef shim::Mat9 Matrix9d;

// ---------------------------------------------------------------------------------------------------
class IMUData {  // src/IMU/imudata.h
public:
    IMUData(double gx, double gy, double gz, double ax, double ay, double az, double t) : _g(gx, gy, gz), _a(ax, ay, az), _t(t) {}
    Vector3d _g, _a;
    double _t;
    static double getGyrBiasRW2() { return 2.0e-5 * 2.0e-5; }  // imudata.cpp:25
    static double getAccBiasRW2() { return 5.0e-3 * 5.0e-3; }  // imudata.cpp:26
};

// ---------------------------------------------------------------------------------------------------
// IMUPreintegrator: value type with the reference's interface.  update() only records the sample; the
// first getter integrates everything recorded since reset() on the GPU (vilba_preintegrate_batch).
class IMUPreintegrator {
public:
    IMUPreintegrator() { reset(); }
    void reset() {  // IMUPreintegrator.cpp:39-56
        mGyro.clear(), mAcc.clear(), mDt.clear();
        std::memset(mState, 0, sizeof(mState));
        mState[VILBA_PI_DR] = mState[VILBA_PI_DR + 4] = mState[VILBA_PI_DR + 8] = 1.0;
        mDirty = false;
    }
    // omega = gyro - bias_g, acc = acc - bias_a (already bias-corrected, as in the reference)
    void update(const Vector3d& omega, const Vector3d& acc, const double& dt) {  // IMUPreintegrator.cpp:63-112
        for (int i = 0; i < 3; ++i) mGyro.push_back(omega[i]), mAcc.push_back(acc[i]);
        mDt.push_back(dt);
        mDirty = true;
    }
    Vector3d getDeltaP() const { return vec(VILBA_PI_DP); }
    Vector3d getDeltaV() const { return vec(VILBA_PI_DV); }
    Matrix3d getDeltaR() const { return mat(VILBA_PI_DR); }
    Matrix3d getJPBiasg() const { return mat(VILBA_PI_JPG); }
    Matrix3d getJPBiasa() const { return mat(VILBA_PI_JPA); }
    Matrix3d getJVBiasg() const { return mat(VILBA_PI_JVG); }
    Matrix3d getJVBiasa() const { return mat(VILBA_PI_JVA); }
    Matrix3d getJRBiasg() const { return mat(VILBA_PI_JRG); }
    Matrix9d getCovPVPhi() const {
        flush();
        Matrix9d c;
        std::memcpy(c.m, mState + VILBA_PI_COV, sizeof(c.m));
        return c;
    }
    double getDeltaTime() const {
        flush();
        return mState[VILBA_PI_DT];
    }
    // flat record (include/vilba.h layout) -- what the optimiser shim copies into the window
    const double* raw() const {
        flush();
        return mState;
    }
    void setRaw(const double* s) {  // batched re-integration writes results back through this
        std::memcpy(mState, s, sizeof(mState));
        mDirty = false;
    }
    // pending samples, for batched integration of many key-frames in one launch
    const std::vector<double>& pendingGyro() const { return mGyro; }
    const std::vector<double>& pendingAcc() const { return mAcc; }
    const std::vector<double>& pendingDt() const { return mDt; }

private:
    void flush() const {
        if (!mDirty) return;
        const int32_t sb[2] = {0, (int32_t)mDt.size()};
        const double zero[3] = {0, 0, 0};
        int st = vilba_preintegrate_batch(shim::context(), 1, sb, mGyro.data(), mAcc.data(), mDt.data(), zero, zero, mState);
        if (st != VILBA_OK) throw std::runtime_error(std::string("vilba_preintegrate_batch: ") + vilba_last_error(shim::context()));
        mDirty = false;
    }
    Vector3d vec(int off) const {
        flush();
        return Vector3d(mState[off], mState[off + 1], mState[off + 2]);
    }
    Matrix3d mat(int off) const {
        flush();
        Matrix3d m;
        std::memcpy(m.m, mState + off, sizeof(m.m));
        return m;
    }
    std::vector<double> mGyro, mAcc, mDt;
    mutable double mState[VILBA_PREINT_DOUBLES];
    mutable bool mDirty;
};

// ---------------------------------------------------------------------------------------------------
class NavState {  // src/IMU/NavState.h:14-140
public:
    Vector3d Get_P() const { return _P; }
    Vector3d Get_V() const { return _V; }
    shim::Quat Get_R() const { return _R; }
    Matrix3d Get_RotMatrix() const { return shim::quat_to_matrix(_R); }
    void Set_Pos(const Vector3d& p) { _P = p; }
    void Set_Vel(const Vector3d& v) { _V = v; }
    void Set_Rot(const Matrix3d& R) { _R = shim::matrix_to_quat(R); }
    void Set_Rot(const shim::Quat& q) { _R = q; }
    Vector3d Get_BiasGyr() const { return _BiasGyr; }
    Vector3d Get_BiasAcc() const { return _BiasAcc; }
    void Set_BiasGyr(const Vector3d& b) { _BiasGyr = b; }
    void Set_BiasAcc(const Vector3d& b) { _BiasAcc = b; }
    Vector3d Get_dBias_Gyr() const { return _dBias_g; }
    Vector3d Get_dBias_Acc() const { return _dBias_a; }
    void Set_DeltaBiasGyr(const Vector3d& d) { _dBias_g = d; }
    void Set_DeltaBiasAcc(const Vector3d& d) { _dBias_a = d; }
    void toFlat(double* s) const {  // VILBA_NS_DOUBLES layout
        for (int i = 0; i < 3; ++i) {
            s[i] = _P[i], s[3 + i] = _V[i], s[10 + i] = _BiasGyr[i], s[13 + i] = _BiasAcc[i];
            s[16 + i] = _dBias_g[i], s[19 + i] = _dBias_a[i];
        }
        s[6] = _R.w, s[7] = _R.x, s[8] = _R.y, s[9] = _R.z;
    }

private:
    Vector3d _P, _V;
    shim::Quat _R;
    Vector3d _BiasGyr, _BiasAcc, _dBias_g, _dBias_a;
};

// ---------------------------------------------------------------------------------------------------
struct ConfigParam {  // src/IMU/configparam.h:21-23 (static globals in the reference)
    static double& EigTbc(int r, int c) {
        static double T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        return T[4 * r + c];
    }
    static const double* GetEigTbc() { return &EigTbc(0, 0); }
    static shim::MatF GetMatTbc() {
        shim::MatF T(4, 4);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) T.at(r, c) = (float)EigTbc(r, c);
        return T;
    }
};

struct Converter {  // src/Converter.cpp:110-160: float32 quantisation at the boundary
    static Vector3d toVector3d(const shim::MatF& v) { return Vector3d(v.at(0), v.at(1), v.at(2)); }
    static shim::MatF toCvMat(const Vector3d& v) {
        shim::MatF m(3, 1);
        for (int i = 0; i < 3; ++i) m.at(i) = (float)v[i];
        return m;
    }
    static shim::MatF toCvMat(const Matrix3d& R) {
        shim::MatF m(3, 3);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) m.at(i, j) = (float)R(i, j);
        return m;
    }
};

class KeyFrame;
struct cmpKeyFrameId {  // src/MapPoint.cpp:19-22
    bool operator()(const KeyFrame* a, const KeyFrame* b) const;
};
typedef std::map<KeyFrame*, size_t, cmpKeyFrameId> mapMapPointObs;  // include/MapPoint.h:28

// ---------------------------------------------------------------------------------------------------
class MapPoint {
public:
    explicit MapPoint(long unsigned int id) : mnId(id), mnBALocalForKF(0), mWorldPos(3, 1) {}
    void SetWorldPos(const shim::MatF& p) { mWorldPos = p; }  // MapPoint.cpp:90-96
    shim::MatF GetWorldPos() const { return mWorldPos; }
    mapMapPointObs GetObservations() const { return mObservations; }
    void AddObservation(KeyFrame* pKF, size_t idx) { mObservations[pKF] = idx; }
    void EraseObservation(KeyFrame* pKF) { mObservations.erase(pKF); }
    bool isBad() const { return mbBad; }
    void UpdateNormalAndDepth() { ++mnNormalUpdates; }  // geometry bookkeeping is outside the path
    long unsigned int mnId;
    long unsigned int mnBALocalForKF;
    shim::MatF mPosGBA;  // include/MapPoint.h:127-128
    long unsigned int mnBAGlobalForKF = 0;
    bool mbBad = false;
    int mnNormalUpdates = 0;

private:
    shim::MatF mWorldPos;
    mapMapPointObs mObservations;
};

struct KeyPoint {  // cv::KeyPoint members used: pt.{x,y}, octave
    struct {
        float x, y;
    } pt;
    int octave;
};

class KeyFrame {
public:
    KeyFrame(long unsigned int id, double t, float fx_, float fy_, float cx_, float cy_)
        : mnId(id), mTimeStamp(t), mnBALocalForKF(0), mnBAFixedForKF(0), fx(fx_), fy(fy_), cx(cx_), cy(cy_), Tcw(4, 4) {}
    // NavState accessors (src/KeyFrame.cpp:124-184)
    const NavState& GetNavState() const { return mNavState; }
    void SetNavState(const NavState& ns) { mNavState = ns; }
    void SetNavStatePos(const Vector3d& p) { mNavState.Set_Pos(p); }
    void SetNavStateVel(const Vector3d& v) { mNavState.Set_Vel(v); }
    void SetNavStateRot(const shim::Quat& q) { mNavState.Set_Rot(q); }
    void SetNavStateRot(const Matrix3d& R) { mNavState.Set_Rot(R); }
    void SetNavStateDeltaBg(const Vector3d& d) { mNavState.Set_DeltaBiasGyr(d); }
    void SetNavStateDeltaBa(const Vector3d& d) { mNavState.Set_DeltaBiasAcc(d); }
    const IMUPreintegrator& GetIMUPreInt() const { return mIMUPreInt; }
    IMUPreintegrator& IMUPreIntRef() { return mIMUPreInt; }
    KeyFrame* GetPrevKeyFrame() const { return mpPrevKeyFrame; }
    void SetPrevKeyFrame(KeyFrame* p) { mpPrevKeyFrame = p; }
    std::vector<MapPoint*> GetMapPointMatches() const { return mvpMapPoints; }
    void EraseMapPointMatch(MapPoint* pMP) {  // include/KeyFrame.h:177
        for (auto& p : mvpMapPoints)
            if (p == pMP) p = nullptr;
    }
    bool isBad() const { return mbBad; }
    // KeyFrame::ComputePreInt (src/KeyFrame.cpp:195-252): same update() sequence, integrated on the GPU
    void ComputePreInt() {
        if (!mpPrevKeyFrame) {
            if (mnId != 0) std::cerr << "previous KeyFrame is NULL, pre-integrator not changed. id: " << mnId << std::endl;
            return;
        }
        mIMUPreInt.reset();
        const Vector3d bg = mpPrevKeyFrame->GetNavState().Get_BiasGyr(), ba = mpPrevKeyFrame->GetNavState().Get_BiasAcc();
        if (mvIMUData.empty()) return;  // the reference dereferences front() here (UB); we leave the reset state
        {
            const IMUData& imu = mvIMUData.front();
            mIMUPreInt.update(imu._g - bg, imu._a - ba, imu._t - mpPrevKeyFrame->mTimeStamp);
        }
        for (size_t i = 0; i < mvIMUData.size(); ++i) {
            const IMUData& imu = mvIMUData[i];
            const double nextt = (i == mvIMUData.size() - 1) ? mTimeStamp : mvIMUData[i + 1]._t;
            mIMUPreInt.update(imu._g - bg, imu._a - ba, nextt - imu._t);
        }
    }
    // KeyFrame::UpdatePoseFromNS (src/KeyFrame.cpp:96-114): camera pose in float32.  The global BA's mTcwGBA
    // (Optimizer.cpp:1646-1652: toCvMatInverse(Twb * Tbc)) is the same float arithmetic on another NavState.
    static shim::MatF PoseFromNS(const NavState& ns, const shim::MatF& Tbc) {
        const shim::MatF Rwb = Converter::toCvMat(ns.Get_RotMatrix()), Pwb = Converter::toCvMat(ns.Get_P());
        float Rwc[9], Pwc[3];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) {
                float s = 0.f;
                for (int k = 0; k < 3; ++k) s += Rwb.at(r, k) * Tbc.at(k, c);
                Rwc[3 * r + c] = s;
            }
            float s = 0.f;
            for (int k = 0; k < 3; ++k) s += Rwb.at(r, k) * Tbc.at(k, 3);
            Pwc[r] = s + Pwb.at(r);
        }
        shim::MatF T(4, 4);
        for (int r = 0; r < 3; ++r) {
            float s = 0.f;
            for (int c = 0; c < 3; ++c) {
                T.at(r, c) = Rwc[3 * c + r];  // Rcw = Rwc^T
                s += Rwc[3 * c + r] * Pwc[c];
            }
            T.at(r, 3) = -s;
        }
        T.at(3, 3) = 1.f;
        return T;
    }
    void UpdatePoseFromNS(const shim::MatF& Tbc) { Tcw = PoseFromNS(mNavState, Tbc); }
    shim::MatF GetPose() const { return Tcw; }

    long unsigned int mnId;
    double mTimeStamp;
    long unsigned int mnBALocalForKF, mnBAFixedForKF;
    // results of a global BA that runs beside the mapping thread (include/KeyFrame.h:263-271)
    NavState mNavStateGBA;
    shim::MatF mTcwGBA;
    long unsigned int mnBAGlobalForKF = 0;
    const float fx, fy, cx, cy;
    std::vector<KeyPoint> mvKeysUn;
    std::vector<float> mvuRight;  // negative => monocular
    std::vector<float> mvInvLevelSigma2;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<IMUData> mvIMUData;
    bool mbBad = false;

private:
    NavState mNavState;
    IMUPreintegrator mIMUPreInt;
    KeyFrame* mpPrevKeyFrame = nullptr;
    shim::MatF Tcw;
};
inline bool cmpKeyFrameId::operator()(const KeyFrame* a, const KeyFrame* b) const { return a->mnId < b->mnId; }

struct Map {
    std::mutex mMutexMapUpdate;  // include/Map.h:72
    // Map::GetAllKeyFrames / GetAllMapPoints (src/Map.cpp:94-105); std::set<T*> order, as in the reference
    void AddKeyFrame(KeyFrame* pKF) { mspKeyFrames.insert(pKF); }
    void AddMapPoint(MapPoint* pMP) { mspMapPoints.insert(pMP); }
    std::vector<KeyFrame*> GetAllKeyFrames() const { return std::vector<KeyFrame*>(mspKeyFrames.begin(), mspKeyFrames.end()); }
    std::vector<MapPoint*> GetAllMapPoints() const { return std::vector<MapPoint*>(mspMapPoints.begin(), mspMapPoints.end()); }
    std::set<KeyFrame*> mspKeyFrames;
    std::set<MapPoint*> mspMapPoints;
};
struct LocalMapping {
    void SetMapUpdateFlagInTracking(bool b) { mbMapUpdateFlagForTracking = b; }  // include/LocalMapping.h:52
    bool mbMapUpdateFlagForTracking = false;
};

// Batched re-integration of many key-frames in ONE kernel launch -- the in-product consumer is
// LocalMapping::TryInitVIO, which re-integrates every key-frame after the bias estimate
// (src/LocalMapping.cpp:650-657,700-706).
inline void ComputePreIntBatch(const std::vector<KeyFrame*>& kfs) {
    std::vector<int32_t> sb(1, 0);
    std::vector<double> g, a, t, zero;
    std::vector<KeyFrame*> todo;
    for (KeyFrame* kf : kfs) {
        if (!kf->GetPrevKeyFrame()) continue;
        kf->ComputePreInt();  // records the samples (lazy); nothing is integrated yet
        const IMUPreintegrator& p = kf->GetIMUPreInt();
        g.insert(g.end(), p.pendingGyro().begin(), p.pendingGyro().end());
        a.insert(a.end(), p.pendingAcc().begin(), p.pendingAcc().end());
        t.insert(t.end(), p.pendingDt().begin(), p.pendingDt().end());
        sb.push_back((int32_t)t.size());
        todo.push_back(kf);
    }
    if (todo.empty()) return;
    zero.assign(3 * todo.size(), 0.0);
    std::vector<double> out((size_t)VILBA_PREINT_DOUBLES * todo.size());
    int st = vilba_preintegrate_batch(shim::context(), (int32_t)todo.size(), sb.data(), g.data(), a.data(), t.data(),
                                      zero.data(), zero.data(), out.data());
    if (st != VILBA_OK) throw std::runtime_error(std::string("vilba_preintegrate_batch: ") + vilba_last_error(shim::context()));
    for (size_t i = 0; i < todo.size(); ++i) todo[i]->IMUPreIntRef().setRaw(&out[(size_t)VILBA_PREINT_DOUBLES * i]);
}

// ---------------------------------------------------------------------------------------------------
class Optimizer {
public:
    // Same signature and behaviour as the reference (include/Optimizer.h:44-46): phases A/B gather and
    // flatten on the host, C..E run on the GPU behind vilba_local_ba, F writes back.
    // Optimizer::GlobalBundleAdjustmentNavState (src/Optimizer.cpp:1392-1668): same signature.  The graph build
    // (:1396-1619) becomes "flatten the whole map into one vilba_window", optimize(nIterations) (:1621-1624) runs
    // on the GPU through vilba_global_ba, and the write-back (:1626-1667) goes to the live state when nLoopKF == 0
    // or to the ...GBA members otherwise.
    static void GlobalBundleAdjustmentNavState(Map* pMap, const shim::MatF& gw, int nIterations, bool* pbStopFlag,
                                               const unsigned long nLoopKF, const bool bRobust,
                                               vilba_result* pTrace = NULL) {
        std::vector<KeyFrame*> vpKFs = pMap->GetAllKeyFrames();
        const std::vector<MapPoint*> vpMP = pMap->GetAllMapPoints();
        const double* Tbc = ConfigParam::GetEigTbc();
        const Vector3d GravityVec = Converter::toVector3d(gw);

        // vertices (:1422-1440): g2o orders the free ones by vertex id = 2 * mnId (+1), i.e. by key-frame id
        std::vector<KeyFrame*> kfs;
        for (KeyFrame* pKF : vpKFs)
            if (!pKF->isBad()) kfs.push_back(pKF);
        std::sort(kfs.begin(), kfs.end(), cmpKeyFrameId());
        const int K = (int)kfs.size();
        if (K == 0) return;
        std::map<KeyFrame*, int> kfIndex;
        std::vector<double> kf_state((size_t)VILBA_NS_DOUBLES * K);
        std::vector<uint8_t> kf_flags(K, VILBA_KF_HAS_BIAS);
        std::vector<int64_t> kf_id(K);
        for (int i = 0; i < K; ++i) {
            kfIndex[kfs[i]] = i;
            kfs[i]->GetNavState().toFlat(&kf_state[(size_t)VILBA_NS_DOUBLES * i]);
            kf_id[i] = (int64_t)kfs[i]->mnId;
            if (kfs[i]->mnId == 0) kf_flags[i] |= VILBA_KF_FIXED;  // vNSPVR / vNSBias ->setFixed(pKF->mnId == 0)
        }
        // IMU edges (:1447-1503), in vpKFs order like the reference's loop (the order does not enter the result
        // beyond floating-point summation order)
        std::vector<int32_t> imu_i, imu_j;
        std::vector<double> imu_preint;
        for (KeyFrame* pKF1 : kfs) {
            KeyFrame* pKF0 = pKF1->GetPrevKeyFrame();
            if (!pKF0) {
                if (pKF1->mnId != 0) std::cerr << "Previous KeyFrame is NULL?" << std::endl;
                continue;
            }
            if (!kfIndex.count(pKF0)) {  // the reference would hand g2o a null vertex here
                std::cerr << "previous KeyFrame of " << pKF1->mnId << " is bad, global BA skipped" << std::endl;
                return;
            }
            imu_i.push_back(kfIndex.at(pKF0));
            imu_j.push_back(kfIndex.at(pKF1));
            const double* raw = pKF1->GetIMUPreInt().raw();
            imu_preint.insert(imu_preint.end(), raw, raw + VILBA_PREINT_DOUBLES);
        }
        // map points and mono edges (:1509-1617); a point without an edge is left out (vbNotIncludedMP)
        std::vector<MapPoint*> pts;
        std::vector<double> pt_xyz;
        std::vector<int32_t> pt_obs_begin(1, 0), obs_kf;
        std::vector<float> obs_uv, obs_is2;
        for (MapPoint* pMP : vpMP) {
            if (pMP->isBad()) continue;
            const size_t n0 = obs_kf.size();
            for (auto& ob : pMP->GetObservations()) {
                KeyFrame* pKF = ob.first;
                if (pKF->isBad()) continue;
                if (!(pKF->mvuRight[ob.second] < 0)) {
                    std::cerr << "Stereo not supported" << std::endl;
                    continue;
                }
                const KeyPoint& kpUn = pKF->mvKeysUn[ob.second];
                obs_kf.push_back(kfIndex.at(pKF));
                obs_uv.push_back(kpUn.pt.x), obs_uv.push_back(kpUn.pt.y);
                obs_is2.push_back(pKF->mvInvLevelSigma2[kpUn.octave]);
            }
            if (obs_kf.size() == n0) continue;
            const Vector3d Pw = Converter::toVector3d(pMP->GetWorldPos());
            for (int d = 0; d < 3; ++d) pt_xyz.push_back(Pw[d]);
            pts.push_back(pMP);
            pt_obs_begin.push_back((int32_t)obs_kf.size());
        }
        vilba_window win;
        std::memset(&win, 0, sizeof(win));
        win.n_kf = K, win.n_imu = (int32_t)imu_i.size(), win.n_pts = (int32_t)pts.size(), win.n_obs = (int32_t)obs_kf.size();
        win.kf_state = kf_state.data(), win.kf_flags = kf_flags.data(), win.kf_id = kf_id.data();
        win.imu_kf_i = imu_i.data(), win.imu_kf_j = imu_j.data(), win.imu_preint = imu_preint.data();
        win.pt_xyz = pt_xyz.data(), win.pt_obs_begin = pt_obs_begin.data();
        win.obs_kf = obs_kf.data(), win.obs_uv = obs_uv.data(), win.obs_inv_sigma2 = obs_is2.data();
        win.fx = kfs.front()->fx, win.fy = kfs.front()->fy, win.cx = kfs.front()->cx, win.cy = kfs.front()->cy;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) win.Rbc[3 * r + c] = Tbc[4 * r + c];
            win.Pbc[r] = Tbc[4 * r + 3];
            win.gravity[r] = GravityVec[r];
        }
        std::vector<double> out_state(kf_state.size()), out_pts(pt_xyz.size());
        vilba_result local_res;
        vilba_result& res = pTrace ? *pTrace : local_res;
        std::memset(&res, 0, sizeof(res));
        res.kf_state = out_state.data(), res.pt_xyz = out_pts.data();

        // ---- optimizer.initializeOptimization(); optimizer.optimize(nIterations) on the GPU ----
        static_assert(sizeof(bool) == 1, "bool* pbStopFlag is polled as a byte");
        const int st = vilba_global_ba(shim::context(), &win, nIterations, bRobust ? 1 : 0, &res,
                                       reinterpret_cast<const volatile uint8_t*>(pbStopFlag));
        if (pTrace) res.kf_state = NULL, res.pt_xyz = NULL;  // scratch does not outlive this call
        if (st != VILBA_OK) {
            std::cerr << "vilba_global_ba failed: " << vilba_last_error(shim::context()) << std::endl;
            return;
        }

        // ---- recover optimised data (:1626-1667) ----
        const shim::MatF matTbc = ConfigParam::GetMatTbc();
        for (int i = 0; i < K; ++i) {
            KeyFrame* pKF = kfs[i];
            const double* s = &out_state[(size_t)VILBA_NS_DOUBLES * i];
            NavState ns_recov = pKF->GetNavState();  // the base biases are not optimised
            ns_recov.Set_Pos(Vector3d(s[0], s[1], s[2]));
            ns_recov.Set_Vel(Vector3d(s[3], s[4], s[5]));
            shim::Quat q;
            q.w = s[6], q.x = s[7], q.y = s[8], q.z = s[9];
            ns_recov.Set_Rot(q);
            ns_recov.Set_DeltaBiasGyr(Vector3d(s[16], s[17], s[18]));
            ns_recov.Set_DeltaBiasAcc(Vector3d(s[19], s[20], s[21]));
            if (nLoopKF == 0) {
                pKF->SetNavState(ns_recov);
                pKF->UpdatePoseFromNS(matTbc);
            } else {
                pKF->mNavStateGBA = ns_recov;
                pKF->mTcwGBA = KeyFrame::PoseFromNS(ns_recov, matTbc);
                pKF->mnBAGlobalForKF = nLoopKF;
            }
        }
        for (size_t p = 0; p < pts.size(); ++p) {
            const Vector3d Pw(out_pts[3 * p], out_pts[3 * p + 1], out_pts[3 * p + 2]);
            if (nLoopKF == 0) {
                pts[p]->SetWorldPos(Converter::toCvMat(Pw));
                pts[p]->UpdateNormalAndDepth();
            } else {
                pts[p]->mPosGBA = Converter::toCvMat(Pw);
                pts[p]->mnBAGlobalForKF = nLoopKF;
            }
        }
    }

    static void LocalBundleAdjustmentNavState(KeyFrame* pCurKF, const std::list<KeyFrame*>& lLocalKeyFrames,
                                              bool* pbStopFlag, Map* pMap, shim::MatF& gw, LocalMapping* pLM = NULL,
                                              vilba_result* pTrace = NULL) {
        if (pCurKF != lLocalKeyFrames.back()) std::cerr << "pCurKF != lLocalKeyFrames.back. check" << std::endl;
        const double* Tbc = ConfigParam::GetEigTbc();
        const Vector3d GravityVec = Converter::toVector3d(gw);  // float -> double (Optimizer.cpp:2332)

        // ---- phase A: gather (Optimizer.cpp:2335-2402) ----
        for (KeyFrame* pKFi : lLocalKeyFrames) pKFi->mnBALocalForKF = pCurKF->mnId;
        std::list<MapPoint*> lLocalMapPoints;
        for (KeyFrame* pKFi : lLocalKeyFrames)
            for (MapPoint* pMP : pKFi->GetMapPointMatches())
                if (pMP && !pMP->isBad() && pMP->mnBALocalForKF != pCurKF->mnId) {
                    lLocalMapPoints.push_back(pMP);
                    pMP->mnBALocalForKF = pCurKF->mnId;
                }
        std::list<KeyFrame*> lFixedCameras;
        KeyFrame* pKFPrevLocal = lLocalKeyFrames.front()->GetPrevKeyFrame();
        if (pKFPrevLocal) {
            pKFPrevLocal->mnBAFixedForKF = pCurKF->mnId;
            if (!pKFPrevLocal->isBad())
                lFixedCameras.push_back(pKFPrevLocal);
            else
                std::cerr << "pKFPrevLocal is Bad?" << std::endl;
        } else
            std::cerr << "pKFPrevLocal is NULL?" << std::endl;
        for (MapPoint* pMP : lLocalMapPoints)
            for (auto& ob : pMP->GetObservations()) {
                KeyFrame* pKFi = ob.first;
                if (pKFi->mnBALocalForKF != pCurKF->mnId && pKFi->mnBAFixedForKF != pCurKF->mnId) {
                    pKFi->mnBAFixedForKF = pCurKF->mnId;
                    if (!pKFi->isBad()) lFixedCameras.push_back(pKFi);
                }
            }

        // ---- phase B: flatten instead of building a g2o graph (Optimizer.cpp:2405-2639) ----
        // key-frames in increasing mnId = g2o's vertex-id order (sparse_optimizer.cpp:166-190)
        std::vector<KeyFrame*> kfs(lLocalKeyFrames.begin(), lLocalKeyFrames.end());
        kfs.insert(kfs.end(), lFixedCameras.begin(), lFixedCameras.end());
        std::sort(kfs.begin(), kfs.end(), [](KeyFrame* a, KeyFrame* b) { return a->mnId < b->mnId; });
        std::map<KeyFrame*, int> kfIndex;
        for (size_t i = 0; i < kfs.size(); ++i) kfIndex[kfs[i]] = (int)i;
        const int K = (int)kfs.size();
        std::vector<double> kf_state((size_t)VILBA_NS_DOUBLES * K);
        std::vector<uint8_t> kf_flags(K, 0);
        std::vector<int64_t> kf_id(K);
        for (int i = 0; i < K; ++i) {
            kfs[i]->GetNavState().toFlat(&kf_state[(size_t)VILBA_NS_DOUBLES * i]);
            kf_id[i] = (int64_t)kfs[i]->mnId;
            const bool local = kfs[i]->mnBALocalForKF == pCurKF->mnId;
            if (!local) kf_flags[i] |= VILBA_KF_FIXED;
            if (local || kfs[i] == pKFPrevLocal) kf_flags[i] |= VILBA_KF_HAS_BIAS;
        }
        std::vector<int32_t> imu_i, imu_j;
        std::vector<double> imu_preint;
        for (KeyFrame* pKF1 : lLocalKeyFrames) {  // EdgeNavStatePVR + EdgeNavStateBias (Optimizer.cpp:2494-2541)
            KeyFrame* pKF0 = pKF1->GetPrevKeyFrame();
            imu_i.push_back(kfIndex.at(pKF0));
            imu_j.push_back(kfIndex.at(pKF1));
            const double* raw = pKF1->GetIMUPreInt().raw();
            imu_preint.insert(imu_preint.end(), raw, raw + VILBA_PREINT_DOUBLES);
            if (pKF1->GetIMUPreInt().getDeltaTime() < 1e-3)
                std::cerr << "IMU pre-integrator delta time between 2 KFs too small: " << pKF1->GetIMUPreInt().getDeltaTime() << std::endl;
        }
        std::vector<MapPoint*> pts(lLocalMapPoints.begin(), lLocalMapPoints.end());
        std::vector<double> pt_xyz(3 * pts.size());
        std::vector<int32_t> pt_obs_begin(1, 0), obs_kf;
        std::vector<float> obs_uv, obs_is2;
        std::vector<KeyFrame*> vpEdgeKFMono;
        std::vector<MapPoint*> vpMapPointEdgeMono;
        for (size_t p = 0; p < pts.size(); ++p) {
            const Vector3d Pw = Converter::toVector3d(pts[p]->GetWorldPos());
            for (int d = 0; d < 3; ++d) pt_xyz[3 * p + d] = Pw[d];
            for (auto& ob : pts[p]->GetObservations()) {  // ordered by key-frame id
                KeyFrame* pKFi = ob.first;
                if (pKFi->isBad()) continue;
                if (!(pKFi->mvuRight[ob.second] < 0)) {
                    std::cerr << "Stereo not supported yet, why here?? check." << std::endl;
                    continue;
                }
                const KeyPoint& kpUn = pKFi->mvKeysUn[ob.second];
                obs_kf.push_back(kfIndex.at(pKFi));
                obs_uv.push_back(kpUn.pt.x), obs_uv.push_back(kpUn.pt.y);
                obs_is2.push_back(pKFi->mvInvLevelSigma2[kpUn.octave]);
                vpEdgeKFMono.push_back(pKFi);
                vpMapPointEdgeMono.push_back(pts[p]);
            }
            pt_obs_begin.push_back((int32_t)obs_kf.size());
        }
        vilba_window win;
        std::memset(&win, 0, sizeof(win));
        win.n_kf = K, win.n_imu = (int32_t)imu_i.size(), win.n_pts = (int32_t)pts.size(), win.n_obs = (int32_t)obs_kf.size();
        win.kf_state = kf_state.data(), win.kf_flags = kf_flags.data(), win.kf_id = kf_id.data();
        win.imu_kf_i = imu_i.data(), win.imu_kf_j = imu_j.data(), win.imu_preint = imu_preint.data();
        win.pt_xyz = pt_xyz.data(), win.pt_obs_begin = pt_obs_begin.data();
        win.obs_kf = obs_kf.data(), win.obs_uv = obs_uv.data(), win.obs_inv_sigma2 = obs_is2.data();
        KeyFrame* any = kfs.front();
        win.fx = any->fx, win.fy = any->fy, win.cx = any->cx, win.cy = any->cy;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) win.Rbc[3 * r + c] = Tbc[4 * r + c];
            win.Pbc[r] = Tbc[4 * r + 3];
            win.gravity[r] = GravityVec[r];
        }
        std::vector<double> out_state(kf_state.size()), out_pts(pt_xyz.size());  // (no per-edge chi2: phase F does not read it)
        std::vector<uint8_t> out_outlier(obs_kf.size());
        vilba_result local_res;
        vilba_result& res = pTrace ? *pTrace : local_res;
        std::memset(&res, 0, sizeof(res));
        res.kf_state = out_state.data(), res.pt_xyz = out_pts.data();
        res.obs_outlier = out_outlier.data();

        // ---- phases C..E on the GPU ----
        static_assert(sizeof(bool) == 1, "bool* pbStopFlag is polled as a byte");
        const int st = vilba_local_ba(shim::context(), &win, &res, reinterpret_cast<const volatile uint8_t*>(pbStopFlag));
        if (st == VILBA_ABORTED) return;  // silent early return, nothing written (Optimizer.cpp:2643-2645)
        if (st != VILBA_OK) {
            std::cerr << "vilba_local_ba failed: " << vilba_last_error(shim::context()) << std::endl;
            return;
        }

        // ---- phase F: write-back (Optimizer.cpp:2680-2769) ----
        std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);
        for (size_t e = 0; e < out_outlier.size(); ++e)
            if (out_outlier[e] && !vpMapPointEdgeMono[e]->isBad()) {
                vpEdgeKFMono[e]->EraseMapPointMatch(vpMapPointEdgeMono[e]);
                vpMapPointEdgeMono[e]->EraseObservation(vpEdgeKFMono[e]);
            }
        const shim::MatF matTbc = ConfigParam::GetMatTbc();
        for (KeyFrame* pKFi : lLocalKeyFrames) {
            const double* s = &out_state[(size_t)VILBA_NS_DOUBLES * kfIndex.at(pKFi)];
            pKFi->SetNavStatePos(Vector3d(s[0], s[1], s[2]));
            pKFi->SetNavStateVel(Vector3d(s[3], s[4], s[5]));
            shim::Quat q;
            q.w = s[6], q.x = s[7], q.y = s[8], q.z = s[9];
            pKFi->SetNavStateRot(q);
            pKFi->SetNavStateDeltaBg(Vector3d(s[16], s[17], s[18]));
            pKFi->SetNavStateDeltaBa(Vector3d(s[19], s[20], s[21]));
            pKFi->UpdatePoseFromNS(matTbc);
        }
        for (size_t p = 0; p < pts.size(); ++p) {
            pts[p]->SetWorldPos(Converter::toCvMat(Vector3d(out_pts[3 * p], out_pts[3 * p + 1], out_pts[3 * p + 2])));
            pts[p]->UpdateNormalAndDepth();
        }
        if (pLM) pLM->SetMapUpdateFlagInTracking(true);
        if (pTrace) res.kf_state = NULL, res.pt_xyz = NULL, res.obs_outlier = NULL, res.obs_chi2 = NULL;  // scratch is gone
    }
};

}  // namespace ORB_SLAM2
#endif

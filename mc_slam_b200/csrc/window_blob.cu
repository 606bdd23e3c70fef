// window_blob.cu -- wire / on-disk format of one local-BA window (host code only).
//
// The reference keeps a window only as a web of KeyFrame / MapPoint objects (Optimizer.cpp:2329-2402 gathers it on
// every call); there is no serialised form.  This blob is the flattened struct-of-arrays window of include/vilba.h
// made self-describing, so that windows captured from a patched reference build (phase A of
// Optimizer::LocalBundleAdjustmentNavState) can be stored, shipped to another process / GPU and replayed through
// vilba_local_ba, and so that solver inputs of a field run can be attached to a bug report.
//
//   offset 0   : char[8]  "VILBAWIN"
//          8   : u32 version (1) | u32 endian tag 0x01020304 (little-endian hosts write 04 03 02 01)
//         16   : i32 n_kf, n_imu, n_pts, n_obs
//         32   : f64 fx, fy, cx, cy, Rbc[9], Pbc[3], gravity[3]                      (19 doubles)
//        184   : u64 payload bytes | u64 FNV-1a 64 of the payload | zero padding up to 256
//        256   : payload = the arrays of vilba_window in declaration order, each padded to 8 bytes:
//                kf_state f64[22 K] | kf_flags u8[K] | kf_id i64[K] | imu_kf_i i32[NI] | imu_kf_j i32[NI] |
//                imu_preint f64[142 NI] | pt_xyz f64[3 P] | pt_obs_begin i32[P + 1] | obs_kf i32[E] |
//                obs_uv f32[2 E] | obs_inv_sigma2 f32[E]
// Deserialisation is zero-copy: the returned vilba_window points into the caller's buffer (which must be 8-byte
// aligned and outlive the view).
#include <cstdint>
#include <cstring>

#include "../../include/vilba.h"

namespace {

constexpr char kMagic[8] = {'V', 'I', 'L', 'B', 'A', 'W', 'I', 'N'};
constexpr uint32_t kVersion = 1;
constexpr uint32_t kEndianTag = 0x01020304u;
constexpr size_t kHeaderBytes = 256;

inline size_t pad8(size_t n) { return (n + 7) & ~(size_t)7; }

struct Sizes {
    size_t kf_state, kf_flags, kf_id, imu_i, imu_j, imu_preint, pt_xyz, pt_begin, obs_kf, obs_uv, obs_is2, total;
};

Sizes payload_sizes(int64_t K, int64_t NI, int64_t P, int64_t E) {
    Sizes s;
    s.kf_state = pad8(sizeof(double) * VILBA_NS_DOUBLES * (size_t)K);
    s.kf_flags = pad8((size_t)K);
    s.kf_id = pad8(sizeof(int64_t) * (size_t)K);
    s.imu_i = pad8(sizeof(int32_t) * (size_t)NI);
    s.imu_j = pad8(sizeof(int32_t) * (size_t)NI);
    s.imu_preint = pad8(sizeof(double) * VILBA_PREINT_DOUBLES * (size_t)NI);
    s.pt_xyz = pad8(sizeof(double) * 3 * (size_t)P);
    s.pt_begin = pad8(sizeof(int32_t) * ((size_t)P + 1));
    s.obs_kf = pad8(sizeof(int32_t) * (size_t)E);
    s.obs_uv = pad8(sizeof(float) * 2 * (size_t)E);
    s.obs_is2 = pad8(sizeof(float) * (size_t)E);
    s.total = s.kf_state + s.kf_flags + s.kf_id + s.imu_i + s.imu_j + s.imu_preint + s.pt_xyz + s.pt_begin + s.obs_kf + s.obs_uv +
              s.obs_is2;
    return s;
}

uint64_t fnv1a64(const unsigned char* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

bool counts_ok(const vilba_window* w) { return w && w->n_kf > 0 && w->n_imu >= 0 && w->n_pts >= 0 && w->n_obs >= 0; }

}  // namespace

extern "C" {

size_t vilba_window_blob_size(const vilba_window* w) {
    if (!counts_ok(w)) return 0;
    return kHeaderBytes + payload_sizes(w->n_kf, w->n_imu, w->n_pts, w->n_obs).total;
}

int vilba_window_serialize(const vilba_window* w, void* buf, size_t capacity, size_t* written) {
    if (written) *written = 0;
    if (!counts_ok(w) || !buf) return VILBA_ERR_ARG;
    if (!w->kf_state || !w->kf_flags) return VILBA_ERR_ARG;
    if (w->n_imu && (!w->imu_kf_i || !w->imu_kf_j || !w->imu_preint)) return VILBA_ERR_ARG;
    if (w->n_pts && (!w->pt_xyz || !w->pt_obs_begin)) return VILBA_ERR_ARG;
    if (w->n_obs && (!w->obs_kf || !w->obs_uv || !w->obs_inv_sigma2)) return VILBA_ERR_ARG;
    const int64_t K = w->n_kf, NI = w->n_imu, P = w->n_pts, E = w->n_obs;
    const Sizes s = payload_sizes(K, NI, P, E);
    const size_t total = kHeaderBytes + s.total;
    if (capacity < total) return VILBA_ERR_ARG;
    unsigned char* b = static_cast<unsigned char*>(buf);
    std::memset(b, 0, total);  // padding bytes are zeros: the blob of a window is unique
    std::memcpy(b, kMagic, 8);
    std::memcpy(b + 8, &kVersion, 4);
    std::memcpy(b + 12, &kEndianTag, 4);
    const int32_t counts[4] = {w->n_kf, w->n_imu, w->n_pts, w->n_obs};
    std::memcpy(b + 16, counts, 16);
    double cal[19];
    cal[0] = w->fx, cal[1] = w->fy, cal[2] = w->cx, cal[3] = w->cy;
    std::memcpy(cal + 4, w->Rbc, sizeof(double) * 9);
    std::memcpy(cal + 13, w->Pbc, sizeof(double) * 3);
    std::memcpy(cal + 16, w->gravity, sizeof(double) * 3);
    std::memcpy(b + 32, cal, sizeof(cal));
    unsigned char* p = b + kHeaderBytes;
    auto put = [&](const void* src, size_t bytes, size_t padded) {
        if (bytes) std::memcpy(p, src, bytes);
        p += padded;
    };
    put(w->kf_state, sizeof(double) * VILBA_NS_DOUBLES * (size_t)K, s.kf_state);
    put(w->kf_flags, (size_t)K, s.kf_flags);
    if (w->kf_id)
        put(w->kf_id, sizeof(int64_t) * (size_t)K, s.kf_id);
    else
        p += s.kf_id;  // informational ids absent: zeros
    put(w->imu_kf_i, sizeof(int32_t) * (size_t)NI, s.imu_i);
    put(w->imu_kf_j, sizeof(int32_t) * (size_t)NI, s.imu_j);
    put(w->imu_preint, sizeof(double) * VILBA_PREINT_DOUBLES * (size_t)NI, s.imu_preint);
    put(w->pt_xyz, sizeof(double) * 3 * (size_t)P, s.pt_xyz);
    if (P) {
        put(w->pt_obs_begin, sizeof(int32_t) * ((size_t)P + 1), s.pt_begin);
    } else {
        p += s.pt_begin;  // one zero
    }
    put(w->obs_kf, sizeof(int32_t) * (size_t)E, s.obs_kf);
    put(w->obs_uv, sizeof(float) * 2 * (size_t)E, s.obs_uv);
    put(w->obs_inv_sigma2, sizeof(float) * (size_t)E, s.obs_is2);
    const uint64_t plen = s.total, sum = fnv1a64(b + kHeaderBytes, s.total);
    std::memcpy(b + 184, &plen, 8);
    std::memcpy(b + 192, &sum, 8);
    if (written) *written = total;
    return VILBA_OK;
}

int vilba_window_deserialize(const void* buf, size_t len, vilba_window* out) {
    if (!buf || !out || len < kHeaderBytes) return VILBA_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(buf) % 8 != 0) return VILBA_ERR_ARG;  // the views are typed pointers into the buffer
    const unsigned char* b = static_cast<const unsigned char*>(buf);
    if (std::memcmp(b, kMagic, 8) != 0) return VILBA_ERR_ARG;
    uint32_t version, endian;
    std::memcpy(&version, b + 8, 4);
    std::memcpy(&endian, b + 12, 4);
    if (version != kVersion || endian != kEndianTag) return VILBA_ERR_ARG;
    int32_t counts[4];
    std::memcpy(counts, b + 16, 16);
    if (counts[0] <= 0 || counts[1] < 0 || counts[2] < 0 || counts[3] < 0) return VILBA_ERR_ARG;
    const Sizes s = payload_sizes(counts[0], counts[1], counts[2], counts[3]);
    uint64_t plen, sum;
    std::memcpy(&plen, b + 184, 8);
    std::memcpy(&sum, b + 192, 8);
    if (plen != s.total || len < kHeaderBytes + s.total) return VILBA_ERR_ARG;
    if (fnv1a64(b + kHeaderBytes, s.total) != sum) return VILBA_ERR_ARG;
    std::memset(out, 0, sizeof(*out));
    out->n_kf = counts[0], out->n_imu = counts[1], out->n_pts = counts[2], out->n_obs = counts[3];
    double cal[19];
    std::memcpy(cal, b + 32, sizeof(cal));
    out->fx = cal[0], out->fy = cal[1], out->cx = cal[2], out->cy = cal[3];
    std::memcpy(out->Rbc, cal + 4, sizeof(double) * 9);
    std::memcpy(out->Pbc, cal + 13, sizeof(double) * 3);
    std::memcpy(out->gravity, cal + 16, sizeof(double) * 3);
    const unsigned char* p = b + kHeaderBytes;
    out->kf_state = reinterpret_cast<const double*>(p), p += s.kf_state;
    out->kf_flags = p, p += s.kf_flags;
    out->kf_id = reinterpret_cast<const int64_t*>(p), p += s.kf_id;
    out->imu_kf_i = reinterpret_cast<const int32_t*>(p), p += s.imu_i;
    out->imu_kf_j = reinterpret_cast<const int32_t*>(p), p += s.imu_j;
    out->imu_preint = reinterpret_cast<const double*>(p), p += s.imu_preint;
    out->pt_xyz = reinterpret_cast<const double*>(p), p += s.pt_xyz;
    out->pt_obs_begin = reinterpret_cast<const int32_t*>(p), p += s.pt_begin;
    out->obs_kf = reinterpret_cast<const int32_t*>(p), p += s.obs_kf;
    out->obs_uv = reinterpret_cast<const float*>(p), p += s.obs_uv;
    out->obs_inv_sigma2 = reinterpret_cast<const float*>(p), p += s.obs_is2;
    // structural checks a solver call would also make (cheap, and a blob comes from outside)
    if (out->n_pts && (out->pt_obs_begin[0] != 0 || out->pt_obs_begin[out->n_pts] != out->n_obs)) return VILBA_ERR_ARG;
    for (int32_t q = 0; q < out->n_pts; ++q)
        if (out->pt_obs_begin[q + 1] < out->pt_obs_begin[q]) return VILBA_ERR_ARG;
    for (int32_t e = 0; e < out->n_obs; ++e)
        if (out->obs_kf[e] < 0 || out->obs_kf[e] >= out->n_kf) return VILBA_ERR_ARG;
    for (int32_t e = 0; e < out->n_imu; ++e)
        if (out->imu_kf_i[e] < 0 || out->imu_kf_i[e] >= out->n_kf || out->imu_kf_j[e] < 0 || out->imu_kf_j[e] >= out->n_kf)
            return VILBA_ERR_ARG;
    return VILBA_OK;
}

}  // extern "C"

// Micro-benchmarks that informed the Cholesky design: FP64 dependent-issue latencies on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench_fp64.cu && /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}
__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_dfma_ilp4(double* out, long long* cyc, double a, double b) {
    double x = out[threadIdx.x], y = x + 1, z = x + 2, w = x + 3;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        x = fma(x, a, b); y = fma(y, a, b); z = fma(z, a, b); w = fma(w, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + y + z + w;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_rcp(double* out, long long* cyc) {
    double x = out[threadIdx.x] + 1.5;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = fast_rcp(x) + 1.0; }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_div(double* out, long long* cyc) {
    double x = out[threadIdx.x] + 1.5;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = 1.0 / x + 1.0; }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_sqrt(double* out, long long* cyc) {
    double x = out[threadIdx.x] + 1.5;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = sqrt(x) + 1.0; }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_shfl(double* out, long long* cyc) {
    double x = out[threadIdx.x] + 1.5;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = __shfl_sync(0xffffffffu, x, (i * 7) & 31) + 1.0; }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_smem_rt(double* out, long long* cyc) {
    __shared__ double s[64];
    double x = out[threadIdx.x] + 1.5;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        s[threadIdx.x] = x;
        __syncwarp();
        x = s[(threadIdx.x + 1) & 31] + 1.0;
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
constexpr int NB = 16;
// LDL^T of a 16x16 block held one row per lane (lane r: a[c] = A(r,c) for c <= r; rows/cols beyond
// the real size are identity-padded by the caller).  On return lane r holds the unit-lower l(r,c) in
// a[c], c < r, and d_r in a[r].  `sm` is 16 + 64 doubles of warp-private shared memory.
__device__ __forceinline__ bool warp_ldlt16_mb4(double (&a)[NB], int lane, double* sm) {
    double* Bc = sm;       // 4 x 4  pivot block
    double* Lb = sm + 16;  // 16 x 4 scaled rows l(r, j..j+3)
    bool ok = true;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int j = 4 * s;
        if (lane >= j && lane < j + 4) {
            double* p = Bc + 4 * (lane - j);
            p[0] = a[j], p[1] = a[j + 1], p[2] = a[j + 2], p[3] = a[j + 3];
        }
        __syncwarp();
        const double b00 = Bc[0], b10 = Bc[4], b20 = Bc[8], b30 = Bc[12];
        double b11 = Bc[5], b21 = Bc[9], b31 = Bc[13], b22 = Bc[10], b32 = Bc[14], b33 = Bc[15];
        // 4x4 LDL^T, redundantly in every lane
        const double d0 = b00, r0 = fast_rcp(d0);
        const double l10 = b10 * r0, l20 = b20 * r0, l30 = b30 * r0;
        b11 -= l10 * b10, b21 -= l20 * b10, b31 -= l30 * b10;
        b22 -= l20 * b20, b32 -= l30 * b20, b33 -= l30 * b30;
        const double d1 = b11, r1 = fast_rcp(d1);
        const double l21 = b21 * r1, l31 = b31 * r1;
        b22 -= l21 * b21, b32 -= l31 * b21, b33 -= l31 * b31;
        const double d2 = b22, r2 = fast_rcp(d2);
        const double l32 = b32 * r2;
        b33 -= l32 * b32;
        const double d3 = b33, r3 = fast_rcp(d3);
        if (!(d0 > 0.0) || !(d1 > 0.0) || !(d2 > 0.0) || !(d3 > 0.0)) ok = false;
        // rows below the pivot block: u = A(r, j..j+3) Lb^-T (unscaled), l = u D^-1
        double u0 = a[j], u1 = a[j + 1], u2 = a[j + 2], u3 = a[j + 3];
        u1 -= u0 * l10;
        u2 -= u0 * l20 + u1 * l21;
        u3 -= u0 * l30 + u1 * l31 + u2 * l32;
        const double q0 = u0 * r0, q1 = u1 * r1, q2 = u2 * r2, q3 = u3 * r3;
        if (lane >= j + 4) {
            a[j] = q0, a[j + 1] = q1, a[j + 2] = q2, a[j + 3] = q3;
            if (lane < NB) {
                double* p = Lb + 4 * lane;
                p[0] = q0, p[1] = q1, p[2] = q2, p[3] = q3;
            }
        } else if (lane >= j) {  // rows of the pivot block itself
            const int i = lane - j;
            a[j] = (i == 0) ? d0 : (i == 1) ? l10 : (i == 2) ? l20 : l30;
            a[j + 1] = (i == 1) ? d1 : (i == 2) ? l21 : (i == 3) ? l31 : 0.0;
            a[j + 2] = (i == 2) ? d2 : (i == 3) ? l32 : 0.0;
            a[j + 3] = (i == 3) ? d3 : 0.0;
        }
        if (s < 3) {
            __syncwarp();
            // trailing part of the block: A(r,c) -= sum_k u(r,k) l(c,k), c = j+4 .. r
#pragma unroll
            for (int c = j + 4; c < NB; ++c) {
                const double* p = Lb + 4 * c;
                if (lane >= c) a[c] -= u0 * p[0] + u1 * p[1] + u2 * p[2] + u3 * p[3];
            }
        }
        __syncwarp();
    }
    return ok;
}


__global__ void k_ldlt(const double* blk, double* out, long long* cyc) {
    __shared__ double sm[80];
    const int lane = threadIdx.x;
    double acc = 0.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        double a[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) a[c] = (lane < 16 && c <= lane) ? blk[lane * 16 + c] + acc * 1e-30 : ((c == lane) ? 1.0 : 0.0);
        bool ok = warp_ldlt16_mb4(a, lane, sm);
#pragma unroll
        for (int c = 0; c < NB; ++c) acc += a[c];
        if (!ok) acc += 1.0;
    }
    long long t1 = clock64();
    out[lane] = acc;
    if (lane == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma_tput(double* out, long long* cyc, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = out[threadIdx.x] + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 128; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
// FP64 tensor path: mma.sync m8n8k4 (DMMA), 8 independent accumulator fragments per warp
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma_tput(double* out, long long* cyc, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = out[threadIdx.x] + i, c[i][1] = 0.5 * i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 128; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_dmma_lat(double* out, long long* cyc, double a, double b) {
    double c0 = out[threadIdx.x], c1 = 1.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) { dmma884(c0, c1, a, b); dmma884(c0, c1, a, b); dmma884(c0, c1, a, b); dmma884(c0, c1, a, b); }
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
// shared memory -> register bandwidth: every lane loads the SAME address (broadcast) or its own
template <int WIDTH, bool BCAST>
__global__ void k_lds_bw(double* out, long long* cyc) {
    __shared__ __align__(16) double s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int base = ((it + 4 * k) & 31) * 64 + (BCAST ? 0 : lane * (WIDTH / 8));
            if (WIDTH == 16) { const double2 v = *reinterpret_cast<const double2*>(s + base); acc += v.x + v.y; }
            else acc += s[base];
        }
    }
    __syncthreads();
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
int main() {
    {
        double* o; long long* cy; cudaMalloc(&o, 1024 * 8); cudaMemset(o, 0, 8192); cudaMallocManaged(&cy, 64);
        for (int nt : {32, 128, 256, 512, 1024}) {
            k_dmma_tput<<<1, nt>>>(o, cy, 1.0000001, 1e-9); cudaDeviceSynchronize();
            double mmas = (double)(nt / 32) * 128 * 32;  // warp-level m8n8k4 (256 FMAs each)
            printf("DMMA m8n8k4 throughput, %4d threads/SM: %.2f cycles per mma per SM, %.1f FMA/clk/SM\n", nt, (double)cy[0] / mmas, mmas * 256 / (double)cy[0]);
        }
        k_dmma_lat<<<1, 32>>>(o, cy, 1.0000001, 1e-9); cudaDeviceSynchronize();
        printf("dependent DMMA m8n8k4: %.1f cycles\n", (double)cy[0] / 1024);
        for (int nt : {32, 256, 1024}) {
            k_lds_bw<8, true><<<1, nt>>>(o, cy); cudaDeviceSynchronize();
            printf("LDS.64 broadcast,  %4d threads: %.2f cycles per warp-load per SM\n", nt, (double)cy[0] / ((nt / 32) * 1024.0));
            k_lds_bw<16, true><<<1, nt>>>(o, cy); cudaDeviceSynchronize();
            printf("LDS.128 broadcast, %4d threads: %.2f cycles per warp-load per SM\n", nt, (double)cy[0] / ((nt / 32) * 1024.0));
            k_lds_bw<8, false><<<1, nt>>>(o, cy); cudaDeviceSynchronize();
            printf("LDS.64 per lane,   %4d threads: %.2f cycles per warp-load per SM\n", nt, (double)cy[0] / ((nt / 32) * 1024.0));
            k_lds_bw<16, false><<<1, nt>>>(o, cy); cudaDeviceSynchronize();
            printf("LDS.128 per lane,  %4d threads: %.2f cycles per warp-load per SM\n", nt, (double)cy[0] / ((nt / 32) * 1024.0));
        }
    }
    {
        double* o; long long* cy; cudaMalloc(&o, 1024 * 8); cudaMemset(o, 0, 8192); cudaMallocManaged(&cy, 64);
        for (int nt : {32, 128, 256, 512, 1024}) {
            k_dfma_tput<<<1, nt>>>(o, cy, 1.0000001, 1e-9); cudaDeviceSynchronize();
            double fmas = (double)nt * 128 * 32;  // thread-level FMAs
            printf("DFMA throughput, %4d threads/SM: %.1f FMA/clk/SM\n", nt, fmas / (double)cy[0]);
        }
    }
    {
        double h[256];
        for (int r = 0; r < 16; ++r) for (int c = 0; c < 16; ++c) h[r * 16 + c] = (r == c) ? 20.0 + r : 1.0 / (1 + r + c);
        double* blk; cudaMalloc(&blk, sizeof(h)); cudaMemcpy(blk, h, sizeof(h), cudaMemcpyHostToDevice);
        double* o; long long* cy; cudaMalloc(&o, 256); cudaMallocManaged(&cy, 64);
        for (int i = 0; i < 2; ++i) { k_ldlt<<<1, 32>>>(blk, o, cy); cudaDeviceSynchronize(); printf("warp_ldlt16_mb4 %.0f cycles/block\n", (double)cy[0] / 64); }
    }

    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMemset(out, 0, 1024 * 8); cudaMallocManaged(&cyc, 64);
    auto rep = [&](const char* name, double per) { cudaDeviceSynchronize(); printf("%-28s %8.1f cycles/op\n", name, (double)cyc[0] / per); };
    for (int it = 0; it < 2; ++it) {
        k_dfma<<<1, 32>>>(out, cyc, 1.0000001, 1e-9); rep("dependent DFMA", 1024);
        k_dfma_ilp4<<<1, 32>>>(out, cyc, 1.0000001, 1e-9); rep("DFMA ilp4 (per fma)", 1024);
        k_dfma<<<1, 512>>>(out, cyc, 1.0000001, 1e-9); rep("dependent DFMA 16 warps", 1024);
        k_rcp<<<1, 32>>>(out, cyc); rep("fast_rcp + add", 256);
        k_div<<<1, 32>>>(out, cyc); rep("1.0/x + add", 256);
        k_sqrt<<<1, 32>>>(out, cyc); rep("sqrt + add", 256);
        k_shfl<<<1, 32>>>(out, cyc); rep("shfl64 + add", 256);
        k_smem_rt<<<1, 32>>>(out, cyc); rep("sts+syncwarp+lds+add", 256);
    }
    return 0;
}

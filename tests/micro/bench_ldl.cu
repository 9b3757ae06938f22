// Micro-benchmark (not a test): cycles of the phases of one log-determinant task
// (stage -> form -> LDL^T -> log) for 1 warp alone and for 12 warps on one SM.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o bench_ldl bench_ldl.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../speaker-diarization_b200/csrc/common.cuh"
#include "../../speaker-diarization_b200/csrc/score.cuh"
using namespace spk;

__global__ void __launch_bounds__(384, 1) k(const double* rec, int ntask, double* out, long long* cyc) {
    extern __shared__ __align__(16) unsigned char sm[];
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch& w = ws[warp];
    long long t_stage = 0, t_form = 0, t_ldl = 0, t_tot = 0;
    for (int it = 0; it < ntask; ++it) {
        const RecSrc X{rec + (size_t)((it * 12 + warp) % 64) * REC};
        const long long c0 = clock64();
        const double* rx = stage_record(X, w.rec[0], lane);
        __syncwarp();
        const long long c1 = clock64();
        double a[Grid<D39>::NSLOT];
        const SmemSrc sx{rx};
        const double n = form_matrix<D39>(a, FORM_X, sx, sx, 1.0, 1.0, w, lane);
        const long long c2 = clock64();
        const double lm = ldl_logdet<D39>(a, w, lane);
        const long long c3 = clock64();
        const double v = finish_logdet(lm, n, D39);
        if (lane == 0) out[warp * ntask + it] = v;
        const long long c4 = clock64();
        t_stage += c1 - c0; t_form += c2 - c1; t_ldl += c3 - c2; t_tot += c4 - c0;
    }
    if (lane == 0) { cyc[warp * 4 + 0] = t_stage; cyc[warp * 4 + 1] = t_form; cyc[warp * 4 + 2] = t_ldl; cyc[warp * 4 + 3] = t_tot; }
}

struct KScr { LdlScratch w; union { double rec[REC]; double Lsm[(D39 * (D39 - 1)) / 2 + 3]; }; double pinv[VS]; double dS[VS], dP[VS]; };
__global__ void __launch_bounds__(384, 1) k2(const double* rec, int ntask, double* out, long long* cyc) {
    extern __shared__ __align__(16) unsigned char sm[];
    KScr* ks = reinterpret_cast<KScr*>(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    KScr& k = ks[warp];
    long long t_tot = 0, t_inv = 0;
    for (int it = 0; it < ntask; ++it) {
        const RecSrc X{rec + (size_t)((it * 12 + warp) % 64) * REC};
        const long long c0 = clock64();
        kl2_side_one(X, k, k.dS, k.dP, lane);
        const long long c1 = clock64();
        double ga = 0, gb = 0;
        const long long c2 = clock64();
        if (lane == 0) out[warp * ntask + it] = k.dP[3] + ga + gb;
        t_tot += c1 - c0; t_inv += c2 - c1;
    }
    if (lane == 0) { cyc[warp * 4 + 0] = t_tot; cyc[warp * 4 + 1] = t_inv; }
}

int main() {
    const int NREC = 64;
    std::vector<double> h((size_t)NREC * REC, 0.0);
    // records of SPD statistics: n = 500 frames of iid N(0,1)-ish data: Q = n*I + noise, s = small
    srand(1);
    for (int r = 0; r < NREC; ++r) {
        double* p = h.data() + (size_t)r * REC;
        for (int i = 0; i < D39; ++i)
            for (int j = 0; j <= i; ++j)
                p[L39::pos(i, j)] = (i == j ? 500.0 + (rand() % 100) : (rand() % 200 - 100) * 0.05);
        for (int j = 0; j < D39; ++j) p[L39::VEC + j] = (rand() % 100 - 50) * 0.1;
        p[L39::CNT] = 500.0;
    }
    double *d, *out; long long* cyc;
    cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 12 * 64 * 8); cudaMalloc(&cyc, 12 * 4 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(12 * sizeof(WarpScratch)));
    for (int nw : {1, 4, 12}) {
        const int ntask = 32;
        for (int rep = 0; rep < 2; ++rep) k<<<1, nw * 32, 12 * sizeof(WarpScratch)>>>(d, ntask, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long hc[48]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        double o; cudaMemcpy(&o, out, 8, cudaMemcpyDeviceToHost);
        printf("%2d warps: cycles/task stage %6.0f form %6.0f ldl %6.0f total %6.0f   (logdet %.6f)\n", nw,
               hc[0] / (double)ntask, hc[1] / (double)ntask, hc[2] / (double)ntask, hc[3] / (double)ntask, o);
    }
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(12 * sizeof(KScr)));
    for (int nw : {1, 4, 12}) {
        const int ntask = 32;
        for (int rep = 0; rep < 2; ++rep) k2<<<1, nw * 32, 12 * sizeof(KScr)>>>(d, ntask, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long hc[48]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        printf("%2d warps: KL2 side cycles/task total %6.0f   (unused %6.0f)\n", nw, hc[0] / (double)ntask, hc[1] / (double)ntask);
    }
    return 0;
}

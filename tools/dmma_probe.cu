// dmma_probe.cu — design-space probe for the FP64 DMMA inner loop of chain_kernel (not product code).
// C[M,N] = A[M,K] * B[N,K]^T, both operands contiguous along K (the "MK / NK" layout of the product kernel).
// Variants differ in CTA tile, warp tile, pipeline depth and copy width; prints TFLOP/s of each so that the
// product kernel is rebuilt around the configuration that keeps the tensor pipe busiest.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ void cp8(double* s, const double* g) {
    unsigned a = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(a), "l"(g));
}
__device__ __forceinline__ void cp16(double* s, const double* g) {
    unsigned a = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(a), "l"(g));
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void waitg() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// BM x BN CTA tile, WR x WC warps, each warp (BM/WR) x (BN/WC); BK deep chunks; NST stages; W16: 16-byte copies
template <int BM, int BN, int WR, int WC, int BK, int NST, bool W16, int MINB>
__global__ void __launch_bounds__(WR* WC * 32, MINB) gemm(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int M, int N, int K) {
    constexpr int NT = WR * WC * 32;
    constexpr int LD = BK + 4;  // padded row stride (doubles), 20 or 36: == 4 mod 16 -> conflict-free fragment reads
    constexpr int TM = BM / WR, TN = BN / WC, FM = TM / 8, FN = TN / 8;
    extern __shared__ __align__(16) double sm[];
    double* As = sm;
    double* Bs = sm + NST * BM * LD;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wr = warp / WC, wc = warp % WC;
    const int bm = blockIdx.y * BM, bn = blockIdx.x * BN;
    const double* Ag = A + (size_t)bm * K;
    const double* Bg = B + (size_t)bn * K;
    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    auto issue = [&](int slot, int k0) {
        if (W16) {
            constexpr int PER_ROW = BK / 2;
            for (int e = tid; e < BM * PER_ROW; e += NT) { int r = e / PER_ROW, c = (e % PER_ROW) * 2; cp16(As + slot * BM * LD + r * LD + c, Ag + (size_t)r * K + k0 + c); }
            for (int e = tid; e < BN * PER_ROW; e += NT) { int r = e / PER_ROW, c = (e % PER_ROW) * 2; cp16(Bs + slot * BN * LD + r * LD + c, Bg + (size_t)r * K + k0 + c); }
        } else {
#pragma unroll
            for (int e = tid; e < BM * BK; e += NT) { int r = e / BK, c = e % BK; cp8(As + slot * BM * LD + r * LD + c, Ag + (size_t)r * K + k0 + c); }
#pragma unroll
            for (int e = tid; e < BN * BK; e += NT) { int r = e / BK, c = e % BK; cp8(Bs + slot * BN * LD + r * LD + c, Bg + (size_t)r * K + k0 + c); }
        }
    };
    const int nch = K / BK;
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) { if (p < nch) issue(p, p * BK); commit(); }
    int cur = 0, fill = NST - 1;
    for (int c = 0; c < nch; ++c) {
        waitg<NST - 2>();
        __syncthreads();
        if (c + NST - 1 < nch) issue(fill, (c + NST - 1) * BK);
        commit();
        const double* as = As + cur * BM * LD + (wr * TM + g) * LD + t;
        const double* bs = Bs + cur * BN * LD + (wc * TN + g) * LD + t;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[FM], b[FN];
#pragma unroll
            for (int i = 0; i < FM; ++i) a[i] = as[i * 8 * LD + kk * 4];
#pragma unroll
            for (int j = 0; j < FN; ++j) b[j] = bs[j * 8 * LD + kk * 4];
#pragma unroll
            for (int j = 0; j < FN; ++j)
#pragma unroll
                for (int i = 0; i < FM; ++i) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        cur = cur + 1 == NST ? 0 : cur + 1;
        fill = fill + 1 == NST ? 0 : fill + 1;
    }
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) {
            const int r = bm + wr * TM + i * 8 + g, cc = bn + wc * TN + j * 8 + 2 * t;
            C[(size_t)r * N + cc] = acc[i][j][0];
            C[(size_t)r * N + cc + 1] = acc[i][j][1];
        }
}

template <int BM, int BN, int WR, int WC, int BK, int NST, bool W16, int MINB>
void run(const char* name, const double* A, const double* B, double* C, int M, int N, int K) {
    constexpr int LD = BK + 4;
    const int smem = NST * (BM + BN) * LD * 8;
    auto kern = gemm<BM, BN, WR, WC, BK, NST, W16, MINB>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { printf("%-44s smem %d too large\n", name, smem); cudaGetLastError(); return; }
    dim3 grid(N / BN, M / BM);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WR * WC * 32, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    for (int i = 0; i < 2; ++i) kern<<<grid, WR * WC * 32, smem>>>(A, B, C, M, N, K);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; ++i) kern<<<grid, WR * WC * 32, smem>>>(A, B, C, M, N, K);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    printf("%-44s regs %3d occ %d smem %6d  %.3f ms  %.2f TFLOP/s %s\n", name, fa.numRegs, occ, smem, ms / reps, 2.0 * M * N * K / (ms / reps * 1e-3) / 1e12,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const int M = 4096, N = 4096 + 128 * 5, K = 2048;  // N chosen so that 64/128-wide tiles both divide it
    double *A, *B, *C;
    cudaMalloc(&A, (size_t)M * K * 8); cudaMalloc(&B, (size_t)N * K * 8); cudaMalloc(&C, (size_t)M * N * 8);
    std::vector<double> h((size_t)N * K);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
    cudaMemcpy(A, h.data(), (size_t)M * K * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(B, h.data(), (size_t)N * K * 8, cudaMemcpyHostToDevice);
#define RUN(...) run<__VA_ARGS__>(#__VA_ARGS__, A, B, C, M, N, K)
    //   BM  BN  WR WC BK NST W16 MINB
    RUN(64, 64, 2, 2, 16, 2, false, 3);
    RUN(64, 64, 2, 2, 16, 3, false, 3);
    RUN(64, 64, 2, 2, 16, 4, false, 2);
    RUN(64, 64, 2, 2, 16, 3, true, 3);
    RUN(64, 64, 2, 2, 32, 3, false, 2);
    RUN(64, 64, 2, 2, 32, 3, true, 2);
    RUN(128, 64, 4, 2, 16, 3, false, 1);
    RUN(128, 64, 4, 2, 16, 3, true, 1);
    RUN(128, 64, 4, 2, 16, 4, true, 1);
    RUN(64, 128, 2, 2, 16, 3, false, 2);
    RUN(64, 128, 2, 2, 16, 3, true, 2);
    RUN(64, 128, 2, 2, 16, 4, true, 2);
    RUN(128, 128, 4, 2, 16, 3, false, 1);
    RUN(128, 128, 4, 2, 16, 3, true, 1);
    RUN(128, 128, 4, 2, 16, 4, true, 1);
    RUN(128, 128, 4, 2, 32, 3, true, 1);
    RUN(128, 128, 2, 4, 16, 4, true, 1);
    RUN(128, 128, 4, 4, 16, 4, true, 1);
    RUN(128, 128, 4, 4, 16, 3, false, 1);
    RUN(128, 64, 2, 2, 16, 3, true, 1);
    RUN(128, 64, 2, 2, 16, 3, false, 2);
    RUN(64, 64, 1, 4, 16, 3, false, 3);
    RUN(64, 64, 4, 1, 16, 3, false, 3);
    RUN(32, 64, 1, 2, 16, 3, false, 4);
    return 0;
}

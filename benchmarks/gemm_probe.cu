// Stand-alone timing probe for gemm_i8.cu (experiments only; not part of the library).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DPA_GEMM_PROBE [-DPA_GEMM_STAGES=n] \
//        -I include -I <pkg>/csrc benchmarks/gemm_probe.cu <pkg>/csrc/page_table.cu -o gpurun_out/gemm_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
unsigned long long* pa_gemm_probe_buf = nullptr;
#include "gemm_i8.cu"

int main(int argc, char** argv) {
    int M = argc > 1 ? atoi(argv[1]) : 256, N = argc > 2 ? atoi(argv[2]) : 16384, K = argc > 3 ? atoi(argv[3]) : 4096;
    int8_t *A, *B[4], *C;
    cudaMalloc(&A, (size_t)M * K); cudaMalloc(&C, (size_t)M * N);
    for (int i = 0; i < 4; ++i) { cudaMalloc(&B[i], (size_t)K * N); cudaMemset(B[i], 1, (size_t)K * N); }
    cudaMemset(A, 1, (size_t)M * K);
    cudaMalloc(&pa_gemm_probe_buf, 64); cudaMemset(pa_gemm_probe_buf, 0, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9, sum = 0; int iters = 20;
    for (int it = 0; it < iters + 5; ++it) {
        cudaEventRecord(e0, 0);
        int st = pa_gemm_i8(A, B[it & 3], C, nullptr, 1, M, N, K, 1.f, 1.f, 1.f, nullptr, 0, nullptr);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        if (st) { printf("status %d\n", st); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 5) { sum += ms; if (ms < best) best = ms; }
    }
    // back-to-back launches (4 weight copies round-robin, no sync in between): device time per launch
    cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    for (int it = 0; it < 40; ++it) pa_gemm_i8(A, B[it & 3], C, nullptr, 1, M, N, K, 1.f, 1.f, 1.f, nullptr, 0, nullptr);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms40; cudaEventElapsedTime(&ms40, e0, e1);
    printf("back-to-back: %.2f us per launch\n", ms40 / 40 * 1e3);
    unsigned long long h[8]; cudaMemcpy(h, pa_gemm_probe_buf, 64, cudaMemcpyDeviceToHost);
    printf("M=%d N=%d K=%d stages=%d: avg %.2f us  min %.2f us | cta5: producer wait %llu / %llu cyc, mma wait %llu / %llu cyc | setup %llu, mainloop(epi view) %llu, epilogue %llu cyc\n",
           M, N, K, pa::gemm::STAGES, sum / iters * 1e3, best * 1e3, h[0], h[2], h[1], h[3], h[4], h[5], h[6]);
    return 0;
}
